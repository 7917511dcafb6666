"""Generate tests/golden/*.json from the LIVE reference (run in the build container only).

    python oracle/make_golden.py [--out tests/golden]

Imports the unmodified reference from /root/reference with the two torch-version shims documented in
SURVEY.md 8(c) (pass-through data_parallel on CPU; float labels in get_labels_for), restates the
~35-line iteration body of gan/trainer.py:199-265 around the IMPORTED CondGan / models / losses
(train() itself needs DALI + a CUDA stream), and records:
  * init parity data: per-tensor checksums of the seed-100 xavier-initialised G / D / caption encoder;
  * one full TGANv2-conditional iteration at B=8, 64x64x16: lossD, lossG, per-parameter gradient
    norms + first elements, level shapes and checksums, the host-RNG draws;
  * the same for TGANv2-unconditional (documented gen_step adapter);
  * index fixtures: Subsample, nearest pyramid, gen_perm, token batches.
It then runs oracle/txt2vid_oracle.py on the same weights/inputs and prints the deviations, so the
oracle is pinned before anything is compared against it.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def import_reference():
    sys.path.insert(0, REF)
    import torch.nn.parallel
    torch.nn.parallel.data_parallel = lambda m, x, *a, **k: m(x)       # shim 1
    import txt2vid.gan.losses as L
    L.get_labels_for = lambda x, label: torch.full(x.size(), float(label), device=x.device)  # shim 2
    return L


def checksum(t):
    t = t.detach().double().reshape(-1)
    idx = torch.arange(t.numel(), dtype=torch.float64)
    return {"sum": float(t.sum()), "abs": float(t.abs().sum()), "wsum": float((t * ((idx % 97) + 1)).sum()),
            "n": int(t.numel()), "first": [float(v) for v in t[:4]]}


def seed_all(seed):
    import random
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def synth_batch(B, V, T=16, S=64, seed=1234):
    """Synthetic inputs of SURVEY 8(d): U(-1,1) video (B,3,T,S,S), MSRVDC-shaped captions."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, T, S, S, generator=g) * 2 - 1
    lengths = sorted([int(v) for v in torch.randint(4, 21, (B,), generator=g)], reverse=True)
    tokens = torch.zeros(B, lengths[0], dtype=torch.long)
    for b, L in enumerate(lengths):
        tokens[b, 0] = 1
        tokens[b, 1:L - 1] = torch.randint(4, V, (L - 2,), generator=g)
        tokens[b, L - 1] = 2
    return x, tokens, lengths


def build_reference_models(conditional, V=1000, seed=100, size=64, frames=16):
    """Construction + init order of train/gan.py:28-70."""
    import contextlib
    import io
    from txt2vid.util.torch.init import init
    seed_all(seed)
    txt = None
    with contextlib.redirect_stdout(io.StringIO()):
        if conditional:
            from txt2vid.models.txt.basic import Seq2Seq
            from txt2vid.models.tganv2_cond.gen import MultiScaleGen
            from txt2vid.models.tganv2_cond.discrim import MultiScaleDiscrim
            txt = Seq2Seq(vocab_size=V)
            init(txt, "xavier")
            gen = MultiScaleGen(width=size, height=size, cond_dim=256, num_frames=frames)
            dis = MultiScaleDiscrim(cond_dim=256)
        else:
            from txt2vid.models.tganv2.gen import MultiScaleGen
            from txt2vid.models.tganv2.discrim import MultiScaleDiscrim
            gen = MultiScaleGen(width=64, height=64, cond_dim=0)
            dis = MultiScaleDiscrim(cond_dim=0)
    init(gen, "xavier")
    init(dis, "xavier")
    return txt, gen, dis


def reference_iteration(L, txt, gen, dis, x, tokens, lengths, conditional, gp_lambda=0.5, lr=2e-4,
                        betas=(0.5, 0.999), frame_sizes=(8, 16, 32, 64)):
    """Iteration body of gan/trainer.py:199-265 around the imported reference objects."""
    import torch.nn.functional as F
    from txt2vid.gan.cond_gan import CondGan
    from txt2vid.gan.losses import MixedGanLoss, RSGANLoss
    from txt2vid.models.layers import Subsample
    gan = CondGan(gen=gen, discrims=[dis], cond_encoder=txt, discrim_names=["video"])
    losses = MixedGanLoss(g_loss=RSGANLoss(), d_loss=RSGANLoss())
    optD = torch.optim.Adam([{"params": dis.parameters()}], lr=lr, betas=betas)
    optG = torch.optim.Adam([{"params": gen.parameters()}], lr=lr, betas=betas)
    B = x.size(0)
    cond = None
    if conditional:
        _, _, cond = txt.encode(tokens, lengths)
        cond = cond.detach()
    # multiscale_data, trainer.py:131-165 (subsample_input = True)
    sub = Subsample()
    xs, conds, bts = [], [], []
    xi, ci = x, cond
    for i in range(len(frame_sizes)):
        T = xi.size(2)
        xs.append(F.interpolate(xi, size=(T, frame_sizes[i], frame_sizes[i])) if i != len(frame_sizes) - 1 else xi)
        if ci is not None:
            conds.append(ci)
        xi, bt = sub(xi)
        bts.append(int(bt))
        if ci is not None:
            ci = ci[::2]
    conds = conds if conds else None
    z = torch.randn(B, gan.gen.latent_size)
    fake = gan(z, cond=conds[0]) if conds is not None else gan(z, cond=None)
    out = {"z": z, "bt_real": bts, "real_levels": xs, "fake": [f.detach().clone() for f in fake]}
    loss = gan.discrim_step(real=xs, fake=[f.detach() for f in fake], cond=conds, loss=losses.discrim_loss,
                            gp_lambda=gp_lambda)
    loss.backward()
    out["lossD"] = float(loss)
    out["gradD"] = {n: p.grad.detach().clone() for n, p in dis.named_parameters() if p.grad is not None}
    optD.step()
    _, _, real_pred = gan.all_discrim_forward(real=xs, cond=conds, fake=None, loss=None)
    if conditional:
        lossg = gan.gen_step(fake=fake, real_pred=real_pred, cond=conds, loss=losses.gen_loss)
    else:
        # documented HEAD adapter: cond_gan.py:102-106 passes tuples to the loss
        gen.zero_grad()
        fake_cc = dis(x=fake, cond=None, xbar=None)
        lossg = torch.stack([losses.gen_loss(fake=ff[0], real=rr) for ff, rr in zip(fake_cc, real_pred[0])]).mean()
    lossg.backward()
    out["lossG"] = float(lossg)
    out["gradG"] = {n: p.grad.detach().clone() for n, p in gen.named_parameters() if p.grad is not None}
    optG.step()
    out["gen_after"] = {k: v.detach().clone() for k, v in gen.state_dict().items()}
    out["dis_after"] = {k: v.detach().clone() for k, v in dis.state_dict().items()}
    return out


def run_config(L, conditional, out_dir, B=8, V=1000, size=64, frames=16, frame_sizes=(8, 16, 32, 64)):
    """size / frames / frame_sizes: BASELINE configs[4] is the conditional model at 128 x 128 x 32 with the pyramid
    16 / 32 / 64 / 128 (the ConvLSTM plane becomes 2 x 2, tganv2_cond/gen.py:32-33)."""
    import oracle.txt2vid_oracle as O
    name = "tganv2_cond" if conditional else "tganv2_uncond"
    if size != 64:
        name += "_%dx%dx%d" % (size, size, frames)
    t0 = time.time()
    txt, gen, dis = build_reference_models(conditional, V, size=size, frames=frames)
    init_sd = {"gen": {k: v.detach().clone() for k, v in gen.state_dict().items()},
               "dis": {k: v.detach().clone() for k, v in dis.state_dict().items()},
               "txt": None if txt is None else {k: v.detach().clone() for k, v in txt.state_dict().items()}}
    x, tokens, lengths = synth_batch(B, V, T=frames, S=size)
    rng_t, rng_n = torch.get_rng_state(), np.random.get_state()
    ref = reference_iteration(L, txt, gen, dis, x, tokens, lengths, conditional, frame_sizes=frame_sizes)
    print("[%s] reference iteration: lossD %.6f lossG %.6f (%.1fs)" % (name, ref["lossD"], ref["lossG"], time.time() - t0))

    # ---- the oracle on the same weights, inputs and RNG stream
    torch.set_rng_state(rng_t)
    np.random.set_state(rng_n)
    sd_g, sd_d = O.as_leaves(init_sd["gen"]), O.as_leaves(init_sd["dis"])
    sd_t = None if init_sd["txt"] is None else O.as_leaves(init_sd["txt"])
    bt_real = O.draw_real(4, True)
    z = torch.randn(B, 256)
    draws = O.draw_rest([B, B // 2, B // 4, B // 8], conditional=conditional, gp=True)
    draws["bt_real"] = bt_real
    assert bt_real == ref["bt_real"], (bt_real, ref["bt_real"])
    assert torch.equal(z, ref["z"])
    opt_g = O.Adam(O.param_names(sd_g), 2e-4, (0.5, 0.999))
    opt_d = O.Adam(O.param_names(sd_d), 2e-4, (0.5, 0.999))
    t1 = time.time()
    orc = O.train_iteration(sd_g, sd_d, sd_t, x, tokens, lengths, z, draws, opt_g=opt_g, opt_d=opt_d,
                            frame_sizes=frame_sizes, num_frames=frames)
    print("[%s] oracle iteration:    lossD %.6f lossG %.6f (%.1fs)" % (name, orc["lossD"], orc["lossG"], time.time() - t1))

    def l2rel(a, b):
        """Per-tensor L2 deviation; tensors whose reference norm is numerically zero (biases in front of a
        BatchNorm, last-layer biases under a relativistic loss) are checked to be ~zero instead."""
        nb = float(b.double().norm())
        if nb < 1e-5:
            assert float(a.double().norm()) < 1e-5, (float(a.double().norm()), nb)
            return 0.0
        return float((a.double() - b.double()).norm()) / nb
    dev = {"lossD": abs(orc["lossD"] - ref["lossD"]), "lossG": abs(orc["lossG"] - ref["lossG"]),
           "fake": max(l2rel(a, b) for a, b in zip(orc["fake"], ref["fake"])),
           "gradD": max(l2rel(orc["gradD"][n], g) for n, g in ref["gradD"].items()),
           "gradG": max(l2rel(orc["gradG"][n], g) for n, g in ref["gradG"].items())}
    print("[%s] oracle vs reference deviations:" % name, json.dumps(dev))
    assert set(orc["gradD"]) == set(ref["gradD"]) and set(orc["gradG"]) == set(ref["gradG"])

    fixture = {
        "config": {"model": name, "B": B, "V": V, "seed": 100, "data_seed": 1234, "frame_sizes": list(frame_sizes),
                   "size": size, "frames": frames, "gp_lambda": 0.5, "loss": "RSGAN", "lr": 2e-4, "betas": [0.5, 0.999]},
        "generated_by": "oracle/make_golden.py from the live reference at /root/reference, torch %s" % torch.__version__,
        "init": {part: (None if sd is None else {k: checksum(v) for k, v in sd.items() if v.dtype.is_floating_point})
                 for part, sd in init_sd.items()},
        "draws": {"bt_real": ref["bt_real"], "bt_fake": draws["bt_fake"],
                  "perm": None if draws["perm"] is None else [int(v) for v in draws["perm"]],
                  "alphas": [[float(v) for v in a.reshape(-1)] for a in draws["alphas"]]},
        "tokens": tokens.tolist(), "lengths": lengths,
        "x": checksum(x), "z": checksum(ref["z"]),
        "lossD": ref["lossD"], "lossG": ref["lossG"],
        "fake": [dict(checksum(f), shape=list(f.shape)) for f in ref["fake"]],
        "real_levels": [dict(checksum(f), shape=list(f.shape)) for f in ref["real_levels"]],
        "gradD": {n: dict(checksum(g), norm=float(g.double().norm())) for n, g in ref["gradD"].items()},
        "gradG": {n: dict(checksum(g), norm=float(g.double().norm())) for n, g in ref["gradG"].items()},
        "oracle_vs_reference": dev,
    }
    with open(os.path.join(out_dir, name + "_B%d.json" % B), "w") as f:
        json.dump(fixture, f)
    return dev


def index_fixtures(out_dir):
    """Bit-exact index work: Subsample (layers.py:106-111), nearest pyramid (trainer.py:149),
    gen_perm (util/misc.py:3-8)."""
    import torch.nn.functional as F
    from txt2vid.models.layers import Subsample
    from txt2vid.util.misc import gen_perm
    fx = {"subsample": [], "nearest": [], "gen_perm": []}
    for shape in ((8, 3, 16, 4, 4), (5, 2, 7, 3, 3), (1, 1, 1, 2, 2), (4, 2, 2, 2, 2)):
        x = torch.arange(int(np.prod(shape)), dtype=torch.float32).view(shape)
        for bt in (0, 1):
            y, _ = Subsample()(x, bt=bt)
            fx["subsample"].append({"shape": list(shape), "bt": bt, "out_shape": list(y.shape),
                                    "values": y.reshape(-1)[:64].tolist(), "sum": float(y.double().sum())})
    for (T, S), fs in (((16, 64), 8), ((8, 64), 16), ((4, 64), 32), ((3, 20), 7), ((2, 10), 16)):
        x = torch.arange(2 * 3 * T * S * S, dtype=torch.float32).view(2, 3, T, S, S)
        y = F.interpolate(x, size=(T, fs, fs))
        fx["nearest"].append({"in": [2, 3, T, S, S], "fs": fs, "first_row": y[0, 0, 0, 0].tolist(),
                              "first_col": y[0, 0, 0, :, 0].tolist(), "sum": float(y.double().sum())})
    for seed, n in ((100, 8), (100, 2), (7, 40), (3, 3)):
        np.random.seed(seed)
        fx["gen_perm"].append({"seed": seed, "n": n, "perm": [int(v) for v in gen_perm(n)],
                               "perm2": [int(v) for v in gen_perm(n)]})
    seed_all(100)
    fx["randint_stream_seed100"] = [int(torch.randint(2, (1,))) for _ in range(16)]
    with open(os.path.join(out_dir, "index_fixtures.json"), "w") as f:
        json.dump(fx, f)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    Lmod = import_reference()
    if a.only in (None, "index"):
        index_fixtures(a.out)
    if a.only in (None, "cond"):
        run_config(Lmod, True, a.out)
    if a.only in (None, "uncond"):
        run_config(Lmod, False, a.out)
    if a.only in (None, "cond128"):
        run_config(Lmod, True, a.out, size=128, frames=32, frame_sizes=(16, 32, 64, 128))
