"""Generate tests/golden/tgan_B8.json and tcwyt_B4.json from the LIVE reference (build container only).

    python oracle/make_golden_families.py [--out tests/golden]

Imports the unmodified TGAN / TCWYT modules and loss classes from /root/reference (shim: float labels in
get_labels_for; TGAN Gen.forward's debug prints are swallowed), drives them directly
(these families do not run through CondGan at HEAD), records initial-weight
checksums, inputs, losses, the fake-clip checksum and per-parameter gradient norms, then runs
oracle/families_oracle.py on the same weights / inputs and prints the deviations, so the oracle is pinned
before anything is compared against it.
"""
import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


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


def grad_norms(module):
    return {k: float(p.grad.norm()) for k, p in module.named_parameters() if p.grad is not None}


def synth(B, T, S, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, T, S, S, generator=g) * 2 - 1


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def cmp_grads(name, ref_norms, oracle_grads):
    worst = 0.0
    for k, n in ref_norms.items():
        o = float(oracle_grads[k].norm())
        worst = max(worst, rel(o, n) if n > 1e-7 else abs(o))
    print("   %s gradient norms: worst relative deviation %.2e over %d tensors" % (name, worst, len(ref_norms)))
    return worst


def run_tgan(out_dir, B=8):
    sys.path.insert(0, REF)
    import txt2vid.gan.losses as L
    L.get_labels_for = lambda x, label: torch.full(x.size(), float(label), device=x.device)
    from txt2vid.models.tgan.gen import Gen
    from txt2vid.models.tgan.discrim import Discrim
    from txt2vid.util.torch.init import init
    import oracle.families_oracle as O
    seed_all(100)
    gen = Gen()
    dis = Discrim(cond_dim=0)
    init(gen, "xavier")
    init(dis, "xavier")
    init_ck = {"gen": {k: checksum(v) for k, v in gen.state_dict().items() if v.dtype.is_floating_point},
               "dis": {k: checksum(v) for k, v in dis.state_dict().items() if v.dtype.is_floating_point}}
    x = synth(B, 16, 64)
    z = torch.randn(B, gen.latent_size)
    loss = L.WassersteinGanLoss()
    with contextlib.redirect_stdout(io.StringIO()):
        fake = gen(z)
    lossD = loss.discrim_loss(fake=dis(x=fake.detach()), real=dis(x=x))
    lossD.backward()
    gD = grad_norms(dis)
    dis.zero_grad()
    lossG = loss.gen_loss(fake=dis(x=fake), real=None)
    lossG.backward()
    gG = grad_norms(gen)
    rec = {"B": B, "seed": 100, "input_seed": 1234, "z": z.tolist(),
           "init": init_ck,
           "lossD": float(lossD), "lossG": float(lossG), "fake": checksum(fake), "fake_shape": list(fake.shape),
           "gradD": gD, "gradG": gG}
    # pin the oracle
    seed_all(100)
    gen2, dis2 = Gen(), Discrim(cond_dim=0)
    init(gen2, "xavier")
    init(dis2, "xavier")
    sd_g, sd_d = O.leaves(gen2.state_dict()), O.leaves(dis2.state_dict())
    o = O.tgan_iteration(sd_g, sd_d, x, z)
    print("TGAN  reference lossD %.6f lossG %.6f | oracle deviation lossD %.2e lossG %.2e fake %.2e" % (
        rec["lossD"], rec["lossG"], rel(o["lossD"], rec["lossD"]), rel(o["lossG"], rec["lossG"]),
        float((o["fake"] - fake.detach()).abs().max())))
    rec["oracle_dev"] = {"lossD": rel(o["lossD"], rec["lossD"]), "lossG": rel(o["lossG"], rec["lossG"]),
                         "gradD": cmp_grads("D", gD, o["gD"]), "gradG": cmp_grads("G", gG, o["gG"])}
    with open(os.path.join(out_dir, "tgan_B%d.json" % B), "w") as f:
        json.dump(rec, f)


def run_tcwyt(out_dir, B=4):
    sys.path.insert(0, REF)
    import txt2vid.gan.losses as L
    L.get_labels_for = lambda x, label: torch.full(x.size(), float(label), device=x.device)
    from txt2vid.models.tcwyt.gen import Gen
    from txt2vid.models.tcwyt.video_discrim import VideoDiscrim
    from txt2vid.models.tcwyt.frame_discrim import FrameMap, FrameDiscrim
    from txt2vid.models.tcwyt.motion_discrim import MotionDiscrim
    from txt2vid.util.torch.init import init
    import oracle.families_oracle as O

    def build():
        seed_all(100)
        mods = [Gen(cond_dim=256), VideoDiscrim(cond_dim=256), FrameDiscrim(cond_dim=256), MotionDiscrim(cond_dim=256),
                FrameMap()]
        for m in mods:
            init(m, "xavier")
        return mods
    gen, dv, df, dm, fm = build()
    init_ck = {n: {k: checksum(v) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
               for n, m in (("gen", gen), ("video", dv), ("frame", df), ("motion", dm), ("map", fm))}
    x = synth(B, 16, 48)
    z = torch.randn(B, gen.latent_size)
    cond = torch.randn(B, 256)
    loss = L.RaLSGANLoss()

    def d_all(vid):
        m = fm(vid)
        return [dv(x=vid, cond=cond), df(x=vid, cond=cond, xbar=m), dm(x=vid, cond=cond, xbar=m)]
    fake = gen(z, cond=cond)
    real_o, fake_o = d_all(x), d_all(fake.detach())
    lossD = sum(loss.discrim_loss(fake=f, real=r) for f, r in zip(fake_o, real_o)) / 3
    lossD.backward()
    gD = {"video": grad_norms(dv), "frame": grad_norms(df), "motion": grad_norms(dm), "map": grad_norms(fm)}
    for m in (dv, df, dm, fm):
        m.zero_grad()
    fake_o2 = d_all(fake)
    lossG = sum(loss.gen_loss(fake=f, real=r.detach()) for f, r in zip(fake_o2, real_o)) / 3
    lossG.backward()
    gG = grad_norms(gen)
    rec = {"B": B, "seed": 100, "input_seed": 1234, "z": z.tolist(), "cond": cond.tolist(),
           "init": init_ck,
           "lossD": float(lossD), "lossG": float(lossG), "fake": checksum(fake), "fake_shape": list(fake.shape),
           "frame_out_shape": list(real_o[1].shape), "motion_out_shape": list(real_o[2].shape),
           "gradD": gD, "gradG": gG}
    gen2, dv2, df2, dm2, fm2 = build()
    sds = [O.leaves(m.state_dict()) for m in (gen2, dv2, df2, dm2, fm2)]
    o = O.tcwyt_iteration(sds[0], sds[1], sds[2], sds[3], sds[4], x, z, cond)
    print("TCWYT reference lossD %.6f lossG %.6f | oracle deviation lossD %.2e lossG %.2e fake %.2e" % (
        rec["lossD"], rec["lossG"], rel(o["lossD"], rec["lossD"]), rel(o["lossG"], rec["lossG"]),
        float((o["fake"] - fake.detach()).abs().max())))
    dev = {"lossD": rel(o["lossD"], rec["lossD"]), "lossG": rel(o["lossG"], rec["lossG"]),
           "gradG": cmp_grads("G", gG, o["gG"])}
    for n in ("video", "frame", "motion", "map"):
        dev["grad_" + n] = cmp_grads(n, gD[n], o["gD"][n])
    rec["oracle_dev"] = dev
    with open(os.path.join(out_dir, "tcwyt_B%d.json" % B), "w") as f:
        json.dump(rec, f)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    run_tgan(a.out)
    run_tcwyt(a.out)
