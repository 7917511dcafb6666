"""CPU: (1) the product's constructors + init reproduce the reference's seed-100 weights (init parity, A16);
(2) the oracle reproduces the reference's recorded iteration (losses, gradients) from those weights.
The fixtures were produced by oracle/make_golden.py from the live reference."""
import numpy as np
import pytest
import torch

from helpers import build_product_models, checksum, golden, l2rel, state_to_cpu, synth_batch


def _close(a, b, rel=1e-6, abs_=1e-6):
    return abs(a - b) <= abs_ + rel * abs(b)


@pytest.mark.parametrize("name,conditional", [("tganv2_cond_B8.json", True), ("tganv2_uncond_B8.json", False),
                                              ("tganv2_cond_128x128x32_B8.json", True)])
def test_init_parity_and_oracle_vs_reference(name, conditional):
    """the third case is BASELINE configs[4] (128 x 128 x 32, pyramid 16 / 32 / 64 / 128, 2 x 2 ConvLSTM plane)"""
    import oracle.txt2vid_oracle as O
    fx = golden(name)
    size, frames = fx["config"].get("size", 64), fx["config"].get("frames", 16)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    txt, gen, dis = build_product_models(conditional, V=fx["config"]["V"], seed=fx["config"]["seed"], width=size,
                                         height=size, num_frames=frames)
    sds = {"gen": state_to_cpu(gen), "dis": state_to_cpu(dis), "txt": None if txt is None else state_to_cpu(txt)}
    # ---- init parity: every tensor of the reference's state_dict, same name, same values
    for part in ("gen", "dis", "txt"):
        ref = fx["init"][part]
        if ref is None:
            assert sds[part] is None
            continue
        mine = {k: v for k, v in sds[part].items() if v.dtype.is_floating_point}
        assert set(mine) == set(ref), (sorted(set(mine) ^ set(ref))[:8])
        for k, c in ref.items():
            m = checksum(mine[k])
            assert m["n"] == c["n"], k
            assert _close(m["sum"], c["sum"]) and _close(m["abs"], c["abs"]) and _close(m["wsum"], c["wsum"], abs_=1e-4), k
            assert np.allclose(m["first"], c["first"], rtol=0, atol=0), k
    # ---- the oracle on those weights against the reference's recorded iteration
    B = fx["config"]["B"]
    x, tokens, lengths = synth_batch(B, fx["config"]["V"], T=frames, S=size, seed=fx["config"]["data_seed"])
    assert tokens.tolist() == fx["tokens"] and lengths == fx["lengths"]
    assert _close(checksum(x)["wsum"], fx["x"]["wsum"], abs_=1e-3)
    d = fx["draws"]
    draws = {"bt_real": d["bt_real"], "bt_fake": d["bt_fake"], "perm": d["perm"],
             "alphas": [torch.tensor(a, dtype=torch.float32).view(-1, 1, 1, 1, 1) for a in d["alphas"]]}
    # z is the first CPU-generator draw after the 4 real-pyramid offsets; recover it from the seed stream
    # is not possible here (weights consumed the stream), so the fixture pins z by checksum and we redraw:
    z = _redraw_z(fx, conditional)
    assert _close(checksum(z)["wsum"], fx["z"]["wsum"], abs_=1e-4)
    sd_g, sd_d = O.as_leaves(sds["gen"]), O.as_leaves(sds["dis"])
    sd_t = None if sds["txt"] is None else O.as_leaves(sds["txt"])
    opt_g = O.Adam(O.param_names(sd_g), 2e-4, (0.5, 0.999))
    opt_d = O.Adam(O.param_names(sd_d), 2e-4, (0.5, 0.999))
    out = O.train_iteration(sd_g, sd_d, sd_t, x, tokens, lengths, z, draws, opt_g=opt_g, opt_d=opt_d,
                            frame_sizes=tuple(fx["config"]["frame_sizes"]), num_frames=frames)
    assert abs(out["lossD"] - fx["lossD"]) < 2e-5 and abs(out["lossG"] - fx["lossG"]) < 2e-5
    for lvl, f in zip(out["fake"], fx["fake"]):
        assert list(lvl.shape) == f["shape"]
        assert _close(checksum(lvl)["abs"], f["abs"], rel=1e-4)
    for part in ("gradD", "gradG"):
        assert set(out[part]) == set(fx[part])
        for k, c in fx[part].items():
            n = float(out[part][k].double().norm())
            if c["norm"] < 1e-5:
                assert n < 1e-4, (k, n)
            else:
                assert abs(n - c["norm"]) <= 2e-3 * c["norm"], (k, n, c["norm"])


def _redraw_z(fx, conditional):
    """Replays the reference's RNG stream up to z: seed, construct + init (done by the caller in the same
    process state is gone), so rebuild once more -- cheap relative to the iteration."""
    size, frames = fx["config"].get("size", 64), fx["config"].get("frames", 16)
    build_product_models(conditional, V=fx["config"]["V"], seed=fx["config"]["seed"], width=size, height=size,
                         num_frames=frames)
    import oracle.txt2vid_oracle as O
    bt = O.draw_real(4, True)
    assert bt == fx["draws"]["bt_real"]
    return torch.randn(fx["config"]["B"], 256)


def test_index_fixtures():
    import oracle.txt2vid_oracle as O
    fx = golden("index_fixtures.json")
    for c in fx["subsample"]:
        x = torch.arange(int(np.prod(c["shape"])), dtype=torch.float32).view(c["shape"])
        y = O.subsample(x, c["bt"])
        assert list(y.shape) == c["out_shape"]
        assert y.reshape(-1)[:64].tolist() == c["values"] and float(y.double().sum()) == c["sum"]
    for c in fx["nearest"]:
        x = torch.arange(int(np.prod(c["in"])), dtype=torch.float32).view(c["in"])
        y = O.nearest_resize(x, (c["in"][2], c["fs"], c["fs"]))
        assert y[0, 0, 0, 0].tolist() == c["first_row"] and y[0, 0, 0, :, 0].tolist() == c["first_col"]
        assert float(y.double().sum()) == c["sum"]
    from txt2vid_b200.util import gen_perm
    for c in fx["gen_perm"]:
        np.random.seed(c["seed"])
        assert [int(v) for v in O.gen_perm(c["n"])] == c["perm"] and [int(v) for v in O.gen_perm(c["n"])] == c["perm2"]
        np.random.seed(c["seed"])
        assert [int(v) for v in gen_perm(c["n"])] == c["perm"] and [int(v) for v in gen_perm(c["n"])] == c["perm2"]
    torch.manual_seed(100)
    assert [int(torch.randint(2, (1,))) for _ in range(16)] == fx["randint_stream_seed100"]
