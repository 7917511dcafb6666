"""What does bf16 COMPUTE cost on this network, independent of any kernel of this repo?

Runs the CPU oracle (oracle/txt2vid_oracle.py, the restatement of the reference pinned to the live reference) twice
on the same weights / inputs / host-RNG draws: in fp32, and under torch.autocast(bfloat16) (conv / linear / bmm
operands and outputs rounded to bf16 by ATen, normalisation and losses in fp32: the standard mixed-precision recipe).
The deviations between the two are the floor any bf16 tensor-core implementation of the step sits on; the GPU
iteration tests pin the product at 1.5x these numbers (tests/test_iteration_gpu.py).

Usage: python scripts/bf16_floor.py [cond|uncond] [attention gamma] [device] -> JSON on stdout (committed as
profiles/r02_bf16_floor_*.json).  On the CPU the bf16 ATen convolutions take ~1 h; on a B200 (device = cuda: cuDNN
bf16 kernels under autocast against cuDNN fp32 with TF32 off) a few seconds."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle.txt2vid_oracle as O  # noqa: E402
from helpers import build_product_models, golden, l2rel, state_to_cpu, synth_batch  # noqa: E402
from test_product_vs_oracle_cpu import grad_stats  # noqa: E402


def one(conditional, autocast, fx, attn_gamma=0.0, device="cpu"):
    B, V = fx["config"]["B"], fx["config"]["V"]
    txt, gen, dis = build_product_models(conditional, V=V, seed=fx["config"]["seed"])
    if attn_gamma:                 # the perturbation of the "attention on" parity tests (gamma = 0.5, BatchNorm affine)
        from test_product_vs_oracle_cpu import perturb_models
        perturb_models(gen, dis)
    mv = lambda sd: None if sd is None else {k: v.to(device) for k, v in sd.items()}
    sds = {"gen": mv(state_to_cpu(gen)), "dis": mv(state_to_cpu(dis)),
           "txt": None if txt is None else mv(state_to_cpu(txt))}
    x, tokens, lengths = synth_batch(B, V, seed=fx["config"]["data_seed"])
    x, tokens = x.to(device), tokens.to(device)
    bt_real = O.draw_real(4, True)
    z = torch.randn(B, 256).to(device)
    draws = O.draw_rest([B, B // 2, B // 4, B // 8], conditional=conditional, gp=True)
    draws["bt_real"] = bt_real
    sd_g, sd_d = O.as_leaves(sds["gen"]), O.as_leaves(sds["dis"])
    sd_t = None if sds["txt"] is None else O.as_leaves(sds["txt"])
    opt_g = O.Adam(O.param_names(sd_g), 2e-4, (0.5, 0.999))
    opt_d = O.Adam(O.param_names(sd_d), 2e-4, (0.5, 0.999))
    with torch.autocast(torch.device(device).type, dtype=torch.bfloat16, enabled=autocast, cache_enabled=False):
        out = O.train_iteration(sd_g, sd_d, sd_t, x, tokens, lengths, z, draws, opt_g=opt_g, opt_d=opt_d)
    cpu = lambda d: {k: v.detach().float().cpu() for k, v in d.items()}
    out["gradD"], out["gradG"] = cpu(out["gradD"]), cpu(out["gradG"])
    out["fake"] = [f.float().cpu() for f in out["fake"]]
    return out


def main():
    conditional = (sys.argv[1] if len(sys.argv) > 1 else "cond") == "cond"
    gamma = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    device = sys.argv[3] if len(sys.argv) > 3 else "cpu"
    # kind "autocast": bf16 floor (default).  kind "fp32dev": the FP32 floor -- the same fp32 oracle on two fp32
    # back ends (ATen CPU kernels vs cuDNN / cuBLAS with TF32 off): how far two correct fp32 implementations of the
    # reference sit from each other on this network (ReLU-mask flips turn 1e-7 forward differences into ~sqrt of that
    # on the gradients), i.e. what a 1e-3 gradient bar means here.
    kind = sys.argv[4] if len(sys.argv) > 4 else "autocast"
    if device != "cpu":
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    fx = golden("tganv2_cond_B8.json" if conditional else "tganv2_uncond_B8.json")
    st_t, st_n = torch.get_rng_state(), np.random.get_state()
    ref = one(conditional, False, fx, gamma, "cpu" if kind == "fp32dev" else device)
    torch.set_rng_state(st_t)
    np.random.set_state(st_n)
    low = one(conditional, kind != "fp32dev", fx, gamma, device)
    what = "oracle in fp32 on %s (TF32 off) vs the same oracle in fp32 on the CPU" % device if kind == "fp32dev" else \
        "oracle under torch.autocast(%s, bfloat16) vs the same oracle in fp32 (TF32 off)" % device
    rep = {"what": what,
           "model": "tganv2_%s_B8" % ("cond" if conditional else "uncond"), "attention_gamma": gamma,
           "lossD": abs(low["lossD"] - ref["lossD"]) / abs(ref["lossD"]),
           "lossG": abs(low["lossG"] - ref["lossG"]) / abs(ref["lossG"]),
           "fake": max(l2rel(a.float(), b) for a, b in zip(low["fake"], ref["fake"])),
           "gradD": grad_stats({k: v.float() for k, v in low["gradD"].items()}, ref["gradD"]),
           "gradG": grad_stats({k: v.float() for k, v in low["gradG"].items()}, ref["gradG"])}
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
