"""Where does the bf16 gradient error of one TGANv2-cond iteration come from?  (CPU, tests/cpu_kernels.py emulation.)

Runs the product's host logic + autograd formulas with the emulated kernels and separately switchable storage
dtypes for (a) forward activations / operand packs and (b) the tensors of the backward chain, against the fp32
oracle.  Usage: python scripts/bf16_error_budget.py [fwd_dtype bwd_dtype]..."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpu_kernels  # noqa: E402
from helpers import golden  # noqa: E402
from test_product_vs_oracle_cpu import grad_stats, run_product_iteration  # noqa: E402
from helpers import l2rel  # noqa: E402

BWD_FUNCS = ["conv_dgrad", "relu_bwd", "avgpool_bwd", "upsample2x_bwd", "bn_backward", "attention_bwd", "render_bwd",
             "broadcast_spatial", "lstm_cell_bwd", "leaky_relu_bwd", "tanh_bwd", "conv_dgrad_sd2", "scatter_frames"]
DT = {"bf16": torch.bfloat16, "fp32": torch.float32}


def install(fwd, bwd):
    from txt2vid_b200 import ops, optim, trainer
    for mod in (ops, optim, trainer):
        mod.K = cpu_kernels
    ops.PACKS.clear()
    cpu_kernels.set_store_dtype(fwd)
    for name in BWD_FUNCS:
        orig = getattr(cpu_kernels, "_orig_" + name, None) or getattr(cpu_kernels, name)
        setattr(cpu_kernels, "_orig_" + name, orig)

        def wrap(*a, _f=orig, **k):
            cpu_kernels.set_store_dtype(bwd)
            try:
                if bwd == torch.float32:
                    a = tuple(t.float() if isinstance(t, torch.Tensor) and t.dtype == torch.bfloat16 and False else t
                              for t in a)
                return _f(*a, **k)
            finally:
                cpu_kernels.set_store_dtype(fwd)
        setattr(cpu_kernels, name, wrap)


def main():
    combos = [("bf16", "bf16"), ("bf16", "fp32"), ("fp32", "bf16")]
    if len(sys.argv) > 2:
        combos = list(zip(sys.argv[1::2], sys.argv[2::2]))
    out = {}
    for f, b in combos:
        install(DT[f], DT[b])
        orc, got = run_product_iteration(True, golden("tganv2_cond_B8.json"), "cpu")
        rep = {"lossD": abs(got["lossD"] - orc["lossD"]) / abs(orc["lossD"]),
               "lossG": abs(got["lossG"] - orc["lossG"]) / abs(orc["lossG"]),
               "fake": max(l2rel(a, c) for a, c in zip(got["fake"], orc["fake"])),
               "gradD": grad_stats(got["gradD"], orc["gradD"]), "gradG": grad_stats(got["gradG"], orc["gradG"])}
        out["fwd_%s_bwd_%s" % (f, b)] = rep
        print(f, b, json.dumps(rep), flush=True)
    return out


if __name__ == "__main__":
    main()
