"""Second error-budget experiment (CPU emulation): which FORWARD rounding points carry the bf16 gradient error?
Variants:  E1  generator conv outputs kept fp32 (conv operands still rounded to bf16), everything else bf16
           E2  generator forward all-fp32, discriminator bf16
           E3  discriminator forward fp32 (D==T>1 convs), generator bf16"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpu_kernels as CK  # noqa: E402
from helpers import golden, l2rel  # noqa: E402
from test_product_vs_oracle_cpu import grad_stats, run_product_iteration  # noqa: E402

BF, F32 = torch.bfloat16, torch.float32
orig_conv = CK.conv_fprop


def run(tag):
    from txt2vid_b200 import ops, optim, trainer
    for mod in (ops, optim, trainer):
        mod.K = CK
    ops.PACKS.clear()
    orc, got = run_product_iteration(True, golden("tganv2_cond_B8.json"), "cpu")
    rep = {"lossD": abs(got["lossD"] - orc["lossD"]) / abs(orc["lossD"]),
           "lossG": abs(got["lossG"] - orc["lossG"]) / abs(orc["lossG"]),
           "fake": max(l2rel(a, c) for a, c in zip(got["fake"], orc["fake"])),
           "gradD": grad_stats(got["gradD"], orc["gradD"]), "gradG": grad_stats(got["gradG"], orc["gradG"])}
    print(tag, json.dumps(rep), flush=True)


def e1():
    def conv(x, w, bias=None, residual=None, k=(3, 3, 3), relu=False, out_f32=False, algo=0):
        is_g = x.shape[1] == 1 and x.shape[2] == x.shape[3] and not out_f32 and torch.is_grad_enabled() is False
        xr = x.to(BF)
        return orig_conv(xr, w, bias, residual, k, relu, out_f32 or (x.shape[1] == 1), algo)
    CK.conv_fprop = conv
    run("E1 G conv outputs fp32")
    CK.conv_fprop = orig_conv


which = sys.argv[1:] or ["e1"]
for wname in which:
    globals()[wname]()
