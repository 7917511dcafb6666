"""One TGAN (config 1) and one TCWYT (config 2) product iteration on the GPU inside a CUDA profiler range -- the target
of `ncu --profile-from-start off` for the launch list of the two families (which kernels their convolutions run on)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import test_families as TF

B = int(os.environ.get("FAM_BATCH", "64"))
gen, dis = TF.build_tgan()
x, z = TF._synth(B, 16, 64, 1), torch.randn(B, 256)
mods = [m.cuda() for m in TF.build_tcwyt()]
x2, z2, cond = TF._synth(B, 16, 48, 2), torch.randn(B, 100), torch.randn(B, 256)
gen, dis = gen.cuda(), dis.cuda()
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    a = TF.product_tgan_iteration(gen, dis, x.cuda(), z.cuda())
    b = TF.product_tcwyt_iteration(mods, x2.cuda(), z2.cuda(), cond.cuda())
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("tgan lossD %.5f lossG %.5f | tcwyt lossD %.5f lossG %.5f" % (a["lossD"], a["lossG"], b["lossD"], b["lossG"]))
