"""Launches the conv-engine kernels on the TGANv2 hot shapes (batch 256 / GPU) -- the target of
`ncu --set full` (see scripts/gpu_ncu_full.sh).  Each kernel runs twice (first = warm-up)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from txt2vid_b200 import kernels as K

def mk(N, D, H, W, Cin, Cout, k, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((N, D, H, W, Cin), device="cuda", generator=g).to(torch.bfloat16)
    taps = k[0] * k[1] * k[2]
    w = (torch.randn((Cout, taps, Cin), device="cuda", generator=g) / (taps * Cin) ** 0.5).to(torch.bfloat16)
    dy = torch.randn((N, D, H, W, Cout), device="cuda", generator=g).to(torch.bfloat16)
    return x, w, dy

shapes = [
    ("stem conv2 64->64 3^3 L1", (128, 8, 16, 16, 64, 64, (3, 3, 3))),
    ("G up0 conv1 1024->512 @2x2", (4096, 1, 2, 2, 1024, 512, (1, 3, 3))),
    ("G up2 conv1 256->128 @8x8", (4096, 1, 8, 8, 256, 128, (1, 3, 3))),
    ("D down0 conv2 64->128", (256, 8, 4, 4, 64, 128, (3, 3, 3))),
]
for name, case in shapes:
    x, w, dy = mk(*case)
    k = case[6]
    for _ in range(2):
        K.conv_fprop(x, w, k=k)
        K.conv_wgrad(dy, x, k=k)
    torch.cuda.synchronize()
    print("ran", name)
