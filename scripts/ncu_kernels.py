"""Launches the conv-engine and attention kernels on the TGANv2 hot shapes (batch 1024 / GPU) -- the target of
`ncu --set full` (see scripts/gpu_ncu_full.sh).  Each kernel runs twice (first = warm-up, outside the profiler range); NVTX-free: the order of
launches is the order of the `ran` lines, the summary script labels them by index."""
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

REPS = 2      # first rep = warm-up (outside the profiler range), second rep is captured
SCALE = int(os.environ.get("BATCH_SCALE", "1"))     # 2: the shapes of the default bench batch (2048 / GPU)
shapes = [
    ("stem conv2 64->64 3^3 level 0 (stride 1)", (1024, 16, 8, 8, 64, 64, (3, 3, 3))),
    ("G up2 conv1 256->128 @8x8", (16384, 1, 8, 8, 256, 128, (1, 3, 3))),
    ("G up0 conv1 1024->512 @2x2", (16384, 1, 2, 2, 1024, 512, (1, 3, 3))),
    ("D down0 conv1 64->64 (512,4,8,8) generic", (512, 4, 8, 8, 64, 64, (3, 3, 3))),
    ("D down0 conv2 64->128 (1024,8,4,4)", (1024, 8, 4, 4, 64, 128, (3, 3, 3))),
    ("G level3 32->32 @64x64", (256, 1, 64, 64, 32, 32, (1, 3, 3))),
    ("stem conv1 as 1x1 GEMM 96->64 level 1", (512, 8, 16, 16, 96, 64, (1, 1, 1))),
]
shapes = [(n, (c[0] * SCALE,) + tuple(c[1:])) for n, c in shapes]
for name, case in shapes:
    x, w, dy = mk(*case)
    k = case[6]
    for r in range(REPS):
        if r == REPS - 1:
            torch.cuda.synchronize(); torch.cuda.profiler.start()
        K.conv_fprop(x, w, k=k)
        K.conv_wgrad(dy, x, k=k)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
    print("ran", name, "| launches per rep: fprop, [memset], wgrad")
    del x, w, dy
# stride-(2,1,1) stem convolution, levels 1 and 3
for shp in [(512, 8, 16, 16), (128, 2, 64, 64)]:
    x, w, _ = mk(*shp, 64, 64, (3, 3, 3))
    dyh = torch.randn((shp[0], shp[1] // 2) + shp[2:] + (64,), device="cuda").to(torch.bfloat16)
    wT = K.pack_dgrad_weight(w.float())
    for r in range(REPS):
        if r == REPS - 1:
            torch.cuda.synchronize(); torch.cuda.profiler.start()
        K.conv_fprop_sd2(x, w)
        K.conv_dgrad_sd2(dyh, wT)
        K.conv_wgrad_sd2(dyh, x)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
    print("ran sd2", shp, "| launches per rep: fprop_sd2, dgrad even planes, dgrad odd planes, wgrad_sd2")
    del x, w, dyh
# RGB stem conv, direct kernels, level 1 at batch 1024
x = torch.rand((512, 3, 8, 16, 16), device="cuda") * 2 - 1
_, xc4 = K.rgb_to_cl(x)
wp = K.stem_pack_weight(torch.randn((64, 27, 3), device="cuda") / 9)
dy = torch.randn((512, 8, 16, 16, 64), device="cuda").to(torch.bfloat16)
for r in range(REPS):
    if r == REPS - 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    K.stem_fprop(xc4, wp)
    K.stem_wgrad(dy, xc4)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ran RGB stem conv (512,8,16,16) | launches per rep: stem_fprop, [memset], stem_wgrad")
# generator non-local block core, 1024 maps of 32x32
def padded(c, N=1024):
    t = torch.zeros(N, 1, 32, 32, 16, device="cuda")
    t[..., :c] = torch.randn(N, 1, 32, 32, c, device="cuda")
    return t.to(torch.bfloat16)
theta, phi, g, do = padded(4), padded(4), padded(16), padded(16)
for r in range(REPS):
    if r == REPS - 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    K.attention_fwd(theta, phi, g, 4, 16)
    K.attention_bwd(theta, phi, g, do, 4, 16)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ran attention core 1024 x (32x32), c8=4 c2=16 | launches per rep: fwd, bwd")
