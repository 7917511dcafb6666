"""GPU debug aid: accuracy of the tcgen05 engine in the fp32 storage mode (bf16 part split, T2V_FP32_SPLIT = 3 | 6)
against an fp64 convolution on the CPU, next to cuDNN fp32 (TF32 off) on the same data."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from txt2vid_b200 import kernels as K, ops  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ops.set_precision("fp32")
    print("split terms", K.SPLIT_TERMS)
    for (N, D, H, W), Cin, Cout, k, positive in [((8, 1, 1, 1), 1024, 4096, (1, 1, 1), False),
                                                 ((128, 1, 1, 1), 1024, 1024, (3, 3, 3), False),
                                                 ((8, 1, 8, 8), 1024, 1024, (1, 3, 3), True),
                                                 ((4, 4, 8, 8), 128, 128, (3, 3, 3), True),
                                                 ((2, 2, 16, 16), 64, 64, (3, 3, 3), True)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn(N, D, H, W, Cin, device="cuda", generator=g)
        if positive:
            x = x.relu()
        w = torch.randn(Cout, k[0] * k[1] * k[2], Cin, device="cuda", generator=g) * (2.0 / (Cin * 9)) ** 0.5
        w5 = w.view(Cout, k[0], k[1], k[2], Cin).permute(0, 4, 1, 2, 3).contiguous()
        pad = tuple(kk // 2 for kk in k)
        ref = F.conv3d(x.permute(0, 4, 1, 2, 3).double().cpu(), w5.double().cpu(), None, padding=pad).permute(0, 2, 3, 4, 1)
        lib = F.conv3d(x.permute(0, 4, 1, 2, 3), w5, None, padding=pad).permute(0, 2, 3, 4, 1)
        cpu = F.conv3d(x.permute(0, 4, 1, 2, 3).cpu(), w5.cpu(), None, padding=pad).permute(0, 2, 3, 4, 1)
        y = K.conv_fprop(x, K.pack_weight(w), None, None, k)
        print("N%d %dx%dx%d Cin %d Cout %d k%s relu-in %s: engine %.2e  cudnn-fp32 %.2e  cpu-fp32 %.2e  (mean signed rel err engine %.2e)"
              % (N, D, H, W, Cin, Cout, k, positive, rel(y, ref), rel(lib, ref), rel(cpu, ref),
                 float(((y.double().cpu() - ref) / ref.abs().clamp_min(1e-3)).mean())))
    ops.set_precision("bf16")


if __name__ == "__main__":
    main()
