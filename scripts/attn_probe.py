"""Times the generator's non-local block core at batch 1024 (1024 maps of 32x32, C = 32): fused kernels vs the
composite (max-pool + bmm + softmax + bmm through ATen)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from txt2vid_b200 import kernels as K

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
def padded(c):
    t = torch.zeros(N, 1, 32, 32, 16, device="cuda")
    t[..., :c] = torch.randn(N, 1, 32, 32, c, device="cuda")
    return t.to(torch.bfloat16)
theta, phi, g, do = padded(4), padded(4), padded(16), padded(16)

def t(fn, n=10):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def composite():
    th = theta[..., :4].float().requires_grad_(True)
    ph = phi[..., :4].float().requires_grad_(True)
    gg = g.float().requires_grad_(True)
    mp = lambda v: F.max_pool3d(v.permute(0, 4, 1, 2, 3), [1, 2, 2]).permute(0, 2, 3, 4, 1)
    beta = torch.softmax(torch.bmm(th.reshape(N, 1024, 4), mp(ph).reshape(N, 256, 4).transpose(1, 2)), -1)
    o = torch.bmm(beta, mp(gg).reshape(N, 256, 16))
    o.backward(do.float().reshape(N, 1024, 16))

print("fused fwd  %.3f ms" % t(lambda: K.attention_fwd(theta, phi, g, 4, 16)))
print("fused bwd  %.3f ms" % t(lambda: K.attention_bwd(theta, phi, g, do, 4, 16)))
print("composite fwd+bwd %.3f ms" % t(composite, 5))
