"""Raw kernel calls: torch tensors in, torch tensors out, one C-ABI call each.

Tensors here are in the kernels' native layouts ("CL" = contiguous (N, D, H, W, C) bf16; 2-D
feature maps use D = 1).  Everything above this file (ops.py autograd formulas, the module mirrors)
only talks to the GPU through these functions.
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvGeom, check, lib, ptr, require_cuda, stream

BF16 = torch.bfloat16


def _geom(N, D, H, W, Cin, Cout, k):
    kd, kh, kw = k
    return ConvGeom(N, D, H, W, Cin, Cout, kd, kh, kw)


def conv_fprop(x, w, bias=None, residual=None, k=(3, 3, 3), relu=False, out_f32=False, algo=0):
    """x (N,D,H,W,Cin) bf16, w (Cout,taps,Cin) bf16 -> y (N,D,H,W,Cout) bf16|f32."""
    require_cuda(x, w, bias, residual)
    N, D, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == k[0] * k[1] * k[2] and w.shape[2] == Cin, (w.shape, k, Cin)
    assert x.is_contiguous() and w.is_contiguous() and x.dtype == BF16 and w.dtype == BF16
    y = torch.empty((N, D, H, W, Cout), device=x.device, dtype=torch.float32 if out_f32 else BF16)
    g = _geom(N, D, H, W, Cin, Cout, k)
    flags = (_lib.EPI_RELU if relu else 0) | (_lib.EPI_OUT_F32 if out_f32 else 0)
    check(lib().t2v_conv_fprop(ctypes.byref(g), ptr(x), ptr(w), ptr(bias), ptr(residual), ptr(y), flags, algo,
                               stream()), "t2v_conv_fprop")
    return y


def conv_dgrad(dy, wT, k=(3, 3, 3), residual=None, relu=False, out_f32=False, algo=0):
    """dy (N,D,H,W,Cout) bf16, wT (Cin,taps,Cout) bf16 (from pack_dgrad_weight) -> dx (N,D,H,W,Cin)."""
    require_cuda(dy, wT, residual)
    N, D, H, W, Cout = dy.shape
    Cin = wT.shape[0]
    assert wT.shape[2] == Cout and dy.is_contiguous() and wT.is_contiguous()
    dx = torch.empty((N, D, H, W, Cin), device=dy.device, dtype=torch.float32 if out_f32 else BF16)
    g = _geom(N, D, H, W, Cin, Cout, k)
    flags = (_lib.EPI_RELU if relu else 0) | (_lib.EPI_OUT_F32 if out_f32 else 0)
    check(lib().t2v_conv_dgrad(ctypes.byref(g), ptr(dy), ptr(wT), ptr(residual), ptr(dx), flags, algo, stream()),
          "t2v_conv_dgrad")
    return dx


def conv_wgrad(dy, x, k=(3, 3, 3), out=None, accumulate=False, algo=0):
    """dw (Cout,taps,Cin) fp32 = sum_pos dy[pos,co] x[pos+tap,ci]."""
    require_cuda(dy, x)
    N, D, H, W, Cout = dy.shape
    Cin = x.shape[-1]
    assert x.shape[:4] == dy.shape[:4] and dy.is_contiguous() and x.is_contiguous()
    taps = k[0] * k[1] * k[2]
    if out is None:
        assert not accumulate
        out = torch.empty((Cout, taps, Cin), device=x.device, dtype=torch.float32)
    g = _geom(N, D, H, W, Cin, Cout, k)
    check(lib().t2v_conv_wgrad(ctypes.byref(g), ptr(dy), ptr(x), ptr(out), 1 if accumulate else 0, algo, stream()),
          "t2v_conv_wgrad")
    return out


def cast_bf16(src):
    require_cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous()
    dst = torch.empty(src.shape, device=src.device, dtype=BF16)
    check(lib().t2v_cast_f32_to_bf16(ptr(src), ptr(dst), src.numel(), stream()), "t2v_cast_f32_to_bf16")
    return dst


def cast_f32(src):
    require_cuda(src)
    assert src.dtype == BF16 and src.is_contiguous()
    dst = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    check(lib().t2v_cast_bf16_to_f32(ptr(src), ptr(dst), src.numel(), stream()), "t2v_cast_bf16_to_f32")
    return dst


def pack_dgrad_weight(w):
    """w (Cout,taps,Cin) fp32 -> (Cin,taps,Cout) bf16 with the tap order reversed."""
    require_cuda(w)
    assert w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 3
    Cout, taps, Cin = w.shape
    wT = torch.empty((Cin, taps, Cout), device=w.device, dtype=BF16)
    check(lib().t2v_pack_dgrad_weight(ptr(w), ptr(wT), Cout, taps, Cin, stream()), "t2v_pack_dgrad_weight")
    return wT
