"""CPU: the ALGORITHM behind the tensor-core route of the kernel-4 / stride-2 / padding-1 layers (csrc/s2d.cu),
checked on the executable spec of its index kernels (tests/cpu_kernels.py): a k4 s2 p1 convolution equals a dense
kernel-2 stride-1 convolution over the shifted space-to-depth block tensor with the embedded weights; the transposed
convolution equals the same contraction with the transposed operand followed by the inverse block permutation; the
weight gradient is the block convolution's weight gradient gathered back.  (The GPU tests check the real kernels against
the same torch references: tests/test_strided_tc_gpu.py.)"""
import pytest
import torch
import torch.nn.functional as F

import cpu_kernels as C

CASES = [((2, 4, 8, 8), 5, 7, (2, 2, 2)), ((3, 1, 6, 10), 3, 4, (1, 2, 2)), ((2, 1, 1, 8), 6, 5, (1, 1, 2)),
         ((1, 2, 4, 6), 16, 8, (2, 2, 2))]


def _ksp(modes):
    return tuple(4 if m == 2 else 1 for m in modes), tuple(2 if m == 2 else 1 for m in modes), \
        tuple(1 if m == 2 else 0 for m in modes)


def _block_conv(xs, we, modes, out_sp, taps):
    """dense convolution over the block tensor: output o reads blocks o + t - 1 for the live taps t of a strided axis
    (zero outside the block tensor) -- what t2v_conv_fprop_win computes"""
    N, Cp = xs.shape[0], xs.shape[-1]
    ke = [3 if m == 2 else 1 for m in modes]
    w5 = we.reshape(we.shape[0], ke[0], ke[1], ke[2], Cp).permute(0, 4, 1, 2, 3)
    xp = F.pad(xs.permute(0, 4, 1, 2, 3), [1 if m == 2 else 0 for m in reversed(modes) for _ in (0, 1)])
    y = F.conv3d(xp, w5)                                   # "same" 3-tap correlation on the padded block tensor
    return y[:, :, :out_sp[0], :out_sp[1], :out_sp[2]].permute(0, 2, 3, 4, 1)


@pytest.mark.parametrize("shape,Cin,Cout,modes", CASES)
def test_strided_conv_is_a_dense_kernel2_conv_over_shifted_blocks(shape, Cin, Cout, modes):
    torch.manual_seed(0)
    N, D, H, W = shape
    k, s, p = _ksp(modes)
    Cp = (Cin + 15) // 16 * 16
    x = torch.zeros(N, D, H, W, Cp, dtype=torch.float64)
    x[..., :Cin] = torch.randn(N, D, H, W, Cin, dtype=torch.float64)
    taps = k[0] * k[1] * k[2]
    w = torch.zeros(Cout, taps, Cp, dtype=torch.float64)
    w[..., :Cin] = torch.randn(Cout, taps, Cin, dtype=torch.float64)
    w5 = w.reshape(Cout, k[0], k[1], k[2], Cp).permute(0, 4, 1, 2, 3)
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3), w5, stride=s, padding=p).permute(0, 2, 3, 4, 1)
    out_sp = ref.shape[1:4]
    xs = C.s2d_shift(x, modes, Cin)
    # the block permutation is a bijection on the real channels
    assert torch.equal(C.d2s_shift(xs, modes, (D, H, W), Cp, Cin)[..., :Cin], x[..., :Cin])
    we = C.s2d_embed_weight(w, modes, Cin)
    # tap 0 of every strided axis is never used: the layer is a DENSE kernel-2 convolution
    ke = [3 if m == 2 else 1 for m in modes]
    w6 = we.reshape(Cout, ke[0], ke[1], ke[2], -1)
    for ax, m in enumerate(modes):
        if m == 2:
            assert float(w6.select(1 + ax, 0).abs().max()) == 0.0
    y = _block_conv(xs, we, modes, out_sp, None)
    assert torch.allclose(y, ref, atol=1e-10)
    # transposed convolution = data gradient: the same contraction with the transposed operand, then the inverse permute
    dy = torch.randn(ref.shape, dtype=torch.float64)
    ref_dx = F.conv_transpose3d(dy.permute(0, 4, 1, 2, 3), w5, stride=s, padding=p).permute(0, 2, 3, 4, 1)
    weT = C.s2d_embed_weight(w, modes, Cin, transposed=True)                 # (Cp_blocks, taps reversed, Cout)
    wT5 = weT.reshape(weT.shape[0], ke[0], ke[1], ke[2], Cout).permute(0, 4, 1, 2, 3)
    # output extents O + 1 on the strided axes (block b reads dy[b - 1], dy[b]): pad one in front, two behind
    dyp = F.pad(dy.permute(0, 4, 1, 2, 3), [v for m in reversed(modes) for v in ((1, 2) if m == 2 else (0, 0))])
    dxs = F.conv3d(dyp, wT5)
    dxs = dxs[:, :, :xs.shape[1], :xs.shape[2], :xs.shape[3]].permute(0, 2, 3, 4, 1)
    dx = C.d2s_shift(dxs.contiguous(), modes, (D, H, W), Cp, Cin)
    assert torch.allclose(dx[..., :Cin], ref_dx[..., :Cin], atol=1e-10)
    # weight gradient: the block convolution's weight gradient, gathered back to the 4-tap layout
    xs_ = xs.clone().requires_grad_(False)
    we_ = we.clone().requires_grad_(True)
    (_block_conv(xs_, we_, modes, out_sp, None) * dy).sum().backward()
    dw = C.s2d_extract_wgrad(we_.grad, modes, Cp, Cin)
    ref_dw = torch.nn.grad.conv3d_weight(x.permute(0, 4, 1, 2, 3), w5.shape, dy.permute(0, 4, 1, 2, 3), stride=s, padding=p)
    ref_dw = ref_dw.permute(0, 2, 3, 4, 1).reshape(Cout, taps, Cp)
    assert torch.allclose(dw[..., :Cin], ref_dw[..., :Cin], atol=1e-9)
