"""GPU parity of the kernel-4 / stride-2 / padding-1 convolutions and transposed convolutions of the TGAN / TCWYT
families on the tcgen05 engine (csrc/s2d.cu + the windowed implicit GEMM), through the C ABI.

Checkers: the block permutation is an index kernel -> bit-exact against a torch pad / reshape / permute composition;
the convolutions -> torch conv3d / conv_transpose3d / conv3d_weight in fp32 (TF32 off) on the same bf16-rounded
operands (reference call sites: models/tcwyt/video_discrim.py:12-25, tcwyt/frame_discrim.py:9-21, tcwyt/gen.py:18-26,
tgan/gen.py:20-23, tgan/temporal_gen.py:112-115), and the CUDA-core general convolution of the same library.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _blocks_ref(x, modes, creal):
    """(N,D,H,W,C) -> (N,D',H',W',round16(P*creal)) with block b of a strided axis = samples (2b-1, 2b)"""
    N, D, H, W, C = x.shape
    xr = x[..., :creal]
    pads = []
    for m in reversed(modes):                     # F.pad order: last axis first (after channels)
        pads += [1, 1] if m == 2 else [0, 0]
    xr = F.pad(xr, [0, 0] + pads)
    f = [2 if m == 2 else 1 for m in modes]
    Db, Hb, Wb = xr.shape[1] // f[0], xr.shape[2] // f[1], xr.shape[3] // f[2]
    xr = xr.reshape(N, Db, f[0], Hb, f[1], Wb, f[2], creal).permute(0, 1, 3, 5, 2, 4, 6, 7)
    xr = xr.reshape(N, Db, Hb, Wb, f[0] * f[1] * f[2] * creal)
    Cp = (xr.shape[-1] + 15) // 16 * 16
    return F.pad(xr, [0, Cp - xr.shape[-1]]).contiguous()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape,modes,creal", [((2, 4, 8, 6, 16), (2, 2, 2), 3), ((3, 1, 6, 12, 32), (1, 2, 2), 32),
                                               ((5, 1, 1, 8, 48), (1, 1, 2), 48), ((2, 2, 4, 4, 16), (2, 2, 2), 10),
                                               ((2, 16, 48, 48, 16), (2, 2, 2), 3)])
def test_block_permutation_is_bit_exact(shape, modes, creal, dtype):
    from txt2vid_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(shape, device="cuda", generator=g).to(dtype)
    xs = K.s2d_shift(x, modes, creal)
    ref = _blocks_ref(x, modes, creal)
    assert xs.shape == ref.shape and torch.equal(xs, ref)
    back = K.d2s_shift(xs, modes, shape[1:4], shape[4], creal)
    want = x.clone()
    want[..., creal:] = 0
    assert torch.equal(back, want)


def _w5(w, k):
    Co, taps, Ci = w.shape
    return w.float().view(Co, k[0], k[1], k[2], Ci).permute(0, 4, 1, 2, 3).contiguous()


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


CASES = [
    # N, in_sp, Cin (padded), cin_real, Cout, k, s, p
    (2, (16, 64, 64), 16, 3, 64, (4, 4, 4), (2, 2, 2), (1, 1, 1)),       # critic layer 1 (RGB), config 1
    (2, (8, 32, 32), 64, 64, 128, (4, 4, 4), (2, 2, 2), (1, 1, 1)),
    (3, (2, 8, 8), 256, 256, 512, (4, 4, 4), (2, 2, 2), (1, 1, 1)),      # (1, 4, 4) output
    (2, (16, 48, 48), 16, 3, 64, (4, 4, 4), (2, 2, 2), (1, 1, 1)),       # TCWYT 48 x 48 (non power-of-two planes)
    (2, (2, 6, 6), 128, 128, 256, (4, 4, 4), (2, 2, 2), (1, 1, 1)),      # (1, 3, 3) output
    (5, (1, 48, 48), 16, 3, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1)),        # frame critic, 2-D
    (32, (1, 8, 8), 256, 256, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1)),     # = ConvTranspose2d(128 -> 256) reversed
    (4, (1, 64, 64), 32, 32, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1)),
    (8, (1, 1, 2), 256, 256, 512, (1, 1, 4), (1, 1, 2), (0, 0, 1)),      # temporal generator, 1-D, length 2 <-> 1
    (8, (1, 1, 16), 256, 256, 128, (1, 1, 4), (1, 1, 2), (0, 0, 1)),
    # stride-1 "same" layers and whole-input kernels of the two families run on the engine as they are
    (6, (1, 64, 64), 16, 3, 32, (1, 3, 3), (1, 1, 1), (0, 1, 1)),        # = ConvTranspose2d(32 -> 3, 3, 1, 1), tgan/gen.py:24
    (2, (16, 48, 48), 16, 3, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0)),       # = ConvTranspose3d(64 -> 3, 1), tcwyt/gen.py:30
    (4, (2, 6, 6), 512, 512, 368, (2, 6, 6), (1, 1, 1), (0, 0, 0)),      # = ConvTranspose3d(356 -> 512, (2,6,6)), tcwyt/gen.py:14
    # single-output-position heads (Cout = 1, padded to 16): the corner window gathered, then a Linear on the engine
    (9, (1, 3, 3), 512, 512, 16, (1, 2, 2), (1, 2, 2), (0, 0, 0)),       # Conv2d(512, 1, 2, 2, 0) on 3 x 3, frame_discrim.py:55
    (9, (1, 4, 4), 512, 512, 16, (1, 3, 3), (2, 2, 2), (0, 0, 0)),       # Conv3d(512, 1, (1,3,3), 2, 0) on (1,4,4), video_discrim.py:46
]


def _mk(case, seed=0):
    N, in_sp, Cin, creal, Cout, k, s, p = case
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((N,) + in_sp + (Cin,), device="cuda", generator=g)
    x[..., creal:] = 0
    taps = k[0] * k[1] * k[2]
    w = torch.randn((Cout, taps, Cin), device="cuda", generator=g) / (taps * creal) ** 0.5
    w[..., creal:] = 0
    out_sp = tuple((i + 2 * pp - kk) // ss + 1 for i, kk, ss, pp in zip(in_sp, k, s, p))
    dy = torch.randn((N,) + out_sp + (Cout,), device="cuda", generator=g)
    return x.to(torch.bfloat16), w.to(torch.bfloat16), dy.to(torch.bfloat16)


@pytest.mark.parametrize("case", CASES)
def test_strided_conv_fprop_on_the_engine(case, monkeypatch):
    from txt2vid_b200 import kernels as K
    N, in_sp, Cin, creal, Cout, k, s, p = case
    x, w, _ = _mk(case)
    assert K.s2d_modes(k, s, p, in_sp) is not None or K._engine_route(x, k, s, p, in_sp) is not None
    bias = torch.linspace(-1, 1, Cout, device="cuda")
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), _w5(w, k), bias, stride=s, padding=p).permute(0, 2, 3, 4, 1)
    n0 = K.lib().t2v_launch_count()
    y32 = K.gconv_fprop(x, w, bias, k, s, p, out_f32=True, cin_real=creal)
    assert K.lib().t2v_launch_count() - n0 >= 1
    y16 = K.gconv_fprop(x, w, bias, k, s, p, cin_real=creal)
    assert y32.dtype == torch.float32 and y16.dtype == torch.bfloat16 and y32.shape == ref.shape
    assert _rel(y32, ref) < 2e-3, _rel(y32, ref)
    assert _rel(y16, ref) < 1e-2, _rel(y16, ref)
    monkeypatch.setattr(K, "GCONV_TC", False)              # the CUDA-core general convolution of the same library
    y_simt = K.gconv_fprop(x, w, bias, k, s, p, out_f32=True, cin_real=creal)
    assert _rel(y32, y_simt) < 2e-3


@pytest.mark.parametrize("case", CASES)
def test_transposed_conv_and_data_gradient_on_the_engine(case, monkeypatch):
    from txt2vid_b200 import kernels as K
    N, in_sp, Cin, creal, Cout, k, s, p = case
    _, w, dy = _mk(case, seed=1)
    bias = torch.zeros(Cin, device="cuda")
    window = K._engine_route(dy, k, s, p, in_sp) == "window"
    if not window:               # (a head's data gradient carries no bias; with one the layer stays on the CUDA cores)
        bias[:creal] = torch.linspace(-0.5, 0.5, creal, device="cuda")
    osp = [(o - 1) * ss - 2 * pp + kk for o, ss, pp, kk in zip(dy.shape[1:4], s, p, k)]
    ref = F.conv_transpose3d(dy.float().permute(0, 4, 1, 2, 3), _w5(w, k), bias, stride=s, padding=p,
                             output_padding=tuple(i - o for i, o in zip(in_sp, osp)))
    ref = ref.permute(0, 2, 3, 4, 1)
    dx32 = K.gconv_dgrad(dy, w, None if window else bias, in_sp, k, s, p, out_f32=True, cin_real=creal)
    dx16 = K.gconv_dgrad(dy, w, None, in_sp, k, s, p, cin_real=creal)
    assert tuple(dx32.shape) == (N,) + in_sp + (Cin,) and dx32.dtype == torch.float32
    assert _rel(dx32, ref) < 2e-3, _rel(dx32, ref)
    assert _rel(dx16.float() + bias, ref) < 1e-2
    assert float(dx32[..., creal:].abs().max()) == 0.0 if creal < Cin else True
    monkeypatch.setattr(K, "GCONV_TC", False)
    dx_simt = K.gconv_dgrad(dy, w, None if window else bias, in_sp, k, s, p, out_f32=True, cin_real=creal)
    assert _rel(dx32, dx_simt) < 2e-3


@pytest.mark.parametrize("case", CASES)
def test_strided_conv_wgrad_on_the_engine(case, monkeypatch):
    from txt2vid_b200 import kernels as K
    N, in_sp, Cin, creal, Cout, k, s, p = case
    x, _, dy = _mk(case, seed=2)
    ref = torch.nn.grad.conv3d_weight(x.float().permute(0, 4, 1, 2, 3), (Cout, Cin) + tuple(k),
                                      dy.float().permute(0, 4, 1, 2, 3), stride=s, padding=p)
    ref = ref.permute(0, 2, 3, 4, 1).reshape(Cout, k[0] * k[1] * k[2], Cin)
    dw = K.gconv_wgrad(dy, x, k, s, p, cin_real=creal)
    assert dw.shape == ref.shape and dw.dtype == torch.float32
    assert _rel(dw, ref) < 2e-3, _rel(dw, ref)
    monkeypatch.setattr(K, "GCONV_TC", False)
    dw_simt = K.gconv_wgrad(dy, x, k, s, p, cin_real=creal)
    assert _rel(dw, dw_simt) < 2e-3
