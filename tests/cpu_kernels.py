"""TEST-ONLY executable specification of txt2vid_b200/kernels.py in plain PyTorch.

Same function names, argument meaning, layouts and rounding points (bf16 storage, fp32 math) as the
C-ABI wrappers, but runs on any device.  Two uses:
  * `-m "not gpu"` tests monkeypatch `txt2vid_b200.ops.K` with this module to check the autograd
    formulas (incl. the double backward the gradient penalty needs) against the oracle on CPU;
  * `-m gpu` tests compare every real kernel against these functions on the same inputs.
The product never imports this file.
"""
import torch
import torch.nn.functional as F

F32 = torch.float32
# Storage dtype of activations / operand packs.  bf16 = the product's rounding points; the CPU tests also
# run with float32 storage to check the autograd formulas against the oracle at the fp32 bar (1e-3).
STORE = torch.bfloat16


def set_store_dtype(dt):
    global STORE
    STORE = dt


def _w5(w, k):
    Cout, taps, Cin = w.shape
    return w.float().view(Cout, k[0], k[1], k[2], Cin).permute(0, 4, 1, 2, 3)


def _pad(k):
    return (k[0] // 2, k[1] // 2, k[2] // 2)


def conv_fprop(x, w, bias=None, residual=None, k=(3, 3, 3), relu=False, out_f32=False, algo=0):
    y = F.conv3d(x.float().permute(0, 4, 1, 2, 3), _w5(w, k), bias, padding=_pad(k)).permute(0, 2, 3, 4, 1)
    if residual is not None:
        y = y + residual.float()
    if relu:
        y = torch.relu(y)
    y = y.contiguous()
    return y if out_f32 else y.to(STORE)


def stem_pack_weight(w3):
    wp = torch.zeros((64, 32, 4), dtype=torch.float32)
    wp[:, :27, :3] = w3
    return wp.reshape(64, 128).to(STORE).contiguous()


def _stem_w5(wp):
    w3 = wp.float()[:, :108].reshape(64, 27, 4)[..., :3]          # (co, tap, c)
    return w3.reshape(64, 3, 3, 3, 3).permute(0, 4, 1, 2, 3).contiguous()


def stem_fprop(xc, wp, bias=None, relu=True):
    y = F.conv3d(xc.float()[..., :3].permute(0, 4, 1, 2, 3), _stem_w5(wp), bias, padding=1).permute(0, 2, 3, 4, 1)
    if relu:
        y = torch.relu(y)
    return y.contiguous().to(STORE)


def stem_wgrad(dy, xc, out=None, accumulate=False):
    xin = xc.float()[..., :3].permute(0, 4, 1, 2, 3)
    g = torch.nn.grad.conv3d_weight(xin, (64, 3, 3, 3, 3), dy.float().permute(0, 4, 1, 2, 3), padding=1)
    g = g.permute(0, 2, 3, 4, 1).reshape(64, 27, 3).contiguous()
    if out is None:
        return g
    if accumulate:
        out += g
    else:
        out.copy_(g)
    return out


def conv_sd2_supported(shape, Cin, Cout, k=(3, 3, 3)):
    N, D, H, W = [int(v) for v in shape[:4]]
    return Cin == 64 and Cout == 64 and tuple(k) == (3, 3, 3) and D % 2 == 0 and D >= 2 and H >= 16 and W >= 8


def conv_fprop_sd2(x, w, bias=None, relu=False):
    y = F.conv3d(x.float().permute(0, 4, 1, 2, 3), _w5(w, (3, 3, 3)), bias, stride=(2, 1, 1),
                 padding=1).permute(0, 2, 3, 4, 1)
    if relu:
        y = torch.relu(y)
    return y.contiguous().to(STORE)


def conv_dgrad_sd2(dy, wT, relu_ref=None):
    # scatter dy onto the even planes of a zero tensor, then the stride-1 adjoint
    N, Dj, H, W, C = dy.shape
    full = torch.zeros((N, 2 * Dj, H, W, C), dtype=dy.dtype)
    full[:, ::2] = dy
    return conv_dgrad(full, wT, (3, 3, 3), relu_ref=relu_ref)


def conv_wgrad_sd2(dy, x, out=None, accumulate=False):
    N, Dj, H, W, C = dy.shape
    full = torch.zeros((N, 2 * Dj, H, W, C), dtype=dy.dtype)
    full[:, ::2] = dy
    return conv_wgrad(full, x, (3, 3, 3), out=out, accumulate=accumulate)


def conv_fprop_skip(x, w, bias, x2, w2, k=(3, 3, 3), relu=False):
    y = conv_fprop(x, w, bias, None, k, False, True) + conv_fprop(x2, w2.reshape(w2.shape[0], 1, -1), None, None,
                                                                  (1, 1, 1), False, True)
    if relu:
        y = torch.relu(y)
    return y.to(STORE)


def conv_dgrad(dy, wT, k=(3, 3, 3), residual=None, relu=False, out_f32=False, algo=0, relu_ref=None):
    # wT (Cin,taps,Cout) with reversed taps == the forward weight of the adjoint convolution
    if relu_ref is None:
        return conv_fprop(dy, wT, None, residual, k, relu, out_f32)
    dx = conv_fprop(dy, wT, None, None, k, False, True) * (relu_ref.float() > 0).float()
    return dx if out_f32 else dx.to(STORE)


def conv_wgrad(dy, x, k=(3, 3, 3), out=None, accumulate=False, algo=0):
    Cout, Cin = dy.shape[-1], x.shape[-1]
    taps = k[0] * k[1] * k[2]
    xin = x.float().permute(0, 4, 1, 2, 3)
    g = torch.nn.grad.conv3d_weight(xin, (Cout, Cin) + tuple(k), dy.float().permute(0, 4, 1, 2, 3), padding=_pad(k))
    g = g.permute(0, 2, 3, 4, 1).reshape(Cout, taps, Cin).contiguous()
    if out is None:
        return g
    if accumulate:
        out += g
    else:
        out.copy_(g)
    return out


def cast_bf16(src):
    return src.to(STORE)


def cast_f32(src):
    return src.float()


def pack_weight(w, CoutP=None, CinP=None):
    Cout, taps, Cin = w.shape
    CoutP, CinP = CoutP or Cout, CinP or Cin
    dst = torch.zeros((CoutP, taps, CinP), dtype=STORE, device=w.device)
    dst[:Cout, :, :Cin] = w.to(STORE)
    return dst


def pack_dgrad_weight(w, CoutP=None, CinP=None):
    Cout, taps, Cin = w.shape
    CoutP, CinP = CoutP or Cout, CinP or Cin
    wT = torch.zeros((CinP, taps, CoutP), dtype=STORE, device=w.device)
    wT[:Cin, :, :Cout] = w.to(STORE).flip(1).permute(2, 1, 0)
    return wT


def unpack_wgrad(dwp, Cout, Cin):
    return dwp[:Cout, :, :Cin].contiguous()


def relu_fwd(x):
    return torch.relu(x)


def relu_bwd(dy, ref):
    return torch.where(ref > 0, dy, torch.zeros_like(dy))


def pool_out_shape(in_shape, kernel, stride, pad):
    N, D, H, W, C = in_shape
    o = [(s + 2 * p - k) // st + 1 for s, k, st, p in zip((D, H, W), kernel, stride, pad)]
    return (N, o[0], o[1], o[2], C)


def avgpool_fwd(x, kernel, stride, pad, residual=None):
    y = F.avg_pool3d(x.float().permute(0, 4, 1, 2, 3), tuple(kernel), tuple(stride), tuple(pad)).permute(0, 2, 3, 4, 1)
    if residual is not None:
        y = y + residual.float()
    return y.contiguous().to(STORE)


def avgpool_bwd(dy, in_shape, kernel, stride, pad):
    with torch.enable_grad():
        x = torch.zeros((in_shape[0], in_shape[4], in_shape[1], in_shape[2], in_shape[3]), requires_grad=True,
                        device=dy.device)
        y = F.avg_pool3d(x, tuple(kernel), tuple(stride), tuple(pad))
        (g,) = torch.autograd.grad(y, x, dy.detach().float().permute(0, 4, 1, 2, 3))
    return g.permute(0, 2, 3, 4, 1).contiguous().to(STORE)


def _w5g(w, k):
    Cout, taps, Cin = w.shape
    return w.float().view(Cout, k[0], k[1], k[2], Cin).permute(0, 4, 1, 2, 3)


def gconv_fprop(x, w, bias, k, s, p, out_f32=False, cin_real=None):
    y = F.conv3d(x.float().permute(0, 4, 1, 2, 3), _w5g(w, k), bias, stride=tuple(s), padding=tuple(p))
    y = y.permute(0, 2, 3, 4, 1).contiguous()
    return y if out_f32 else y.to(STORE)


def gconv_dgrad(dy, w, bias, in_sp, k, s, p, out_f32=False, cin_real=None):
    osp = [(o - 1) * ss - 2 * pp + kk for o, ss, pp, kk in zip(dy.shape[1:4], s, p, k)]
    opad = [i - o for i, o in zip(in_sp, osp)]
    dx = F.conv_transpose3d(dy.float().permute(0, 4, 1, 2, 3), _w5g(w, k), bias, stride=tuple(s), padding=tuple(p),
                            output_padding=tuple(opad))
    dx = dx.permute(0, 2, 3, 4, 1).contiguous()
    return dx if out_f32 else dx.to(STORE)


def gconv_wgrad(dy, x, k, s, p, cin_real=None):
    Cout, Cin = dy.shape[-1], x.shape[-1]
    g = torch.nn.grad.conv3d_weight(x.float().permute(0, 4, 1, 2, 3), (Cout, Cin) + tuple(k),
                                    dy.float().permute(0, 4, 1, 2, 3), stride=tuple(s), padding=tuple(p))
    return g.permute(0, 2, 3, 4, 1).reshape(Cout, k[0] * k[1] * k[2], Cin).contiguous()


def leaky_relu_fwd(x, slope):
    xf = x.float()
    return torch.where(xf > 0, xf, xf * slope).to(STORE)


def leaky_relu_bwd(dy, ref, slope):
    g = dy.float()
    return torch.where(ref.float() > 0, g, g * slope).to(STORE)


def tanh_fwd(x):
    return torch.tanh(x.float()).to(STORE)


def tanh_bwd(dy, y):
    yf = y.float()
    return (dy.float() * (1 - yf * yf)).to(STORE)


def upsample2x_fwd(x):
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3).contiguous()


def upsample2x_bwd(dy):
    N, D, H2, W2, C = dy.shape
    return dy.float().view(N, D, H2 // 2, 2, W2 // 2, 2, C).sum(dim=(3, 5)).to(STORE)


def nchw_to_cl(x, Cp):
    N, C, D, H, W = x.shape
    y = torch.zeros((N, D, H, W, Cp), dtype=STORE, device=x.device)
    y[..., :C] = x.permute(0, 2, 3, 4, 1).to(STORE)
    return y


def rgb_to_cl(x, want4=True):
    return nchw_to_cl(x, 16), (nchw_to_cl(x, 4) if want4 else None)


def cl_to_nchw(x, C):
    return x[..., :C].permute(0, 4, 1, 2, 3).float().contiguous()


def im2col3(x, Kp):
    N, C, D, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    col = torch.zeros((N, D, H, W, Kp), dtype=F32, device=x.device)
    for tap in range(27):
        a, b, c = tap // 9, (tap // 3) % 3, tap % 3
        col[..., tap * C:(tap + 1) * C] = xp[:, :, a:a + D, b:b + H, c:c + W].permute(0, 2, 3, 4, 1)
    return col.to(STORE)


def col2im3(dcol, C):
    N, D, H, W, Kp = dcol.shape
    dxp = torch.zeros((N, C, D + 2, H + 2, W + 2), dtype=F32, device=dcol.device)
    g = dcol.float()
    for tap in range(27):
        a, b, c = tap // 9, (tap // 3) % 3, tap % 3
        dxp[:, :, a:a + D, b:b + H, c:c + W] += g[..., tap * C:(tap + 1) * C].permute(0, 4, 1, 2, 3)
    return dxp[:, :, 1:-1, 1:-1, 1:-1].contiguous()


def sum_rows(x, out=None):
    s = x.float().reshape(-1, x.shape[-1]).sum(0)
    if out is None:
        return s
    out += s
    return out


def sum_spatial(x):
    return x.float().reshape(x.shape[0], -1, x.shape[-1]).sum(1)


def broadcast_spatial(g, shape):
    N, C = g.shape
    return g.view(N, 1, 1, 1, C).expand(tuple(shape)).to(STORE).contiguous()


def bn_forward(x, gamma, beta, running_mean, running_var, relu, up, eps=1e-5, momentum=0.1, training=True):
    N, D, H, W, C = x.shape
    xf = x.float().reshape(-1, C)
    if training:
        P = xf.shape[0]
        mean = xf.mean(0)
        var = (xf * xf).mean(0) - mean * mean
        var = var.clamp_min(0)
        if running_mean is not None:
            unbiased = var * P / (P - 1) if P > 1 else var
            running_mean.mul_(1 - momentum).add_(momentum * mean)
            running_var.mul_(1 - momentum).add_(momentum * unbiased)
    else:
        mean, var = running_mean.float(), running_var.float()
    invstd = torch.rsqrt(var + eps)
    scale = gamma * invstd
    shift = beta - mean * scale
    y = x.float() * scale + shift
    if relu:                                    # activation code: 1 = ReLU, 2 = LeakyReLU(0.2)
        y = torch.where(y > 0, y, y * (0.0 if relu == 1 else 0.2))
    y = y.to(STORE)
    if up == 2:
        y = upsample2x_fwd(y)
    return y.contiguous(), torch.cat((mean, invstd)), torch.cat((scale, shift))


def bn_backward(dy, x, mean_invstd, scale_shift, relu, up):
    N, D, H, W, C = x.shape
    mean, invstd = mean_invstd[:C], mean_invstd[C:]
    scale, shift = scale_shift[:C], scale_shift[C:]
    g = dy.float()
    if up == 2:
        g = g.view(N, D, H, 2, W, 2, C).sum(dim=(3, 5))
    xf = x.float()
    if relu:
        g = torch.where(xf * scale + shift > 0, g, g * (0.0 if relu == 1 else 0.2))
    xhat = (xf - mean) * invstd
    P = N * D * H * W
    dbeta = g.reshape(-1, C).sum(0)
    dgamma = (g * xhat).reshape(-1, C).sum(0)
    dx = scale * (g - dbeta / P - xhat * dgamma / P)
    return dx.to(STORE), dgamma, dbeta


def _attn_parts(theta, phi, g, c8, c2):
    N, D, H, W, _ = theta.shape
    th = theta.float()[..., :c8].reshape(N, D * H * W, c8)

    def mp(t, c):
        tt = t.float()[..., :c].permute(0, 4, 1, 2, 3)
        return F.max_pool3d(tt, [1, 2, 2]).permute(0, 2, 3, 4, 1).reshape(N, -1, c)
    return th, mp(phi, c8), mp(g, c2)


def attention_fused_ok(D, H, W, c8, c2):
    if c8 > 8 or c2 > 16:
        return False
    if D * H * W <= 1024:
        return True
    return c8 == 4 and c2 == 16 and D * (H // 2) * (W // 2) <= 2560


def attention_fwd(theta, phi, g, c8, c2):
    N, D, H, W, _ = theta.shape
    th, ph, gp = _attn_parts(theta, phi, g, c8, c2)
    beta = torch.softmax(torch.bmm(th, ph.transpose(1, 2)), -1)
    o = torch.zeros((N, D, H, W, g.shape[-1]), dtype=F32, device=theta.device)
    o[..., :c2] = torch.bmm(beta, gp).reshape(N, D, H, W, c2)
    return o.to(STORE)


def attention_bwd(theta, phi, g, dout, c8, c2):
    th_, ph_, g_ = (t.detach().float().requires_grad_(True) for t in (theta, phi, g))
    with torch.enable_grad():
        N, D, H, W, _ = theta.shape
        th, ph, gp = _attn_parts(th_, ph_, g_, c8, c2)
        beta = torch.softmax(torch.bmm(th, ph.transpose(1, 2)), -1)
        o = torch.bmm(beta, gp).reshape(N, D, H, W, c2)
        grads = torch.autograd.grad(o, (th_, ph_, g_), dout.float()[..., :c2])
    return tuple(x.to(STORE) for x in grads)


def render_fwd(pre, B, T, C):
    BT, D, H, W, Cp = pre.shape
    return torch.tanh(pre.float()[..., :C]).view(B, T, H, W, C).permute(0, 4, 1, 2, 3).contiguous()


def render_bwd(dy, y, Cp):
    B, C, T, H, W = y.shape
    g = (dy * (1 - y * y)).permute(0, 2, 3, 4, 1).reshape(B * T, 1, H, W, C)
    out = torch.zeros((B * T, 1, H, W, Cp), dtype=STORE, device=y.device)
    out[..., :C] = g.to(STORE)
    return out


def gather_frames(x, B, T, bt, sn=2, st=2):
    bt = int(bt)
    x5 = x.view((B, T) + tuple(x.shape[1:]))
    y = x5[::sn, bt::st]
    return y.reshape((-1,) + tuple(x.shape[1:])).contiguous()


def scatter_frames(dy, B, T, bt, sn=2, st=2):
    bt = int(bt)
    dx = torch.zeros((B, T) + tuple(dy.shape[1:]), dtype=dy.dtype, device=dy.device)
    Bo = (B + sn - 1) // sn
    dx[::sn, bt::st] = dy.view((Bo, -1) + tuple(dy.shape[1:]))
    return dx.view((B * T,) + tuple(dy.shape[1:]))


def pyramid_level(x, Ho, Wo, sn=1, st=1, bt=0):
    bt = int(bt)
    B, C, T, H, W = x.shape
    ih = (torch.arange(Ho, device=x.device) * H) // Ho
    iw = (torch.arange(Wo, device=x.device) * W) // Wo
    return x[::sn, :, bt::st][:, :, :, ih][:, :, :, :, iw].contiguous()


def lstm_cell_fwd(gates, c_prev, want_h32=False):
    gi, gf, gg, go = gates.chunk(4, dim=-1)
    cp = c_prev if c_prev is not None else torch.zeros_like(gi)
    c = torch.sigmoid(gf) * cp + torch.sigmoid(gi) * torch.tanh(gg)
    h = torch.sigmoid(go) * torch.tanh(c)
    return c.contiguous(), h.to(STORE).contiguous(), (h.contiguous() if want_h32 else None)


def lstm_cell_bwd(gates, c_prev, c, dh, dc_next):
    gi, gf, gg, go = [t for t in gates.chunk(4, dim=-1)]
    i, f, g, o = torch.sigmoid(gi), torch.sigmoid(gf), torch.tanh(gg), torch.sigmoid(go)
    cp = c_prev if c_prev is not None else torch.zeros_like(c)
    tc = torch.tanh(c)
    dhv = dh if dh is not None else torch.zeros_like(c)
    dc = (dc_next if dc_next is not None else torch.zeros_like(c)) + dhv * o * (1 - tc * tc)
    dg = torch.cat((dc * g * i * (1 - i), dc * cp * f * (1 - f), dc * i * (1 - g * g), dhv * tc * o * (1 - o)), dim=-1)
    return dg.to(STORE).contiguous(), (dc * f).contiguous()


def adam_step(params, grads, ms, vs, lr, beta1, beta2, eps, step, grad_scale=1.0, dyn=None):
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    with torch.no_grad():
        for p, g, m, v in zip(params, grads, ms, vs):
            g = g * grad_scale
            m.mul_(beta1).add_(g, alpha=1 - beta1)
            v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            p.addcdiv_(m, v.sqrt() / (bc2 ** 0.5) + eps, value=-lr / bc1)


def multi_copy(srcs, dsts):
    with torch.no_grad():
        for s, d in zip(srcs, dsts):
            d.view(-1)[:] = torch.as_strided(s, (s.numel(),), (1,), s.storage_offset()) if not s.is_contiguous() \
                else s.reshape(-1)


# ------------------------------------------------------------------------------------- round-2 kernels
def gconv_pack(w3):
    return w3.contiguous().to(STORE)


def scale(x, s):
    return (x.float() * s.float()).to(x.dtype)


def scale_add(o, x, s=None):
    return (o.float() * (1.0 if s is None else s.float()) + x.float()).to(x.dtype)


def dot(a, b):
    return (a.float() * b.float()).sum()


def cl_slice_f32(x, c):
    return x[..., :c].float().contiguous()


def f32_pad_cl(x, Cp, dtype=None):
    y = torch.zeros(tuple(x.shape[:-1]) + (Cp,), dtype=dtype or STORE, device=x.device)
    y[..., :x.shape[-1]] = x.to(y.dtype)
    return y


def maxpool122_fwd(x):
    M, H, W, c = x.shape
    win = x.view(M, H // 2, 2, W // 2, 2, c).permute(0, 1, 3, 5, 2, 4).reshape(M, H // 2, W // 2, c, 4)
    y, idx = win.max(dim=-1)
    # first maximum, as the kernel: torch.max returns an arbitrary one on ties; recompute deterministically
    first = (win == y.unsqueeze(-1)).float().argmax(dim=-1)
    return y.contiguous(), first.to(torch.uint8).contiguous()


def pool122_gather(x, idx):
    M, H, W, c = x.shape
    win = x.view(M, H // 2, 2, W // 2, 2, c).permute(0, 1, 3, 5, 2, 4).reshape(M, H // 2, W // 2, c, 4)
    return win.gather(-1, idx.long().unsqueeze(-1)).squeeze(-1).contiguous()


def pool122_scatter(dy, idx):
    M, Hp, Wp, c = dy.shape
    win = torch.zeros((M, Hp, Wp, c, 4), dtype=dy.dtype, device=dy.device)
    win.scatter_(-1, idx.long().unsqueeze(-1), dy.unsqueeze(-1))
    return win.view(M, Hp, Wp, c, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(M, 2 * Hp, 2 * Wp, c).contiguous()


def bmm(a, b, ta=False, tb=False):
    return torch.bmm(a.transpose(1, 2) if ta else a, b.transpose(1, 2) if tb else b).contiguous()


def softmax_fwd(s):
    return torch.softmax(s, -1)


def softmax_bwd(beta, dbeta):
    return beta * (dbeta - (beta * dbeta).sum(-1, keepdim=True))


def softmax_bwd_bwd(beta, dbeta, u):
    s = (beta * dbeta).sum(-1, keepdim=True)
    t = (beta * u).sum(-1, keepdim=True)
    return u * (dbeta - s) - dbeta * t, beta * (u - t)


def head_fwd(feat, cond, w, bias):
    x = feat if cond is None else torch.cat((feat, cond), dim=1)
    out = x @ w
    return out + bias.reshape(()) if bias is not None else out


def head_bwd_data(dpred, w, F_, E):
    g = dpred.reshape(-1, 1) * w.reshape(1, -1)
    return g[:, :F_].contiguous(), (g[:, F_:F_ + E].contiguous() if E else None)


def head_bwd_weight(dpred, feat, cond, want_bias=True):
    x = feat if cond is None else torch.cat((feat, cond), dim=1)
    return dpred.reshape(1, -1) @ x, (dpred.sum().reshape(1) if want_bias else None)


def rel_loss_fwd(a_list, b_list, weights, mode):
    tot = 0
    for a, b, w in zip(a_list, b_list, weights):
        d = b - a
        tot = tot + w * (F.softplus(d) if mode == 0 else d).mean()
    return tot.reshape(())


def rel_loss_bwd(a_list, b_list, da_list, db_list, weights, mode, gout):
    for a, b, da, db, w in zip(a_list, b_list, da_list, db_list, weights):
        d = b - a
        fp = (torch.sigmoid(d) if mode == 0 else torch.ones_like(d)) * (gout * w / a.numel())
        if da is not None:
            da -= fp
        if db is not None:
            db += fp


def lerp_rows(real, fake, alpha):
    a = alpha.view([-1] + [1] * (real.dim() - 1))
    return a * real + (1 - a) * fake


def lstm_pack_whh(whh):
    return whh.transpose(1, 2).contiguous()


def lstm_seq_fwd(gx, whhT, lengths, h0, c0, save=True):
    B, L = gx.shape[0], gx.shape[1]
    ndir, H = whhT.shape[0], whhT.shape[1]
    g4 = gx.view(B, L, ndir, 4 * H)
    out = torch.zeros((B, L, ndir * H))
    hprev = torch.zeros((B, L, ndir * H))
    gates = torch.zeros((B, L, ndir, 4 * H))
    cells = torch.zeros((B, L, ndir, H))
    hn, cn = torch.zeros((ndir, B, H)), torch.zeros((ndir, B, H))
    lens = lengths.long()
    for d in range(ndir):
        h = h0[d].clone() if h0 is not None else torch.zeros(B, H)
        c = c0[d].clone() if c0 is not None else torch.zeros(B, H)
        for s in range(L):
            t = s if d == 0 else L - 1 - s
            live = (lens > t).unsqueeze(1)
            pre = g4[:, t, d] + h @ whhT[d]
            gi, gf, gg, go = pre.chunk(4, dim=1)
            gi, gf, gg, go = torch.sigmoid(gi), torch.sigmoid(gf), torch.tanh(gg), torch.sigmoid(go)
            c_new = gf * c + gi * gg
            h_new = go * torch.tanh(c_new)
            hprev[:, t, d * H:(d + 1) * H] = torch.where(live, h, torch.zeros_like(h))
            gates[:, t, d] = torch.where(live, torch.cat((gi, gf, gg, go), 1), torch.zeros(B, 4 * H))
            cells[:, t, d] = torch.where(live, c_new, torch.zeros_like(c))
            out[:, t, d * H:(d + 1) * H] = torch.where(live, h_new, torch.zeros_like(h))
            h = torch.where(live, h_new, h)
            c = torch.where(live, c_new, c)
        hn[d], cn[d] = h, c
    if not save:
        return out.to(STORE), None, None, None, hn, cn
    return out.to(STORE), hprev.to(STORE), gates.view(B, L, ndir * 4 * H), cells.view(B, L, ndir * H), hn, cn


def lstm_seq_bwd(whh, lengths, c0, gates, cells, dout, dhn, dcn):
    ndir, H = whh.shape[0], whh.shape[2]
    B, L = gates.shape[0], gates.shape[1]
    g4, c4 = gates.view(B, L, ndir, 4 * H), cells.view(B, L, ndir, H)
    dg = torch.zeros((B, L, ndir, 4 * H))
    dh0, dc0 = torch.zeros((ndir, B, H)), torch.zeros((ndir, B, H))
    lens = lengths.long()
    for d in range(ndir):
        dh = dhn[d].clone() if dhn is not None else torch.zeros(B, H)
        dc = dcn[d].clone() if dcn is not None else torch.zeros(B, H)
        for s in range(L - 1, -1, -1):
            t = s if d == 0 else L - 1 - s
            live = (lens > t).unsqueeze(1)
            gi, gf, gg, go = g4[:, t, d].chunk(4, dim=1)
            tp = t - 1 if d == 0 else t + 1
            cp0 = c0[d] if c0 is not None else torch.zeros(B, H)
            if 0 <= tp < L:
                cp = torch.where((lens > tp).unsqueeze(1), c4[:, tp, d], cp0)
            else:
                cp = cp0
            tc = torch.tanh(c4[:, t, d])
            dhv = dh + (dout[:, t, d * H:(d + 1) * H].float() if dout is not None else 0)
            dcv = dc + dhv * go * (1 - tc * tc)
            step = torch.cat((dcv * gg * gi * (1 - gi), dcv * cp * gf * (1 - gf), dcv * gi * (1 - gg * gg),
                              dhv * tc * go * (1 - go)), 1)
            step = torch.where(live, step, torch.zeros_like(step))
            dg[:, t, d] = step
            dh = torch.where(live, step @ whh[d], dh)
            dc = torch.where(live, dcv * gf, dc)
        dh0[d], dc0[d] = dh, dc
    return dg.view(B, L, ndir * 4 * H).to(STORE), dh0, dc0


def embedding_fwd(tokens, weight):
    return weight[tokens].to(STORE)


def embedding_bwd(tokens, dout, V):
    dw = torch.zeros((V, dout.shape[-1]))
    dw.index_add_(0, tokens.reshape(-1), dout.float().reshape(-1, dout.shape[-1]))
    return dw


# ---- executable spec of csrc/s2d.cu (stride-2 layers as dense kernel-2 convolutions over shifted blocks) -----------
def _s2d_phases(modes):
    return 2 ** sum(1 for m in modes if m == 2)


def s2d_shift(x, modes, creal=None):
    """(N,D,H,W,C) -> (N,D',H',W',round16(P*creal)): block b of a strided axis = samples (2b-1, 2b), zero outside"""
    N, D, H, W, C = x.shape
    creal = C if creal is None else creal
    xr = x[..., :creal]
    pads = []
    for m in reversed(modes):
        pads += [1, 1] if m == 2 else [0, 0]
    xr = F.pad(xr, [0, 0] + pads)
    f = [2 if m == 2 else 1 for m in modes]
    Db, Hb, Wb = xr.shape[1] // f[0], xr.shape[2] // f[1], xr.shape[3] // f[2]
    xr = xr.reshape(N, Db, f[0], Hb, f[1], Wb, f[2], creal).permute(0, 1, 3, 5, 2, 4, 6, 7)
    xr = xr.reshape(N, Db, Hb, Wb, f[0] * f[1] * f[2] * creal)
    Cp = (xr.shape[-1] + 15) // 16 * 16
    return F.pad(xr, [0, Cp - xr.shape[-1]]).contiguous()


def d2s_shift(xs, modes, sp, C, creal=None):
    """inverse of s2d_shift (channels >= creal zero)"""
    N = xs.shape[0]
    creal = C if creal is None else creal
    f = [2 if m == 2 else 1 for m in modes]
    Db, Hb, Wb = xs.shape[1:4]
    t = xs[..., :f[0] * f[1] * f[2] * creal].reshape(N, Db, Hb, Wb, f[0], f[1], f[2], creal)
    t = t.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(N, Db * f[0], Hb * f[1], Wb * f[2], creal)
    sl = [slice(1, 1 + e) if m == 2 else slice(0, e) for m, e in zip(modes, sp)]
    t = t[:, sl[0], sl[1], sl[2]]
    return F.pad(t, [0, C - creal]).contiguous()


def s2d_embed_weight(w, modes, creal=None, transposed=False):
    """w (Co, k taps, Ci), k = 4 per strided axis -> (Co, 3^s, Cp): engine tap t = block offset j + 1, channel =
    phase * creal + c with kernel index 2 j + phase per axis; transposed: (Cp, 3^s reversed, Co)"""
    Co, taps, Ci = w.shape
    creal = Ci if creal is None else creal
    kk = [4 if m == 2 else 1 for m in modes]
    ke = [3 if m == 2 else 1 for m in modes]
    P = _s2d_phases(modes)
    Cp = (P * creal + 15) // 16 * 16
    w6 = w.reshape(Co, kk[0], kk[1], kk[2], Ci)
    we = torch.zeros((Co, ke[0], ke[1], ke[2], Cp), dtype=w.dtype, device=w.device)
    rng = [(0, 1) if m == 2 else (0,) for m in modes]
    for jd in rng[0]:
        for jh in rng[1]:
            for jw in rng[2]:
                for pd in rng[0]:
                    for ph in rng[1]:
                        for pw in rng[2]:
                            phase = 0
                            for m, pp in zip(modes, (pd, ph, pw)):
                                if m == 2:
                                    phase = phase * 2 + pp
                            td, th, tw = [j + 1 if m == 2 else 0 for m, j in zip(modes, (jd, jh, jw))]
                            ad, ah, aw = [2 * j + pp if m == 2 else 0 for m, j, pp in zip(modes, (jd, jh, jw), (pd, ph, pw))]
                            we[:, td, th, tw, phase * creal:(phase + 1) * creal] = w6[:, ad, ah, aw, :creal]
    we = we.reshape(Co, ke[0] * ke[1] * ke[2], Cp)
    if transposed:
        we = we.flip(1).permute(2, 1, 0).contiguous()
    return we


def s2d_extract_wgrad(dwe, modes, Ci, creal=None):
    """(Co, 3^s, Cp) -> (Co, k taps, Ci): the adjoint gather of s2d_embed_weight"""
    Co, ntap, Cp = dwe.shape
    creal = Ci if creal is None else creal
    kk = [4 if m == 2 else 1 for m in modes]
    ke = [3 if m == 2 else 1 for m in modes]
    d5 = dwe.reshape(Co, ke[0], ke[1], ke[2], Cp)
    dw = torch.zeros((Co, kk[0], kk[1], kk[2], Ci), dtype=dwe.dtype, device=dwe.device)
    for ad in range(kk[0]):
        for ah in range(kk[1]):
            for aw in range(kk[2]):
                phase, t = 0, []
                for m, a in zip(modes, (ad, ah, aw)):
                    if m == 2:
                        phase = phase * 2 + (a & 1)
                        t.append(a // 2 + 1)
                    else:
                        t.append(0)
                dw[:, ad, ah, aw, :creal] = d5[:, t[0], t[1], t[2], phase * creal:(phase + 1) * creal]
    return dw.reshape(Co, kk[0] * kk[1] * kk[2], Ci)
