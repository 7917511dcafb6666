"""CPU oracle for txt2vid's GAN training step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, fp32, plain-PyTorch restatement of the reference's TGANv2 (conditional and
unconditional) generator / discriminator / caption encoder / losses / gradient penalty / training
iteration.  It works on state_dicts that use the reference's parameter names, so the same weights can
be fed to the reference (in the build container), to this oracle, and to the sm_100a product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this file.  The product (txt2vid_b200/) never does.

Parity pin: the reference has no tests or golden vectors of its own, so this
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports the live
reference from /root/reference, runs the same seeds and writes tests/golden/*.json; tests/test_oracle_*.py
check this file against those fixtures.

Every function cites the reference file:line it restates (paths relative to the reference tree).
Host-RNG draws (frame offsets `bt`, caption permutation, GP alphas) are explicit inputs; `draw_*`
helpers below reproduce the reference's draw order.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------ indices


def subsample(x, bt, sn=2, st=2):
    """txt2vid/models/layers.py:106-111 -- x[::sn, :, bt::st] on a (B,C,T,H,W) tensor."""
    return x[::sn, :, int(bt)::st]


def draw_bt(st=2):
    """layers.py:107-108 -- one CPU-generator draw."""
    return int(torch.randint(st, (1,)))


def nearest_resize(x, size):
    """gan/trainer.py:149 -- F.interpolate(x, size=(T, fs, fs)) (mode 'nearest'): src = floor(dst*in/out)."""
    B, C, T, H, W = x.shape
    t, h, w = size
    it = (torch.arange(t, device=x.device) * T) // t
    ih = (torch.arange(h, device=x.device) * H) // h
    iw = (torch.arange(w, device=x.device) * W) // w
    return x[:, :, it][:, :, :, ih][:, :, :, :, iw]


def multiscale_data(x, cond, frame_sizes, subsample_input, bts):
    """gan/trainer.py:131-165.  bts: one frame offset per level (the last one is drawn and unused)."""
    n = len(frame_sizes)
    if n == 1:
        return [x], (None if cond is None else [cond])
    xs, conds = [], []
    for i in range(n):
        T = x.size(2)
        xs.append(nearest_resize(x, (T, frame_sizes[i], frame_sizes[i])) if i != n - 1 else x)
        if cond is not None:
            conds.append(cond)
        if subsample_input:
            x = subsample(x, bts[i])
            if cond is not None:
                cond = cond[::2]
    return xs, (conds if conds else None)


def gen_perm(n):
    """util/misc.py:3-8 -- numpy permutation re-drawn until it is not the identity."""
    old = np.array(range(n))
    new = np.random.permutation(old)
    while (new == old).all():
        new = np.random.permutation(old)
    return new


# ------------------------------------------------------------------------------------------ blocks


def _conv(x, sd, name, pad):
    w = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    if w.dim() == 4:
        return F.conv2d(x, w, b, padding=pad)
    return F.conv3d(x, w, b, padding=pad)


def _bn_train(x, sd, name, new_buffers):
    """nn.BatchNorm2d in train(): batch statistics, eps 1e-5, momentum 0.1."""
    rm = sd[name + ".running_mean"].clone()
    rv = sd[name + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, sd[name + ".weight"], sd[name + ".bias"], True, 0.1, 1e-5)
    if new_buffers is not None:
        new_buffers[name + ".running_mean"] = rm
        new_buffers[name + ".running_var"] = rv
    return y


def _bn_eval(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], False, 0.1, 1e-5)


def _bn(x, sd, name, training, new_buffers):
    return _bn_train(x, sd, name, new_buffers) if training else _bn_eval(x, sd, name)


def attention2d(x, sd, p):
    """models/layers.py:23-36 (SA-GAN non-local block, 2-D)."""
    ch = x.shape[1]
    theta = F.conv2d(x, sd[p + ".theta.weight"])
    phi = F.max_pool2d(F.conv2d(x, sd[p + ".phi.weight"]), [2, 2])
    g = F.max_pool2d(F.conv2d(x, sd[p + ".g.weight"]), [2, 2])
    hw = x.shape[2] * x.shape[3]
    theta = theta.view(-1, ch // 8, hw)
    phi = phi.view(-1, ch // 8, hw // 4)
    g = g.view(-1, ch // 2, hw // 4)
    beta = F.softmax(torch.bmm(theta.transpose(1, 2), phi), -1)
    o = torch.bmm(g, beta.transpose(1, 2)).view(-1, ch // 2, x.shape[2], x.shape[3])
    o = F.conv2d(o, sd[p + ".o.weight"])
    return sd[p + ".gamma"] * o + x


def attention3d(x, sd, p):
    """models/layers.py:52-68 (non-local block, 3-D, max-pool (1,2,2))."""
    B, ch = x.shape[0], x.shape[1]
    theta = F.conv3d(x, sd[p + ".theta.weight"])
    phi = F.max_pool3d(F.conv3d(x, sd[p + ".phi.weight"]), [1, 2, 2])
    g = F.max_pool3d(F.conv3d(x, sd[p + ".g.weight"]), [1, 2, 2])
    theta = theta.view(B, ch // 8, -1)
    phi = phi.view(B, ch // 8, -1)
    g = g.view(B, ch // 2, -1)
    beta = F.softmax(torch.bmm(theta.transpose(1, 2), phi), -1)
    o = torch.bmm(g, beta.transpose(1, 2)).view(B, -1, x.shape[2], x.shape[3], x.shape[4])
    o = F.conv3d(o, sd[p + ".o.weight"])
    return sd[p + ".gamma"] * o + x


def up_block(x, sd, p, training, new_buffers):
    """models/layers.py:152-195: BN-ReLU-Up2-conv3-BN-ReLU-conv3 + (Up2 [-conv1]) skip [+ attention]."""
    m = p + ".main.inner_module"
    h = _bn(x, sd, m + ".0", training, new_buffers)
    h = F.relu(h)
    h = F.interpolate(h, scale_factor=2)
    h = _conv(h, sd, m + ".3", 1)
    h = _bn(h, sd, m + ".4", training, new_buffers)
    h = F.relu(h)
    h = _conv(h, sd, m + ".6", 1)
    s = F.interpolate(x, scale_factor=2)
    if (p + ".main.identity_map.1.weight") in sd:
        s = _conv(s, sd, p + ".main.identity_map.1", 0)
    out = s + h
    if (p + ".attn.gamma") in sd:
        out = attention2d(out, sd, p + ".attn")
    return out


def render_block(x, sd, p, training, new_buffers):
    """models/layers.py:245-259: BN-ReLU-conv3(C->3)-tanh."""
    h = F.relu(_bn(x, sd, p + ".bn", training, new_buffers))
    return torch.tanh(_conv(h, sd, p + ".conv", 1))


def conv_lstm(x, sd, p, steps):
    """models/conv_lstm.py:32-38,75-97: one cell, peepholes identically zero (:46-49), input is zero
    after step 0 (:79) -- so Wx*(0) contributes only its bias."""
    B, _, fh, fw = x.shape
    hid = sd[p + ".Whi.weight"].shape[0]
    h = torch.zeros(B, hid, fh, fw, device=x.device)
    c = torch.zeros(B, hid, fh, fw, device=x.device)
    outs = []
    for step in range(steps):
        xi = x if step == 0 else torch.zeros_like(x)
        gi = torch.sigmoid(_conv(xi, sd, p + ".Wxi", 1) + _conv(h, sd, p + ".Whi", 1))
        gf = torch.sigmoid(_conv(xi, sd, p + ".Wxf", 1) + _conv(h, sd, p + ".Whf", 1))
        c = gf * c + gi * torch.tanh(_conv(xi, sd, p + ".Wxc", 1) + _conv(h, sd, p + ".Whc", 1))
        go = torch.sigmoid(_conv(xi, sd, p + ".Wxo", 1) + _conv(h, sd, p + ".Who", 1))
        h = go * torch.tanh(c)
        outs.append(h)
    return outs


def gen_forward(sd, z, cond, bts, training=True, num_frames=16, new_buffers=None, abstract=None):
    """models/tganv2_cond/gen.py:64-124 (cond) and models/tganv2/gen.py:62-119 (uncond, cond=None).

    bts: the 3 frame offsets drawn before levels 1..3 (training only).  Returns the list of rendered
    levels (B_i, 3, T_i, H_i, W_i), coarse to fine."""
    x = torch.cat((z, cond), dim=1) if cond is not None else z
    x = F.linear(x, sd["fc.weight"], sd["fc.bias"])
    fm_ch = sd["clstm.cell0.Whi.weight"].shape[0]
    fm = int(round(math.sqrt(x.shape[1] // fm_ch)))
    x = x.view(x.size(0), fm_ch, fm, fm)
    frames = conv_lstm(x, sd, "clstm.cell0", num_frames)
    x = torch.stack(frames).permute(1, 0, 2, 3, 4)          # (B,T,C,h,w)
    T = num_frames
    x = x.contiguous().view(-1, x.size(2), x.size(3), x.size(4))
    n_levels = len([k for k in sd if k.startswith("render_blocks.") and k.endswith(".conv.weight")])
    rendered = []
    for i in range(n_levels):
        if i != 0 and training:
            x5 = x.contiguous().view(-1, T, x.size(1), x.size(2), x.size(3)).permute(0, 2, 1, 3, 4)
            x5 = subsample(x5, bts[i - 1])
            x5 = x5.permute(0, 2, 1, 3, 4)
            x = x5.contiguous().view(-1, x5.size(2), x5.size(3), x5.size(4))
            T //= 2
        ab = "abstract_blocks.%d" % i
        if i == 0:
            for u in ("up0", "up1", "up2"):
                x = up_block(x, sd, ab + "." + u, training, new_buffers)
        else:
            x = up_block(x, sd, ab, training, new_buffers)
        if abstract is not None:                        # debugging aid: the abstract map of every level (B_i*T_i,C,H,W)
            abstract.append(x)
        if i == n_levels - 1 or training:
            r = render_block(x, sd, "render_blocks.%d" % i, training, new_buffers)
            r = r.contiguous().view(-1, T, r.size(1), r.size(2), r.size(3)).permute(0, 2, 1, 3, 4)
            rendered.append(r)
    return rendered


def down_sample(x):
    """models/layers.py:202-217: avg-pool 2 on every dim of extent > 1 (pad 1 when odd,
    count_include_pad)."""
    k, s, pd = [1, 1, 1], [1, 1, 1], [0, 0, 0]
    for i in range(3):
        size = x.size(i + 2)
        if size == 1:
            continue
        k[i] = s[i] = 2
        if size % 2:
            pd[i] = 1
    return F.avg_pool3d(x, kernel_size=k, stride=s, padding=pd)


def down_block(x, sd, p):
    """models/layers.py:219-243: ReLU-conv3^3-ReLU-conv3^3-pool  +  conv1^3-pool."""
    m = p + ".main.inner_module"
    h = _conv(F.relu(x), sd, m + ".1", 1)
    h = _conv(F.relu(h), sd, m + ".3", 1)
    h = down_sample(h)
    s = down_sample(_conv(x, sd, p + ".main.identity_map.0", 0))
    return s + h


def resnet3d(sd, p, x, cond):
    """models/resnet3d.py:38-57.  Returns (uncond, cond_or_None, features)."""
    m = p + "res_block.inner_module"
    h = _conv(x, sd, m + ".0", 1)
    h = _conv(F.relu(h), sd, m + ".2", 1)
    h = F.avg_pool3d(h, (1, 2, 2), 2)                         # resnet3d.py:16 -- stride 2 in ALL dims
    s = _conv(F.avg_pool3d(x, (1, 2, 2), 2), sd, p + "res_block.identity_map.1", 0)
    x = s + h
    i = 0
    while (p + "down.%d.main.inner_module.1.weight" % i) in sd or (p + "down.%d.gamma" % i) in sd:
        if (p + "down.%d.gamma" % i) in sd:
            x = attention3d(x, sd, p + "down.%d" % i)
        else:
            x = down_block(x, sd, p + "down.%d" % i)
        i += 1
    feat = torch.sum(x, [2, 3, 4])
    u = F.linear(feat, sd[p + "fc_uncond.weight"], sd[p + "fc_uncond.bias"])
    if cond is not None:
        c = F.linear(torch.cat((feat, cond), dim=1), sd[p + "fc.weight"], sd[p + "fc.bias"])
        return u, c, feat
    return u, None, feat


def discrim_prefix(sd):
    """cond D wraps the shared trunk in nn.DataParallel -> 'single_discrim.module.' keys
    (models/tganv2_cond/discrim.py:15); the uncond one does not (models/tganv2/discrim.py:16)."""
    return "single_discrim.module." if any(k.startswith("single_discrim.module.") for k in sd) \
        else "single_discrim."


def discrim_forward(sd, xs, conds):
    """models/tganv2_cond/discrim.py:28-48: the shared Resnet3D applied to every level."""
    p = discrim_prefix(sd)
    return [resnet3d(sd, p, x, None if conds is None else conds[i]) for i, x in enumerate(xs)]


# ------------------------------------------------------------------------------------------ text


def seq2seq_encode(sd, tokens, lengths, num_layers=4, p="encoder."):
    """models/txt/basic.py:49-70: Embedding -> packed 4-layer Bi-LSTM -> cat(last-layer final h fwd, bwd).

    Restates cuDNN/ATen packed-sequence semantics explicitly: sample b only advances while t < len_b;
    the backward direction starts at each sample's own last token.  Returns (out, hn)."""
    emb = F.embedding(tokens, sd[p + "embed.weight"])         # (B, L, E)
    B, L = tokens.shape
    lens = torch.as_tensor(lengths, device=tokens.device)
    H = sd[p + "lstm.weight_hh_l0"].shape[1]
    inp = emb
    h_last = None
    for layer in range(num_layers):
        outs = []
        finals = []
        for direction, suffix in ((0, ""), (1, "_reverse")):
            w_ih = sd[p + "lstm.weight_ih_l%d%s" % (layer, suffix)]
            w_hh = sd[p + "lstm.weight_hh_l%d%s" % (layer, suffix)]
            b = sd[p + "lstm.bias_ih_l%d%s" % (layer, suffix)] + sd[p + "lstm.bias_hh_l%d%s" % (layer, suffix)]
            h = torch.zeros(B, H, device=emb.device)
            c = torch.zeros(B, H, device=emb.device)
            out = [None] * L
            order = range(L) if direction == 0 else range(L - 1, -1, -1)
            for t in order:
                gates = F.linear(inp[:, t], w_ih) + F.linear(h, w_hh) + b
                gi, gf, gg, go = gates.chunk(4, dim=1)
                c_new = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
                h_new = torch.sigmoid(go) * torch.tanh(c_new)
                live = (lens > t).unsqueeze(1).to(h.dtype)
                h = live * h_new + (1 - live) * h
                c = live * c_new + (1 - live) * c
                out[t] = live * h_new
            outs.append(torch.stack(out, dim=1))
            finals.append(h)
        inp = torch.cat(outs, dim=2)
        h_last = finals
    hn = torch.cat((h_last[0], h_last[1]), dim=1)
    return inp[:, :int(lengths[0])], hn


# ------------------------------------------------------------------------------------------ losses


def _bce_logits_ones(x):
    """BCEWithLogitsLoss(x, 1) with mean reduction = mean(softplus(-x))."""
    return F.softplus(-x).mean()


class RSGAN:
    """gan/losses.py:74-85."""

    @staticmethod
    def discrim_loss(fake, real):
        return _bce_logits_ones(real - fake)

    @staticmethod
    def gen_loss(fake, real):
        return _bce_logits_ones(fake - real)


class Wasserstein:
    """gan/losses.py:55-68."""

    @staticmethod
    def discrim_loss(fake, real):
        return -(real.mean() - fake.mean())

    @staticmethod
    def gen_loss(fake, real):
        return -fake.mean()


class RaLSGAN:
    """gan/losses.py:113-133."""

    @staticmethod
    def discrim_loss(fake, real):
        return (torch.mean((real - fake.mean() - 1) ** 2) + torch.mean((fake - real.mean() + 1) ** 2)) / 2

    @staticmethod
    def gen_loss(fake, real):
        return (torch.mean((real - fake.mean() + 1) ** 2) + torch.mean((fake - real.mean() - 1) ** 2)) / 2


class Vanilla:
    """gan/losses.py:19-46 -- NOTE the reference swaps the labels in LabelledGanLoss.__init__ (:26-27):
    fake_label <- real_label (1), real_label <- fake_label (0).  Kept as observable behaviour."""

    @staticmethod
    def discrim_loss(fake, real):
        return F.binary_cross_entropy_with_logits(fake, torch.ones_like(fake)) + \
            F.binary_cross_entropy_with_logits(real, torch.zeros_like(real))

    @staticmethod
    def gen_loss(fake, real):
        return F.binary_cross_entropy_with_logits(fake, torch.zeros_like(fake))


def draw_gp_alphas(batch_sizes):
    """gan/losses.py:140-145: torch.rand(B_i,1,1,1,1) on the CPU generator, levels in order."""
    return [torch.rand(b, 1, 1, 1, 1) for b in batch_sizes]


def gradient_penalty(sd, real_x, fake_x, real_cond, fake_cond, alphas):
    """gan/losses.py:135-209 for a multi-scale D: per level, zero-centred, combine = sum over samples,
    summed over levels.  Only d/dx_hat is kept (`grad(...)[0]`, :178)."""
    p = discrim_prefix(sd)
    total = 0
    for i in range(len(real_x)):
        a = alphas[i].to(real_x[i].device).clone().requires_grad_(True)
        ax = a.expand_as(real_x[i])
        xh = ax * real_x[i] + (1 - ax) * fake_x[i]
        ch = None
        if real_cond is not None:
            ac = a.view(-1, 1).expand_as(real_cond[i])
            ch = ac * real_cond[i] + (1 - ac) * fake_cond[i]
        u, c, _ = resnet3d(sd, p, xh, ch)
        outs = [u] + ([c] if c is not None else [])
        ins = [xh] + ([ch] if ch is not None else [])
        g = torch.autograd.grad(outs, ins, [torch.ones_like(o) for o in outs], create_graph=True,
                                retain_graph=True)[0]
        total = total + (g.reshape(g.size(0), -1).norm(2, dim=1) ** 2).sum()
    return total


def discrim_loss(sd, real, fake, cond, fake_cond, alphas, loss=RSGAN, gp_lambda=0.5):
    """gan/cond_gan.py:34-87 (+ :121-164 wrapper).  fake_cond: per-level mismatched captions
    (cond_gan.py:133-134), None for the unconditional model."""
    if cond is not None:
        real_cc = discrim_forward(sd, real, cond)
        real_ic = discrim_forward(sd, real, fake_cond)
        fake_cc = discrim_forward(sd, fake, cond)
        lu = torch.stack([loss.discrim_loss(f[0], r[0]) for f, r in zip(fake_cc, real_cc)]).mean()
        l1 = torch.stack([loss.discrim_loss(f[1], r[1]) for f, r in zip(fake_cc, real_cc)]).mean()
        l2 = torch.stack([loss.discrim_loss(f[1], r[1]) for f, r in zip(real_ic, real_cc)]).mean()
        l = (lu + (l1 + l2) / 2) / 2.0
    else:
        rp = [r[0] for r in discrim_forward(sd, real, None)]
        fp = [f[0] for f in discrim_forward(sd, fake, None)]
        l = torch.stack([loss.discrim_loss(f, r) for f, r in zip(fp, rp)]).mean()
    if gp_lambda > 0:
        l = l + gp_lambda * gradient_penalty(sd, real, fake, cond, fake_cond, alphas)
    return l


def gen_loss(sd_d, fake, real_pred, cond, loss=RSGAN):
    """gan/cond_gan.py:90-118.  real_pred: D(real) tuples from all_discrim_forward(real only)
    (trainer.py:247).  For the unconditional model the reference passes tuples to the loss and raises
    (cond_gan.py:102-106); the documented adapter uses element [0]."""
    fake_cc = discrim_forward(sd_d, fake, cond)
    if cond is None:
        return torch.stack([loss.gen_loss(f[0], r[0]) for f, r in zip(fake_cc, real_pred)]).mean()
    lu = torch.stack([loss.gen_loss(f[0], r[0]) for f, r in zip(fake_cc, real_pred)]).mean()
    lc = torch.stack([loss.gen_loss(f[1], r[1]) for f, r in zip(fake_cc, real_pred)]).mean()
    return (lc + lu) / 2.0


# ------------------------------------------------------------------------------------------ optimiser


class Adam:
    """torch.optim.Adam as configured at train/gan.py:93-94 (eps 1e-8, no weight decay, no amsgrad)."""

    def __init__(self, names, lr, betas):
        self.names, self.lr, self.betas = list(names), lr, betas
        self.t = 0
        self.m, self.v = {}, {}

    def step(self, sd, grads):
        self.t += 1
        b1, b2 = self.betas
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        with torch.no_grad():
            for n in self.names:
                g = grads.get(n)
                if g is None:
                    continue
                m = self.m.setdefault(n, torch.zeros_like(g))
                v = self.v.setdefault(n, torch.zeros_like(g))
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (v.sqrt() / math.sqrt(bc2)).add_(1e-8)
                sd[n].addcdiv_(m, denom, value=-self.lr / bc1)


# ------------------------------------------------------------------------------------------ iteration


def param_names(sd):
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


def as_leaves(sd):
    out = {}
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            out[k] = v
        elif k.endswith("running_mean") or k.endswith("running_var"):
            out[k] = v.detach().clone()
        else:
            out[k] = v.detach().clone().float().requires_grad_(True)
    return out


def train_iteration(sd_g, sd_d, sd_txt, x, tokens, lengths, z, draws, loss=RSGAN, gp_lambda=0.5,
                    frame_sizes=(8, 16, 32, 64), subsample_input=True, opt_g=None, opt_d=None,
                    num_frames=16, reduce=None):
    """One iteration of gan/trainer.py:199-265 (discrim_steps = gen_steps = 1, end2end = False).
    reduce: optional callable {name: grad} -> {name: grad} applied to the D and to the G gradients right before their
    optimiser step (data-parallel oracle, SURVEY 8e: the fp32 average over ranks).

    x: (B,3,T,H,W) fp32 ("channel_first", trainer.py:203-204).  draws: dict with
      'bt_real' (len(frame_sizes) offsets, trainer.py:145-160), 'bt_fake' (3 offsets, gen.py:101-105),
      'perm' (caption permutation for the D step, cond_gan.py:133), 'alphas' (GP, losses.py:140-145).
    sd_*: dicts of leaf tensors (as_leaves).  Returns a dict with losses, gradients, outputs; applies
    Adam in place when optimisers are given."""
    out = {}
    cond0 = None
    if sd_txt is not None:
        _, hn = seq2seq_encode(sd_txt, tokens, lengths)
        cond0 = hn.detach()                                   # trainer.py:214-215 (not end2end)
    xs, conds = multiscale_data(x, cond0, list(frame_sizes), subsample_input, draws["bt_real"])
    new_buf = {}
    fake = gen_forward(sd_g, z, None if conds is None else conds[0], draws["bt_fake"], True, num_frames, new_buf)
    out["fake"] = [f.detach() for f in fake]
    out["real_levels"] = xs
    # ---- D step (cond_gan.py:156-164, trainer.py:230-243)
    fake_cond = None
    if conds is not None:
        fc0 = conds[0][torch.as_tensor(draws["perm"], device=conds[0].device)]
        fake_cond = [fc0[0:c.size(0)] for c in conds]
    ld = discrim_loss(sd_d, xs, [f.detach() for f in fake], conds, fake_cond, draws["alphas"], loss, gp_lambda)
    d_names = param_names(sd_d)
    d_grads = torch.autograd.grad(ld, [sd_d[n] for n in d_names], allow_unused=True)
    out["lossD"] = float(ld.detach())
    out["gradD"] = {n: g for n, g in zip(d_names, d_grads) if g is not None}
    if reduce is not None:
        out["gradD"] = reduce(out["gradD"])
    if opt_d is not None:
        opt_d.step(sd_d, out["gradD"])
    # ---- real_pred with the UPDATED D (trainer.py:247), then G step (cond_gan.py:90-118)
    real_pred = discrim_forward(sd_d, xs, conds)
    lg = gen_loss(sd_d, fake, real_pred, conds, loss)
    g_names = param_names(sd_g)
    g_grads = torch.autograd.grad(lg, [sd_g[n] for n in g_names], allow_unused=True)
    out["lossG"] = float(lg.detach())
    out["gradG"] = {n: g for n, g in zip(g_names, g_grads) if g is not None}
    if reduce is not None:
        out["gradG"] = reduce(out["gradG"])
    if opt_g is not None:
        opt_g.step(sd_g, out["gradG"])
    for k, v in new_buf.items():
        sd_g[k] = v
    return out


def draw_real(n_levels=4, subsample_input=True):
    return [draw_bt() for _ in range(n_levels)] if (subsample_input and n_levels > 1) else []


def draw_rest(level_batches, conditional=True, gp=True):
    bt_fake = [draw_bt() for _ in range(len(level_batches) - 1)]
    perm = gen_perm(level_batches[0]) if conditional else None
    alphas = draw_gp_alphas(level_batches) if gp else None
    perm2 = gen_perm(level_batches[0]) if conditional else None   # trainer.py:247 draws again (unused)
    return {"bt_fake": bt_fake, "perm": perm, "alphas": alphas, "perm_unused": perm2}
