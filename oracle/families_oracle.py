"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference's TGAN and TCWYT model families
(BASELINE configs 1 and 2).  Plain functional PyTorch on state dicts; every
function cites the reference lines it restates.  Pinned against the LIVE reference by
oracle/make_golden_families.py (tests/golden/tgan_B*.json, tcwyt_B*.json); only tests/, smoke() and
bench.py's CPU legs may import this file.

At HEAD these two families cannot run through CondGan.discrim_step as-is (scalar-returning discriminator,
txt2vid/models/tcwyt/video_discrim.py:57), so -- as the survey prescribes -- the
iteration here drives G, the discriminators and the loss classes directly:

    fake = G(z[, cond]);  D-phase on (x, fake.detach());  G-phase on D(fake)

with one loss term per discriminator averaged over discriminators (cond_gan.py:26-31 default).
"""
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ helpers
def _bn(x, sd, p, new_buffers=None, eps=1e-5, momentum=0.1):
    """nn.BatchNorm{1,2,3}d in train(): batch statistics, running-stat update returned in new_buffers."""
    dims = [0] + list(range(2, x.dim()))
    mean = x.mean(dims)
    var = x.var(dims, unbiased=False)
    if new_buffers is not None:
        n = x.numel() // x.shape[1]
        new_buffers[p + "running_mean"] = (1 - momentum) * sd[p + "running_mean"] + momentum * mean.detach()
        new_buffers[p + "running_var"] = (1 - momentum) * sd[p + "running_var"] + \
            momentum * var.detach() * (n / max(n - 1, 1))
    shape = [1, -1] + [1] * (x.dim() - 2)
    xhat = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps)
    return xhat * sd[p + "weight"].view(shape) + sd[p + "bias"].view(shape)


def _lrelu(x, slope=0.2):
    return F.leaky_relu(x, slope)


# ------------------------------------------------------------------------------------------------ TGAN
def frame_seed_generator(sd, z_slow, p="_fsgen.", nb=None):
    """FrameSeedGenerator.forward (models/tgan/temporal_gen.py:26-34): (B, zs) -> (B, zf, 16)."""
    h = z_slow.view(z_slow.size(0), -1, 1)
    h = F.relu(_bn(F.conv_transpose1d(h, sd[p + "dc0.weight"], sd[p + "dc0.bias"], 1, 0), sd, p + "bn0.", nb))
    for i in (1, 2, 3):
        h = F.relu(_bn(F.conv_transpose1d(h, sd[p + "dc%d.weight" % i], sd[p + "dc%d.bias" % i], 2, 1), sd,
                       p + "bn%d." % i, nb))
    return torch.tanh(F.conv_transpose1d(h, sd[p + "dc4.weight"], sd[p + "dc4.bias"], 2, 1))


def video_frame_generator(sd, z_slow, z_fast, p="_vgen.", nb=None, ch=512, bw=4):
    """VideoFrameGenerator.forward (models/tgan/gen.py:34-48): per-frame latents -> (n, 3, 64, 64)."""
    n = z_slow.size(0)
    hs = F.relu(_bn(F.linear(z_slow, sd[p + "l0s.weight"], sd[p + "l0s.bias"]), sd, p + "bn0s.", nb)).view(n, ch // 2, bw, bw)
    hf = F.relu(_bn(F.linear(z_fast, sd[p + "l0f.weight"], sd[p + "l0f.bias"]), sd, p + "bn0f.", nb)).view(n, ch // 2, bw, bw)
    h = torch.cat((hs, hf), 1)
    for i in (1, 2, 3, 4):
        h = F.relu(_bn(F.conv_transpose2d(h, sd[p + "dc%d.weight" % i], sd[p + "dc%d.bias" % i], 2, 1), sd,
                       p + "bn%d." % i, nb))
    return torch.tanh(F.conv_transpose2d(h, sd[p + "dc5.weight"], sd[p + "dc5.bias"], 1, 1))


def tgan_gen(sd, z_slow, cond=None, nb=None):
    """Gen.forward (models/tgan/gen.py:56-74) -> (B, 3, 16, 64, 64); the two debug prints are dropped."""
    if cond is not None:
        z_slow = torch.cat((z_slow, cond), dim=-1)
    z_fast = frame_seed_generator(sd, z_slow, nb=nb)
    B, nzf, T = z_fast.shape
    z_fast = z_fast.permute(0, 2, 1).contiguous().view(B * T, nzf)
    zs = z_slow.unsqueeze(1).repeat(1, T, 1).contiguous().view(B * T, -1)
    out = video_frame_generator(sd, zs, z_fast, nb=nb)
    return out.view(B, T, out.shape[1], 64, 64).permute(0, 2, 1, 3, 4)


def video_discrim(sd, x, cond=None, p="", nb=None):
    """VideoDiscrim.forward (models/tcwyt/video_discrim.py:48-57); TGAN's Discrim is the same class
    (models/tgan/discrim.py:2).  Returns the scalar out.mean()."""
    h = _lrelu(F.conv3d(x, sd[p + "x_map.0.weight"], None, 2, 1))
    for i in (2, 5, 8):
        h = _lrelu(_bn(F.conv3d(h, sd[p + "x_map.%d.weight" % i], None, 2, 1), sd, p + "x_map.%d." % (i + 1), nb))
    if cond is not None:
        c = _lrelu(_bn(F.linear(cond, sd[p + "cond_map.0.weight"], sd[p + "cond_map.0.bias"]), sd, p + "cond_map.1.", nb))
        c = c.view(c.size(0), -1, 1, 1, 1).expand(-1, -1, h.size(2), h.size(3), h.size(4))
        h = torch.cat((h, c), dim=1)
        h = _lrelu(_bn(F.conv3d(h, sd[p + "pred.0.weight"]), sd, p + "pred.1.", nb))
        out = F.conv3d(h, sd[p + "pred.3.weight"])
    else:
        out = F.conv3d(h, sd[p + "pred.weight"], None, 2, 0)
    return out.view(out.size(0), -1).mean()


# ------------------------------------------------------------------------------------------------ TCWYT
def tcwyt_gen(sd, z, cond=None, nb=None):
    """Gen.forward (models/tcwyt/gen.py:41-49) -> (B, 3, 16, 48, 48)."""
    x = torch.cat((z, cond), dim=1) if cond is not None else z
    x = _lrelu(_bn(F.linear(x, sd["input_map.0.weight"], sd["input_map.0.bias"]), sd, "input_map.1.", nb))
    h = x.view(x.size(0), x.size(1), 1, 1, 1)
    h = _lrelu(_bn(F.conv_transpose3d(h, sd["seq.0.weight"]), sd, "seq.1.", nb))
    for i in (3, 6, 9):
        h = _lrelu(_bn(F.conv_transpose3d(h, sd["seq.%d.weight" % i], None, 2, 1), sd, "seq.%d." % (i + 1), nb))
    return torch.tanh(F.conv_transpose3d(h, sd["seq.12.weight"]))


def frame_map(sd, videos, nb=None):
    """FrameMap.forward (models/tcwyt/frame_discrim.py:25-36): per-frame Conv2d stack (BatchNorm statistics
    per frame call) -> (T, B, 512, h, w)."""
    out = []
    for t in range(videos.size(2)):
        h = videos[:, :, t]
        h = _lrelu(_bn(F.conv2d(h, sd["frame_map.0.weight"], None, 2, 1), sd, "frame_map.1.", nb))
        h = _lrelu(_bn(F.conv2d(h, sd["frame_map.3.weight"], None, 2, 1), sd, "frame_map.4.", nb))
        h = _lrelu(_bn(F.conv2d(h, sd["frame_map.6.weight"], None, 2, 1), sd, "frame_map.7.", nb))
        out.append(F.conv2d(h, sd["frame_map.9.weight"], None, 2, 1))
        if nb is not None:               # running statistics chain through the 16 per-frame calls
            sd = dict(sd)
            sd.update(nb)
    return torch.stack(out)


def _per_frame_head(sd, frames, cond, trunk, nb):
    """shared body of FrameDiscrim / MotionDiscrim .forward (frame_discrim.py:62-84, motion_discrim.py:33-52)."""
    sent = _lrelu(_bn(F.linear(cond, sd["sent_map.0.weight"], sd["sent_map.0.bias"]), sd, "sent_map.1.", nb))
    outs = []
    for i in range(frames.size(0)):
        f = _lrelu(_bn(F.conv2d(frames[i], sd[trunk + ".0.weight"]), sd, trunk + ".1.", nb))
        s = sent.view(sent.size(0), -1, 1, 1).expand(-1, -1, f.size(2), f.size(3))
        h = torch.cat((f, s), dim=1)
        h = _lrelu(_bn(F.conv2d(h, sd["predictor.0.weight"]), sd, "predictor.1.", nb))
        o = F.conv2d(h, sd["predictor.3.weight"], None, 2, 0)
        outs.append(o.view(o.size(0), -1).squeeze(1))
        if nb is not None:
            sd = dict(sd)
            sd.update(nb)
    return torch.stack(outs, 0)


def frame_discrim(sd, cond, xbar, nb=None):
    return _per_frame_head(sd, xbar, cond, "frame_map", nb)


def motion_discrim(sd, cond, xbar, nb=None):
    return _per_frame_head(sd, xbar[1:] - xbar[0:-1], cond, "motion_map", nb)


# ------------------------------------------------------------------------------------------------ losses
def wgan_d(fake, real):           # gan/losses.py:60-63
    return -(real.mean() - fake.mean())


def wgan_g(fake, real=None):      # gan/losses.py:65-68
    return -fake.mean()


def ralsgan_d(fake, real):        # gan/losses.py:117-124
    return (torch.mean((real - torch.mean(fake) - 1) ** 2) + torch.mean((fake - torch.mean(real) + 1) ** 2)) / 2


def ralsgan_g(fake, real):        # gan/losses.py:126-133
    return (torch.mean((real - torch.mean(fake) + 1) ** 2) + torch.mean((fake - torch.mean(real) - 1) ** 2)) / 2


# ------------------------------------------------------------------------------------------------ iterations
def leaves(sd, dtype=torch.float32):
    """state dict -> differentiable leaves; dtype=torch.float64 gives the rounding-free checker the formula
    tests use (the fp32 run of this oracle is itself 1e-2 away from it on cancellation-dominated tensors)."""
    out = {}
    for k, v in sd.items():
        t = v.detach().clone().to(dtype) if v.dtype.is_floating_point else v.detach().clone()
        if t.dtype.is_floating_point and "running_" not in k and "num_batches" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


def grads_of(sd):
    return {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}


def tgan_iteration(sd_g, sd_d, x, z):
    """One direct-drive TGAN iteration (WGAN loss, scripts/run_tgan.sh:17 without its gradient penalty):
    returns losses, the fake clip and gradients of D (from lossD) and G (from lossG)."""
    fake = tgan_gen(sd_g, z)
    d_fake, d_real = video_discrim(sd_d, fake.detach()), video_discrim(sd_d, x)
    lossD = wgan_d(d_fake, d_real)
    lossD.backward()
    gD = {k: v.clone() for k, v in grads_of(sd_d).items()}
    for v in sd_d.values():
        v.grad = None
    lossG = wgan_g(video_discrim(sd_d, fake))
    lossG.backward()
    return {"lossD": float(lossD), "lossG": float(lossG), "fake": fake.detach(), "gD": gD, "gG": grads_of(sd_g),
            "d_real": float(d_real), "d_fake": float(d_fake)}


def tcwyt_iteration(sd_g, sd_v, sd_f, sd_m, sd_map, x, z, cond):
    """One direct-drive TCWYT iteration: three discriminators (video / frame / motion) on FrameMap features
    (scripts/run.sh:17: RaLSGAN), losses averaged over discriminators (cond_gan.py:26-31)."""
    fake = tcwyt_gen(sd_g, z, cond)

    def d_all(vid):
        m = frame_map(sd_map, vid)
        return [video_discrim(sd_v, vid, cond), frame_discrim(sd_f, cond, m), motion_discrim(sd_m, cond, m)]
    real_o, fake_o = d_all(x), d_all(fake.detach())
    lossD = sum(ralsgan_d(f, r) for f, r in zip(fake_o, real_o)) / 3
    lossD.backward()
    gD = {}
    for name, sd in (("video", sd_v), ("frame", sd_f), ("motion", sd_m), ("map", sd_map)):
        gD[name] = {k: v.clone() for k, v in grads_of(sd).items()}
        for v in sd.values():
            v.grad = None
    fake_o2 = d_all(fake)
    lossG = sum(ralsgan_g(f, r.detach()) for f, r in zip(fake_o2, real_o)) / 3
    lossG.backward()
    return {"lossD": float(lossD), "lossG": float(lossG), "fake": fake.detach(), "gD": gD, "gG": grads_of(sd_g)}
