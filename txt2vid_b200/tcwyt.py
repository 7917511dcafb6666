"""TCWYT ("to create what you tell") model family on the sm_100a kernels (BASELINE config 2).

Mirrors txt2vid/models/tcwyt/{gen,video_discrim,frame_discrim,motion_discrim}.py: same constructor
arguments, nn.Sequential layouts (identical state_dict keys) and return conventions (VideoDiscrim returns
the scalar out.mean(), video_discrim.py:57; FrameDiscrim (T,B), MotionDiscrim (T-1,B)).  nn.* members hold
parameters only; compute goes through txt2vid_b200.ops."""
import torch
import torch.nn as nn

from . import ops


def _bn_lrelu(x, bn, slope):
    """BatchNorm + LeakyReLU: fused activation code 2 for the reference's slope 0.2, two kernels otherwise."""
    if abs(slope - 0.2) < 1e-12:
        return ops.bn_act(x, bn, 2)
    return ops.leaky_relu(ops.bn_act(x, bn, 0), slope)


def _conv_bn_lrelu(x, conv, bn, slope):
    """Conv3d -> BatchNorm3d -> LeakyReLU with the BatchNorm fed from the convolution's fp32 accumulator"""
    if abs(slope - 0.2) < 1e-12:
        return ops.conv_bn_act(x, conv, bn, 2)
    return ops.leaky_relu(ops.conv_bn_act(x, conv, bn, 0), slope)


def _tile_cond(c_cl, like):
    """(B,1,1,1,C) -> broadcast over the spatial extent of `like` (the reference's expand / sent_dupe loops)."""
    B, D, H, W, _ = like.shape
    return c_cl.expand(B, D, H, W, c_cl.shape[-1])


class Gen(nn.Module):
    """(z, cond) -> (B, 3, 16, 48, 48) (models/tcwyt/gen.py:5-49)."""

    def __init__(self, z_size=100, cond_dim=0, num_channels=3, scale_factor=1):
        super().__init__()
        self.cond_dim = cond_dim
        self.latent_size = z_size
        self.input_size = self.latent_size + self.cond_dim
        self.num_channels = num_channels
        s = scale_factor
        self.seq = nn.Sequential(
            nn.ConvTranspose3d(self.input_size, int(512 * s), kernel_size=(2, 6, 6), padding=0, bias=False),
            nn.BatchNorm3d(int(512 * s)), nn.LeakyReLU(0.2, True),
            nn.ConvTranspose3d(int(512 * s), int(256 * s), kernel_size=4, stride=2, padding=1, bias=False),
            nn.BatchNorm3d(int(256 * s)), nn.LeakyReLU(0.2, True),
            nn.ConvTranspose3d(int(256 * s), int(128 * s), kernel_size=4, stride=2, padding=1, bias=False),
            nn.BatchNorm3d(int(128 * s)), nn.LeakyReLU(0.2, True),
            nn.ConvTranspose3d(int(128 * s), int(64 * s), kernel_size=4, stride=2, padding=1, bias=False),
            nn.BatchNorm3d(int(64 * s)), nn.LeakyReLU(0.2, True),
            nn.ConvTranspose3d(int(64 * s), num_channels, kernel_size=1, stride=1, padding=0, bias=False),
            nn.Tanh())
        self.input_map = nn.Sequential(nn.Linear(self.input_size, self.input_size), nn.BatchNorm1d(self.input_size),
                                       nn.LeakyReLU(0.2, True))

    def forward(self, x, cond=None):
        if cond is not None:
            x = torch.cat((x, cond), dim=1)
        B = x.size(0)
        h = ops.vec_to_cl(x.reshape(B, x.size(1)))
        h = ops.bn_act(ops.linear_cl(h, self.input_map[0]), self.input_map[1], 2)
        for i in (0, 3, 6, 9):
            h = ops.conv_bn_act(h, self.seq[i], self.seq[i + 1], 2, transpose=True)
        pre = ops.gconv_transpose(h, self.seq[12])                    # (B, T, H, W, Cp)
        _, T, H, W, Cp = pre.shape
        return ops.render_tail(pre.reshape(B * T, 1, H, W, Cp), B, T, self.num_channels)


class VideoDiscrim(nn.Module):
    """3-D conv discriminator; returns the scalar mean of its output map (models/tcwyt/video_discrim.py:4-57).
    Also TGAN's Discrim (models/tgan/discrim.py:2)."""

    def __init__(self, cond_dim=256, mid_ch=64, num_channels=3, negative_slope=0.2, which_conv=nn.Conv3d):
        super().__init__()
        self.slope = negative_slope
        self.f = nn.LeakyReLU(negative_slope, True)
        self.x_map = nn.Sequential(
            which_conv(num_channels, mid_ch, 4, 2, 1, bias=False), self.f,
            which_conv(mid_ch, mid_ch * 2, 4, 2, 1, bias=False), nn.BatchNorm3d(mid_ch * 2), self.f,
            which_conv(mid_ch * 2, mid_ch * 4, 4, 2, 1, bias=False), nn.BatchNorm3d(mid_ch * 4), self.f,
            which_conv(mid_ch * 4, mid_ch * 8, 4, 2, 1, bias=False), nn.BatchNorm3d(mid_ch * 8), self.f)
        if cond_dim > 0:
            self.cond_map = nn.Sequential(nn.Linear(cond_dim, cond_dim), nn.BatchNorm1d(cond_dim), self.f)
            self.pred = nn.Sequential(which_conv(mid_ch * 8 + cond_dim, 512, 1, 1, 0, bias=False), nn.BatchNorm3d(512),
                                      self.f, which_conv(mid_ch * 8, 1, (1, 3, 3), 1, 0, bias=False))
        else:
            self.pred = which_conv(mid_ch * 8, 1, (1, 3, 3), 2, 0, bias=False)

    def forward(self, x=None, cond=None, xbar=None):
        h = ops.leaky_relu(ops.gconv(ops.to_cl(x), self.x_map[0]), self.slope)
        for i in (2, 5, 8):
            h = _conv_bn_lrelu(h, self.x_map[i], self.x_map[i + 1], self.slope)
        if cond is not None:
            c = _bn_lrelu(ops.linear_cl(ops.vec_to_cl(cond), self.cond_map[0]), self.cond_map[1], self.slope)
            h = torch.cat((h, _tile_cond(c, h)), dim=-1)
            h = _conv_bn_lrelu(h, self.pred[0], self.pred[1], self.slope)
            out = ops.gconv(h, self.pred[3], out_f32=True)
        else:
            out = ops.gconv(h, self.pred, out_f32=True)
        return out[..., 0].float().reshape(out.shape[0], -1).mean()


class FrameMap(nn.Module):
    """per-frame Conv2d k4 s2 p1 stack, BatchNorm statistics per frame call (frame_discrim.py:4-36):
    (B,3,T,H,W) -> (T,B,512,h,w)."""

    def __init__(self, num_channels=3):
        super().__init__()
        self.frame_map = nn.Sequential(
            nn.Conv2d(num_channels, 64, 4, 2, 1, bias=False), nn.BatchNorm2d(64), nn.LeakyReLU(0.2, True),
            nn.Conv2d(64, 128, 4, 2, 1, bias=False), nn.BatchNorm2d(128), nn.LeakyReLU(0.2, True),
            nn.Conv2d(128, 256, 4, 2, 1, bias=False), nn.BatchNorm2d(256), nn.LeakyReLU(0.2, True),
            nn.Conv2d(256, 512, 4, 2, 1, bias=False))

    def forward_cl(self, videos):
        """-> list of T CL maps (B,1,h,w,512)"""
        m = self.frame_map
        xc = ops.to_cl(videos)                                        # (B,T,H,W,16)
        out = []
        for t in range(xc.shape[1]):
            h = xc[:, t:t + 1].contiguous()
            h = ops.conv_bn_act(h, m[0], m[1], 2)
            h = ops.conv_bn_act(h, m[3], m[4], 2)
            h = ops.conv_bn_act(h, m[6], m[7], 2)
            out.append(ops.gconv(h, m[9]))
        return out

    def forward(self, videos):
        return torch.stack([ops.from_cl(f, 512)[:, :, 0] for f in self.forward_cl(videos)])


class _PerFrameDiscrim(nn.Module):
    """shared body of FrameDiscrim / MotionDiscrim (frame_discrim.py:39-84, motion_discrim.py:4-52)."""
    trunk_name = "frame_map"

    def __init__(self, cond_dim=256):
        super().__init__()
        setattr(self, self.trunk_name, nn.Sequential(nn.Conv2d(512, 512, 1, 1, 0, bias=False), nn.BatchNorm2d(512),
                                                     nn.LeakyReLU(0.2, True)))
        self.predictor = nn.Sequential(nn.Conv2d(512 + cond_dim, 512, 1, 1, 0, bias=False), nn.BatchNorm2d(512),
                                       nn.LeakyReLU(0.2, True), nn.Conv2d(512, 1, 2, 2, 0, bias=False))
        self.sent_map = nn.Sequential(nn.Linear(cond_dim, cond_dim), nn.BatchNorm1d(cond_dim), nn.LeakyReLU(0.2, True))

    def _frames_cl(self, xbar):
        if isinstance(xbar, (list, tuple)):
            return list(xbar)
        return [ops.to_cl(xbar[i].unsqueeze(2)) for i in range(xbar.size(0))]     # (B,512,h,w) -> (B,1,h,w,512)

    def _heads(self, frames, cond):
        trunk = getattr(self, self.trunk_name)
        sent = ops.bn_act(ops.linear_cl(ops.vec_to_cl(cond), self.sent_map[0]), self.sent_map[1], 2)
        outs = []
        for f in frames:
            h = ops.bn_act(ops.conv(f, trunk[0].weight), trunk[1], 2)
            h = torch.cat((h, _tile_cond(sent, h)), dim=-1)
            h = ops.bn_act(ops.conv(h, self.predictor[0].weight), self.predictor[1], 2)
            o = ops.gconv(h, self.predictor[3], out_f32=True)
            outs.append(o[..., 0].float().reshape(o.shape[0], -1).squeeze(1))
        return torch.stack(outs, 0)


class FrameDiscrim(_PerFrameDiscrim):
    trunk_name = "frame_map"

    def forward(self, x=None, cond=None, xbar=None):
        return self._heads(self._frames_cl(xbar), cond)


class MotionDiscrim(_PerFrameDiscrim):
    trunk_name = "motion_map"

    def forward(self, x=None, cond=None, xbar=None):
        fr = self._frames_cl(xbar)
        return self._heads([fr[i + 1] - fr[i] for i in range(len(fr) - 1)], cond)
