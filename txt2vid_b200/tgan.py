"""TGAN model family on the sm_100a kernels (BASELINE config 1).

Mirrors txt2vid/models/tgan/{gen,temporal_gen,discrim}.py: same constructor arguments, sub-module names and
parameter shapes (identical state_dict keys and constructor-time RNG consumption); the nn.* members are
parameter containers only -- every FLOP goes through txt2vid_b200.ops (transposed convolutions =
t2v_gconv_dgrad, Linear = the tcgen05 1x1x1 engine, BatchNorm+ReLU fused kernels, render tail)."""
import torch
import torch.nn as nn

from . import ops
from .tcwyt import VideoDiscrim as Discrim  # noqa: F401  (models/tgan/discrim.py:2: "literally the same")


class FrameSeedGenerator(nn.Module):
    """z_slow (B, zs) -> 16 per-frame latents (B, zf, 16): a 1-D transposed-conv stack
    (models/tgan/temporal_gen.py:10-34)."""

    def __init__(self, z_slow_dim, z_fast_dim):
        super().__init__()
        self.z_slow_dim = z_slow_dim
        self.z_fast_dim = z_fast_dim
        self.dc0 = nn.ConvTranspose1d(z_slow_dim, 512, 1, 1, 0)
        self.dc1 = nn.ConvTranspose1d(512, 256, 4, 2, 1)
        self.dc2 = nn.ConvTranspose1d(256, 128, 4, 2, 1)
        self.dc3 = nn.ConvTranspose1d(128, 128, 4, 2, 1)
        self.dc4 = nn.ConvTranspose1d(128, z_fast_dim, 4, 2, 1)
        self.bn0 = nn.BatchNorm1d(512)
        self.bn1 = nn.BatchNorm1d(256)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(128)

    def forward_cl(self, z_cl):
        """CL (B,1,1,1,zsP) -> CL (B,1,1,16,zfP), tanh applied."""
        h = ops.conv_bn_act(z_cl, self.dc0, self.bn0, 1, transpose=True)
        h = ops.conv_bn_act(h, self.dc1, self.bn1, 1, transpose=True)
        h = ops.conv_bn_act(h, self.dc2, self.bn2, 1, transpose=True)
        h = ops.conv_bn_act(h, self.dc3, self.bn3, 1, transpose=True)
        return ops.tanh(ops.gconv_transpose(h, self.dc4))

    def forward(self, z_slow):
        y = self.forward_cl(ops.vec_to_cl(z_slow.reshape(z_slow.size(0), -1)))
        return ops.from_cl(y, self.z_fast_dim).reshape(z_slow.size(0), self.z_fast_dim, -1)


class VideoFrameGenerator(nn.Module):
    """per-frame (z_slow, z_fast) -> 64x64 frame: Linear+BN+ReLU x2, four k4 s2 p1 transposed convs with
    BN+ReLU, a 3x3 transposed conv and tanh (models/tgan/gen.py:10-48)."""

    def __init__(self, z_slow_dim, z_fast_dim, out_channels=3, bottom_width=4, conv_ch=512):
        super().__init__()
        self.ch = conv_ch
        self.bottom_width = bottom_width
        self.out_channels = out_channels
        slow_mid_dim = bottom_width * bottom_width * conv_ch // 2
        fast_mid_dim = bottom_width * bottom_width * conv_ch // 2
        self.l0s = nn.Linear(z_slow_dim, slow_mid_dim)
        self.l0f = nn.Linear(z_fast_dim, fast_mid_dim)
        self.dc1 = nn.ConvTranspose2d(conv_ch, conv_ch // 2, 4, 2, 1)
        self.dc2 = nn.ConvTranspose2d(conv_ch // 2, conv_ch // 4, 4, 2, 1)
        self.dc3 = nn.ConvTranspose2d(conv_ch // 4, conv_ch // 8, 4, 2, 1)
        self.dc4 = nn.ConvTranspose2d(conv_ch // 8, conv_ch // 16, 4, 2, 1)
        self.dc5 = nn.ConvTranspose2d(conv_ch // 16, out_channels, 3, 1, 1)
        self.bn0s = nn.BatchNorm1d(slow_mid_dim)
        self.bn0f = nn.BatchNorm1d(fast_mid_dim)
        self.bn1 = nn.BatchNorm2d(conv_ch // 2)
        self.bn2 = nn.BatchNorm2d(conv_ch // 4)
        self.bn3 = nn.BatchNorm2d(conv_ch // 8)
        self.bn4 = nn.BatchNorm2d(conv_ch // 16)

    def _bottom(self, z_cl, lin, bn):
        """Linear + BN1d + ReLU, then the reference's .view(n, ch/2, bw, bw) (channel-major features) as a
        CL map (n, 1, bw, bw, ch/2)."""
        n, bw, c = z_cl.shape[0], self.bottom_width, self.ch // 2
        h = ops.bn_act(ops.linear_cl(z_cl, lin), bn, 1)
        return h.reshape(n, c, bw, bw).permute(0, 2, 3, 1).reshape(n, 1, bw, bw, c).contiguous()

    def forward_cl(self, zs_cl, zf_cl):
        """-> pre-tanh CL (n, 1, 64, 64, round16(out_channels))"""
        h = torch.cat((self._bottom(zs_cl, self.l0s, self.bn0s), self._bottom(zf_cl, self.l0f, self.bn0f)), dim=-1)
        h = ops.conv_bn_act(h, self.dc1, self.bn1, 1, transpose=True)
        h = ops.conv_bn_act(h, self.dc2, self.bn2, 1, transpose=True)
        h = ops.conv_bn_act(h, self.dc3, self.bn3, 1, transpose=True)
        h = ops.conv_bn_act(h, self.dc4, self.bn4, 1, transpose=True)
        return ops.gconv_transpose(h, self.dc5)

    def forward(self, z_slow, z_fast):
        pre = self.forward_cl(ops.vec_to_cl(z_slow), ops.vec_to_cl(z_fast))
        n = pre.shape[0]
        return ops.render_tail(pre, n, 1, self.out_channels).reshape(n, self.out_channels, pre.shape[2], pre.shape[3])


class Gen(nn.Module):
    """TGAN generator (models/tgan/gen.py:50-78): z_slow (B, 256) [+ cond] -> (B, 3, 16, 64, 64)."""

    def __init__(self, z_slow_dim=256, z_fast_dim=256, cond_dim=0, out_channels=3, bottom_width=4, conv_ch=512):
        super().__init__()
        self.z_slow_plus_cond_dim = z_slow_dim + cond_dim
        self.z_slow_dim = z_slow_dim
        self.z_fast_dim = z_fast_dim
        self.out_channels = out_channels
        self._fsgen = FrameSeedGenerator(self.z_slow_plus_cond_dim, z_fast_dim)
        self._vgen = VideoFrameGenerator(self.z_slow_plus_cond_dim, z_fast_dim, out_channels, bottom_width, conv_ch)

    def forward(self, z_slow, cond=None):
        if cond is not None:
            z_slow = torch.cat((z_slow, cond), dim=-1)
        B = z_slow.size(0)
        zs = ops.vec_to_cl(z_slow)                                    # (B,1,1,1,P)
        zf = self._fsgen.forward_cl(zs)                               # (B,1,1,T,zfP): frame t of sample b
        T = zf.shape[3]
        zf = zf.reshape(B * T, 1, 1, 1, zf.shape[-1])                 # (b, t) frame order, as the reference
        zs = zs.expand(B, T, 1, 1, zs.shape[-1]).reshape(B * T, 1, 1, 1, zs.shape[-1]).contiguous()
        pre = self._vgen.forward_cl(zs, zf)
        return ops.render_tail(pre, B, T, self.out_channels)          # tanh + (B, C, T, 64, 64)

    @property
    def latent_size(self):
        return self.z_slow_dim
