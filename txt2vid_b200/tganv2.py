"""TGANv2 multi-scale generator / discriminator (unconditional and text-conditional), B200-native.

Mirrors txt2vid/models/tganv2/{gen,discrim}.py and txt2vid/models/tganv2_cond/{gen,discrim}.py of the
reference: same constructor kwargs, attribute names, state_dict keys and forward signatures /
return structures.  Internally everything between the latent and the rendered frames stays in
channels-last bf16 on the sm_100a kernels; there is no torch.nn.parallel anywhere: data parallelism is
one process per GPU (txt2vid_b200/parallel.py).
"""
import torch
import torch.nn as nn

from . import ops
from .blocks import ConvLSTM, RenderBlock, Resnet3D, Subsample, UpBlock


class BaseFrameGen(nn.Module):
    """Three UpBlocks 1024 -> 512 -> 256 -> 128, x8 upscaling (tganv2_cond/gen.py:7-20)."""

    def __init__(self, in_channels=1024, out_channels=128):
        super().__init__()
        self.out_channels = 128
        self.up0 = UpBlock(in_channels=in_channels, out_channels=512)
        self.up1 = UpBlock(in_channels=512, out_channels=256)
        self.up2 = UpBlock(in_channels=256, out_channels=out_channels)

    def forward_cl(self, x):
        return self.up2.forward_cl(self.up1.forward_cl(self.up0.forward_cl(x)))

    def forward(self, x, cond=None):
        return self.up2(self.up1(self.up0(x)))


class _MultiScaleGenBase(nn.Module):
    """Shared body of the conditional (tganv2_cond/gen.py:22-124) and unconditional
    (tganv2/gen.py:22-119) generators."""

    def _build(self, latent_size, width, height, num_channels, additional_blocks, fm_channels, num_frames, cond_dim,
               no_lstm, non_local):
        self.subsample = Subsample()
        self.latent_size = latent_size
        self.fm_channels = fm_channels
        self.fm_width = max(1, (width // 64))
        self.fm_height = max(1, (height // 64))
        self.latent_plane_ch = self.fm_channels
        self.fm_size = self.fm_width * self.fm_height * self.latent_plane_ch
        self.num_frames = num_frames
        self.num_channels = num_channels
        self.fc = nn.Linear(latent_size + cond_dim, self.fm_size)
        self.no_lstm = no_lstm
        assert not no_lstm, "the FrameSeedGenerator variant (no_lstm=True) is outside the TGANv2 hot path"
        self.clstm = ConvLSTM(input_channels=self.latent_plane_ch, hidden_channels=[self.fm_channels], kernel_size=3,
                              step=num_frames, effective_step=range(num_frames))
        base = BaseFrameGen()
        self.render_blocks = [RenderBlock(in_channels=base.out_channels, out_channels=num_channels)]
        self.abstract_blocks = [base]
        for i, block in enumerate(additional_blocks):
            prev = self.abstract_blocks[i].out_channels
            self.abstract_blocks.append(UpBlock(in_channels=prev, out_channels=block,
                                                with_non_local=non_local and i == len(additional_blocks) - 2))
            self.render_blocks.append(RenderBlock(in_channels=block, out_channels=num_channels))
        self.abstract_blocks = nn.ModuleList(self.abstract_blocks)
        self.render_blocks = nn.ModuleList(self.render_blocks)

    def forward(self, x, cond=None, return_abstract_maps=False, output_blocks=None):
        if cond is not None and self.fc.in_features != self.latent_size:
            x = torch.cat((x, cond), dim=1)
        B = x.size(0)
        # fc -> (B, fm_ch, fh, fw) plane (gen.py:70-72).  Linear == 1x1x1 conv on B positions.
        zc = ops.to_cl(x.view(B, x.size(1), 1, 1, 1))
        plane = ops.conv(zc, self.fc.weight, self.fc.bias)                       # (B,1,1,1,fm_size)
        if self.fm_height * self.fm_width > 1:
            plane = ops.from_cl(plane, self.fm_size).view(B, self.latent_plane_ch, 1, self.fm_height, self.fm_width)
            plane = ops.to_cl(plane)
        h = self.clstm.forward_cl(plane)                                          # (B*T,1,fh,fw,C), (b,t) order
        T = self.num_frames
        Bc = B
        rendered, abstract = [], []
        n = len(self.render_blocks)
        for i in range(n):
            if i != 0 and self.training:
                bt = self.subsample.draw()                                        # gen.py:101-109
                h = ops.gather_frames(h, Bc, T, bt)
                Bc = (Bc + 1) // 2
                if isinstance(bt, torch.Tensor):
                    T //= 2                                                       # graph mode: T is even
                else:
                    T = (T - bt + 1) // 2 if T > bt else 0
            blk = self.abstract_blocks[i]
            h = blk.forward_cl(h)
            abstract.append(h)
            if i == n - 1 or self.training or (output_blocks is not None and i in output_blocks):
                rendered.append(self.render_blocks[i].forward_cl(h, Bc, T))
        if return_abstract_maps:
            return rendered, [ops.from_cl(a, a.shape[-1]).squeeze(2) for a in abstract]
        return rendered


class MultiScaleGen(_MultiScaleGenBase):
    """txt2vid.models.tganv2_cond.gen.MultiScaleGen"""

    def __init__(self, latent_size=256, width=64, height=64, num_channels=3, additional_blocks=[64, 32, 32],
                 fm_channels=1024, num_frames=16, cond_dim=256, no_lstm=False):
        super().__init__()
        self._build(latent_size, width, height, num_channels, additional_blocks, fm_channels, num_frames, cond_dim,
                    no_lstm, non_local=True)


class MultiScaleGenUncond(_MultiScaleGenBase):
    """txt2vid.models.tganv2.gen.MultiScaleGen (note the reference's different defaults: 128x128,
    cond_dim ignored by fc, no non-local block)."""

    def __init__(self, latent_size=256, width=128, height=128, num_channels=3, additional_blocks=[64, 32, 32],
                 fm_channels=1024, num_frames=16, cond_dim=0, no_lstm=False):
        super().__init__()
        self._build(latent_size, width, height, num_channels, additional_blocks, fm_channels, num_frames, 0,
                    no_lstm, non_local=False)


class _PassThrough(nn.Module):
    """Keeps the `single_discrim.module.*` state_dict keys that nn.DataParallel gives the reference's
    conditional discriminator (tganv2_cond/discrim.py:15) without any replication machinery."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


class MultiScaleDiscrim(nn.Module):
    """txt2vid.models.tganv2_cond.discrim.MultiScaleDiscrim: one shared Resnet3D applied per level."""

    _wrap = True

    def __init__(self, discrim_down_blocks=[4, 4, 4, 4], num_channels=3, cond_dim=0, underlying_discrim=Resnet3D,
                 single_discrim=True):
        super().__init__()
        wrap = _PassThrough if self._wrap else (lambda m: m)
        if single_discrim:
            self.single_discrim = wrap(underlying_discrim(cond_dim=cond_dim, num_down_blocks=discrim_down_blocks[-1],
                                                          num_channels=num_channels))
            self.sub_discrims = [self.single_discrim for _ in range(len(discrim_down_blocks))]
        else:
            self.single_discrim = None
            self.sub_discrims = nn.ModuleList(
                [wrap(underlying_discrim(cond_dim=cond_dim, num_down_blocks=db, num_channels=num_channels))
                 for db in discrim_down_blocks])

    def forward(self, x=None, cond=None, xbar=None, computed_features=None):
        out = []
        for i, r in enumerate(x):
            cond_i = cond[i] if cond is not None else None
            xbar_i = xbar[i] if xbar is not None else None
            # The reference accepts `computed_features` but never forwards it (cf_i stays None,
            # discrim.py:35,40-41) and recomputes an identical trunk pass.  Here the shortcut works:
            # same values, same gradients (the features node is shared), one trunk pass fewer.
            cf_i = computed_features[i] if computed_features is not None else None
            out.append(self.sub_discrims[i](r, cond=cond_i, xbar=xbar_i, computed_features=cf_i))
        return out


    def forward_pair(self, x_a=None, x_b=None, cond_a=None, cond_b=None):
        """Two inputs through the shared trunk in ONE pass per level: D(x_a, cond_a) and D(x_b, cond_b) with the
        clips concatenated along the batch.  The discriminator has no normalisation layer, so
        every sample is processed independently and the results equal two separate calls; the D step of
        cond_gan.py:42-53 (real and fake pair) then launches half as many kernels on twice the positions."""
        out_a, out_b = [], []
        for i, (ra, rb) in enumerate(zip(x_a, x_b)):
            d = self.sub_discrims[i]
            d = d.module if hasattr(d, "module") else d
            na = ra.size(0)
            feat = d.features(torch.cat((ra, rb), dim=0))
            fa, fb = feat[:na], feat[na:]
            ca = cond_a[i] if cond_a is not None else None
            cb = cond_b[i] if cond_b is not None else None
            out_a.append(d.heads(fa, ca))
            out_b.append(d.heads(fb, cb))
        return out_a, out_b


class MultiScaleDiscrimUncond(MultiScaleDiscrim):
    """txt2vid.models.tganv2.discrim.MultiScaleDiscrim: no DataParallel wrapper (plain keys); forward
    takes (x, cond=None, xbar=None) and ignores cond (tganv2/discrim.py:23-31)."""

    _wrap = False

    def forward(self, x=None, cond=None, xbar=None):
        return [self.sub_discrims[i](r) for i, r in enumerate(x)]
