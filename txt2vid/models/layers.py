"""Drop-in path of the reference's txt2vid/models/layers.py -> B200-native blocks."""
from txt2vid_b200.blocks import (Attention, Attention3d, DownBlock, DownSample, Identity, RenderBlock,  # noqa: F401
                                 ResidualBlock, Subsample, UpBlock)
