"""Drop-in path of the reference's txt2vid/models/resnet3d.py."""
from txt2vid_b200.blocks import Resnet3D  # noqa: F401
