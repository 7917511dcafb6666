"""Drop-in path of the reference's txt2vid/models/tgan/temporal_gen.py."""
from txt2vid_b200.tgan import FrameSeedGenerator  # noqa: F401
