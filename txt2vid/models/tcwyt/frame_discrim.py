"""--D txt2vid.models.tcwyt.frame_discrim.FrameDiscrim, --M ...FrameMap (scripts/run.sh:17)."""
from txt2vid_b200.tcwyt import FrameDiscrim, FrameMap  # noqa: F401
