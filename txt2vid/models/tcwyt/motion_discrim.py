"""--D txt2vid.models.tcwyt.motion_discrim.MotionDiscrim (scripts/run.sh:17)."""
from txt2vid_b200.tcwyt import MotionDiscrim  # noqa: F401
