"""--D txt2vid.models.tcwyt.video_discrim.VideoDiscrim (scripts/run.sh:17)."""
from txt2vid_b200.tcwyt import VideoDiscrim  # noqa: F401
