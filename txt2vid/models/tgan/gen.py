"""--G txt2vid.models.tgan.gen.Gen (scripts/run_tgan.sh:17)."""
from txt2vid_b200.tgan import Gen, VideoFrameGenerator  # noqa: F401
