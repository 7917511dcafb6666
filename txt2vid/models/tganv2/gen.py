"""--G txt2vid.models.tganv2.gen.MultiScaleGen (scripts/run_tganv2.sh:18)."""
from txt2vid_b200.tganv2 import BaseFrameGen  # noqa: F401
from txt2vid_b200.tganv2 import MultiScaleGenUncond as MultiScaleGen  # noqa: F401
