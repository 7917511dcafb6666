"""--G txt2vid.models.tganv2_cond.gen.MultiScaleGen (scripts/run_tganv2_cond.sh:20)."""
from txt2vid_b200.tganv2 import BaseFrameGen, MultiScaleGen  # noqa: F401
