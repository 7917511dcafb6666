"""--D txt2vid.models.tganv2_cond.discrim.MultiScaleDiscrim (scripts/run_tganv2_cond.sh:20)."""
from txt2vid_b200.tganv2 import MultiScaleDiscrim  # noqa: F401
