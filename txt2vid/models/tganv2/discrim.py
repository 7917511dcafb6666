"""--D txt2vid.models.tganv2.discrim.MultiScaleDiscrim (scripts/run_tganv2.sh:18)."""
from txt2vid_b200.tganv2 import MultiScaleDiscrimUncond as MultiScaleDiscrim  # noqa: F401
