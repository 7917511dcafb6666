"""--G txt2vid.models.tcwyt.gen.Gen (scripts/run.sh:17)."""
from txt2vid_b200.tcwyt import Gen  # noqa: F401
