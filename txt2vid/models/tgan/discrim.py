"""--D txt2vid.models.tgan.discrim.Discrim (the reference re-exports tcwyt's VideoDiscrim, tgan/discrim.py:2)."""
from txt2vid_b200.tcwyt import VideoDiscrim as Discrim  # noqa: F401
