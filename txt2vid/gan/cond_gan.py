"""txt2vid.gan.cond_gan.CondGan (train/gan.py:111)."""
from txt2vid_b200.gan import CondGan  # noqa: F401
