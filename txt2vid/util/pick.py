from txt2vid_b200.util import load  # noqa: F401
