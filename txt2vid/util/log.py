from txt2vid_b200.util import error, status, warn  # noqa: F401
