from txt2vid_b200.util import ensure_exists  # noqa: F401
