from txt2vid_b200.util import RollingAvg  # noqa: F401
