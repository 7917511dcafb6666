from txt2vid_b200.util import Stopwatch  # noqa: F401
