from txt2vid_b200.util import _weight_init, init  # noqa: F401
