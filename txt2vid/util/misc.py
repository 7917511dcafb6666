from txt2vid_b200.util import count_params, gen_perm  # noqa: F401
