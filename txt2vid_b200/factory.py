"""Model construction in the reference's order (txt2vid/train/gan.py:28-70): seed -> caption encoder (+init)
-> generator -> discriminator -> init(G) -> init(D).  The order fixes the RNG stream, hence the weights."""
import contextlib
import io
import random

import numpy as np
import torch


def seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def build_models(conditional=True, vocab_size=1000, seed=100, width=64, height=64, num_frames=16,
                 init_method="xavier"):
    from . import tganv2, text
    from .util import init
    if seed is not None:
        seed_all(seed)
    txt = None
    with contextlib.redirect_stdout(io.StringIO()):
        if conditional:
            txt = text.Seq2Seq(vocab_size=vocab_size)
            init(txt, init_method)
            gen = tganv2.MultiScaleGen(width=width, height=height, cond_dim=256, num_frames=num_frames)
            dis = tganv2.MultiScaleDiscrim(cond_dim=256)
        else:
            gen = tganv2.MultiScaleGenUncond(width=width, height=height, cond_dim=0, num_frames=num_frames)
            dis = tganv2.MultiScaleDiscrimUncond(cond_dim=0)
    init(gen, init_method)
    init(dis, init_method)
    return txt, gen, dis
