"""txt2vid.train.setup (seed + device), train/setup.py:7-31."""
import random

import numpy as np
import torch

from txt2vid_b200.util import status, warn


def set_seed(seed):
    if seed is None:
        seed = random.randint(1, 100000)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


def set_cuda(use_cuda=False):
    if torch.cuda.is_available() and not use_cuda:
        warn('cuda is available')
    return torch.device("cuda:%d" % torch.cuda.current_device() if use_cuda else "cpu")


def setup(args):
    seed = set_seed(args.seed)
    device = set_cuda(use_cuda=args.cuda)
    status('Seed: %d' % seed)
    status('Device set to: %s' % device)
    return seed, device
