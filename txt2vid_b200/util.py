"""Small host utilities with the reference's names and behaviour (txt2vid/util/*)."""
import math
import time
from functools import partial
from pathlib import Path

import numpy as np
import torch.nn.init as tinit


# ---- util/misc.py
def gen_perm(n):
    """A numpy permutation of range(n) that is not the identity (util/misc.py:3-8); consumes the global
    numpy RNG exactly like the reference (one draw per attempt)."""
    ident = np.arange(n)
    while True:
        perm = np.random.permutation(ident)
        if not (perm == ident).all():
            return perm


def count_params(model):
    return sum(p.numel() for p in model.parameters())


# ---- util/torch/init.py
def _weight_init(layer, init_func=None):
    name = layer.__class__.__name__
    if 'Linear' in name or 'Conv' in name or 'Embedding' in name:
        w = getattr(layer, 'weight', None)
        if w is not None:
            if getattr(layer, 'is_residual', False):
                init_func(w, gain=math.sqrt(2))
            else:
                init_func(w)
        b = getattr(layer, 'bias', None)
        if b is not None:
            b.data.fill_(0.0)
    elif 'BatchNorm' in name:
        if getattr(layer, 'weight', None) is not None:
            layer.weight.data.fill_(1.0)
        if getattr(layer, 'bias', None) is not None:
            layer.bias.data.fill_(0.0)


def init(model, init_method=None):
    """xavier / ortho / normal initialisation in module.apply order (util/torch/init.py:4-39)."""
    funcs = {'xavier': tinit.xavier_normal_, 'ortho': tinit.orthogonal_,
             'normal': partial(tinit.normal_, mean=0, std=0.02)}
    assert init_method in funcs
    model.apply(partial(_weight_init, init_func=funcs[init_method]))


# ---- util/reflection.py
def get_class(kls):
    parts = kls.split('.')
    m = __import__(".".join(parts[:-1]))
    for comp in parts[1:]:
        m = getattr(m, comp)
    return m


def create_object_json(json_obj, **kwargs):
    clz = get_class(json_obj['class'])
    args = dict(json_obj.get('args', {}))
    args.update(kwargs)
    return clz(**args)


def create_object_file(json_file_path, **kwargs):
    import json
    with open(json_file_path) as f:
        params = json.load(f)
    assert 'class' in params
    return create_object(params, **kwargs)


def create_object(json_or_file, **kwargs):
    """"pkg.mod.Class" | path to {"class":..,"args":..} json | dict -> instance (util/reflection.py:12-50)."""
    if isinstance(json_or_file, str):
        if Path(json_or_file).exists():
            return create_object_file(json_or_file, **kwargs)
        return create_object_json({'class': json_or_file}, **kwargs)
    assert isinstance(json_or_file, dict)
    return create_object_json(json_or_file, **kwargs)


# ---- util/metrics.py, stopwatch.py, log.py, dir.py, pick.py
class RollingAvg(object):
    """Mean of the last `window_size` values (util/metrics.py:3-23); get() on an empty window asserts."""

    def __init__(self, window_size=100):
        import collections
        self.window_size = window_size
        self.window = collections.deque(maxlen=window_size)

    def update(self, x):
        self.window.append(x)

    def get(self):
        assert len(self.window) != 0
        return sum(self.window) / len(self.window)


class Stopwatch(object):
    """Wall-clock stopwatch (util/stopwatch.py:3-22)."""

    def __init__(self, should_start=False):
        self.reset()
        if should_start:
            self.start()

    def start(self):
        self.t1 = time.time()

    def stop(self):
        self.t2 = time.time()

    def reset(self):
        self.t1 = self.t2 = 0

    @property
    def elapsed_time(self):
        return self.t2 - self.t1


def _stamp(kind, msg):
    import datetime
    print('%s [%s]: %s' % (datetime.datetime.now(), kind, msg))


def status(msg):
    _stamp('STATUS', msg)


def warn(msg):
    _stamp('WARNING', msg)


def error(msg):
    _stamp('ERROR', msg)


def ensure_exists(path):
    import os
    os.makedirs(path, exist_ok=True)


def load(path):
    import pickle
    with open(path, 'rb') as f:
        return pickle.load(f)
