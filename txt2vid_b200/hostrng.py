"""Host-RNG draws of one training iteration: frame offsets `bt`
(models/layers.py:107-108, torch CPU generator), the mismatched-caption permutation
(gan/cond_gan.py:133, numpy), the gradient-penalty alphas (gan/losses.py:140-145, torch CPU generator).

EagerDraws reproduces the reference: values are drawn at the call site, in call order.
StaticDraws serves the same calls from device buffers so that a captured CUDA graph can be replayed:
`refresh()` draws the whole iteration's values on the host -- same generators, same order -- and copies them
into the buffers the captured kernels read.
"""
import numpy as np
import torch

from .util import gen_perm


class EagerDraws(object):
    def begin_iteration(self):
        pass

    def bt(self, st=2):
        return int(torch.randint(st, (1,)))

    def perm(self, n, device):
        return torch.as_tensor(gen_perm(n), device=device)

    def alpha(self, B, device):
        return torch.rand(B).to(device)


class RecordingDraws(EagerDraws):
    """Eager draws that also record the (kind, size) sequence of one iteration."""

    def __init__(self):
        self.calls = []

    def begin_iteration(self):
        self.calls = []

    def bt(self, st=2):
        self.calls.append(("bt", st))
        return EagerDraws.bt(self, st)

    def perm(self, n, device):
        self.calls.append(("perm", n))
        return EagerDraws.perm(self, n, device)

    def alpha(self, B, device):
        self.calls.append(("alpha", B))
        return EagerDraws.alpha(self, B, device)


class StaticDraws(EagerDraws):
    RING = 4      # pinned staging slots: the host may run this many replays ahead of the device

    def __init__(self, calls, device):
        self.calls = list(calls)
        self.device = device
        self.host, self.dev = [[] for _ in range(self.RING)], []
        for kind, n in self.calls:
            if kind == "bt":
                h = torch.zeros(1, dtype=torch.int32)
            elif kind == "perm":
                h = torch.zeros(n, dtype=torch.int64)
            else:
                h = torch.zeros(n, dtype=torch.float32)
            for r in range(self.RING):
                self.host[r].append(h.clone().pin_memory() if device.type == "cuda" else h.clone())
            self.dev.append(torch.zeros_like(h, device=device))
        self.i = 0

    def refresh(self, slot=0):
        """Draw the next iteration's values in the recorded (= reference) order and stage them on the device.
        The copies are asynchronous: the caller owns `slot` (a pinned staging set nobody is still reading)."""
        hs = self.host[slot % self.RING]
        for (kind, n), h in zip(self.calls, hs):
            if kind == "bt":
                h[0] = int(torch.randint(n, (1,)))
            elif kind == "perm":
                h.copy_(torch.from_numpy(np.ascontiguousarray(gen_perm(n))))
            else:
                h.copy_(torch.rand(n))
        if self.device.type == "cuda":
            # one SM copy kernel reading the pinned slots over UVA (not ~14 cudaMemcpyAsync on the compute stream:
            # those queue behind the prefetcher's H2D pieces on the copy engine and stall the step)
            from . import kernels as K
            K.multi_copy([h.view(torch.float32) for h in hs], [d.view(torch.float32) for d in self.dev])
        else:
            for h, d in zip(hs, self.dev):
                d.copy_(h)

    def begin_iteration(self):
        self.i = 0

    def _next(self, kind, n):
        k, m = self.calls[self.i]
        assert (k, m) == (kind, n), "host-RNG call sequence changed: got %s, recorded %s" % ((kind, n), (k, m))
        d = self.dev[self.i]
        self.i += 1
        return d

    def bt(self, st=2):
        return self._next("bt", st)

    def perm(self, n, device):
        return self._next("perm", n)

    def alpha(self, B, device):
        return self._next("alpha", B)


CURRENT = EagerDraws()


def set_current(d):
    global CURRENT
    CURRENT = d
    return d
