"""Data parallelism: one process per GPU, weights resident, ONE bucketed NCCL all-reduce per optimiser
step (D gradients after the D backward, G gradients after the G backward), SURVEY.md 8(e).

Replaces the reference's per-call torch.nn.parallel.data_parallel / nn.DataParallel replicate-scatter-
gather (models/tganv2_cond/gen.py:111,116; discrim.py:15), which re-broadcasts 29 M parameters 24 times
per iteration.  Per-rank semantics match the reference's own multi-GPU behaviour: BatchNorm statistics,
`x[::2]` batch striding and the caption permutation act on the local shard; the frame offsets `bt` come
from an identically seeded CPU generator on every rank; the gradient-penalty term (a SUM over samples,
gan/losses.py:203) is multiplied by the world size so that averaging gradients reproduces the reference's
full-batch loss.
"""
import os

import torch
import torch.distributed as dist

from . import kernels as K


class DistContext(object):
    def __init__(self, backend=None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.enabled = self.world > 1
        self._flat = {}
        if self.enabled and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend=backend, init_method="env://")

    @property
    def gp_scale(self):
        """factor for gp_lambda under gradient averaging (see module docstring)"""
        return float(self.world)

    def gp_lambda_for(self, gp_lambda, discrims):
        """gp_lambda under gradient AVERAGING across ranks: the multi-scale penalty is a SUM over samples
        (gan/losses.py:203, combine=torch.sum), so it is scaled by the world size; discriminators without
        `sub_discrims` use the batch MEAN (losses.py:135, combine=torch.mean), which averaging reproduces as is."""
        if not self.enabled or gp_lambda <= 0:
            return gp_lambda
        if all(hasattr(d, "sub_discrims") for d in discrims):
            return gp_lambda * self.gp_scale
        return gp_lambda

    def seed_ranks(self, seed):
        """Per-rank random streams for everything that must differ between shards (z on the CUDA generator, the
        caption permutation on numpy), while the torch CPU generator stays SHARED: the frame offsets `bt` of
        models/layers.py:107-108 are one draw per level per iteration in the reference."""
        import random
        import numpy as np
        torch.manual_seed(seed)                       # CPU generator: identical on every rank
        if torch.cuda.is_available():
            torch.cuda.manual_seed(seed + self.rank)
        np.random.seed(seed + self.rank)
        random.seed(seed + self.rank)

    def broadcast_module(self, module):
        """parameters and buffers of rank 0 to every rank (replicas must start identical)"""
        if not self.enabled or module is None:
            return
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=0)

    def shard(self, dataset):
        """every rank iterates its own stride of the batches (rank, rank + world, ...), equal counts on all ranks"""
        if not self.enabled:
            return dataset
        return _RankStride(dataset, self.rank, self.world)

    @property
    def is_main(self):
        return self.rank == 0

    def barrier(self):
        if self.enabled:
            dist.barrier()

    def reduce_grads(self, optimizer):
        """Average the gradients of every parameter of `optimizer` across ranks, in place."""
        if not self.enabled:
            return
        params = [p for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
        if not params:
            return
        # a 4-D weight tagged with `_t2v_live_tap` (ConvLSTM kernels on a 1x1 plane) only has gradient in that one
        # 3x3 tap: exchange the (Cout, Cin) slice, leave the structurally-zero taps alone
        def live(p):
            tap = getattr(p, "_t2v_live_tap", None)
            return tap if (tap is not None and p.dim() == 4) else None
        dense = [p for p in params if live(p) is None]
        sliced = [p for p in params if live(p) is not None]
        key = id(optimizer)
        total = sum(p.numel() for p in dense) + sum(p.shape[0] * p.shape[1] for p in sliced)
        flat = self._flat.get(key)
        if flat is None or flat.numel() != total or flat.device != params[0].device:
            flat = torch.empty(total, device=params[0].device, dtype=torch.float32)
            self._flat[key] = flat
        views, off = [], 0
        for p in dense:
            views.append(flat[off:off + p.numel()])
            off += p.numel()
        grads = [p.grad for p in dense]
        slices = []
        for p in sliced:
            n = p.shape[0] * p.shape[1]
            a, b = live(p)
            slices.append((p.grad[:, :, a, b], flat[off:off + n].view(p.shape[0], p.shape[1])))
            off += n
        if flat.is_cuda:
            K.multi_copy(grads, views)
        else:                                   # gloo / CPU tests of the host logic
            for g, v in zip(grads, views):
                v.copy_(_memory_order(g))
        for g, v in slices:
            v.copy_(g)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if hasattr(optimizer, "grad_scale"):
            optimizer.grad_scale = 1.0 / self.world           # the 1/N is folded into the fused Adam kernel
        else:
            flat.mul_(1.0 / self.world)
        if flat.is_cuda:
            K.multi_copy(views, grads)
        else:
            for g, v in zip(grads, views):
                _memory_order(g).copy_(v)
        for g, v in slices:
            g.copy_(v)
        self.last_bucket_bytes = 4 * total

    def all_reduce_max(self, value):
        if not self.enabled:
            return value
        t = torch.tensor([value], dtype=torch.float64, device="cuda" if torch.cuda.is_available() and
                         dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])


def _memory_order(t):
    """1-D view of a dense tensor in memory order (works for channels-last parameter gradients)."""
    return torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())


class _RankStride(object):
    """Iterable view of a batch iterable: batches rank, rank + world, ... ; truncated so that all ranks see the same
    number of batches (a collective per step must be entered by everyone)."""

    def __init__(self, dataset, rank, world):
        self.dataset, self.rank, self.world = dataset, rank, world

    def __len__(self):
        return len(self.dataset) // self.world

    def __iter__(self):
        n = len(self)
        for i, batch in enumerate(self.dataset):
            if i // self.world >= n:
                return
            if i % self.world == self.rank:
                yield batch
