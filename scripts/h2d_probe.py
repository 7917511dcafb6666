"""Diagnosis: how long does the 805 MB pinned H2D copy of one batch take while a training step runs?"""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import train_params
from txt2vid_b200.data import SyntheticVideoCaptions
from txt2vid_b200.factory import build_models
from txt2vid_b200.gan import CondGan, MixedGanLoss, RSGANLoss
from txt2vid_b200.optim import FusedAdam
from txt2vid_b200.trainer import GraphedTrainStep

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
with contextlib.redirect_stdout(io.StringIO()):
    txt, gen, dis = build_models(True, vocab_size=1000, seed=100)
txt, gen, dis = txt.to(dev), gen.to(dev), dis.to(dev)
gan = CondGan(gen=gen, discrims=[dis], cond_encoder=txt, discrim_names=["video"])
losses = MixedGanLoss(g_loss=RSGANLoss(), d_loss=RSGANLoss())
optD = FusedAdam([{"params": dis.parameters()}], lr=2e-4, betas=(0.5, 0.999))
optG = FusedAdam([{"params": gen.parameters()}], lr=2e-4, betas=(0.5, 0.999))
step = GraphedTrainStep(gan, optD, optG, train_params(0.5), losses, dev, warmup=2)
x, t, l = SyntheticVideoCaptions(b, 1, vocab_size=1000).batch(0)
hx = x.contiguous().pin_memory()
dx, dt = x.to(dev), t.to(dev)
for _ in range(4):
    step(dx, [dt, l])
torch.cuda.synchronize()
side = torch.cuda.Stream()
stage = torch.empty_like(dx)

def run(n, copy, chunks=1, wait_main=False, after=False, mainwait=False):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s0.record()
    for i in range(n):
        if after:
            step(dx, [dt, l])
        if mainwait:
            torch.cuda.current_stream().wait_stream(side)
        if copy:
            if wait_main:
                side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                ev[i][0].record()
                if chunks == 1:
                    stage.copy_(hx, non_blocking=True)
                else:
                    for c, h in zip(stage.chunk(chunks), hx.chunk(chunks)):
                        c.copy_(h, non_blocking=True)
                ev[i][1].record()
        if not after:
            step(dx, [dt, l])
    s1.record()
    torch.cuda.synchronize()
    return s0.elapsed_time(s1) / n, [a.elapsed_time(b_) for a, b_ in ev] if copy else []

def run_pf(n, sync_every_step):
    """the bench's e2e structure: data_prefetcher + graphed step"""
    from txt2vid_b200.data import data_prefetcher
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s0.record()
    pf = data_prefetcher(((hx, t, l) for _ in range(n)), device=dev)
    ev = None
    for i in range(n):
        xx, yy = pf.next()
        step(xx, [yy[0], yy[1]])
        if sync_every_step:
            if ev is not None:
                ev.synchronize()
            ev = torch.cuda.Event(); ev.record()
    s1.record()
    torch.cuda.synchronize()
    return s0.elapsed_time(s1) / n

t = t.pin_memory()
torch.cuda.synchronize()
hs = []
for i in range(6):
    a_ = time.perf_counter(); step(dx, [dt, l]); hs.append((time.perf_counter() - a_) * 1e3)
torch.cuda.synchronize()
print("host time of one step() call (enqueue only): %s ms" % ["%.1f" % v for v in hs])
print("steps alone: %.2f ms/step" % run(10, False)[0])
from txt2vid_b200.data import data_prefetcher as _pf
for guard in ("none", "host", "none", "host"):
    _pf.GUARD = guard
    print("prefetcher loop (guard=%s): %.2f ms/step | host one step ahead: %.2f ms/step"
          % (guard, run_pf(10, False), run_pf(10, True)))
ms, c = run(10, True)
print("steps + concurrent H2D: %.2f ms/step; copy durations %s" % (ms, ["%.1f" % v for v in c]))
ms, c = run(10, True, 16)
print("steps + concurrent H2D in 16 chunks: %.2f ms/step; copy durations %s" % (ms, ["%.1f" % v for v in c]))
def run_v(n, variant):
    """build-up from the overlapping variant towards the prefetcher"""
    slots = [torch.empty_like(dx) for _ in range(3)]
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s0.record()
    with torch.cuda.stream(side):
        for c, h in zip(slots[0].chunk(16), hx.chunk(16)):
            c.copy_(h, non_blocking=True)
    for i in range(n):
        cur, nxt = slots[i % 3], slots[(i + 1) % 3]
        if variant >= 2:
            torch.cuda.current_stream().wait_stream(side)
        if variant >= 3:
            e = torch.cuda.Event(); e.record()
        with torch.cuda.stream(side):
            for c, h in zip(nxt.chunk(16), hx.chunk(16)):
                c.copy_(h, non_blocking=True)
            if variant >= 4:
                dtt = t.to(dev, non_blocking=True)
        step(cur if variant >= 1 else dx, [dt, l])
    s1.record()
    torch.cuda.synchronize()
    return s0.elapsed_time(s1) / n

for v in (4, 4):
    print("variant %d: %.2f ms/step" % (v, run_v(10, v)))
for kw in (dict(wait_main=True), dict(after=True), dict(mainwait=True), dict(wait_main=True, mainwait=True)):
    ms, c = run(10, True, 16, **kw)
    print("16 chunks %s: %.2f ms/step; copies %s" % (kw, ms, ["%.0f" % v for v in c]))
torch.cuda.synchronize()
a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); stage.copy_(hx, non_blocking=True); b_.record(); torch.cuda.synchronize()
print("H2D alone: %.2f ms" % a.elapsed_time(b_))
