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

def run(n, copy, chunks=1):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s0.record()
    for i in range(n):
        if copy:
            with torch.cuda.stream(side):
                ev[i][0].record()
                if chunks == 1:
                    stage.copy_(hx, non_blocking=True)
                else:
                    for c, h in zip(stage.chunk(chunks), hx.chunk(chunks)):
                        c.copy_(h, non_blocking=True)
                ev[i][1].record()
        step(dx, [dt, l])
    s1.record()
    torch.cuda.synchronize()
    return s0.elapsed_time(s1) / n, [a.elapsed_time(b_) for a, b_ in ev] if copy else []

print("steps alone: %.2f ms/step" % run(6, False)[0])
ms, c = run(6, True)
print("steps + concurrent H2D: %.2f ms/step; copy durations %s" % (ms, ["%.1f" % v for v in c]))
ms, c = run(6, True, 16)
print("steps + concurrent H2D in 16 chunks: %.2f ms/step; copy durations %s" % (ms, ["%.1f" % v for v in c]))
torch.cuda.synchronize()
a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); stage.copy_(hx, non_blocking=True); b_.record(); torch.cuda.synchronize()
print("H2D alone: %.2f ms" % a.elapsed_time(b_))
