"""Runs a few TGANv2-cond training iterations (same setup as bench.py) with the CUDA profiler range around
the last one -- the target of `ncu --profile-from-start off`."""
import argparse
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import train_params
from txt2vid_b200.data import SyntheticVideoCaptions
from txt2vid_b200.factory import build_models
from txt2vid_b200.gan import CondGan, MixedGanLoss, RSGANLoss
from txt2vid_b200.optim import FusedAdam
from txt2vid_b200.trainer import train_iteration

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--convprof", action="store_true", help="time every conv-engine launch (T2V_PROFILE_DUMP=csv)")
a = ap.parse_args()
device = torch.device("cuda", 0)
with contextlib.redirect_stdout(io.StringIO()):
    txt, gen, dis = build_models(True, vocab_size=1000, seed=100)
txt, gen, dis = txt.to(device), gen.to(device), dis.to(device)
gan = CondGan(gen=gen, discrims=[dis], cond_encoder=txt, discrim_names=["video"])
losses = MixedGanLoss(g_loss=RSGANLoss(), d_loss=RSGANLoss())
optD = FusedAdam([{"params": dis.parameters()}], lr=2e-4, betas=(0.5, 0.999))
optG = FusedAdam([{"params": gen.parameters()}], lr=2e-4, betas=(0.5, 0.999))
params = train_params(0.5)
x, t, l = SyntheticVideoCaptions(a.batch, 1).batch(0)
x, t = x.to(device), t.to(device)
for i in range(a.warm):
    train_iteration(gan, x, [t, l], device, optD, optG, params, losses, end2end=False)
torch.cuda.synchronize()
import ctypes
from txt2vid_b200 import _lib
if a.convprof:
    _lib.lib().t2v_profile_enable(1)
else:
    torch.cuda.profiler.start()
ld, lg, _, _, _ = train_iteration(gan, x, [t, l], device, optD, optG, params, losses, end2end=False)
torch.cuda.synchronize()
if a.convprof:
    _lib.lib().t2v_profile_enable(0)
    _lib.lib().t2v_profile_read((ctypes.c_double * 6)())
else:
    torch.cuda.profiler.stop()
print("lossD %.5f lossG %.5f" % (float(ld), float(lg)))
