"""Shared test helpers: build the product models in the reference's construction order, compare with the
oracle, golden-fixture access."""
import contextlib
import io
import json
import os
import random
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def checksum(t):
    t = t.detach().double().cpu().reshape(-1)
    idx = torch.arange(t.numel(), dtype=torch.float64)
    return {"sum": float(t.sum()), "abs": float(t.abs().sum()), "wsum": float((t * ((idx % 97) + 1)).sum()),
            "n": int(t.numel()), "first": [float(v) for v in t[:4]]}


def build_product_models(conditional, V=1000, seed=100, width=64, height=64, num_frames=16):
    """Construction + init order of txt2vid/train/gan.py:28-70, on the product classes."""
    from txt2vid_b200.factory import build_models
    return build_models(conditional, vocab_size=V, seed=seed, width=width, height=height, num_frames=num_frames)


def synth_batch(B, V, T=16, S=64, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, T, S, S, generator=g) * 2 - 1
    lengths = sorted([int(v) for v in torch.randint(4, 21, (B,), generator=g)], reverse=True)
    tokens = torch.zeros(B, lengths[0], dtype=torch.long)
    for b, L in enumerate(lengths):
        tokens[b, 0] = 1
        tokens[b, 1:L - 1] = torch.randint(4, V, (L - 2,), generator=g)
        tokens[b, L - 1] = 2
    return x, tokens, lengths


def train_params(gp_lambda=0.5, frame_sizes=(8, 16, 32, 64)):
    return SimpleNamespace(data_is_imgs=False, img_model=False, frame_sizes=list(frame_sizes), subsample_input=True,
                           discrim_steps=1, gen_steps=1, gp_lambda=gp_lambda, no_mean_discrim_loss=False,
                           no_mean_gen_loss=True)


def l2rel(a, b, zero_tol=1e-5):
    """Per-tensor relative L2 deviation; numerically-zero reference tensors must stay ~zero."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    nb = float(b.norm())
    if nb < zero_tol:
        return 0.0 if float(a.norm()) < 50 * max(zero_tol, nb) else float("inf")
    return float((a - b).norm()) / nb


def state_to_cpu(module):
    return {k: v.detach().float().cpu().clone() if v.dtype.is_floating_point else v.detach().cpu().clone()
            for k, v in module.state_dict().items()}


def bf16_bars(tag, factor=1.5):
    """Bars of the bf16-mode full-iteration parity tests: `factor` x the measured bf16 floor of this network --
    the ORACLE itself under torch.autocast(bfloat16) (cuDNN / cuBLAS bf16 kernels, fp32 accumulation) against the same
    oracle in fp32 (TF32 off) on the same weights, inputs and host-RNG stream, measured on a B200 by
    scripts/bf16_floor.py and committed as profiles/r02_bf16_floor_<tag>.json.  Losses keep BASELINE north_star's own
    2e-2 bar (they sit far inside it); gradients and generated clips of this un-normalised random-init network are
    bf16-noise amplified (the floor shows by how much), so their bar is the floor, not a free parameter.
    -> (loss_tol, gradG_l2_tol, gradG_cos_min, fake_tol, gradD_l2_tol, gradD_cos_min)"""
    import json
    with open(os.path.join(ROOT, "profiles", "r02_bf16_floor_%s.json" % tag)) as f:
        fl = json.load(f)
    cos_bar = lambda c: 1.0 - factor * factor * (1.0 - c)          # 1 - cos ~ err^2 / 2
    return (2e-2, factor * fl["gradG"]["l2"], cos_bar(fl["gradG"]["cos"]), factor * fl["fake"],
            factor * fl["gradD"]["l2"], cos_bar(fl["gradD"]["cos"]))


def fp32_bars():
    """Bars of the fp32-mode parity tests: BASELINE north_star's 1e-3 on losses, generated clips and gradients.
    -> (loss_tol, gradG_l2_tol, gradG_cos_min, fake_tol, gradD_l2_tol, gradD_cos_min)"""
    return (1e-3, 1e-3, 0.999999, 1e-3, 1e-3, 0.999999)
