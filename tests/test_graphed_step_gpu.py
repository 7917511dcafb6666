"""GPU: the CUDA-graph replay of the training iteration (trainer.GraphedTrainStep) must follow the eager
iteration: same seeds -> same host-RNG draws -> same losses (up to fp32 atomics order / bf16 noise), and the
weights must keep training across replays.  Also checks tcgen05 wgrad on < 64-channel layers (TMA zero fill)."""
import contextlib
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(B, seed=100):
    from bench import train_params
    from txt2vid_b200 import ops
    from txt2vid_b200.data import SyntheticVideoCaptions
    from txt2vid_b200.factory import build_models
    from txt2vid_b200.gan import CondGan, MixedGanLoss, RSGANLoss
    from txt2vid_b200.optim import FusedAdam
    ops.PACKS.clear()
    dev = torch.device("cuda", 0)
    with contextlib.redirect_stdout(io.StringIO()):
        txt, gen, dis = build_models(True, vocab_size=1000, seed=seed)
    txt, gen, dis = txt.to(dev), gen.to(dev), dis.to(dev)
    torch.manual_seed(7)
    torch.cuda.manual_seed(7)
    np.random.seed(7)
    gan = CondGan(gen=gen, discrims=[dis], cond_encoder=txt, discrim_names=["video"])
    losses = MixedGanLoss(g_loss=RSGANLoss(), d_loss=RSGANLoss())
    optD = FusedAdam([{"params": dis.parameters()}], lr=2e-4, betas=(0.5, 0.999))
    optG = FusedAdam([{"params": gen.parameters()}], lr=2e-4, betas=(0.5, 0.999))
    data = SyntheticVideoCaptions(B, 6, vocab_size=1000)
    batches = [(x.to(dev), t.to(dev), l) for x, t, l in (data.batch(i) for i in range(6))]
    return dev, gan, losses, optD, optG, train_params(0.5), batches


def test_graph_replay_matches_eager():
    from txt2vid_b200.trainer import GraphedTrainStep, train_iteration
    B = 8
    dev, gan, losses, optD, optG, params, batches = _setup(B)
    eager = []
    for x, t, l in batches:
        ld, lg, _, _, _ = train_iteration(gan, x, [t, l], dev, optD, optG, params, losses, end2end=False)
        eager.append((float(ld), float(lg)))
    dev, gan, losses, optD, optG, params, batches = _setup(B)
    step = GraphedTrainStep(gan, optD, optG, params, losses, dev, warmup=2)
    graphed = []
    for x, t, l in batches:
        ld, lg = step(x, [t, l])
        graphed.append((float(ld), float(lg)))
    print("eager  ", eager)
    print("graphed", graphed)
    assert step.graphs is not None
    # The two runs differ by the order of fp32 atomics only, but GAN training amplifies that: Adam's first steps move
    # every weight by ~lr whatever the gradient's size, so a near-zero gradient whose sign flips changes the
    # trajectory.  Tight for the first iterations, loose afterwards.
    for i, ((a, b), (c, d)) in enumerate(zip(eager, graphed)):
        tol = 3e-2 if i < 3 else 1.5e-1
        assert abs(a - c) < tol * max(1.0, abs(a)) and abs(b - d) < tol * max(1.0, abs(b)), (i, eager, graphed)
    # the losses move (training is happening) and stay finite
    assert all(np.isfinite(v) for pair in graphed for v in pair)
    assert len({round(p[0], 4) for p in graphed}) > 2


@pytest.mark.parametrize("case", [(8, 1, 16, 16, 32, 48, (1, 3, 3)), (4, 8, 8, 8, 16, 64, (3, 3, 3)),
                                  (6, 1, 32, 32, 32, 16, (1, 3, 3)), (16, 1, 8, 8, 128, 16, (1, 1, 1))])
def test_tc_wgrad_small_channels(case):
    import torch.nn.functional as F
    from txt2vid_b200 import kernels as K
    N, D, H, W, Cin, Cout, k = case
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((N, D, H, W, Cin), device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn((N, D, H, W, Cout), device="cuda", generator=g).to(torch.bfloat16)
    dw = K.conv_wgrad(dy, x, k=k, algo=1)
    ref = K.conv_wgrad(dy, x, k=k, algo=2)
    err = float((dw - ref).abs().max() / ref.abs().max())
    print(case, err)
    assert err < 1e-4


def test_public_train_loop_with_graphs(tmp_path, capsys):
    """trainer.train (gan/trainer.py:111-333) end to end on synthetic batches: pinned host data through data_prefetcher,
    CUDA-graph iterations, lagged loss logging.  Every iteration's losses must reach the rolling averages (finite,
    moving) and the log lines must be printed."""
    from txt2vid_b200 import trainer as T
    B = 8
    dev, gan, losses, optD, optG, params, _ = _setup(B)
    seen = []
    real_update = T.RollingAvg.update

    class Data(object):
        def __init__(self, n):
            from txt2vid_b200.data import SyntheticVideoCaptions
            self.src = SyntheticVideoCaptions(B, n, vocab_size=1000)

        def __len__(self):
            return len(self.src)

        def __iter__(self):
            return iter(self.src)

    for k, v in dict(batch_size=B, sample_batch_size=B, out=str(tmp_path / "out"), out_samples=str(tmp_path / "samples"),
                     loss_window_size=4, log_period=3, save_initial=False, save_initial_examples=False,
                     save_example_period=10 ** 9, save_model_period=10 ** 9, cuda_graphs=True, debug=False).items():
        setattr(params, k, v)
    orig = T.LaggedLosses._deliver

    def spy(self, k):
        self.events[k].synchronize()
        seen.append((float(self.slots[k][0]), float(self.slots[k][1])))
        return orig(self, k)
    T.LaggedLosses._deliver = spy
    try:
        T.train(gan=gan, num_epoch=1, dataset=Data(7), device=dev, optD=optD, optG=optG, params=params, vocab=None,
                losses=losses, channel_first=True, end2end=False)
    finally:
        T.LaggedLosses._deliver = orig
    out = capsys.readouterr().out
    # 2 eager warm-up iterations push through the same path; all 7 iterations are delivered exactly once
    assert len(seen) == 7, seen
    assert all(np.isfinite(a) and np.isfinite(b) for a, b in seen)
    assert len({round(a, 4) for a, _ in seen}) > 3
    assert "Iter 3" in out and "Iter 6" in out


def test_uint8_frames_are_normalised_into_the_graph_input():
    """Frames handed out as stored (data_prefetcher(normalize=False)): eager warm-ups and the capture normalise them
    with t2v_u8_normalize, replays fuse ToTensor + Normalize(0.5, 0.5) (data/__init__.py:362-364) into the fill of the
    graph's static input -- bit-identical to the CPU transform."""
    from txt2vid_b200.data import SyntheticVideoCaptions, data_prefetcher
    from txt2vid_b200.trainer import GraphedTrainStep
    B = 8
    dev, gan, losses, optD, optG, params, _ = _setup(B)
    step = GraphedTrainStep(gan, optD, optG, params, losses, dev, warmup=2)
    pf = data_prefetcher(SyntheticVideoCaptions(B, 5, vocab_size=1000, as_uint8=True), device=dev, normalize=False)
    last, seen = None, []
    while True:
        x, y = pf.next()
        if x is None:
            break
        assert x.dtype == torch.uint8 and x.is_cuda
        ld, lg = step(x, y)
        seen.append((float(ld), float(lg)))
        last = x.clone()
    assert step.graphs is not None and step.replays >= 3 and all(np.isfinite(v) for p in seen for v in p)
    torch.cuda.synchronize()
    want = last.cpu().float().div(255.0).sub(0.5).div(0.5)
    assert step.static_x.dtype == torch.float32 and torch.equal(step.static_x.cpu(), want)
