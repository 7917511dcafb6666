"""CPU: host-side pieces added around the kernels -- checkpoint compatibility with the reference, the
position-pair identity behind the small-channel weight gradients, lagged loss logging, the prefetcher's CPU path."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from helpers import build_product_models

REF = "/root/reference"


def test_checkpoint_roundtrip_through_cond_gan(tmp_path):
    """CondGan.save_dict / load_from_dict (gan/cond_gan.py:186-217): keys gen / cond / <discriminator name>, the
    discriminator's parameters under `single_discrim.module.*`; torch.save -> torch.load -> fresh modules, exact."""
    from txt2vid_b200.gan import CondGan
    txt, gen, dis = build_product_models(True, V=50, seed=3)
    gan = CondGan(gen=gen, discrims=[dis], cond_encoder=txt, discrim_names=["video"])
    path = str(tmp_path / "iter_1")
    torch.save(gan.save_dict(), path)
    loaded = torch.load(path)
    assert set(loaded) == {"gen", "cond", "video"}
    assert any(k.startswith("single_discrim.module.") for k in loaded["video"])
    txt2, gen2, dis2 = build_product_models(True, V=50, seed=4)
    gan2 = CondGan(gen=gen2, discrims=[dis2], cond_encoder=txt2, discrim_names=["video"])
    gan2.load_from_dict(loaded)
    for a, b in ((gen, gen2), (dis, dis2), (txt, txt2)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(torch.equal(sa[k], sb[k]) for k in sa)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the build container")
def test_state_dicts_interchange_with_the_reference_modules():
    """A reference-trained checkpoint loads into the product modules and vice versa (strict): same keys, same shapes
    (models/tganv2_cond/{gen,discrim}.py, models/txt/basic.py)."""
    # import the REFERENCE's txt2vid package in isolation from this repo's namespace package of the same name
    stash = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "txt2vid" or k.startswith("txt2vid.")}
    sys.path.insert(0, REF)
    try:
        from txt2vid.models.tganv2_cond.gen import MultiScaleGen as RefGen
        from txt2vid.models.tganv2_cond.discrim import MultiScaleDiscrim as RefDis
        from txt2vid.models.txt.basic import Seq2Seq as RefTxt
        assert sys.modules["txt2vid.models.tganv2_cond.gen"].__file__.startswith(REF)
    except ImportError as e:
        pytest.skip("reference modules not importable here: %r" % (e,))
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "txt2vid" or k.startswith("txt2vid.")]:
            del sys.modules[k]
        sys.modules.update(stash)
    txt, gen, dis = build_product_models(True, V=50, seed=3)
    rgen, rdis, rtxt = RefGen(width=64, height=64, cond_dim=256), RefDis(cond_dim=256), RefTxt(vocab_size=50)
    for prod, ref in ((gen, rgen), (dis, rdis), (txt, rtxt)):
        ref.load_state_dict(prod.state_dict(), strict=True)            # product checkpoint -> reference
        prod.load_state_dict(ref.state_dict(), strict=True)            # reference checkpoint -> product
        sp, sr = prod.state_dict(), ref.state_dict()
        assert list(sp) == list(sr)
        assert all(sp[k].shape == sr[k].shape and torch.equal(sp[k].float(), sr[k].float()) for k in sp)


def test_position_pair_identity_of_small_channel_weight_gradients():
    """kernels.conv_wgrad runs 32-channel (1,3,3) weight gradients on PAIRS of adjacent w voxels (64-channel views)
    and folds dw2[(pw,co)][a_h, s][(qw,ci)] into the real taps: shift a_w - 1 = 2 s + qw - pw (t2v_wgrad_fold_pairs).
    Here the same algebra in plain torch against the direct weight gradient."""
    g = torch.Generator().manual_seed(0)
    N, H, W, Ci, Co = 3, 6, 8, 4, 5
    x = torch.randn(N, H, W, Ci, generator=g, dtype=torch.float64)
    dy = torch.randn(N, H, W, Co, generator=g, dtype=torch.float64)
    direct = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (Co, Ci, 3, 3), dy.permute(0, 3, 1, 2), padding=1)
    x2, dy2 = x.reshape(N, H, W // 2, 2 * Ci), dy.reshape(N, H, W // 2, 2 * Co)
    dw2 = torch.nn.grad.conv2d_weight(x2.permute(0, 3, 1, 2), (2 * Co, 2 * Ci, 3, 3), dy2.permute(0, 3, 1, 2), padding=1)
    folded = torch.zeros_like(direct)
    for a_w in range(3):
        for s in (-1, 0, 1):
            for pw in (0, 1):
                for qw in (0, 1):
                    if 2 * s + qw - pw == a_w - 1:
                        folded[:, :, :, a_w] += dw2[pw * Co:(pw + 1) * Co, qw * Ci:(qw + 1) * Ci, :, s + 1]
    assert torch.allclose(folded, direct, rtol=1e-12, atol=1e-12)


def test_lagged_losses_cpu_path_is_immediate():
    from txt2vid_b200.trainer import LaggedLosses
    seen = []
    ll = LaggedLosses(lambda d, g: seen.append((d, g)), lag=2, device="cpu")
    ll.push(torch.tensor(1.5), torch.tensor(2.5))
    assert seen == [(1.5, 2.5)]
    ll.drain()
    assert seen == [(1.5, 2.5)]


def test_prefetcher_cpu_path_and_deferred_preload():
    """data_prefetcher on a CPU device is a plain iterator; next(preload=False) + preload() deliver the same batches."""
    from txt2vid_b200.data import data_prefetcher
    batches = [(torch.full((2, 3), float(i)), torch.full((2, 4), i, dtype=torch.long), [4, 4]) for i in range(5)]
    for deferred in (False, True):
        pf = data_prefetcher(iter(batches), device="cpu")
        got = []
        x, y = pf.next(preload=not deferred)
        while x is not None:
            got.append((float(x[0, 0]), int(y[0][0, 0]), y[1]))
            if deferred:
                pf.preload()
            x, y = pf.next(preload=not deferred)
        assert got == [(float(i), i, [4, 4]) for i in range(5)]


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the build container")
def test_caption_pretraining_step_matches_the_reference(monkeypatch):
    """one optimisation step of train/txt.py:166-181 (encode -> teacher-forced decode -> cross entropy
    -> Adam) on identical weights and sentences: same loss, same decoded symbols, same updated weights.  The product's
    embedding / LSTM recurrence / GEMM kernels are replaced by their executable spec (tests/cpu_kernels.py, fp32
    storage): this checks the host logic and the autograd formulas of text.py against the LIVE reference."""
    import cpu_kernels
    from txt2vid_b200 import ops
    monkeypatch.setattr(ops, "K", cpu_kernels)
    cpu_kernels.set_store_dtype(torch.float32)
    ops.PACKS.clear()
    try:
        _caption_pretraining_body()
    finally:
        cpu_kernels.set_store_dtype(torch.bfloat16)
        ops.PACKS.clear()


def _caption_pretraining_body():
    stash = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "txt2vid" or k.startswith("txt2vid.")}
    sys.path.insert(0, REF)
    try:
        from txt2vid.models.txt.basic import Seq2Seq as RefTxt
        assert sys.modules["txt2vid.models.txt.basic"].__file__.startswith(REF)
    except ImportError as e:
        pytest.skip("reference modules not importable here: %r" % (e,))
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "txt2vid" or k.startswith("txt2vid.")]:
            del sys.modules[k]
        sys.modules.update(stash)
    import contextlib, io
    from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
    from txt2vid_b200.text import Seq2Seq
    from txt2vid_b200.train_txt import collate_fn, pretrain_step
    V = 40
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefTxt(vocab_size=V)
    prod = Seq2Seq(vocab_size=V)
    prod.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(1)
    sents = []
    for L in (9, 7, 7, 4):
        s = torch.randint(4, V, (L,), generator=g).float()
        s[0], s[-1] = 1, 2
        sents.append(s)
    sent, lengths = collate_fn(list(sents))
    for tf in (True, False):
        opt_r = torch.optim.Adam(ref.parameters(), lr=1e-3)
        opt_p = torch.optim.Adam(prod.parameters(), lr=1e-3)
        # the reference's loop body, verbatim order (train/txt.py:166-181)
        ref.zero_grad()
        _, hid, _ = ref.encode(sent, lengths=lengths)
        targets, _ = pad_packed_sequence(pack_padded_sequence(sent, lengths, batch_first=True), batch_first=True,
                                         total_length=lengths[0])
        dec, sym = ref.decode(true_inputs=sent, initial_hidden=hid, max_seq_len=lengths[0], teacher_force=tf)
        loss_r = torch.nn.CrossEntropyLoss()(dec.permute(0, 2, 1), targets)
        loss_r.backward()
        loss_p, sym_p = pretrain_step(prod, sent, lengths, None, teacher_force=tf)
        assert abs(float(loss_p) - float(loss_r)) <= 2e-6 * abs(float(loss_r)), (float(loss_p), float(loss_r))
        assert torch.equal(sym_p, sym)
        # gradients of every parameter (the kernels sum in a different order than ATen: 1e-4 relative; Adam's first
        # step is sign-like, so the updated weights themselves are compared on the tensors' scale)
        for (n, a), (_, b) in zip(prod.named_parameters(), ref.named_parameters()):
            assert a.grad is not None and b.grad is not None, n
            err = float((a.grad - b.grad).norm() / (b.grad.norm() + 1e-12))
            assert err < 1e-4, (n, err)
        opt_r.step()
        opt_p.step()
        for (n, a), (_, b) in zip(prod.state_dict().items(), ref.state_dict().items()):
            assert float((a - b).abs().max()) <= 2.1e-3, n          # at most one sign flip of a 1e-3 Adam step
        prod.load_state_dict(ref.state_dict(), strict=True)


def test_caption_pretraining_golden_fixture_through_the_executable_spec(monkeypatch):
    """The committed fixture of oracle/make_golden_caption_pretrain.py (recorded from the live reference) against the
    product's host logic + autograd formulas on the CPU spec kernels (fp32): the same check the GPU test
    test_round2_gpu.py::test_caption_pretraining_step_vs_reference_golden runs through the real kernels."""
    import cpu_kernels
    from txt2vid_b200 import ops
    monkeypatch.setattr(ops, "K", cpu_kernels)
    cpu_kernels.set_store_dtype(torch.float32)
    ops.PACKS.clear()
    try:
        from test_round2_gpu import run_caption_pretrain_against_golden
        rep = run_caption_pretrain_against_golden("cpu", 1e-5, 1e-4, 1e-3)
        assert rep["teacher_force"]["loss"] < 1e-5
    finally:
        cpu_kernels.set_store_dtype(torch.bfloat16)
        ops.PACKS.clear()


@pytest.fixture()
def emulated_fp32(monkeypatch):
    import cpu_kernels
    from txt2vid_b200 import ops, optim, trainer
    for mod in (ops, optim, trainer):
        monkeypatch.setattr(mod, "K", cpu_kernels)
    monkeypatch.setattr(ops, "BF16", torch.float32)
    cpu_kernels.set_store_dtype(torch.float32)
    ops.PACKS.clear()
    yield
    cpu_kernels.set_store_dtype(torch.bfloat16)
    ops.PACKS.clear()


def test_trainer_test_writes_the_reference_sample_files(tmp_path, emulated_fp32):
    """trainer.test() (gan/trainer.py:44-90) -- eval-mode generator, ONE full-resolution clip per
    sample, the reference's file names: real_<i>.png, sentences_<i>_<j>.txt (vocab.to_words of the token rows),
    <h>x<w>_<i>_<j>.jpg.  Host logic on the kernels' executable spec."""
    from types import SimpleNamespace
    from txt2vid_b200.data import build_vocab, collate_fn
    from txt2vid_b200.gan import CondGan
    from txt2vid_b200.trainer import test as run_test
    sents = ["digit 3 is left and right.", "digit 7 is top and bottom."]
    vocab = build_vocab(sents)
    txt, gen, dis = build_product_models(True, V=len(vocab), seed=5)
    gan = CondGan(gen=gen, discrims=[dis], cond_encoder=txt, discrim_names=["video"])
    g = torch.Generator().manual_seed(0)
    vids = [torch.rand(16, 3, 64, 64, generator=g) * 2 - 1 for _ in sents]          # loader order (T, C, H, W)
    batch = collate_fn([(v, vocab.encode(s)) for v, s in zip(vids, sents)])
    params = SimpleNamespace(out_samples=str(tmp_path / "samples"), img_model=False)
    run_test(gan=gan, num_samples=2, dataset=[batch], device=torch.device("cpu"), params=params, vocab=vocab)
    assert not gan.gen.training
    names = sorted(os.listdir(params.out_samples))
    assert names == ["64x64_0_0.jpg", "64x64_1_0.jpg", "real_0.png", "real_1.png", "sentences_0_0.txt",
                     "sentences_1_0.txt"], names
    with open(os.path.join(params.out_samples, "sentences_0_0.txt")) as f:
        lines = f.read().splitlines()
    assert lines == ["<start> digit 3 is left and right<end>", "<start> digit 7 is top and bottom<end>"]
    from PIL import Image
    im = Image.open(os.path.join(params.out_samples, "64x64_0_0.jpg"))
    assert im.size == (16 * 64 + 17 * 2, 2 * 64 + 3 * 2)                             # nrow = 16 frames, 2 clips, padding 2


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the build container")
def test_pickled_reference_sentence_model_loads_into_the_product(tmp_path, emulated_fp32):
    """train/txt.py:190-191 saves the WHOLE caption model with torch.save(model) and train/gan.py:38-42
    loads that pickle (`--sent_weights`).  A pickle written by the live reference (subprocess with /root/reference on
    the path) must unpickle under this repo's `txt2vid` namespace into the product's Seq2Seq and encode identically."""
    import subprocess
    path, out = str(tmp_path / "sent.pth"), str(tmp_path / "ref_out.pt")
    code = ("import sys, torch; sys.path.insert(0, %r); torch.manual_seed(7)\n"
            "from txt2vid.models.txt.basic import Seq2Seq\n"
            "m = Seq2Seq(vocab_size=40)\n"
            "tok = torch.tensor([[1, 5, 6, 7, 2], [1, 9, 2, 0, 0]]); lens = [5, 3]\n"
            "o, (h, c), hn = m.encode(tok, lens)\n"
            "torch.save(m, %r); torch.save({'tok': tok, 'lens': lens, 'out': o, 'h': h, 'c': c, 'hn': hn}, %r)\n"
            % (REF, path, out))
    subprocess.check_call([sys.executable, "-c", code], cwd=str(tmp_path))
    import txt2vid_b200.text as T
    m = torch.load(path, weights_only=False)
    assert type(m) is T.Seq2Seq and type(m.encoder) is T.RecurrentModel
    ref = torch.load(out)
    o, (h, c), hn = m.encode(ref["tok"], ref["lens"])
    for a, b in ((o, ref["out"]), (h, ref["h"]), (c, ref["c"]), (hn, ref["hn"])):
        assert a.shape == b.shape and float((a.float() - b).abs().max()) < 1e-5
