"""the on-device synthetic moving-MNIST generator (txt2vid_b200/data.MovingDigits, csrc/synth.cu) against
the fixture recorded from the LIVE reference generator (oracle/make_golden_moving_digits.py;
txt2vid/data/synthetic/generate.py:18-47,59-182): same random decisions, same sentences, bit-identical frames, and the
token rows the reference's Vocab / collate_fn produce for those sentences."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def fixture():
    with open(os.path.join(HERE, "golden", "moving_digits.json")) as f:
        return json.load(f)


def digit_bank(seed, per_class):
    rng = np.random.RandomState(seed)
    return {c: np.stack([rng.randint(0, 256, (28, 28)).astype(np.uint8) for _ in range(per_class)]) for c in range(10)}


def make(fx, device, **kw):
    from txt2vid_b200.data import MovingDigits
    return MovingDigits(len(fx["examples"]), 1, digit_bank(fx["bank_seed"], fx["per_class"]), frames=fx["T"],
                        size=fx["W"], device=device, seed=fx["seed"], **kw)


def test_host_draws_and_positions_reproduce_the_reference_generator():
    """decisions in the reference's RNG order + generate_frames' position walk, rendered by a numpy paste"""
    fx = fixture()
    md = make(fx, "cpu")
    bank = md.bank.numpy()
    for e in fx["examples"]:
        d = md.draw()
        assert md.sentence(d) == e["sentence"]
        pos = md.frame_positions(d["a"], d["b"], fx["T"], d["anim"])
        clip = np.zeros((fx["T"], fx["H"], fx["W"], 3), dtype=np.uint8)
        for t, (x, y) in enumerate(pos):
            clip[t, y:y + 28, x:x + 28, :] = bank[d["digit"]][:, :, None]
        assert hashlib.sha256(clip.tobytes()).hexdigest() == e["sha256"], e["sentence"]


def test_grammar_sentences_cover_the_vocabulary_and_encode_to_eight_tokens():
    from txt2vid_b200.data import MovingDigits, build_vocab
    sents = MovingDigits.all_sentences()
    v = build_vocab(sents)
    assert len(v) == 4 + 17 and len(sents) == 40          # 21 words incl. the 4 specials
    for s in sents:
        t = v.encode(s)
        assert t.numel() == 8 and int(t[0]) == v(v.START) and int(t[-1]) == v(v.END)


@pytest.mark.gpu
def test_moving_digits_kernels_are_bit_exact_vs_the_reference_fixture():
    fx = fixture()
    from txt2vid_b200 import _lib
    md = make(fx, "cuda", as_uint8=True)
    n0 = _lib.lib().t2v_launch_count()
    draws = [md.draw() for _ in fx["examples"]]
    clips, tokens, lengths = md.batch(draws)
    assert _lib.lib().t2v_launch_count() - n0 == 2
    assert clips.dtype == torch.uint8 and tuple(clips.shape) == (len(draws), fx["T"], 3, fx["H"], fx["W"])
    host = clips.permute(0, 1, 3, 4, 2).contiguous().cpu().numpy()             # (B, T, H, W, 3) like the recorded frames
    for i, e in enumerate(fx["examples"]):
        assert hashlib.sha256(host[i].tobytes()).hexdigest() == e["sha256"], e["sentence"]
        assert torch.equal(tokens[i].cpu(), md.vocab.encode(e["sentence"]))
    assert lengths == [8] * len(draws)
    # fp32 output = transforms.ToTensor() + Normalize(0.5, 0.5) of those frames, same IEEE operations (CPU torch)
    md32 = make(fx, "cuda")
    clips32, _, _ = md32.batch(draws)
    want = clips.cpu().float().div(255.0).sub(0.5).div(0.5)
    assert clips32.dtype == torch.float32 and torch.equal(clips32.cpu(), want)
    # training-order layout
    from txt2vid_b200 import kernels as K
    pos = torch.tensor([md.frame_positions(d["a"], d["b"], fx["T"], d["anim"]) for d in draws], dtype=torch.int32).cuda()
    dig = torch.tensor([d["digit"] for d in draws], dtype=torch.int32).cuda()
    alt = K.moving_digits(md.bank, dig, pos, fx["T"], fx["H"], fx["W"], out_f32=False, layout=1)
    assert torch.equal(alt, clips.permute(0, 2, 1, 3, 4))


@pytest.mark.gpu
def test_u8_normalize_kernel_is_bit_exact():
    from txt2vid_b200 import kernels as K
    g = torch.Generator().manual_seed(3)
    for n in (7, 256, 4099, 3 * 16 * 64 * 64):
        u = torch.randint(0, 256, (n,), generator=g, dtype=torch.uint8)
        assert torch.equal(K.u8_normalize(u.cuda()).cpu(), u.float().div(255.0).sub(0.5).div(0.5))
    u = torch.arange(256, dtype=torch.uint8)
    assert torch.equal(K.u8_normalize(u.cuda()).cpu(), u.float().div(255.0).sub(0.5).div(0.5))


@pytest.mark.gpu
def test_moving_digits_feed_the_public_train_loop_through_the_prefetcher():
    """data_prefetcher hands resident batches through: next() returns the generator's tensors and [tokens, lengths]"""
    fx = fixture()
    from txt2vid_b200.data import data_prefetcher
    md = make(fx, "cuda")
    md.n = 3
    pf = data_prefetcher(md, device="cuda")
    seen = 0
    while True:
        x, y = pf.next()
        if x is None:
            break
        assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape[1:]) == (fx["T"], 3, fx["H"], fx["W"])
        assert y[0].is_cuda and tuple(y[0].shape) == (x.shape[0], 8) and y[1] == [8] * x.shape[0]
        assert float(x.min()) == -1.0 and float(x.max()) <= 1.0
        seen += 1
    assert seen == 3
    # frames handed out as stored (uint8, resident): the prefetcher normalises them on its side stream, ordered behind
    # the generator's kernels -- same values as the generator's own fp32 output for the same draws
    md8, md32 = make(fx, "cuda", as_uint8=True), make(fx, "cuda")
    md8.n = 2
    draws = []
    real_draw = md8.draw
    md8.draw = lambda: draws.append(real_draw()) or draws[-1]
    pf = data_prefetcher(md8, device="cuda")
    B = len(fx["examples"])
    for k in range(2):
        x, _ = pf.next()
        torch.cuda.synchronize()
        want, _, _ = md32.batch(draws[k * B:(k + 1) * B])
        assert x.dtype == torch.float32 and torch.equal(x, want)


@pytest.mark.gpu
def test_cli_entry_point_trains_on_device_generated_clips(tmp_path, capsys):
    """The reference's launch line (scripts/run_tganv2_cond.sh:20) through txt2vid.train.gan.main -- reflection on the
    dotted class paths, xavier init, RSGAN + gradient penalty, fused Adam, CUDA-graph iterations -- fed by the
    on-device moving-MNIST dataset named by a `--data` json: checkpoints and samples appear, losses are logged."""
    import json
    import pickle
    from txt2vid.train.gan import build_parser, main
    from txt2vid_b200.data import MovingDigits, build_vocab
    vocab_path, data_path = tmp_path / "vocab.pickle", tmp_path / "data.json"
    with open(vocab_path, "wb") as f:
        pickle.dump(build_vocab(MovingDigits.all_sentences()), f)
    with open(data_path, "w") as f:
        json.dump({"class": "txt2vid.data.MovingDigitsDataset",
                   "args": {"batch_size": 8, "num_batches": 6, "seed": 300}}, f)
    out, samples = tmp_path / "out", tmp_path / "samples"
    argv = ["--cuda", "--seed", "100", "--G", "txt2vid.models.tganv2_cond.gen.MultiScaleGen",
            "--D", "txt2vid.models.tganv2_cond.discrim.MultiScaleDiscrim", "--sent", "txt2vid.models.txt.basic.Seq2Seq",
            "--D_loss", "txt2vid.gan.losses.RSGANLoss", "--frame_sizes", "8", "16", "32", "64", "--D_names", "video",
            "--G_lr", "0.0002", "--D_lr", "0.0002", "--D_beta1", "0.5", "--D_beta2", ".999", "--G_beta1", "0.5",
            "--G_beta2", ".999", "--init_method", "xavier", "--discrim_steps", "1", "--gp_lambda", ".5",
            "--no_mean_discrim_loss", "--subsample_input", "--data", str(data_path), "--vocab", str(vocab_path),
            "--epochs", "1", "--batch_size", "8", "--log_period", "2", "--save_example_period", "4",
            "--save_model_period", "4", "--out", str(out), "--out_samples", str(samples), "--cuda_graphs"]
    from txt2vid_b200 import _lib
    n0 = _lib.lib().t2v_launch_count()
    main(build_parser().parse_args(argv))
    assert _lib.lib().t2v_launch_count() - n0 > 3 * 1500     # two eager warm-ups + the capture (replays are not counted)
    text = capsys.readouterr().out
    assert "Iter 6, Loss_D:" in text and "Loss_G:" in text
    ckpts = [p for p in out.iterdir() if p.name.startswith("iter_4_")]
    assert len(ckpts) == 1
    saved = torch.load(ckpts[0], weights_only=False)
    assert {"gen", "cond", "video", "optG", "optD"} <= set(saved.keys())
    names = {p.name for p in samples.iterdir()}
    assert "real_samples.png" in names and any(n.startswith("fake_samples_epoch_000_iter_000004_64x64") for n in names)
    assert any(n.startswith("sentences_epoch000_iter_000004") for n in names)
    words = (samples / "sentences_epoch000_iter_000004.txt").read_text().splitlines()
    assert len(words) == 8 and all(w.startswith("<start> digit ") for w in words)
