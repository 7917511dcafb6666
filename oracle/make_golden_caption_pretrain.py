"""SURVEY 8(f4) fixture: one caption auto-encoder pre-training step of the LIVE reference (txt2vid/train/txt.py:166-181:
encode -> teacher-forced / greedy decode from the encoder's final state -> cross entropy), recorded so that the GPU box
(which has no /root/reference) can check the product's kernels against it.

    python oracle/make_golden_caption_pretrain.py      # build container only; writes tests/golden/caption_pretrain.json

Test infrastructure only.  Model: txt2vid.models.txt.basic.Seq2Seq(vocab_size=V) built under torch.manual_seed(5) (the
product's constructor consumes the RNG identically; the fixture carries weight checksums to prove it).  Recorded per
mode (teacher forcing on / off): the loss, the decoded symbols, every parameter's gradient norm and a 64-entry probe of
every gradient tensor."""
import contextlib
import io
import json
import os
import sys

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
V = 40
LENGTHS = (9, 7, 7, 4, 4, 3)


def sentences():
    g = torch.Generator().manual_seed(1)
    sents = []
    for L in LENGTHS:
        s = torch.randint(4, V, (L,), generator=g).float()
        s[0], s[-1] = 1, 2
        sents.append(s)
    return sents


def collate(data):                                   # train/txt.py:44-52
    data.sort(key=lambda x: len(x), reverse=True)
    lengths = [len(s) for s in data]
    targets = torch.zeros(len(data), max(lengths)).long()
    for i, s in enumerate(data):
        targets[i, :lengths[i]] = s[:lengths[i]]
    return targets, lengths


def probe_index(numel):
    n = min(64, numel)
    return [(i * 2654435761) % numel for i in range(n)]


def main():
    sys.path.insert(0, REF)
    from txt2vid.models.txt.basic import Seq2Seq
    from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = Seq2Seq(vocab_size=V)
    sent, lengths = collate(sentences())
    fx = {"source": "txt2vid/train/txt.py:166-181 + models/txt/basic.py (live reference, torch %s CPU fp32)" % torch.__version__,
          "V": V, "seed": 5, "sent": sent.tolist(), "lengths": lengths,
          "weights": {n: [float(p.double().sum()), float(p.double().abs().sum())] for n, p in ref.state_dict().items()},
          "modes": {}}
    for tf in (True, False):
        ref.zero_grad()
        _, hid, _ = ref.encode(sent, lengths=lengths)
        targets, _ = pad_packed_sequence(pack_padded_sequence(sent, lengths, batch_first=True), batch_first=True,
                                         total_length=lengths[0])
        dec, sym = ref.decode(true_inputs=sent, initial_hidden=hid, max_seq_len=lengths[0], teacher_force=tf)
        loss = torch.nn.CrossEntropyLoss()(dec.permute(0, 2, 1), targets)
        loss.backward()
        grads = {}
        for n, p in ref.named_parameters():
            g = p.grad.detach().reshape(-1)
            grads[n] = {"norm": float(g.double().norm()), "probe": [float(g[i]) for i in probe_index(g.numel())]}
        fx["modes"]["teacher_force" if tf else "greedy"] = {"loss": float(loss), "symbols": sym.tolist(), "grads": grads,
                                                            "logits_sum": float(dec.double().sum()),
                                                            "logits_abs": float(dec.double().abs().sum())}
    out = os.path.join(HERE, "..", "tests", "golden", "caption_pretrain.json")
    with open(out, "w") as f:
        json.dump(fx, f)
    print("wrote", os.path.normpath(out), {k: v["loss"] for k, v in fx["modes"].items()})


if __name__ == "__main__":
    main()
