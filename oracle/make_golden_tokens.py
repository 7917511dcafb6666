"""Row T fixtures (token indexing) from the LIVE reference's own Vocab / build_vocab / collate_fn
(txt2vid/data/__init__.py:260-355) and the caption part of Dataset.__getitem__ (:250-254).

    python oracle/make_golden_tokens.py        # build container only; writes tests/golden/token_fixtures.json

txt2vid.data imports nvidia.dali at module level (data/__init__.py:16-18), which is absent here; the three DALI
modules are stubbed in sys.modules (nothing on the token path touches them).  Sentences: the synthetic moving-MNIST
grammar (data/synthetic/generate.py:102-182) plus hand-written edge cases (upper case, inner full stops, a missing
final full stop, unknown words, repeated spaces)."""
import json
import os
import random
import sys
import types

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference_data():
    for name in ("nvidia", "nvidia.dali", "nvidia.dali.pipeline", "nvidia.dali.ops", "nvidia.dali.types"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["nvidia.dali.pipeline"].Pipeline = object
    sys.path.insert(0, REF)
    import txt2vid.data as D
    return D


def sentences():
    digits = list(range(10))
    moves = ["left and right", "right and left", "top and bottom", "bottom and top"]
    out = ["digit %d is %s." % (d, m) for d in digits for m in moves]
    out += ["A man is Playing a guitar.", "two dogs run. then they stop", "the END", "no stop here",
            "spaces   between    words.", "Digit 7 IS LEFT and right.", "x."]
    return out


def main():
    D = import_reference_data()
    sents = sentences()
    vocab = D.build_vocab(sents[:40] + sents[40:43])          # the last edge cases stay out of the vocabulary
    enc = []
    for s in sents + ["completely unknown words only.", "digit 3 is nowhere."]:
        cap = [vocab(t) for t in vocab.tokenize(s)]            # Dataset.__getitem__, data/__init__.py:250-254
        if cap[-1] != vocab(vocab.END):
            cap.append(vocab(vocab.END))
        enc.append({"sentence": s, "tokens": cap, "words": vocab.to_words(cap)})
    rnd = random.Random(5)
    batches = []
    for B in (1, 4, 8, 13):
        picks = [rnd.randrange(len(enc)) for _ in range(B)]
        data = [(torch.full((2,), float(i)), torch.Tensor(enc[i]["tokens"])) for i in picks]
        vids, targets, lengths = D.collate_fn(data)
        batches.append({"picks": picks, "order": [int(v[0]) for v in vids], "targets": targets.tolist(),
                        "lengths": [int(l) for l in lengths]})
    fx = {"source": "txt2vid/data/__init__.py:250-254,260-355 (live reference, nvidia.dali stubbed)",
          "vocab_sentences": sents[:43], "word2idx": vocab.word2idx, "len": len(vocab), "encoded": enc,
          "batches": batches}
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "token_fixtures.json")
    with open(out, "w") as f:
        json.dump(fx, f)
    print("wrote", out, "vocab", len(vocab), "sentences", len(enc))


if __name__ == "__main__":
    main()
