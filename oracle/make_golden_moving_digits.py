"""SURVEY 8(f1) fixture: the LIVE reference's synthetic moving-MNIST generator
(txt2vid/data/synthetic/generate.py:18-47 generate_frames, :59-182 generate_examples) run on a small synthetic digit
bank, recorded so that the product's on-device generator (txt2vid_b200/data.MovingDigits: host draws in the
reference's RNG order, t2v_moving_digits / t2v_grammar_tokens on the device) can be checked bit-exactly, here and on
the GPU box.

    python oracle/make_golden_moving_digits.py     # build container only; writes tests/golden/moving_digits.json

Test infrastructure only.  nvidia.dali (imported by txt2vid/data/__init__.py:16-18) is stubbed; save_video (cv2 XVID
writer, generate.py:50-57) is replaced by a function that keeps the frames the generator yields -- the fixture pins the
frames BEFORE the lossy codec.  MNIST itself cannot be downloaded here: the bank is 10 classes x 3 random 28x28 'L'
images from numpy seed 7 (the test rebuilds it from the same seed)."""
import hashlib
import json
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
W = H = 64
T = 16
N_EXAMPLES = 12
SEED = 300                     # generate.py:214-215


def digit_bank(seed=7, per_class=3):
    rng = np.random.RandomState(seed)
    return {c: [rng.randint(0, 256, (28, 28)).astype(np.uint8) for _ in range(per_class)] for c in range(10)}


def main():
    for name in ("nvidia", "nvidia.dali", "nvidia.dali.pipeline", "nvidia.dali.ops", "nvidia.dali.types"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["nvidia.dali.pipeline"].Pipeline = object
    sys.path.insert(0, REF)
    import txt2vid.data.synthetic.generate as gen
    from PIL import Image
    assert gen.__file__.startswith(REF)
    bank = digit_bank()
    objects = {c: [Image.fromarray(a, mode="L") for a in imgs] for c, imgs in bank.items()}
    object_classes = set(bank.keys())
    object_names = {c: "digit {}".format(c) for c in object_classes}
    videos = {}

    def keep(frames, path, fps, frame_size):
        videos[int(os.path.basename(path).split(".")[0])] = np.stack([np.array(f) for f in frames])   # (T, H, W, 3)

    gen.save_video = keep
    gen.FRAME_SIZE = np.array([W, H])          # module global only set under __main__ (generate.py:224)
    random.seed(SEED)
    np.random.seed(SEED)
    sent_out = "/tmp/_moving_digits_sent.pickle"
    gen.generate_examples("/tmp/_moving_digits_videos", sent_out, num_examples=N_EXAMPLES, fps=30,
                          frame_size=gen.FRAME_SIZE, num_frames=T, object_classes=object_classes, objects=objects,
                          object_names=object_names)
    import pickle
    with open(sent_out, "rb") as f:
        sent_map = pickle.load(f)
    ex = []
    for i in range(N_EXAMPLES):
        v = videos[i]
        assert v.shape == (T, H, W, 3) and v.dtype == np.uint8
        pos = []
        for t in range(T):                      # top-left corner of the pasted patch (non-black bounding box is not
            ys, xs = np.nonzero(v[t, :, :, 0])  # reliable for random patches with zero rows: recorded for information)
            pos.append([int(xs.min()) if len(xs) else -1, int(ys.min()) if len(ys) else -1])
        ex.append({"sentence": sent_map[i][0], "sha256": hashlib.sha256(v.tobytes()).hexdigest(), "bbox_min": pos,
                   "frame0_row_sums": v[0, :, :, 0].sum(axis=1).astype(int).tolist()})
    fx = {"source": "txt2vid/data/synthetic/generate.py:18-47,59-182 (live reference; save_video replaced, DALI stubbed)",
          "seed": SEED, "bank_seed": 7, "per_class": 3, "W": W, "H": H, "T": T, "examples": ex}
    out = os.path.join(HERE, "..", "tests", "golden", "moving_digits.json")
    with open(out, "w") as f:
        json.dump(fx, f)
    print("wrote", os.path.normpath(out), [e["sentence"] for e in ex[:4]])


if __name__ == "__main__":
    main()
