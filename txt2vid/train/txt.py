"""txt2vid/train/txt.py (caption auto-encoder pre-training) on the B200-native stack: same names."""
from txt2vid_b200.train_txt import SentenceDataset, collate_fn, evaluate, pretrain_step, train  # noqa: F401
from txt2vid_b200.text import Seq2Seq  # noqa: F401
