"""Hot-path part of txt2vid.data: token indexing + prefetcher (the file / LMDB datasets are out of scope)."""
from txt2vid_b200.data import (MovingDigits, MovingDigitsDataset, SyntheticVideoCaptions, Vocab, build_vocab, collate_fn,  # noqa: F401
                               data_prefetcher)
