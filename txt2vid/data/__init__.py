"""Hot-path part of txt2vid.data: token indexing + prefetcher (the file / LMDB datasets are out of scope)."""
from txt2vid_b200.data import SyntheticVideoCaptions, Vocab, build_vocab, collate_fn, data_prefetcher  # noqa: F401
