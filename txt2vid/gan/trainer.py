"""txt2vid.gan.trainer entry points (train/gan.py:15)."""
from txt2vid_b200.trainer import (add_params_to_parser, multiscale_data, save_frames, save_sentences,  # noqa: F401
                                  test, train, train_iteration)
