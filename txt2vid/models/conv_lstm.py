"""Drop-in path of the reference's txt2vid/models/conv_lstm.py."""
from txt2vid_b200.blocks import ConvLSTM, ConvLSTMCell  # noqa: F401
