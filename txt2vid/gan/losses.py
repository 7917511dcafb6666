"""--D_loss / --G_loss txt2vid.gan.losses.* (train/gan.py:156)."""
from txt2vid_b200.gan import (HingeGanLoss, LabelledGanLoss, MixedGanLoss, RaLSGANLoss, RaSGANLoss,  # noqa: F401
                              RSGANLoss, VanillaGanLoss, WassersteinGanLoss, _gradient_penalty, get_labels_for,
                              gradient_penalty)
