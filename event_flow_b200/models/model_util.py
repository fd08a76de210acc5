"""
Helpers shared by the model classes (role of models/model_util.py): state copies for `model.states` (:82-102), the skip
connections of the U-Nets (:14-27) and the pad / crop bookkeeping for inputs whose size the encoder pyramid does not divide
(:30-80).
"""
import copy
import math

import torch
import torch.nn.functional as F


def recursive_clone(tensor):
    """clone() of a tensor, or of every tensor inside nested tuples / lists (LSTM states are (hidden, cell) tuples)."""
    if hasattr(tensor, "clone"):
        return tensor.clone()
    try:
        return type(tensor)(recursive_clone(t) for t in tensor)
    except TypeError:
        print("{} is not iterable and has no clone() method.".format(tensor))


def copy_states(states):
    """`model.states` hands out copies: a deepcopy of the all-None list, clones otherwise."""
    return copy.deepcopy(states) if states[0] is None else recursive_clone(states)


def _centre_pad_to(x, like):
    """Zero-pad x symmetrically (extra pixel on the bottom / right) to the spatial size of `like`."""
    dy, dx = like.shape[2] - x.shape[2], like.shape[3] - x.shape[3]
    if dy == 0 and dx == 0:
        return x
    return F.pad(x, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))


def skip_concat(x1, x2):
    return torch.cat([_centre_pad_to(x1, x2), x2], dim=1)


def skip_sum(x1, x2):
    return _centre_pad_to(x1, x2) + x2


def optimal_crop_size(max_size, max_subsample_factor, safety_margin=0):
    """Smallest size >= max_size that 2^max_subsample_factor divides, plus `safety_margin` such blocks."""
    block = 2 ** max_subsample_factor
    return int(block * math.ceil(max_size / block)) + safety_margin * block


class CropParameters:
    """
    Input padding to a size the encoder pyramid divides and the window that crops the output back.  Attribute names follow the
    reference (the models read ix0, ix1, iy0, iy1 and call .pad / .crop).
    """

    def __init__(self, width, height, num_encoders, safety_margin=0):
        self.width, self.height, self.num_encoders = width, height, num_encoders
        self.width_crop_size = optimal_crop_size(width, num_encoders, safety_margin)
        self.height_crop_size = optimal_crop_size(height, num_encoders, safety_margin)
        extra_h, extra_w = self.height_crop_size - height, self.width_crop_size - width
        # the larger half of an odd surplus goes to the top / left
        self.padding_top, self.padding_bottom = math.ceil(0.5 * extra_h), math.floor(0.5 * extra_h)
        self.padding_left, self.padding_right = math.ceil(0.5 * extra_w), math.floor(0.5 * extra_w)
        self.cx, self.cy = self.width_crop_size // 2, self.height_crop_size // 2
        self.ix0, self.ix1 = self.cx - width // 2, self.cx + math.ceil(width / 2)
        self.iy0, self.iy1 = self.cy - height // 2, self.cy + math.ceil(height / 2)

    def pad(self, x):
        return F.pad(x, (self.padding_left, self.padding_right, self.padding_top, self.padding_bottom))

    def crop(self, img):
        return img[..., self.iy0:self.iy1, self.ix0:self.ix1]
