"""State-copy helpers with the semantics of models/model_util.py:82-102 (model.states returns clones)."""
import copy


def recursive_clone(tensor):
    if hasattr(tensor, "clone"):
        return tensor.clone()
    try:
        return type(tensor)(recursive_clone(t) for t in tensor)
    except TypeError:
        print("{} is not iterable and has no clone() method.".format(tensor))


def copy_states(states):
    if states[0] is None:
        return copy.deepcopy(states)
    return recursive_clone(states)


# ---- helpers of the U-Net family (models/model_util.py:14-80) ------------------------------------------------------------
from math import ceil, floor  # noqa: E402

import torch  # noqa: E402
from torch.nn import ZeroPad2d  # noqa: E402


def skip_concat(x1, x2):
    diffY = x2.size()[2] - x1.size()[2]
    diffX = x2.size()[3] - x1.size()[3]
    if diffX or diffY:
        x1 = ZeroPad2d((diffX // 2, diffX - diffX // 2, diffY // 2, diffY - diffY // 2))(x1)
    return torch.cat([x1, x2], dim=1)


def skip_sum(x1, x2):
    diffY = x2.size()[2] - x1.size()[2]
    diffX = x2.size()[3] - x1.size()[3]
    if diffX or diffY:
        x1 = ZeroPad2d((diffX // 2, diffX - diffX // 2, diffY // 2, diffY - diffY // 2))(x1)
    return x1 + x2


def optimal_crop_size(max_size, max_subsample_factor, safety_margin=0):
    """Smallest size >= max_size divisible by 2^max_subsample_factor (+ margin)."""
    crop_size = int(pow(2, max_subsample_factor) * ceil(max_size / pow(2, max_subsample_factor)))
    crop_size += safety_margin * pow(2, max_subsample_factor)
    return crop_size


class CropParameters:
    """Zero-padding of the input to a size the encoder pyramid divides, and the crop back (models/model_util.py:40-80)."""

    def __init__(self, width, height, num_encoders, safety_margin=0):
        self.height = height
        self.width = width
        self.num_encoders = num_encoders
        self.width_crop_size = optimal_crop_size(self.width, num_encoders, safety_margin)
        self.height_crop_size = optimal_crop_size(self.height, num_encoders, safety_margin)
        self.padding_top = ceil(0.5 * (self.height_crop_size - self.height))
        self.padding_bottom = floor(0.5 * (self.height_crop_size - self.height))
        self.padding_left = ceil(0.5 * (self.width_crop_size - self.width))
        self.padding_right = floor(0.5 * (self.width_crop_size - self.width))
        self.pad = ZeroPad2d((self.padding_left, self.padding_right, self.padding_top, self.padding_bottom))
        self.cx = floor(self.width_crop_size / 2)
        self.cy = floor(self.height_crop_size / 2)
        self.ix0 = self.cx - floor(self.width / 2)
        self.ix1 = self.cx + ceil(self.width / 2)
        self.iy0 = self.cy - floor(self.height / 2)
        self.iy1 = self.cy + ceil(self.height / 2)

    def crop(self, img):
        return img[..., self.iy0:self.iy1, self.ix0:self.ix1]
