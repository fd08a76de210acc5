"""
IWE primitives with the signatures of utils/iwe.py.  deblur_events / compute_pol_iwe (:95-153) run as one CUDA kernel
(ef_iwe_image).  The intermediate helpers get_interpolation / interpolate (:20-92) have no stand-alone device version:
inside this package they only exist fused into the loss and image kernels.
"""
import torch

from .. import ops


def compute_pol_iwe(flow, event_list, res, pos_mask, neg_mask, flow_scaling=128, round_idx=True):
    """
    :param flow: [batch_size x 2 x H x W] optical flow map
    :param event_list: [batch_size x N x 4] input events (ts, y, x, p)
    :param pos_mask / neg_mask: [batch_size x N x 1]
    :return iwe: [batch_size x 2 x H x W] per-polarity image of warped events (tref = 1)
    """
    pol_mask = torch.cat([pos_mask, neg_mask], dim=2)
    return ops.iwe_image(event_list, pol_mask, res, flow=flow, tref=1.0, flow_scaling=flow_scaling, round_idx=round_idx)


def deblur_events(flow, event_list, res, flow_scaling=128, round_idx=True, polarity_mask=None):
    """:return iwe: [batch_size x 1 x H x W] image of warped events (utils/iwe.py:95-129)."""
    if polarity_mask is None:
        polarity_mask = torch.ones(event_list.shape[0], event_list.shape[1], 1, device=event_list.device)
    pol_mask = torch.cat([polarity_mask, torch.zeros_like(polarity_mask)], dim=2)
    return ops.iwe_image(event_list, pol_mask, res, flow=flow, tref=1.0, flow_scaling=flow_scaling, round_idx=round_idx)[:, 0:1]


def get_interpolation(*args, **kwargs):
    raise NotImplementedError("get_interpolation is fused into ef_iwe_loss_fwd / ef_iwe_image in event_flow_b200")


def interpolate(*args, **kwargs):
    raise NotImplementedError("interpolate is fused into ef_iwe_loss_fwd / ef_iwe_image in event_flow_b200")


def purge_unfeasible(*args, **kwargs):
    raise NotImplementedError("purge_unfeasible is fused into ef_iwe_loss_fwd / ef_iwe_image in event_flow_b200")
