"""
IWE primitives with the signatures of utils/iwe.py.  deblur_events / compute_pol_iwe (:95-153) run as one CUDA kernel
(ef_iwe_image).  The intermediate helpers purge_unfeasible / get_interpolation / interpolate (:4-92) are fused into the loss,
metric and image kernels on the training / evaluation paths; the stand-alone versions below (one small kernel each) return
the reference's intermediate tensors for callers that use the helpers directly.
"""
import torch

from .. import _lib as L
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


def purge_unfeasible(x, res):
    """
    utils/iwe.py:4-17 (one kernel, ef_iwe_purge_unfeasible).
    :param x: [batch_size x N x 2] locations (y, x) of motion-compensated events
    :return x * mask, mask [batch_size x N x 1]: 0 where a location lies outside the image
    """
    x = _f32(x)
    out = torch.empty_like(x)
    mask = torch.empty((x.shape[0], x.shape[1], 1), device=x.device, dtype=torch.float32)
    L.LAUNCHES += 1
    L.check(L.lib().ef_iwe_purge_unfeasible(L.ptr(x), x.shape[0] * x.shape[1], int(res[0]), int(res[1]), L.ptr(out), L.ptr(mask), L.stream()),
            "ef_iwe_purge_unfeasible")
    return out, mask


def get_interpolation(events, flow, tref, res, flow_scaling, round_idx=False):
    """
    utils/iwe.py:20-74 (one kernel, ef_iwe_get_interpolation): warp the events with their per-event flow and split them
    bilinearly (or round them) onto the pixel grid.
    :param events: [batch_size x N x 4] input events (ts, y, x, p)
    :param flow: [batch_size x N x 2] optical flows (y, x)
    :return idx, weights: [batch_size x 4N x 1] (corner order top-left, top-right, bottom-left, bottom-right along N), or
            [batch_size x N x 1] with round_idx.  Forward only: for gradients use loss.flow.EventWarping.
    """
    events, flow = _f32(events), _f32(flow)
    B, N = events.shape[:2]
    M = N if round_idx else 4 * N
    idx = torch.empty((B, M, 1), device=events.device, dtype=torch.float32)
    weights = torch.empty((B, M, 1), device=events.device, dtype=torch.float32)
    p = L.IweInterpParams()
    p.B, p.N, p.H, p.W, p.round_idx = B, N, int(res[0]), int(res[1]), int(bool(round_idx))
    p.tref, p.flow_scaling = float(tref), float(flow_scaling)
    p.events, p.flow, p.idx, p.weights = L.ptr(events), L.ptr(flow), L.ptr(idx), L.ptr(weights)
    L.call("ef_iwe_get_interpolation", p)
    return idx, weights


def interpolate(idx, weights, res, polarity_mask=None):
    """
    utils/iwe.py:77-92 (memset + one kernel, ef_iwe_interpolate): image-like representation of the warped events.
    :param idx: [batch_size x N x 1] warped event locations (flat pixel index; float or long like the reference's callers pass)
    :param weights: [batch_size x N x 1] interpolation weights
    :param polarity_mask: [batch_size x N x 1] or None
    :return [batch_size x 1 x H x W]
    """
    idx, weights = _f32(idx), _f32(weights)
    pm = None if polarity_mask is None else _f32(polarity_mask.expand_as(weights))
    B, M = idx.shape[:2]
    iwe = torch.empty((B, 1, int(res[0]), int(res[1])), device=idx.device, dtype=torch.float32)
    L.LAUNCHES += 1
    L.check(L.lib().ef_iwe_interpolate(L.ptr(idx), L.ptr(weights), L.ptr(pm), B, M, int(res[0]), int(res[1]), L.ptr(iwe), L.stream()),
            "ef_iwe_interpolate")
    return iwe


def _f32(t):
    if not t.is_cuda:
        raise L.EventFlowError("event_flow_b200 has no CPU path: tensors must live on a CUDA device (got %s)" % t.device)
    return t.detach().to(torch.float32).contiguous()
