"""
Event encodings with the signatures of dataloader/encodings.py (events_to_image :30, events_to_voxel :48,
events_to_channels :70), computed on the GPU by ef_encode_events; plus `encode_batch`, the batched one-launch form.
"""
import torch

from .. import ops


def _as_batch(xs, ys, ts, ps):
    return torch.stack([ts, ys, xs, ps], dim=-1).unsqueeze(0).float()


def events_to_channels(xs, ys, ps, sensor_size=(180, 240)):
    ev = _as_batch(xs, ys, torch.zeros_like(ps), ps)
    return ops.encode_events(ev, sensor_size, 1, want=("cnt",))["event_cnt"][0]


def events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size=(180, 240), round_ts=False):
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    ev = _as_batch(xs, ys, ts, ps)
    return ops.encode_events(ev, sensor_size, num_bins, round_ts=round_ts, want=("voxel",))["event_voxel"][0]


def events_to_mask(xs, ys, ps, sensor_size=(180, 240)):
    """create_mask_encoding, dataloader/base.py:159-172."""
    ev = _as_batch(xs, ys, torch.zeros_like(ps), ps)
    return ops.encode_events(ev, sensor_size, 1, want=("mask",))["event_mask"][0]


def encode_batch(event_list, sensor_size, num_bins, round_ts=False):
    """
    One launch for a whole batch.  event_list [B,N,4] (ts,y,x,p) -> dict with event_cnt [B,2,H,W], event_voxel
    [B,bins,H,W], event_mask [B,1,H,W], event_list_pol_mask [B,N,2] (the loader's batch dict, h5.py:330-341).
    """
    return ops.encode_events(event_list, sensor_size, num_bins, round_ts=round_ts)
