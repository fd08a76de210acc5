"""
CPU oracle for the event encodings feeding the hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates dataloader/encodings.py:30-85 (events_to_image / events_to_voxel / events_to_channels) and
dataloader/base.py:148-222 (cnt, mask, voxel, list and polarity-mask encodings) for ONE sample; plus the synthetic
event generator every test and bench.py share.
"""

import torch


def events_to_image(xs, ys, ps, sensor_size, accumulate=True):
    """encodings.py:30-45: img[y,x] (+)= p."""
    img = torch.zeros(list(sensor_size), dtype=ps.dtype)
    img.index_put_((ys.long(), xs.long()), ps, accumulate=accumulate)
    return img


def events_to_channels(xs, ys, ps, sensor_size):
    """encodings.py:70-85: per-polarity event counts [2,H,W] (integer-valued, order independent)."""
    pos = torch.where(ps < 0, torch.zeros_like(ps), ps)
    neg = torch.where(ps > 0, torch.zeros_like(ps), ps)
    return torch.stack([events_to_image(xs, ys, ps * pos, sensor_size), events_to_image(xs, ys, ps * neg, sensor_size)])


def events_to_voxel(xs, ys, ts, ps, num_bins, sensor_size, round_ts=False):
    """encodings.py:48-67: temporal bilinear voxel grid [bins,H,W]."""
    ts = ts * (num_bins - 1)
    if round_ts:
        ts = torch.round(ts)
    out = []
    for b in range(num_bins):
        w = torch.max(torch.zeros_like(ts), 1.0 - torch.abs(ts - b))
        out.append(events_to_image(xs, ys, ps * w, sensor_size))
    return torch.stack(out)


def event_mask(xs, ys, ps, sensor_size):
    """base.py:159-172: binary mask [1,H,W] (index_put_ without accumulate)."""
    return events_to_image(xs, ys, ps.abs(), sensor_size, accumulate=False).unsqueeze(0)


def polarity_mask(ps):
    """base.py:207-222 (after custom_collate's transpose): [N,2] = (p>0, p<0)."""
    return torch.stack([(ps > 0).to(ps.dtype), (ps < 0).to(ps.dtype)], dim=1)


def synthetic_events(B, N, H, W, seed):
    """
    Deterministic synthetic window: x,y uniform integer pixels (as fp32), p=+-1 Bernoulli(.5),
    ts sorted uniform normalised to [0,1] (base.py:84-85).  Returns ts, ys, xs, ps each [B,N].
    """
    g = torch.Generator().manual_seed(seed)
    ts = torch.sort(torch.rand(B, N, generator=g))[0]
    ts = (ts - ts[:, :1]) / (ts[:, -1:] - ts[:, :1])
    ys = torch.randint(0, H, (B, N), generator=g).float()
    xs = torch.randint(0, W, (B, N), generator=g).float()
    ps = (torch.randint(0, 2, (B, N), generator=g) * 2 - 1).float()
    return ts, ys, xs, ps


def encode_window(ts, ys, xs, ps, H, W, num_bins, round_ts=False):
    """
    Batch dict the loader would hand to the model / loss (h5.py:330-341 after custom_collate):
    event_voxel [B,bins,H,W], event_cnt [B,2,H,W], event_mask [B,1,H,W], event_list [B,N,4], event_list_pol_mask [B,N,2].
    """
    B = ts.shape[0]
    res = (H, W)
    return {
        "event_voxel": torch.stack([events_to_voxel(xs[b], ys[b], ts[b], ps[b], num_bins, res, round_ts) for b in range(B)]),
        "event_cnt": torch.stack([events_to_channels(xs[b], ys[b], ps[b], res) for b in range(B)]),
        "event_mask": torch.stack([event_mask(xs[b], ys[b], ps[b], res) for b in range(B)]),
        "event_list": torch.stack([ts, ys, xs, ps], dim=2),
        "event_list_pol_mask": torch.stack([polarity_mask(ps[b]) for b in range(B)]),
    }
