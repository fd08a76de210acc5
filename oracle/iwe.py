"""
CPU oracle for hot path B: event warping, image of warped events (IWE) and the contrast-maximisation loss.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, in torch-CPU fp32 with the reference's op order (so autograd reproduces the reference's sub-gradient choices
at ties):
  * purge_unfeasible / get_interpolation / interpolate      utils/iwe.py:4-92
  * deblur_events / compute_pol_iwe                          utils/iwe.py:95-153
  * per-event flow gather                                    loss/flow.py:65-79
  * EventWarping.forward                                     loss/flow.py:176-301
The window is given in "map form": every pass t contributes its flow map and its events; this is the form the CUDA
path consumes (it never materialises the reference's concatenated per-event flow lists, loss/flow.py:81-116).
"""

import torch


def warp_and_split(events, flow, tref, res, flow_scaling, round_idx=False):
    """
    utils/iwe.py:20-74.  events [B,N,4] (ts,y,x,p); flow [B,N,2] (fy,fx).
    Returns flat pixel index [B,4N,1] (or [B,N,1] when rounding) and bilinear weights of the same shape.  Corner
    order TL,TR,BL,BR concatenated along N; out-of-bounds corners get weight 0 and index 0 (iwe.py:13-17,65-72).
    """
    warped = events[:, :, 1:3] + (tref - events[:, :, 0:1]) * flow * flow_scaling
    if round_idx:
        idx = torch.round(warped)  # half-to-even
        weights = torch.ones_like(idx)
    else:
        top_y = torch.floor(warped[:, :, 0:1])
        bot_y = torch.floor(warped[:, :, 0:1] + 1)
        left_x = torch.floor(warped[:, :, 1:2])
        right_x = torch.floor(warped[:, :, 1:2] + 1)
        idx = torch.cat(
            [
                torch.cat([top_y, left_x], dim=2),
                torch.cat([top_y, right_x], dim=2),
                torch.cat([bot_y, left_x], dim=2),
                torch.cat([bot_y, right_x], dim=2),
            ],
            dim=1,
        )
        warped4 = torch.cat([warped] * 4, dim=1)
        weights = torch.max(torch.zeros_like(warped4), 1 - torch.abs(warped4 - idx))
    mask = torch.ones((idx.shape[0], idx.shape[1], 1), dtype=idx.dtype)
    oob = (idx[:, :, 0:1] < 0) + (idx[:, :, 0:1] >= res[0]) + (idx[:, :, 1:2] < 0) + (idx[:, :, 1:2] >= res[1])
    mask[oob] = 0
    idx = idx * mask
    weights = torch.prod(weights, dim=-1, keepdim=True) * mask
    flat = idx[:, :, 0:1] * res[1] + idx[:, :, 1:2]
    return flat, weights


def scatter_image(idx, weights, res, polarity_mask=None):
    """utils/iwe.py:77-92: scatter_add of (masked) weights into a [B,1,H,W] image."""
    if polarity_mask is not None:
        weights = weights * polarity_mask
    iwe = torch.zeros((idx.shape[0], res[0] * res[1], 1), dtype=weights.dtype)
    iwe = iwe.scatter_add(1, idx.long(), weights)
    return iwe.view(idx.shape[0], 1, res[0], res[1])


def gather_event_flow(flow_map, events, res):
    """loss/flow.py:65-79: flow [B,2,H,W] (x,y channels) sampled at each event's pixel -> [B,N,2] ordered (fy,fx)."""
    flat = (events[:, :, 1] * res[1] + events[:, :, 2]).long()
    f = flow_map.reshape(flow_map.shape[0], 2, -1)
    fy = torch.gather(f[:, 1, :], 1, flat)
    fx = torch.gather(f[:, 0, :], 1, flat)
    return torch.stack([fy, fx], dim=2)


def pol_iwe(flow_map, events, res, pos_mask, neg_mask, flow_scaling=128, round_idx=True):
    """utils/iwe.py:95-153 (deblur_events + compute_pol_iwe): per-polarity IWE at tref=1 -> [B,2,H,W]."""
    ev_flow = gather_event_flow(flow_map, events, res)
    idx, w = warp_and_split(events, ev_flow, 1, res, flow_scaling, round_idx=round_idx)
    out = []
    for m in (pos_mask, neg_mask):
        if not round_idx and m is not None:
            m = torch.cat([m] * 4, dim=1)
        out.append(scatter_image(idx, w, res, m))
    return torch.cat(out, dim=1)


def _direction_loss(events, ev_flow, pol4, ts_w, tref, max_ts, res, flow_scaling, loss_scaling):
    """One warping direction of loss/flow.py:196-226 (fw) / :229-259 (bw).  Returns a scalar summed over batch."""
    idx, w = warp_and_split(events, ev_flow, tref, res, flow_scaling)
    iwe_pos = scatter_image(idx, w, res, pol4[:, :, 0:1])
    iwe_neg = scatter_image(idx, w, res, pol4[:, :, 1:2])
    ts_pos = scatter_image(idx, w * ts_w, res, pol4[:, :, 0:1])
    ts_neg = scatter_image(idx, w * ts_w, res, pol4[:, :, 1:2])
    ts_pos = ts_pos / (iwe_pos + 1e-9)
    ts_neg = ts_neg / (iwe_neg + 1e-9)
    ts_pos = ts_pos / max_ts
    ts_neg = ts_neg / max_ts
    ts_pos = ts_pos.view(ts_pos.shape[0], -1)
    ts_neg = ts_neg.view(ts_neg.shape[0], -1)
    loss = torch.sum(ts_pos**2, dim=1) + torch.sum(ts_neg**2, dim=1)
    if loss_scaling:
        nonzero = iwe_pos + iwe_neg
        nonzero[nonzero > 0] = 1  # in place on purpose: pixels that stay 0 keep gradient
        nonzero = nonzero.view(nonzero.shape[0], -1)
        loss = loss / torch.sum(nonzero, dim=1)
    return torch.sum(loss)


def smoothness_loss(fx, fy, event_mask, smoothing_mask, overwrite_intermediate):
    """loss/flow.py:262-294 (+ masks :184-190).  fx, fy, event_mask: [B,T,H,W]."""

    def charb(a, b, sl1, sl2):
        return torch.sqrt(((a[sl1] - a[sl2]) + (b[sl1] - b[sl2])) ** 2 + 1e-6)

    S = slice(None)
    pairs = {
        "dx": ((S, S, S, slice(None, -1)), (S, S, S, slice(1, None))),
        "dy": ((S, S, slice(None, -1), S), (S, S, slice(1, None), S)),
        "dr": ((S, S, slice(None, -1), slice(None, -1)), (S, S, slice(1, None), slice(1, None))),
        "ur": ((S, S, slice(1, None), slice(None, -1)), (S, S, slice(None, -1), slice(1, None))),
        "dt": ((S, slice(None, -1), S, S), (S, slice(1, None), S, S)),
    }
    total = 0
    components = 0
    for name, (a, b) in pairs.items():
        if name == "dt" and overwrite_intermediate:
            continue
        c = charb(fx, fy, a, b)
        if smoothing_mask:
            c = event_mask[a] * event_mask[b] * c
        total = total + c.sum()
        components += 1
    return total / components / fx.shape[1]


def event_warping_loss(
    events,
    pol_mask,
    pass_of_event,
    flow_maps,
    event_mask,
    res,
    *,
    flow_scaling=None,
    weight=0.001,
    loss_scaling=True,
    smoothing_mask=True,
    overwrite_intermediate=False,
    passes=None,
):
    """
    EventWarping.forward (loss/flow.py:176-301) on a window in map form.
    :param events: [B,Ntot,4] (ts,y,x,p), ts already offset by its pass index (loss/flow.py:90)
    :param pol_mask: [B,Ntot,2]
    :param pass_of_event: LongTensor [Ntot], pass index of each event column (ignored when overwrite_intermediate)
    :param flow_maps: list over scales of [B,T,2,H,W] (x,y channels); T=1 when overwrite_intermediate
    :param event_mask: [B,T,H,W]
    """
    if flow_scaling is None:
        flow_scaling = max(res)
    max_ts = passes if passes is not None else int(pass_of_event.max().item()) + 1
    pol4 = torch.cat([pol_mask] * 4, dim=1)
    ts4 = torch.cat([events[:, :, 0:1]] * 4, dim=1)
    B, N = events.shape[:2]
    loss = 0
    for fm in flow_maps:
        T = fm.shape[1]
        # per-event flow, gathered from the map of the event's own pass (loss/flow.py:65-87)
        flat = (events[:, :, 1] * res[1] + events[:, :, 2]).long()
        t_idx = pass_of_event.view(1, N).expand(B, N) if T > 1 else torch.zeros(B, N, dtype=torch.long)
        f = fm.reshape(B, T, 2, -1)
        bi = torch.arange(B).view(B, 1).expand(B, N)
        fy = f[bi, t_idx, 1, flat]
        fx = f[bi, t_idx, 0, flat]
        ev_flow = torch.stack([fy, fx], dim=2)
        fw = _direction_loss(events, ev_flow, pol4, ts4, max_ts, max_ts, res, flow_scaling, loss_scaling)
        bw = _direction_loss(events, ev_flow, pol4, max_ts - ts4, 0, max_ts, res, flow_scaling, loss_scaling)
        sm = smoothness_loss(fm[:, :, 0], fm[:, :, 1], event_mask, smoothing_mask, overwrite_intermediate)
        loss = loss + fw + bw + weight * sm
    return loss / len(flow_maps)


def spatial_variance(x):
    """loss/flow.py:13-23: unbiased variance over the pixels of each [B,C,H,W] channel."""
    return torch.var(x.view(x.shape[0], x.shape[1], 1, -1), dim=3, keepdim=True)


def fwl_rsat(events, pol_mask, ev_flow, passes, res, flow_scaling):
    """
    FWL.forward and RSAT.forward (loss/flow.py:481-500, 514-579) on an accumulated validation window.
    events [B,N,4] (ts offset by pass), ev_flow [B,N,2] per-event flow (fy,fx).  Returns (FWL [B], RSAT [B]).
    """
    max_ts = passes
    ts = events[:, :, 0:1]
    out = []
    for flow in (ev_flow, ev_flow * 0):
        idx, w = warp_and_split(events, flow, max_ts, res, flow_scaling, round_idx=True)
        iwe = scatter_image(idx, w, res)
        pos = scatter_image(idx, w, res, pol_mask[:, :, 0:1])
        neg = scatter_image(idx, w, res, pol_mask[:, :, 1:2])
        pos_ts = scatter_image(idx, w * ts, res, pol_mask[:, :, 0:1]) / (pos + 1e-9) / max_ts
        neg_ts = scatter_image(idx, w * ts, res, pol_mask[:, :, 1:2]) / (neg + 1e-9) / max_ts
        ssum = torch.sum(pos_ts.view(pos.shape[0], -1) ** 2, dim=1) + torch.sum(neg_ts.view(neg.shape[0], -1) ** 2, dim=1)
        nz = pos + neg
        nz[nz > 0] = 1
        out.append((spatial_variance(iwe).view(-1), ssum / torch.sum(nz.view(nz.shape[0], -1), dim=1)))
    return out[0][0] / out[1][0], out[0][1] / out[1][1]


def aee(flow, gtflow, event_mask, dt_gt, dt_input, flow_scaling):
    """AEE.forward (loss/flow.py:597-628).  flow, gtflow [B,2,H,W]; event_mask [B,H,W] (last pass).  Returns (AEE [B], pct [B])."""
    flow = flow * flow_scaling
    flow = flow * (dt_gt / dt_input)
    flow_mag = flow.pow(2).sum(1).sqrt()
    error = (flow - gtflow).pow(2).sum(1).sqrt()
    gt_mask = ~((gtflow[:, 0] == 0.0) * (gtflow[:, 1] == 0.0))
    mask = (event_mask.bool() * gt_mask).view(flow.shape[0], -1)
    error = error.view(flow.shape[0], -1) * mask
    flow_mag = flow_mag.view(flow.shape[0], -1) * mask
    n_valid = torch.sum(mask, dim=1)
    outliers = (error > 3.0) * (error > 0.05 * flow_mag)
    return torch.sum(error, dim=1) / (n_valid + 1e-9), outliers.sum() / (n_valid + 1e-9)
