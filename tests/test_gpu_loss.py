"""
GPU parity of hot path B (T5, T6): event-warping loss value and analytic gradient against the reference's golden numbers
(random flows, zero flow = all ties, half-integer flows + out-of-bounds), IWE images, encodings, and a known-answer test.
Tolerances: loss rel 1e-5, d loss / d flow rel 1e-3 (north-star), counts / rounded IWEs bit-exact.
"""
import glob
import os

import pytest
import torch

from oracle import encodings as oenc
from oracle import iwe as oiwe
from tests.conftest import GOLDEN, load_golden
from tests.util import assert_rel

pytestmark = pytest.mark.gpu
LOSSES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "loss_*.npz")))
DEV = "cuda"


def cuda_loss_from_golden(g):
    from event_flow_b200.loss.flow import EventWarping

    scaling, smask, overwrite, weight, T, N = g["cfg"].tolist()
    T, N = int(T), int(N)
    flow = g["flow"].to(DEV)
    H, W = flow.shape[-2:]
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": weight, "overwrite_intermediate": bool(overwrite)},
           "model": {"mask_output": bool(smask)}}
    L = EventWarping(cfg, DEV, loss_scaling=bool(scaling))
    flows = [flow[:, t].clone().requires_grad_(True) for t in range(T)]
    ev = g["events"].clone()
    for t in range(T):
        e = ev[:, t * N:(t + 1) * N].clone()
        e[:, :, 0] -= t  # the fixture stores offset timestamps; the API offsets them itself (loss/flow.py:90)
        L.event_flow_association([flows[t]], e.to(DEV), g["pol_mask"][:, t * N:(t + 1) * N].to(DEV), g["event_mask"][:, t:t + 1].to(DEV))
    assert L.num_events == T * N
    if overwrite:
        L.overwrite_intermediate_flow([flows[-1]])
    loss = L()
    loss.backward()
    return loss, flows


@pytest.mark.parametrize("name", LOSSES)
def test_event_warping_loss_and_gradient_match_reference(name):
    g = load_golden(name)
    loss, flows = cuda_loss_from_golden(g)
    assert_rel(loss, g["loss"], 1e-5, "loss")
    for t, f in enumerate(flows):
        if f"grad_{t}" in g:
            assert_rel(f.grad, g[f"grad_{t}"], 1e-3, f"grad[{t}]")
        else:
            assert f.grad is None or f.grad.abs().max() == 0


def test_iwe_image_matches_reference():
    from event_flow_b200.utils.iwe import compute_pol_iwe

    g = load_golden("iwe_image")
    H, W = g["flow"].shape[-2:]
    pm = g["pol_mask"].to(DEV)
    for rnd, key in ((True, "iwe_round"), (False, "iwe_bilinear")):
        out = compute_pol_iwe(g["flow"].to(DEV), g["events"].to(DEV), (H, W), pm[:, :, 0:1], pm[:, :, 1:2], flow_scaling=max(H, W), round_idx=rnd)
        if rnd:
            assert torch.equal(out.cpu(), g[key])
        else:
            torch.testing.assert_close(out.cpu(), g[key], rtol=1e-5, atol=1e-6)


def test_encodings_match_reference_counts_bit_exact():
    from event_flow_b200.dataloader import encodings as E

    g = load_golden("encodings")
    B = g["ts"].shape[0]
    H, W = g["cnt_0"].shape[-2:]
    ev = torch.stack([g["ts"], g["ys"], g["xs"], g["ps"]], dim=2).to(DEV)
    for bins in (2, 5):
        d = E.encode_batch(ev, (H, W), bins)
        for b in range(B):
            assert torch.equal(d["event_cnt"][b].cpu(), g[f"cnt_{b}"])
            assert torch.equal(d["event_mask"][b, 0].cpu(), g[f"mask_{b}"])
            torch.testing.assert_close(d["event_voxel"][b].cpu(), g[f"voxel{bins}_{b}"], rtol=1e-5, atol=1e-6)
            assert torch.equal(d["event_list_pol_mask"][b].cpu(), oenc.polarity_mask(g["ps"][b]))
    one = E.events_to_channels(g["xs"][0].to(DEV), g["ys"][0].to(DEV), g["ps"][0].to(DEV), (H, W))
    assert torch.equal(one.cpu(), g["cnt_0"])


def test_full_size_window_against_oracle_and_properties():
    """cfg-2 size (B=8, 128x128, T=10, N=1000): loss vs oracle; invariances the domain offers (size-independent checks)."""
    from event_flow_b200 import ops

    B, H, W, T, N = 8, 128, 128, 10, 1000
    g = torch.Generator().manual_seed(0)
    evs, pms, masks = [], [], []
    for t in range(T):
        d = oenc.encode_window(*oenc.synthetic_events(B, N, H, W, 50 + t), H, W, 2)
        e = d["event_list"].clone()
        e[:, :, 0] += t
        evs.append(e), pms.append(d["event_list_pol_mask"]), masks.append(d["event_mask"])
    events, pol, mask = torch.cat(evs, 1), torch.cat(pms, 1), torch.cat(masks, 1)
    flow = (torch.rand((1, B, T, 2, H, W), generator=g) - 0.5) * 0.1
    kw = dict(passes=T, n_per_pass=N, flow_scaling=128, weight=0.001)
    fd = flow.to(DEV).requires_grad_(True)
    loss = ops.event_warping_loss(fd, events.to(DEV), pol.to(DEV), mask.to(DEV), **kw)
    loss.backward()
    fo = flow[0].clone().requires_grad_(True)
    loss_o = oiwe.event_warping_loss(events, pol, torch.arange(T).repeat_interleave(N), [fo], mask, (H, W), weight=0.001, passes=T)
    loss_o.backward()
    assert_rel(loss, loss_o, 1e-5, "loss")
    assert_rel(fd.grad[0], fo.grad, 1e-3, "grad")
    # property: the loss is invariant to the order of events inside a pass (scatter-add commutes)
    perm = torch.cat([t * N + torch.randperm(N, generator=g) for t in range(T)])
    loss_p = ops.event_warping_loss(flow.to(DEV), events[:, perm].to(DEV), pol[:, perm].to(DEV), mask.to(DEV), **kw)
    assert_rel(loss_p, loss, 1e-5, "permutation invariance")
    # property: the loss is a sum over samples -> batch halves add up (smoothness weight 0 isolates the event terms)
    kw0 = dict(kw, weight=0.0)
    full = ops.event_warping_loss(flow.to(DEV), events.to(DEV), pol.to(DEV), mask.to(DEV), **kw0)
    h1 = ops.event_warping_loss(flow[:, :4].contiguous().to(DEV), events[:4].to(DEV), pol[:4].to(DEV), mask[:4].to(DEV), **kw0)
    h2 = ops.event_warping_loss(flow[:, 4:].contiguous().to(DEV), events[4:].to(DEV), pol[4:].to(DEV), mask[4:].to(DEV), **kw0)
    assert_rel(h1 + h2, full, 1e-5, "additivity over the batch")


def test_known_answer_constant_flow_minimises_loss():
    """tools/demo_iwe.py:69-91 idea: events generated by a known constant flow -> the loss over a (u,v) grid is minimal there."""
    from event_flow_b200 import ops

    H = W = 64
    N, T = 4000, 1
    g = torch.Generator().manual_seed(1)
    u_true, v_true = 6.0, -4.0  # px per window
    x0 = torch.rand(N, generator=g) * (W - 20) + 10
    y0 = torch.rand(N, generator=g) * (H - 20) + 10
    pts = torch.randint(0, 40, (N,), generator=g)  # 40 point sources -> sharp IWE when compensated
    x0, y0 = x0[pts], y0[pts]
    ts = torch.sort(torch.rand(N, generator=g))[0]
    xs, ys = torch.round(x0 + u_true * ts), torch.round(y0 + v_true * ts)
    ps = torch.ones(N)
    events = torch.stack([ts, ys, xs, ps], 1).unsqueeze(0).to(DEV)
    pol = oenc.polarity_mask(ps).unsqueeze(0).to(DEV)
    mask = torch.ones(1, 1, H, W, device=DEV)
    best, best_uv = None, None
    for u in range(-8, 9, 2):
        for v in range(-8, 9, 2):
            flow = torch.zeros(1, 1, 1, 2, H, W, device=DEV)
            flow[..., 0, :, :], flow[..., 1, :, :] = u, v
            l = ops.event_warping_loss(flow, events, pol, mask, passes=T, n_per_pass=N, flow_scaling=1.0, weight=0.0).item()
            if best is None or l < best:
                best, best_uv = l, (u, v)
    assert best_uv == (6, -4), best_uv


def test_empty_and_ragged_windows():
    from event_flow_b200 import ops
    from event_flow_b200.loss.flow import EventWarping

    H, W = 16, 20
    # ragged passes (different event counts per pass) against the oracle
    Ns = [40, 75, 10]
    g = torch.Generator().manual_seed(4)
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.01, "overwrite_intermediate": False}, "model": {"mask_output": True}}
    L = EventWarping(cfg, DEV)
    evs, pms, masks, flows = [], [], [], []
    for t, n in enumerate(Ns):
        d = oenc.encode_window(*oenc.synthetic_events(2, n, H, W, 70 + t), H, W, 2)
        f = (torch.rand((2, 2, H, W), generator=g) - 0.5) * 0.2
        L.event_flow_association([f.to(DEV)], d["event_list"].clone().to(DEV), d["event_list_pol_mask"].to(DEV), d["event_mask"].to(DEV))
        e = d["event_list"].clone()
        e[:, :, 0] += t
        evs.append(e), pms.append(d["event_list_pol_mask"]), masks.append(d["event_mask"]), flows.append(f)
    pass_of = torch.cat([torch.full((n,), t) for t, n in enumerate(Ns)])
    loss_o = oiwe.event_warping_loss(torch.cat(evs, 1), torch.cat(pms, 1), pass_of, [torch.stack(flows, 1)], torch.cat(masks, 1), (H, W),
                                     weight=0.01, passes=3)
    assert_rel(L(), loss_o, 1e-5, "ragged passes")
    # empty event list: IWE image is all zeros, no launch failure
    out = ops.iwe_image(torch.zeros(1, 0, 4, device=DEV), torch.zeros(1, 0, 2, device=DEV), (H, W), flow=torch.zeros(1, 2, H, W, device=DEV))
    assert out.abs().sum().item() == 0


@pytest.mark.parametrize("overwrite", [False, True])
def test_validation_metrics_match_reference(overwrite):
    """FWL / RSAT / AEE classes (loss/flow.py:468-628) through the drop-in API against the reference's golden numbers."""
    from event_flow_b200.loss.flow import AEE, FWL, RSAT

    g = load_golden("metrics")
    B, T = g["flows"].shape[:2]
    H, W = g["flows"].shape[-2:]
    N = g["events"].shape[1] // T
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"overwrite_intermediate": overwrite}}
    metrics = [cls(cfg, DEV, flow_scaling=max(H, W)) for cls in (FWL, RSAT, AEE)]
    for t in range(T):
        e = g["events"][:, t * N:(t + 1) * N].clone()
        e[:, :, 0] -= t
        inputs = {"event_list": e, "event_list_pol_mask": g["pol_mask"][:, t * N:(t + 1) * N], "event_mask": g["event_masks"][:, t:t + 1],
                  "gtflow": g["gtflow"], "dt_input": g["dt_input"], "dt_gt": g["dt_gt"]}
        for m in metrics:
            m.event_flow_association([g["flows"][:, t].to(DEV)], inputs)
    assert metrics[0].num_events == T * N and metrics[2].num_events == float("inf")
    if overwrite:
        for m in metrics:
            m.overwrite_intermediate_flow([g["flows"][:, -1].to(DEV)])
    key = "ow" if overwrite else "seq"
    assert_rel(metrics[0](), g[f"{key}_fwl"], 1e-5, "FWL")
    assert_rel(metrics[1](), g[f"{key}_rsat"], 1e-5, "RSAT")
    aee, pct = metrics[2]()
    assert_rel(aee, g[f"{key}_aee"], 1e-5, "AEE")
    assert_rel(pct, g[f"{key}_pct"], 1e-5, "percent_AEE")
    # window images used by eval_flow.py for visualisation
    iwe = metrics[0].compute_window_iwe()
    ev_img = metrics[0].compute_window_events()
    assert iwe.shape == (B, 2, H, W) and ev_img.shape == (B, 2, H, W)
    assert ev_img.sum().item() == B * T * N  # every event lands in bounds when it is not warped
    assert metrics[0].compute_masked_window_flow().shape == (B, 2, H, W)


# ---------------------------------------------------------------------------------------------------------------------
# stand-alone primitives of utils/iwe.py (idx + weights exact) and models/spiking_util.py
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("round_idx", [False, True])
def test_get_interpolation_interpolate_purge_exact(round_idx):
    """utils.iwe.get_interpolation / interpolate / purge_unfeasible against the oracle restatement: bit-exact idx and weights."""
    from event_flow_b200.utils import iwe as U

    B, N, H, W = 3, 700, 24, 40
    g = torch.Generator().manual_seed(11)
    ts = torch.rand((B, N, 1), generator=g) * 3
    ys = torch.randint(0, H, (B, N, 1), generator=g).float()
    xs = torch.randint(0, W, (B, N, 1), generator=g).float()
    ps = (torch.rand((B, N, 1), generator=g) < 0.5).float() * 2 - 1
    events = torch.cat([ts, ys, xs, ps], 2)
    flow = (torch.rand((B, N, 2), generator=g) - 0.5) * 0.2  # +-10 px over the window at scaling 40: plenty of out-of-bounds corners
    flow[:, :50] = 0.0                                       # exactly-integer warps (ties)
    flow[:, 50:100] = 0.5 / (3.0 * 40)                       # half-pixel cases for torch.round (half to even)
    idx_o, w_o = oiwe.warp_and_split(events, flow, 3.0, (H, W), 40.0, round_idx=round_idx)
    idx, w = U.get_interpolation(events.to(DEV), flow.to(DEV), 3.0, (H, W), 40.0, round_idx=round_idx)
    assert idx.shape == idx_o.shape and w.shape == w_o.shape
    assert torch.equal(idx.cpu(), idx_o) and torch.equal(w.cpu(), w_o)
    assert (w_o == 0).any() and (w_o > 0).any()
    pm = (ps > 0).float() if round_idx else torch.cat([(ps > 0).float()] * 4, 1)
    for mask in (None, pm):
        img = U.interpolate(idx.long(), w, (H, W), polarity_mask=None if mask is None else mask.to(DEV))
        img_o = oiwe.scatter_image(idx_o, w_o, (H, W), mask)
        assert img.shape == (B, 1, H, W)
        torch.testing.assert_close(img.cpu(), img_o, rtol=1e-5, atol=1e-6)
    x = torch.stack([ys[..., 0] + (torch.rand((B, N), generator=g) - 0.5) * 60, xs[..., 0] + (torch.rand((B, N), generator=g) - 0.5) * 90], 2)
    out, m = U.purge_unfeasible(x.to(DEV), (H, W))
    m_o = torch.ones(B, N, 1)
    m_o[((x[:, :, 0:1] < 0) + (x[:, :, 0:1] >= H) + (x[:, :, 1:2] < 0) + (x[:, :, 1:2] >= W))] = 0
    assert torch.equal(m.cpu(), m_o) and torch.equal(out.cpu(), x * m_o)


@pytest.mark.parametrize("name,width", [("arctanspike", 10.0), ("superspike", 10.0), ("trianglespike", 1.0), ("mgspike", 0.5)])
def test_standalone_spike_functions(name, width):
    """models.spiking_util.<fn>(x, thresh, width) called like the reference's cells do: Heaviside forward, surrogate backward."""
    from event_flow_b200.models import spiking_util as S
    from oracle import spiking as osp

    g = torch.Generator().manual_seed(2)
    x = (torch.randn((2, 6, 9, 11), generator=g)).requires_grad_(True)
    th = (torch.rand((6, 1, 1), generator=g) + 0.3).requires_grad_(True)
    xd, thd = x.detach().to(DEV).requires_grad_(True), th.detach().to(DEV).requires_grad_(True)
    z = getattr(S, name)(xd, thd, torch.tensor(width))
    up = torch.randn((2, 6, 9, 11), generator=g)
    z.backward(up.to(DEV))
    u = x.detach() - th.detach()
    assert torch.equal(z.cpu(), (u > 0).float())
    sg = osp.surrogate_grad(u, width, name)
    torch.testing.assert_close(xd.grad.cpu(), up * sg, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(thd.grad.cpu(), -(up * sg).sum(dim=(0, 2, 3)).view(6, 1, 1), rtol=1e-4, atol=1e-6)
    # scalar threshold, default width
    z2 = getattr(S, name)(xd.detach())
    assert torch.equal(z2.cpu(), (x.detach() - 1.0 > 0).float())


def test_pass_form_equals_map_form_and_workspace_is_reusable():
    """
    The two public window forms (concatenated map form / per-pass pointer tables) run the same kernels: equal results.  The
    workspace's barrier counters are restored by every call, so repeated calls on a recycled workspace agree to the bit-level
    noise of the atomics (rel 1e-6); ragged passes and overwrite_intermediate go through both forms.
    """
    from event_flow_b200 import ops

    B, H, W, T = 3, 32, 48, 4
    Ns = [300, 120, 0, 511]
    g = torch.Generator().manual_seed(3)
    evs, pms, masks, flows = [], [], [], []
    for t, n in enumerate(Ns):
        d = oenc.encode_window(*oenc.synthetic_events(B, max(n, 1), H, W, 90 + t), H, W, 2)
        e = d["event_list"][:, :n].clone()
        e[:, :, 0] += t
        evs.append(e.to(DEV)), pms.append(d["event_list_pol_mask"][:, :n].contiguous().to(DEV)), masks.append(d["event_mask"].to(DEV))
        flows.append(((torch.rand((B, 2, H, W), generator=g) - 0.5) * 0.2).to(DEV))
    offs = [0]
    for n in Ns:
        offs.append(offs[-1] + n)
    for overwrite in (False, True):
        fl = [flows[-1]] if overwrite else flows
        mk = [torch.cat(masks, 1).sum(1, keepdim=True).clamp(max=1)] if overwrite else masks
        leafs = [f.clone().requires_grad_(True) for f in fl]
        kw = dict(flow_scaling=48.0, weight=0.01, overwrite_intermediate=overwrite)
        lp = ops.event_warping_loss_passes([leafs], evs, pms, mk, **kw)
        lp.backward()
        maps = torch.stack(fl, 1).unsqueeze(0).clone().requires_grad_(True)
        lm = ops.event_warping_loss(maps, torch.cat(evs, 1), torch.cat(pms, 1), torch.cat(mk, 1), passes=T, n_per_pass=Ns[0],
                                    pass_offsets=offs, **kw)
        lm.backward()
        assert_rel(lp, lm, 1e-6, "pass form vs map form")
        for t, f in enumerate(leafs):
            assert_rel(f.grad, maps.grad[0, :, t], 1e-5, f"grad of map {t}")
        # the same oracle number
        pass_of = torch.cat([torch.full((n,), t) for t, n in enumerate(Ns)])
        lo = oiwe.event_warping_loss(torch.cat(evs, 1).cpu(), torch.cat(pms, 1).cpu(), pass_of, [torch.stack(fl, 1).cpu()], torch.cat(mk, 1).cpu(),
                                     (H, W), flow_scaling=48.0, weight=0.01, passes=T, overwrite_intermediate=overwrite)
        assert_rel(lp, lo, 1e-5, "vs oracle")
        for _ in range(3):  # recycled workspace (ops._WS_POOL): counters must have been restored
            again = ops.event_warping_loss_passes([[f.detach() for f in fl]], evs, pms, mk, **kw)
            assert_rel(again, lp, 1e-6, "repeated call")
    torch.cuda.synchronize()


def test_fwl_counts_events_without_polarity():
    """FWL scatters weight 1 per event regardless of the polarity mask (loss/flow.py:488-494); RSAT uses the mask.  Padded events
    (p = 0 -> mask (0,0)) must count for FWL only; the encoder must not let them clear the event mask."""
    from event_flow_b200 import ops
    from event_flow_b200.dataloader import encodings as E

    B, H, W, N = 2, 20, 28, 400
    ts, ys, xs, ps = oenc.synthetic_events(B, N, H, W, 5)
    ps[:, ::5] = 0.0  # every fifth event carries no polarity
    events = torch.stack([ts, ys, xs, ps], 2)
    pol = torch.stack([(ps > 0).float(), (ps < 0).float()], 2)
    g = torch.Generator().manual_seed(6)
    flow = (torch.rand((B, 1, 2, H, W), generator=g) - 0.5) * 0.3
    ev_flow = oiwe.gather_event_flow(flow[:, 0], events, (H, W))
    fwl_o, rsat_o = oiwe.fwl_rsat(events, pol, ev_flow, 1, (H, W), float(max(H, W)))
    fwl, rsat = ops.iwe_metrics(flow.to(DEV), events.to(DEV), pol.to(DEV), passes=1, n_per_pass=N, flow_scaling=float(max(H, W)))
    assert_rel(fwl, fwl_o, 1e-5, "FWL with unpolarised events")
    assert_rel(rsat, rsat_o, 1e-5, "RSAT with unpolarised events")
    # encoder: a p = 0 event at the pixel of a real event leaves the mask set
    ev = events.clone()
    ev[:, 1, 1:3] = ev[:, 0, 1:3]
    ev[:, 0, 3], ev[:, 1, 3] = 1.0, 0.0
    d = E.encode_batch(ev.to(DEV), (H, W), 2)
    for b in range(B):
        assert d["event_mask"][b, 0, int(ev[b, 0, 1]), int(ev[b, 0, 2])].item() == 1.0
