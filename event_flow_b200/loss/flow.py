"""
Event-warping (contrast maximisation) loss with the class / method contract of loss/flow.py (EventWarping :26-301).
The window is kept in "map form" -- per pass: the flow map(s), the event list, the polarity mask and the event mask --
and the whole forward (flow gather, forward+backward warping, bilinear scatter to 8 IWE images, contrast reduction,
Charbonnier smoothness) runs in libeventflow.so (ef_iwe_loss_fwd), its analytic gradient in ef_iwe_loss_bwd.
"""
import torch

from .. import ops


class EventWarping(torch.nn.Module):
    def __init__(self, config, device, flow_scaling=None, loss_scaling=True):
        super().__init__()
        self.loss_scaling = loss_scaling
        self.res = config["loader"]["resolution"]
        self.flow_scaling = flow_scaling if flow_scaling is not None else max(config["loader"]["resolution"])
        self.weight = config["loss"]["flow_regul_weight"]
        self.smoothing_mask = False if "mask_output" not in config["model"].keys() else config["model"]["mask_output"]
        self.overwrite_intermediate = (
            False if "overwrite_intermediate" not in config["loss"].keys() else config["loss"]["overwrite_intermediate"]
        )
        self.device = device
        self.reset()

    def reset(self):
        self._passes = 0
        self._events = []        # per pass [B,N_t,4] (ts already offset by the pass index)
        self._pol_masks = []     # per pass [B,N_t,2]
        self._masks = []         # per pass [B,1,H,W]
        self._flow_maps = None   # per scale: list over passes of [B,2,H,W]
        self._num_events = 0
        self._overwritten = False

    def event_flow_association(self, flow_list, event_list, pol_mask, event_mask):
        """
        :param flow_list: [[batch_size x 2 x H x W]] list of optical flow (x, y) maps
        :param event_list: [batch_size x N x 4] input events (ts, y, x, p)
        :param pol_mask: [batch_size x N x 2] polarity mask (pos, neg)
        :param event_mask: [batch_size x 1 x H x W] event mask
        """
        if self._flow_maps is None:
            self._flow_maps = [[] for _ in flow_list]
        for i, flow in enumerate(flow_list):
            self._flow_maps[i].append(flow)
        if self._passes > 0:
            event_list[:, :, 0:1] += self._passes  # in place on the caller's tensor, like loss/flow.py:90
        self._events.append(event_list)
        self._pol_masks.append(pol_mask)
        self._masks.append(event_mask)
        self._num_events += event_list.shape[1]
        self._passes += 1

    def overwrite_intermediate_flow(self, flow_list):
        """loss/flow.py:118-146: every event of the window is warped with the final flow map(s)."""
        self._flow_maps = [[flow] for flow in flow_list]
        mask = torch.sum(torch.cat(self._masks, dim=1), dim=1, keepdim=True)
        mask[mask > 1] = 1
        self._masks = [mask]
        self._overwritten = True

    @property
    def num_events(self):
        return self._num_events

    @property
    def event_mask(self):
        if self.overwrite_intermediate:
            return torch.cat(self._masks, dim=1)
        return self._masks[-1]

    def forward(self):
        """
        loss/flow.py:176-301 in ONE kernel launch (ef_iwe_loss_fwd_passes), differentiable wrt the flow maps (one more launch,
        ef_iwe_loss_bwd_passes).  The window is handed over in pass form: the tensors event_flow_association received stay where
        they are -- no torch.cat / torch.stack (the reference re-concatenates its lists every pass).
        """
        f32 = lambda ts: [t if t.dtype == torch.float32 else t.float() for t in ts]  # noqa: E731
        return ops.event_warping_loss_passes(
            [f32(per_scale) for per_scale in self._flow_maps],
            f32(self._events),
            f32(self._pol_masks),
            f32(self._masks),
            flow_scaling=self.flow_scaling,
            weight=self.weight,
            loss_scaling=self.loss_scaling,
            smoothing_mask=self.smoothing_mask,
            overwrite_intermediate=self._overwritten,
        )


class BaseValidationLoss(torch.nn.Module):
    """
    Bookkeeping of the validation metrics (loss/flow.py:304-465): accumulates the window in map form like EventWarping;
    the metric itself is one call into libeventflow.so.
    """

    def __init__(self, config, device, flow_scaling=128):
        super().__init__()
        self.res = config["loader"]["resolution"]
        self.flow_scaling = flow_scaling  # should be specified by the user
        self.overwrite_intermediate = (
            False if "overwrite_intermediate" not in config["loss"].keys() else config["loss"]["overwrite_intermediate"]
        )
        self.device = device
        self.reset()

    def reset(self):
        self._passes = 0
        self._events, self._pol_masks, self._masks, self._flow_map = [], [], [], []
        self._num_events = 0
        self._overwritten = False
        self._gtflow = None

    @property
    def num_events(self):
        return self._num_events

    def event_flow_association(self, flow_list, inputs):
        """
        :param flow_list: [[batch_size x 2 x H x W]] list of optical flow (x, y) maps
        :param inputs: dataloader dictionary
        """
        event_list = inputs["event_list"].to(self.device)
        pol_mask = inputs["event_list_pol_mask"].to(self.device)
        event_mask = inputs["event_mask"].to(self.device)
        self._gtflow = inputs["gtflow"].to(self.device) if "gtflow" in inputs.keys() else None
        flow = flow_list[-1]  # only highest resolution flow
        if self._passes > 0:
            event_list = event_list.clone()  # to prevent issues with other metrics (loss/flow.py:367)
            event_list[:, :, 0:1] += self._passes
        self._events.append(event_list)
        self._pol_masks.append(pol_mask)
        self._masks.append(event_mask)
        self._flow_map.append(flow.view(flow.shape[0], 2, self.res[0], self.res[1]))
        self._num_events += event_list.shape[1]
        self._dt_input = inputs["dt_input"]
        self._dt_gt = inputs["dt_gt"]
        self._passes += 1

    def overwrite_intermediate_flow(self, flow_list):
        flow = flow_list[-1]
        self._flow_map = [flow.view(flow.shape[0], 2, self.res[0], self.res[1])]
        mask = torch.sum(torch.cat(self._masks, dim=1), dim=1, keepdim=True)
        mask[mask > 1] = 1
        self._masks = [mask]
        self._overwritten = True

    def _window(self):
        events = self._events[0] if len(self._events) == 1 else torch.cat(self._events, dim=1)
        pol = self._pol_masks[0] if len(self._pol_masks) == 1 else torch.cat(self._pol_masks, dim=1)
        maps = torch.stack(self._flow_map, dim=1)  # [B,Tm,2,H,W]
        counts = [e.shape[1] for e in self._events]
        offsets = None
        if not self._overwritten and len(set(counts)) > 1:
            offsets = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32, device=events.device)
        return events.float(), pol.float(), maps.float(), counts[0], offsets

    def compute_window_events(self):
        """Per-polarity image of the (non-warped) events of the window (loss/flow.py:425-435)."""
        events, pol, _, _, _ = self._window()
        zero = torch.zeros(events.shape[0], 2, self.res[0], self.res[1], device=events.device)
        return ops.iwe_image(events, pol, self.res, flow=zero, tref=float(self._passes), flow_scaling=self.flow_scaling, round_idx=True)

    def compute_masked_window_flow(self):
        """loss/flow.py:437-447 (visualisation helper: elementwise tensor arithmetic on the flow maps)."""
        masks = torch.cat(self._masks, dim=1)
        if self.overwrite_intermediate:
            return self._flow_map[-1] * masks
        avg_flow = self._flow_map[0] * masks[:, 0:1]
        for i in range(1, masks.shape[1]):
            avg_flow = avg_flow + self._flow_map[i] * masks[:, i:i + 1]
        return avg_flow / (torch.sum(masks, dim=1, keepdim=True) + 1e-9)

    def compute_window_iwe(self, round_idx=True):
        """Per-polarity image of warped events of the window (loss/flow.py:449-465)."""
        events, pol, maps, n0, offsets = self._window()
        B, N = events.shape[:2]
        # per-event flow gathered from the map of the event's own pass (as event_flow_association does in the reference)
        if maps.shape[1] == 1:
            t_idx = torch.zeros(N, dtype=torch.long, device=events.device)
        elif offsets is not None:
            t_idx = torch.bucketize(torch.arange(N, device=events.device), offsets[1:].long(), right=True).clamp(max=maps.shape[1] - 1)
        else:
            t_idx = (torch.arange(N, device=events.device) // n0).clamp(max=maps.shape[1] - 1)
        flat = (events[:, :, 1] * self.res[1] + events[:, :, 2]).long()
        f = maps.reshape(B, maps.shape[1], 2, -1)
        bi = torch.arange(B, device=events.device).view(B, 1).expand(B, N)
        ev_flow = torch.stack([f[bi, t_idx.view(1, N).expand(B, N), 1, flat], f[bi, t_idx.view(1, N).expand(B, N), 0, flat]], dim=2)
        return ops.iwe_image(events, pol, self.res, event_flow=ev_flow, tref=float(self._passes), flow_scaling=self.flow_scaling,
                             round_idx=round_idx)


class FWL(BaseValidationLoss):
    """Flow Warp Loss (loss/flow.py:468-500): spatial variance of the IWE over that of the image of events, per sample."""

    def forward(self):
        events, pol, maps, n0, offsets = self._window()
        fwl, _ = ops.iwe_metrics(maps, events, pol, passes=self._passes, n_per_pass=n0, flow_scaling=self.flow_scaling, pass_offsets=offsets)
        return fwl


class RSAT(BaseValidationLoss):
    """Ratio of the squared averaged timestamps (loss/flow.py:503-579), per sample."""

    def forward(self):
        events, pol, maps, n0, offsets = self._window()
        _, rsat = ops.iwe_metrics(maps, events, pol, passes=self._passes, n_per_pass=n0, flow_scaling=self.flow_scaling, pass_offsets=offsets)
        return rsat


class AEE(BaseValidationLoss):
    """Average endpoint error and outlier percentage (loss/flow.py:582-628)."""

    @property
    def num_events(self):
        return float("inf")

    def forward(self):
        masks = torch.cat(self._masks, dim=1)
        dt_ratio = torch.as_tensor(self._dt_gt, dtype=torch.float32) / torch.as_tensor(self._dt_input, dtype=torch.float32)
        return ops.aee(self._flow_map[-1], self._gtflow, masks[:, -1], dt_ratio, self.flow_scaling)
