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
        T = self._passes
        events = self._events[0] if T == 1 else torch.cat(self._events, dim=1)
        pol = self._pol_masks[0] if T == 1 else torch.cat(self._pol_masks, dim=1)
        masks = self._masks[0] if len(self._masks) == 1 else torch.cat(self._masks, dim=1)
        flow_maps = torch.stack([torch.stack(per_pass, dim=1) for per_pass in self._flow_maps], dim=0)  # [S,B,Tm,2,H,W]
        overwrite = self._overwritten
        counts = [e.shape[1] for e in self._events]
        offsets = None
        if not overwrite and len(set(counts)) > 1:
            offsets = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32, device=events.device)
        return ops.event_warping_loss(
            flow_maps.float(),
            events.float(),
            pol.float(),
            masks.float(),
            passes=T,
            n_per_pass=counts[0],
            flow_scaling=self.flow_scaling,
            weight=self.weight,
            loss_scaling=self.loss_scaling,
            smoothing_mask=self.smoothing_mask,
            overwrite_intermediate=overwrite,
            pass_offsets=offsets,
        )
