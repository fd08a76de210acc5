"""
Data-parallel training loop around the hot path: what train_flow.py:97-171 does for one process, for one rank of a
`torchrun` job.  The batch dimension is sharded over the ranks (each rank draws its own batch slots from the loader), the
model and the loss run on the rank's GPU, and DataParallelTrainer replaces clip_grad_norm_ + optimizer.step() + zero_grad()
with all-reduce(SUM) -> clip -> Adam.  Loop semantics kept from the reference:
  * one loader item = one timestep for all batch slots; `new_seq` (any slot starting a new recording) resets the loss window,
    ALL neuron states and the gradients (train_flow.py:100-105);
  * the loss fires when `loss_function.num_events >= window_loss` (:141), optionally after overwrite_intermediate_flow (:144-145);
  * after the optimiser step: `model.detach_states()` and `loss_function.reset()` (:170-171) -- truncated BPTT.
The loader is any iterable of batch dicts with the reference's keys (dataloader/h5.py:330-341); `SyntheticEventStream` is the
stand-in used by the benchmarks and tests (the HDF5 loader itself is out of scope).
"""
import torch
import torch.distributed as dist

from .dataloader.encodings import encode_batch
from .parallel import DataParallelTrainer


class SyntheticEventStream:
    """
    Per-rank synthetic event stream with the H5Loader's contract: iterating yields one timestep (window of `n_events` events per
    batch slot) as a dict of device tensors; `.new_seq` is True for the first item of a sequence (h5.py:51-54,184).  Seeds follow
    SURVEY 8d (1234 + 1000*rank + step), so the ranks see disjoint shards and a single process can re-create every rank's data.
    Raw events are generated on the host (pinned) and encoded on the device by ef_encode_events -- the loader's CPU encodings
    (dataloader/encodings.py) are not on the training path here.
    """

    def __init__(self, batch_size, n_events, resolution, num_bins, device, rank=0, seq_len=None, n_items=1000, seed=1234):
        self.batch_size, self.n_events, self.res, self.num_bins = batch_size, n_events, tuple(resolution), num_bins
        self.device, self.rank, self.seq_len, self.n_items, self.seed = device, rank, seq_len, n_items, seed
        self.new_seq = False
        self.step = 0

    def host_events(self, step):
        H, W = self.res
        g = torch.Generator().manual_seed(self.seed + 1000 * self.rank + step)
        B, N = self.batch_size, self.n_events
        ts = torch.sort(torch.rand(B, N, generator=g))[0]
        ts = (ts - ts[:, :1]) / (ts[:, -1:] - ts[:, :1])
        ys = torch.randint(0, H, (B, N), generator=g).float()
        xs = torch.randint(0, W, (B, N), generator=g).float()
        ps = (torch.randint(0, 2, (B, N), generator=g) * 2 - 1).float()
        return torch.stack([ts, ys, xs, ps], dim=2)

    def __len__(self):
        return self.n_items

    def __iter__(self):
        for _ in range(self.n_items):
            self.new_seq = self.step == 0 or (self.seq_len is not None and self.step % self.seq_len == 0)
            ev = self.host_events(self.step)
            if torch.device(self.device).type == "cuda":
                ev = ev.pin_memory().to(self.device, non_blocking=True)
            d = encode_batch(ev, self.res, self.num_bins)
            d["event_list"] = ev
            self.step += 1
            yield d


def _finish_window(model, loss_function, trainer, last_flow, overwrite_intermediate, losses, log):
    if overwrite_intermediate:
        loss_function.overwrite_intermediate_flow(last_flow)
    loss = loss_function()
    loss.backward()
    trainer.step()  # all-reduce(SUM) -> clip -> Adam -> zero grads
    model.detach_states()
    loss_function.reset()
    value = loss.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(value, op=dist.ReduceOp.SUM)  # for logging only: the global loss is the sum over the shards
    losses.append(value)
    if log is not None:
        log(len(losses), value)


def train_windows(model, loss_function, trainer, loader, window_loss, n_windows, overwrite_intermediate=False, log=None, staged=False):
    """
    train_flow.py:97-171 for `n_windows` loss windows.  `loader` yields batch dicts and exposes `.new_seq`; `trainer` is a
    DataParallelTrainer (or anything with step() / zero_grad()).  Returns the list of (summed over ranks) window losses.

    staged=True: the loader items of a loss window are STAGED (the loss fires on the event count alone, :141, so the window's extent is
    known without running the model) and the model runs the whole window through `model.forward_window` -- layer-major, the time loop
    inside the kernels of the feed-forward cells.  Same arithmetic, same losses and gradients as the step-by-step loop; a `new_seq` in
    the middle of a window drops the staged steps exactly like the reference drops its partial window (:100-105).
    """
    losses = []
    model.train()
    staged = staged and hasattr(model, "forward_window")
    pending, n_pending = [], 0
    for inputs in loader:
        if getattr(loader, "new_seq", False):
            loss_function.reset()
            model.reset_states()
            trainer.zero_grad()
            pending, n_pending = [], 0
        if staged:
            pending.append(inputs)
            n_pending += inputs["event_list"].shape[1]
            if n_pending < window_loss:
                continue
            outs = model.forward_window(torch.stack([d["event_voxel"] for d in pending]), torch.stack([d["event_cnt"] for d in pending]))
            for d, x in zip(pending, outs):
                loss_function.event_flow_association(x["flow"], d["event_list"], d["event_list_pol_mask"], d["event_mask"])
            pending, n_pending = [], 0
            _finish_window(model, loss_function, trainer, outs[-1]["flow"], overwrite_intermediate, losses, log)
        else:
            x = model(inputs["event_voxel"], inputs["event_cnt"])
            loss_function.event_flow_association(x["flow"], inputs["event_list"], inputs["event_list_pol_mask"], inputs["event_mask"])
            if loss_function.num_events >= window_loss:
                _finish_window(model, loss_function, trainer, x["flow"], overwrite_intermediate, losses, log)
        if len(losses) >= n_windows:
            break
    return [v.item() for v in losses]


def build_trainer(model, config):
    """Optimiser settings of the reference's yml (configs/train_SNN.yml: optimizer.lr, loss.clip_grad)."""
    return DataParallelTrainer(model, lr=config["optimizer"]["lr"], clip_grad=config["loss"].get("clip_grad", 100.0))
