"""
Data-parallel training launcher for the hot path: the loop of train_flow.py:97-171 on a synthetic event stream,
one process per GPU.

    python tools/train_dp.py --windows 20                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_dp.py --windows 20 --batch-size 64                       # batch 64 sharded 8 x 8

Model / loss / optimiser settings are those of configs/train_SNN.yml (LIFFireNet, voxel encoding with 5 bins, window 1000
events, window_loss 10000, Adam lr 2e-4, clip 100); --model picks any class of event_flow_b200.models.model.
Rank 0 prints one line per loss window (loss summed over all ranks = the loss of the global batch) and a final JSON summary.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def model_config(name, num_bins):
    """
    config["model"] for any class of event_flow_b200.models.model.  spiking_neuron: the LIF settings of configs/train_SNN.yml for the LIF
    models (the other neuron kinds have their own keyword sets: the class defaults), None for the ANN models, whose activations are
    those of the reference's ANN configs.
    """
    lif = name.startswith("LIF") or name == "SpikingRecEVFlowNet"
    spiking = lif or any(name.startswith(k) for k in ("PLIF", "ALIF", "XLIF"))
    neuron = dict(leak=[-4.0, 0.1], thresh=[0.8, 0.1], learn_leak=True, learn_thresh=True, hard_reset=True) if lif else ({} if spiking else None)
    return dict(name=name, encoding="voxel", round_encoding=False, norm_input=False, num_bins=num_bins, base_num_channels=32, kernel_size=3,
                activations=["arctanspike", "arctanspike"] if spiking else ["relu", None], mask_output=True, spiking_neuron=neuron)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="LIFFireNet")
    ap.add_argument("--batch-size", type=int, default=None, help="GLOBAL batch size (default 8 per rank)")
    ap.add_argument("--windows", type=int, default=20, help="loss windows (optimiser steps) to run")
    ap.add_argument("--window", type=int, default=1000, help="events per timestep and sample (data.window)")
    ap.add_argument("--window-loss", type=int, default=10000, help="events per loss window and sample (data.window_loss)")
    ap.add_argument("--resolution", type=int, nargs=2, default=[128, 128])
    ap.add_argument("--num-bins", type=int, default=5)
    ap.add_argument("--seq-len", type=int, default=None, help="timesteps per synthetic recording (new_seq resets states); default: one long recording")
    ap.add_argument("--lr", type=float, default=2e-4)
    ap.add_argument("--stepwise", action="store_true", help="call the model once per timestep (the reference's loop) instead of staging loss windows")
    ap.add_argument("--weight-gain", type=float, default=2.5, help="conv-weight gain so that spikes propagate on synthetic events")
    a = ap.parse_args()

    from event_flow_b200 import _lib
    from event_flow_b200.loss.flow import EventWarping
    from event_flow_b200.models import model as M
    from event_flow_b200.train import SyntheticEventStream, build_trainer, train_windows

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.lib().ef_device_ok() == 1, "needs a compute-capability 10.x GPU (B200)"
    global_batch = a.batch_size or 8 * world
    assert global_batch % world == 0, "the global batch must divide over the ranks"
    config = {
        "model": model_config(a.model, a.num_bins),
        "loss": {"flow_regul_weight": 0.001, "clip_grad": 100.0, "overwrite_intermediate": False},
        "optimizer": {"name": "Adam", "lr": a.lr},
        "loader": {"resolution": a.resolution, "batch_size": global_batch},
    }
    torch.manual_seed(0)  # identical initial parameters on every rank
    model = getattr(M, a.model)(dict(config["model"]))
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(a.weight_gain)
        if hasattr(model, "pred"):
            model.pred.conv2d.weight.mul_(20.0)
    model = model.to(dev)
    loss_function = EventWarping(config, dev)
    trainer = build_trainer(model, config)
    loader = SyntheticEventStream(global_batch // world, a.window, a.resolution, a.num_bins, dev, rank=rank, seq_len=a.seq_len,
                                  n_items=(a.windows + 1) * (a.window_loss // a.window))

    def log(i, v):
        if rank == 0:
            print(f"window {i:4d}  loss {v.item():.6f}", flush=True)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    losses = train_windows(model, loss_function, trainer, loader, a.window_loss, a.windows, log=log, staged=not a.stepwise)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        ev = global_batch * a.window_loss * len(losses)
        print(json.dumps({"model": a.model, "world_size": world, "global_batch": global_batch, "windows": len(losses), "staged": not a.stepwise, "seconds": dt,
                          "events_per_s_incl_host_generation": ev / dt, "first_loss": losses[0], "last_loss": losses[-1],
                          "param_checksum": float(trainer.flat_param.double().sum().item())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
