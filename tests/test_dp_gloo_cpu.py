"""
CPU, world_size 2 over gloo: the data-parallel contract of SURVEY 8e / T7 -- gradients of a batch equal the all-reduced SUM
of the per-shard gradients (no division by the world size), and after the reduce every rank holds the same buffer.  The
CUDA kernels cannot run here, so the per-rank gradients come from the CPU oracle; what is under test is the sharding and
reduction logic bench.py / DataParallelTrainer rely on.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import encodings as oenc
from oracle import iwe as oiwe
from oracle import spiking as osp

H, W, T, N, BINS = 16, 16, 2, 200, 2


def window_grads(params, batch_slice, seed0):
    for lp in params.values():
        for v in lp.values():
            v.grad = None
    states, flows, evs, pms, masks = [None] * 7, [], [], [], []
    for t in range(T):
        d = oenc.encode_window(*oenc.synthetic_events(4, N, H, W, seed0 + t), H, W, BINS)
        d = {k: v[batch_slice] for k, v in d.items()}
        flow, states, _ = osp.firenet_step("lif", params, states, d["event_cnt"])
        flows.append(flow)
        e = d["event_list"].clone()
        e[:, :, 0] += t
        evs.append(e), pms.append(d["event_list_pol_mask"]), masks.append(d["event_mask"])
    loss = oiwe.event_warping_loss(torch.cat(evs, 1), torch.cat(pms, 1), torch.arange(T).repeat_interleave(N), [torch.stack(flows, 1)],
                                   torch.cat(masks, 1), (H, W), weight=0.0, passes=T)  # weight 0: the smoothness term is a batch sum too,
    loss.backward()                                                                  # but its /T normalisation is shared -> keep it simple
    return torch.cat([v.grad.reshape(-1) for lp in params.values() for v in lp.values()]), loss.detach()


def make_params():
    torch.manual_seed(0)
    params = osp.init_firenet_params("lif", BINS, 32, seed=0, weight_gain=2.5)
    params["pred"]["weight"] = params["pred"]["weight"] * 30.0
    for lp in params.values():
        for k in lp:
            lp[k] = lp[k].clone().requires_grad_(True)
    return params


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = make_params()
    shard = slice(rank * 2, rank * 2 + 2)  # batch 4 -> 2 samples per rank
    flat, loss = window_grads(params, shard, 40)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)  # SUM, not MEAN: the loss sums over the batch (loss/flow.py:226,259)
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out["flat"], out["loss"], out["same"] = flat, loss, all(torch.equal(g, flat) for g in gathered)
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_gradients_sum_to_full_batch_gradients():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(2, free_port(), out), nprocs=2, join=True)
    params = make_params()
    full, loss = window_grads(params, slice(0, 4), 40)
    assert out["same"]
    torch.testing.assert_close(out["loss"], loss, rtol=1e-5, atol=0)
    assert full.abs().max() > 0
    torch.testing.assert_close(out["flat"], full, rtol=1e-4, atol=1e-6 * full.abs().max().item())
