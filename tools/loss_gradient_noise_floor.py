import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import encodings as oenc, iwe as oiwe, spiking as osp
from tests.util import rel_err
torch.set_num_threads(8)
B,H,W,T,N,bins=4,128,128,10,1000,5
params=osp.init_firenet_params("lif",bins,32,seed=0,weight_gain=2.5)
params["pred"]["weight"]=params["pred"]["weight"]*20
data=[oenc.encode_window(*oenc.synthetic_events(B,N,H,W,7000+t),H,W,bins) for t in range(T)]
def run(dtype, forced=None):
    leaves={l:{k:v.detach().clone().to(dtype).requires_grad_(True) for k,v in lp.items()} for l,lp in params.items()}
    states=[None]*7; flows=[]; spikes=[]
    for t,d in enumerate(data):
        flow,states,acts=osp.firenet_step("lif",leaves,states,d["event_voxel"].to(dtype), forced=None if forced is None else forced[t])
        flows.append(flow); spikes.append([s[1].detach() for s in states])
    for f in flows: f.retain_grad()
    evs=[]
    for t,d in enumerate(data):
        e=d["event_list"].clone().to(dtype); e[:,:,0]+=t; evs.append(e)
    loss=oiwe.event_warping_loss(torch.cat(evs,1),torch.cat([d["event_list_pol_mask"] for d in data],1).to(dtype),torch.arange(T).repeat_interleave(N),[torch.stack(flows,1)],torch.cat([d["event_mask"] for d in data],1).to(dtype),(H,W),weight=0.001,passes=T)
    loss.backward()
    return loss.item(), leaves, [f.grad for f in flows], spikes, [f.detach() for f in flows]
l32,p32,gf32,sp,fl32=run(torch.float32)
l64,p64,gf64,_,fl64=run(torch.float64, forced=sp)
print("loss",l32,l64)
for t in range(T):
    print(t,"flow rel",rel_err(fl32[t].double(),fl64[t]) if fl64[t].abs().max()>0 else 0,"g_flow rel",rel_err(gf32[t].double(),gf64[t]), "max|gflow|",gf64[t].abs().max().item())
for l in p32:
    print(l," ".join(f"{k}={rel_err(p32[l][k].grad.double(),p64[l][k].grad):.1e}" for k in p32[l] if p64[l][k].grad is not None and p64[l][k].grad.abs().max()>0))
