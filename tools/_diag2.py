import sys, torch
sys.path.insert(0, "/root/repo")
from event_flow_b200 import ops, _lib as L
from oracle import spiking as osp
DEV="cuda"
B,H,W=8,128,128
g=torch.Generator().manual_seed(B*7+H)
params=osp.init_firenet_params("lif",32,32,seed=B*7+H,weight_gain=2.0)["R1a"]
x=(torch.rand((B,32,H,W),generator=g)<0.3).float()
st=torch.rand((2,B,32,H,W),generator=g)*1.2-0.1
st[1]=(st[1]<0.3).float()
pd={k:v.to(DEV).contiguous() for k,v in params.items()}
x_cl=ops.pack_cl(x.to(DEV)); v_in=st[0].to(DEV).contiguous(); z_in=ops.pack_cl(st[1].to(DEV))
ws=ops.split_weights(pd["ff"],None)
leak,thresh=pd["leak"].reshape(-1),pd["thresh"].reshape(-1)
v_cc,_=ops.lif_step_cl(x_cl,v_in,z_in,pd["ff"],None,leak,thresh,hard_reset=False,w_split=None)
for mask in (0, 4096, 4096, 1024, 2048, 0):
    L.lib().ef_debug_tc_skip(mask)
    nb=[]
    for trial in range(10):
        v_tc,z_tc=ops.lif_step_cl(x_cl,v_in,z_in,pd["ff"],None,leak,thresh,hard_reset=False,w_split=ws)
        torch.cuda.synchronize()
        nb.append(int(((v_tc-v_cc).abs()>1e-3).sum()))
    print("mask",mask,"bad counts",nb)
L.lib().ef_debug_tc_skip(0)
