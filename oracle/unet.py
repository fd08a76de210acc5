"""
CPU restatement (torch fp32) of the spiking multi-resolution recurrent U-Net.  TEST INFRASTRUCTURE ONLY: nothing under
event_flow_b200/ imports this file.  Pinned against the unmodified reference by oracle/pin_against_reference.py (pin_unet).

Follows models/unet.py:436-465 (SpikingMultiResUNetRecurrent.forward), the layer blocks of
models/spiking_submodules.py:878-1013 (SpikingRecurrentConvLayer, SpikingResidualBlock, SpikingUpsampleConvLayer),
models/model_util.py:14-19 (skip_concat) and the flow upsampling of models/model.py:525-539 (RecEVFlowNet.forward).
"""
import torch
import torch.nn.functional as F

from . import spiking as osp

CELL_PARAM_NAMES = {
    "lif": ("leak", "thresh"),
    "plif": ("leak_v", "leak_pt", "add_pt", "thresh"),
    "alif": ("leak_v", "leak_t", "t0", "t1"),
    "xlif": ("leak_v", "leak_pt", "t0", "t1"),
}


def cell_params(sd, prefix, neuron):
    """Parameter dict of one cell (oracle.spiking.cell_step format) from a state_dict."""
    p = {"ff": sd[prefix + "ff.weight"]}
    if prefix + "rec.weight" in sd:
        p["rec"] = sd[prefix + "rec.weight"]
    for n in CELL_PARAM_NAMES[neuron]:
        p[n] = sd[prefix + n]
    return p


def unet_params(sd, neuron, num_encoders=4, num_residual_blocks=2, prefix="multires_unetrec."):
    """Groups a RecEVFlowNet state_dict (reference key names) by layer."""
    P = {"enc": [], "res": [], "dec": [], "pred": []}
    for i in range(num_encoders):
        P["enc"].append((cell_params(sd, f"{prefix}encoders.{i}.conv.", neuron), cell_params(sd, f"{prefix}encoders.{i}.recurrent_block.", neuron)))
    for i in range(num_residual_blocks):
        P["res"].append((cell_params(sd, f"{prefix}resblocks.{i}.conv1.", neuron), cell_params(sd, f"{prefix}resblocks.{i}.conv2.", neuron)))
    for i in range(num_encoders):
        P["dec"].append(cell_params(sd, f"{prefix}decoders.{i}.conv2d.", neuron))
        P["pred"].append((sd[f"{prefix}preds.{i}.conv2d.weight"], sd[f"{prefix}preds.{i}.conv2d.bias"]))
    return P


def skip_concat(x1, x2):
    """models/model_util.py:14-19."""
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
    return torch.cat([x1, x2], dim=1)


def unet_step(neuron, P, states, x, trace=None, **cell_kwargs):
    """
    One forward pass.  states: list of 2*E + R entries (None or stacked tensors exactly like the reference keeps them).
    trace: optional list that receives (name, cell_input, state_in, out, state_out) of every cell, in execution order.
    Returns (multires predictions coarse->fine, flows upsampled to the input resolution, new states).
    """
    E, R = len(P["enc"]), len(P["res"])
    new_states = list(states)

    def run(name, p, xin, st, stride=1, residual=0):
        out, st_out = osp.cell_step(neuron, xin, st, p, stride=stride, residual=residual, **cell_kwargs)
        if trace is not None:
            trace.append((name, xin, st, out, st_out))
        return out, st_out

    blocks = []
    for i, (pc, pr) in enumerate(P["enc"]):  # spiking_submodules.py:922-930
        ff, rec = (None, None) if states[i] is None else states[i]
        x1, ff = run(f"encoders.{i}.conv", pc, x, ff, stride=2)
        x, rec = run(f"encoders.{i}.recurrent_block", pr, x1, rec)
        new_states[i] = torch.stack([ff, rec])
        blocks.append(x)
    for i, (p1, p2) in enumerate(P["res"]):  # spiking_submodules.py:965-975
        c1, c2 = (None, None) if states[E + i] is None else states[E + i]
        x1, c1 = run(f"resblocks.{i}.conv1", p1, x, c1)
        x, c2 = run(f"resblocks.{i}.conv2", p2, x1, c2, residual=x)
        new_states[E + i] = torch.stack([c1, c2])
    preds = []
    for i, pd in enumerate(P["dec"]):  # unet.py:456-463
        x = skip_concat(x, blocks[E - i - 1])
        if i > 0:
            x = skip_concat(preds[-1], x)
        x_up = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        x, new_states[E + R + i] = run(f"decoders.{i}.conv2d", pd, x_up, states[E + R + i])
        w, b = P["pred"][i]
        preds.append(osp.pred_head(x, w, b))
    flows = [F.interpolate(f, scale_factor=(preds[-1].shape[2] / f.shape[2], preds[-1].shape[3] / f.shape[3])) for f in preds]
    return preds, flows, new_states


def ann_unet_forward(sd, x, act="relu", num_encoders=4, num_residual_blocks=2, prefix="multires_unet.", trace=None, states=None):
    """
    EV-FlowNet's ANN U-Net (models/unet.py:224-311 with the blocks of models/submodules.py:12-61, 140-185, 238-312):
    stride-2 conv+bias+ReLU encoders, residual blocks, bilinear-upsampling decoders with concat skips, 1x1 tanh predictions.
    Returns (multires predictions, flows upsampled to the input resolution as in models/model.py:370-383).
    trace: optional list receiving (name, input, output) of every 3x3 conv layer.
    states: None for EV-FlowNet; for the recurrent variant (RecEVFlowNet, models/unet.py:314-416 with RecurrentConvLayer,
    submodules.py:188-235) a list of num_encoders ConvGRU hidden states (None = zeros), updated IN PLACE; the state_dict prefix
    is then "multires_unetrec." and every encoder is conv (stride 2) + ConvGRU.
    """
    f_act = {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, None: (lambda t: t)}[act]

    def conv(name, t, stride=1, residual=None):
        out = F.conv2d(t, sd[prefix + name + ".weight"], sd[prefix + name + ".bias"], stride, 1)
        if residual is not None:
            out = out + residual
        out = f_act(out)
        if trace is not None:
            trace.append((name, t, out))
        return out

    blocks = []
    for i in range(num_encoders):
        if states is None:
            x = conv(f"encoders.{i}.conv2d", x, stride=2)
        else:
            x = conv(f"encoders.{i}.conv.conv2d", x, stride=2)
            g = prefix + f"encoders.{i}.recurrent_block."
            gp = {"update_w": sd[g + "update_gate.weight"], "update_b": sd[g + "update_gate.bias"], "reset_w": sd[g + "reset_gate.weight"],
                  "reset_b": sd[g + "reset_gate.bias"], "out_w": sd[g + "out_gate.weight"], "out_b": sd[g + "out_gate.bias"]}
            x = osp.conv_gru_step(x, states[i], gp)
            states[i] = x
        blocks.append(x)
    for i in range(num_residual_blocks):
        out1 = conv(f"resblocks.{i}.conv1", x)
        x = conv(f"resblocks.{i}.conv2", out1, residual=x)
    preds = []
    for i in range(num_encoders):
        x = skip_concat(x, blocks[num_encoders - i - 1])
        if i > 0:
            x = skip_concat(preds[-1], x)
        x = conv(f"decoders.{i}.conv2d", F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))
        preds.append(torch.tanh(F.conv2d(x, sd[prefix + f"preds.{i}.conv2d.weight"], sd[prefix + f"preds.{i}.conv2d.bias"])))
    flows = [F.interpolate(f, scale_factor=(preds[-1].shape[2] / f.shape[2], preds[-1].shape[3] / f.shape[3])) for f in preds]
    return preds, flows
