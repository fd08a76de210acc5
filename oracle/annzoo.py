"""
CPU restatement (torch fp32) of the remaining ANN cell zoo and the models built from it.  TEST INFRASTRUCTURE
ONLY.  Pinned against the unmodified reference by oracle/pin_against_reference.py (pin_ann_zoo).

Cells: models/submodules.py:314-374 (ConvLSTM), :421-451 (ConvRecurrent), :454-499 (ConvLeakyRecurrent), :502-554 (ConvLeaky),
blocks :557-686.  Models: models/model.py:29-145 (E2VID -> unet.py:148-222 UNetRecurrent), :594-633 and :696-704 (RNN / Leaky
FireNets and EV-FlowNets), unet.py:314-480.
Everything is driven by a state_dict with the reference's key names.
"""
import torch
import torch.nn.functional as F

from . import unet as ounet

_ACT = {None: (lambda t: t), "relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}


def conv(sd, name, x, stride=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride, 1)


def conv_layer(sd, name, x, act="relu", stride=1, residual=0):
    return _ACT[act](conv(sd, name + ".conv2d", x, stride) + residual)


def conv_recurrent(sd, name, x, h):
    if h is None:
        h = torch.zeros(x.shape[0], sd[name + ".ff.weight"].shape[0], *x.shape[2:], dtype=x.dtype)
    state = torch.tanh(conv(sd, name + ".ff", x) + conv(sd, name + ".rec", h))
    return torch.relu(conv(sd, name + ".out", state)), state


def conv_leaky(sd, name, x, s, act="relu", stride=1, residual=0):
    ff = conv(sd, name + ".ff", x, stride)
    if s is None:
        s = torch.zeros_like(ff)
    leak = torch.sigmoid(sd[name + ".leak"])
    state = s * leak + (1 - leak) * (ff + residual)
    return _ACT[act](state), state


def conv_leaky_recurrent(sd, name, x, s):
    ff = conv(sd, name + ".ff", x)
    if s is None:
        s = torch.zeros_like(ff)
    leak = torch.sigmoid(sd[name + ".leak"])
    state = torch.tanh(s * leak + (1 - leak) * (ff + conv(sd, name + ".rec", s)))
    return torch.relu(conv(sd, name + ".out", state)), state


def conv_lstm(sd, name, x, state):
    C = sd[name + ".Gates.weight"].shape[0] // 4
    if state is None:
        z = torch.zeros(x.shape[0], C, *x.shape[2:], dtype=x.dtype)
        state = (z, z)
    h, c = state
    i, r, o, g = conv(sd, name + ".Gates", torch.cat((x, h), 1)).chunk(4, 1)
    cell = torch.sigmoid(r) * c + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(cell), cell


FIRE = ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b")


def firenet_zoo_step(kind, sd, states, x, ff_act="relu", rec_act=None):
    """models/model.py:254-265 for kind in {"rnn", "leaky", "leakyflow"}.  Returns (flow, new states)."""
    new, h = [], x
    for i, name in enumerate(FIRE):
        rec = name in ("G1", "G2")
        if kind == "rnn":
            if rec:
                h, s = conv_recurrent(sd, name, h, states[i])
            else:
                h, s = conv_layer(sd, name, h, ff_act), states[i] if states[i] is not None else torch.tensor(0)
        elif kind == "leaky" and rec:
            h, s = conv_leaky_recurrent(sd, name, h, states[i])
        else:
            h, s = conv_leaky(sd, name, h, states[i], rec_act if rec else ff_act)
        new.append(s)
    flow = torch.tanh(F.conv2d(h, sd["pred.conv2d.weight"], sd["pred.conv2d.bias"]))
    return flow, new


def rnn_unet_forward(sd, x, states, prefix="multires_unetrec."):
    """RNNRecEVFlowNet: ounet.ann_unet_forward with ConvRecurrent after every encoder conv (unet.py:314-416)."""
    E = len(states)
    blocks = []
    for i in range(E):
        x = torch.relu(conv(sd, prefix + f"encoders.{i}.conv.conv2d", x, 2))
        x, states[i] = conv_recurrent(sd, prefix + f"encoders.{i}.recurrent_block", x, states[i])
        blocks.append(x)
    return _ann_tail(sd, x, blocks, prefix)


def _ann_tail(sd, x, blocks, prefix):
    E = len(blocks)
    for i in range(2):
        out1 = torch.relu(conv(sd, prefix + f"resblocks.{i}.conv1", x))
        x = torch.relu(conv(sd, prefix + f"resblocks.{i}.conv2", out1) + x)
    preds = []
    for i in range(E):
        x = ounet.skip_concat(x, blocks[E - i - 1])
        if i > 0:
            x = ounet.skip_concat(preds[-1], x)
        x = torch.relu(conv(sd, prefix + f"decoders.{i}.conv2d", F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)))
        preds.append(torch.tanh(conv_1x1(sd, prefix + f"preds.{i}.conv2d", x)))
    return [F.interpolate(f, scale_factor=(preds[-1].shape[2] / f.shape[2], preds[-1].shape[3] / f.shape[3])) for f in preds]


def conv_1x1(sd, name, x):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"])


def leaky_unet_step(sd, states, x, prefix="multires_unetrec."):
    """LeakyRecEVFlowNet (unet.py:436-480 wiring with the leaky blocks).  states: list of 10, updated in place.  Returns flows."""
    E, R = 4, 2
    blocks = []
    for i in range(E):
        ff, rec = (None, None) if states[i] is None else states[i]
        x1, ff = conv_leaky(sd, prefix + f"encoders.{i}.conv", x, ff, "relu", stride=2)
        x, rec = conv_leaky_recurrent(sd, prefix + f"encoders.{i}.recurrent_block", x1, rec)
        states[i] = torch.stack([ff, rec])
        blocks.append(x)
    for i in range(R):
        c1, c2 = (None, None) if states[E + i] is None else states[E + i]
        x1, c1 = conv_leaky(sd, prefix + f"resblocks.{i}.conv1", x, c1, "relu")
        x, c2 = conv_leaky(sd, prefix + f"resblocks.{i}.conv2", x1, c2, "relu", residual=x)
        states[E + i] = torch.stack([c1, c2])
    preds = []
    for i in range(E):
        x = ounet.skip_concat(x, blocks[E - i - 1])
        if i > 0:
            x = ounet.skip_concat(preds[-1], x)
        x, states[E + R + i] = conv_leaky(sd, prefix + f"decoders.{i}.conv2d", F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False),
                                          states[E + R + i], "relu")
        preds.append(torch.tanh(conv_1x1(sd, prefix + f"preds.{i}.conv2d", x)))
    return [F.interpolate(f, scale_factor=(preds[-1].shape[2] / f.shape[2], preds[-1].shape[3] / f.shape[3])) for f in preds]


def e2vid_step(sd, states, x, prefix="unetrecurrent."):
    """E2VID (unet.py:194-222): states = list of 3 (hidden, cell) tuples or None, updated in place.  Returns the flow map."""
    x = torch.relu(conv(sd, prefix + "head.conv2d", x))
    head, blocks = x, []
    for i in range(3):
        x = torch.relu(conv(sd, prefix + f"encoders.{i}.conv.conv2d", x, 2))
        x, cell = conv_lstm(sd, prefix + f"encoders.{i}.recurrent_block", x, states[i])
        states[i] = (x, cell)
        blocks.append(x)
    for i in range(2):
        out1 = torch.relu(conv(sd, prefix + f"resblocks.{i}.conv1", x))
        x = torch.relu(conv(sd, prefix + f"resblocks.{i}.conv2", out1) + x)
    for i in range(3):
        x = x + blocks[3 - i - 1]  # skip_sum (same sizes here)
        x = torch.relu(conv(sd, prefix + f"decoders.{i}.conv2d", F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)))
    return torch.tanh(conv_1x1(sd, prefix + "pred.conv2d", x + head))
