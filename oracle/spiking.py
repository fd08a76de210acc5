"""
CPU oracle for hot path A: spiking conv cells and the FireNet chain.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A functional (stateless) restatement in torch-CPU fp32 of
  * the Heaviside spike with surrogate gradients          models/spiking_util.py:13-109
  * the LIF / PLIF / ALIF / XLIF cell updates (ff + rec)   models/spiking_submodules.py:96-126, 191-227, 299-334,
                                                           399-435, 516-551, 618-657, 730-768, 836-875
  * the 1x1 tanh prediction head                           models/submodules.py:52-61, models/model.py:197-199
  * the 7-cell FireNet chain                               models/model.py:229-286
The float-op ORDER of every pointwise expression follows the reference line by line, so the CPU results are bit-equal
to the reference's on the same torch build (checked by oracle/pin_against_reference.py).
"""

import math

import torch
import torch.nn.functional as F

SURROGATES = ("arctanspike", "superspike", "trianglespike", "mgspike")
NEURONS = ("lif", "plif", "alif", "xlif")


def _gaussian(x, mu, sigma):
    # spiking_util.py:6-10
    return torch.exp(-((x - mu) * (x - mu)) / (2 * sigma * sigma)) / (sigma * math.sqrt(2 * math.pi))


def surrogate_grad(x, width, kind):
    """d spike / d (v - thresh) for each surrogate (spiking_util.py:39-43, 56-65, 75-79, 89-93)."""
    if kind == "arctanspike":
        return 1 / (1 + width * x * x)
    if kind == "superspike":
        return 1 / (1 + width * x.abs()) ** 2
    if kind == "trianglespike":
        return F.relu(1 - width * x.abs())
    if kind == "mgspike":
        return (
            1.15 * _gaussian(x, 0.0, width) - 0.15 * _gaussian(x, width, 6 * width) - 0.15 * _gaussian(x, -width, 6 * width)
        )
    raise ValueError(kind)


class _Spike(torch.autograd.Function):
    """Heaviside forward (spiking_util.py:19-21), surrogate backward."""

    @staticmethod
    def forward(ctx, x, width, kind):
        ctx.save_for_backward(x)
        ctx.width, ctx.kind = width, kind
        return x.gt(0).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * surrogate_grad(x, ctx.width, ctx.kind), None, None


def spike(v, thresh, width=10.0, kind="arctanspike"):
    """spiking_util.py:96-109: the surrogate is applied to (v - thresh)."""
    return _Spike.apply(v - thresh, float(width), kind)


def presyn_trace_input(x, ksize=3, stride=1):
    """P(x) = avgpool_k(mean_c |x|), count_include_pad (spiking_submodules.py:164,212)."""
    return F.avg_pool2d(x.abs().mean(1, keepdim=True), ksize, stride, padding=ksize // 2)


def cell_step(
    neuron,
    x,
    state,
    p,
    *,
    hard_reset=None,
    detach=True,
    surrogate="arctanspike",
    width=10.0,
    stride=1,
    residual=0,
    forced_z=None,
):
    """
    One step of a conv spiking cell.
    :param forced_z: teacher forcing for BPTT comparisons: the returned spikes take these VALUES (the spikes another
                     implementation emitted) while the gradient still flows through this step's own surrogate, so a borderline
                     spike that flipped elsewhere cannot cascade into the comparison (threshold chaos).
    :param neuron: "lif" | "plif" | "alif" | "xlif"
    :param x: [B,Cin,H,W]
    :param state: None or stacked [2|3,B,C,H',W'] (v, z[, trace])
    :param p: dict of tensors: "ff" [C,Cin,k,k], optional "rec" [C,C,k,k] and the per-channel [C,1,1] params
              lif: leak, thresh | plif: leak_v, leak_pt, add_pt, thresh | alif: leak_v, leak_t, t0, t1 |
              xlif: leak_v, leak_pt, t0, t1
    :return (out, new_state) exactly like the reference cells.
    """
    if hard_reset is None:
        hard_reset = neuron in ("lif", "plif")  # reference defaults: spiking_submodules.py:51,145 vs 260,358
    k = p["ff"].shape[-1]
    ff = F.conv2d(x, p["ff"], None, stride, k // 2)
    n_state = 2 if neuron == "lif" else 3
    if state is None:
        state = torch.zeros(n_state, *ff.shape, dtype=ff.dtype)
    v, z = state[0], state[1]
    aux = state[2] if n_state == 3 else None

    cur = ff
    if "rec" in p and p["rec"] is not None:
        cur = ff + F.conv2d(z, p["rec"], None, 1, k // 2)  # rec path is differentiable (:530 precedes the detach :539)

    if neuron == "lif":
        thresh = p["thresh"].clamp_min(0.01)
        leak = torch.sigmoid(p["leak"])
        zr = z.detach() if detach else z
        if hard_reset:
            v_out = v * leak * (1 - zr) + (1 - leak) * cur
        else:
            v_out = v * leak + (1 - leak) * cur - zr * thresh
        z_out = spike(v_out, thresh, width, surrogate)
        new_state = torch.stack([v_out, z_out])
    elif neuron == "plif":
        thresh = p["thresh"].clamp_min(0.01)
        leak_v = torch.sigmoid(p["leak_v"])
        leak_pt = torch.sigmoid(p["leak_pt"])
        add_pt = torch.sigmoid(p["add_pt"])
        pt_out = aux * leak_pt + (1 - leak_pt) * presyn_trace_input(x, k, stride)
        zr = z.detach() if detach else z
        if hard_reset:
            v_out = v * leak_v * (1 - zr) + (1 - leak_v) * (cur - add_pt * pt_out)
        else:
            v_out = v * leak_v + (1 - leak_v) * (cur - add_pt * pt_out) - zr * thresh
        z_out = spike(v_out, thresh, width, surrogate)
        new_state = torch.stack([v_out, z_out, pt_out])
    elif neuron == "alif":
        t0 = p["t0"].clamp_min(0.01)
        t1 = p["t1"].clamp_min(0)
        leak_v = torch.sigmoid(p["leak_v"])
        leak_t = torch.sigmoid(p["leak_t"])
        t_out = aux * leak_t + (1 - leak_t) * z  # non-detached z (:317 precedes :322)
        thresh = t0 + t1 * t_out
        zr = z.detach() if detach else z
        if hard_reset:
            v_out = v * leak_v * (1 - zr) + (1 - leak_v) * cur
        else:
            v_out = v * leak_v + (1 - leak_v) * cur - zr * (t0 + t1 * aux)
        z_out = spike(v_out, thresh, width, surrogate)
        new_state = torch.stack([v_out, z_out, t_out])
    elif neuron == "xlif":
        t0 = p["t0"].clamp_min(0.01)
        t1 = p["t1"].clamp_min(0)
        leak_v = torch.sigmoid(p["leak_v"])
        leak_pt = torch.sigmoid(p["leak_pt"])
        pt_out = aux * leak_pt + (1 - leak_pt) * presyn_trace_input(x, k, stride)
        thresh = t0 + t1 * pt_out
        zr = z.detach() if detach else z
        if hard_reset:
            v_out = v * leak_v * (1 - zr) + (1 - leak_v) * cur
        else:
            v_out = v * leak_v + (1 - leak_v) * cur - zr * (t0 + t1 * aux)
        z_out = spike(v_out, thresh, width, surrogate)
        new_state = torch.stack([v_out, z_out, pt_out])
    else:
        raise ValueError(neuron)
    if forced_z is not None:
        z_out = forced_z.detach().to(z_out.dtype) + (z_out - z_out.detach())
        new_state = torch.stack([new_state[0], z_out] + ([new_state[2]] if new_state.shape[0] == 3 else []))
    return z_out + residual, new_state


def pred_head(x, weight, bias):
    """1x1 conv + bias + tanh (models/submodules.py:52-61 with activation="tanh", model.py:197-199)."""
    return torch.tanh(F.conv2d(x, weight, bias))


FIRENET_LAYERS = ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b")
FIRENET_RECURRENT = ("G1", "G2")


def firenet_step(neuron, params, states, x, forced=None, **cell_kwargs):
    """
    One forward pass of a spiking FireNet (model.py:254-265).
    :param params: {"head": {...}, "G1": {... incl. "rec"}, ..., "pred": {"weight","bias"}}
    :param states: list of 7 (None or stacked state)
    :param forced: optional list of 7 spike tensors (teacher forcing, see cell_step)
    :return flow [B,2,H,W], new states list, list of layer outputs (for activity / per-layer parity)
    """
    new_states, acts = [], []
    h = x
    for i, name in enumerate(FIRENET_LAYERS):
        h, s = cell_step(neuron, h, states[i], params[name], forced_z=None if forced is None else forced[i], **cell_kwargs)
        new_states.append(s)
        acts.append(h)
    flow = pred_head(h, params["pred"]["weight"], params["pred"]["bias"])
    return flow, new_states, acts


def firenet_layerwise_check(neuron, params, captured, **cell_kwargs):
    """
    Per-LAYER teacher-forced check of one FireNet step computed elsewhere (the CUDA path).
    :param captured: {layer: (x_in, state_in | None, out, state_out)} CPU tensors captured from the path under test
    :return list of (layer, max|dv| / max(1, max|v|/6.67), spike flips outside the tolerance band around threshold, flips inside, neurons)
            -- the membrane error is reported relative to the magnitude-scaled fp32 summation noise (3e-6*max|v|, floor 2e-5)
    Every layer's oracle output is computed from the inputs the tested path actually fed to that layer, so a single
    borderline spike cannot cascade into the next layer's comparison (threshold chaos).
    """
    report = []
    for name in FIRENET_LAYERS:
        x_in, st_in, out, st_out = captured[name]
        out_o, st_o = cell_step(neuron, x_in, st_in, params[name], **cell_kwargs)
        p = params[name]
        if neuron in ("lif", "plif"):
            thr = p["thresh"].clamp_min(0.01)
        else:
            thr = p["t0"].clamp_min(0.01) + p["t1"].clamp_min(0) * st_o[2]
        dv = (st_out[0] - st_o[0]).abs().max().item() / max(1.0, st_o[0].abs().max().item() * 3e-6 / 2e-5)
        near = (st_o[0] - thr).abs() < max(2e-5, 3e-6 * st_o[0].abs().max().item())  # band = membrane tolerance
        diff = st_out[1] != st_o[1]
        report.append((name, dv, int((diff & ~near).sum()), int((diff & near).sum()), st_o[1].numel()))
    return report


def init_firenet_params(neuron, num_bins, channels=32, ksize=3, seed=0, weight_gain=1.0, thresh=(0.8, 0.1)):
    """
    Random parameters with the reference's initialisers (spiking_submodules.py:60-75,487-490; model.py:197-199 with
    w_scale_pred=0.01).  weight_gain>1 keeps spikes alive through 7 layers on sparse synthetic input.
    """
    g = torch.Generator().manual_seed(seed)

    def U(shape, s):
        return (torch.rand(shape, generator=g) * 2 - 1) * s

    def Nrm(mu, sd):
        return torch.randn(channels, 1, 1, generator=g) * sd + mu

    params = {}
    for name in FIRENET_LAYERS:
        cin = num_bins if name == "head" else channels
        p = {"ff": U((channels, cin, ksize, ksize), math.sqrt(1 / cin)) * weight_gain}
        if name in FIRENET_RECURRENT:
            p["rec"] = U((channels, channels, ksize, ksize), math.sqrt(1 / channels)) * weight_gain
        if neuron == "lif":
            p["leak"], p["thresh"] = Nrm(-4.0, 0.1), Nrm(*thresh)
        elif neuron == "plif":
            p["leak_v"], p["leak_pt"], p["add_pt"], p["thresh"] = Nrm(-4.0, 0.1), Nrm(-4.0, 0.1), Nrm(-2.0, 0.1), Nrm(*thresh)
        elif neuron == "alif":
            p["leak_v"], p["leak_t"], p["t0"], p["t1"] = Nrm(-4.0, 0.1), Nrm(-4.0, 0.1), Nrm(0.01, 0.0), Nrm(1.8, 0.0)
        elif neuron == "xlif":
            p["leak_v"], p["leak_pt"], p["t0"], p["t1"] = Nrm(-4.0, 0.1), Nrm(-4.0, 0.1), Nrm(0.01, 0.0), Nrm(1.8, 0.0)
        params[name] = p
    params["pred"] = {"weight": U((2, channels, 1, 1), 0.01), "bias": torch.zeros(2)}
    return params


# ---------------------------------------------------------------------------------------------------------------------
# ANN cells of the FireNet family (models/submodules.py:64-83, 377-418) and the ANN chain (models/model.py:254-265)
# ---------------------------------------------------------------------------------------------------------------------
def conv_layer_step(x, weight, bias, activation="relu", residual=0):
    """ConvLayer_.forward (submodules.py:69-83): conv + bias, += residual, activation."""
    out = F.conv2d(x, weight, bias, 1, weight.shape[-1] // 2)
    out = out + residual
    if activation is not None:
        out = getattr(torch, activation)(out)
    return out


def conv_gru_step(x, h, p):
    """ConvGRU.forward (submodules.py:400-418).  p: update_w/b, reset_w/b, out_w/b; h may be None (zeros)."""
    if h is None:
        h = torch.zeros(x.shape[0], p["out_w"].shape[0], *x.shape[2:], dtype=x.dtype)
    stacked = torch.cat([x, h], dim=1)
    update = torch.sigmoid(F.conv2d(stacked, p["update_w"], p["update_b"], 1, 1))
    reset = torch.sigmoid(F.conv2d(stacked, p["reset_w"], p["reset_b"], 1, 1))
    out = torch.tanh(F.conv2d(torch.cat([x, h * reset], dim=1), p["out_w"], p["out_b"], 1, 1))
    return h * (1 - update) + out * update


def firenet_ann_step(params, states, x, ff_act="relu", rec_act=None, recurrent=True):
    """
    ANN FireNet (recurrent=True: ConvGRU at G1/G2) or FireFlowNet (all ConvLayer_) forward pass.
    params[layer] = {"w","b"} for conv cells, the 6 gate tensors for GRU cells; params["pred"] = {"weight","bias"}.
    states: list of 7 (GRU hidden state or None).  Returns flow, new states, per-layer outputs.
    """
    new_states, acts = [], []
    h = x
    for i, name in enumerate(FIRENET_LAYERS):
        if recurrent and name in FIRENET_RECURRENT:
            h = conv_gru_step(h, states[i], params[name])
            new_states.append(h)
        else:  # FireFlowNet builds G1/G2 as ConvLayer_ with the *recurrent* activation of the config (model.py:175-187)
            h = conv_layer_step(h, params[name]["w"], params[name]["b"], rec_act if name in FIRENET_RECURRENT else ff_act)
            new_states.append(None)
        acts.append(h)
    flow = pred_head(h, params["pred"]["weight"], params["pred"]["bias"])
    return flow, new_states, acts
