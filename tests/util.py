"""Shared helpers for the GPU parity tests."""
import torch


def spike_band_compare(v_mine, z_mine, v_ref, z_ref, thresh, band=1e-5, v_atol=None):
    """
    SURVEY T1: membrane potentials agree to summation-order noise; spikes agree exactly outside a band |v - thresh| < band.
    Summation-order noise of an fp32 accumulation scales with the magnitude of the sum: measured against an fp64 oracle
    (tools/tc_accuracy_probe.py, profiles/r01_tc_accuracy.txt) it is 0.3e-6*max|v| for the CPU path, 0.7e-6*max|v| for the
    CUDA-core kernel and 1.6e-6*max|v| for the tensor-core kernel (truncating fp32 accumulate), hence 3e-6*max|v|, floor 2e-5.
    Returns (max|dv|, flips outside band, flips inside band).
    """
    if v_atol is None:
        v_atol = max(2e-5, 3e-6 * v_ref.abs().max().item())
    band = max(band, v_atol)  # a neuron closer to threshold than the membrane tolerance may legitimately flip
    dv = (v_mine - v_ref).abs().max().item()
    near = (v_ref - thresh).abs() < band
    diff = z_mine != z_ref
    out_flips = (diff & ~near).sum().item()
    in_flips = (diff & near).sum().item()
    assert dv <= v_atol, f"max|dv| = {dv:.3e} > {v_atol}"
    assert out_flips == 0, f"{out_flips} spike flips outside the |v-thresh|<{band} band"
    return dv, out_flips, in_flips


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_rel(a, b, tol, what=""):
    r = rel_err(a.detach().cpu().double(), b.detach().cpu().double())
    assert r <= tol, f"{what}: max-abs rel err {r:.3e} > {tol}"
    return r


def firenet_cfg(bins, encoding, neuron="lif"):
    sn = dict(leak=[-4.0, 0.1], thresh=[0.8, 0.1], learn_leak=True, learn_thresh=True, hard_reset=True) if neuron == "lif" else {}
    return dict(name="x", encoding=encoding, round_encoding=False, norm_input=False, num_bins=bins, base_num_channels=32,
                kernel_size=3, activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron=sn)


def capture_layers(model):
    """Forward hooks recording (x_in, state_in, out, state_out) of the 7 FireNet cells as CPU tensors."""
    from oracle.spiking import FIRENET_LAYERS

    captured, handles = {}, []

    def mk(name):
        def hook(mod, inputs, output):
            st = inputs[1] if len(inputs) > 1 else None
            captured[name] = (inputs[0].detach().cpu(), None if st is None else st.detach().cpu(), output[0].detach().cpu(),
                              output[1].detach().cpu())
        return hook

    for l in FIRENET_LAYERS:
        handles.append(getattr(model, l).register_forward_hook(mk(l)))
    model._capture = captured  # the fused fast path does not call the cells' forward(); it fills the same dict itself
    return captured, handles


def oracle_params_of(model):
    from oracle.spiking import FIRENET_LAYERS

    params = {}
    for l in FIRENET_LAYERS:
        cell = getattr(model, l)
        params[l] = {"ff": cell.ff.weight.detach().cpu().clone()}
        if hasattr(cell, "rec"):
            params[l]["rec"] = cell.rec.weight.detach().cpu().clone()
        for k in ("leak", "thresh", "leak_v", "leak_pt", "leak_t", "add_pt", "t0", "t1"):
            if hasattr(cell, k):
                params[l][k] = getattr(cell, k).detach().cpu().clone()
    params["pred"] = {"weight": model.pred.conv2d.weight.detach().cpu().clone(), "bias": model.pred.conv2d.bias.detach().cpu().clone()}
    return params


def oracle_bptt_teacher_forced(neuron, params, xs, spikes, loss_of_flows, states0=None, forced_flows=None, **cell_kwargs):
    """
    BPTT of the oracle FireNet over the inputs `xs` with every layer's spikes FORCED to `spikes[t][layer]` (the spikes the path under
    test emitted, e.g. `model.states[i][1]` after step t): both implementations then differentiate the same trajectory, and a
    borderline spike that flipped in a free-running rollout cannot invalidate the gradient comparison.
    :param forced_flows: optionally also force the VALUES of the flow maps (gradient still through the oracle's prediction head).
           Needed when the loss is the event-warping loss: its gradient is discontinuous in the flow wherever a warped coordinate
           crosses an integer and blows up like 1/I^2 at pixels with a tiny accumulated weight, so that the reference's own fp32 and
           fp64 gradients disagree by tens of percent at config size (profiles/r02_loss_gradient_noise_floor.txt); flows that agree
           to 1e-7 are not enough, they have to be identical.
    :param loss_of_flows: callable(list of flows) -> scalar
    :return (loss, {layer: {param: grad}}, flows)
    """
    from oracle import spiking as osp

    leaves = {l: {k: v.detach().clone().requires_grad_(True) for k, v in lp.items()} for l, lp in params.items()}
    states = [None] * 7 if states0 is None else [None if s is None else s.detach().clone() for s in states0]
    flows = []
    for t, x in enumerate(xs):
        flow, states, _ = osp.firenet_step(neuron, leaves, states, x, forced=spikes[t], **cell_kwargs)
        if forced_flows is not None:
            flow = forced_flows[t].detach() + (flow - flow.detach())
        flows.append(flow)
    loss = loss_of_flows(flows)
    loss.backward()
    return loss.detach(), {l: {k: v.grad for k, v in lp.items()} for l, lp in leaves.items()}, [f.detach() for f in flows]


def model_grads_by_layer(model):
    """Parameter gradients of a FireNet model keyed like the oracle's parameter dict."""
    from oracle.spiking import FIRENET_LAYERS

    out = {}
    for l in FIRENET_LAYERS:
        cell = getattr(model, l)
        out[l] = {"ff": cell.ff.weight.grad}
        if hasattr(cell, "rec"):
            out[l]["rec"] = cell.rec.weight.grad
        for k in ("leak", "thresh", "leak_v", "leak_pt", "leak_t", "add_pt", "t0", "t1"):
            if hasattr(cell, k) and isinstance(getattr(cell, k), torch.nn.Parameter):
                out[l][k] = getattr(cell, k).grad
    out["pred"] = {"weight": model.pred.conv2d.weight.grad, "bias": model.pred.conv2d.bias.grad}
    return out


def compare_grads_by_layer(mine, ref, tol, min_ref=0.0):
    """Every parameter gradient: max-abs error relative to the largest entry of the reference gradient (all reported on failure)."""
    worst, report, bad = 0.0, [], []
    for l, lp in ref.items():
        for k, g in lp.items():
            if g is None or g.abs().max().item() <= min_ref or k not in mine[l]:  # (buffers, e.g. the ALIF t0 / t1, are leaves only in the oracle)
                continue
            assert mine[l][k] is not None, f"no gradient for {l}.{k}"
            r = rel_err(mine[l][k].detach().cpu().double().reshape(g.shape), g.detach().double())
            report.append(f"{l}.{k}={r:.1e}")
            if not r <= tol:
                bad.append(f"{l}.{k}")
            worst = max(worst, r)
    assert not bad, f"gradients beyond rel {tol}: {bad}; all: {' '.join(report)}"
    return worst
