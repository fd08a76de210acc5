"""
GPU parity of the fused conv + neuron kernels (T1 forward, T2 backward), teacher-forced:
identical (x, prev_state, params) go to the CUDA path (through the C ABI) and to the reference numbers
(golden fixtures written from the reference itself) or to the CPU oracle.
Tolerances: v abs <= 2e-5 (fp32 summation order), spikes exact outside |v-thresh| < 1e-5, gradients rel 1e-3 (north-star).
"""
import glob
import os

import pytest
import torch

from oracle import spiking as osp
from tests.conftest import GOLDEN, load_golden
from tests.util import assert_rel, spike_band_compare

pytestmark = pytest.mark.gpu
CELLS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "cell_*.npz")))
DEV = "cuda"


def _thresh_map(neuron, p, new_state):
    if neuron in ("lif", "plif"):
        return p["thresh"].clamp_min(0.01)
    return p["t0"].clamp_min(0.01) + p["t1"].clamp_min(0) * new_state[2]


def run_cuda_cell(neuron, x, state, p, hard, surrogate="arctanspike", width=10.0, g_out=None, g_state=None, residual=None, detach=True):
    from event_flow_b200 import ops

    xd = x.detach().to(DEV).requires_grad_(True)
    sd = None if state is None else state.detach().to(DEV).requires_grad_(True)
    pd = {k: v.detach().to(DEV).requires_grad_(True) for k, v in p.items()}
    rd = None if residual is None else residual.to(DEV)
    chan = {k: v for k, v in pd.items() if k not in ("ff", "rec")}
    out, ns = ops.cell_step(neuron, xd, sd, pd["ff"], pd.get("rec"), chan, hard_reset=hard, surrogate=surrogate, width=width, residual=rd,
                            detach=detach)
    grads = None
    if g_out is not None:
        ((out * g_out.to(DEV)).sum() + (ns * g_state.to(DEV)).sum()).backward()
        grads = {"x": xd.grad, "state": None if sd is None else sd.grad}
        grads.update({k: v.grad for k, v in pd.items()})
    return out.detach().cpu(), ns.detach().cpu(), grads


@pytest.mark.parametrize("name", CELLS)
def test_cell_step_matches_reference_golden(name):
    g = load_golden(name)
    _, neuron, _, reset, _ = name.split("_")
    p = {k[2:]: v for k, v in g.items() if k.startswith("p_")}
    out, ns, grads = run_cuda_cell(neuron, g["x"], g["state"], p, reset == "hard", width=float(g["width"]), g_out=g["g_out"],
                                   g_state=g["g_state"])
    thr = _thresh_map(neuron, p, g["new_state"])
    spike_band_compare(ns[0], ns[1], g["new_state"][0], g["new_state"][1], thr)
    if ns.shape[0] == 3:
        torch.testing.assert_close(ns[2], g["new_state"][2], rtol=1e-5, atol=1e-6)
    assert torch.equal(out, ns[1])
    # The backward of ONE cell step reads its inputs, the new membrane potential and the parameters -- never the emitted spikes
    # (the surrogate is evaluated at v - thresh, a continuous function) -- so the gradients are comparable UNCONDITIONALLY: a
    # borderline output spike that flipped (|v - thresh| < 1e-5) changes them by O(1e-5), far inside the tolerance.
    assert_rel(grads["x"], g["grad_x"], 1e-3, "g_x")
    assert_rel(grads["state"], g["grad_state"], 1e-3, "g_state")
    for k in p:
        if "grad_" + k in g:
            assert_rel(grads[k], g["grad_" + k], 1e-3, "g_" + k)


CASES = [(n, rec, hard) for n in osp.NEURONS for rec in (False, True) for hard in (True, False)]


@pytest.mark.parametrize("neuron,rec,hard", CASES)
@pytest.mark.parametrize("state_given", [True, False])
def test_cell_step_matches_oracle_ragged_shape(neuron, rec, hard, state_given):
    # 8 cells x {hard,soft} x {prev_state None, given}; H, W not multiples of the 16x16 tile; B=3
    B, Cin, C, H, W = 3, 32, 32, 37, 53
    g = torch.Generator().manual_seed(hash((neuron, rec, hard)) % 1000)
    params = osp.init_firenet_params(neuron, Cin, C, seed=3, weight_gain=2.0)["G1" if rec else "R1a"]
    if neuron in ("alif", "xlif"):
        params["t0"] = params["t0"] + 0.05
    x = (torch.rand((B, Cin, H, W), generator=g) < 0.3).float()
    n_state = 2 if neuron == "lif" else 3
    st = None
    if state_given:
        st = torch.rand((n_state, B, C, H, W), generator=g)
        st[1] = (st[1] < 0.3).float()
    g_out, g_state = torch.rand((B, C, H, W), generator=g), torch.rand((n_state, B, C, H, W), generator=g)
    g_state[1] = 0
    out, ns, grads = run_cuda_cell(neuron, x, st, params, hard, g_out=g_out, g_state=g_state)
    xo = x.clone().requires_grad_(True)
    so = None if st is None else st.clone().requires_grad_(True)
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out_o, ns_o = osp.cell_step(neuron, xo, so, po, hard_reset=hard)
    spike_band_compare(ns[0], ns[1], ns_o[0].detach(), ns_o[1].detach(), _thresh_map(neuron, params, ns_o.detach()))
    ((out_o * g_out).sum() + (ns_o * g_state).sum()).backward()  # unconditional: see test_cell_step_matches_reference_golden
    assert_rel(grads["x"], xo.grad, 1e-3, "g_x")
    if so is not None:
        assert_rel(grads["state"], so.grad, 1e-3, "g_state")
    for k, v in po.items():
        if v.grad is not None and v.grad.abs().max() > 0:
            assert_rel(grads[k], v.grad, 1e-3, "g_" + k)


@pytest.mark.parametrize("surrogate,width", [("superspike", 10.0), ("trianglespike", 1.0), ("mgspike", 0.5)])
def test_other_surrogates_backward(surrogate, width):
    B, Cin, C, H, W = 2, 32, 32, 20, 24
    g = torch.Generator().manual_seed(5)
    params = osp.init_firenet_params("lif", Cin, C, seed=4, weight_gain=2.0)["G1"]
    x = (torch.rand((B, Cin, H, W), generator=g) < 0.3).float()
    st = torch.rand((2, B, C, H, W), generator=g)
    st[1] = (st[1] < 0.3).float()
    g_out, g_state = torch.rand((B, C, H, W), generator=g), torch.zeros((2, B, C, H, W))
    out, ns, grads = run_cuda_cell("lif", x, st, params, True, surrogate, width, g_out, g_state)
    xo, so = x.clone().requires_grad_(True), st.clone().requires_grad_(True)
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out_o, ns_o = osp.cell_step("lif", xo, so, po, hard_reset=True, surrogate=surrogate, width=width)
    assert torch.equal(ns[1], ns_o[1].detach())
    (out_o * g_out).sum().backward()
    assert_rel(grads["x"], xo.grad, 1e-3, "g_x")
    assert_rel(grads["ff"], po["ff"].grad, 1e-3, "g_ff")
    assert_rel(grads["thresh"], po["thresh"].grad, 1e-3, "g_thresh")


def test_head_layer_fractional_inputs_stride2_and_residual():
    # head: Cin=5 voxel input with fractional values; stride 2 (U-Net encoders); residual added to the spikes
    B, Cin, C, H, W = 2, 5, 32, 33, 46
    g = torch.Generator().manual_seed(9)
    params = osp.init_firenet_params("lif", Cin, C, seed=1, weight_gain=3.0)["head"]
    x = torch.randn((B, Cin, H, W), generator=g) * (torch.rand((B, Cin, H, W), generator=g) < 0.4)
    for stride in (1, 2):
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        st = torch.rand((2, B, C, Ho, Wo), generator=g)
        st[1] = (st[1] < 0.3).float()
        res = (torch.rand((B, C, Ho, Wo), generator=g) < 0.5).float()
        from event_flow_b200 import ops

        pd = {k: v.to(DEV) for k, v in params.items()}
        out, ns = ops.cell_step("lif", x.to(DEV), st.to(DEV), pd["ff"], None, {"leak": pd["leak"], "thresh": pd["thresh"]},
                                hard_reset=True, stride=stride, residual=res.to(DEV))
        out_o, ns_o = osp.cell_step("lif", x, st, params, hard_reset=True, stride=stride, residual=res)
        spike_band_compare(ns[0].cpu(), ns[1].cpu(), ns_o[0], ns_o[1], params["thresh"].clamp_min(0.01))
        assert torch.equal(out.cpu() - res, ns[1].cpu())


def test_cl_layout_roundtrip_and_cl_inputs():
    from event_flow_b200 import ops

    g = torch.Generator().manual_seed(2)
    x = torch.randint(0, 3, (2, 32, 19, 23), generator=g).float().to(DEV)
    packed = ops.pack_cl(x)
    assert packed.shape == (2, 19, 23, 32) and packed.dtype == torch.bfloat16
    assert torch.equal(packed.float().permute(0, 3, 1, 2), x)
    assert torch.equal(ops.unpack_cl(packed), x)


@pytest.mark.parametrize("neuron,rec,hard", CASES)
def test_differentiable_reset_matches_oracle_autograd(neuron, rec, hard):
    """detach=False (spiking_submodules.py:110-112 and twins): the previous spikes also receive gradient through the reset term."""
    B, Cin, C, H, W = 2, 32, 32, 21, 28
    g = torch.Generator().manual_seed(17)
    params = osp.init_firenet_params(neuron, Cin, C, seed=6, weight_gain=2.0)["G1" if rec else "R1a"]
    if neuron in ("alif", "xlif"):
        params["t0"] = params["t0"] + 0.05
    x = (torch.rand((B, Cin, H, W), generator=g) < 0.3).float()
    n_state = 2 if neuron == "lif" else 3
    st = torch.rand((n_state, B, C, H, W), generator=g)
    st[1] = (st[1] < 0.3).float()
    g_out, g_state = torch.rand((B, C, H, W), generator=g), torch.rand((n_state, B, C, H, W), generator=g)
    g_state[1] = 0
    out, ns, grads = run_cuda_cell(neuron, x, st, params, hard, g_out=g_out, g_state=g_state, detach=False)
    _, _, grads_det = run_cuda_cell(neuron, x, st, params, hard, g_out=g_out, g_state=g_state, detach=True)
    xo, so = x.clone().requires_grad_(True), st.clone().requires_grad_(True)
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out_o, ns_o = osp.cell_step(neuron, xo, so, po, hard_reset=hard, detach=False)
    ((out_o * g_out).sum() + (ns_o * g_state).sum()).backward()
    assert_rel(grads["state"], so.grad, 1e-3, "g_state")
    assert_rel(grads["x"], xo.grad, 1e-3, "g_x")
    assert (grads["state"][1] - grads_det["state"][1]).abs().max() > 1e-3 * so.grad[1].abs().max()  # the reset path carries gradient
    for k, v in po.items():
        if v.grad is not None and v.grad.abs().max() > 0:
            assert_rel(grads[k], v.grad, 1e-3, "g_" + k)


@pytest.mark.parametrize("rec", [False, True])
@pytest.mark.parametrize("norm", ["weight", "group"])
def test_lif_cell_normalisation_options_match_torch_composition(rec, norm):
    """
    ConvLIF / ConvLIFRecurrent with norm="weight" | "group" (spiking_submodules.py:86-99, 501-529): the CUDA cell behind torch's own
    weight-norm parametrisation / GroupNorm modules against the oracle cell step behind the same modules, values and gradients.
    """
    import event_flow_b200.models.spiking_submodules as S

    torch.manual_seed(2)
    cls = S.ConvLIFRecurrent if rec else S.ConvLIF
    cell = cls(8, 16, 3, norm=norm).to(DEV)
    with torch.no_grad():
        for n, p in cell.named_parameters():
            if n.endswith("weight_g") or n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.mul_(3.0)
    g = torch.Generator().manual_seed(4)
    x = torch.rand((2, 8, 20, 24), generator=g)
    st = torch.rand((2, 2, 16, 20, 24), generator=g)
    st[1] = (st[1] < 0.3).float()
    g_out = torch.rand((2, 16, 20, 24), generator=g)
    xd, sd = x.to(DEV).requires_grad_(True), st.to(DEV).requires_grad_(True)
    out, ns = cell(xd, sd)
    (out * g_out.to(DEV)).sum().backward()
    # the same composition on the CPU: torch modules around the oracle's cell step
    ref = cls(8, 16, 3, norm=norm)  # (a weight-normalised module holds a computed `weight` tensor: no deepcopy)
    ref.load_state_dict({k: v.cpu() for k, v in cell.state_dict().items()})
    xo, so = x.clone().requires_grad_(True), st.clone().requires_grad_(True)
    xin, sin = xo, so
    n_in = ref._modules.get("norm_ff" if rec else "norm")
    if n_in is not None:
        xin = n_in(xo)
    if ref._modules.get("norm_rec") is not None:
        sin = torch.stack([so[0], ref.norm_rec(so[1])])
    po = {"ff": ref._kernel_of(ref.ff), "leak": ref.leak, "thresh": ref.thresh}
    if rec:
        po["rec"] = ref._kernel_of(ref.rec)
    out_o, ns_o = osp.cell_step("lif", xin, sin, po, hard_reset=True)
    (out_o * g_out).sum().backward()
    spike_band_compare(ns[0].detach().cpu(), ns[1].detach().cpu(), ns_o[0].detach(), ns_o[1].detach(), ref.thresh.detach().clamp_min(0.01))
    assert ns[1].mean() > 0.01
    assert_rel(xd.grad, xo.grad, 1e-3, "g_x")
    assert_rel(sd.grad, so.grad, 1e-3, "g_state")
    for (n, pa), pb in zip(cell.named_parameters(), ref.parameters()):
        if pb.grad is not None and pb.grad.abs().max() > 0:
            assert_rel(pa.grad, pb.grad, 1e-3, "g_" + n)
