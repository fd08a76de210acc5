"""CPU: the C-ABI library loads, exports every symbol include/eventflow.h declares, and the ctypes structs match the C layout."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "eventflow.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ef_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from event_flow_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _lib


def test_library_exports_every_declared_symbol(lib):
    h = ctypes.CDLL(lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(h, n), f"{n} declared in eventflow.h but not exported"
    assert set(names) == set(lib.EXPORTS), "ctypes binding and header disagree"


def test_version_and_error_string(lib):
    assert lib.lib().ef_version() == 200
    assert isinstance(lib.lib().ef_last_error(), bytes)


def test_argument_errors_are_reported_without_a_gpu(lib):
    h = lib.lib()
    assert h.ef_lif_conv_fwd(None, None) == -3 and b"NULL" in h.ef_last_error()
    p = lib.LifConvParams()
    assert h.ef_lif_conv_fwd(ctypes.byref(p), None) == -1  # non-positive dims
    p.B = p.Cin = p.C = p.H = p.W = 8
    p.ksize, p.stride = 5, 1
    assert h.ef_lif_conv_fwd(ctypes.byref(p), None) == -2 and b"kernel_size" in h.ef_last_error()
    q = lib.IweLossParams()
    assert h.ef_iwe_loss_fwd(ctypes.byref(q), None) == -1
    assert h.ef_iwe_image(None, None) == -3
    assert h.ef_encode_events(None, None) == -3
    assert h.ef_pack_cl(None, None, 1, 8, 4, 4, None) == -3
    # round-2 entry points: the general tensor-core gradients, the fused data-parallel step, the IPC helpers
    assert h.ef_conv32_bwd_tc(None, None) == -3
    assert h.ef_split2_pack_cl(None, None, None, 1, 32, 8, 8, 8, 8, None) == -3
    assert h.ef_wgrad_tcg_partial_elems(1, 16, 16, 33, 32) == 0 and h.ef_wgrad_tcg_partial_elems(4, 16, 16, 512, 512) == 256 * 2 * 9 * 32 * 32
    assert h.ef_wgrad_tcg(None, None, None, 1, 16, 16, 32, 32, None, None, 32, 0, None) == -3
    assert h.ef_dp_step(None, None) == -3
    d = lib.DpStepParams()
    d.world, d.rank, d.n = 9, 0, 10
    assert h.ef_dp_step(ctypes.byref(d), None) == -1 and b"world" in h.ef_last_error()
    assert h.ef_ipc_alloc(0, None, None) == -3 and h.ef_ipc_open(None, None) == -3


def test_ctypes_structs_match_c_layout(lib, tmp_path):
    structs = {
        "ef_lif_conv_params": lib.LifConvParams,
        "ef_lif_conv_bwd_params": lib.LifConvBwdParams,
        "ef_pred_params": lib.PredParams,
        "ef_iwe_loss_params": lib.IweLossParams,
        "ef_iwe_loss_pass_params": lib.IweLossPassParams,
        "ef_iwe_interp_params": lib.IweInterpParams,
        "ef_lif_bwd_tc_params": lib.LifBwdTcParams,
        "ef_conv32_bwd_tc_params": lib.Conv32BwdTcParams,
        "ef_iwe_metrics_params": lib.IweMetricsParams,
        "ef_aee_params": lib.AeeParams,
        "ef_conv_ann_params": lib.ConvAnnParams,
        "ef_ann_gate_bwd_params": lib.AnnGateBwdParams,
        "ef_dp_step_params": lib.DpStepParams,
        "ef_iwe_image_params": lib.IweImageParams,
        "ef_encode_params": lib.EncodeParams,
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)])  # the header is plain C
    out = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_product_path_refuses_cpu_tensors(lib):
    import torch

    from event_flow_b200 import ops

    with pytest.raises(lib.EventFlowError):
        ops.pack_cl(torch.zeros(1, 8, 4, 4))
    from event_flow_b200.models.model import LIFFireNet

    cfg = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=32, kernel_size=3,
               activations=["arctanspike", "arctanspike"], mask_output=True, spiking_neuron={})
    m = LIFFireNet(cfg)
    with pytest.raises(lib.EventFlowError):
        m(torch.zeros(1, 2, 16, 16), torch.zeros(1, 2, 16, 16))


def test_dropin_names_and_state_dict_keys(lib):
    import event_flow_b200

    event_flow_b200.install_dropin()
    from models.model import LIFFireNet  # noqa: resolves to this package

    assert LIFFireNet.__module__ == "event_flow_b200.models.model"
    cfg = dict(name="x", encoding="cnt", round_encoding=False, norm_input=False, num_bins=2, base_num_channels=32, kernel_size=3,
               activations=["arctanspike", "arctanspike"], mask_output=True,
               spiking_neuron=dict(leak=[-4.0, 0.1], thresh=[0.8, 0.1], learn_leak=True, learn_thresh=True, hard_reset=True))
    m = LIFFireNet(cfg)
    keys = set(m.state_dict().keys())
    # {head,R1a,R1b,R2a,R2b}.{leak,thresh,ff.weight,act_width}, {G1,G2}.{..,rec.weight}, pred.conv2d.{weight,bias}
    want = {f"{l}.{k}" for l in ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b") for k in ("leak", "thresh", "ff.weight", "act_width")}
    want |= {"G1.rec.weight", "G2.rec.weight", "pred.conv2d.weight", "pred.conv2d.bias"}
    assert keys == want
    assert m.G1.rec.weight.shape == (32, 32, 3, 3) and m.head.leak.shape == (32, 1, 1) and m.head.act_width.shape == ()
    assert "Trainable parameters: 74818" in str(m)
    for s in ("models", "loss", "utils", "dataloader"):
        sys.modules.pop(s, None)
    for s in [k for k in sys.modules if k.split(".")[0] in ("models", "loss", "utils", "dataloader")]:
        sys.modules.pop(s, None)


def test_dropin_resolves_the_reference_import_list_and_falls_back_to_reference_files(tmp_path):
    """
    train_flow.py:8-34 imports 19 model names and modules this package does not replace (utils.utils, dataloader.h5, ...):
    after install_dropin(reference_root=...) the former resolve here, the latter to the reference's own files.
    """
    import subprocess
    import sys

    ref = tmp_path / "ref"
    (ref / "utils").mkdir(parents=True)
    (ref / "dataloader").mkdir()
    (ref / "utils" / "utils.py").write_text("def load_model(*a):\n    return 'reference utils.utils'\n")
    (ref / "dataloader" / "h5.py").write_text("class H5Loader:\n    pass\n")
    (ref / "train_flow.py").write_text("")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import event_flow_b200\n"
        "event_flow_b200.install_dropin(reference_root=%r)\n"
        "from dataloader.h5 import H5Loader\n"
        "from loss.flow import EventWarping\n"
        "from models.model import (FireNet, RNNFireNet, LeakyFireNet, FireFlowNet, LeakyFireFlowNet, E2VID, EVFlowNet, RecEVFlowNet,\n"
        "    LeakyRecEVFlowNet, RNNRecEVFlowNet, LIFFireNet, PLIFFireNet, ALIFFireNet, XLIFFireNet, LIFFireFlowNet, SpikingRecEVFlowNet,\n"
        "    PLIFRecEVFlowNet, ALIFRecEVFlowNet, XLIFRecEVFlowNet)\n"
        "from utils.utils import load_model\n"
        "from utils.iwe import compute_pol_iwe\n"
        "import models.unet, models.spiking_submodules\n"
        "assert load_model() == 'reference utils.utils' and EventWarping.__module__.startswith('event_flow_b200')\n"
        "assert models.unet.SpikingMultiResUNetRecurrent.__module__ == 'event_flow_b200.models.unet'\n"
        "assert E2VID.__module__ == 'event_flow_b200.models.model'\n"
        "print('ok')\n"
    ) % (ROOT, str(ref))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="needs the reference tree (build container only)")
def test_whole_module_checkpoint_written_by_the_reference_unpickles_into_this_package(tmp_path):
    """
    utils/utils.py:19-20,36-37 saves and loads WHOLE modules.  A LIFFireNet pickled by the unmodified reference must unpickle,
    under the drop-in aliases, into this package's classes with identical parameters and a usable cell / state API.
    """
    ckpt = str(tmp_path / "ref_model.pt")
    cfg = ("dict(name='LIFFireNet', encoding='voxel', round_encoding=False, norm_input=False, num_bins=5, base_num_channels=32, kernel_size=3, "
           "activations=['arctanspike', 'arctanspike'], mask_output=True, spiking_neuron=dict(leak=[-4.0, 0.1], thresh=[0.8, 0.1], "
           "learn_leak=True, learn_thresh=True, hard_reset=True))")
    save = ("import sys, torch; sys.path.insert(0, '/root/reference')\n"
            "import models.model as M\n"
            "M.LIFFireNet.kwargs = [{}] * 7\n"
            "torch.manual_seed(3); m = M.LIFFireNet(%s)\n"
            "torch.save(m, %r); torch.save(m.state_dict(), %r)\n") % (cfg, ckpt, ckpt + ".sd")
    load = ("import sys, torch; sys.path.insert(0, %r)\n"
            "import event_flow_b200; event_flow_b200.install_dropin()\n"
            "m = torch.load(%r, weights_only=False)\n"
            "sd = torch.load(%r)\n"
            "assert type(m).__module__ == 'event_flow_b200.models.model' and type(m).__name__ == 'LIFFireNet', type(m)\n"
            "assert type(m.G1).__module__ == 'event_flow_b200.models.spiking_submodules'\n"
            "mine = m.state_dict()\n"
            "assert list(mine) == list(sd) and all(torch.equal(mine[k], sd[k]) for k in sd)\n"
            "assert m.G1.stride == 1 and m.head.activation == 'arctanspike' and m.G1.recurrent and m.head.neuron == 'lif'\n"
            "assert m._fast is None and m.states == [None] * 7\n"
            "m.reset_states(); print('ok')\n") % (ROOT, ckpt, ckpt + ".sd")
    for code in (save, load):
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().endswith("ok")
