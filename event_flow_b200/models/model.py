"""
Model zoo behind the class names, constructor contract (`config["model"]` dict), state API and output dict of the reference's
models/model.py, so that `eval(config["model"]["name"])(config["model"])` in train_flow.py:78-81 / eval_flow.py:93-101 keeps
working.  Two families:

* FireNet chain (models/model.py:148-286 and its subclasses :398-409, :614-704): seven conv cells + a 1x1 tanh prediction.  The
  LIF variants run on the fused tcgen05 path (event_flow_b200/fast.py), everything else cell by cell through ops.*.
* U-Net models (models/model.py:29-145, :289-395, :410-611): input selection / normalisation / padding, one of the U-Nets of
  models/unet.py, flow post-processing (nearest upsampling of the coarse scales, crop).

The nineteen public classes are declared at the bottom from two small tables (which cell types a FireNet variant uses; which
U-Net and recurrent block a U-Net model uses) instead of one hand-written subclass each.
"""
import torch

from .. import fast, graphed, ops
from . import spiking_submodules as snn
from . import submodules as ann
from . import unet as nets
from .base import BaseModel
from .model_util import CropParameters, copy_states

_CHAIN = ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b")   # execution (and construction) order of the FireNet cells
_GATED = ("G1", "G2")                                         # built from `rec_neuron`; their output feeds the optional residuals
_RESIDUAL_INTO = ("R1b", "R2b")                               # receive the last gated output as residual when `residual` is set
_ACTIVITY_KEYS = ["0:input", "1:head", "2:G1", "3:R1a", "4:R1b", "5:G2", "6:R2a", "7:R2b", "8:pred"]


def _network_input(model, event_voxel, event_cnt):
    """Encoding select + optional in-place normalisation of the non-zero entries (models/model.py:236-252, identical in every model)."""
    if model.encoding == "voxel":
        x = event_voxel
    elif model.encoding == "cnt" and model.num_bins == 2:
        x = event_cnt
    else:
        print("Model error: Incorrect input encoding.")
        raise AttributeError
    if model.norm_input:  # on the caller's tensor, like the reference
        nz = x != 0
        mean, stddev = x[nz].mean(), x[nz].std()
        x[nz] = (x[nz] - mean) / stddev
    return x


def _detach_all(states):
    """detach() of every state entry; LSTM states are (hidden, cell) tuples (models/model.py:211-221)."""
    return [tuple(h.detach() for h in s) if type(s) is tuple else s.detach() for s in states]


def _common_options(model, cfg):
    model.num_bins = cfg["num_bins"]
    model.encoding = cfg["encoding"]
    model.norm_input = cfg["norm_input"] if "norm_input" in cfg.keys() else False
    model.mask = cfg["mask_output"]


# =====================================================================================================================
# FireNet family
# =====================================================================================================================
class FireNet(BaseModel):
    """
    head - G1 - R1a - R1b - G2 - R2a - R2b + 1x1 tanh prediction (models/model.py:148-286).  This base class is the ANN FireNet
    (ConvLayer_ cells, ConvGRU at G1 / G2); the variants below only swap the three cell types.
    """

    head_neuron = ann.ConvLayer_
    ff_neuron = ann.ConvLayer_
    rec_neuron = ann.ConvGRU
    residual = False
    num_recurrent_units = len(_CHAIN)
    w_scale_pred = None

    def __init__(self, unet_kwargs):
        super().__init__()
        _common_options(self, unet_kwargs)
        width, ksize = unet_kwargs["base_num_channels"], unet_kwargs["kernel_size"]
        ff_act, rec_act = unet_kwargs["activations"]
        # the reference keeps ONE class-level list of kwargs dicts shared by all instances (model.py:159,171-173); per instance here
        extra = dict(unet_kwargs["spiking_neuron"]) if type(unet_kwargs.get("spiking_neuron")) is dict else {}
        for name in _CHAIN:  # creation order = the reference's, so the same torch seed gives the same initial values
            make = self.head_neuron if name == "head" else (self.rec_neuron if name in _GATED else self.ff_neuron)
            cell = make(self.num_bins if name == "head" else width, width, ksize, activation=rec_act if name in _GATED else ff_act, **extra)
            setattr(self, name, cell)
        self._mark_inputs()
        self.pred = ann.ConvLayer(width, out_channels=2, kernel_size=1, activation="tanh", w_scale=self.w_scale_pred)
        self.reset_states()

    def _mark_inputs(self):
        """
        Tell the spiking cells what their input is, so that their convolution may run on the tensor cores with exact products: the head
        sees the (fractional) event encoding -> exact bf16 split; every later cell sees the spikes of a spiking cell (plus, in the
        residual variants, another spike tensor: small integers) -> exact in bf16 as they are.
        """
        prev_spiking = False
        for name in _CHAIN:
            cell = getattr(self, name)
            spiking = isinstance(cell, snn._SpikingConvCell)
            if spiking:
                cell.__dict__["_x_kind"] = "split" if name == "head" else ("spikes" if prev_spiking else None)
            prev_spiking = spiking

    def __getattr__(self, name):
        if name == "_fast":  # a FireNet unpickled from a checkpoint the reference wrote has no fast-path state yet
            return None
        return super().__getattr__(name)

    # Run-time caches of the fast path (CUDA graphs, ctypes argument structs, activation slabs, weight images): none of them
    # can or should travel with a checkpoint.  The reference checkpoints by pickling the whole module (utils/utils.py:36,
    # mlflow.pytorch.log_model) and callers deepcopy models; both go through __getstate__.
    _RUNTIME_KEYS = ("_fast_params", "_fast_cells", "_fast_eligible", "_w_split_cache", "_arena", "_capture", "_last_spikes", "_w_epoch", "_grad_sink") + BaseModel._GRAPH_KEYS

    def __getstate__(self):
        state = self.__dict__.copy()
        if state.get("_fast") is not None:  # internal state -> the reference's stacked fp32 tensors (detached copies)
            state["_states"] = [None if s is None else s.detach() for s in fast.states_of(self)]
        state["_fast"] = None
        for k in self._RUNTIME_KEYS:
            state.pop(k, None)
        return state

    # ---- state API (models/model.py:203-227) ----
    @property
    def states(self):
        if self._fast is not None:  # internal (channels-last spike) state -> the reference's stacked fp32 format; fresh tensors = clones
            return fast.states_of(self)
        return copy_states(self._states)

    @states.setter
    def states(self, states):
        self._states = states
        self._fast = None  # converted to the internal format again by the next forward pass

    def detach_states(self):
        if self._fast is not None:  # cut the BPTT chain without copying any state
            fast.detach(self)
            return
        self.states = _detach_all(self.states)

    def reset_states(self):
        self._states = [None] * self.num_recurrent_units
        if getattr(self, "_fast", None) is not None:
            fast.detach(self)  # a new sequence is also a window boundary for the activation arena
        self._fast = None

    def init_cropping(self, width, height):
        pass

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float(): parameter storage moves, so the pointer-keyed caches of the fast path are dropped
        out = super()._apply(fn, *args, **kwargs)
        for k in ("_fast_params", "_fast_cells", "_fast_eligible", "_w_split_cache", "_arena") + BaseModel._GRAPH_KEYS:
            self.__dict__.pop(k, None)
        if getattr(self, "_fast", None) is not None:
            self._fast.param_sig = None
        return out

    def forward(self, event_voxel, event_cnt, log=False):
        """
        :param event_voxel: N x num_bins x H x W
        :param event_cnt: N x 2 x H x W per-polarity event counts
        :return {"flow": [N x 2 x H x W], "activity": dict | None}
        """
        x = _network_input(self, event_voxel, event_cnt)
        if fast.eligible(self, x):  # LIF, 32 channels: tcgen05 kernels on the internal spike format, one autograd node per step
            return fast.forward(self, x, log)
        if graphed.usable(self, x, log):  # evaluation loop (no_grad): the cell-by-cell step below, replayed as one CUDA graph
            return graphed.step(self, x, self, "_states", self._cell_step)
        graphed.leave(self, self, "_states")
        return self._cell_step(x, log)

    def _cell_step(self, x, log=False):
        """The chain cell by cell (models/model.py:255-286); reads and replaces the entries of self._states."""
        seen, h, gated = [x], x, None
        for i, name in enumerate(_CHAIN):
            cell = getattr(self, name)
            if name in _RESIDUAL_INTO:
                h, self._states[i] = cell(h, self._states[i], residual=gated if self.residual else 0)
            else:
                h, self._states[i] = cell(h, self._states[i])
            if name in _GATED:
                gated = h
            seen.append(h)
        flow = self.pred(h)
        seen.append(flow)

        activity = None
        if log:
            activity = {key: t.detach().ne(0).float().mean().item() for key, t in zip(_ACTIVITY_KEYS, seen)}
        return {"flow": [flow], "activity": activity}


def _forward_window(self, event_voxels, event_cnts, log=False):
    """
    The forward passes of a whole loss window at once: `event_voxels` [T x N x num_bins x H x W], `event_cnts` [T x N x 2 x H x W].
    Returns the list of the T per-step output dicts -- numerically the steps `model(event_voxels[t], event_cnts[t])`, t = 0..T-1, would
    produce.  On the LIF fast path the window runs layer-major with the time loop INSIDE the kernels of the feed-forward cells (their
    state stays on chip over the T steps) and back-propagates as one window; other models simply loop over the steps.  Must start at
    a window boundary (after reset_states() / detach_states()), like the loop of train_flow.py:97-171 does.
    """
    if not log and not self.norm_input:  # (input normalisation is per step: the statistics of one step's tensor, model.py:246-252)
        x = _network_input(self, event_voxels, event_cnts)
        if torch.is_tensor(x) and x.dim() == 5 and fast.eligible(self, x[0]) and x.shape[2] <= fast.L.EF_HEAD_MAX_CIN:
            return [{"flow": [f], "activity": None} for f in fast.forward_window(self, x)]
    vox = event_voxels if event_voxels is not None else [None] * len(event_cnts)
    cnt = event_cnts if event_cnts is not None else [None] * len(event_voxels)
    return [self(v, c, log=log) for v, c in zip(vox, cnt)]


FireNet.forward_window = _forward_window


def _firenet_variant(name, head, ff, rec, w_scale_pred, where):
    """A FireNet subclass that only swaps the cell types (what every subclass in the reference does)."""
    cls = type(name, (FireNet,), {"head_neuron": head, "ff_neuron": ff, "rec_neuron": rec, "residual": False,
                                  "w_scale_pred": w_scale_pred, "__doc__": f"{where}.", "__module__": __name__})
    cls.__qualname__ = name  # picklable under models.model.<name> (utils/utils.py:19-20 pickles whole modules)
    return cls


# name                      head cell          feed-forward cell   G1 / G2 cell             w_scale_pred  reference
FireFlowNet = _firenet_variant("FireFlowNet", ann.ConvLayer_, ann.ConvLayer_, ann.ConvLayer_, 0.01, "EV-FireFlowNet: all-feed-forward ANN FireNet (models/model.py:398-409)")
RNNFireNet = _firenet_variant("RNNFireNet", ann.ConvLayer_, ann.ConvLayer_, ann.ConvRecurrent, None, "models/model.py:614-622")
LeakyFireNet = _firenet_variant("LeakyFireNet", ann.ConvLeaky, ann.ConvLeaky, ann.ConvLeakyRecurrent, None, "models/model.py:625-633")
LeakyFireFlowNet = _firenet_variant("LeakyFireFlowNet", ann.ConvLeaky, ann.ConvLeaky, ann.ConvLeaky, None, "models/model.py:696-704")
LIFFireNet = _firenet_variant("LIFFireNet", snn.ConvLIF, snn.ConvLIF, snn.ConvLIFRecurrent, 0.01, "models/model.py:636-645")
PLIFFireNet = _firenet_variant("PLIFFireNet", snn.ConvPLIF, snn.ConvPLIF, snn.ConvPLIFRecurrent, 0.01, "models/model.py:648-657")
ALIFFireNet = _firenet_variant("ALIFFireNet", snn.ConvALIF, snn.ConvALIF, snn.ConvALIFRecurrent, 0.01, "models/model.py:660-669")
XLIFFireNet = _firenet_variant("XLIFFireNet", snn.ConvXLIF, snn.ConvXLIF, snn.ConvXLIFRecurrent, 0.01, "models/model.py:672-681")
LIFFireFlowNet = _firenet_variant("LIFFireFlowNet", snn.ConvLIF, snn.ConvLIF, snn.ConvLIF, 0.01, "models/model.py:684-693")


# =====================================================================================================================
# U-Net models
# =====================================================================================================================
class _NoStates:
    """State holder of the stateless U-Net (EVFlowNet) for graphed.step."""

    def __init__(self):
        self.states = []


class _UNetFlowModel(BaseModel):
    """
    What EVFlowNet (models/model.py:289-395), RecEVFlowNet (:410-547) and E2VID (:29-145) share: the option handling of the
    constructor (the architecture constants are written into the caller's dict and the driver-only keys popped, as the reference
    does), zero-padding of the input to a size the encoder pyramid divides, and the flow post-processing.
    """

    net_attr = None        # attribute name of the wrapped U-Net (part of the state_dict keys)
    net_type = None
    num_pyramid_levels = 4
    drop_keys = ("name", "encoding", "round_encoding", "norm_input", "mask_output")

    def architecture(self, cfg):
        """Architecture constants of this model (merged into the constructor dict)."""
        raise NotImplementedError

    def __init__(self, unet_kwargs):
        super().__init__()
        fixed = self.architecture(unet_kwargs)
        self.crop = None
        _common_options(self, unet_kwargs)
        self.num_encoders = fixed["num_encoders"]
        unet_kwargs.update(fixed)  # in place on the caller's dict, like the reference
        for key in self.drop_keys:
            unet_kwargs.pop(key, None)
        setattr(self, self.net_attr, self.network_class()(unet_kwargs))
        unet_kwargs.pop("final_activation", None)  # the reference's U-Nets pop it from the caller's dict (unet.py:233,152,326)

    def network_class(self):
        return self.net_type

    @property
    def net(self):
        return getattr(self, self.net_attr)

    def init_cropping(self, width, height, safety_margin=0):
        self.crop = CropParameters(width, height, self.num_encoders, safety_margin)

    def detach_states(self):
        pass

    def reset_states(self):
        pass

    def forward(self, event_voxel, event_cnt, log=False):
        """
        :param event_voxel: N x num_bins x H x W
        :param event_cnt: N x 2 x H x W per-polarity event counts
        :return {"flow": list of N x 2 x H x W maps at the input resolution (coarse to fine), "activity": None}
        """
        x = _network_input(self, event_voxel, event_cnt)
        if self.crop is not None:
            x = self.crop.pad(x)
        if log:
            raise NotImplementedError("Activity logging not implemented")
        where = nets.state_list_of(self.net, x)  # None: the spiking U-Net's tensor-core inference path (own state format, own launches)
        if where is not None:
            holder, attr = where if where[0] is not None else (self.__dict__.setdefault("_no_states", _NoStates()), "states")
            if graphed.usable(self, x):  # evaluation loop (no_grad): the step below replayed as one CUDA graph
                return graphed.step(self, x, holder, attr, self._net_step)
            graphed.leave(self, holder, attr)
        return self._net_step(x)

    def _net_step(self, x):
        out = self.net.forward(x)
        if isinstance(out, (list, tuple)):  # multi-resolution estimates: nearest-neighbour upsampling to the finest one
            full_h, full_w = out[-1].shape[2], out[-1].shape[3]
            flows = [ops.upsample_nearest(f, full_h // f.shape[2], full_w // f.shape[3]) for f in out]
        else:
            flows = [out]
        if self.crop is not None:
            c = self.crop
            flows = [f[:, :, c.iy0:c.iy1, c.ix0:c.ix1].contiguous() for f in flows]
        return {"flow": flows, "activity": None}


class _RecurrentUNetFlowModel(_UNetFlowModel):
    """State API of the recurrent U-Net models (models/model.py:466-487): the states live in the wrapped network."""

    @property
    def states(self):
        return copy_states(self.net.states)

    @states.setter
    def states(self, states):
        self.net.states = states

    def detach_states(self):
        self.net.states = _detach_all(self.net.states)

    def reset_states(self):
        self.net.states = [None] * self.net.num_states


def _optional(cfg, key, default):
    return cfg[key] if key in cfg.keys() else default


class EVFlowNet(_UNetFlowModel):
    """EV-FlowNet (models/model.py:289-395): the stateless ANN multi-resolution U-Net."""

    net_attr, net_type = "multires_unet", nets.MultiResUNet
    drop_keys = _UNetFlowModel.drop_keys + ("eval", "spiking_neuron")

    def architecture(self, cfg):
        return {"base_num_channels": cfg["base_num_channels"], "num_encoders": 4, "num_residual_blocks": 2, "num_output_channels": 2,
                "skip_type": "concat", "norm": None, "use_upsample_conv": True, "kernel_size": cfg["kernel_size"],
                "channel_multiplier": 2, "final_activation": "tanh"}


class RecEVFlowNet(_RecurrentUNetFlowModel):
    """
    Recurrent EV-FlowNet (models/model.py:410-547): a recurrent block after every encoder conv.  The base class uses ConvGRU;
    the subclasses below select other recurrent blocks or the spiking / leaky U-Nets.
    """

    net_attr, net_type = "multires_unetrec", nets.MultiResUNetRecurrent
    unet_type = nets.MultiResUNetRecurrent
    recurrent_block_type = "convgru"
    spiking_feedforward_block_type = None

    def network_class(self):
        return self.unet_type  # the attribute name the reference's subclasses override (models/model.py:418,556)

    def architecture(self, cfg):
        return {"base_num_channels": cfg["base_num_channels"], "num_encoders": 4, "num_residual_blocks": 2, "num_output_channels": 2,
                "skip_type": "concat", "norm": _optional(cfg, "norm", None), "use_upsample_conv": _optional(cfg, "use_upsample_conv", True),
                "kernel_size": cfg["kernel_size"], "channel_multiplier": 2, "recurrent_block_type": self.recurrent_block_type,
                "final_activation": "tanh", "spiking_feedforward_block_type": self.spiking_feedforward_block_type,
                "spiking_neuron": cfg["spiking_neuron"]}


def _rec_evflownet_variant(name, unet_type, block, spiking_block, where):
    cls = type(name, (RecEVFlowNet,), {"unet_type": unet_type, "recurrent_block_type": block, "spiking_feedforward_block_type": spiking_block,
                                       "__doc__": f"{where}.", "__module__": __name__})
    cls.__qualname__ = name
    return cls


SpikingRecEVFlowNet = _rec_evflownet_variant("SpikingRecEVFlowNet", nets.SpikingMultiResUNetRecurrent, "lif", "lif", "models/model.py:550-558")
PLIFRecEVFlowNet = _rec_evflownet_variant("PLIFRecEVFlowNet", nets.SpikingMultiResUNetRecurrent, "plif", "plif", "models/model.py:561-569")
ALIFRecEVFlowNet = _rec_evflownet_variant("ALIFRecEVFlowNet", nets.SpikingMultiResUNetRecurrent, "alif", "alif", "models/model.py:572-580")
XLIFRecEVFlowNet = _rec_evflownet_variant("XLIFRecEVFlowNet", nets.SpikingMultiResUNetRecurrent, "xlif", "xlif", "models/model.py:583-591")
RNNRecEVFlowNet = _rec_evflownet_variant("RNNRecEVFlowNet", nets.MultiResUNetRecurrent, "convrnn", None, "models/model.py:594-601")
LeakyRecEVFlowNet = _rec_evflownet_variant("LeakyRecEVFlowNet", nets.LeakyMultiResUNetRecurrent, "convleaky", None, "models/model.py:604-611")


class E2VID(_RecurrentUNetFlowModel):
    """E2VID adapted for flow (models/model.py:29-145): recurrent U-Net with ConvLSTM encoders and sum skips, one flow map."""

    net_attr, net_type = "unetrecurrent", nets.UNetRecurrent
    num_pyramid_levels = 3
    drop_keys = _UNetFlowModel.drop_keys + ("spiking_neuron",)

    def architecture(self, cfg):
        return {"base_num_channels": cfg["base_num_channels"], "num_encoders": 3, "num_residual_blocks": 2, "num_output_channels": 2,
                "skip_type": "sum", "norm": _optional(cfg, "norm", None), "use_upsample_conv": _optional(cfg, "use_upsample_conv", True),
                "kernel_size": cfg["kernel_size"], "channel_multiplier": 2, "recurrent_block_type": "convlstm", "final_activation": "tanh"}
