"""
Model zoo with the class names, constructor contract, state API and output dict of models/model.py.
Round 1: the spiking FireNet family (FireNet base :148-286; LIFFireNet :636, PLIFFireNet :648, ALIFFireNet :660,
XLIFFireNet :672, LIFFireFlowNet :684).  Each forward pass is 7 fused conv+neuron kernels and one prediction-head kernel.
"""
from .. import fast, ops
from .base import BaseModel
from .model_util import CropParameters, copy_states
from .spiking_submodules import (
    ConvALIF,
    ConvALIFRecurrent,
    ConvLIF,
    ConvLIFRecurrent,
    ConvPLIF,
    ConvPLIFRecurrent,
    ConvXLIF,
    ConvXLIFRecurrent,
)
from .submodules import ConvGRU, ConvLayer, ConvLayer_, ConvLeaky, ConvLeakyRecurrent, ConvRecurrent
from .unet import LeakyMultiResUNetRecurrent, MultiResUNet, MultiResUNetRecurrent, SpikingMultiResUNetRecurrent, UNetRecurrent


class FireNet(BaseModel):
    """
    7-cell chain head-G1-R1a-R1b-G2-R2a-R2b + 1x1 tanh prediction (models/model.py:148-286).
    The base class is the ANN FireNet (ConvLayer_ + ConvGRU cells, forward only in this version).
    """

    head_neuron = ConvLayer_
    ff_neuron = ConvLayer_
    rec_neuron = ConvGRU
    residual = False
    num_recurrent_units = 7
    w_scale_pred = None

    def __init__(self, unet_kwargs):
        super().__init__()
        self.num_bins = unet_kwargs["num_bins"]
        base_num_channels = unet_kwargs["base_num_channels"]
        kernel_size = unet_kwargs["kernel_size"]
        self.encoding = unet_kwargs["encoding"]
        self.norm_input = False if "norm_input" not in unet_kwargs.keys() else unet_kwargs["norm_input"]
        self.mask = unet_kwargs["mask_output"]
        ff_act, rec_act = unet_kwargs["activations"]
        # the reference shares one class-level list of dicts between all models (model.py:159,171-173); per-instance here
        kwargs = dict(unet_kwargs["spiking_neuron"]) if type(unet_kwargs.get("spiking_neuron")) is dict else {}

        self.head = self.head_neuron(self.num_bins, base_num_channels, kernel_size, activation=ff_act, **kwargs)
        self.G1 = self.rec_neuron(base_num_channels, base_num_channels, kernel_size, activation=rec_act, **kwargs)
        self.R1a = self.ff_neuron(base_num_channels, base_num_channels, kernel_size, activation=ff_act, **kwargs)
        self.R1b = self.ff_neuron(base_num_channels, base_num_channels, kernel_size, activation=ff_act, **kwargs)
        self.G2 = self.rec_neuron(base_num_channels, base_num_channels, kernel_size, activation=rec_act, **kwargs)
        self.R2a = self.ff_neuron(base_num_channels, base_num_channels, kernel_size, activation=ff_act, **kwargs)
        self.R2b = self.ff_neuron(base_num_channels, base_num_channels, kernel_size, activation=ff_act, **kwargs)
        self.pred = ConvLayer(base_num_channels, out_channels=2, kernel_size=1, activation="tanh", w_scale=self.w_scale_pred)
        self.reset_states()

    @property
    def states(self):
        if self._fast is not None:  # internal (cl spike) state -> the reference's stacked fp32 format; fresh tensors = clones
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
        detached_states = []
        for state in self.states:
            if type(state) is tuple:
                detached_states.append(tuple(hidden.detach() for hidden in state))
            else:
                detached_states.append(state.detach())
        self.states = detached_states

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
        for k in ("_fast_params", "_fast_cells", "_fast_eligible", "_w_split_cache", "_arena"):
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
        if self.encoding == "voxel":
            x = event_voxel
        elif self.encoding == "cnt" and self.num_bins == 2:
            x = event_cnt
        else:
            print("Model error: Incorrect input encoding.")
            raise AttributeError

        if self.norm_input:  # model.py:247-252 (in place on the caller's tensor, like the reference)
            mean, stddev = x[x != 0].mean(), x[x != 0].std()
            x[x != 0] = (x[x != 0] - mean) / stddev

        if fast.eligible(self, x):  # LIF, 32 channels: tcgen05 kernels on the internal spike format, one autograd node per step
            return fast.forward(self, x, log)

        x1, self._states[0] = self.head(x, self._states[0])
        x2, self._states[1] = self.G1(x1, self._states[1])
        x3, self._states[2] = self.R1a(x2, self._states[2])
        x4, self._states[3] = self.R1b(x3, self._states[3], residual=x2 if self.residual else 0)
        x5, self._states[4] = self.G2(x4, self._states[4])
        x6, self._states[5] = self.R2a(x5, self._states[5])
        x7, self._states[6] = self.R2b(x6, self._states[6], residual=x5 if self.residual else 0)
        flow = self.pred(x7)

        if log:
            activity = {}
            name = ["0:input", "1:head", "2:G1", "3:R1a", "4:R1b", "5:G2", "6:R2a", "7:R2b", "8:pred"]
            for n, l in zip(name, [x, x1, x2, x3, x4, x5, x6, x7, flow]):
                activity[n] = l.detach().ne(0).float().mean().item()
        else:
            activity = None

        return {"flow": [flow], "activity": activity}


class FireFlowNet(FireNet):
    """EV-FireFlowNet: all-feed-forward ANN FireNet (models/model.py:398-409)."""

    head_neuron = ConvLayer_
    ff_neuron = ConvLayer_
    rec_neuron = ConvLayer_
    residual = False
    w_scale_pred = 0.01


class LIFFireNet(FireNet):
    """models/model.py:636-645."""

    head_neuron = ConvLIF
    ff_neuron = ConvLIF
    rec_neuron = ConvLIFRecurrent
    residual = False
    w_scale_pred = 0.01


class PLIFFireNet(FireNet):
    """models/model.py:648-657."""

    head_neuron = ConvPLIF
    ff_neuron = ConvPLIF
    rec_neuron = ConvPLIFRecurrent
    residual = False
    w_scale_pred = 0.01


class ALIFFireNet(FireNet):
    """models/model.py:660-669."""

    head_neuron = ConvALIF
    ff_neuron = ConvALIF
    rec_neuron = ConvALIFRecurrent
    residual = False
    w_scale_pred = 0.01


class XLIFFireNet(FireNet):
    """models/model.py:672-681."""

    head_neuron = ConvXLIF
    ff_neuron = ConvXLIF
    rec_neuron = ConvXLIFRecurrent
    residual = False
    w_scale_pred = 0.01


class LIFFireFlowNet(FireNet):
    """models/model.py:684-693."""

    head_neuron = ConvLIF
    ff_neuron = ConvLIF
    rec_neuron = ConvLIF
    residual = False
    w_scale_pred = 0.01


class EVFlowNet(BaseModel):
    """EV-FlowNet (models/model.py:289-395): the stateless ANN U-Net; forward only in this version."""

    def __init__(self, unet_kwargs):
        super().__init__()
        EVFlowNet_kwargs = {
            "base_num_channels": unet_kwargs["base_num_channels"],
            "num_encoders": 4,
            "num_residual_blocks": 2,
            "num_output_channels": 2,
            "skip_type": "concat",
            "norm": None,
            "use_upsample_conv": True,
            "kernel_size": unet_kwargs["kernel_size"],
            "channel_multiplier": 2,
            "final_activation": "tanh",
        }
        self.crop = None
        self.mask = unet_kwargs["mask_output"]
        self.norm_input = False if "norm_input" not in unet_kwargs.keys() else unet_kwargs["norm_input"]
        self.encoding = unet_kwargs["encoding"]
        self.num_bins = unet_kwargs["num_bins"]
        self.num_encoders = EVFlowNet_kwargs["num_encoders"]

        unet_kwargs.update(EVFlowNet_kwargs)  # in place on the caller's dict, like the reference (model.py:318-325)
        for k in ("name", "eval", "encoding", "round_encoding", "mask_output", "norm_input", "spiking_neuron"):
            unet_kwargs.pop(k, None)
        self.multires_unet = MultiResUNet(unet_kwargs)

    def detach_states(self):
        pass

    def reset_states(self):
        pass

    def init_cropping(self, width, height, safety_margin=0):
        self.crop = CropParameters(width, height, self.num_encoders, safety_margin)

    def forward(self, event_voxel, event_cnt, log=False):
        if self.encoding == "voxel":
            x = event_voxel
        elif self.encoding == "cnt" and self.num_bins == 2:
            x = event_cnt
        else:
            print("Model error: Incorrect input encoding.")
            raise AttributeError
        if self.norm_input:
            mean, stddev = x[x != 0].mean(), x[x != 0].std()
            x[x != 0] = (x[x != 0] - mean) / stddev
        if self.crop is not None:
            x = self.crop.pad(x)
        multires_flow = self.multires_unet.forward(x)
        if log:
            raise NotImplementedError("Activity logging not implemented")
        flow_list = []
        full_h, full_w = multires_flow[-1].shape[2], multires_flow[-1].shape[3]
        for flow in multires_flow:
            flow_list.append(ops.upsample_nearest(flow, full_h // flow.shape[2], full_w // flow.shape[3]))
        if self.crop is not None:
            for i, flow in enumerate(flow_list):
                flow_list[i] = flow[:, :, self.crop.iy0:self.crop.iy1, self.crop.ix0:self.crop.ix1].contiguous()
        return {"flow": flow_list, "activity": None}


class RecEVFlowNet(BaseModel):
    """
    Recurrent EV-FlowNet (models/model.py:410-547): input encoding select, optional input normalisation and padding, the
    multi-resolution recurrent U-Net, nearest-neighbour upsampling of every flow estimate to the input resolution, crop.
    The spiking subclasses (SpikingRecEVFlowNet :550, PLIF :561, ALIF :572, XLIF :583) run on the fused spiking cells, the
    ANN base class on the ConvLayer / ConvGRU kernels (ConvLSTM / ConvRecurrent encoders raise).
    """

    unet_type = MultiResUNetRecurrent
    recurrent_block_type = "convgru"
    spiking_feedforward_block_type = None

    def __init__(self, unet_kwargs):
        super().__init__()
        norm = None
        use_upsample_conv = True
        if "norm" in unet_kwargs.keys():
            norm = unet_kwargs["norm"]
        if "use_upsample_conv" in unet_kwargs.keys():
            use_upsample_conv = unet_kwargs["use_upsample_conv"]

        RecEVFlowNet_kwargs = {
            "base_num_channels": unet_kwargs["base_num_channels"],
            "num_encoders": 4,
            "num_residual_blocks": 2,
            "num_output_channels": 2,
            "skip_type": "concat",
            "norm": norm,
            "use_upsample_conv": use_upsample_conv,
            "kernel_size": unet_kwargs["kernel_size"],
            "channel_multiplier": 2,
            "recurrent_block_type": self.recurrent_block_type,
            "final_activation": "tanh",
            "spiking_feedforward_block_type": self.spiking_feedforward_block_type,
            "spiking_neuron": unet_kwargs["spiking_neuron"],
        }

        self.crop = None
        self.mask = unet_kwargs["mask_output"]
        self.norm_input = False if "norm_input" not in unet_kwargs.keys() else unet_kwargs["norm_input"]
        self.encoding = unet_kwargs["encoding"]
        self.num_bins = unet_kwargs["num_bins"]
        self.num_encoders = RecEVFlowNet_kwargs["num_encoders"]

        unet_kwargs.update(RecEVFlowNet_kwargs)  # in place on the caller's dict, like the reference (model.py:456-461)
        unet_kwargs.pop("name", None)
        unet_kwargs.pop("encoding", None)
        unet_kwargs.pop("round_encoding", None)
        unet_kwargs.pop("norm_input", None)
        unet_kwargs.pop("mask_output", None)

        self.multires_unetrec = self.unet_type(unet_kwargs)

    @property
    def states(self):
        return copy_states(self.multires_unetrec.states)

    @states.setter
    def states(self, states):
        self.multires_unetrec.states = states

    def detach_states(self):
        detached_states = []
        for state in self.multires_unetrec.states:
            if type(state) is tuple:
                detached_states.append(tuple(hidden.detach() for hidden in state))
            else:
                detached_states.append(state.detach())
        self.multires_unetrec.states = detached_states

    def reset_states(self):
        self.multires_unetrec.states = [None] * self.multires_unetrec.num_states

    def init_cropping(self, width, height, safety_margin=0):
        self.crop = CropParameters(width, height, self.num_encoders, safety_margin)

    def forward(self, event_voxel, event_cnt, log=False):
        """
        :param event_voxel: N x num_bins x H x W
        :param event_cnt: N x 2 x H x W per-polarity event counts
        :return {"flow": [N x 2 x H x W] * num_encoders (coarse to fine, all at input resolution), "activity": None}
        """
        if self.encoding == "voxel":
            x = event_voxel
        elif self.encoding == "cnt" and self.num_bins == 2:
            x = event_cnt
        else:
            print("Model error: Incorrect input encoding.")
            raise AttributeError

        if self.norm_input:
            mean, stddev = x[x != 0].mean(), x[x != 0].std()
            x[x != 0] = (x[x != 0] - mean) / stddev

        if self.crop is not None:
            x = self.crop.pad(x)

        multires_flow = self.multires_unetrec.forward(x)

        if log:
            raise NotImplementedError("Activity logging not implemented")
        activity = None

        flow_list = []
        full_h, full_w = multires_flow[-1].shape[2], multires_flow[-1].shape[3]
        for flow in multires_flow:
            flow_list.append(ops.upsample_nearest(flow, full_h // flow.shape[2], full_w // flow.shape[3]))

        if self.crop is not None:
            for i, flow in enumerate(flow_list):
                flow_list[i] = flow[:, :, self.crop.iy0:self.crop.iy1, self.crop.ix0:self.crop.ix1].contiguous()

        return {"flow": flow_list, "activity": activity}


class SpikingRecEVFlowNet(RecEVFlowNet):
    """models/model.py:550-558."""

    unet_type = SpikingMultiResUNetRecurrent
    recurrent_block_type = "lif"
    spiking_feedforward_block_type = "lif"


class PLIFRecEVFlowNet(RecEVFlowNet):
    """models/model.py:561-569."""

    unet_type = SpikingMultiResUNetRecurrent
    recurrent_block_type = "plif"
    spiking_feedforward_block_type = "plif"


class ALIFRecEVFlowNet(RecEVFlowNet):
    """models/model.py:572-580."""

    unet_type = SpikingMultiResUNetRecurrent
    recurrent_block_type = "alif"
    spiking_feedforward_block_type = "alif"


class XLIFRecEVFlowNet(RecEVFlowNet):
    """models/model.py:583-591."""

    unet_type = SpikingMultiResUNetRecurrent
    recurrent_block_type = "xlif"
    spiking_feedforward_block_type = "xlif"


class RNNRecEVFlowNet(RecEVFlowNet):
    """models/model.py:594-601: ConvRecurrent instead of ConvGRU after every encoder conv."""

    unet_type = MultiResUNetRecurrent
    recurrent_block_type = "convrnn"


class LeakyRecEVFlowNet(RecEVFlowNet):
    """models/model.py:604-611: leaky (stateful ANN) cells throughout."""

    unet_type = LeakyMultiResUNetRecurrent
    recurrent_block_type = "convleaky"


class RNNFireNet(FireNet):
    """models/model.py:614-622."""

    head_neuron = ConvLayer_
    ff_neuron = ConvLayer_
    rec_neuron = ConvRecurrent
    residual = False


class LeakyFireNet(FireNet):
    """models/model.py:625-633."""

    head_neuron = ConvLeaky
    ff_neuron = ConvLeaky
    rec_neuron = ConvLeakyRecurrent
    residual = False


class LeakyFireFlowNet(FireNet):
    """models/model.py:696-704."""

    head_neuron = ConvLeaky
    ff_neuron = ConvLeaky
    rec_neuron = ConvLeaky
    residual = False


class E2VID(BaseModel):
    """E2VID adapted for flow (models/model.py:29-145): recurrent U-Net with ConvLSTM encoders and sum skips, one flow map."""

    def __init__(self, unet_kwargs):
        super().__init__()
        norm = None
        use_upsample_conv = True
        if "norm" in unet_kwargs.keys():
            norm = unet_kwargs["norm"]
        if "use_upsample_conv" in unet_kwargs.keys():
            use_upsample_conv = unet_kwargs["use_upsample_conv"]
        E2VID_kwargs = {
            "base_num_channels": unet_kwargs["base_num_channels"],
            "num_encoders": 3,
            "num_residual_blocks": 2,
            "num_output_channels": 2,
            "skip_type": "sum",
            "norm": norm,
            "use_upsample_conv": use_upsample_conv,
            "kernel_size": unet_kwargs["kernel_size"],
            "channel_multiplier": 2,
            "recurrent_block_type": "convlstm",
            "final_activation": "tanh",
        }
        self.crop = None
        self.mask = unet_kwargs["mask_output"]
        self.norm_input = False if "norm_input" not in unet_kwargs.keys() else unet_kwargs["norm_input"]
        self.encoding = unet_kwargs["encoding"]
        self.num_bins = unet_kwargs["num_bins"]
        self.num_encoders = E2VID_kwargs["num_encoders"]
        unet_kwargs.update(E2VID_kwargs)
        for k in ("name", "encoding", "round_encoding", "norm_input", "mask_output", "spiking_neuron"):
            unet_kwargs.pop(k, None)
        self.unetrecurrent = UNetRecurrent(unet_kwargs)

    @property
    def states(self):
        return copy_states(self.unetrecurrent.states)

    @states.setter
    def states(self, states):
        self.unetrecurrent.states = states

    def detach_states(self):
        detached_states = []
        for state in self.unetrecurrent.states:
            if type(state) is tuple:
                detached_states.append(tuple(hidden.detach() for hidden in state))
            else:
                detached_states.append(state.detach())
        self.unetrecurrent.states = detached_states

    def reset_states(self):
        self.unetrecurrent.states = [None] * self.unetrecurrent.num_states

    def init_cropping(self, width, height, safety_margin=0):
        self.crop = CropParameters(width, height, self.num_encoders, safety_margin)

    def forward(self, event_voxel, event_cnt, log=False):
        if self.encoding == "voxel":
            x = event_voxel
        elif self.encoding == "cnt" and self.num_bins == 2:
            x = event_cnt
        else:
            print("Model error: Incorrect input encoding.")
            raise AttributeError
        if self.norm_input:
            mean, stddev = x[x != 0].mean(), x[x != 0].std()
            x[x != 0] = (x[x != 0] - mean) / stddev
        if self.crop is not None:
            x = self.crop.pad(x)
        flow = self.unetrecurrent.forward(x)
        if log:
            raise NotImplementedError("Activity logging not implemented")
        if self.crop is not None:
            flow = flow[:, :, self.crop.iy0:self.crop.iy1, self.crop.ix0:self.crop.ix1].contiguous()
        return {"flow": [flow], "activity": None}
