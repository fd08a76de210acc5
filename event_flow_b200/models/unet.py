"""
Spiking multi-resolution recurrent U-Net with the constructor contract, module names (state_dict keys) and state layout of
models/unet.py: BaseUNet (:28-120), MultiResUNetRecurrent (:314-416), SpikingMultiResUNetRecurrent (:418-465).
Every encoder / residual / decoder stage is built from the fused conv+neuron cells (ef_lif_conv_fwd, backward ef_lif_conv_bwd incl.
stride 2); the 2x bilinear upsampling is ef_upsample_bilinear2x(_bwd) and the per-scale prediction ef_pred_fwd / ef_pred_bwd: forward
and BPTT.  Without gradient tracking the LIF variant runs on the general tensor-core cell kernel (fast_unet.py: no concat, stride 2 through
space-to-depth inputs).  Also here: the ANN twins (MultiResUNet, MultiResUNetRecurrent), the leaky U-Net and E2VID's UNetRecurrent.
"""
import torch.nn as nn

from .. import fast_unet
from .model_util import skip_concat, skip_sum  # noqa: F401  (resolved by name like the reference: "skip_" + skip_type)
from .spiking_submodules import (SpikingRecurrentConvLayer, SpikingResidualBlock, SpikingTransposedConvLayer, SpikingUpsampleConvLayer,
                                 _SpikingConvCell)
from .submodules import (ConvLayer, LeakyRecurrentConvLayer, LeakyResidualBlock, LeakyTransposedConvLayer, LeakyUpsampleConvLayer,
                         RecurrentConvLayer, ResidualBlock, TransposedConvLayer, UpsampleConvLayer)


class SpikingMultiResUNetRecurrent(nn.Module):
    ff_type = ConvLayer
    res_type = SpikingResidualBlock
    upsample_type = SpikingUpsampleConvLayer
    transpose_type = SpikingTransposedConvLayer
    rec_type = SpikingRecurrentConvLayer
    w_scale_pred = 0.01

    def __init__(self, unet_kwargs):
        super().__init__()
        kw = dict(unet_kwargs)
        self.final_activation = kw.pop("final_activation", None)
        self.base_num_channels = kw["base_num_channels"]
        self.num_encoders = kw["num_encoders"]
        self.num_residual_blocks = kw["num_residual_blocks"]
        self.num_output_channels = kw["num_output_channels"]
        self.kernel_size = kw.get("kernel_size", 5)
        self.skip_type = kw["skip_type"]
        self.norm = kw["norm"]
        self.num_bins = kw["num_bins"]
        self.recurrent_block_type = kw.get("recurrent_block_type")
        self.channel_multiplier = kw.get("channel_multiplier", 2)
        self.ff_act, self.rec_act = kw.get("activations", ["relu", None])
        if self.kernel_size != 3:
            raise NotImplementedError("event_flow_b200: the spiking U-Net is built for kernel_size 3 (got %r)" % (self.kernel_size,))

        self.spiking_kwargs = {}
        if kw.get("spiking_feedforward_block_type") is not None:
            self.spiking_kwargs["spiking_feedforward_block_type"] = kw["spiking_feedforward_block_type"]
        if type(kw.get("spiking_neuron")) is dict:
            self.spiking_kwargs.update(kw["spiking_neuron"])

        self.skip_ftn = {"concat": skip_concat, "sum": skip_sum}[self.skip_type]
        self.UpsampleLayer = self.upsample_type if kw["use_upsample_conv"] else self.transpose_type
        assert self.num_output_channels > 0

        self.encoder_input_sizes = [int(self.base_num_channels * pow(self.channel_multiplier, i)) for i in range(self.num_encoders)]
        self.encoder_output_sizes = [int(self.base_num_channels * pow(self.channel_multiplier, i + 1)) for i in range(self.num_encoders)]
        self.max_num_channels = self.encoder_output_sizes[-1]

        # construction order = the reference's (unet.py:329-332): same torch seed, same initial values
        self.encoders = self.build_recurrent_encoders()
        self.resblocks = self.build_resblocks()
        self.decoders = self.build_multires_prediction_decoders()
        self.preds = self.build_multires_prediction_layer()
        self.num_states = self.num_encoders * 2 + self.num_residual_blocks
        self.states = [None] * self.num_states
        self._mark_inputs()

    def _mark_inputs(self):
        """
        Tell the cells what their input is (as FireNet._mark_inputs does), so that LIF cells may run on the tensor cores with exact
        products under autograd as well: every cell after the first sees spikes, sums of spikes or their bilinear x2 upsampling (multiples
        of 1/16) -- exact in bf16; the decoders after the first additionally see the upsampled flow prediction in their first
        `num_output_channels` input channels (cat order [prediction, x, skip], forward() below), which enters as its exact split.
        """
        if self.skip_type != "concat":
            return

        def mark(owner, name, kind):
            cell = getattr(owner, name, None)
            if isinstance(cell, _SpikingConvCell):  # (the leaky ANN twins share this wiring with other cell classes)
                cell.__dict__["_x_kind"] = kind

        for i, enc in enumerate(self.encoders):
            if i > 0:
                mark(enc, "conv", "spikes")
            mark(enc, "recurrent_block", "spikes")
        for r in self.resblocks:
            mark(r, "conv1", "spikes"), mark(r, "conv2", "spikes")
        for i, d in enumerate(self.decoders):
            mark(d, "conv2d", "spikes" if i == 0 else ("mixed", self.num_output_channels))

    # The states live either in the reference's format (list of stacked fp32 tensors, what the cell-by-cell path reads and writes) or,
    # after a step of the tensor-core inference path (event_flow_b200/fast_unet.py), in the internal format (membrane fp32 + spikes
    # bf16 channels-last); the property converts lazily at this API boundary.
    @property
    def states(self):
        if self.__dict__.get("_tc_state") is not None:
            self._states = fast_unet.export_states(self)
            self.__dict__["_tc_state"] = None
        return self._states

    @states.setter
    def states(self, value):
        self._states = value
        self.__dict__["_tc_state"] = None

    def __getstate__(self):  # checkpoints / deepcopy: states in the reference format, no run-time caches
        state = self.__dict__.copy()
        if state.get("_tc_state") is not None:
            state["_states"] = [None if t is None else t.detach() for t in fast_unet.export_states(self)]
        for k in ("_tc_state", "_tc_images", "_tc_eligible"):
            state.pop(k, None)
        return state

    def build_recurrent_encoders(self):
        encoders = nn.ModuleList()
        for i, (input_size, output_size) in enumerate(zip(self.encoder_input_sizes, self.encoder_output_sizes)):
            if i == 0:
                input_size = self.num_bins
            kw = {k: v for k, v in self.spiking_kwargs.items()}
            encoders.append(self.rec_type(input_size, output_size, kernel_size=self.kernel_size, stride=2,
                                          recurrent_block_type=self.recurrent_block_type, activation_ff=self.ff_act, activation_rec=self.rec_act,
                                          norm=self.norm, **kw))
        return encoders

    def build_resblocks(self):
        resblocks = nn.ModuleList()
        for _ in range(self.num_residual_blocks):
            resblocks.append(self.res_type(self.max_num_channels, self.max_num_channels, activation=self.ff_act, norm=self.norm,
                                           **self.spiking_kwargs))
        return resblocks

    def build_multires_prediction_layer(self):
        preds = nn.ModuleList()
        for output_size in reversed(self.encoder_input_sizes):
            preds.append(self.ff_type(output_size, self.num_output_channels, 1, activation=self.final_activation, norm=self.norm,
                                      w_scale=self.w_scale_pred))
        return preds

    def build_multires_prediction_decoders(self):
        decoders = nn.ModuleList()
        for i, (input_size, output_size) in enumerate(zip(reversed(self.encoder_output_sizes), reversed(self.encoder_input_sizes))):
            prediction_channels = 0 if i == 0 else self.num_output_channels
            decoders.append(self.UpsampleLayer(2 * input_size + prediction_channels, output_size, kernel_size=self.kernel_size,
                                               activation=self.ff_act, norm=self.norm, **self.spiking_kwargs))
        return decoders

    def forward(self, x):
        """
        :param x: N x num_input_channels x H x W
        :return: [N x num_output_channels x H x W for i in range(self.num_encoders)]
        """
        if fast_unet.eligible(self, x):  # LIF cells, no gradient tracked: tcgen05 cells on the internal spike format
            if self.__dict__.get("_tc_state") is None:
                self._states = self.states  # (materialised reference-format list the importer reads)
            return fast_unet.forward(self, x)
        blocks = []
        offset = 0
        for i, encoder in enumerate(self.encoders):
            x, self.states[i] = encoder(x, self.states[i])
            blocks.append(x)

        offset += self.num_encoders
        for i, resblock in enumerate(self.resblocks):
            x, self.states[offset + i] = resblock(x, self.states[offset + i])

        predictions = []
        offset += self.num_residual_blocks
        for i, (decoder, pred) in enumerate(zip(self.decoders, self.preds)):
            x = self.skip_ftn(x, blocks[self.num_encoders - i - 1])
            if i > 0:
                x = self.skip_ftn(predictions[-1], x)
            x, self.states[offset + i] = decoder(x, self.states[offset + i])
            predictions.append(pred(x))
        return predictions


def state_list_of(net, x):
    """
    Where a U-Net keeps the plain list of its recurrent states, as (object, attribute name) -- (None, None) for the stateless U-Net --,
    or None if the step on `x` runs on the spiking U-Net's tensor-core inference path (states in its internal format).
    """
    if isinstance(net, SpikingMultiResUNetRecurrent):
        if fast_unet.eligible(net, x):
            return None
        net.states  # (materialises the reference-format list if the last step left the internal format)
        return net, "_states"
    return (net, "states") if hasattr(net, "states") else (None, None)


class MultiResUNet(nn.Module):
    """
    ANN multi-resolution U-Net of EV-FlowNet (models/unet.py:224-311): four stride-2 conv encoders, two residual blocks, four
    bilinear-upsampling (or, with use_upsample_conv=False, transposed-convolution) decoders with concat skips, a 1x1 tanh prediction
    per scale; optional BN / IN normalisation of every layer.
    """

    def __init__(self, unet_kwargs):
        super().__init__()
        kw = dict(unet_kwargs)
        self.final_activation = kw.pop("final_activation", None)
        self.base_num_channels = kw["base_num_channels"]
        self.num_encoders = kw["num_encoders"]
        self.num_residual_blocks = kw["num_residual_blocks"]
        self.num_output_channels = kw["num_output_channels"]
        self.kernel_size = kw.get("kernel_size", 5)
        self.skip_type = kw["skip_type"]
        self.norm = kw["norm"]
        self.num_bins = kw["num_bins"]
        self.channel_multiplier = kw.get("channel_multiplier", 2)
        self.ff_act, self.rec_act = kw.get("activations", ["relu", None])
        if self.norm not in (None, "BN", "IN") or self.kernel_size != 3:
            raise NotImplementedError("event_flow_b200 MultiResUNet: kernel_size 3, norm None / 'BN' / 'IN'")
        self.use_upsample_conv = kw["use_upsample_conv"]
        self.skip_ftn = {"concat": skip_concat, "sum": skip_sum}[self.skip_type]
        self.encoder_input_sizes = [int(self.base_num_channels * pow(self.channel_multiplier, i)) for i in range(self.num_encoders)]
        self.encoder_output_sizes = [int(self.base_num_channels * pow(self.channel_multiplier, i + 1)) for i in range(self.num_encoders)]
        self.max_num_channels = self.encoder_output_sizes[-1]
        self.recurrent_block_type = kw.get("recurrent_block_type")
        # construction order = the reference's (unet.py:236-239, 329-332)
        self.encoders = self.build_encoders()
        self.resblocks = nn.ModuleList([ResidualBlock(self.max_num_channels, self.max_num_channels, activation=self.ff_act, norm=self.norm)
                                        for _ in range(self.num_residual_blocks)])
        self.decoders = nn.ModuleList()
        for i, (cin, cout) in enumerate(zip(reversed(self.encoder_output_sizes), reversed(self.encoder_input_sizes))):
            up = UpsampleConvLayer if self.use_upsample_conv else TransposedConvLayer  # models/unet.py:77-81
            self.decoders.append(up(2 * cin + (0 if i == 0 else self.num_output_channels), cout, kernel_size=self.kernel_size,
                                    activation=self.ff_act, norm=self.norm))
        self.preds = nn.ModuleList([ConvLayer(cout, self.num_output_channels, 1, activation=self.final_activation, norm=self.norm)
                                    for cout in reversed(self.encoder_input_sizes)])

    def build_encoders(self):
        encoders = nn.ModuleList()
        for i, (cin, cout) in enumerate(zip(self.encoder_input_sizes, self.encoder_output_sizes)):
            encoders.append(ConvLayer(self.num_bins if i == 0 else cin, cout, kernel_size=self.kernel_size, stride=2, activation=self.ff_act,
                                      norm=self.norm))
        return encoders

    def encode(self, x):
        blocks = []
        for encoder in self.encoders:
            x = encoder(x)
            blocks.append(x)
        return x, blocks

    def forward(self, x):
        x, blocks = self.encode(x)
        for resblock in self.resblocks:
            x, _ = resblock(x)
        predictions = []
        for i, (decoder, pred) in enumerate(zip(self.decoders, self.preds)):
            x = self.skip_ftn(x, blocks[self.num_encoders - i - 1])
            if i > 0:
                x = self.skip_ftn(predictions[-1], x)
            x = decoder(x)
            predictions.append(pred(x))
        return predictions


class MultiResUNetRecurrent(MultiResUNet):
    """Recurrent ANN U-Net (models/unet.py:314-416): every stride-2 encoder conv is followed by a ConvGRU; one state per encoder."""

    def __init__(self, unet_kwargs):
        super().__init__(unet_kwargs)
        self.num_states = self.num_encoders
        self.states = [None] * self.num_states

    def build_encoders(self):
        encoders = nn.ModuleList()
        for i, (cin, cout) in enumerate(zip(self.encoder_input_sizes, self.encoder_output_sizes)):
            encoders.append(RecurrentConvLayer(self.num_bins if i == 0 else cin, cout, kernel_size=self.kernel_size, stride=2,
                                               recurrent_block_type=self.recurrent_block_type, activation_ff=self.ff_act,
                                               activation_rec=self.rec_act, norm=self.norm))
        return encoders

    def encode(self, x):
        blocks = []
        for i, encoder in enumerate(self.encoders):
            x, self.states[i] = encoder(x, self.states[i])
            blocks.append(x)
        return x, blocks


class LeakyMultiResUNetRecurrent(SpikingMultiResUNetRecurrent):
    """Leaky (stateful ANN) twin of the spiking U-Net (models/unet.py:468-480): same wiring, ConvLeaky / ConvLeakyRecurrent cells."""

    res_type = LeakyResidualBlock
    upsample_type = LeakyUpsampleConvLayer
    transpose_type = LeakyTransposedConvLayer
    rec_type = LeakyRecurrentConvLayer


class UNetRecurrent(nn.Module):
    """
    E2VID's recurrent U-Net (models/unet.py:148-222): stride-1 head conv, stride-2 conv + ConvLSTM encoders, residual blocks,
    upsampling decoders with SUM skips, one full-resolution 1x1 prediction + final activation.
    """

    def __init__(self, unet_kwargs):
        super().__init__()
        kw = dict(unet_kwargs)
        self.final_activation = kw.pop("final_activation", "none")
        if self.final_activation != "tanh":
            raise NotImplementedError("event_flow_b200 UNetRecurrent: final_activation tanh (E2VID) only")
        self.base_num_channels = kw["base_num_channels"]
        self.num_encoders = kw["num_encoders"]
        self.num_residual_blocks = kw["num_residual_blocks"]
        self.num_output_channels = kw["num_output_channels"]
        self.kernel_size = kw.get("kernel_size", 5)
        self.skip_type = kw["skip_type"]
        self.norm = kw["norm"]
        self.num_bins = kw["num_bins"]
        self.recurrent_block_type = kw.get("recurrent_block_type")
        self.channel_multiplier = kw.get("channel_multiplier", 2)
        self.ff_act, self.rec_act = kw.get("activations", ["relu", None])
        if self.norm is not None or self.kernel_size != 3 or not kw["use_upsample_conv"]:
            raise NotImplementedError("event_flow_b200 UNetRecurrent: kernel_size 3, no norm, upsample-conv decoders only")
        self.skip_ftn = {"concat": skip_concat, "sum": skip_sum}[self.skip_type]
        self.encoder_input_sizes = [int(self.base_num_channels * pow(self.channel_multiplier, i)) for i in range(self.num_encoders)]
        self.encoder_output_sizes = [int(self.base_num_channels * pow(self.channel_multiplier, i + 1)) for i in range(self.num_encoders)]
        self.max_num_channels = self.encoder_output_sizes[-1]
        mult = 1 if self.skip_type == "sum" else 2
        # construction order = the reference's (unet.py:161-172)
        self.head = ConvLayer(self.num_bins, self.base_num_channels, kernel_size=self.kernel_size, stride=1)
        self.encoders = nn.ModuleList([RecurrentConvLayer(cin, cout, kernel_size=self.kernel_size, stride=2,
                                                          recurrent_block_type=self.recurrent_block_type, activation_ff=self.ff_act,
                                                          activation_rec=self.rec_act, norm=self.norm)
                                       for cin, cout in zip(self.encoder_input_sizes, self.encoder_output_sizes)])
        self.resblocks = nn.ModuleList([ResidualBlock(self.max_num_channels, self.max_num_channels, activation=self.ff_act, norm=self.norm)
                                        for _ in range(self.num_residual_blocks)])
        self.decoders = nn.ModuleList([UpsampleConvLayer(mult * cin, cout, kernel_size=self.kernel_size, activation=self.ff_act, norm=self.norm)
                                       for cin, cout in zip(reversed(self.encoder_output_sizes), reversed(self.encoder_input_sizes))])
        # reference: ConvLayer(k=1, activation=None) followed by the final tanh -- one fused 1x1 + tanh launch here
        self.pred = ConvLayer(mult * self.base_num_channels, self.num_output_channels, 1, activation="tanh", norm=self.norm)
        self.num_states = self.num_encoders
        self.states = [None] * self.num_states

    def forward(self, x):
        x = self.head(x)
        head = x
        blocks = []
        for i, encoder in enumerate(self.encoders):
            x, state = encoder(x, self.states[i])
            blocks.append(x)
            self.states[i] = state
        for resblock in self.resblocks:
            x, _ = resblock(x)
        for i, decoder in enumerate(self.decoders):
            x = decoder(self.skip_ftn(x, blocks[self.num_encoders - i - 1]))
        return self.pred(self.skip_ftn(x, head))
