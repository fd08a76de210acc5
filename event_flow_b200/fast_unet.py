"""
Tensor-core inference path of the spiking EV-FlowNet (SpikingRecEVFlowNet: models/unet.py:418-465, spiking_submodules.py:878-1013).

Like the FireNet fast path, spikes travel between the cells as bf16 channels-last tensors and membrane potentials stay fp32 NCHW.
Per step:
  * encoders  : ConvLIF stride 2 (CUDA-core kernel, cl in / cl out) + ConvLIFRecurrent C -> C on the general tcgen05 cell kernel
                (ef_lif_conv_fwd_g; the recurrent convolution is a second input source of the same launch);
  * resblocks : two tcgen05 cells, the second one adds the block input to its spikes inside the kernel;
  * decoders  : the reference upsamples cat[prediction, x, skip] bilinearly and convolves it (unet.py:451-462) -- here every part is
                upsampled on its own (spike tensors exactly in bf16, the fractional flow prediction in fp32 followed by an exact
                hi/mid/lo split) and enters the cell kernel as its own source: the concat is never built;
  * predictions: ef_pred_fwd on the decoder spikes.
Used when no gradient is tracked (evaluation, eval_flow.py:119) for LIF U-Nets whose pyramid divides the input; training keeps the
fp32 cell path (its backward kernels need the fp32 activations).  States are kept in the internal format and converted lazily at
the `states` API boundary.
"""
import torch

from . import _lib as L
from . import ops


def eligible(net, x):
    if not x.is_cuda or torch.is_grad_enabled() and any(p.requires_grad for p in net.parameters()):
        return False
    ok = net.__dict__.get("_tc_eligible")
    if ok is None:
        try:
            cells = _all_cells(net)
            ok = all(getattr(c, "neuron", None) == "lif" and c.ff.kernel_size == (3, 3) and c.plain() for c in cells)
            ok = ok and all(c.hidden_size % 32 == 0 for c in cells) and net.skip_type == "concat" and net.num_output_channels <= L.EF_HEAD_MAX_CIN
            ok = ok and all(type(d).__name__ == "SpikingUpsampleConvLayer" for d in net.decoders)
        except AttributeError:  # other cell families (leaky ANN twins) share the wiring but not the kernels
            ok = False
        net.__dict__["_tc_eligible"] = ok
    div = 2 ** net.num_encoders
    H, W = x.shape[-2:]
    return bool(ok) and net.__dict__.get("_use_tc", True) and H % div == 0 and W % div == 0 and (W // div) % 4 == 0


def _all_cells(net):
    cells = []
    for e in net.encoders:
        cells += [e.conv, e.recurrent_block]
    for r in net.resblocks:
        cells += [r.conv1, r.conv2]
    for d in net.decoders:
        cells.append(d.conv2d)
    return cells


class _Cell:
    """Internal state of one cell: membrane fp32 NCHW, spikes bf16 channels-last."""

    __slots__ = ("v", "z")

    def __init__(self, v=None, z=None):
        self.v, self.z = v, z


def _weights(net, key, cell, sources):
    """Cached weight image of a cell for a given source list; rebuilt when a weight tensor changed."""
    cache = net.__dict__.setdefault("_tc_images", {})
    ver = tuple((src[0]._version, src[0].data_ptr()) for src in sources)
    hit = cache.get(key)
    if hit is None or hit[0] != ver:
        hit = (ver, ops.split_weights_g(sources, cell.hidden_size))
        cache[key] = hit
    return hit[1]


def _chan(cell):
    return cell.leak.detach().reshape(-1), cell.thresh.detach().reshape(-1)


def _ref_state(st):
    return None if st.v is None else torch.stack([st.v, ops.unpack_cl(st.z)]).cpu()


def _capture(net, name, x_in, st_before, residual, out_cl, st, stride):
    """Test hook (net._capture = list): what the cell consumed and produced, in the reference's tensor formats, as CPU tensors."""
    cap = net.__dict__.get("_capture")
    if cap is not None:
        cap.append((net.__dict__.get("_capture_prefix", "") + name, x_in.detach().cpu(), st_before, 0 if residual is None else ops.unpack_cl(residual).cpu(), ops.unpack_cl(out_cl).cpu(),
                    _ref_state(st), stride))


def _src_as_f32(src, wsrc):
    """A source tensor back in fp32 NCHW with its real channel count (split sources: hi + mid + lo)."""
    _, _, n, split = wsrc
    x = src.float()
    if split:
        SL = 8 if n <= 8 else 10
        x = x[..., 0:n] + x[..., SL:SL + n] + x[..., 2 * SL:2 * SL + n]
    else:
        x = x[..., :n]
    return x.permute(0, 3, 1, 2).contiguous()


def _step_g(net, key, cell, st, srcs, wsrcs, residual=None, name=None):
    """One cell on the general tensor-core kernel; the recurrent convolution joins as a source when there is a previous state."""
    capturing = net.__dict__.get("_capture") is not None
    if capturing:
        x_in, before = torch.cat([_src_as_f32(a, b) for a, b in zip(srcs, wsrcs)], 1), _ref_state(st)
    out = _step_g_impl(net, key, cell, st, srcs, wsrcs, residual)
    if capturing:
        _capture(net, name, x_in, before, residual, out, st, 1)
    return out


def _step_g_impl(net, key, cell, st, srcs, wsrcs, residual):
    if getattr(cell, "recurrent", False) and st.z is not None:
        srcs = srcs + [st.z]
        wsrcs = wsrcs + [(cell.rec.weight, 0, cell.hidden_size, False)]
        key = key + ("rec",)
    img = _weights(net, key, cell, wsrcs)
    leak, thresh = _chan(cell)
    st.v, st.z, out = ops.lif_step_g(srcs, st.v, st.z, img, leak, thresh, cell.hidden_size, hard_reset=cell.hard_reset, residual_cl=residual)
    return st.z if out is None else out


def forward(net, x):
    """One forward pass of a SpikingMultiResUNetRecurrent.  x [B,num_bins,H,W] fp32 -> list of predictions [B,2,h,w] (coarse to fine)."""
    S = net.__dict__.get("_tc_state")
    if S is None:
        S = net.__dict__["_tc_state"] = _import_states(net)
    blocks = []
    h = None  # cl spikes
    n_enc = net.num_encoders
    for i, enc in enumerate(net.encoders):
        ff, rec = S[2 * i], S[2 * i + 1]
        c = enc.conv
        leak, thresh = _chan(c)
        capturing = net.__dict__.get("_capture") is not None
        before = _ref_state(ff) if capturing else None
        x_in = x if i == 0 else h
        cin = c.input_size
        if c.stride == 2 and (i > 0 or 4 * cin <= L.EF_HEAD_MAX_CIN):
            # stride-2 cell on the tensor cores: a stride-1 cell at the output resolution over the space-to-depth form of the input
            # (the fp32 network input additionally as its exact hi/mid/lo split)
            src = ops.pack_split_s2d_cl(x) if i == 0 else ops.space_to_depth_cl(h)
            img = _weights(net, ("enc_ff", i), c, [(c.ff.weight, 0, cin, i == 0, True)])
            ff.v, ff.z, _ = ops.lif_step_g([src], ff.v, ff.z, img, leak, thresh, c.hidden_size, hard_reset=c.hard_reset, s2d=True)
        elif i == 0:
            ff.v, ff.z = ops.lif_step_cl(None, ff.v, ff.z, c.ff.weight, None, leak, thresh, hard_reset=c.hard_reset, x_f32=x.contiguous(), stride=c.stride)
        else:
            ff.v, ff.z = ops.lif_step_cl(h, ff.v, ff.z, c.ff.weight, None, leak, thresh, hard_reset=c.hard_reset, stride=c.stride)
        if capturing:
            _capture(net, f"encoders.{i}.conv", x_in if i == 0 else ops.unpack_cl(x_in), before, None, ff.z, ff, c.stride)
        rb = enc.recurrent_block
        h = _step_g(net, ("enc", i), rb, rec, [ff.z], [(rb.ff.weight, 0, rb.input_size, False)], name=f"encoders.{i}.recurrent_block")
        blocks.append(h)
    off = 2 * n_enc
    for i, rbk in enumerate(net.resblocks):
        s1, s2 = S[off + 2 * i], S[off + 2 * i + 1]
        x1 = _step_g(net, ("res1", i), rbk.conv1, s1, [h], [(rbk.conv1.ff.weight, 0, rbk.conv1.input_size, False)], name=f"resblocks.{i}.conv1")
        h = _step_g(net, ("res2", i), rbk.conv2, s2, [x1], [(rbk.conv2.ff.weight, 0, rbk.conv2.input_size, False)], residual=h,
                    name=f"resblocks.{i}.conv2")
    off += 2 * len(net.resblocks)
    preds = []
    for i, (dec, pred) in enumerate(zip(net.decoders, net.preds)):
        cell = dec.conv2d
        skip = blocks[n_enc - i - 1]
        cx, cs = h.shape[3], skip.shape[3]
        srcs, wsrcs, ch = [], [], 0
        if i > 0:  # cat[prediction, x, skip]: the prediction is fractional fp32 -> upsample in fp32, then the exact three-way split
            npred = preds[-1].shape[1]
            srcs.append(ops.pack_split_cl(ops.upsample_bilinear2x(preds[-1])))
            wsrcs.append((cell.ff.weight, 0, npred, True))
            ch = npred
        srcs += [ops.upsample_bilinear2x_cl(h), ops.upsample_bilinear2x_cl(skip)]
        wsrcs += [(cell.ff.weight, ch, cx, False), (cell.ff.weight, ch + cx, cs, False)]
        h = _step_g(net, ("dec", i), cell, S[off + i], srcs, wsrcs, name=f"decoders.{i}.conv2d")
        w = pred.conv2d.weight.detach()
        preds.append(_pred(h, w.reshape(w.shape[0], -1), pred.conv2d.bias.detach()))
    return preds


def _pred(x_cl, w, b):
    B, H, W, Cin = x_cl.shape
    y = torch.empty((B, w.shape[0], H, W), device=x_cl.device, dtype=torch.float32)
    p = L.PredParams()
    p.B, p.Cin, p.Cout, p.H, p.W = B, Cin, w.shape[0], H, W
    p.x_cl, p.w, p.b, p.y = L.ptr(x_cl), L.ptr(w.contiguous()), L.ptr(b.contiguous()), L.ptr(y)
    L.call("ef_pred_fwd", p)
    return y


# ---- state API boundary: reference format <-> internal format ---------------------------------------------------------------------
def _n_cells(net):
    return 2 * net.num_encoders + 2 * len(net.resblocks) + len(net.decoders)


def _import_states(net):
    """Reference-format states (net._states: stacked fp32 tensors, two cells per encoder / residual block) -> internal cells."""
    out = []
    ref = net._states
    n_pair = net.num_encoders + len(net.resblocks)
    for i in range(n_pair):
        s = ref[i]
        for k in range(2):
            if s is None or s[k] is None:
                out.append(_Cell())
            else:
                out.append(_Cell(s[k][0].detach().contiguous(), ops.pack_cl(s[k][1].detach())))
    for i in range(len(net.decoders)):
        s = ref[n_pair + i]
        out.append(_Cell() if s is None else _Cell(s[0].detach().contiguous(), ops.pack_cl(s[1].detach())))
    return out


def export_states(net):
    """Internal cells -> the reference's state list (fresh tensors): stack([stack([v, z]) x 2]) per encoder / residual block, stack([v, z]) per decoder."""
    S = net.__dict__["_tc_state"]
    one = lambda c: None if c.v is None else torch.stack([c.v, ops.unpack_cl(c.z)])  # noqa: E731
    out = []
    n_pair = net.num_encoders + len(net.resblocks)
    for i in range(n_pair):
        a, b = one(S[2 * i]), one(S[2 * i + 1])
        out.append(None if a is None else torch.stack([a, b]))
    for i in range(len(net.decoders)):
        out.append(one(S[2 * n_pair + i]))
    return out
