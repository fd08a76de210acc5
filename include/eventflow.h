/*
 * eventflow.h -- C ABI of libeventflow.so: B200 (sm_100a) kernels for the two hot paths of tudelft/event_flow.
 *
 * The reference is pure Python/PyTorch and has no FFI; what each entry point REPLACES is therefore a span of the
 * reference's Python (file:line into the reference tree), listed with every declaration.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - plain C types only; every tensor pointer is a DEVICE pointer into caller-owned memory (torch storage);
 *     the library never allocates, frees or keeps a pointer after the call returns;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); work is enqueued asynchronously;
 *   - returns 0 on success, <0 for an argument / shape error, >0 = the cudaError_t of a failed launch;
 *     ef_last_error() returns a thread-local message for the last non-zero return of the calling thread;
 *   - fp32 tensors are dense NCHW.  "cl" tensors are the internal spike format: bf16, channels-last [B, H, W, C]
 *     (all channels of a pixel contiguous: the layout TMA boxes and tcgen05 K-major operands want), exact for the
 *     values a spiking layer emits ({0,1}, or {0,1,2} with a residual); C must be a multiple of 8;
 *   - a NULL optional pointer means "absent" (zero state, no residual, output not wanted).
 */
#ifndef EVENTFLOW_H_
#define EVENTFLOW_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EF_VERSION 200 /* 0.2.0 */

/* neuron models: models/spiking_submodules.py ConvLIF :24, ConvPLIF :129, ConvALIF :230, ConvXLIF :337 (+Recurrent) */
enum { EF_LIF = 0, EF_PLIF = 1, EF_ALIF = 2, EF_XLIF = 3 };
/* surrogate gradients: models/spiking_util.py ArctanSpike :82, SuperSpike :28, TriangleSpike :68, MultiGaussSpike :46 */
enum { EF_ARCTAN = 0, EF_SUPERSPIKE = 1, EF_TRIANGLE = 2, EF_MULTIGAUSS = 3 };
/* error codes */
enum { EF_OK = 0, EF_EINVAL = -1, EF_EUNSUPPORTED = -2, EF_ENULL = -3 };

int ef_version(void);
const char* ef_last_error(void);
/* 1 if the visible device is compute capability 10.x (tcgen05/TMA kernels usable), 0 if another GPU, <0 on error. */
int ef_device_ok(void);
/* number of CUDA kernels this library has launched in the calling process (monotonic; memsets not counted). */
uint64_t ef_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused conv3x3 + spiking-neuron update, one timestep of one cell.
 * Replaces the whole forward() of ConvLIF / ConvLIFRecurrent / ConvPLIF(Recurrent) / ConvALIF(Recurrent) /
 * ConvXLIF(Recurrent): models/spiking_submodules.py:96-126, 191-227, 299-334, 399-435, 516-551, 618-657, 730-768,
 * 836-875, including the spike function models/spiking_util.py:19-21,96-109.
 *
 *   I      = conv(x, w_ff, stride, pad k/2) [+ conv(z_in, w_rec, 1, pad k/2)]
 *   v_out  = LIF/PLIF/ALIF/XLIF update of (v_in, z_in, aux_in) with hard or soft reset
 *   z_out  = (v_out - thresh_t > 0);   out = z_out + residual
 *
 * Inputs may be given as fp32 NCHW or as cl bf16; when BOTH x_cl (and z_in_cl) are given, C == 32, Cin == 32,
 * ksize == 3, stride == 1, the tcgen05 tensor-core kernel runs (3-way bf16 split of the weights, fp32-exact products);
 * otherwise the fp32 CUDA-core kernel runs.  Per-channel parameter arrays are the RAW parameters of the reference
 * module ([C] floats; sigmoid / clamp are applied inside, spiking_submodules.py:108-112).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_lif_conv_params {
  int32_t B, Cin, C, H, W;      /* x is [B,Cin,H,W]; outputs are [B,C,Ho,Wo], Ho = (H-1)/stride+1                    */
  int32_t ksize, stride;        /* ksize 3 (pad 1); stride 1 or 2                                                     */
  int32_t neuron, hard_reset;   /* EF_LIF.. ; 1 = hard reset, 0 = soft reset                                          */
  int32_t surrogate;            /* EF_ARCTAN.. (backward only)                                                        */
  float act_width;              /* surrogate width (buffer act_width)                                                 */
  /* inputs */
  const float* x;               /* [B,Cin,H,W] fp32, or NULL when x_cl is given                                       */
  const uint16_t* x_cl;         /* [B,H,W,Cin] bf16, or NULL                                                          */
  const float* v_in;            /* [B,C,Ho,Wo] or NULL (zeros)                                                        */
  const float* z_in;            /* [B,C,Ho,Wo] fp32 previous spikes or NULL                                           */
  const uint16_t* z_in_cl;      /* same in cl, or NULL                                                                */
  const float* aux_in;          /* third state (PLIF/XLIF trace pt, ALIF trace t) or NULL                             */
  const float* w_ff;            /* [C,Cin,k,k]                                                                        */
  const float* w_rec;           /* [C,C,k,k] or NULL (non-recurrent cell)                                             */
  const float* leak;            /* [C] leak (LIF) / leak_v                                                            */
  const float* thresh;          /* [C] thresh (LIF, PLIF) or NULL                                                     */
  const float* leak_aux;        /* [C] leak_pt (PLIF, XLIF) / leak_t (ALIF) or NULL                                   */
  const float* add_pt;          /* [C] add_pt (PLIF) or NULL                                                          */
  const float* t0;              /* [C] t0 (ALIF, XLIF) or NULL                                                        */
  const float* t1;              /* [C] t1 (ALIF, XLIF) or NULL                                                        */
  const float* residual;        /* [B,C,Ho,Wo] fp32 or NULL                                                           */
  /* tensor-core path only: weights pre-split by ef_split_weights (bf16 hi/mid/lo, UMMA layout) or NULL (CUDA cores)  */
  const uint16_t* w_split;
  /* outputs (each optional except v_out) */
  float* v_out;                 /* [B,C,Ho,Wo]                                                                        */
  float* z_out;                 /* [B,C,Ho,Wo] fp32 spikes (state plane 1) or NULL                                    */
  uint16_t* z_out_cl;           /* spikes in cl or NULL                                                               */
  float* aux_out;               /* third state or NULL                                                                */
  float* out;                   /* z_out + residual, fp32, or NULL                                                    */
  uint16_t* out_cl;             /* z_out + residual in cl or NULL (only differs from z_out_cl with a residual)        */
} ef_lif_conv_params;

int ef_lif_conv_fwd(const ef_lif_conv_params* p, void* stream);

/* The neuron update of a cell step on a synaptic current computed elsewhere: cur [B,C,H,W] = conv(x, w_ff) (+ conv(z_in, w_rec)), e.g. the
 * membrane output of the tensor-core kernel run as a pure convolution (zero state, leak = -inf: v_out = current).  Same parameter block as
 * ef_lif_conv_fwd -- state in / out, per-channel parameters, residual, out / z_out (fp32 NCHW and / or channels-last bf16) -- for all four
 * neuron kinds (models/spiking_submodules.py:108-126, 200-227, 310-334, 409-435 and the recurrent twins); stride 1; x (fp32 NCHW) is
 * only read for the pre-synaptic trace of PLIF / XLIF; w_ff / w_rec are not read.  This is how 32-channel PLIF / ALIF / XLIF cells get
 * their convolution onto the tensor cores (event_flow_b200/ops.py, _CellStep). */
int ef_lif_neuron_fwd(const ef_lif_conv_params* p, const float* cur, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * A FEED-FORWARD 32 -> 32 LIF cell (or the head layer on split inputs, ef_pack_split_cl) over a whole window of T steps in ONE launch
 * (the time loop of train_flow.py:98-141 / models/model.py:255-265 moved inside the kernel for the cells without a recurrent
 * convolution: head, R1a, R1b, R2a, R2b).  Every tile runs its T steps back to back; the membrane potential and the previous spikes
 * stay in REGISTERS from step to step -- only step 0 reads a state from memory.  x_cl [T,B,H,W,32] (the inputs of all steps must exist:
 * layer-major execution of a window), v_in / z_in_cl [B,...] = the state before step 0 (NULL = zero), z_out_cl [T,B,H,W,32], v_out
 * [T,B,32,H,W] with save_all_v (training: the backward needs every step) or [B,32,H,W] = the last step only (inference).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_lif_conv_window_params {
  int32_t B, T, H, W, hard_reset, save_all_v;
  const uint16_t* x_cl;
  const float* v_in;
  const uint16_t* z_in_cl;
  const float* leak;             /* [32] raw parameters                                                                 */
  const float* thresh;           /* [32]                                                                                */
  const uint16_t* w_split;       /* ef_split_weights (no recurrent part) / ef_split_weights_head image                 */
  float* v_out;
  uint16_t* z_out_cl;
} ef_lif_conv_window_params;

int ef_lif_conv_fwd_window(const ef_lif_conv_window_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Backward of one cell-step (what autograd derives from the spans above; recurrences in SURVEY.md 8a).
 * Given g_out (dL/d out), g_v_out, g_z_out, g_aux_out (dL/d new state from step t+1, NULL = 0) and the tensors the
 * forward read/wrote, produces g_x, g_v_in, g_z_in, g_aux_in and ACCUMULATES (+=) weight / per-channel gradients.
 * scratch_gI: caller-provided [B,C,Ho,Wo] fp32 workspace (receives g_I = (1-leak) g_v).
 * scratch_gP: caller-provided [B,Ho,Wo] fp32 workspace, only needed for PLIF / XLIF cells when g_x is wanted (holds the
 * channel-summed gradient of the pre-synaptic trace input); may be NULL otherwise.
 * Inputs x / z_in may be given in either layout (fp32 NCHW or cl), gradients are fp32 NCHW.  Limits of this version: stride 1.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_lif_conv_bwd_params {
  ef_lif_conv_params f;         /* the forward call: its inputs and the v_out / aux_out it wrote (other outputs ignored) */
  const float* g_out;           /* [B,C,Ho,Wo] or NULL                                                                */
  const float* g_v_out;         /* or NULL                                                                            */
  const float* g_z_out;         /* or NULL                                                                            */
  const float* g_aux_out;       /* or NULL                                                                            */
  float* scratch_gI;            /* [B,C,Ho,Wo] workspace                                                              */
  float* scratch_gP;            /* [B,Ho,Wo] workspace (PLIF / XLIF with g_x) or NULL                                 */
  float* g_x;                   /* [B,Cin,H,W] or NULL (head layer)                                                   */
  float* g_v_in;                /* [B,C,Ho,Wo] or NULL                                                                */
  float* g_z_in;                /* [B,C,Ho,Wo] or NULL                                                                */
  float* g_aux_in;              /* or NULL                                                                            */
  float* g_w_ff;                /* [C,Cin,k,k] += , or NULL                                                           */
  float* g_w_rec;               /* [C,C,k,k] +=, or NULL                                                              */
  float* g_leak;                /* [C] += raw-parameter gradients (through sigmoid / clamp), each may be NULL         */
  float* g_thresh;
  float* g_leak_aux;
  float* g_add_pt;
  float* g_t0;
  float* g_t1;
  float* scratch_gI_up;         /* stride 2 only: [B,C,H,W] workspace (g_I zero-inserted to the input resolution)         */
  float* scratch_gP_up;         /* stride 2, PLIF / XLIF with g_x: [B,H,W] workspace                                      */
  int32_t reset_grad;           /* 1: the reset term is differentiable (cells built with detach=False, spiking_submodules.py:110-112): */
                                /* dL/dz_in also receives -leak v_in g_v (hard reset) / -thresh g_v (soft reset); needs g_z_in          */
  int32_t neuron_only;          /* 1: stop after the neuron backward -- scratch_gI (and scratch_gP, if given), g_v_in, the direct terms */
                                /* of g_z_in and the per-channel gradients are produced, the convolution gradients are the caller's     */
                                /* (ef_conv32_bwd_tc)                                                                                   */
} ef_lif_conv_bwd_params;

int ef_lif_conv_bwd(const ef_lif_conv_bwd_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Convolution gradients of a 32 -> 32, 3x3, stride-1 cell step on the tensor cores, for ANY neuron kind: what
 * ef_lif_conv_bwd computes after its neuron backward, given that backward's g_I (scratch_gI of a neuron_only call).  The
 * cell's inputs must be exact in bf16 (spikes, sums of spikes).  g_I is split into two bf16 terms (16 significant bits,
 * as in ef_lif_bwd_tc).  g_x is overwritten; g_z_in is ADDED to (it holds the direct terms of the neuron backward);
 * the weight gradients follow the wg_partial protocol of ef_lif_bwd_tc.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_conv32_bwd_tc_params {
  int32_t B, H, W, has_rec;
  const float* gI;               /* [B,32,H,W] dL/d(synaptic current)                                                   */
  const uint16_t* x_cl;          /* [B,H,W,32] bf16 input of the step                                                   */
  const uint16_t* z_in_cl;       /* [B,H,W,32] bf16 previous spikes, or NULL                                            */
  const uint16_t* w_bwd;         /* ef_split_weights_bwd image                                                          */
  uint16_t* gI_hi;               /* [B,H,W,32] workspace                                                                */
  uint16_t* gI_mid;              /* [B,H,W,32] workspace                                                                */
  float* g_x;                    /* [B,32,H,W]                                                                          */
  float* g_z_in;                 /* [B,32,H,W] += or NULL                                                               */
  float* g_z_tmp;                /* [B,32,H,W] workspace (with g_z_in)                                                  */
  float* wg_partial;             /* ef_lif_wgrad_partial_elems() floats (with g_w_ff / g_w_rec)                         */
  int32_t wg_flags;              /* EF_WG_*                                                                             */
  float* g_w_ff;                 /* [32,32,3,3] += or NULL                                                              */
  float* g_w_rec;                /* [32,32,3,3] += or NULL                                                              */
  const float* gP_sum;           /* [B,H,W] or NULL: PLIF / XLIF trace gradient (scratch_gP of the neuron backward):    */
  const float* x_f32;            /* [B,32,H,W] fp32 input (sign of x), needed with gP_sum; g_x += sign(x)/32 pool^T(gP) */
} ef_conv32_bwd_tc_params;
int ef_conv32_bwd_tc(const ef_conv32_bwd_tc_params* p, void* stream);
/* g fp32 NCHW [B,C,Hs,Ws] (C % 8 == 0) -> two bf16 channels-last terms hi / mid [B,H,W,C] (hi + mid = g to 16 significant bits), at the
 * same resolution or zero-inserted to the input resolution (H, W) of a stride-2 convolution (Hs = (H-1)/2 + 1): the two sources of a
 * data gradient computed as a plain convolution by ef_lif_conv_fwd_g (autograd of the cells with other channel counts than 32). */
int ef_split2_pack_cl(const float* src, uint16_t* hi, uint16_t* mid, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Hs, int32_t Ws, void* stream);
/* Tensor-core weight gradient of a 3x3 convolution for channel counts that are multiples of 32 (autograd of the U-Net cells):
 * g_w[co][ci_off + ci][dy][dx] += sum_{b,y,x} x_cl[b,y+dy-1,x+dx-1,ci] (g_hi + g_mid)[b,y,x,co].  x_cl [B,H,W,cin] bf16 (exact values:
 * spikes), g_hi / g_mid [B,H,W,cout] from ef_split2_pack_cl (a stride-2 convolution passes the zero-inserted form at its input resolution),
 * g_w [cout][cin_total][3][3] fp32, partial = ef_wgrad_tcg_partial_elems() floats of workspace.  Fixed summation order. */
int64_t ef_wgrad_tcg_partial_elems(int32_t B, int32_t H, int32_t W, int32_t cin, int32_t cout);
int ef_wgrad_tcg(const uint16_t* x_cl, const uint16_t* g_hi, const uint16_t* g_mid, int32_t B, int32_t H, int32_t W, int32_t cin, int32_t cout,
                 float* partial, float* g_w, int32_t cin_total, int32_t ci_off, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Backward of a 32 -> 32 LIF cell-step on the fast-path formats (same maths as ef_lif_conv_bwd; tensor-core data gradient).
 * x_cl / z_in_cl: channels-last bf16 spikes the forward consumed; v_in / v_out: fp32 NCHW membrane before / after the step.
 * g_out: dL/d(output spikes) from the layer above; g_v_out / g_z_out: dL/d(new state) from step t+1 (NULL = 0).
 * gI_hi / gI_mid: caller-provided bf16 [B,H,W,32] workspaces (receive g_I = (1-leak) g_v as two bf16 terms).
 * w_bwd: flipped / transposed weight image from ef_split_weights_bwd.  g_x (and g_z_in for a recurrent cell) are overwritten,
 * g_v_in is overwritten, weight / per-channel gradients are accumulated (+=).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_lif_bwd_tc_params {
  int32_t B, H, W, has_rec, hard_reset, surrogate;
  float act_width;
  const uint16_t* x_cl;          /* [B,H,W,32]                                                                         */
  const uint16_t* z_in_cl;       /* [B,H,W,32] or NULL (no previous state)                                             */
  const float* v_in;             /* [B,32,H,W] or NULL                                                                 */
  const float* v_out;            /* [B,32,H,W]                                                                         */
  const float* g_out;            /* [B,32,H,W] or NULL                                                                 */
  const float* g_v_out;          /* or NULL                                                                            */
  const float* g_z_out;          /* or NULL                                                                            */
  const float* leak;             /* [32] raw parameters                                                                */
  const float* thresh;           /* [32]                                                                               */
  const uint16_t* w_bwd;         /* ef_split_weights_bwd image                                                         */
  uint16_t* gI_hi;               /* [B,H,W,32] workspace                                                               */
  uint16_t* gI_mid;              /* [B,H,W,32] workspace                                                               */
  float* g_x;                    /* [B,32,H,W]                                                                         */
  float* g_v_in;                 /* [B,32,H,W] or NULL                                                                 */
  float* g_z_in;                 /* [B,32,H,W] or NULL (recurrent cells)                                               */
  float* g_w_ff;                 /* [32,32,3,3] += or NULL                                                             */
  float* g_w_rec;                /* [32,32,3,3] += or NULL                                                             */
  float* g_leak;                 /* [32] += or NULL                                                                    */
  float* g_thresh;               /* [32] += or NULL                                                                    */
  float* wg_partial;             /* ef_lif_wgrad_partial_elems() floats or NULL: selects the tensor-core weight gradient */
  int32_t wg_flags;              /* EF_WG_* below                                                                      */
  /* Head-layer mode (x_f32 != NULL): the cell's input is the fp32 NCHW network input with Cin <= 8 channels (event counts /  */
  /* voxel bins); no data gradient is produced, x_cl / w_bwd / gI_hi / gI_mid / g_x / wg_partial are ignored.                */
  int32_t Cin;                   /* input channels of the head layer                                                   */
  const float* x_f32;            /* [B,Cin,H,W] or NULL (= 32 -> 32 cell)                                              */
  float* gI_f32;                 /* [B,32,H,W] workspace (head mode)                                                   */
} ef_lif_bwd_tc_params;

/* Tensor-core weight gradient (wg_partial != NULL): every CTA keeps its share of
 * sum_{b,y,x} x[b,y+dy-1,x+dx-1,ci] g_I[b,y,x,co] in tensor memory and writes it to its own slice of wg_partial; nothing is
 * added to g_w_ff / g_w_rec until a call carries EF_WG_FINALIZE, which reduces the slices in a fixed order (bit-reproducible)
 * and adds the result to g_w_*.  A BPTT sweep passes the same wg_partial for every step of one cell: first call without
 * flags, later calls with EF_WG_ACCUMULATE, the last one with EF_WG_FINALIZE as well.
 * wg_partial == NULL: CUDA-core kernel, g_w_* += at once. */
#define EF_WG_ACCUMULATE 1       /* wg_partial already holds the sums of earlier calls: add to them                      */
#define EF_WG_FINALIZE 2         /* after this call: g_w_ff / g_w_rec += reduced wg_partial                              */
int64_t ef_lif_wgrad_partial_elems(int32_t B, int32_t H, int32_t W, int32_t has_rec);

int ef_lif_bwd_tc(const ef_lif_bwd_tc_params* p, void* stream);
int64_t ef_split_weights_bwd_elems(int32_t has_rec);
/* Backward of a FEED-FORWARD 32 -> 32 LIF cell over a whole BPTT window of T steps (same maths as T calls of ef_lif_bwd_tc walking
 * t = T-1 ... 0 from a zero incoming state gradient).  Three launches for the whole window: (1) the neuron backward with the time
 * loop INSIDE the kernel -- dL/dv stays in registers from step to step, every membrane tensor is read once; (2) the tensor-core data
 * gradient and (3) the tensor-core weight gradient over T*B images at once.  Tensors hold the window step-major and dense:
 * x_cl / z_cl / gI_* [T,B,H,W,32], v / g_out / g_x [T,B,32,H,W]; the state before step 0 comes separately (NULL = zero state).
 * Head mode (x_f32 != NULL): the Cin <= 8 input cell, x_f32 [T,B,Cin,H,W], gI_f32 [T,B,32,H,W] workspace, no data gradient.
 * Split-input head (x_f32 == NULL, 0 < Cin < 32): x_cl [T,B,H,W,32] from ef_pack_split_cl, weight gradient on the tensor cores. */
typedef struct ef_lif_bwd_window_params {
  int32_t B, T, H, W, hard_reset, surrogate;
  float act_width;
  const uint16_t* x_cl;          /* input spikes of the cell at steps 0..T-1                                            */
  const uint16_t* z_cl;          /* spikes the cell emitted at steps 0..T-1 (step t reads z_cl[t-1] as previous spikes) */
  const uint16_t* z_prev_cl;     /* [B,H,W,32] spikes before step 0, or NULL                                            */
  const float* v;                /* membrane potential after steps 0..T-1                                               */
  const float* v_prev;           /* [B,32,H,W] before step 0, or NULL                                                   */
  const float* g_out;            /* dL/d(output spikes) of steps 0..T-1                                                 */
  const float* leak;             /* [32] raw parameters                                                                 */
  const float* thresh;           /* [32]                                                                                */
  const uint16_t* w_bwd;         /* ef_split_weights_bwd image                                                          */
  uint16_t* gI_hi;               /* [T,B,H,W,32] workspace                                                              */
  uint16_t* gI_mid;              /* [T,B,H,W,32] workspace                                                              */
  float* g_x;                    /* [T,B,32,H,W] overwritten, or NULL                                                   */
  float* g_v_prev;               /* [B,32,H,W] dL/d(membrane before step 0), overwritten, or NULL                       */
  float* g_w_ff;                 /* [32,32,3,3] ([32,Cin,3,3] in head mode) += or NULL                                  */
  float* g_leak;                 /* [32] += or NULL                                                                     */
  float* g_thresh;               /* [32] += or NULL                                                                     */
  float* wg_partial;             /* ef_lif_wgrad_partial_elems(T*B, H, W, 0) floats (needed with g_w_ff, not in head mode) */
  int32_t Cin;                   /* head modes: input channels of the head layer (0 = a 32 -> 32 cell)                  */
  const float* x_f32;            /* head mode: [T,B,Cin,H,W]                                                            */
  float* gI_f32;                 /* head mode: [T,B,32,H,W] workspace                                                   */
} ef_lif_bwd_window_params;

int ef_lif_bwd_window(const ef_lif_bwd_window_params* p, void* stream);
/* The weight-gradient stage of ef_lif_bwd_tc alone, for B images whose g_I = gI_hi + gI_mid exists already (a recurrent cell runs
 * pointwise + data gradient step by step, then this once over the window).  wg_partial / wg_flags as for ef_lif_bwd_tc. */
int ef_lif_wgrad_tc(const uint16_t* x_cl, const uint16_t* z_in_cl, const uint16_t* gI_hi, const uint16_t* gI_mid, int32_t has_rec, int32_t B,
                    int32_t H, int32_t W, float* wg_partial, int32_t wg_flags, float* g_w_ff, float* g_w_rec, void* stream);

int ef_split_weights_bwd(const float* w_ff, const float* w_rec, uint16_t* out, void* stream);

/* Split fp32 conv weights [C,Cin,3,3] (+ optional recurrent [C,C,3,3]) into three bf16 terms hi+mid+lo == w exactly,
 * laid out as the tcgen05 B operand.  out: uint16 [ef_split_weights_elems(Cin, C, has_rec)].  (no reference analogue) */
int64_t ef_split_weights_elems(int32_t Cin, int32_t C, int32_t has_rec);
int ef_split_weights(const float* w_ff, const float* w_rec, int32_t Cin, int32_t C, uint16_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused conv3x3 + LIF step on the tensor cores for GENERAL channel counts and several concatenated input sources: the cells of the
 * EV-FlowNet family (models/unet.py:418-465 with spiking_submodules.py:878-1013; ConvLIF / ConvLIFRecurrent at 64..512 channels,
 * residual blocks, decoders on cat[prediction, x, skip]).  Stride 1, kernel 3, LIF, C a multiple of 32, W a multiple of 4.
 *   I      = sum over sources s of conv(src[s], W_s)          (the recurrent convolution of a ConvLIFRecurrent is one more source:
 *            src = z_in_cl, W = rec.weight; the channel concat of a decoder is never materialised)
 *   v_out  = LIF update of (v_in, z_in_cl) with I; z_out = (v_out - thresh > 0); out = z_out + residual (optional)
 * Sources are cl bf16 tensors [B,H,W,src_c[s]] with src_c a multiple of 32 (unused channels zero) whose values are EXACT in bf16
 * (spikes, residual sums, their bilinear x2 upsampling, event counts); a fractional fp32 source is passed as its exact hi/mid/lo
 * split (ef_pack_split_cl, src_c = 32).  The weights come as one image built by ef_split_weights_g from the same source list
 * (three exact bf16 terms per weight, tcgen05 B-operand layout, blocked [C/32][K/32]); products are fp32-exact.
 * STRIDE 2 (the encoder cells, unet.py:83-89 / spiking_submodules.py:878-930) runs as a stride-1 cell at the OUTPUT resolution over the
 * space-to-depth form of the input, s2d[b, Y, X, (py*2 + px)*n + c] = in[b, 2Y + py, 2X + px, c] (ef_space_to_depth_cl, or
 * ef_pack_split_cl with s2d = 1 for fp32 inputs): ef_wsrc.s2d = 1 remaps the 3x3 stride-2 weights onto the four taps (ky, kx) in {0,1}^2
 * of that input, ef_lif_conv_g_params.s2d = 1 makes the kernel skip the five all-zero taps.  B, H, W are then the output's.
 * ------------------------------------------------------------------------------------------------------------------ */
#define EF_TCG_MAX_SRC 4
typedef struct ef_wsrc {
  const float* w;                /* fp32 conv weight [C, c_total, 3, 3] this source's channels are a slice of            */
  int32_t c_total;               /* input channels of that weight tensor                                              */
  int32_t ch0, n;                /* first channel and channel count of the slice                                      */
  int32_t split;                 /* 1: the source tensor is an ef_pack_split_cl split (n <= EF_HEAD_MAX_CIN)            */
  int32_t s2d;                   /* 1: the source tensor is the SPACE-TO-DEPTH form of the stride-2 cell's input (below)  */
} ef_wsrc;
/* out: uint16[ef_split_weights_g_elems(C, n_src, srcs)] */
int64_t ef_split_weights_g_elems(int32_t C, int32_t n_src, const ef_wsrc* srcs);
int ef_split_weights_g(const ef_wsrc* srcs, int32_t n_src, int32_t C, uint16_t* out, void* stream);

typedef struct ef_lif_conv_g_params {
  int32_t B, H, W, C;            /* outputs: membrane [B,C,H,W] fp32, spikes [B,H,W,C] cl                              */
  int32_t n_src;                 /* 1..EF_TCG_MAX_SRC input sources, in the order given to ef_split_weights_g          */
  int32_t hard_reset;
  int32_t s2d;                   /* 1: stride-2 cell on a space-to-depth source (a single source)                      */
  const uint16_t* src[EF_TCG_MAX_SRC];  /* [B,H,W,src_c[s]] cl                                                         */
  int32_t src_c[EF_TCG_MAX_SRC]; /* channels of each source tensor, multiples of 32                                   */
  const float* v_in;             /* [B,C,H,W] or NULL (zero state; then z_in_cl is NULL too)                           */
  const uint16_t* z_in_cl;       /* [B,H,W,C] previous spikes (reset term) or NULL                                     */
  const uint16_t* residual_cl;   /* [B,H,W,C] or NULL                                                                  */
  const float* leak;             /* [C] raw parameters                                                                 */
  const float* thresh;           /* [C]                                                                                */
  const uint16_t* w_image;       /* ef_split_weights_g image                                                           */
  float* v_out;                  /* [B,C,H,W]                                                                          */
  uint16_t* z_out_cl;            /* [B,H,W,C]                                                                          */
  uint16_t* out_cl;              /* [B,H,W,C] = z_out + residual (given iff residual_cl is)                            */
} ef_lif_conv_g_params;

int ef_lif_conv_fwd_g(const ef_lif_conv_g_params* p, void* stream);

/* Debug aid: subsequent tensor-core launches write a per-CTA clock64 timeline into buf (device int64 [n_ctas][32][8]);
 * NULL switches it off (default).  Not part of the reference's interface. */
int ef_debug_tc_trace(long long* buf);
/* Debug aid: ablation mask for the tensor-core kernel (results become wrong; timing experiments only).  0 = off. */
int ef_debug_tc_skip(int mask);
/* Epilogue shape of the tensor-core forward kernel: 16 channels per thread (8 epilogue warps, default) or 8 (16 warps). */
int ef_debug_tc_cpt(int cpt);
/* Programmatic dependent launch of the forward kernels of a model step (head, fused cells, prediction): on by default, EF_PDL=0
 * in the environment or ef_debug_pdl(0) falls back to plain stream-ordered launches. */
int ef_debug_pdl(int on);

/* Resampling glue of the U-Net family.  src [n_planes,H,W] fp32 (n_planes = B*C), dst [n_planes,2H,2W]:
 * F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False) of models/spiking_submodules.py:1010 /
 * submodules.py UpsampleConvLayer; and the nearest-neighbour upsampling by integer factors of the multi-resolution flow
 * maps (models/model.py:528-539), dst [n_planes,H*fy,W*fx]. */
int ef_upsample_bilinear2x(const float* src, float* dst, int64_t n_planes, int32_t H, int32_t W, void* stream);
int ef_upsample_nearest(const float* src, float* dst, int64_t n_planes, int32_t H, int32_t W, int32_t fy, int32_t fx, void* stream);
/* The bilinear x2 upsampling on the internal spike format: src [B,H,W,C] cl bf16 -> dst [B,2H,2W,C]; exact for spikes and residual sums
 * (results are multiples of 1/16).  C a multiple of 8. */
int ef_upsample_bilinear2x_cl(const uint16_t* src, uint16_t* dst, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* Their adjoints (what autograd derives for the two F.interpolate calls): g_dst has the upsampled shape, g_src [n_planes,H,W]. */
int ef_upsample_bilinear2x_bwd(const float* g_dst, float* g_src, int64_t n_planes, int32_t H, int32_t W, void* stream);
int ef_upsample_nearest_bwd(const float* g_dst, float* g_src, int64_t n_planes, int32_t H, int32_t W, int32_t fy, int32_t fx, void* stream);

/* Head layer (Cin <= EF_HEAD_MAX_CIN fractional fp32 inputs: voxel grids) on the tensor-core cell kernel with fp32-exact products:
 * ef_pack_split_cl turns x [B,Cin,H,W] into a cl tensor [B,H,W,32] whose channel s*EF_HEAD_SLOT(Cin) + c holds term s (hi, mid, lo) of
 * the exact three-way bf16 split of x[.,c] (hi + mid + lo == x), ef_split_weights_head builds the matching weight image (w[.,c] repeated
 * in the three slots of c; out: uint16[ef_split_weights_elems(32, 32, 0)]).  ef_lif_conv_fwd is then called with Cin = 32, x_cl = the packed
 * tensor, w_split = that image; ef_lif_bwd_window with 0 < Cin < 32 and the packed x_cl computes the head's weight gradient on the
 * tensor cores (no data gradient). */
#define EF_HEAD_MAX_CIN 10
#define EF_HEAD_SLOT(cin) ((cin) <= 8 ? 8 : 10)
int ef_pack_split_cl(const float* src, uint16_t* dst, int32_t B, int32_t Cin, int32_t H, int32_t W, void* stream);
/* The same split of the space-to-depth form of x (stride-2 first encoder): dst [B,H/2,W/2,32], virtual channel (py*2 + px)*Cin + c,
 * 4*Cin <= EF_HEAD_MAX_CIN, H and W even. */
int ef_pack_split_s2d_cl(const float* src, uint16_t* dst, int32_t B, int32_t Cin, int32_t H, int32_t W, void* stream);
/* Space-to-depth of a cl tensor: src [B,H,W,C] -> dst [B,H/2,W/2,4C], dst[b,Y,X,(py*2+px)*C + c] = src[b,2Y+py,2X+px,c]; C % 8 == 0. */
int ef_space_to_depth_cl(const uint16_t* src, uint16_t* dst, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
int ef_split_weights_head(const float* w_ff, int32_t Cin, uint16_t* out, void* stream);

/* fp32 NCHW <-> cl bf16 layout conversion at the API boundary (model.states getter/setter, first input). */
int ef_pack_cl(const float* src, uint16_t* dst, int32_t B, int32_t C, int32_t H, int32_t W, void* stream);
int ef_unpack_cl(const uint16_t* src, float* dst, int32_t B, int32_t C, int32_t H, int32_t W, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Prediction head: flow = tanh(conv1x1(x, w) + b).  Replaces ConvLayer.forward, models/submodules.py:52-61 as built
 * at models/model.py:197-199 (32 -> 2 channels).  x may be fp32 NCHW or cl.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_pred_params {
  int32_t B, Cin, Cout, H, W;
  const float* x;               /* [B,Cin,H,W] or NULL                                                                */
  const uint16_t* x_cl;         /* or NULL                                                                            */
  const float* w;               /* [Cout,Cin]                                                                         */
  const float* b;               /* [Cout]                                                                             */
  float* y;                     /* [B,Cout,H,W] = tanh(...)                                                           */
  /* backward only */
  const float* g_y;             /* [B,Cout,H,W]                                                                       */
  float* g_x;                   /* [B,Cin,H,W]                                                                        */
  float* g_w;                   /* [Cout,Cin] +=                                                                      */
  float* g_b;                   /* [Cout] +=                                                                          */
} ef_pred_params;

int ef_pred_fwd(const ef_pred_params* p, void* stream);
int ef_pred_bwd(const ef_pred_params* p, void* stream); /* needs y (forward output), x, g_y */

/* ------------------------------------------------------------------------------------------------------------------
 * ANN cells: 3x3 convolution (pad 1, stride 1) + bias + residual + activation, optionally over cat([x1, x2 * x2_scale])
 * and with a gated blend out = h*(1-u) + act(...)*u.  Replaces ConvLayer.forward / ConvLayer_.forward
 * (models/submodules.py:52-83) and, in two calls, ConvGRU.forward (models/submodules.py:400-418):
 *   call 1: x1 = input, x2 = h, w = [update_gate.weight; reset_gate.weight], act = sigmoid      -> ur [B, 2C, H, W]
 *   call 2: x1 = input, x2 = h, x2_scale = reset, w = out_gate.weight, act = tanh, blend_h = h, blend_u = update -> new h
 * Tensors are fp32 NCHW planes; the *_bstride fields are the element distance between samples (lets a channel slice of a
 * larger tensor be passed).  act: 0 none, 1 relu, 2 sigmoid, 3 tanh.  Forward only in this version.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_conv_ann_params {
  int32_t B, C1, C2, Cout, H, W, act;
  const float* x1;               /* [B,C1,H,W]                                                                         */
  const float* x2;               /* [B,C2,H,W] or NULL (C2 = 0)                                                        */
  const float* x2_scale;         /* [B,C2,H,W] elementwise multiplier of x2 or NULL                                    */
  int64_t x1_bstride, x2_bstride, x2_scale_bstride;
  const float* w;                /* [Cout, C1 + C2, 3, 3]                                                              */
  const float* bias;             /* [Cout] or NULL                                                                     */
  const float* residual;         /* [B,Cout,H,W] added before the activation, or NULL                                  */
  const float* blend_h;          /* [B,Cout,H,W] or NULL                                                               */
  const float* blend_u;          /* [B,Cout,H,W] or NULL                                                               */
  int64_t blend_h_bstride, blend_u_bstride;
  float* out;                    /* [B,Cout,Ho,Wo]                                                                     */
  int32_t stride;                /* 1 (0 = 1) or 2: output Ho = (H-1)/stride + 1; residual / blend / out at the output size */
  float* act_out;                /* [B,Cout,Ho,Wo] or NULL: the activation BEFORE the blend (kept for the backward)      */
  int32_t inference;             /* != 0: no backward pass follows -- small launches may split the input-channel sum over thread
                                  * groups (latency shape; sums differ in the last bit from the sequential channel order)     */
} ef_conv_ann_params;

int ef_conv_ann_fwd(const ef_conv_ann_params* p, void* stream);

/* Backward helpers of the ANN cells: everything autograd derives AROUND the convolution, one launch each, so that the backward of a
 * ConvLayer / ConvGRU step is kernels only (models/submodules.py:52-61, 400-418).
 *  ef_ann_gate_bwd   y = h (1 - u) + o u with o = act(pre) (or y = o without a blend): g_pre = g_y [u] act'(o), g_h = g_y (1 - u),
 *                    g_u = g_y (o - h); g_bias[c] += sum g_pre.  act_out = o as the forward stored it (ef_conv_ann_params.act_out, or
 *                    its `out` when there is no blend).
 *  ef_ann_cat_scale  out [B,C1+C2,H,W] = cat([x1, x2 * scale]): the convolution's input as one tensor (ef_conv3x3_bwd reads it for the
 *                    weight gradient); scale may be NULL.
 *  ef_ann_scale_bwd  through the gate product: g_x2 = g_xcat[:, C1:] * scale, g_scale = g_xcat[:, C1:] * x2 (either output may be NULL). */
typedef struct ef_ann_gate_bwd_params {
  int32_t B, C, H, W, act;
  const float* g_y;              /* [B,C,H,W]                                                                          */
  const float* act_out;          /* [B,C,H,W]                                                                          */
  const float* blend_h;          /* [B,C,H,W] (batch stride blend_h_bstride) or NULL                                   */
  const float* blend_u;
  int64_t blend_h_bstride, blend_u_bstride;
  float* g_pre;                  /* [B,C,H,W]                                                                          */
  float* g_h;                    /* [B,C,H,W] or NULL                                                                  */
  float* g_u;                    /* [B,C,H,W] or NULL                                                                  */
  float* g_bias;                 /* [C] += or NULL                                                                     */
} ef_ann_gate_bwd_params;
int ef_ann_gate_bwd(const ef_ann_gate_bwd_params* p, void* stream);
int ef_ann_cat_scale(const float* x1, const float* x2, const float* scale, float* out, int32_t B, int32_t C1, int32_t C2, int32_t H, int32_t W,
                     int64_t x1_bstride, int64_t x2_bstride, int64_t scale_bstride, void* stream);
int ef_ann_scale_bwd(const float* g_xcat, const float* x2, const float* scale, float* g_x2, float* g_scale, int32_t B, int32_t C1, int32_t C2,
                     int32_t H, int32_t W, int64_t x2_bstride, int64_t scale_bstride, void* stream);

/* Gradients of the 3x3 stride-1 convolution inside the ANN cells (what autograd derives for the nn.Conv2d of
 * models/submodules.py:22,159,256,281,386-388): g_pre = dL/d(conv output) [B,C,H,W]; g_x [B,Cin,H,W] is overwritten (NULL = skip),
 * g_w [C,Cin,3,3] is accumulated (NULL = skip; needs x [B,Cin,H,W]). */
int ef_conv3x3_bwd(const float* g_pre, const float* x, const float* w, float* g_x, float* g_w, int32_t B, int32_t Cin, int32_t C,
                   int32_t H, int32_t W, void* stream);
/* The same for a stride-2 convolution: g_pre [B,C,Ho,Wo] is zero-inserted into scratch_up [B,C,H,W] (caller-provided workspace) and the
 * stride-1 gradient kernels run on it (what the transposed convolution is); stride 1 ignores scratch_up. */
int ef_conv3x3_bwd_s(const float* g_pre, const float* x, const float* w, float* g_x, float* g_w, float* scratch_up, int32_t B, int32_t Cin,
                     int32_t C, int32_t H, int32_t W, int32_t stride, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Contrast-maximisation event-warping loss over one training window.
 * Replaces EventWarping.forward (loss/flow.py:176-301) with utils/iwe.py:4-92 (purge_unfeasible, get_interpolation,
 * interpolate) and the per-event flow gather of event_flow_association (loss/flow.py:65-79), and -- ef_iwe_loss_bwd --
 * the autograd graph of all of it (analytic gradient).
 *
 * Window in "map form": events [B,Ntot,4] (ts,y,x,p; ts already offset by pass index, loss/flow.py:90),
 * pol_mask [B,Ntot,2], flow maps [S,B,T,2,H,W] (scale, sample, pass, (x,y), H, W), event_mask [B,T,H,W].
 * Events of pass t occupy columns [t*n_per_pass, (t+1)*n_per_pass) (or [pass_offsets[t], pass_offsets[t+1]) when passes
 * have different lengths).  With overwrite_intermediate, T_maps == 1 and every
 * event reads map 0 (loss/flow.py:118-146).
 *
 * workspace: float[ef_iwe_loss_workspace_elems(S,B,H,W)]: grid-barrier counters, the 8 IWE images per (scale, sample), the
 * per-(scale,sample,direction) sums and pixel counts (kept for the backward), per-CTA partial sums of the smoothness term.
 * CONTRACT: the first 16 words of a workspace buffer must be ZERO when the buffer is first used (cudaMemset the whole buffer
 * once, or torch.zeros); every forward / backward call leaves them zero again, so a buffer can be reused call after call
 * without being cleared.  Everything else in the buffer is initialised by the forward call itself.
 *
 * ONE kernel launch per direction.  Forward: zero accumulators + smoothness -> (grid barrier) -> per-event warp and bilinear
 * scatter with 8/16-byte vector atomics -> (grid barrier) -> contrast reduction -> last CTA writes the scalar.  Backward:
 * smoothness gradient -> (grid barrier) -> per-event analytic gradient, adjoint images evaluated on the fly per corner.
 * The launches are cooperative (all CTAs co-resident); limits: T <= EF_IWE_MAX_PASSES, S <= EF_IWE_MAX_SCALES, H*W even.
 * ------------------------------------------------------------------------------------------------------------------ */
#define EF_IWE_MAX_PASSES 32
#define EF_IWE_MAX_SCALES 4

typedef struct ef_iwe_loss_params {
  int32_t S, B, T, T_maps, H, W; /* T = number of passes (max_ts); T_maps = T, or 1 with overwrite_intermediate       */
  int32_t n_total, n_per_pass;   /* Ntot events per sample; events per pass                                           */
  float flow_scaling;            /* loss/flow.py:40 (default max(H,W))                                                */
  float weight;                  /* flow_regul_weight                                                                 */
  int32_t loss_scaling;          /* divide by #pixels with events (loss/flow.py:221-225)                              */
  int32_t smoothing_mask;        /* mask smoothness terms with the event masks (:184-190, :280-286)                   */
  int32_t overwrite_intermediate;
  const float* events;           /* [B,Ntot,4]                                                                        */
  const float* pol_mask;         /* [B,Ntot,2]                                                                        */
  const float* flow_maps;        /* [S,B,T_maps,2,H,W]                                                                */
  const float* event_mask;       /* [B,T_maps,H,W] (only read when smoothing_mask)                                    */
  const int32_t* pass_offsets;   /* HOST int32[T+1], first event column of each pass (ragged passes), or NULL =        */
                                 /* uniform passes of n_per_pass events.  Read during the call only.                  */
  float* workspace;
  float* loss;                   /* [1]                                                                               */
  /* backward only */
  const float* g_loss;           /* [1] upstream gradient (device)                                                    */
  float* g_flow_maps;            /* [S,B,T_maps,2,H,W], overwritten                                                   */
} ef_iwe_loss_params;

int64_t ef_iwe_loss_workspace_elems(int32_t S, int32_t B, int32_t H, int32_t W);
int ef_iwe_loss_fwd(const ef_iwe_loss_params* p, void* stream);
int ef_iwe_loss_bwd(const ef_iwe_loss_params* p, void* stream);

/* The same loss on a window in "pass form": what EventWarping.event_flow_association (loss/flow.py:56-116) receives pass
 * by pass is handed over as pointer tables -- nothing is concatenated or copied (the reference grows its lists with
 * torch.cat every pass: O(T^2) copying).  Pass t: events[t] [B,n_pass[t],4] (ts already offset by t),
 * pol_mask[t] [B,n_pass[t],2]; flow[s*T_maps + m] [B,2,H,W] = scale s, map m; event_mask[m] [B,1,H,W]; g_flow likewise.  All
 * dense.  With overwrite_intermediate T_maps == 1: flow[s] is the final map, event_mask[0] the union mask (:118-146). */
typedef struct ef_iwe_loss_pass_params {
  int32_t S, B, T, T_maps, H, W;
  float flow_scaling, weight;
  int32_t loss_scaling, smoothing_mask, overwrite_intermediate;
  int32_t n_pass[EF_IWE_MAX_PASSES];
  const float* events[EF_IWE_MAX_PASSES];
  const float* pol_mask[EF_IWE_MAX_PASSES];
  const float* flow[EF_IWE_MAX_SCALES * EF_IWE_MAX_PASSES];
  const float* event_mask[EF_IWE_MAX_PASSES];
  float* workspace;              /* as above                                                                          */
  float* loss;                   /* [1]                                                                               */
  const float* g_loss;           /* backward only: [1]                                                                */
  float* g_flow[EF_IWE_MAX_SCALES * EF_IWE_MAX_PASSES]; /* backward only: [B,2,H,W] each, overwritten                 */
} ef_iwe_loss_pass_params;

int ef_iwe_loss_fwd_passes(const ef_iwe_loss_pass_params* p, void* stream);
int ef_iwe_loss_bwd_passes(const ef_iwe_loss_pass_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Per-polarity image of warped events.  Replaces compute_pol_iwe / deblur_events (utils/iwe.py:95-153) and the
 * IWE half of BaseValidationLoss.compute_window_iwe (loss/flow.py:452-465).
 * iwe: [B,2,H,W], overwritten.  round_idx: nearest pixel, half-to-even (torch.round) instead of bilinear.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_iwe_image_params {
  int32_t B, N, H, W;
  int32_t round_idx;
  float tref, flow_scaling;
  const float* events;           /* [B,N,4]                                                                           */
  const float* pol_mask;         /* [B,N,2]                                                                           */
  const float* flow;             /* [B,2,H,W] flow map, gathered per event; or NULL if event_flow given               */
  const float* event_flow;       /* [B,N,2] (fy,fx) per event or NULL                                                 */
  float* iwe;                    /* [B,2,H,W]                                                                         */
} ef_iwe_image_params;

int ef_iwe_image(const ef_iwe_image_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Stand-alone IWE primitives, for callers that use the reference's helpers directly (utils/iwe.py:4-92).  Inside the loss /
 * metric / image entry points above they are fused; these reproduce the reference's intermediate tensors one launch each.
 * Forward only (the differentiable path is ef_iwe_loss_fwd / ef_iwe_loss_bwd).
 *   ef_iwe_purge_unfeasible  purge_unfeasible (:4-17):  x [n,2] (y,x) -> x_out = x * mask, mask [n,1]
 *   ef_iwe_get_interpolation get_interpolation (:20-74): events [B,N,4], flow [B,N,2] (fy,fx) -> idx, weights [B,4N,1]
 *                            (corner order TL,TR,BL,BR along N; [B,N,1] with round_idx); out-of-bounds: weight 0, index 0
 *   ef_iwe_interpolate       interpolate (:77-92): scatter-add of weights (* polarity_mask [B,M,1] or NULL) at idx -> iwe
 *                            [B,1,H,W], overwritten
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_iwe_interp_params {
  int32_t B, N, H, W, round_idx;
  float tref, flow_scaling;
  const float* events;           /* [B,N,4]                                                                           */
  const float* flow;             /* [B,N,2] (fy,fx) per event                                                         */
  float* idx;                    /* [B,4N,1] ([B,N,1] with round_idx) flat pixel index, as fp32 like the reference    */
  float* weights;                /* same shape                                                                        */
} ef_iwe_interp_params;

int ef_iwe_purge_unfeasible(const float* x, int64_t n, int32_t H, int32_t W, float* x_out, float* mask, void* stream);
int ef_iwe_get_interpolation(const ef_iwe_interp_params* p, void* stream);
int ef_iwe_interpolate(const float* idx, const float* weights, const float* polarity_mask, int32_t B, int32_t M, int32_t H, int32_t W,
                       float* iwe, void* stream);

/* Stand-alone spike functions (models/spiking_util.py:13-109): z = (x - thresh > 0) as fp32; backward g_x = g * sg(x - thresh)
 * with sg the surrogate EF_ARCTAN.. of the given width.  thresh_mode 0: one scalar; 1: per channel (thresh [C], x [B,C,hw]);
 * 2: thresh has x's shape.  n = number of elements of x.  Inside ef_lif_conv_fwd / _bwd these are fused. */
int ef_spike_fwd(const float* x, const float* thresh, int32_t thresh_mode, int32_t C, int64_t hw, int64_t n, float* z, void* stream);
int ef_spike_bwd(const float* x, const float* thresh, int32_t thresh_mode, int32_t C, int64_t hw, int64_t n, const float* g,
                 int32_t surrogate, float width, float* g_x, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Validation metrics on rounded-index IWEs.  Replaces FWL.forward and RSAT.forward (loss/flow.py:468-579) with
 * get_interpolation(round_idx=True) / interpolate (utils/iwe.py:20-92) and spatial_variance (loss/flow.py:13-23).
 * Per-event flow comes from flow maps as in ef_iwe_loss_fwd (pass layout identical); with T_maps == 1 every event reads map 0.
 * out: [B][4] = FWL, RSAT, and the two RSAT terms (warped, unwarped) for inspection.  workspace:
 * ef_iwe_metrics_workspace_elems(B,H,W) floats.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_iwe_metrics_params {
  int32_t B, T, T_maps, H, W, n_total, n_per_pass;
  float flow_scaling;
  const float* events;           /* [B,Ntot,4] (ts offset by pass index)                                               */
  const float* pol_mask;         /* [B,Ntot,2]                                                                         */
  const float* flow_maps;        /* [B,T_maps,2,H,W]                                                                   */
  const int32_t* pass_offsets;   /* int32[T+1] or NULL                                                                 */
  float* workspace;
  float* out;                    /* [B,4]                                                                              */
} ef_iwe_metrics_params;

int64_t ef_iwe_metrics_workspace_elems(int32_t B, int32_t H, int32_t W);
int ef_iwe_metrics(const ef_iwe_metrics_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Average endpoint error.  Replaces AEE.forward (loss/flow.py:597-628): flow * flow_scaling * dt_gt/dt_input against the
 * ground truth, on pixels with events and valid ground truth; outliers: error > 3 px and > 5 % of the flow magnitude.
 * out: [2][B] = AEE per sample, then percent_AEE per sample (batch-wide outlier count / per-sample valid pixels, as the
 * reference computes it, loss/flow.py:625-626).  workspace: 2*B + 1 floats.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_aee_params {
  int32_t B, H, W;
  float flow_scaling;
  const float* flow;             /* [B,2,H,W] network output (last flow map)                                           */
  const float* gtflow;           /* [B,2,H,W]                                                                          */
  const float* event_mask;       /* [B,H,W] mask of the last pass                                                      */
  const float* dt_ratio;         /* [B] dt_gt / dt_input (device)                                                      */
  float* workspace;
  float* out;                    /* [2,B]                                                                              */
} ef_aee_params;

int ef_aee(const ef_aee_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Event encodings for one batch.  Replaces dataloader/encodings.py:30-85 (events_to_image/voxel/channels) and
 * dataloader/base.py:148-222 (cnt, mask, voxel, polarity mask).  events: [B,N,4] (ts,y,x,p).  Outputs overwritten;
 * any may be NULL.  Counts are integer-valued fp32 and bit-exact.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ef_encode_params {
  int32_t B, N, H, W, num_bins, round_ts;
  const float* events;
  float* cnt;                    /* [B,2,H,W]                                                                         */
  float* voxel;                  /* [B,num_bins,H,W]                                                                  */
  float* mask;                   /* [B,1,H,W]                                                                         */
  float* pol_mask;               /* [B,N,2]                                                                           */
} ef_encode_params;

int ef_encode_events(const ef_encode_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused gradient-norm clip + Adam on one flat fp32 parameter buffer.  Replaces clip_grad_norm_ + Adam.step
 * (train_flow.py:157-163).  Two launches: ef_grad_sqnorm accumulates sum(g^2) into sqnorm[1] (caller zeroes), then
 * ef_clip_adam applies g *= min(1, max_norm/(sqrt(sqnorm)+1e-6)) and the Adam update (bias-corrected, eps outside sqrt).
 * ------------------------------------------------------------------------------------------------------------------ */
int ef_grad_sqnorm(const float* g, int64_t n, float* sqnorm, void* stream);

/* The data-parallel optimiser step as ONE launch per rank (csrc/dp_step.cu): one-shot all-reduce(SUM) of the flat gradient over NVLink
 * peer memory -> global-norm clip -> Adam -> zero the own gradient; every rank reduces all buffers in rank order (identical replicas).
 * grads[r] / signals[r]: the flat fp32 gradient [n] and the uint32 signal words [16] of rank r as mapped into THIS process (CUDA IPC;
 * [rank] = the own buffers).  signal words must be zero before the first step; epoch = 1, 2, 3, ... the same on all ranks, one launch
 * per epoch (epoch_launches = launches so far incl. this one: the grid barrier counts arrivals cumulatively).  scratch:
 * float[ef_dp_step_grid(n)]; status: uint32, 0 = fine, 1 / 2 / 3 = a wait of phase A / B / C timed out (graceful = 1: the kernel then
 * returns without updating; graceful = 0: it traps).  bc1 / bc2_sqrt are filled in by the call. */
#define EF_DP_MAX_RANKS 8
typedef struct ef_dp_step_params {
  int32_t world, rank, n, step;
  float clip, lr, beta1, beta2, eps;
  float bc1, bc2_sqrt;           /* set by ef_dp_step                                                                   */
  uint32_t epoch, epoch_launches;
  int32_t graceful, grid_expected;
  const float* grads[EF_DP_MAX_RANKS];
  uint32_t* signals[EF_DP_MAX_RANKS];
  float* param;
  float* m;
  float* v;
  float* sqnorm;                 /* [1] receives sum g^2 of the reduced gradient                                        */
  float* scratch;
  uint32_t* status;
  int32_t timeout_ms;            /* how long a rank waits for its peers before it gives up (status word, then trap unless graceful);   */
                                 /* <= 0: 600 000 (ten minutes, NCCL's watchdog default) -- peers may legitimately be late (I/O, validation) */
} ef_dp_step_params;
int ef_dp_step(const ef_dp_step_params* p, void* stream);
int32_t ef_dp_step_grid(int32_t n);
/* Buffers shared between the ranks through CUDA IPC: ef_ipc_alloc = cudaMalloc (zero-filled) + cudaIpcGetMemHandle (64-byte handle for the
 * peers); ef_ipc_open = cudaIpcOpenMemHandle with lazy peer access on the CURRENT device; ef_ipc_close / ef_ipc_free undo them. */
int ef_ipc_alloc(int64_t bytes, void** ptr, unsigned char* handle64);
int ef_ipc_open(const unsigned char* handle64, void** ptr);
int ef_ipc_close(void* ptr);
int ef_ipc_free(void* ptr);
int ef_clip_adam(float* param, const float* g, float* m, float* v, int64_t n, const float* sqnorm, float max_norm, float lr,
                 float beta1, float beta2, float eps, int32_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVENTFLOW_H_ */
