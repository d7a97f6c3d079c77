/*
 * gwnet_b200 -- C ABI of the B200-native Graph WaveNet forward/backward hot path.
 *
 * This is the drop-in boundary: the reference (sklin93/Graph-WaveNet) is pure
 * Python/PyTorch and has no FFI of its own, so each entry point below names the
 * reference call site it replaces (file:line in /root/reference).  A reference
 * maintainer binds these with ctypes; the host-side mirror
 * of model.py in graph-wavenet_b200/ does exactly that.
 *
 * Conventions
 *   - plain C: raw device pointers, ints, floats; no torch / C++ types.
 *   - every function returns 0 on success, else a gwn_status / cudaError code;
 *     gwn_last_error() returns a thread-local message.  Nothing throws or exits.
 *   - the caller owns every buffer (parameters, activations, workspace); the
 *     library allocates no device memory (one explicit exception: gwn_p2p_alloc,
 *     the IPC-exportable gradient buffer of the data-parallel exchange).
 *   - all work is ordered on the cudaStream_t passed as `stream` (void*), on the
 *     caller's current device; no implicit synchronisation; re-entrant (forward
 *     is called from the main thread, backward from PyTorch's autograd thread).
 *     A plan additionally owns ONE internal non-blocking stream (+ four events,
 *     created at the first gwn_plan_forward, released by gwn_plan_destroy): the
 *     forward pass forks it from `stream` with an event, runs its parameter- and
 *     support-only operand preparation there, and joins it back with events
 *     before the first consumer and before it returns.  To the caller the call
 *     behaves as if everything ran on `stream` (also under stream capture, where
 *     the preparation becomes parallel graph branches); GWNET_B200_SIDE_STREAM=0
 *     keeps every launch on `stream`.  Kernels are launched with programmatic
 *     stream serialization and wait (griddepcontrol.wait) for the previous kernel
 *     of the stream before touching memory; GWNET_B200_PDL=0 turns that off.
 *   - activation tensors are fp32 in the "BLNC" physical layout
 *         x[b][l][n][c]   (batch, time, node, channel; channel innermost)
 *     which is the reference's logical NCHW tensor [B,C,N,L] viewed with strides
 *     (L*N*C, 1, C, N*C).  gwn_permute4d converts from/to any 4-D strided tensor.
 *   - there is NO CPU fallback: every entry point fails if no CUDA device backs
 *     the pointers.
 */
#ifndef GWNET_B200_H
#define GWNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWN_ABI_VERSION 4

typedef enum gwn_status {
  GWN_OK = 0,
  GWN_ERR_INVALID = 10001,     /* bad argument / unsupported shape */
  GWN_ERR_CUDA = 10002,        /* a CUDA runtime call failed (message has the detail) */
  GWN_ERR_UNSUPPORTED = 10003, /* valid but not implemented by this build */
  GWN_ERR_NO_DEVICE = 10004    /* no CUDA device: there is no CPU fallback */
} gwn_status;

/* Precision tier of the contractions (activations, parameters and gradients are stored in fp32 in every tier). */
typedef enum gwn_precision {
  GWN_PREC_FP32 = 0,  /* fp32 FMA everywhere (generic SIMT kernels): reference tier, 1e-4 parity             */
  GWN_PREC_TF32 = 1,  /* every contraction on tcgen05 kind::tf32 (operands truncated to TF32 by the tensor
                         core), fp32 accumulate in TMEM: the 2e-2 tier (measured ~2e-4 output error)          */
  GWN_PREC_BF16 = 2,  /* reserved: bf16 operands (not in this build; entry points reject it)                  */
  GWN_PREC_FP32X3 = 3 /* DEFAULT.  fp32-grade on tcgen05: every contraction as a 3xTF32 split
                         (A.B + A.B_lo + A_lo.B, remainders exact in fp32), fp32 accumulate: 1e-4 parity      */
} gwn_precision;

/* Dropout source for gcn (model.py:54). */
typedef enum gwn_dropout_mode {
  GWN_DROPOUT_NONE = 0,   /* eval mode or p == 0                                    */
  GWN_DROPOUT_MASK = 1,   /* caller supplies uint8 keep-masks (parity tests, G7)     */
  GWN_DROPOUT_PHILOX = 2  /* in-kernel Philox4x32-10 keyed by (seed, layer, element) */
} gwn_dropout_mode;

const char* gwn_last_error(void);
int gwn_abi_version(void);
/* Kernels launched by this library so far in this process (optionally reset). */
long long gwn_launch_count(int reset);
/* Number of CUDA devices visible; fills name (<=255 chars) and SM count of the current one. */
int gwn_device_info(int* n_devices, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor);

/* Per-operator device timing for the roofline report (bench.py): between gwn_profile_begin() and
 * gwn_profile_end() every operator of the plan is bracketed by CUDA events on its launching stream.
 * gwn_profile_end synchronises the device and writes a JSON array of
 * {"op", "calls", "ms", "bytes", "flops"} (bytes / flops = algorithmic work).
 * Do not use while a CUDA graph is being captured.                                                    */
int gwn_profile_begin(void);
int gwn_profile_end(char* buf, int len);

/* ------------------------------------------------------------------ layout helpers */
/* dst[i0,i1,i2,i3] = src[i0,i1,i2,i3] for two arbitrarily strided fp32 4-D tensors
 * (element strides).  Used at true layout boundaries only (NCHW <-> BLNC).        */
int gwn_permute4d(const float* src, const int64_t src_strides[4], float* dst, const int64_t dst_strides[4],
                  const int64_t sizes[4], void* stream);

/* Node contraction with the support given K-contiguous: y[b,l,m,c] = sum_k S[m*ld + k] * x[b,l,k,c].
 * (nconv forward is this with S = A^T, its input gradient with S = A.)  precision GWN_PREC_TF32 runs the
 * tcgen05/TMEM/TMA kernel (needs C == 32, ld % 4 == 0, 16-byte aligned pointers); GWN_PREC_FP32 the FMA tier. */
int gwn_node_contract(const float* x, const float* S, int64_t ld, float* y, int B, int L, int V, int C, int precision,
                      void* stream);
/* The same contraction in the fp32-grade 3xTF32 tier (GWN_PREC_FP32X3): S_lo = S - tf32_trunc(S) from gwn_split_lo. */
int gwn_node_contract_x3(const float* x, const float* S, const float* S_lo, int64_t ld, float* y, int B, int L, int V, int C,
                         void* stream);
/* lo[i] = src[i] - tf32_trunc(src[i]) (exact in fp32): the remainder plane of a 3xTF32 operand. */
int gwn_split_lo(const float* src, float* lo, int64_t n, void* stream);
/* First pipeline time-out recorded by the tcgen05 kernels (0 = none); synchronises the device. Debug aid. */
int gwn_tc_error_flag(int reset);
/* Debug aid: device buffer that receives a dump of the first pipeline stage of subsequent tcgen05 launches (NULL = off). */
void gwn_tc_debug_buffer(float* p);
void gwn_tc_debug_mode(int mode);   /* 0 = normal; non-zero = kernel self-test modes (development only) */

/* ------------------------------------------------------------------ stand-alone operators
 * nconv / gcn / nconv2 / gcn2 run in any tier.  GWN_PREC_FP32 needs no workspace (pass NULL).  The tensor-core tiers
 * (GWN_PREC_FP32X3, GWN_PREC_TF32: tcgen05 + TMA kernels, c_in = c_out = 32) prepare their operands -- zero-padded
 * K-contiguous supports A / A^T, 3xTF32 remainders, transposed weights, partial-result slots of the reductions -- in a
 * caller-owned, 16-byte aligned workspace whose size the *_workspace_floats / *_scratch_floats functions return.      */

/* ------------------------------------------------------------------ nconv (model.py:8-14)
 * y[b,l,w,c] = sum_v x[b,l,v,c] * A[v,w].   x,y: BLNC [B,L,V,C]; A: [V,V] row-major, ld = lda.
 * Replaces torch.einsum('ncvl,vw->ncwl') + .contiguous().                          */
/* floats of workspace for gwn_nconv_* (n_sets = 1) / gwn_nconv2_* (n_sets = B); support_grad = 1 when dA is wanted. */
size_t gwn_nconv_workspace_floats(int n_sets, int V, int precision, int support_grad);
int gwn_nconv_fwd(const float* x, const float* A, int64_t lda, float* y, int B, int L, int V, int C,
                  int precision, void* workspace, void* stream);
/* Autograd of nconv: dx = dy . A^T ; dA += x^T dy summed over (b,l,c) when dA != NULL
 * (dA is accumulated into -- zero it first).                                        */
int gwn_nconv_bwd(const float* dy, const float* x, const float* A, int64_t lda, float* dx, float* dA, int64_t ldda,
                  int B, int L, int V, int C, int precision, void* workspace, void* stream);

/* ------------------------------------------------------------------ linear (model.py:24-30)
 * 1x1 Conv2d with bias: y[p,co] = sum_ci W[co,ci] x[p,ci] + b[co], p over B*L*N positions (fp32 FMA kernels; inside
 * gcn and gwnet the same contraction runs on tcgen05). */
int gwn_linear_fwd(const float* x, const float* W, const float* bias, float* y, int64_t positions, int c_in, int c_out,
                   void* stream);
int gwn_linear_bwd(const float* dy, const float* x, const float* W, float* dx, float* dW, float* dbias,
                   int64_t positions, int c_in, int c_out, void* stream);

/* ------------------------------------------------------------------ gcn (model.py:32-55)
 * h = dropout(mlp(cat([x, A1 x, A1^2 x, ..., As^order x], channel))).
 * hops: caller buffer for the order*S diffused tensors, [order*S][B*L*V*C] (saved for backward).
 * keep_mask: uint8 [B*L*V*c_out] when dropout_mode == GWN_DROPOUT_MASK.             */
typedef struct gwn_gcn_desc {
  int B, L, V, C;          /* input BLNC dims; C = c_in per hop                      */
  int c_out;
  int n_supports;          /* S                                                      */
  int order;               /* K hops per support                                     */
  int precision;           /* gwn_precision                                          */
  int dropout_mode;        /* gwn_dropout_mode                                       */
  float dropout_p;
  uint64_t seed;           /* Philox key                                             */
  uint64_t offset;         /* Philox stream offset (layer id)                        */
} gwn_gcn_desc;

/* floats of forward workspace (0 in the fp32 tier); per_sample_supports = 1 for gwn_gcn2_fwd. */
size_t gwn_gcn_workspace_floats(const gwn_gcn_desc* d, int per_sample_supports);
int gwn_gcn_fwd(const gwn_gcn_desc* d, const float* x, const float* const* supports, const int64_t* lds,
                const float* W, const float* bias, const uint8_t* keep_mask, float* hops, float* y, void* workspace,
                void* stream);
/* dsupports[s] may be NULL (no gradient wanted for that support); non-NULL ones are accumulated into.
 * scratch: caller buffer of gwn_gcn_bwd_scratch_floats(d) floats (16-byte aligned; covers gwn_gcn_bwd and gwn_gcn2_bwd). */
size_t gwn_gcn_bwd_scratch_floats(const gwn_gcn_desc* d);
int gwn_gcn_bwd(const gwn_gcn_desc* d, const float* dy, const float* x, const float* const* supports, const int64_t* lds,
                const float* W, const uint8_t* keep_mask, const float* hops, float* dx, float* dW, float* dbias,
                float* const* dsupports, const int64_t* ldds, float* scratch, void* stream);

/* ------------------------------------------------------------------ nconv2 / gcn2 (model.py:16-22, 57-80)
 * The per-sample-graph operators of the fork: y[b,l,w,c] = sum_v x[b,l,v,c] * A[b][v,w], one [V,V] support per sample
 * (A: [B,V,V], sample stride lda_b, row stride lda).  Same semantics as gwn_nconv_* / gwn_gcn_* otherwise; the supports
 * of gwn_gcn2_* are `n_supports` tensors [B,V,V]; dsupports[s] (nullable) are accumulated into, same layout.  In the
 * tensor-core tiers all samples' graphs go through one launch (batched tensor maps).                                */
int gwn_nconv2_fwd(const float* x, const float* A, int64_t lda_b, int64_t lda, float* y, int B, int L, int V, int C,
                   int precision, void* workspace, void* stream);
int gwn_nconv2_bwd(const float* dy, const float* x, const float* A, int64_t lda_b, int64_t lda, float* dx, float* dA,
                   int64_t ldda_b, int64_t ldda, int B, int L, int V, int C, int precision, void* workspace, void* stream);
int gwn_gcn2_fwd(const gwn_gcn_desc* d, const float* x, const float* const* supports, const int64_t* lds_b, const int64_t* lds,
                 const float* W, const float* bias, const uint8_t* keep_mask, float* hops, float* y, void* workspace,
                 void* stream);
int gwn_gcn2_bwd(const gwn_gcn_desc* d, const float* dy, const float* x, const float* const* supports, const int64_t* lds_b,
                 const int64_t* lds, const float* W, const uint8_t* keep_mask, const float* hops, float* dx, float* dW,
                 float* dbias, float* const* dsupports, const int64_t* ldds_b, const int64_t* ldds, float* scratch,
                 void* stream);

/* ------------------------------------------------------------------ gwnet (model.py:82-241)
 * Whole-network plan: replaces gwnet.forward and its autograd graph.                 */
typedef struct gwn_config {
  int batch;
  int num_nodes;
  int seq_len;             /* T of the tensor handed to gwnet.forward                 */
  int in_dim, out_dim;
  int residual_channels, dilation_channels, skip_channels, end_channels;
  int kernel_size;         /* only 2 (the reference default) is supported             */
  int blocks, layers;
  int n_static_supports;   /* len(self.supports) without the adaptive one             */
  int gcn_bool;            /* constructor flag: gconv modules exist (model.py:156)    */
  int adaptive;            /* gcn_bool and addaptadj (model.py:114)                   */
  int gcn;                 /* gcn_bool and self.supports is not None (model.py:225)   */
  int order;               /* 2 (model.py:33)                                         */
  int apt_rank;            /* 10 (model.py:117-118)                                   */
  int precision;           /* gwn_precision                                           */
  float dropout;
  float bn_eps, bn_momentum;
  /* per-sample-graph variant gwnet_diff_G (model.py:244-407); all 0 for gwnet */
  int dilation_base;       /* first dilation of every block: 0/1 = gwnet (model.py:132), 4 = gwnet_diff_G (model.py:273) */
  int per_sample_supports; /* supports are [B,N,N], one graph per sample (model.py:313)                              */
  int adaptive_input;      /* an extra per-sample support softmax(relu(E1 E2)) from node embeddings passed to forward
                              (model.py:324-329,345-346: re-drawn every forward, not parameters, no gradient)      */
} gwn_config;

typedef struct gwn_plan gwn_plan;

int gwn_plan_create(const gwn_config* cfg, gwn_plan** out);
void gwn_plan_destroy(gwn_plan* p);
/* Bytes of caller-owned workspace: `fwd` holds everything forward saves for backward,
 * `bwd` is backward scratch.                                                         */
int gwn_plan_workspace_bytes(const gwn_plan* p, size_t* fwd, size_t* bwd);
/* Number of entries of the parameter table (reference state_dict order, App. F of SURVEY.md:
 * parameters and BN buffers) and of floats in the flat gradient buffer.              */
int gwn_plan_param_count(const gwn_plan* p, int* n_entries, int64_t* grad_floats);
/* Name, element offset into the flat gradient buffer (-1 for buffers) and element count of entry i. */
int gwn_plan_param_info(const gwn_plan* p, int i, char* name, int name_len, int64_t* grad_offset, int64_t* numel);
int gwn_plan_out_len(const gwn_plan* p, int* t_out, int* receptive_field);
/* Debug aid: text listing "<fwd|bwd> <name> <float offset> <floats>" of the workspace / scratch regions. */
int gwn_plan_debug_layout(const gwn_plan* p, char* buf, int len);

typedef struct gwn_forward_args {
  const void* const* params;     /* n_entries device pointers, state_dict order            */
  const float* const* supports;  /* n_static_supports device pointers [N,N]                */
  const int64_t* support_strides;/* 2 per support: (row stride, col stride) in elements; with per_sample_supports
                                    3 per support: (sample stride, row stride, col stride)  */
  const float* input;            /* [B,in_dim,N,T] fp32, any strides                        */
  int64_t input_strides[4];
  float* output;                 /* [B,out_dim,N,T_out] fp32 contiguous NCHW                */
  void* workspace;               /* fwd bytes                                               */
  int training;                  /* BN batch stats + running-stat update + dropout          */
  int dropout_mode;              /* gwn_dropout_mode (ignored when !training)               */
  const uint8_t* const* keep_masks; /* n_layers pointers [B*L_i*N*C] (GWN_DROPOUT_MASK)     */
  uint64_t seed;                 /* Philox key (GWN_DROPOUT_PHILOX)                         */
  void* stream;
  const uint64_t* seed_device;   /* optional: read the Philox key from device memory instead (CUDA graphs)  */
  const float* apt_e1;           /* adaptive_input: [B,N,apt_rank] contiguous                                */
  const float* apt_e2;           /* adaptive_input: [B,apt_rank,N] contiguous                                */
} gwn_forward_args;

int gwn_plan_forward(gwn_plan* p, const gwn_forward_args* a);

typedef struct gwn_backward_args {
  const void* const* params;
  const float* const* supports;
  const int64_t* support_strides;
  const float* input;
  int64_t input_strides[4];
  const float* grad_output;      /* [B,out_dim,N,T_out] fp32 contiguous NCHW                */
  const void* workspace;         /* the forward workspace of the matching forward call      */
  void* scratch;                 /* bwd bytes                                               */
  float* grad_flat;              /* grad_floats floats; zeroed by this call                  */
  float* grad_input;             /* [B,in_dim,N,T] contiguous NCHW or NULL                   */
  int training;                  /* the `training` flag of the matching forward call         */
  int dropout_mode;
  const uint8_t* const* keep_masks;
  uint64_t seed;
  void* stream;
  const uint64_t* seed_device;   /* as in the matching forward call                                         */
} gwn_backward_args;

int gwn_plan_backward(gwn_plan* p, const gwn_backward_args* a);

/* ------------------------------------------------------------------ trainer.train (engine.py:41-58)
 * One optimisation step as a fixed sequence of launches with no host round trip, so that it can be captured in
 * a CUDA graph:  gwn_plan_train_fwd_bwd  = forward + inverse_transform + masked MAE/MAPE/RMSE (Utils/util.py:116-117,
 * 510-552) + backward into the flat gradient buffer;  [the caller's gradient all-reduce for data parallelism];
 * gwn_adam_step = clip_grad_norm_ (engine.py:53-54) + Adam with L2 weight decay (engine.py:33,55) on flat buffers.
 * A small device control block (gwn_train_ctrl_bytes) carries the dropout key, the Adam step count and the
 * reduction accumulators from launch to launch.                                                              */
size_t gwn_train_ctrl_bytes(void);
/* Synchronous; (re)initialises the control block: dropout key, Adam step count. */
int gwn_train_ctrl_init(void* ctrl, uint64_t seed, int64_t step);
int gwn_train_ctrl_read(const void* ctrl, uint64_t* seed, int64_t* step);

typedef struct gwn_train_args {
  gwn_forward_args fwd;          /* as for gwn_plan_forward (training = 1); output = caller buffer [B,out_dim,N,T_out];
                                    seed / seed_device are ignored: the key comes from the control block            */
  void* scratch;                 /* bwd bytes                                                                       */
  float* grad_flat;              /* grad_floats floats; written                                                     */
  const float* target;           /* real_val [B,N,out_dim] fp32, any strides (engine.py:48)                         */
  int64_t target_strides[3];
  float scaler_mean, scaler_std; /* StandardScaler.inverse_transform (Utils/util.py:116-117)                        */
  void* ctrl;                    /* control block                                                                    */
  float* metrics;                /* device float[4]: masked MAE (the loss), MAPE, RMSE, [3] = gradient norm, written by
                                    gwn_adam_step                                                                    */
} gwn_train_args;
int gwn_plan_train_fwd_bwd(gwn_plan* p, const gwn_train_args* a);

/* trainer.eval (engine.py:119-130): forward in eval mode (BatchNorm running statistics, no dropout) + the three masked
 * metrics, no backward.  Uses fwd (training is forced to 0), target, target_strides, scaler_*, ctrl and metrics of
 * gwn_train_args; scratch / grad_flat are ignored.  metrics[0..2] = masked MAE, MAPE, RMSE.                        */
int gwn_plan_eval_metrics(gwn_plan* p, const gwn_train_args* a);

typedef struct gwn_adam_args {
  float* param_flat;             /* n floats, the layout of the plan's flat gradient buffer; updated in place        */
  float* grad_flat;              /* n floats; on return holds the (scaled, clipped) gradient, like p.grad           */
  float* exp_avg;                /* n floats, Adam first moment                                                       */
  float* exp_avg_sq;             /* n floats, Adam second moment                                                      */
  const uint8_t* live4;          /* n/4 bytes: 1 = this group of 4 floats belongs to a parameter that has a gradient
                                    (dead parameters are skipped by clip and Adam alike)                 */
  int64_t n;                     /* multiple of 4                                                                     */
  const float* hyper;            /* device float[8]: lr, beta1, beta2, eps, weight_decay, max_norm (<=0: no clip),
                                    grad_scale (1/world after a summing all-reduce), unused                          */
  void* ctrl;
  float* metrics;                /* device float[4] or NULL; [3] receives the total gradient norm                    */
  void* stream;
} gwn_adam_args;
int gwn_adam_step(const gwn_adam_args* a);

/* ------------------------------------------------------------------ data-parallel step tail over NVLink peer memory
 * (no reference counterpart: SURVEY.md section 8(e)).  One process per GPU; every rank's flat gradient buffer lives in an
 * allocation made by gwn_p2p_alloc -- gwn_p2p_header_bytes() of flags followed by the gradient floats, which is what the
 * rank passes as grad_flat to gwn_plan_train_fwd_bwd -- and is mapped into its peers with gwn_p2p_open from the 64-byte
 * IPC handle (exchanged by the host code, e.g. through torch.distributed).  gwn_allreduce_adam_step then replaces
 * [all-reduce; gwn_adam_step]: ONE kernel sums all ranks' gradients over peer memory (rank order: bit-identical on every
 * rank), writes the sum to sum_out and accumulates its squared norm; the Adam kernel reads sum_out (grad_scale = 1/world
 * averages) and leaves the clipped gradient in grad_flat.  Capturable in a CUDA graph.  All ranks must call it the same
 * number of times (the protocol's epoch is the Adam step count of the control block).  These are the only entry points
 * that allocate device memory; the caller frees it with gwn_p2p_close (peers) / gwn_p2p_free (own).                   */
size_t gwn_p2p_header_bytes(void);
int gwn_p2p_alloc(size_t bytes, void** base, unsigned char* handle64);
int gwn_p2p_open(const unsigned char* handle64, void** base);
int gwn_p2p_close(void* base);
int gwn_p2p_free(void* base);
typedef struct gwn_p2p_args {
  void* base[8];                 /* every rank's allocation as mapped in this process (base[rank] = own)               */
  int rank, world;               /* world in [2, 8]                                                                    */
  float* sum_out;                /* n floats, local                                                                    */
} gwn_p2p_args;
int gwn_allreduce_adam_step(const gwn_adam_args* a, const gwn_p2p_args* p);

#ifdef __cplusplus
}
#endif
#endif /* GWNET_B200_H */
