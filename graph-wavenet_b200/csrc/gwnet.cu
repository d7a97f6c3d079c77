// gwnet_b200: whole-network plan (gwn_plan_*) and the op-level C ABI (include/gwnet_b200.h).
// Replaces gwnet.forward (model.py:175-241) and its autograd graph with explicit fused launches.
#include "ops.cuh"
#include "nconv_tc_impl.cuh"
#include "train_tail.cuh"
#include "p2p_allreduce.cuh"

#include <atomic>
#include <cstdarg>
#include <map>
#include <mutex>
#include <vector>
#include <string>

namespace gwn {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
#if !GWN_EMU
bool pdl_enabled() {   // GWNET_B200_PDL=0 launches every kernel with plain stream ordering (A/B switch, read once)
  static const bool on = [] {
    const char* e = getenv("GWNET_B200_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
#endif
static thread_local int g_math = 0;
int current_math() { return g_math; }
void set_current_math(int m) { g_math = m; }
// gwn_precision -> GEMM math mode (gemm.cuh)
static int math_of(int precision) { return precision == GWN_PREC_TF32 ? 1 : (precision == GWN_PREC_FP32X3 ? 3 : 0); }

// ---- per-op device timing (see common.cuh)
#if !GWN_EMU
struct ProfRec {
  const char* tag;
  cudaEvent_t e0, e1;
  double bytes, flops;
};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
ProfScope::ProfScope(const char* tag, cudaStream_t stream, double bytes, double flops) : idx(-1), st(stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r;
  r.tag = tag; r.bytes = bytes; r.flops = flops;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, stream);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  idx = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx < (int)g_prof.size()) cudaEventRecord(g_prof[idx].e1, st);
}
#else
ProfScope::ProfScope(const char*, cudaStream_t stream, double, double) : idx(-1), st(stream) {}
ProfScope::~ProfScope() {}
#endif

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#if GWN_EMU
static int dev_memset(void* p, int v, size_t bytes, cudaStream_t) {
  if (bytes) memset(p, v, bytes);
  return 0;
}
#else
static int dev_memset(void* p, int v, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return 0;
  GWN_CUDA(cudaMemsetAsync(p, v, bytes, s));
  return 0;
}
#endif

static int require_device() {
#if GWN_EMU
  return 0;
#else
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no CUDA device available (%s); gwnet_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return GWN_ERR_NO_DEVICE;
  }
  return 0;
#endif
}

static int permute4d(const float* src, const int64_t* ss, float* dst, const int64_t* ds, const int64_t* sz,
                     cudaStream_t stream) {
  Strides4 s, d;
  Sizes4 n, ord;
  i64 total = 1;
  for (int i = 0; i < 4; ++i) {
    s.s[i] = ss[i]; d.s[i] = ds[i]; n.n[i] = sz[i]; ord.n[i] = i;
    total *= sz[i];
  }
  for (int i = 0; i < 4; ++i)   // iterate dst-major so that writes coalesce
    for (int j = i + 1; j < 4; ++j)
      if (d.s[ord.n[j]] > d.s[ord.n[i]]) std::swap(ord.n[i], ord.n[j]);
  GWN_LAUNCH_1D(permute4d_kernel, total, stream, src, s, dst, d, n, ord, total);
  return 0;
}

struct Entry {
  std::string name;
  i64 numel;
  i64 grad_off;  // -1 for buffers
};
struct LayerIdx {
  int fw, fb, gw, gb, rw, rb, sw, sb, bnw, bnb, bnm, bnv, bnt, mw, mb;
};

}  // namespace gwn

struct gwn_plan {
  gwn_config c;
  int nL, S, nseg, nseg_mod, RF, L0, pad, T_out, ld, ldo;
  int Bs;                          // support sets: 1, or batch with per_sample_supports
  std::vector<int> dil, L;  // per layer dilation and output length
  std::vector<gwn::Entry> entries;
  std::vector<gwn::LayerIdx> li;
  int i_nv1, i_nv2, i_startw, i_startb, i_e1w, i_e1b, i_e2w, i_e2b;
  gwn::i64 grad_floats;
  // forward workspace offsets (floats)
  gwn::i64 o_sup, o_supT, o_sup_lo, sup_span, o_x0, o_skip, o_e1, fwd_floats;
  // packed head weights for the tcgen05 position GEMMs (forward workspace / backward scratch)
  gwn::i64 o_hk_wcat, o_hk_wcat_lo, o_hk_bsum, o_hk_e1lo, o_hk_e2lo;
  gwn::i64 o_hb_wt, o_hb_wt_lo, o_hb_w1t, o_hb_w1t_lo, o_hb_w2t, o_hb_w2t_lo;
  bool head_tc;
  std::vector<gwn::i64> o_g, o_u, o_ac, o_mr, o_sums, o_pack;
  gwn::i64 pk_wp, pk_bf, pk_bg, pk_wd, pk_wt;   // offsets inside a layer's pack region (tensor-core tiers)
  gwn::i64 pk_wp_lo, pk_wd_lo, pk_wt_lo, pk_wm_lo;   // 3xTF32 remainders of the packed weights (fp32x3 tier)
  // backward scratch offsets (floats)
  gwn::i64 o_part, part_floats, o_buf0, o_buf1, o_dh, o_dg, o_dpre, o_dgh, o_dout, o_de1, o_dskip, o_dA, o_dR, o_bsum, bwd_floats;
  std::vector<gwn::i64> o_dsegs;   // per layer when the support gradient is deferred to ONE launch per backward pass
  bool defer_dA;
  std::vector<gwn::i64> o_dpre_l, o_dh_l;   // per layer when the weight gradients are deferred (else all = o_dpre / o_dh)
  bool defer_wgrad;
  // Side stream for the operand preparation that depends only on parameters / supports (weight packing, 3xTF32
  // remainders, support packing, adaptive adjacency): forked from the caller's stream at the start of a forward
  // pass and joined before the first consumer, so ~30 tiny launches leave the critical path (also as parallel
  // branches of a captured CUDA graph).  Created lazily on the device of the first forward call.
  cudaStream_t side = nullptr;
  cudaEvent_t side_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  gwn::i64 P(int i) const { return (gwn::i64)c.batch * L[i] * c.num_nodes; }
  gwn::i64 P0() const { return (gwn::i64)c.batch * L0 * c.num_nodes; }
  gwn::i64 PT() const { return (gwn::i64)c.batch * T_out * c.num_nodes; }
  int Lin(int i) const { return i == 0 ? L0 : L[i - 1]; }
};

namespace gwn {

static i64 align_up(i64 x) { return (x + 63) / 64 * 64; }

static int build_plan(gwn_plan* p) {
  const gwn_config& c = p->c;
  GWN_CHECK_ARG(c.batch >= 1 && c.num_nodes >= 1 && c.seq_len >= 1 && c.in_dim >= 1 && c.out_dim >= 1, "plan: bad dims");
  GWN_CHECK_ARG(c.kernel_size == 2, "plan: kernel_size %d unsupported (the reference default 2 only)", c.kernel_size);
  GWN_CHECK_ARG(c.residual_channels % 4 == 0 && c.dilation_channels % 4 == 0 && c.skip_channels % 4 == 0 &&
                    c.end_channels % 4 == 0,
                "plan: channel counts must be multiples of 4");
  GWN_CHECK_ARG(c.blocks >= 1 && c.layers >= 1 && c.blocks * c.layers <= MAXSEG, "plan: blocks*layers must be in [1,%d]", MAXSEG);
  GWN_CHECK_ARG(c.precision == GWN_PREC_FP32 || c.precision == GWN_PREC_TF32 || c.precision == GWN_PREC_FP32X3,
                "plan: precision %d not available in this build", c.precision);
  GWN_CHECK_ARG(c.precision != GWN_PREC_TF32 || !c.gcn || c.dilation_channels == 32,
                "plan: the tcgen05 (tf32) tier needs dilation_channels == 32");
  GWN_CHECK_ARG(c.order >= 1 && c.order <= MAXSUP, "plan: order must be in [1,%d]", MAXSUP);
  GWN_CHECK_ARG(c.dropout >= 0.f && c.dropout < 1.f, "plan: dropout must be in [0,1)");
  GWN_CHECK_ARG(c.n_static_supports >= 0, "plan: negative support count");
  GWN_CHECK_ARG(!c.gcn || c.gcn_bool, "plan: gcn active without gcn_bool");
  GWN_CHECK_ARG(!c.adaptive || c.gcn, "plan: adaptive adjacency implies an active gcn");
  p->nL = c.blocks * c.layers;
  GWN_CHECK_ARG(!(c.adaptive && c.adaptive_input), "plan: adaptive (parameters) and adaptive_input are exclusive");
  GWN_CHECK_ARG(!c.adaptive_input || c.per_sample_supports, "plan: adaptive_input implies per_sample_supports");
  GWN_CHECK_ARG(!c.per_sample_supports || !c.adaptive, "plan: per-sample supports have no trainable adaptive adjacency");
  const int supports_len = c.n_static_supports + ((c.adaptive || c.adaptive_input) ? 1 : 0);  // model.py:109-128
  p->Bs = c.per_sample_supports ? c.batch : 1;
  p->S = c.gcn ? supports_len : 0;
  GWN_CHECK_ARG(p->S <= MAXSUP, "plan: too many supports");
  GWN_CHECK_ARG(!c.gcn || p->S >= 1, "plan: gcn enabled without supports");
  GWN_CHECK_ARG(!(c.adaptive || c.adaptive_input) || (c.apt_rank >= 1 && c.apt_rank <= 16), "plan: apt_rank must be in [1,16]");
  p->nseg_mod = 1 + c.order * supports_len;   // c_in multiplier of the gconv modules (model.py:36)
  p->nseg = c.gcn ? p->nseg_mod : 1;
  GWN_CHECK_ARG(p->nseg_mod <= MAXSEG, "plan: too many gcn segments");
  p->dil.clear();
  int rf = 1;
  for (int b = 0; b < c.blocks; ++b) {   // model.py:130-155
    int d = c.dilation_base > 0 ? c.dilation_base : 1, scope = c.kernel_size - 1;
    for (int l = 0; l < c.layers; ++l) {
      p->dil.push_back(d);
      d *= 2;
      rf += scope;
      scope *= 2;
    }
  }
  p->RF = rf;
  p->L0 = std::max(c.seq_len, rf);
  p->pad = p->L0 - c.seq_len;
  p->L.resize(p->nL);
  int Lc = p->L0;
  for (int i = 0; i < p->nL; ++i) {
    Lc -= p->dil[i];
    p->L[i] = Lc;
  }
  p->T_out = Lc;
  GWN_CHECK_ARG(p->T_out >= 1, "plan: sequence shorter than the receptive field after padding");
  p->ld = round_up(c.num_nodes, 4);
  p->ldo = round_up(c.out_dim, 4);

  // ---- parameter table in reference state_dict order
  const int C = c.residual_channels, D = c.dilation_channels, Sk = c.skip_channels, E = c.end_channels;
  const int nL = p->nL;
  p->entries.clear();
  p->li.assign(nL, LayerIdx());
  i64 goff = 0;
  auto add = [&](const std::string& name, i64 numel, bool is_param) {
    Entry e;
    e.name = name;
    e.numel = numel;
    e.grad_off = -1;
    if (is_param) {
      e.grad_off = goff;
      goff += round_up64(numel, 4);
    }
    p->entries.push_back(e);
    return (int)p->entries.size() - 1;
  };
  auto nm = [](const char* grp, int i, const char* leaf) {
    char buf[128];
    snprintf(buf, sizeof(buf), "%s.%d.%s", grp, i, leaf);
    return std::string(buf);
  };
  p->i_nv1 = p->i_nv2 = -1;
  if (c.adaptive) {
    p->i_nv1 = add("nodevec1", (i64)c.num_nodes * c.apt_rank, true);
    p->i_nv2 = add("nodevec2", (i64)c.apt_rank * c.num_nodes, true);
  }
  for (int i = 0; i < nL; ++i) {
    p->li[i].fw = add(nm("filter_convs", i, "weight"), (i64)D * C * 2, true);
    p->li[i].fb = add(nm("filter_convs", i, "bias"), D, true);
  }
  for (int i = 0; i < nL; ++i) {
    p->li[i].gw = add(nm("gate_convs", i, "weight"), (i64)D * C * 2, true);
    p->li[i].gb = add(nm("gate_convs", i, "bias"), D, true);
  }
  for (int i = 0; i < nL; ++i) {
    p->li[i].rw = add(nm("residual_convs", i, "weight"), (i64)C * D, true);
    p->li[i].rb = add(nm("residual_convs", i, "bias"), C, true);
  }
  for (int i = 0; i < nL; ++i) {
    p->li[i].sw = add(nm("skip_convs", i, "weight"), (i64)Sk * D, true);
    p->li[i].sb = add(nm("skip_convs", i, "bias"), Sk, true);
  }
  for (int i = 0; i < nL; ++i) {
    p->li[i].bnw = add(nm("bn", i, "weight"), C, true);
    p->li[i].bnb = add(nm("bn", i, "bias"), C, true);
    p->li[i].bnm = add(nm("bn", i, "running_mean"), C, false);
    p->li[i].bnv = add(nm("bn", i, "running_var"), C, false);
    p->li[i].bnt = add(nm("bn", i, "num_batches_tracked"), 1, false);
  }
  for (int i = 0; i < nL; ++i) {
    p->li[i].mw = p->li[i].mb = -1;
    if (c.gcn_bool) {
      p->li[i].mw = add(nm("gconv", i, "mlp.mlp.weight"), (i64)C * p->nseg_mod * D, true);
      p->li[i].mb = add(nm("gconv", i, "mlp.mlp.bias"), C, true);
    }
  }
  p->i_startw = add("start_conv.weight", (i64)C * c.in_dim, true);
  p->i_startb = add("start_conv.bias", C, true);
  p->i_e1w = add("end_conv_1.weight", (i64)E * Sk, true);
  p->i_e1b = add("end_conv_1.bias", E, true);
  p->i_e2w = add("end_conv_2.weight", (i64)c.out_dim * E, true);
  p->i_e2b = add("end_conv_2.bias", c.out_dim, true);
  p->grad_floats = goff;

  // ---- forward workspace
  const i64 N = c.num_nodes;
  i64 o = 0;
  auto take = [&](i64 n) {
    i64 r = o;
    o += align_up(n);
    return r;
  };
  p->o_sup = take((i64)std::max(p->S, 1) * p->Bs * N * p->ld);     // [support][sample set][N][ld]
  p->o_supT = take((i64)std::max(p->S, 1) * p->Bs * N * p->ld);
  p->sup_span = p->o_supT + (i64)std::max(p->S, 1) * p->Bs * N * p->ld - p->o_sup;   // both packed support regions
  p->o_sup_lo = take(c.precision == GWN_PREC_FP32X3 ? p->sup_span : 0);        // their 3xTF32 remainders, same layout
  p->o_x0 = take(p->P0() * C);
  p->o_g.resize(nL); p->o_u.resize(nL); p->o_ac.resize(nL); p->o_mr.resize(nL); p->o_sums.resize(nL); p->o_pack.resize(nL);
  p->pk_wp = 0;
  p->pk_bf = p->pk_wp + align_up((i64)2 * D * 2 * C);
  p->pk_bg = p->pk_bf + align_up(D);
  p->pk_wd = p->pk_bg + align_up(D);
  p->pk_wt = p->pk_wd + align_up((i64)C * 4 * D);
  p->pk_wp_lo = p->pk_wt + align_up((i64)p->nseg * D * C);
  p->pk_wd_lo = p->pk_wp_lo + align_up((i64)2 * D * 2 * C);
  p->pk_wt_lo = p->pk_wd_lo + align_up((i64)C * 4 * D);
  p->pk_wm_lo = p->pk_wt_lo + align_up((i64)p->nseg * D * C);
  const i64 pack_floats = p->pk_wm_lo + align_up((i64)p->nseg * D * C);
  {   // BatchNorm forward sums of all layers: contiguous, so one memset clears them
    const i64 per = (i64)4 * C * GWN_STAT_REPL;   // GWN_STAT_REPL replicas of 2*C doubles
    const i64 o_s0 = take(per * nL);
    for (int i = 0; i < nL; ++i) p->o_sums[i] = o_s0 + per * i;
  }
  for (int i = 0; i < nL; ++i) {
    p->o_g[i] = take(p->P(i) * D * p->nseg);  // g_i followed by its hop tensors
    p->o_u[i] = take(p->P(i) * C);
    p->o_ac[i] = take(2 * C);
    p->o_mr[i] = take(2 * C);
    p->o_pack[i] = take(pack_floats);         // K-major packed weights for the tcgen05 position GEMMs
  }
  p->o_skip = take(p->PT() * Sk);
  p->o_e1 = take(p->PT() * E);
  // head on tcgen05: default widths only (K segments of 32, column tiles of <= 256)
  p->head_tc = (c.precision == GWN_PREC_TF32 || c.precision == GWN_PREC_FP32X3) && D == 32 && nL <= TP_MAXSEG &&
               Sk % 32 == 0 && Sk <= 512 && (Sk <= 128 || Sk % 256 == 0) && E % 32 == 0 && E <= 512 &&
               (E <= 128 || E % 256 == 0) && c.out_dim <= 256 && (nL * D <= 128 || (nL * D) % 256 == 0);
  {
    const bool hx3 = p->head_tc && c.precision == GWN_PREC_FP32X3;
    p->o_hk_wcat = take(p->head_tc ? (i64)Sk * nL * D : 0);
    p->o_hk_wcat_lo = take(hx3 ? (i64)Sk * nL * D : 0);
    p->o_hk_bsum = take(p->head_tc ? Sk : 0);
    p->o_hk_e1lo = take(hx3 ? (i64)E * Sk : 0);
    p->o_hk_e2lo = take(hx3 ? (i64)c.out_dim * E : 0);
  }
  p->fwd_floats = o;

  // ---- backward scratch
  o = 0;
  i64 maxP = p->P0();
  for (int i = 0; i < nL; ++i) maxP = std::max(maxP, p->P(i));
  i64 maxPi = 0;
  for (int i = 0; i < nL; ++i) maxPi = std::max(maxPi, p->P(i));
  // tcgen05 reductions (tf32 tier): per-CTA partial results, 160 slots of the largest accumulator tile in use
  const bool tc_tier = (c.precision == GWN_PREC_TF32 || c.precision == GWN_PREC_FP32X3) && C == 32 && D == 32;
  p->defer_dA = tc_tier && c.adaptive && c.num_nodes <= 512 && 2 * nL <= TR_MAXSRC;
  {
    // adaptive support gradient: one slot = at most 512 x 128 accumulator floats; large graphs tile the output into
    // (V/256) x (V/256) tiles with at least one CTA each
    const i64 ot = (i64)((N + 255) / 256) * ((N + 255) / 256);
    const i64 slots = std::max<i64>(160, ot);
    p->part_floats = tc_tier ? (c.adaptive ? slots * 512 * 128 : (i64)160 * 256 * 64) : 0;
  }
  p->o_part = take(p->part_floats);
  p->o_buf0 = take(maxP * C);
  p->o_buf1 = take(maxP * C);
  // Weight gradients of all layers in ONE tcgen05 reduction per kind at the end of the backward pass (like the support
  // gradient): a per-layer launch pays ~10 us of prologue / pipeline fill / slot write plus an ~8 us slot reduction,
  // 26 launches per step at the METR-LA shape; deferred, dpre_i and dh_i stay alive in per-layer buffers instead.
  {
    const char* e = getenv("GWNET_B200_DEFER_WGRAD");
    p->defer_wgrad = tc_tier && nL >= 2 && p->nseg <= 7 && !(e && e[0] == '0');
  }
  p->o_dh_l.assign(nL, 0);
  p->o_dpre_l.assign(nL, 0);
  p->o_dh = take(maxPi * C);   // du * dropout keep-mask (gradient wrt the pre-dropout mlp output)
  for (int i = 0; i < nL; ++i) p->o_dh_l[i] = p->defer_wgrad ? take(p->P(i) * C) : p->o_dh;
  p->o_dsegs.assign(nL, 0);
  if (p->defer_dA) {
    for (int i = 0; i < nL; ++i) p->o_dsegs[i] = take(p->P(i) * D * p->nseg);   // t tensors stay alive until the dA launch
  } else {
    const i64 o_shared = take(maxPi * D * p->nseg);
    for (int i = 0; i < nL; ++i) p->o_dsegs[i] = o_shared;
  }
  p->o_dg = take(maxPi * D);
  p->o_dpre = take(maxPi * 2 * D);
  for (int i = 0; i < nL; ++i) p->o_dpre_l[i] = p->defer_wgrad ? take(p->P(i) * 2 * D) : p->o_dpre;
  p->o_dgh = take((i64)nL * p->PT() * D);
  p->o_dout = take(p->PT() * p->ldo);
  p->o_de1 = take(p->PT() * E);
  p->o_dskip = take(p->PT() * Sk);
  p->o_dA = take(N * p->ld);
  p->o_dR = take(N * p->ld);
  p->o_bsum = take((i64)nL * 4 * C * GWN_STAT_REPL);   // per layer GWN_STAT_REPL replicas of 2*C doubles
  {
    const bool hx3 = p->head_tc && c.precision == GWN_PREC_FP32X3;
    p->o_hb_wt = take(p->head_tc ? (i64)Sk * nL * D : 0);
    p->o_hb_wt_lo = take(hx3 ? (i64)Sk * nL * D : 0);
    p->o_hb_w1t = take(p->head_tc ? (i64)E * Sk : 0);
    p->o_hb_w1t_lo = take(hx3 ? (i64)E * Sk : 0);
    p->o_hb_w2t = take(p->head_tc ? (i64)E * p->ldo : 0);
    p->o_hb_w2t_lo = take(hx3 ? (i64)E * p->ldo : 0);
  }
  p->bwd_floats = o;
  return 0;
}

template <class T>
static const T* P_(const void* const* params, int idx) {
  return idx < 0 ? nullptr : reinterpret_cast<const T*>(params[idx]);
}

static int check_ptr_table(const gwn_plan* p, const void* const* params) {
  GWN_CHECK_ARG(params != nullptr, "null parameter table");
  for (size_t i = 0; i < p->entries.size(); ++i) {
    GWN_CHECK_ARG(params[i] != nullptr, "parameter %s is null", p->entries[i].name.c_str());
    GWN_CHECK_ARG((reinterpret_cast<uintptr_t>(params[i]) & 15) == 0 || p->entries[i].numel == 1,
                  "parameter %s is not 16-byte aligned", p->entries[i].name.c_str());
  }
  return 0;
}

struct TcnGeom {
  int L_in, L_out, d;
};

// A operand of the gated conv (forward and the recompute in backward): two taps of the previous layer's
// pre-BN tensor with the BatchNorm affine folded into the load.
static LdRows tcn_rows(const gwn_plan* p, const float* prev, const float* prev_ac, int i) {
  LdRows a;
  memset(&a, 0, sizeof(a));
  const int N = p->c.num_nodes;
  a.p[0] = prev; a.p[1] = prev;
  a.rm[0] = make_remap(p->L[i], p->Lin(i), 0, N);
  a.rm[1] = make_remap(p->L[i], p->Lin(i), p->dil[i], N);
  a.set_wd(p->c.residual_channels);
  a.use_remap = 1;
  a.ac = prev_ac;
  return a;
}

// The same operand for posgemm.cuh (tensor-core tiers, C == 32): two K segments = the two taps.
static ARows tcn_arows(const gwn_plan* p, const float* prev, const float* prev_ac, int i) {
  ARows a;
  memset(&a, 0, sizeof(a));
  const int N = p->c.num_nodes;
  a.P[0] = prev; a.P[1] = prev;
  a.rmap[0] = 0; a.rmap[1] = 1;
  a.rm[0] = rowmap_shift(p->L[i], p->Lin(i), 0, N);
  a.rm[1] = rowmap_shift(p->L[i], p->Lin(i), p->dil[i], N);
  a.nseg = 2;
  a.rs = p->c.residual_channels;
  a.ac = prev_ac;
  return a;
}
// tcgen05 + TMA position GEMMs: tf32 tier with the reference's default widths.
static bool tcpos_ok(const gwn_plan* p) {
  return (p->c.precision == GWN_PREC_TF32 || p->c.precision == GWN_PREC_FP32X3) && p->c.residual_channels == 32 &&
         p->c.dilation_channels == 32;
}
static bool x3(const gwn_plan* p) { return p->c.precision == GWN_PREC_FP32X3; }
static TcPosArgs tcn_tcpos_args(const gwn_plan* p, const float* prev, const float* Wp, int i) {
  TcPosArgs t;
  memset(&t, 0, sizeof(t));
  const int N = p->c.num_nodes;
  t.seg[0] = TcPosSeg{prev, p->Lin(i) * N, 32, 0, 0};
  t.seg[1] = TcPosSeg{prev, p->Lin(i) * N, 32, 0, p->dil[i] * N};
  t.nseg = 2; t.nb = p->c.batch; t.rows_out = p->L[i] * N; t.Wp = Wp; t.N = 64;
  return t;
}
static bool pg_ok(const gwn_plan* p) {
  return current_math() != 0 && p->c.residual_channels == PG_WD && p->c.dilation_channels == PG_WD;
}

static DropoutSrc layer_dropout(const gwn_plan* p, int training, int mode, const uint8_t* const* masks, uint64_t seed, int i,
                                const uint64_t* seed_dev) {
  if (!training || !p->c.gcn) return make_dropout(GWN_DROPOUT_NONE, nullptr, 0, 0, 0.f);
  return make_dropout(mode, (mode == GWN_DROPOUT_MASK && masks) ? masks[i] : nullptr, seed, (uint64_t)i, p->c.dropout,
                      reinterpret_cast<const unsigned long long*>(seed_dev));
}

// Fork: returns the stream to launch preparation work on (the caller's own stream when the side stream is disabled).
static int side_fork(gwn_plan* p, cudaStream_t st, cudaStream_t* out) {
  *out = st;
#if !GWN_EMU
  static const bool off = [] {
    const char* e = getenv("GWNET_B200_SIDE_STREAM");
    return e && e[0] == '0';
  }();
  if (off) return 0;
  if (!p->side) {
    GWN_CUDA(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
    for (int k = 0; k < 4; ++k) GWN_CUDA(cudaEventCreateWithFlags(&p->side_ev[k], cudaEventDisableTiming));
  }
  GWN_CUDA(cudaEventRecord(p->side_ev[0], st));
  GWN_CUDA(cudaStreamWaitEvent(p->side, p->side_ev[0], 0));
  *out = p->side;
#else
  (void)p;
#endif
  return 0;
}
static int side_mark(gwn_plan* p, cudaStream_t sd, cudaStream_t st, int k) {   // "everything launched on sd so far"
#if !GWN_EMU
  if (sd != st) GWN_CUDA(cudaEventRecord(p->side_ev[k], sd));
#else
  (void)p; (void)sd; (void)st; (void)k;
#endif
  return 0;
}
static int side_join(gwn_plan* p, cudaStream_t sd, cudaStream_t st, int k) {   // st continues after mark k
#if !GWN_EMU
  if (sd != st) GWN_CUDA(cudaStreamWaitEvent(st, p->side_ev[k], 0));
#else
  (void)p; (void)sd; (void)st; (void)k;
#endif
  return 0;
}

static int plan_forward(gwn_plan* p, const gwn_forward_args* a) {
  MathScope math_scope(math_of(p->c.precision));
  GWN_TRY(require_device());
  GWN_TRY(check_ptr_table(p, a->params));
  GWN_CHECK_ARG(a->input && a->output && a->workspace, "forward: null input/output/workspace");
  const gwn_config& c = p->c;
  GWN_CHECK_ARG(!(a->training && c.dropout > 0.f && c.gcn) || a->dropout_mode != GWN_DROPOUT_MASK || a->keep_masks,
                "forward: GWN_DROPOUT_MASK without keep_masks");
  cudaStream_t st = (cudaStream_t)a->stream;
  float* ws = reinterpret_cast<float*>(a->workspace);
  const int N = c.num_nodes, C = c.residual_channels, D = c.dilation_channels, Sk = c.skip_channels, E = c.end_channels;
  const int nL = p->nL, B = c.batch;
  const void* const* prm = a->params;

  // ---- preparation that depends only on parameters / supports runs on the plan's side stream (see gwn_plan::side)
  cudaStream_t sd = st;
  GWN_TRY(side_fork(p, st, &sd));
  if (tcpos_ok(p)) {   // layer 0's packed gated-conv weights (no BatchNorm fold below the first layer)
    float* pk = ws + p->o_pack[0];
    GWN_LAUNCH_WARP_ROWS(pack_tcn_fwd_kernel, 2 * D, sd, P_<float>(prm, p->li[0].fw), P_<float>(prm, p->li[0].gw),
                  P_<float>(prm, p->li[0].fb), P_<float>(prm, p->li[0].gb), (const float*)nullptr, pk + p->pk_wp, pk + p->pk_bf,
                  pk + p->pk_bg, D, C, x3(p) ? pk + p->pk_wp_lo : (float*)nullptr);
  }
  // ---- supports: pack static ones, compute the adaptive one (model.py:185-188); with per-sample graphs one set per
  // sample (model.py:313,345-346).  Layout of both packed regions: [support][sample set][N][ld].
  const int Bs = p->Bs;
  const i64 sup_sz = (i64)N * p->ld;
  SupportView supF[MAXSUP], supB[MAXSUP];
  TcSupports tcF;
  memset(&tcF, 0, sizeof(tcF));
  const int sstr = c.per_sample_supports ? 3 : 2;
  for (int s = 0; s < p->S; ++s) {
    for (int b = 0; b < Bs; ++b) {
      float* Ap = ws + p->o_sup + ((i64)s * Bs + b) * sup_sz;
      float* ATp = ws + p->o_supT + ((i64)s * Bs + b) * sup_sz;
      if (s < c.n_static_supports) {
        GWN_CHECK_ARG(a->supports && a->supports[s] && a->support_strides, "forward: support %d missing", s);
        const int64_t* ss = a->support_strides + sstr * s;
        const float* As = a->supports[s] + (c.per_sample_supports ? (i64)b * ss[0] : 0);
        GWN_LAUNCH_1D(support_pack_kernel, sup_sz, sd, As, (i64)ss[sstr - 2], (i64)ss[sstr - 1], Ap, ATp, N, p->ld);
      } else if (c.adaptive_input) {
        GWN_CHECK_ARG(a->apt_e1 && a->apt_e2, "forward: adaptive_input without node embeddings");
        GWN_LAUNCH_WARP_ROWS(adp_fwd_kernel, N, sd, a->apt_e1 + (i64)b * N * c.apt_rank, a->apt_e2 + (i64)b * c.apt_rank * N,
                             c.apt_rank, Ap, ATp, N, p->ld);
      } else {
        GWN_LAUNCH_WARP_ROWS(adp_fwd_kernel, N, sd, P_<float>(prm, p->i_nv1), P_<float>(prm, p->i_nv2), c.apt_rank, Ap, ATp, N,
                             p->ld);
      }
    }
    float* Ap0 = ws + p->o_sup + (i64)s * Bs * sup_sz;
    float* ATp0 = ws + p->o_supT + (i64)s * Bs * sup_sz;
    supF[s] = support_padded(Ap0, p->ld);
    supB[s] = support_padded(ATp0, p->ld);
    tcF.S[s] = ATp0;   // forward contraction y[w] = sum_v A[v,w] x[v]: K-contiguous rows are those of A^T
    if (x3(p) && tcpos_ok(p)) tcF.Slo[s] = ws + p->o_sup_lo + (ATp0 - (ws + p->o_sup));
  }
  if (x3(p) && tcpos_ok(p) && p->S > 0)
    GWN_LAUNCH_1D(split_lo_kernel, p->sup_span, sd, (const float*)(ws + p->o_sup), ws + p->o_sup_lo, p->sup_span);
  GWN_TRY(side_mark(p, sd, st, 1));   // layer 0's gated conv and the node contractions may start
  if (tcpos_ok(p)) {
    if (x3(p)) {   // 3xTF32 remainders of every layer's mlp weights
      for (int i = 0; i < nL; ++i) {
        const float* Wm = P_<float>(prm, c.gcn ? p->li[i].mw : p->li[i].rw);
        GWN_LAUNCH_1D(split_lo_kernel, (i64)C * p->nseg * D, sd, Wm, ws + p->o_pack[i] + p->pk_wm_lo, (i64)C * p->nseg * D);
      }
    }
    GWN_TRY(side_mark(p, sd, st, 2));   // the mlp of layer 0 may start
#if !GWN_EMU
    if (p->head_tc) {   // head: concatenated skip weights + summed bias, remainders of the end convs
      const bool h3 = x3(p);
      SkipWeights sw;
      memset(&sw, 0, sizeof(sw));
      for (int i = 0; i < nL; ++i) { sw.w[i] = P_<float>(prm, p->li[i].sw); sw.b[i] = P_<float>(prm, p->li[i].sb); }
      GWN_LAUNCH_1D(pack_skip_kernel, (i64)Sk * nL * D + Sk, sd, sw, nL, D, Sk, ws + p->o_hk_wcat,
                    h3 ? ws + p->o_hk_wcat_lo : (float*)nullptr, (float*)nullptr, (float*)nullptr, ws + p->o_hk_bsum);
      if (h3) {
        GWN_LAUNCH_1D(split_lo_kernel, (i64)E * Sk, sd, P_<float>(prm, p->i_e1w), ws + p->o_hk_e1lo, (i64)E * Sk);
        GWN_LAUNCH_1D(split_lo_kernel, (i64)c.out_dim * E, sd, P_<float>(prm, p->i_e2w), ws + p->o_hk_e2lo, (i64)c.out_dim * E);
      }
    }
#endif
    // operands of the backward pass that live in the forward workspace: transposed mlp weights, gated-conv dgrad weights
    for (int i = 0; i < nL; ++i) {
      float* pk = ws + p->o_pack[i];
      if (i < nL - 1) {
        const float* Wm = P_<float>(prm, c.gcn ? p->li[i].mw : p->li[i].rw);
        GWN_LAUNCH_1D(transpose_kernel, (i64)C * p->nseg * D, sd, Wm, pk + p->pk_wt, C, p->nseg * D,
                      x3(p) ? pk + p->pk_wt_lo : (float*)nullptr);
      }
      GWN_LAUNCH_1D(pack_tcn_dgrad_kernel, (i64)C * 4 * D, sd, P_<float>(prm, p->li[i].fw), P_<float>(prm, p->li[i].gw),
                    pk + p->pk_wd, D, C, x3(p) ? pk + p->pk_wd_lo : (float*)nullptr);
    }
    GWN_TRY(side_mark(p, sd, st, 3));
  }
  tcF.ld = p->ld;
  tcF.precision = c.precision;
  (void)supB;
  // ---- BN bookkeeping
  if (a->training) {
    GWN_TRY(dev_memset(ws + p->o_sums[0], 0, sizeof(double) * 2 * C * GWN_STAT_REPL * nL, st));
  } else {
    for (int i = 0; i < nL; ++i)
      GWN_LAUNCH_1D(bn_eval_kernel, C, st, P_<float>(prm, p->li[i].bnw), P_<float>(prm, p->li[i].bnb),
                    P_<float>(prm, p->li[i].bnm), P_<float>(prm, p->li[i].bnv), c.bn_eps, ws + p->o_ac[i], ws + p->o_mr[i], C);
  }
  // ---- start conv (+ left pad)
  {
    ProfScope prof("start_conv_fwd", st, 4.0 * p->P0() * (c.in_dim + C), 2.0 * p->P0() * c.in_dim * C);
    Strides4 is;
    for (int k = 0; k < 4; ++k) is.s[k] = a->input_strides[k];
#if !GWN_EMU
    if (C == 32 && c.in_dim <= START_MAXF && p->P0() < 2147483647LL) {
      GWN_CUDA(launch_kernel(start_fwd32b_kernel, dim3((unsigned)(B * p->L0)), dim3(256), 0, st, a->input, is,
                             P_<float>(prm, p->i_startw), P_<float>(prm, p->i_startb), ws + p->o_x0, c.in_dim, N, p->L0, p->pad));
      count_launch();
    } else
#endif
    GWN_LAUNCH_1D(start_fwd_kernel, p->P0() * C, st, a->input, is, P_<float>(prm, p->i_startw), P_<float>(prm, p->i_startb),
                  ws + p->o_x0, B, c.in_dim, N, p->L0, p->pad, C);
  }
  GWN_TRY(side_join(p, sd, st, 1));
  // ---- WaveNet layers (model.py:192-236)
  bool packed_next = false;
  for (int i = 0; i < nL; ++i) {
    const float* prev = i == 0 ? ws + p->o_x0 : ws + p->o_u[i - 1];
    const float* prev_ac = i == 0 ? nullptr : ws + p->o_ac[i - 1];
    float* g = ws + p->o_g[i];
    const i64 Pi = p->P(i);
    {  // gated dilated conv
      ProfScope prof("gated_tcn_fwd", st, 4.0 * ((double)B * p->Lin(i) * N * C + (double)Pi * D), 2.0 * Pi * 2 * C * 2 * D);
      LdRows la = tcn_rows(p, prev, prev_ac, i);
      LdWTcn lb{P_<float>(prm, p->li[i].fw), P_<float>(prm, p->li[i].gw), C};
      EpGate ep{g, P_<float>(prm, p->li[i].fb), P_<float>(prm, p->li[i].gb), D};
      int pst = -1;
      if (tcpos_ok(p)) {
        float* pk = ws + p->o_pack[i];
        if (i > 0 && !packed_next)   // BatchNorm of the layer below folded in (layer 0 was packed on the side stream)
          GWN_LAUNCH_WARP_ROWS(pack_tcn_fwd_kernel, 2 * D, st, P_<float>(prm, p->li[i].fw), P_<float>(prm, p->li[i].gw),
                        P_<float>(prm, p->li[i].fb), P_<float>(prm, p->li[i].gb), prev_ac, pk + p->pk_wp, pk + p->pk_bf,
                        pk + p->pk_bg, D, C, x3(p) ? pk + p->pk_wp_lo : (float*)nullptr);
        RowGate eg;
        memset(&eg, 0, sizeof(eg));
        eg.y = g; eg.bf = pk + p->pk_bf; eg.bg = pk + p->pk_bg;
        TcPosArgs ta = tcn_tcpos_args(p, prev, pk + p->pk_wp, i);
        ta.out = g; ta.out_width = D; ta.out_nblk = 1;
        if (x3(p)) ta.Wp_lo = pk + p->pk_wp_lo;
        pst = launch_tcpos<64>(ta, eg, st);
        if (pst > 0) return pst;
      }
      if (pst < 0 && pg_ok(p)) {
        pst = launch_posgemm<TPG, 64>(tcn_arows(p, prev, prev_ac, i), lb, ep, Pi, 2 * D, st);
        if (pst > 0) return pst;
      }
      if (pst < 0) {
        GemmShape sh{Pi, 2 * D, 2 * C, 1, 1};
        GWN_TRY((launch_gemm<TPos64>(la, lb, ep, sh, st)));
      }
    }
    const float* segs[MAXSEG];
    for (int q = 0; q < p->nseg; ++q) segs[q] = g + (i64)q * Pi * D;
    MlpFwdArgs m;
    memset(&m, 0, sizeof(m));
    m.tf32_tc = tcpos_ok(p) ? 1 : 0;
    float* pk_m = ws + p->o_pack[i];
    if (c.gcn) {
      const bool ps_tc = c.per_sample_supports && tcpos_ok(p);   // tcgen05 tiers: all samples' graphs in one launch
      if (!c.per_sample_supports || ps_tc) {
        GcnShape gs{B, p->L[i], N, D, C, p->S, c.order};
        TcSupports tf = tcF;
        tf.per_sample = ps_tc ? 1 : 0;
        tf.batch_stride = sup_sz;
        GWN_TRY(gcn_hops_forward(gs, g, supF, g + Pi * D, st, &tf));
      } else {   // fp32 reference tier: the sample's slab against its own support set, one launch group per sample
        GcnShape g1{1, p->L[i], N, D, C, p->S, c.order};
        const i64 slab = (i64)p->L[i] * N * D;
        for (int b = 0; b < B; ++b) {
          SupportView sf[MAXSUP];
          TcSupports tb = tcF;
          for (int s = 0; s < p->S; ++s) {
            sf[s] = support_padded(supF[s].p + (i64)b * sup_sz, p->ld);
            tb.S[s] = tcF.S[s] + (i64)b * sup_sz;
            if (tcF.Slo[s]) tb.Slo[s] = tcF.Slo[s] + (i64)b * sup_sz;
          }
          GWN_TRY(gcn_hops_forward(g1, g + b * slab, sf, g + Pi * D + b * slab, st, &tb, Pi * D));
        }
      }
      m.W = P_<float>(prm, p->li[i].mw);
      m.bias = P_<float>(prm, p->li[i].mb);
    } else {  // model.py:232
      m.W = P_<float>(prm, p->li[i].rw);
      m.bias = P_<float>(prm, p->li[i].rb);
    }
    m.segs = segs; m.nseg = p->nseg; m.P = Pi; m.D = D; m.C_out = C;
    m.drop = layer_dropout(p, a->training, a->dropout_mode, a->keep_masks, a->seed, i, a->seed_device);
    m.res = prev;
    m.nb = B; m.rows_per_sample = p->L[i] * N; m.res_rows_src = p->Lin(i) * N; m.res_rshift = (p->Lin(i) - p->L[i]) * N;
    m.rrm = make_remap(p->L[i], p->Lin(i), p->Lin(i) - p->L[i], N);
    m.rac = prev_ac;
    m.stats = a->training ? reinterpret_cast<double*>(ws + p->o_sums[i]) : nullptr;
    m.y = ws + p->o_u[i];
    if (tcpos_ok(p) && x3(p)) m.W_lo = pk_m + p->pk_wm_lo;   // split on the side stream
    if (i == 0 && tcpos_ok(p)) GWN_TRY(side_join(p, sd, st, 2));
    GWN_TRY(mlp_forward(m, st));
    packed_next = false;
    if (a->training && tcpos_ok(p) && i + 1 < nL) {   // finalize + the next layer's packed gated-conv weights
      float* pk = ws + p->o_pack[i + 1];
      GWN_LAUNCH_WARP_ROWS(bn_finalize_pack_kernel, 2 * D + 1, st, reinterpret_cast<const double*>(ws + p->o_sums[i]), (double)Pi,
                    P_<float>(prm, p->li[i].bnw), P_<float>(prm, p->li[i].bnb), const_cast<float*>(P_<float>(prm, p->li[i].bnm)),
                    const_cast<float*>(P_<float>(prm, p->li[i].bnv)), const_cast<long long*>(P_<long long>(prm, p->li[i].bnt)),
                    c.bn_eps, c.bn_momentum, ws + p->o_ac[i], ws + p->o_mr[i], C, P_<float>(prm, p->li[i + 1].fw),
                    P_<float>(prm, p->li[i + 1].gw), P_<float>(prm, p->li[i + 1].fb), P_<float>(prm, p->li[i + 1].gb),
                    pk + p->pk_wp, pk + p->pk_bf, pk + p->pk_bg, D, x3(p) ? pk + p->pk_wp_lo : (float*)nullptr);
      packed_next = true;
    } else if (a->training) {
      GWN_LAUNCH_1D(bn_finalize_kernel, C, st, reinterpret_cast<const double*>(ws + p->o_sums[i]), (double)Pi,
                    P_<float>(prm, p->li[i].bnw), P_<float>(prm, p->li[i].bnb), const_cast<float*>(P_<float>(prm, p->li[i].bnm)),
                    const_cast<float*>(P_<float>(prm, p->li[i].bnv)), const_cast<long long*>(P_<long long>(prm, p->li[i].bnt)),
                    c.bn_eps, c.bn_momentum, ws + p->o_ac[i], ws + p->o_mr[i], C);
    }
  }
  // ---- head: skip sum over the live columns (G5), relu, end convs (model.py:216-222,238-240)
  if (tcpos_ok(p)) GWN_TRY(side_join(p, sd, st, 3));
  const i64 PT = p->PT();
  ProfScope prof_head("head_fwd", st, 4.0 * PT * ((double)nL * D + 2.0 * Sk + 2.0 * E + c.out_dim),
                      2.0 * PT * ((double)nL * D * Sk + (double)Sk * E + (double)E * c.out_dim));
  bool head_done = false;
#if !GWN_EMU
  if (p->head_tc && PT < 2147483647LL) {   // the three head layers as tcgen05 position GEMMs with streamed weights
    const bool h3 = x3(p);
    const int TN = h3 ? 128 : 256;   // output columns per tile: two stages of [A | W blk] (x2 planes in 3xTF32 mode) must fit
    int hs;
    {  // skip = relu(sum_i W_i g_i[live columns] + sum_i b_i)
      TcPosArgs t;
      memset(&t, 0, sizeof(t));
      for (int i = 0; i < nL; ++i) t.seg[i] = TcPosSeg{ws + p->o_g[i], p->L[i] * N, 32, 0, (p->L[i] - p->T_out) * N};
      t.nseg = nL; t.nb = B; t.rows_out = p->T_out * N; t.Wp = ws + p->o_hk_wcat; t.Wp_lo = h3 ? ws + p->o_hk_wcat_lo : nullptr;
      t.N = Sk <= TN ? Sk : TN; t.wstream = 1; t.N_total = Sk;
      t.out = ws + p->o_skip; t.out_width = Sk; t.out_nblk = Sk / 32;
      RowDense ep;
      memset(&ep, 0, sizeof(ep));
      ep.bias = ws + p->o_hk_bsum; ep.relu = 1;
      hs = launch_tcpos<0>(t, ep, st);
      if (hs > 0) return hs;
    }
    if (hs == 0) {  // e1 = relu(W1 skip + b1)
      TcPosArgs t;
      memset(&t, 0, sizeof(t));
      for (int q = 0; q < Sk / 32; ++q) t.seg[q] = TcPosSeg{ws + p->o_skip, (int)PT, Sk, 32 * q, 0};
      t.nseg = Sk / 32; t.nb = 1; t.rows_out = (int)PT; t.Wp = P_<float>(prm, p->i_e1w); t.Wp_lo = h3 ? ws + p->o_hk_e1lo : nullptr;
      t.N = E <= TN ? E : TN; t.wstream = 1; t.N_total = E;
      t.out = ws + p->o_e1; t.out_width = E; t.out_nblk = E / 32;
      RowDense ep;
      memset(&ep, 0, sizeof(ep));
      ep.bias = P_<float>(prm, p->i_e1b); ep.relu = 1;
      hs = launch_tcpos<0>(t, ep, st);
      if (hs > 0) return hs;
      GWN_CHECK_ARG(hs == 0, "forward: head layer 2 not eligible for the tcgen05 path after layer 1 ran on it");
    }
    if (hs == 0) {  // out = W2 e1 + b2, written in the reference's NCHW layout
      TcPosArgs t;
      memset(&t, 0, sizeof(t));
      for (int q = 0; q < E / 32; ++q) t.seg[q] = TcPosSeg{ws + p->o_e1, (int)PT, E, 32 * q, 0};
      t.nseg = E / 32; t.nb = 1; t.rows_out = (int)PT; t.Wp = P_<float>(prm, p->i_e2w); t.Wp_lo = h3 ? ws + p->o_hk_e2lo : nullptr;
      t.w_rows = c.out_dim;
      RowNCHW ep;
      memset(&ep, 0, sizeof(ep));
      ep.y = a->output; ep.bias = P_<float>(prm, p->i_e2b); ep.O = c.out_dim; ep.N = N; ep.T = p->T_out;
      if (c.out_dim <= 16) {   // the 12-step configurations: 16 output columns, weights resident
        t.N = 16;
        hs = launch_tcpos<16>(t, ep, st);
      } else {                 // longer horizons (the fork's seq_length 48): wider tile, weights streamed per k-block
        t.N = round_up(c.out_dim, 16); t.wstream = 1; t.N_total = t.N;
        hs = launch_tcpos<0>(t, ep, st);
      }
      if (hs > 0) return hs;
      GWN_CHECK_ARG(hs == 0, "forward: head layer 3 not eligible for the tcgen05 path after layers 1-2 ran on it");
      head_done = true;
    }
  }
#endif
  if (head_done) return 0;
  {
    LdRows la;
    memset(&la, 0, sizeof(la));
    LdWK lb;
    memset(&lb, 0, sizeof(lb));
    EpRows ep;
    memset(&ep, 0, sizeof(ep));
    for (int i = 0; i < nL; ++i) {
      la.p[i] = ws + p->o_g[i];
      la.rm[i] = make_remap(p->T_out, p->L[i], p->L[i] - p->T_out, N);
      lb.p[i] = P_<float>(prm, p->li[i].sw);
      ep.bias[i] = P_<float>(prm, p->li[i].sb);
    }
    la.set_wd(D); la.use_remap = 1;
    lb.set_wd(D); lb.ldw = D;
    ep.y = ws + p->o_skip; ep.ldy = Sk; ep.M = PT; ep.nbias = nL; ep.relu = 1;
    GemmShape sh{PT, Sk, nL * D, 1, 1};
    GWN_TRY((launch_gemm<TBig>(la, lb, ep, sh, st)));
  }
  {
    LdRows la;
    memset(&la, 0, sizeof(la));
    la.p[0] = ws + p->o_skip; la.set_wd(Sk);
    LdWK lb;
    memset(&lb, 0, sizeof(lb));
    lb.p[0] = P_<float>(prm, p->i_e1w); lb.set_wd(Sk); lb.ldw = Sk;
    EpRows ep;
    memset(&ep, 0, sizeof(ep));
    ep.y = ws + p->o_e1; ep.ldy = E; ep.M = PT; ep.bias[0] = P_<float>(prm, p->i_e1b); ep.nbias = 1; ep.relu = 1;
    GemmShape sh{PT, E, Sk, 1, 1};
    GWN_TRY((launch_gemm<TBig>(la, lb, ep, sh, st)));
  }
  {
    LdRows la;
    memset(&la, 0, sizeof(la));
    la.p[0] = ws + p->o_e1; la.set_wd(E);
    LdWK lb;
    memset(&lb, 0, sizeof(lb));
    lb.p[0] = P_<float>(prm, p->i_e2w); lb.set_wd(E); lb.ldw = E;
    EpNCHW ep;
    ep.y = a->output; ep.bias = P_<float>(prm, p->i_e2b); ep.N = N; ep.T = p->T_out;
    ep.sb = (i64)c.out_dim * N * p->T_out; ep.so = (i64)N * p->T_out; ep.sn = p->T_out; ep.st = 1;
    GemmShape sh{PT, c.out_dim, E, 1, 1};
    GWN_TRY((launch_gemm<TPos32>(la, lb, ep, sh, st)));
  }
  return 0;
}

static int plan_backward(gwn_plan* p, const gwn_backward_args* a, bool dout_ready = false) {
  MathScope math_scope(math_of(p->c.precision));
  GWN_TRY(require_device());
  GWN_TRY(check_ptr_table(p, a->params));
  GWN_CHECK_ARG((a->grad_output || dout_ready) && a->workspace && a->scratch && a->grad_flat && a->input, "backward: null argument");
  const gwn_config& c = p->c;
  cudaStream_t st = (cudaStream_t)a->stream;
  const float* ws = reinterpret_cast<const float*>(a->workspace);
  float* sc = reinterpret_cast<float*>(a->scratch);
  float* gf = a->grad_flat;
  const int N = c.num_nodes, C = c.residual_channels, D = c.dilation_channels, Sk = c.skip_channels, E = c.end_channels;
  const int nL = p->nL, B = c.batch, O = c.out_dim;
  const void* const* prm = a->params;
  const i64 PT = p->PT();
  auto G = [&](int idx) { return gf + p->entries[idx].grad_off; };
  const int training = a->training ? 1 : 0;
  const int dmode = a->dropout_mode;

  GWN_TRY(dev_memset(gf, 0, sizeof(float) * p->grad_floats, st));
  GWN_TRY(dev_memset(sc + p->o_bsum, 0, sizeof(float) * (i64)nL * 4 * C * GWN_STAT_REPL, st));
  if (c.adaptive) GWN_TRY(dev_memset(sc + p->o_dA, 0, sizeof(float) * (i64)N * p->ld, st));
  if (!dout_ready) GWN_TRY(dev_memset(sc + p->o_dout, 0, sizeof(float) * PT * p->ldo, st));

  SupportView supB[MAXSUP];
  TcSupports tcB;
  memset(&tcB, 0, sizeof(tcB));
  for (int s = 0; s < p->S; ++s) {
    supB[s] = support_padded(ws + p->o_supT + (i64)s * p->Bs * N * p->ld, p->ld);
    tcB.S[s] = ws + p->o_sup + (i64)s * p->Bs * N * p->ld;   // dx[v] = sum_w A[v,w] dy[w]: K-contiguous rows are those of A
    if (x3(p) && tcpos_ok(p)) tcB.Slo[s] = ws + p->o_sup_lo + (i64)s * p->Bs * N * p->ld;
  }
  tcB.ld = p->ld;
  tcB.precision = c.precision;

  // ---- head backward
  ProfScope* prof_hb = new ProfScope("head_bwd", st, 4.0 * PT * (c.out_dim + 3.0 * E + 3.0 * Sk + 2.0 * nL * D),
                                     2.0 * PT * (2.0 * E * c.out_dim + 2.0 * Sk * E + 2.0 * nL * D * Sk));
  if (!dout_ready) {
    // grad_output [B,O,N,T_out] contiguous -> dout [P_T, ldo]
    int64_t sz[4] = {B, O, N, p->T_out};
    int64_t ss[4] = {(int64_t)O * N * p->T_out, (int64_t)N * p->T_out, p->T_out, 1};
    int64_t ds[4] = {(int64_t)p->T_out * N * p->ldo, 1, p->ldo, (int64_t)N * p->ldo};
    GWN_TRY(permute4d(a->grad_output, ss, sc + p->o_dout, ds, sz, st));
  }
  const float* dout = sc + p->o_dout;
  const float* e1 = ws + p->o_e1;
  const float* skip = ws + p->o_skip;
  float* de1 = sc + p->o_de1;
  float* dskip = sc + p->o_dskip;
  float* dgh = sc + p->o_dgh;
  GWN_CHECK_ARG(PT < 2147483647LL, "backward: too many output positions");
  TcScratch tsc{p->part_floats > 0 ? sc + p->o_part : nullptr, p->part_floats, x3(p) ? 1 : 0};
  bool hb_tc = false;
#if !GWN_EMU
  if (p->head_tc) {   // the three input-gradient GEMMs of the head on tcgen05 (weight gradients below stay on the generic kernel)
    const bool h3 = x3(p);
    const int TN = h3 ? 128 : 256;
    GWN_LAUNCH_1D(pack_e2t_kernel, (i64)E * p->ldo, st, P_<float>(prm, p->i_e2w), sc + p->o_hb_w2t,
                  h3 ? sc + p->o_hb_w2t_lo : (float*)nullptr, O, E, p->ldo);
    GWN_LAUNCH_1D(transpose_kernel, (i64)E * Sk, st, P_<float>(prm, p->i_e1w), sc + p->o_hb_w1t, E, Sk,
                  h3 ? sc + p->o_hb_w1t_lo : (float*)nullptr);
    SkipWeights sw;
    memset(&sw, 0, sizeof(sw));
    for (int i = 0; i < nL; ++i) { sw.w[i] = P_<float>(prm, p->li[i].sw); sw.b[i] = P_<float>(prm, p->li[i].sb); }
    GWN_LAUNCH_1D(pack_skip_kernel, (i64)Sk * nL * D, st, sw, nL, D, Sk, (float*)nullptr, (float*)nullptr, sc + p->o_hb_wt,
                  h3 ? sc + p->o_hb_wt_lo : (float*)nullptr, (float*)nullptr);
    int hs;
    {  // (a) de1 = (dout . W2) * (e1 > 0)
      TcPosArgs t;
      memset(&t, 0, sizeof(t));
      const int nsd = (p->ldo + 31) / 32;   // K = out_dim columns of dout in 32-wide segments (the last one reads zeros past the row)
      for (int q = 0; q < nsd; ++q) t.seg[q] = TcPosSeg{dout, (int)PT, p->ldo, 32 * q, 0};
      t.nseg = nsd; t.nb = 1; t.rows_out = (int)PT; t.Wp = sc + p->o_hb_w2t; t.Wp_lo = h3 ? sc + p->o_hb_w2t_lo : nullptr;
      t.N = E <= TN ? E : TN; t.wstream = 1; t.N_total = E; t.w_k = p->ldo;
      t.out = de1; t.out_width = E; t.out_nblk = E / 32;
      RowDense ep;
      memset(&ep, 0, sizeof(ep));
      ep.gate = e1; ep.ldg = E;
      hs = launch_tcpos<0>(t, ep, st);
      if (hs > 0) return hs;
    }
    if (hs == 0) {  // (c) dskip = (de1 . W1) * (skip > 0)
      TcPosArgs t;
      memset(&t, 0, sizeof(t));
      for (int q = 0; q < E / 32; ++q) t.seg[q] = TcPosSeg{de1, (int)PT, E, 32 * q, 0};
      t.nseg = E / 32; t.nb = 1; t.rows_out = (int)PT; t.Wp = sc + p->o_hb_w1t; t.Wp_lo = h3 ? sc + p->o_hb_w1t_lo : nullptr;
      t.N = Sk <= TN ? Sk : TN; t.wstream = 1; t.N_total = Sk;
      t.out = dskip; t.out_width = Sk; t.out_nblk = Sk / 32;
      RowDense ep;
      memset(&ep, 0, sizeof(ep));
      ep.gate = skip; ep.ldg = Sk;
      hs = launch_tcpos<0>(t, ep, st);
      if (hs > 0) return hs;
      GWN_CHECK_ARG(hs == 0, "backward: head gradient (c) not eligible for the tcgen05 path");
    }
    if (hs == 0) {  // (f) gradient into the live columns of every g_i: dgh[i] = dskip . W_i
      TcPosArgs t;
      memset(&t, 0, sizeof(t));
      for (int q = 0; q < Sk / 32; ++q) t.seg[q] = TcPosSeg{dskip, (int)PT, Sk, 32 * q, 0};
      t.nseg = Sk / 32; t.nb = 1; t.rows_out = (int)PT; t.Wp = sc + p->o_hb_wt; t.Wp_lo = h3 ? sc + p->o_hb_wt_lo : nullptr;
      t.N = nL * D <= TN ? nL * D : TN; t.wstream = 1; t.N_total = nL * D;
      t.out = dgh; t.out_width = 32; t.out_nblk = nL * D / 32; t.out_blk_dim2 = 1;
      RowSeg ep;
      memset(&ep, 0, sizeof(ep));
      ep.out = dgh; ep.M = PT;
      hs = launch_tcpos<0>(t, ep, st);
      if (hs > 0) return hs;
      GWN_CHECK_ARG(hs == 0, "backward: head gradient (f) not eligible for the tcgen05 path");
      hb_tc = true;
    }
  }
#endif
  if (!hb_tc) {  // (a) de1 = (dout . W2) * (e1 > 0)
    LdRows la;
    memset(&la, 0, sizeof(la));
    la.p[0] = dout; la.set_wd(p->ldo);
    LdWN lb;
    memset(&lb, 0, sizeof(lb));
    lb.p[0] = P_<float>(prm, p->i_e2w); lb.set_wd(E); lb.ldw = E;
    EpRows ep;
    memset(&ep, 0, sizeof(ep));
    ep.y = de1; ep.ldy = E; ep.M = PT; ep.gate = e1;
    GemmShape sh{PT, E, O, 1, 1};
    GWN_TRY((launch_gemm<TBig>(la, lb, ep, sh, st)));
  }
  bool hw_tc = false;   // head weight gradients on the tcgen05 reduction kernel
#if !GWN_EMU
  auto head_wgrad_tc = [&](const TcRedSrc* ablk, int na, const TcRedSrc& bsrc, int Ntot, int n_valid, int nb_, int rows_,
                           float* const* blk, i64 sn, float* const* bias, int nbias) -> int {
    if (!tsc.partial || na > TR_MAXSRC) return -1;
    TcRedArgs t;
    memset(&t, 0, sizeof(t));
    t.mode = 0; t.na = na; t.x3 = tsc.x3;
    for (int j = 0; j < na; ++j) t.a[j] = ablk[j];
    t.b[0] = bsrc;
    t.ab = tsc.x3 ? 3 : 7;   // 3 real blocks + the all-ones block = ONE 128-row tile (4 + ones made two, 3/8 of the MMA rows zero)
    t.N = Ntot <= (tsc.x3 ? 128 : 256) ? Ntot : (tsc.x3 ? 128 : 256);
    t.N_total = Ntot; t.nb = nb_; t.rows = rows_; t.partial = tsc.partial; t.partial_floats = tsc.floats;
    TcRedResult r;
    int rs = launch_tcred(t, st, &r);
    if (rs != 0) return rs;
    tc::SlotGridOut f;
    memset(&f, 0, sizeof(f));
    for (int j = 0; j < na; ++j) f.blk[j] = blk[j];
    for (int q = 0; q < nbias; ++q) f.bias[q] = bias[q];
    f.nbias = nbias; f.na = na; f.ab = t.ab < na ? t.ab : na; f.n_ag = r.n_mg; f.n_bg = r.n_nt; f.Ntile = r.N; f.Ntot = Ntot;
    f.n_valid = n_valid; f.sn = sn; f.si = 1;
    const i64 nout = (i64)na * 32 * Ntot + Ntot;
    if (Ntot % 4 == 0 && r.N % 4 == 0 && r.slot_floats % 4 == 0 && (reinterpret_cast<uintptr_t>(tsc.partial) & 15) == 0)
      return launch_slot_reduce4(tsc.partial, r, nout, f, st);
    return launch_slot_reduce(tsc.partial, r, nout, f, st);
  };
  if (hb_tc && E % 32 == 0 && E / 32 <= TR_MAXSRC && Sk / 32 <= TR_MAXSRC && O <= 256) {
    int rs;
    {  // (b) dW2[o][e] = sum_p dout[p][o] e1[p][e], db2[o] = sum_p dout[p][o]
      TcRedSrc ab_[TR_MAXSRC];
      float* blk[TR_MAXSRC];
      for (int j = 0; j < E / 32; ++j) { ab_[j] = TcRedSrc{e1, (int)PT, E, 32 * j, 0, 0}; blk[j] = G(p->i_e2w) + 32 * j; }
      float* bias[1] = {G(p->i_e2b)};
      rs = head_wgrad_tc(ab_, E / 32, TcRedSrc{dout, (int)PT, p->ldo, 0, 0, 0}, round_up(O, 32), O, 1, (int)PT, blk, E, bias, 1);
      if (rs > 0) return rs;
    }
    if (rs == 0) {  // (d) dW1[e][sk] = sum_p de1[p][e] skip[p][sk], db1[e] = sum_p de1[p][e]
      TcRedSrc ab_[TR_MAXSRC];
      float* blk[TR_MAXSRC];
      for (int j = 0; j < Sk / 32; ++j) { ab_[j] = TcRedSrc{skip, (int)PT, Sk, 32 * j, 0, 0}; blk[j] = G(p->i_e1w) + 32 * j; }
      float* bias[1] = {G(p->i_e1b)};
      rs = head_wgrad_tc(ab_, Sk / 32, TcRedSrc{de1, (int)PT, E, 0, 0, 0}, E, E, 1, (int)PT, blk, Sk, bias, 1);
      if (rs > 0) return rs;
      GWN_CHECK_ARG(rs == 0, "backward: head weight gradient (d) not eligible for the tcgen05 path");
    }
    if (rs == 0) {  // (e) dWskip_i[sk][c] = sum_p dskip[p][sk] g_i[live p][c], dbskip_i[sk] = sum_p dskip[p][sk]
      TcRedSrc ab_[TR_MAXSRC];
      float* blk[TR_MAXSRC];
      float* bias[TR_MAXSRC];
      for (int i = 0; i < nL; ++i) {
        ab_[i] = TcRedSrc{ws + p->o_g[i], p->L[i] * N, 32, 0, (p->L[i] - p->T_out) * N, 0};
        blk[i] = G(p->li[i].sw);
        bias[i] = G(p->li[i].sb);
      }
      rs = head_wgrad_tc(ab_, nL, TcRedSrc{dskip, p->T_out * N, Sk, 0, 0, 0}, Sk, Sk, B, p->T_out * N, blk, D, bias, nL);
      if (rs > 0) return rs;
      GWN_CHECK_ARG(rs == 0, "backward: head weight gradient (e) not eligible for the tcgen05 path");
      hw_tc = true;
    }
  }
#endif
  if (!hw_tc) {  // (b) dW2, db2
    LdCols la;
    memset(&la, 0, sizeof(la));
    la.p[0] = dout; la.set_wd(p->ldo); la.nseg = 1;
    LdCols lb;
    memset(&lb, 0, sizeof(lb));
    lb.p[0] = e1; lb.set_wd(E); lb.nseg = 1; lb.ones = 1;
    EpWgrad ep;
    memset(&ep, 0, sizeof(ep));
    ep.dw[0] = G(p->i_e2w); ep.db[0] = G(p->i_e2b); ep.set_wd(E); ep.nseg = 1; ep.ldw = E; ep.nbias = 1;
    GemmShape sh{(i64)O, E + 1, (int)PT, pick_ksplit(O, E + 1, PT, TW32::BM, TW32::BN, kTargetBlocks), 1};
    GWN_TRY((launch_gemm<TW32>(la, lb, ep, sh, st)));
  }
  if (!hb_tc) {  // (c) dskip = (de1 . W1) * (skip > 0)
    LdRows la;
    memset(&la, 0, sizeof(la));
    la.p[0] = de1; la.set_wd(E);
    LdWN lb;
    memset(&lb, 0, sizeof(lb));
    lb.p[0] = P_<float>(prm, p->i_e1w); lb.set_wd(Sk); lb.ldw = Sk;
    EpRows ep;
    memset(&ep, 0, sizeof(ep));
    ep.y = dskip; ep.ldy = Sk; ep.M = PT; ep.gate = skip;
    GemmShape sh{PT, Sk, E, 1, 1};
    GWN_TRY((launch_gemm<TBig>(la, lb, ep, sh, st)));
  }
  if (!hw_tc) {  // (d) dW1, db1
    LdCols la;
    memset(&la, 0, sizeof(la));
    la.p[0] = de1; la.set_wd(E); la.nseg = 1;
    LdCols lb;
    memset(&lb, 0, sizeof(lb));
    lb.p[0] = skip; lb.set_wd(Sk); lb.nseg = 1; lb.ones = 1;
    EpWgrad ep;
    memset(&ep, 0, sizeof(ep));
    ep.dw[0] = G(p->i_e1w); ep.db[0] = G(p->i_e1b); ep.set_wd(Sk); ep.nseg = 1; ep.ldw = Sk; ep.nbias = 1;
    GemmShape sh{(i64)E, Sk + 1, (int)PT, pick_ksplit(E, Sk + 1, PT, TBig::BM, TBig::BN, kTargetBlocks), 1};
    GWN_TRY((launch_gemm<TBig>(la, lb, ep, sh, st)));
  }
  {  // (e) skip conv weight / bias gradients of every layer; (f) gradient into the live columns of every g_i
    LdCols la;
    memset(&la, 0, sizeof(la));
    la.p[0] = dskip; la.set_wd(Sk); la.nseg = 1;
    LdCols lb;
    memset(&lb, 0, sizeof(lb));
    EpWgrad ep;
    memset(&ep, 0, sizeof(ep));
    LdWN lw;
    memset(&lw, 0, sizeof(lw));
    for (int i = 0; i < nL; ++i) {
      lb.p[i] = ws + p->o_g[i];
      lb.rm[i] = make_remap(p->T_out, p->L[i], p->L[i] - p->T_out, N);
      ep.dw[i] = G(p->li[i].sw);
      ep.db[i] = G(p->li[i].sb);
      lw.p[i] = P_<float>(prm, p->li[i].sw);
    }
    lb.set_wd(D); lb.nseg = nL; lb.use_remap = 1; lb.ones = 1;
    ep.set_wd(D); ep.nseg = nL; ep.ldw = D; ep.nbias = nL;
    GemmShape sh{(i64)Sk, nL * D + 1, (int)PT, pick_ksplit(Sk, nL * D + 1, PT, TBig::BM, TBig::BN, kTargetBlocks), 1};
    if (!hw_tc) GWN_TRY((launch_gemm<TBig>(la, lb, ep, sh, st)));

    if (!hb_tc) {
      LdRows lr;
      memset(&lr, 0, sizeof(lr));
      lr.p[0] = dskip; lr.set_wd(Sk);
      lw.set_wd(D); lw.ldw = D;
      EpRows es;
      memset(&es, 0, sizeof(es));
      es.y = dgh; es.M = PT; es.set_seg(D);
      GemmShape sh2{PT, nL * D, Sk, 1, 1};
      GWN_TRY((launch_gemm<TBig>(lr, lw, es, sh2, st)));
    }
  }

  delete prof_hb;
  // ---- layers in reverse
  float* cur = sc + p->o_buf0;   // holds d(loss)/d(x_{i+1}) on entry of layer i (unused for the last layer)
  float* oth = sc + p->o_buf1;
  float* dg = sc + p->o_dg;
  const float* dA_X[TR_MAXSRC];
  const float* dA_T[TR_MAXSRC];
  int dA_slabs[TR_MAXSRC], dA_pairs = 0;
  const bool defer_w = !GWN_EMU && p->defer_wgrad && tcpos_ok(p) && tsc.partial != nullptr;
  struct WJob { int layer; const float* prev; const float* prev_ac; const float* dpre; const float* g; const float* dh; };
  WJob wj_tcn[64], wj_mlp[64];
  int n_wj_tcn = 0, n_wj_mlp = 0;
  for (int i = nL - 1; i >= 0; --i) {
    float* dpre = sc + (defer_w ? p->o_dpre_l[i] : p->o_dpre);
    float* dh_buf = sc + (defer_w ? p->o_dh_l[i] : p->o_dh);
    const bool live = i < nL - 1;   // the last layer's gcn/bn output is discarded (model.py:238)
    const i64 Pi = p->P(i);
    const float* prev = i == 0 ? ws + p->o_x0 : ws + p->o_u[i - 1];
    const float* prev_ac = i == 0 ? nullptr : ws + p->o_ac[i - 1];
    const float* g = ws + p->o_g[i];
    const float* dgh_i = dgh + (i64)i * PT * D;
    float* dsegs = sc + p->o_dsegs[i];
    const float* dgp;   // gradient wrt g_i
    if (live) {
      const DropoutSrc ldrop = layer_dropout(p, training, dmode, a->keep_masks, a->seed, i, a->seed_device);
      // materialise du * keep once instead of regenerating masks; deferred weight gradients need dh_i kept anyway
      const bool use_dh = ldrop.mode != GWN_DROPOUT_NONE || (defer_w && nL <= 64);
      {
      ProfScope prof("bn_bwd_apply", st, 4.0 * Pi * C * (use_dh ? 4.0 : 3.0), 0.0);
#if !GWN_EMU
      if (C >= 8 && C <= 512 && (C & (C - 1)) == 0 && (Pi * C) % 8 == 0) {
        // 8 resident blocks per SM, grid-stride: the per-block prologue (replica sums, channel constants) is paid once
        const i64 n8 = Pi * C / 8;
        const i64 blocks = std::min<i64>((n8 + 255) / 256, 148 * 8);
        GWN_CUDA(launch_kernel(bn_bwd_apply8_kernel, dim3((unsigned)blocks), dim3(256), 0, st, cur, (const float*)(ws + p->o_u[i]),
                               (const float*)(ws + p->o_ac[i]), (const float*)(ws + p->o_mr[i]),
                               reinterpret_cast<const double*>(sc + p->o_bsum + (i64)i * 4 * C * GWN_STAT_REPL), (double)Pi, training,
                               G(p->li[i].bnw), G(p->li[i].bnb), n8, C, use_dh ? dh_buf : (float*)nullptr, ldrop));
        count_launch();
      } else
#endif
      {
        GWN_LAUNCH_1D(bn_bwd_apply_kernel, Pi * C / 4, st, cur, ws + p->o_u[i], ws + p->o_ac[i], ws + p->o_mr[i],
                    reinterpret_cast<const double*>(sc + p->o_bsum + (i64)i * 4 * C * GWN_STAT_REPL), (double)Pi, training,
                    G(p->li[i].bnw), G(p->li[i].bnb), Pi, C, use_dh ? dh_buf : (float*)nullptr, ldrop);
      }

      }
      const float* segs[MAXSEG];
      for (int q = 0; q < p->nseg; ++q) segs[q] = g + (i64)q * Pi * D;
      MlpBwdArgs m;
      memset(&m, 0, sizeof(m));
      m.dh = use_dh ? dh_buf : cur;
      m.drop = make_dropout(GWN_DROPOUT_NONE, nullptr, 0, 0, 0.f);
      m.segs = segs; m.nseg = p->nseg; m.P = Pi; m.D = D; m.C_out = C;
      m.W = P_<float>(prm, c.gcn ? p->li[i].mw : p->li[i].rw);
      m.dsegs = dsegs;
      if (tcpos_ok(p)) {
        float* pk = const_cast<float*>(ws) + p->o_pack[i];   // W^T (+ remainders): packed by the forward pass
        m.WT = pk + p->pk_wt;
        if (x3(p)) m.WT_lo = pk + p->pk_wt_lo;
      }
      m.dW = G(c.gcn ? p->li[i].mw : p->li[i].rw);
      m.dbias = G(c.gcn ? p->li[i].mb : p->li[i].rb);
      if (defer_w && nL <= 64 && D == 32 && C == 32 && Pi < 2147483647LL) {   // weight gradient after the layer loop
        wj_mlp[n_wj_mlp++] = WJob{i, nullptr, nullptr, nullptr, g, m.dh};
        m.dW = nullptr;
        m.dbias = nullptr;
      }
      m.ts = tsc;
      GWN_TRY(mlp_backward(m, st));
      if (c.gcn) {
        GcnShape gs{B, p->L[i], N, D, C, p->S, c.order};
        float* dsup[MAXSUP];
        i64 ldds[MAXSUP];
        for (int s = 0; s < p->S; ++s) { dsup[s] = nullptr; ldds[s] = p->ld; }
        if (c.adaptive && !p->defer_dA) dsup[p->S - 1] = sc + p->o_dA;
        const bool ps_tc = c.per_sample_supports && tcpos_ok(p);
        if (!c.per_sample_supports || ps_tc) {
          TcSupports tb = tcB;
          tb.per_sample = ps_tc ? 1 : 0;
          tb.batch_stride = (i64)N * p->ld;
          GWN_TRY(gcn_hops_backward(gs, g, g + Pi * D, supB, dsegs, dg, dgh_i, p->T_out, dsup, ldds, st, &tb, &tsc));
        } else {   // fp32 tier; per-sample graphs have no support gradient (the supports are inputs, model.py:313-329)
          GcnShape g1{1, p->L[i], N, D, C, p->S, c.order};
          const i64 slab = (i64)p->L[i] * N * D, sup_sz = (i64)N * p->ld;
          for (int b = 0; b < B; ++b) {
            SupportView sb[MAXSUP];
            TcSupports tb = tcB;
            for (int s = 0; s < p->S; ++s) {
              sb[s] = support_padded(supB[s].p + (i64)b * sup_sz, p->ld);
              tb.S[s] = tcB.S[s] + (i64)b * sup_sz;
              if (tcB.Slo[s]) tb.Slo[s] = tcB.Slo[s] + (i64)b * sup_sz;
            }
            GWN_TRY(gcn_hops_backward(g1, g + b * slab, g + Pi * D + b * slab, sb, dsegs + b * slab, dg + b * slab,
                                      dgh_i + (i64)b * p->T_out * N * D, p->T_out, nullptr, ldds, st, &tb, &tsc, Pi * D));
          }
        }
        if (c.adaptive && p->defer_dA) {   // (hop input, chained gradient) pairs of the adaptive support, used after the loop
          const int sA = p->S - 1;
          for (int k = 1; k <= c.order; ++k) {
            dA_X[dA_pairs] = (k == 1) ? g : g + (i64)hop_index(gs, sA, k - 1) * Pi * D;
            dA_T[dA_pairs] = dsegs + (i64)hop_index(gs, sA, k) * Pi * D;
            dA_slabs[dA_pairs] = B * p->L[i];
            ++dA_pairs;
          }
        }
      } else {
        GWN_LAUNCH_1D(add_window_kernel, Pi * D, st, dg, (const float*)dsegs, dgh_i, B, p->L[i], N, D, p->T_out);
      }
      dgp = dg;
    } else {
      GWN_LAUNCH_1D(add_window_kernel, Pi * D, st, dg, (const float*)nullptr, dgh_i, B, p->L[i], N, D, p->T_out);
      dgp = dg;
    }
    // gated conv backward: recompute pre-activations -> dpre
    {
      ProfScope prof("gated_tcn_bwd_gate", st, 4.0 * ((double)B * p->Lin(i) * N * C + (double)Pi * 3 * D), 2.0 * Pi * 2 * C * 2 * D);
      LdRows la = tcn_rows(p, prev, prev_ac, i);
      LdWTcn lb{P_<float>(prm, p->li[i].fw), P_<float>(prm, p->li[i].gw), C};
      EpGateBwd ep{dpre, dgp, P_<float>(prm, p->li[i].fb), P_<float>(prm, p->li[i].gb), D};
      int pst = -1;
      if (tcpos_ok(p) && training) {   // packed weights (BN fold included) were written by the forward pass
        const float* pk = ws + p->o_pack[i];
        RowGateBwd eg;
        memset(&eg, 0, sizeof(eg));
        eg.dpre = dpre; eg.dg = dgp; eg.bf = pk + p->pk_bf; eg.bg = pk + p->pk_bg;
        TcPosArgs ta = tcn_tcpos_args(p, prev, pk + p->pk_wp, i);
        ta.out = dpre; ta.out_width = 2 * D; ta.out_nblk = 2;
        if (x3(p)) ta.Wp_lo = pk + p->pk_wp_lo;
        ta.addend[0] = TcPosSeg{dgp, p->L[i] * N, 32, 0, 0};
        pst = launch_tcpos<64>(ta, eg, st);
        if (pst > 0) return pst;
      }
      if (pst < 0 && pg_ok(p)) pst = launch_posgemm<TPG, 64>(tcn_arows(p, prev, prev_ac, i), lb, ep, Pi, 2 * D, st);
      if (pst > 0) return pst;
      if (pst < 0) {
        GemmShape sh{Pi, 2 * D, 2 * C, 1, 1};
        GWN_TRY((launch_gemm<TPos64>(la, lb, ep, sh, st)));
      }
    }
    const i64 Pin = (i64)B * p->Lin(i) * N;
    {  // input gradient (+ residual path, + BN-backward sums of the layer below)
      ProfScope prof("gated_tcn_dgrad", st, 4.0 * ((double)Pi * 2 * D + (live ? (double)Pi * C : 0.0) + (double)Pin * C * (i > 0 ? 2 : 1)),
                     2.0 * Pin * 4 * D * C);
      LdDpreTaps la{dpre, 2 * D, N, p->Lin(i), p->L[i], p->dil[i]};
      LdWTcnT lb{P_<float>(prm, p->li[i].fw), P_<float>(prm, p->li[i].gw), C, 2 * D};
      EpTcnDgrad<TPos32> ep;
      memset(&ep, 0, sizeof(ep));
      ep.dx = oth; ep.du = live ? cur : nullptr; ep.C = C; ep.N = N; ep.L_in = p->Lin(i); ep.L_out = p->L[i];
      if (i > 0) {
        ep.uprev = ws + p->o_u[i - 1];
        ep.mr = ws + p->o_mr[i - 1];
        ep.bsum = reinterpret_cast<double*>(sc + p->o_bsum + (i64)(i - 1) * 4 * C * GWN_STAT_REPL);
      }
      int pst = -1;
      if (tcpos_ok(p)) {
        float* pk = const_cast<float*>(ws) + p->o_pack[i];   // dgrad weights: packed by the forward pass
        RowTcnDgrad eg;
        memset(&eg, 0, sizeof(eg));
        eg.dx = ep.dx; eg.du = ep.du; eg.N = N; eg.L_in = ep.L_in; eg.L_out = ep.L_out;
        eg.uprev = ep.uprev; eg.mr = ep.mr; eg.bsum = ep.bsum;
        TcPosArgs t;
        memset(&t, 0, sizeof(t));
        for (int q = 0; q < 4; ++q) t.seg[q] = TcPosSeg{dpre, p->L[i] * N, 2 * D, (q & 1) * 32, -(q >> 1) * p->dil[i] * N};
        t.nseg = 4; t.nb = B; t.rows_out = p->Lin(i) * N; t.Wp = pk + p->pk_wd; t.N = 32;
        t.out = eg.dx; t.out_width = 32; t.out_nblk = 1;
        if (x3(p)) t.Wp_lo = pk + p->pk_wd_lo;
        if (eg.du) t.addend[0] = TcPosSeg{eg.du, p->L[i] * N, 32, 0, -(p->Lin(i) - p->L[i]) * N};
        if (eg.uprev) t.addend[1] = TcPosSeg{eg.uprev, p->Lin(i) * N, 32, 0, 0};
        pst = launch_tcpos<32>(t, eg, st);
        if (pst > 0) return pst;
      }
      if (pst < 0 && pg_ok(p)) {
        EpTcnDgrad<TPG> eg;
        memset(&eg, 0, sizeof(eg));
        eg.dx = ep.dx; eg.du = ep.du; eg.C = C; eg.N = N; eg.L_in = ep.L_in; eg.L_out = ep.L_out;
        eg.uprev = ep.uprev; eg.mr = ep.mr; eg.bsum = ep.bsum;
        ARows ar;
        memset(&ar, 0, sizeof(ar));
        for (int q = 0; q < 4; ++q) { ar.P[q] = dpre + (q & 1) * PG_WD; ar.rmap[q] = (unsigned char)(q >> 1); }
        ar.rm[0] = rowmap_shift(p->Lin(i), p->L[i], 0, N);
        ar.rm[1] = rowmap_shift(p->Lin(i), p->L[i], -p->dil[i], N);
        ar.nseg = 4; ar.rs = 2 * D;
        pst = launch_posgemm<TPG, 32>(ar, lb, eg, Pin, C, st);
        if (pst > 0) return pst;
      }
      if (pst < 0) {
        GemmShape sh{Pin, C, 4 * D, 1, 1};
        GWN_TRY((launch_gemm<TPos32>(la, lb, ep, sh, st)));
      }
    }
    bool wgrad_done = false;
    if (defer_w && nL <= 64) {
      wj_tcn[n_wj_tcn++] = WJob{i, prev, prev_ac, dpre, nullptr, nullptr};
      wgrad_done = true;
    }
    ProfScope prof_w("gated_tcn_wgrad", st, wgrad_done ? 0.0 : 4.0 * ((double)Pin * C + (double)Pi * 2 * D),
                     wgrad_done ? 0.0 : 2.0 * Pi * 2 * D * (2.0 * C + 1));
    if (!wgrad_done && tcpos_ok(p) && tsc.partial) {   // tcgen05 + TMA reduction over all positions; the slot reduction folds the
#if !GWN_EMU                            // BatchNorm affine of the layer below and scatters to the four gradients
      TcRedArgs t;
      memset(&t, 0, sizeof(t));
      t.mode = 0; t.na = 2; t.x3 = tsc.x3;
      t.a[0] = TcRedSrc{prev, p->Lin(i) * N, 32, 0, 0, 0};
      t.a[1] = TcRedSrc{prev, p->Lin(i) * N, 32, 0, p->dil[i] * N, 0};
      t.b[0] = TcRedSrc{dpre, p->L[i] * N, 2 * D, 0, 0, 0};
      t.N = 2 * D; t.nb = B; t.rows = p->L[i] * N; t.partial = tsc.partial; t.partial_floats = tsc.floats;
      TcRedResult rr;
      int rst = launch_tcred(t, st, &rr);
      if (rst > 0) return rst;
      if (rst == 0) {
        tc::SlotTcnOut f{prev_ac, G(p->li[i].fw), G(p->li[i].gw), G(p->li[i].fb), G(p->li[i].gb), C, D};
        GWN_TRY(launch_slot_reduce(tsc.partial, rr, (i64)2 * C * 2 * D + 2 * D, f, st));
        wgrad_done = true;
      }
#endif
    }
    if (!wgrad_done) {  // filter / gate weight and bias gradients
      LdCols la;
      memset(&la, 0, sizeof(la));
      la.p[0] = dpre; la.set_wd(2 * D); la.nseg = 1;
      LdCols lb;
      memset(&lb, 0, sizeof(lb));
      lb.p[0] = prev; lb.p[1] = prev;
      lb.rm[0] = make_remap(p->L[i], p->Lin(i), 0, N);
      lb.rm[1] = make_remap(p->L[i], p->Lin(i), p->dil[i], N);
      lb.set_wd(C); lb.nseg = 2; lb.use_remap = 1; lb.ones = 1; lb.ac = prev_ac;
      EpWgradTcn ep{G(p->li[i].fw), G(p->li[i].gw), G(p->li[i].fb), G(p->li[i].gb), C};
      GWN_CHECK_ARG(Pi < 2147483647LL, "backward: too many positions");
      GemmShape sh{(i64)2 * D, 2 * C + 1, (int)Pi, pick_ksplit(2 * D, 2 * C + 1, Pi, TW64::BM, TW64::BN, kTargetBlocks), 1};
      GWN_TRY((launch_gemm<TW64>(la, lb, ep, sh, st)));
    }
    std::swap(cur, oth);
  }
#if !GWN_EMU
  // ---- deferred weight gradients: all layers' gated-conv (then all layers' mlp) reductions as jobs of one launch each
  for (int kind = 0; kind < 2; ++kind) {
    const WJob* wj = kind == 0 ? wj_tcn : wj_mlp;
    const int nj = kind == 0 ? n_wj_tcn : n_wj_mlp;
    if (nj == 0) continue;
    double bytes = 0, flops = 0;
    for (int q = 0; q < nj; ++q) {
      const int i = wj[q].layer;
      const double Pi = (double)p->P(i), Pin = (double)B * p->Lin(i) * N;
      if (kind == 0) { bytes += 4.0 * (Pin * C + Pi * 2 * D); flops += 2.0 * Pi * 2 * D * (2.0 * C + 1); }
      else { bytes += 4.0 * Pi * ((double)p->nseg * D + C); flops += 2.0 * Pi * (p->nseg * D + 1.0) * C; }
    }
    ProfScope prof(kind == 0 ? "gated_tcn_wgrad" : "gcn_mlp_wgrad", st, bytes, flops);
    for (int q0 = 0; q0 < nj; q0 += TR_MAXJOBS) {
      const int nq = std::min(TR_MAXJOBS, nj - q0);
      if (nq == 1) {   // a lone job: the single-reduction launch
        const WJob& w = wj[q0];
        const int i = w.layer;
        TcRedArgs t;
        memset(&t, 0, sizeof(t));
        t.mode = 0; t.x3 = tsc.x3; t.partial = tsc.partial; t.partial_floats = tsc.floats;
        TcRedResult rr;
        if (kind == 0) {
          t.na = 2;
          t.a[0] = TcRedSrc{w.prev, p->Lin(i) * N, 32, 0, 0, 0};
          t.a[1] = TcRedSrc{w.prev, p->Lin(i) * N, 32, 0, p->dil[i] * N, 0};
          t.b[0] = TcRedSrc{w.dpre, p->L[i] * N, 2 * D, 0, 0, 0};
          t.N = 2 * D; t.nb = B; t.rows = p->L[i] * N;
          int rst = launch_tcred(t, st, &rr);
          GWN_CHECK_ARG(rst <= 0, "backward: deferred gated-conv weight gradient failed");
          GWN_CHECK_ARG(rst == 0, "backward: deferred gated-conv weight gradient not eligible for the tcgen05 path");
          tc::SlotTcnOut f{w.prev_ac, G(p->li[i].fw), G(p->li[i].gw), G(p->li[i].fb), G(p->li[i].gb), C, D};
          GWN_TRY(launch_slot_reduce(tsc.partial, rr, (i64)2 * C * 2 * D + 2 * D, f, st));
        } else {
          t.na = p->nseg;
          for (int k = 0; k < p->nseg; ++k) t.a[k] = TcRedSrc{w.g + (i64)k * p->P(i) * D, (int)p->P(i), 32, 0, 0, 0};
          t.b[0] = TcRedSrc{w.dh, (int)p->P(i), 32, 0, 0, 0};
          t.N = 32; t.nb = 1; t.rows = (int)p->P(i);
          int rst = launch_tcred(t, st, &rr);
          GWN_CHECK_ARG(rst == 0, "backward: deferred mlp weight gradient not eligible for the tcgen05 path");
          tc::SlotMlpOut f{G(c.gcn ? p->li[i].mw : p->li[i].rw), G(c.gcn ? p->li[i].mb : p->li[i].rb), p->nseg * D, p->nseg};
          GWN_TRY(launch_slot_reduce(tsc.partial, rr, (i64)p->nseg * 32 * 32 + 32, f, st));
        }
        continue;
      }
      TcRedJobsArgs ja;
      memset(&ja, 0, sizeof(ja));
      ja.njobs = nq; ja.x3 = tsc.x3; ja.partial = tsc.partial; ja.partial_floats = tsc.floats;
      if (kind == 0) { ja.na = 2; ja.N = 2 * D; ja.seg[0] = ja.seg[1] = 0; }
      else { ja.na = p->nseg; ja.N = 32; for (int k = 0; k < p->nseg; ++k) ja.seg[k] = k; }
      for (int q = 0; q < nq; ++q) {
        const WJob& w = wj[q0 + q];
        const int i = w.layer;
        TcRedJob& j = ja.job[q];
        if (kind == 0) {
          j.a_src = w.prev; j.a_rows_src = p->Lin(i) * N; j.nseg_src = 1; j.a_seg_stride = 0;
          j.b_src = w.dpre; j.b_width = 2 * D; j.nb = B; j.rows = p->L[i] * N;
          j.rshift[0] = 0; j.rshift[1] = p->dil[i] * N;
        } else {
          j.a_src = w.g; j.a_rows_src = (int)p->P(i); j.nseg_src = p->nseg; j.a_seg_stride = p->P(i) * D;
          j.b_src = w.dh; j.b_width = 32; j.nb = 1; j.rows = (int)p->P(i);
        }
      }
      TcRedResult rr;
      int rst = launch_tcred_jobs(ja, st, &rr);
      if (rst > 0) return rst;
      GWN_CHECK_ARG(rst == 0, "backward: deferred weight gradients not eligible for the tcgen05 path");
      if (kind == 0) {
        tc::SlotJobs<tc::SlotTcnOut> f;
        memset(&f, 0, sizeof(f));
        for (int q = 0; q < nq; ++q) {
          const WJob& w = wj[q0 + q];
          const int i = w.layer;
          f.f[q] = tc::SlotTcnOut{w.prev_ac, G(p->li[i].fw), G(p->li[i].gw), G(p->li[i].fb), G(p->li[i].gb), C, D};
        }
        for (int q = 0; q <= nq; ++q) f.cta0[q] = rr.job_cta0[q];
        f.nout_job = (i64)2 * C * 2 * D + 2 * D;
        GWN_TRY(launch_slot_reduce(tsc.partial, rr, f.nout_job * nq, f, st));
      } else {
        tc::SlotJobs<tc::SlotMlpOut> f;
        memset(&f, 0, sizeof(f));
        for (int q = 0; q < nq; ++q) {
          const int i = wj[q0 + q].layer;
          f.f[q] = tc::SlotMlpOut{G(c.gcn ? p->li[i].mw : p->li[i].rw), G(c.gcn ? p->li[i].mb : p->li[i].rb), p->nseg * D, p->nseg};
        }
        for (int q = 0; q <= nq; ++q) f.cta0[q] = rr.job_cta0[q];
        f.nout_job = (i64)p->nseg * 32 * 32 + 32;
        GWN_TRY(launch_slot_reduce(tsc.partial, rr, f.nout_job * nq, f, st));
      }
    }
  }
#endif
  // ---- adaptive-support gradient of ALL layers in one tcgen05 reduction (dA sums over 7 layers x 2 hops)
  if (dA_pairs > 0) {
    double slabs = 0;
    for (int q = 0; q < dA_pairs; ++q) slabs += dA_slabs[q];
    ProfScope prof("nconv_bwd_dA", st, 4.0 * slabs * N * D * 2.0, 2.0 * slabs * N * D * N);
    int dst = support_grad_tc(dA_X, dA_T, dA_slabs, dA_pairs, sc + p->o_dA, p->ld, N, D, tsc, st);
    GWN_CHECK_ARG(dst >= 0, "backward: deferred support gradient not eligible for the tcgen05 path");
    if (dst > 0) return dst;
  }
  // ---- start conv backward: cur = d(loss)/d(x0)
  {
    ProfScope prof("start_conv_bwd", st, 4.0 * p->P0() * (C + c.in_dim), 2.0 * p->P0() * C * (c.in_dim + 1));
    LdCols la;
    memset(&la, 0, sizeof(la));
    la.p[0] = cur; la.set_wd(C); la.nseg = 1;
    LdInputCols lb;
    lb.in = a->input;
    lb.sb = a->input_strides[0]; lb.sf = a->input_strides[1]; lb.sn = a->input_strides[2]; lb.st = a->input_strides[3];
    lb.F = c.in_dim; lb.N = N; lb.L0 = p->L0; lb.pad = p->pad; lb.ones = 1;
    EpWgrad ep;
    memset(&ep, 0, sizeof(ep));
    ep.dw[0] = G(p->i_startw); ep.db[0] = G(p->i_startb); ep.set_wd(c.in_dim); ep.nseg = 1; ep.ldw = c.in_dim; ep.nbias = 1;
    GWN_CHECK_ARG(p->P0() < 2147483647LL, "backward: too many input positions");
    bool sw_done = false;
#if !GWN_EMU
    if (C == 32 && c.in_dim <= START_MAXF && p->P0() < 2147483647LL) {   // warp-per-position reduction (elementwise.cuh)
      Strides4 is;
      for (int k = 0; k < 4; ++k) is.s[k] = a->input_strides[k];
      const int nbt = B * p->L0;
      GWN_CUDA(launch_kernel(start_wgrad32b_kernel, dim3((unsigned)std::min(nbt, 148 * 8)), dim3(256), 0, st, (const float*)cur,
                             a->input, is, G(p->i_startw), G(p->i_startb), c.in_dim, N, p->L0, p->pad, nbt));
      count_launch();
      sw_done = true;
    }
#endif
    if (!sw_done) {
      GemmShape sh{(i64)C, c.in_dim + 1, (int)p->P0(), pick_ksplit(C, c.in_dim + 1, p->P0(), TW32::BM, TW32::BN, kTargetBlocks), 1};
      GWN_TRY((launch_gemm<TW32>(la, lb, ep, sh, st)));
    }
    if (a->grad_input)
      GWN_LAUNCH_1D(start_dgrad_kernel, (i64)B * c.in_dim * N * c.seq_len, st, (const float*)cur, P_<float>(prm, p->i_startw),
                    a->grad_input, B, c.in_dim, N, c.seq_len, p->L0, p->pad, C);
  }
  // ---- adaptive adjacency backward (model.py:187)
  if (c.adaptive) {
    const float* Ap = ws + p->o_sup + (i64)(p->S - 1) * N * p->ld;
    GWN_LAUNCH_WARP_ROWS(adp_bwd_rows_kernel, N, st, (const float*)(sc + p->o_dA), Ap, P_<float>(prm, p->i_nv1),
                         P_<float>(prm, p->i_nv2), c.apt_rank, sc + p->o_dR, G(p->i_nv1), N, p->ld);
#if !GWN_EMU
    if (c.apt_rank <= 16) {   // G(nv2) was zeroed with the flat gradient buffer: the slices add into it
      const int slices = 32, vchunk = (N + slices - 1) / slices;
      GWN_CUDA(launch_kernel(adp_bwd_cols_split_kernel, dim3((unsigned)((N + 63) / 64), (unsigned)slices), dim3(64), 0, st,
                             (const float*)(sc + p->o_dR), P_<float>(prm, p->i_nv1), c.apt_rank, G(p->i_nv2), N, p->ld, vchunk));
      count_launch();
    } else
#endif
    GWN_LAUNCH_1D(adp_bwd_cols_kernel, (i64)c.apt_rank * N, st, (const float*)(sc + p->o_dR), P_<float>(prm, p->i_nv1),
                  c.apt_rank, G(p->i_nv2), N, p->ld);
  }
  return 0;
}

}  // namespace gwn

// =============================================================================== C ABI
using namespace gwn;

extern "C" {

const char* gwn_last_error(void) { return g_err; }
int gwn_abi_version(void) { return GWN_ABI_VERSION; }
long long gwn_launch_count(int reset) {
  long long v = g_launches.load(std::memory_order_relaxed);
  if (reset) g_launches.store(0, std::memory_order_relaxed);
  return v;
}

int gwn_profile_begin(void) {
#if !GWN_EMU
  GWN_TRY(require_device());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_prof_on.store(1);
#endif
  return 0;
}

int gwn_profile_end(char* buf, int len) {
  GWN_CHECK_ARG(buf && len > 2, "profile_end: bad buffer");
  std::string s = "[";
#if !GWN_EMU
  g_prof_on.store(0);
  GWN_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Agg { long long calls = 0; double ms = 0, bytes = 0, flops = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { ms = 0.f; cudaGetLastError(); }
    if (!agg.count(r.tag)) order.push_back(r.tag);
    Agg& a = agg[r.tag];
    a.calls += 1; a.ms += ms; a.bytes += r.bytes; a.flops += r.flops;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  char t[256];
  for (size_t i = 0; i < order.size(); ++i) {
    const Agg& a = agg[order[i]];
    snprintf(t, sizeof(t), "%s{\"op\":\"%s\",\"calls\":%lld,\"ms\":%.6f,\"bytes\":%.0f,\"flops\":%.0f}", i ? "," : "",
             order[i].c_str(), a.calls, a.ms, a.bytes, a.flops);
    s += t;
  }
#endif
  s += "]";
  GWN_CHECK_ARG((int)s.size() < len, "profile_end: buffer too small (%d needed)", (int)s.size() + 1);
  snprintf(buf, len, "%s", s.c_str());
  return 0;
}

int gwn_device_info(int* n_devices, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor) {
#if GWN_EMU
  if (n_devices) *n_devices = 0;
  if (name && name_len > 0) snprintf(name, name_len, "host-emulation (tests only)");
  if (sm_count) *sm_count = 0;
  if (cc_major) *cc_major = 0;
  if (cc_minor) *cc_minor = 0;
  return 0;
#else
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { n = 0; cudaGetLastError(); }
  if (n_devices) *n_devices = n;
  if (n <= 0) {
    if (name && name_len > 0) name[0] = 0;
    return 0;
  }
  int dev = 0;
  GWN_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp pr;
  GWN_CUDA(cudaGetDeviceProperties(&pr, dev));
  if (name && name_len > 0) snprintf(name, name_len, "%s", pr.name);
  if (sm_count) *sm_count = pr.multiProcessorCount;
  if (cc_major) *cc_major = pr.major;
  if (cc_minor) *cc_minor = pr.minor;
  return 0;
#endif
}

int gwn_permute4d(const float* src, const int64_t src_strides[4], float* dst, const int64_t dst_strides[4],
                  const int64_t sizes[4], void* stream) {
  GWN_TRY(require_device());
  GWN_CHECK_ARG(src && dst, "permute4d: null pointer");
  return permute4d(src, src_strides, dst, dst_strides, sizes, (cudaStream_t)stream);
}

int gwn_tc_error_flag(int reset) { return tc_error_flag(reset); }
void gwn_tc_debug_buffer(float* p) { tc_set_debug_buffer(p); }
void gwn_tc_debug_mode(int m) { tc_set_debug_mode(m); }

int gwn_node_contract(const float* x, const float* S, int64_t ld, float* y, int B, int L, int V, int C, int precision,
                      void* stream) {
  GWN_TRY(require_device());
  GWN_CHECK_ARG(x && S && y, "node_contract: null pointer");
  MathScope math_scope(math_of(precision));
  SupportView sv = SupportView{S, 1, ld, 0};   // op(k, m) = S[m*ld + k]
  TcSupports tcs;
  memset(&tcs, 0, sizeof(tcs));
  tcs.S[0] = S; tcs.ld = (int)ld; tcs.precision = precision;
  const float* X[1] = {x};
  float* Y[1] = {y};
  return node_gemm(&sv, 1, false, X, Y, nullptr, nullptr, B, L, 0, V, C, (cudaStream_t)stream, &tcs);
}

int gwn_split_lo(const float* src, float* lo, int64_t n, void* stream) {
  GWN_TRY(require_device());
  GWN_CHECK_ARG(src && lo && n >= 0, "split_lo: bad argument");
  if (n == 0) return 0;
  GWN_LAUNCH_1D(split_lo_kernel, (i64)n, (cudaStream_t)stream, src, lo, (i64)n);
  return 0;
}

int gwn_node_contract_x3(const float* x, const float* S, const float* S_lo, int64_t ld, float* y, int B, int L, int V, int C,
                         void* stream) {
  GWN_TRY(require_device());
  GWN_CHECK_ARG(x && S && S_lo && y, "node_contract_x3: null pointer");
  MathScope math_scope(math_of(GWN_PREC_FP32X3));
  SupportView sv = SupportView{S, 1, ld, 0};
  TcSupports tcs;
  memset(&tcs, 0, sizeof(tcs));
  tcs.S[0] = S; tcs.Slo[0] = S_lo; tcs.ld = (int)ld; tcs.precision = GWN_PREC_FP32X3;
  const float* X[1] = {x};
  float* Y[1] = {y};
  return node_gemm(&sv, 1, false, X, Y, nullptr, nullptr, B, L, 0, V, C, (cudaStream_t)stream, &tcs);
}

// ---- tensor-core tiers of the stand-alone operators: operand preparation in the caller's workspace
namespace gwn {
static bool tc_tier(int precision) { return precision == GWN_PREC_TF32 || precision == GWN_PREC_FP32X3; }
static i64 op_part_floats(int V) {   // partial-result slots of the tcgen05 reductions (as in build_plan)
  const i64 ot = (i64)((V + 255) / 256) * ((V + 255) / 256);
  return std::max<i64>(160, ot) * 512 * 128;
}
// Packed supports of one operator call: [set][support] matrices A (ld-padded) and A^T, then their 3xTF32 remainders.
struct OpSupports {
  float *Ap, *ATp, *lo;   // lo = remainders of [Ap | ATp] (fp32x3) or nullptr
  i64 mat, span;          // floats per matrix, floats of Ap + ATp
  int ld;
};
static i64 op_support_floats(int nsets, int nsup, int V, int precision) {
  const i64 mat = align_up((i64)V * round_up(V, 4));
  return (i64)nsets * nsup * mat * 2 * (precision == GWN_PREC_FP32X3 ? 2 : 1);
}
// supports[s]: [nsets][V][V] with sample stride lds_b[s] (ignored for nsets == 1) and row stride lds[s]
static int op_pack_supports(float* ws, const float* const* supports, const int64_t* lds_b, const int64_t* lds, int nsets, int nsup,
                            int V, int precision, cudaStream_t st, OpSupports* o) {
  o->ld = round_up(V, 4);
  o->mat = align_up((i64)V * o->ld);
  o->span = (i64)nsets * nsup * o->mat * 2;
  o->Ap = ws;
  o->ATp = ws + (i64)nsets * nsup * o->mat;
  o->lo = precision == GWN_PREC_FP32X3 ? ws + o->span : nullptr;
  for (int s = 0; s < nsup; ++s)
    for (int b = 0; b < nsets; ++b) {
      const float* A = supports[s] + (nsets > 1 ? (i64)b * lds_b[s] : 0);
      const i64 off = ((i64)s * nsets + b) * o->mat;
      GWN_LAUNCH_1D(support_pack_kernel, (i64)V * o->ld, st, A, (i64)lds[s], (i64)1, o->Ap + off, o->ATp + off, V, o->ld);
    }
  if (o->lo) GWN_LAUNCH_1D(split_lo_kernel, o->span, st, (const float*)ws, o->lo, o->span);
  return 0;
}
// forward = true: y[w] = sum_v A[v,w] x[v] (K-contiguous rows are those of A^T); false: dx[v] = sum_w A[v,w] dy[w]
static TcSupports op_tc_supports(const OpSupports& o, int nsets, int nsup, bool forward, int precision) {
  TcSupports t;
  memset(&t, 0, sizeof(t));
  for (int s = 0; s < nsup; ++s) {
    const i64 off = (i64)s * nsets * o.mat;
    t.S[s] = (forward ? o.ATp : o.Ap) + off;
    if (o.lo) t.Slo[s] = o.lo + ((forward ? o.ATp : o.Ap) - o.Ap) + off;
  }
  t.ld = o.ld; t.precision = precision; t.per_sample = nsets > 1 ? 1 : 0; t.batch_stride = o.mat;
  return t;
}
}  // namespace gwn

size_t gwn_nconv_workspace_floats(int n_sets, int V, int precision, int support_grad) {
  if (!tc_tier(precision) || V <= 0 || n_sets <= 0) return 0;
  return (size_t)(op_support_floats(n_sets, 1, V, precision) + (support_grad ? op_part_floats(V) : 0));
}

static int nconv_tc_args_ok(const char* what, int C, int precision, const void* workspace) {
  GWN_CHECK_ARG(precision == GWN_PREC_FP32 || tc_tier(precision), "%s: precision %d not available in this build", what, precision);
  if (!tc_tier(precision)) return 0;
  GWN_CHECK_ARG(C == 32, "%s: the tcgen05 tiers need 32 channels per node row (got %d); use GWN_PREC_FP32", what, C);
  GWN_CHECK_ARG(workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                "%s: the tcgen05 tiers need a 16-byte aligned workspace of gwn_nconv_workspace_floats() floats", what);
  return 0;
}

int gwn_nconv_fwd(const float* x, const float* A, int64_t lda, float* y, int B, int L, int V, int C, int precision,
                  void* workspace, void* stream) {
  GWN_TRY(require_device());
  GWN_CHECK_ARG(x && A && y, "nconv_fwd: null pointer");
  GWN_TRY(nconv_tc_args_ok("nconv_fwd", C, precision, workspace));
  cudaStream_t st = (cudaStream_t)stream;
  SupportView sv = support_fwd(A, lda, 1);
  const float* X[1] = {x};
  float* Y[1] = {y};
  if (tc_tier(precision)) {
    MathScope math_scope(math_of(precision));
    OpSupports os;
    const int64_t zero = 0;
    GWN_TRY(op_pack_supports(reinterpret_cast<float*>(workspace), &A, &zero, &lda, 1, 1, V, precision, st, &os));
    TcSupports tcs = op_tc_supports(os, 1, 1, true, precision);
    return node_gemm(&sv, 1, false, X, Y, nullptr, nullptr, B, L, 0, V, C, st, &tcs);
  }
  return node_gemm(&sv, 1, false, X, Y, nullptr, nullptr, B, L, 0, V, C, st);
}

int gwn_nconv_bwd(const float* dy, const float* x, const float* A, int64_t lda, float* dx, float* dA, int64_t ldda, int B,
                  int L, int V, int C, int precision, void* workspace, void* stream) {
  GWN_TRY(require_device());
  GWN_CHECK_ARG(dy && A, "nconv_bwd: null pointer");
  GWN_TRY(nconv_tc_args_ok("nconv_bwd", C, precision, workspace));
  GWN_CHECK_ARG(!dA || x, "nconv_bwd: x needed for dA");
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_tier(precision)) {
    MathScope math_scope(math_of(precision));
    float* ws = reinterpret_cast<float*>(workspace);
    OpSupports os;
    const int64_t zero = 0;
    GWN_TRY(op_pack_supports(ws, &A, &zero, &lda, 1, 1, V, precision, st, &os));
    if (dx) {
      TcSupports tcs = op_tc_supports(os, 1, 1, false, precision);
      SupportView sv = support_bwd(A, lda, 1);
      const float* X[1] = {dy};
      float* Y[1] = {dx};
      GWN_TRY(node_gemm(&sv, 1, false, X, Y, nullptr, nullptr, B, L, 0, V, C, st, &tcs));
    }
    if (dA) {   // caller sized the workspace with support_grad = 1
      TcScratch ts{ws + op_support_floats(1, 1, V, precision), op_part_floats(V), precision == GWN_PREC_FP32X3 ? 1 : 0};
      const float* Xp[1] = {x};
      const float* Yp[1] = {dy};
      GWN_TRY(support_grad_gemm(Xp, Yp, 1, dA, ldda, B, L, V, C, st, &ts));
    }
    return 0;
  }
  if (dx) {
    SupportView sv = support_bwd(A, lda, 1);
    const float* X[1] = {dy};
    float* Y[1] = {dx};
    GWN_TRY(node_gemm(&sv, 1, false, X, Y, nullptr, nullptr, B, L, 0, V, C, st));
  }
  if (dA) {
    const float* Xp[1] = {x};
    const float* Yp[1] = {dy};
    GWN_TRY(support_grad_gemm(Xp, Yp, 1, dA, ldda, B, L, V, C, st));
  }
  return 0;
}

int gwn_linear_fwd(const float* x, const float* W, const float* bias, float* y, int64_t positions, int c_in, int c_out,
                   void* stream) {
  GWN_TRY(require_device());
  GWN_CHECK_ARG(x && W && bias && y, "linear_fwd: null pointer");
  GWN_CHECK_ARG(c_in % 4 == 0, "linear_fwd: c_in must be a multiple of 4");
  const float* segs[1] = {x};
  MlpFwdArgs m;
  memset(&m, 0, sizeof(m));
  m.segs = segs; m.nseg = 1; m.P = positions; m.D = c_in; m.C_out = c_out; m.W = W; m.bias = bias;
  m.drop = make_dropout(GWN_DROPOUT_NONE, nullptr, 0, 0, 0.f);
  m.y = y;
  return mlp_forward(m, (cudaStream_t)stream);
}

int gwn_linear_bwd(const float* dy, const float* x, const float* W, float* dx, float* dW, float* dbias, int64_t positions,
                   int c_in, int c_out, void* stream) {
  GWN_TRY(require_device());
  GWN_CHECK_ARG(dy && x && W, "linear_bwd: null pointer");
  GWN_CHECK_ARG((dW == nullptr) == (dbias == nullptr), "linear_bwd: dW and dbias must be given together");
  const float* segs[1] = {x};
  MlpBwdArgs m;
  memset(&m, 0, sizeof(m));
  m.dh = dy; m.drop = make_dropout(GWN_DROPOUT_NONE, nullptr, 0, 0, 0.f);
  m.segs = segs; m.nseg = 1; m.P = positions; m.D = c_in; m.C_out = c_out; m.W = W;
  m.dsegs = dx; m.dW = dW; m.dbias = dbias;
  if (dW) {
    GWN_TRY(dev_memset(dW, 0, sizeof(float) * (size_t)c_in * c_out, (cudaStream_t)stream));
    GWN_TRY(dev_memset(dbias, 0, sizeof(float) * (size_t)c_out, (cudaStream_t)stream));
  }
  return mlp_backward(m, (cudaStream_t)stream);
}

static int gcn_check(const gwn_gcn_desc* d) {
  GWN_CHECK_ARG(d != nullptr, "gcn: null descriptor");
  GWN_CHECK_ARG(d->B >= 1 && d->L >= 1 && d->V >= 1 && d->C >= 4 && d->C % 4 == 0 && d->c_out >= 1, "gcn: bad dims");
  GWN_CHECK_ARG(d->c_out % 4 == 0, "gcn: c_out must be a multiple of 4");
  GWN_CHECK_ARG(d->n_supports >= 1 && d->n_supports <= MAXSUP && d->order >= 1 && d->order <= MAXSUP &&
                    1 + d->n_supports * d->order <= MAXSEG,
                "gcn: bad support count / order");
  GWN_CHECK_ARG(d->precision == GWN_PREC_FP32 || tc_tier(d->precision), "gcn: precision %d not available in this build", d->precision);
  GWN_CHECK_ARG(!tc_tier(d->precision) || (d->C == 32 && d->c_out == 32 && d->n_supports <= TC_MAXSUP && 1 + d->n_supports * d->order <= 7),
                "gcn: the tcgen05 tiers need c_in = c_out = 32 and at most 7 concatenated segments; use GWN_PREC_FP32");
  GWN_CHECK_ARG(d->dropout_p >= 0.f && d->dropout_p < 1.f, "gcn: dropout must be in [0,1)");
  return 0;
}

// Workspace layout of the gcn operators in the tensor-core tiers (floats):
//   [packed supports (+ remainders)] [W_lo] [W^T] [W^T_lo] [dh * keep-mask (backward)] [reduction slots (backward)]
namespace gwn {
struct GcnOpWs {
  i64 o_sup, o_wlo, o_wt, o_wtlo, o_dh, o_part, total;
};
static GcnOpWs gcn_op_ws(const gwn_gcn_desc* d, int nsets, bool backward) {
  GcnOpWs w;
  const i64 nw = align_up((i64)d->c_out * (1 + d->n_supports * d->order) * d->C);
  i64 o = 0;
  w.o_sup = o; o += op_support_floats(nsets, d->n_supports, d->V, d->precision);
  w.o_wlo = o; o += nw;
  w.o_wt = o; o += nw;
  w.o_wtlo = o; o += nw;
  w.o_dh = o; o += backward ? align_up((i64)d->B * d->L * d->V * d->c_out) : 0;
  w.o_part = o; o += backward ? op_part_floats(d->V) : 0;
  w.total = o;
  return w;
}

// nsets = 1: supports shared by all samples (gcn); nsets = B: one set per sample (gcn2)
static int gcn_op_fwd(const gwn_gcn_desc* d, int nsets, const float* x, const float* const* supports, const int64_t* lds_b,
                      const int64_t* lds, const float* W, const float* bias, const uint8_t* keep_mask, float* hops, float* y,
                      void* workspace, cudaStream_t st) {
  MathScope math_scope(math_of(d->precision));
  const bool tc = tc_tier(d->precision);
  GWN_CHECK_ARG(!tc || (workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0),
                "gcn_fwd: the tcgen05 tiers need a 16-byte aligned workspace of gwn_gcn_workspace_floats() floats");
  const i64 slab = (i64)d->L * d->V * d->C, PD = (i64)d->B * slab;
  const i64 P = (i64)d->B * d->L * d->V;
  const int nseg = 1 + d->n_supports * d->order;
  float* ws = reinterpret_cast<float*>(workspace);
  const GcnOpWs w = gcn_op_ws(d, nsets, false);
  if (tc) {
    OpSupports os;
    GWN_TRY(op_pack_supports(ws + w.o_sup, supports, lds_b, lds, nsets, d->n_supports, d->V, d->precision, st, &os));
    TcSupports tcs = op_tc_supports(os, nsets, d->n_supports, true, d->precision);
    GcnShape gs{d->B, d->L, d->V, d->C, d->c_out, d->n_supports, d->order};
    SupportView sv[MAXSUP];
    for (int s = 0; s < d->n_supports; ++s) sv[s] = support_fwd(supports[s], lds[s], 1);
    GWN_TRY(gcn_hops_forward(gs, x, sv, hops, st, &tcs));
    if (d->precision == GWN_PREC_FP32X3)
      GWN_LAUNCH_1D(split_lo_kernel, (i64)d->c_out * nseg * d->C, st, W, ws + w.o_wlo, (i64)d->c_out * nseg * d->C);
  } else if (nsets == 1) {
    GcnShape gs{d->B, d->L, d->V, d->C, d->c_out, d->n_supports, d->order};
    SupportView sv[MAXSUP];
    for (int s = 0; s < d->n_supports; ++s) sv[s] = support_fwd(supports[s], lds[s], 1);
    GWN_TRY(gcn_hops_forward(gs, x, sv, hops, st));
  } else {   // fp32 tier, per-sample graphs: one launch group per sample over the sample's slab
    GcnShape g1{1, d->L, d->V, d->C, d->c_out, d->n_supports, d->order};
    for (int b = 0; b < d->B; ++b) {
      SupportView sv[MAXSUP];
      for (int s = 0; s < d->n_supports; ++s) sv[s] = support_fwd(supports[s] + (i64)b * lds_b[s], lds[s], 1);
      GWN_TRY(gcn_hops_forward(g1, x + b * slab, sv, hops + b * slab, st, nullptr, PD));
    }
  }
  const float* segs[MAXSEG];
  segs[0] = x;
  for (int q = 1; q < nseg; ++q) segs[q] = hops + (i64)(q - 1) * P * d->C;
  MlpFwdArgs m;
  memset(&m, 0, sizeof(m));
  m.segs = segs; m.nseg = nseg; m.P = P; m.D = d->C; m.C_out = d->c_out; m.W = W; m.bias = bias;
  m.drop = make_dropout(d->dropout_mode, keep_mask, d->seed, d->offset, d->dropout_p);
  m.y = y;
  if (tc) {
    m.tf32_tc = 1;
    if (d->precision == GWN_PREC_FP32X3) m.W_lo = ws + w.o_wlo;
  }
  return mlp_forward(m, st);
}

static int gcn_op_bwd(const gwn_gcn_desc* d, int nsets, const float* dy, const float* x, const float* const* supports,
                      const int64_t* lds_b, const int64_t* lds, const float* W, const uint8_t* keep_mask, const float* hops,
                      float* dx, float* dW, float* dbias, float* const* dsupports, const int64_t* ldds_b, const int64_t* ldds,
                      float* scratch, cudaStream_t st) {
  MathScope math_scope(math_of(d->precision));
  const bool tc = tc_tier(d->precision);
  const i64 slab = (i64)d->L * d->V * d->C, PD = (i64)d->B * slab;
  const i64 P = (i64)d->B * d->L * d->V;
  const int nseg = 1 + d->n_supports * d->order;
  const float* segs[MAXSEG];
  segs[0] = x;
  for (int q = 1; q < nseg; ++q) segs[q] = hops + (i64)(q - 1) * P * d->C;
  // scratch: [nseg][P][C] segment gradients, then (tensor-core tiers) the operator workspace
  float* ws = scratch + (i64)nseg * P * d->C;
  const GcnOpWs w = gcn_op_ws(d, nsets, true);
  MlpBwdArgs m;
  memset(&m, 0, sizeof(m));
  m.dh = dy;
  m.drop = make_dropout(d->dropout_mode, keep_mask, d->seed, d->offset, d->dropout_p);
  m.segs = segs; m.nseg = nseg; m.P = P; m.D = d->C; m.C_out = d->c_out; m.W = W;
  m.dsegs = scratch; m.dW = dW; m.dbias = dbias;
  OpSupports os;
  TcScratch tsc{nullptr, 0, 0};
  if (tc) {
    GWN_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "gcn_bwd: scratch must be 16-byte aligned");
    GWN_TRY(op_pack_supports(ws + w.o_sup, supports, lds_b, lds, nsets, d->n_supports, d->V, d->precision, st, &os));
    const bool x3_ = d->precision == GWN_PREC_FP32X3;
    GWN_LAUNCH_1D(transpose_kernel, (i64)d->c_out * nseg * d->C, st, W, ws + w.o_wt, d->c_out, nseg * d->C,
                  x3_ ? ws + w.o_wtlo : (float*)nullptr);
    m.WT = ws + w.o_wt;
    if (x3_) m.WT_lo = ws + w.o_wtlo;
    tsc = TcScratch{ws + w.o_part, op_part_floats(d->V), x3_ ? 1 : 0};
    m.ts = tsc;
    if (m.drop.mode != GWN_DROPOUT_NONE) {   // dh = dy * keep / (1 - p) once, so that the tcgen05 kernels see a plain operand
      GWN_LAUNCH_1D(apply_dropout_kernel, P * d->c_out, st, dy, ws + w.o_dh, P * d->c_out, m.drop);
      m.dh = ws + w.o_dh;
      m.drop = make_dropout(GWN_DROPOUT_NONE, nullptr, 0, 0, 0.f);
    }
  }
  if (dW) {
    GWN_TRY(dev_memset(dW, 0, sizeof(float) * (size_t)d->c_out * nseg * d->C, st));
    GWN_TRY(dev_memset(dbias, 0, sizeof(float) * (size_t)d->c_out, st));
  }
  GWN_TRY(mlp_backward(m, st));
  if (tc || nsets == 1) {
    GcnShape gs{d->B, d->L, d->V, d->C, d->c_out, d->n_supports, d->order};
    SupportView sv[MAXSUP];
    float* dsup[MAXSUP];
    i64 ldd[MAXSUP];
    for (int s = 0; s < d->n_supports; ++s) {
      sv[s] = support_bwd(supports[s], lds[s], 1);
      dsup[s] = (dsupports && nsets == 1) ? dsupports[s] : nullptr;
      ldd[s] = ldds ? ldds[s] : d->V;
    }
    if (tc) {
      TcSupports tcs = op_tc_supports(os, nsets, d->n_supports, false, d->precision);
      GWN_TRY(gcn_hops_backward(gs, x, hops, sv, scratch, dx, nullptr, 0, dsup, ldd, st, &tcs, &tsc));
      if (nsets > 1 && dsupports) {   // per-sample support gradients: one tcgen05 reduction per (sample, support)
        GWN_CHECK_ARG(d->order <= MAXSUP, "gcn2 bwd: order too large");
        for (int s = 0; s < d->n_supports; ++s) {
          if (!dsupports[s]) continue;
          for (int b = 0; b < d->B; ++b) {
            const float* Xp[MAXSUP];
            const float* Yp[MAXSUP];
            for (int k = 1; k <= d->order; ++k) {
              Xp[k - 1] = ((k == 1) ? x : hops + (i64)(hop_index(gs, s, k - 1) - 1) * PD) + b * slab;
              Yp[k - 1] = scratch + (i64)hop_index(gs, s, k) * PD + b * slab;
            }
            GWN_TRY(support_grad_gemm(Xp, Yp, d->order, dsupports[s] + (i64)b * ldds_b[s], ldd[s], 1, d->L, d->V, d->C, st, &tsc));
          }
        }
      }
      return 0;
    }
    return gcn_hops_backward(gs, x, hops, sv, scratch, dx, nullptr, 0, dsup, ldd, st);
  }
  GcnShape g1{1, d->L, d->V, d->C, d->c_out, d->n_supports, d->order};
  for (int b = 0; b < d->B; ++b) {
    SupportView sv[MAXSUP];
    float* dsup[MAXSUP];
    i64 ldd[MAXSUP];
    for (int s = 0; s < d->n_supports; ++s) {
      sv[s] = support_bwd(supports[s] + (i64)b * lds_b[s], lds[s], 1);
      dsup[s] = (dsupports && dsupports[s]) ? dsupports[s] + (i64)b * ldds_b[s] : nullptr;
      ldd[s] = ldds ? ldds[s] : d->V;
    }
    GWN_TRY(gcn_hops_backward(g1, x + b * slab, hops + b * slab, sv, scratch + b * slab, dx + b * slab, nullptr, 0, dsup, ldd,
                              st, nullptr, nullptr, PD));
  }
  return 0;
}
}  // namespace gwn

size_t gwn_gcn_workspace_floats(const gwn_gcn_desc* d, int per_sample_supports) {
  if (!d || !tc_tier(d->precision)) return 0;
  return (size_t)gcn_op_ws(d, per_sample_supports ? d->B : 1, false).total;
}

int gwn_gcn_fwd(const gwn_gcn_desc* d, const float* x, const float* const* supports, const int64_t* lds, const float* W,
                const float* bias, const uint8_t* keep_mask, float* hops, float* y, void* workspace, void* stream) {
  GWN_TRY(require_device());
  GWN_TRY(gcn_check(d));
  GWN_CHECK_ARG(x && supports && lds && W && bias && hops && y, "gcn_fwd: null pointer");
  GWN_CHECK_ARG(d->dropout_mode != GWN_DROPOUT_MASK || keep_mask, "gcn_fwd: GWN_DROPOUT_MASK without keep_mask");
  return gcn_op_fwd(d, 1, x, supports, lds, lds, W, bias, keep_mask, hops, y, workspace, (cudaStream_t)stream);
}

size_t gwn_gcn_bwd_scratch_floats(const gwn_gcn_desc* d) {
  if (!d) return 0;
  const size_t segs = (size_t)(1 + d->n_supports * d->order) * d->B * d->L * d->V * d->C;
  if (!tc_tier(d->precision)) return segs;
  return segs + (size_t)gcn_op_ws(d, d->B, true).total;   // sized for per-sample supports too (gwn_gcn2_bwd)
}

int gwn_gcn_bwd(const gwn_gcn_desc* d, const float* dy, const float* x, const float* const* supports, const int64_t* lds,
                const float* W, const uint8_t* keep_mask, const float* hops, float* dx, float* dW, float* dbias,
                float* const* dsupports, const int64_t* ldds, float* scratch, void* stream) {
  GWN_TRY(require_device());
  GWN_TRY(gcn_check(d));
  GWN_CHECK_ARG(dy && x && supports && lds && W && hops && dx && scratch, "gcn_bwd: null pointer");
  GWN_CHECK_ARG((dW == nullptr) == (dbias == nullptr), "gcn_bwd: dW and dbias must be given together");
  return gcn_op_bwd(d, 1, dy, x, supports, lds, lds, W, keep_mask, hops, dx, dW, dbias, dsupports, ldds, ldds, scratch,
                    (cudaStream_t)stream);
}

// ---- per-sample-graph operators (model.py:16-22, 57-80).  fp32 tier: one launch group per sample over the sample's
// slab; tensor-core tiers: all samples' graphs in one launch (batched tensor maps of nconv_tc_kernel)
int gwn_nconv2_fwd(const float* x, const float* A, int64_t lda_b, int64_t lda, float* y, int B, int L, int V, int C,
                   int precision, void* workspace, void* stream) {
  GWN_TRY(require_device());
  GWN_CHECK_ARG(x && A && y, "nconv2_fwd: null pointer");
  GWN_TRY(nconv_tc_args_ok("nconv2_fwd", C, precision, workspace));
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_tier(precision)) {
    MathScope math_scope(math_of(precision));
    OpSupports os;
    GWN_TRY(op_pack_supports(reinterpret_cast<float*>(workspace), &A, &lda_b, &lda, B, 1, V, precision, st, &os));
    TcSupports tcs = op_tc_supports(os, B, 1, true, precision);
    SupportView sv = support_fwd(A, lda, 1);
    const float* X[1] = {x};
    float* Y[1] = {y};
    return node_gemm(&sv, 1, false, X, Y, nullptr, nullptr, B, L, 0, V, C, st, &tcs);
  }
  const i64 slab = (i64)L * V * C;
  for (int b = 0; b < B; ++b) {
    SupportView sv = support_fwd(A + (i64)b * lda_b, lda, 1);
    const float* X[1] = {x + b * slab};
    float* Y[1] = {y + b * slab};
    GWN_TRY(node_gemm(&sv, 1, false, X, Y, nullptr, nullptr, 1, L, 0, V, C, st));
  }
  return 0;
}

int gwn_nconv2_bwd(const float* dy, const float* x, const float* A, int64_t lda_b, int64_t lda, float* dx, float* dA,
                   int64_t ldda_b, int64_t ldda, int B, int L, int V, int C, int precision, void* workspace, void* stream) {
  GWN_TRY(require_device());
  GWN_CHECK_ARG(dy && A, "nconv2_bwd: null pointer");
  GWN_TRY(nconv_tc_args_ok("nconv2_bwd", C, precision, workspace));
  GWN_CHECK_ARG(!dA || x, "nconv2_bwd: x needed for dA");
  cudaStream_t st = (cudaStream_t)stream;
  const i64 slab = (i64)L * V * C;
  if (tc_tier(precision)) {
    MathScope math_scope(math_of(precision));
    float* ws = reinterpret_cast<float*>(workspace);
    OpSupports os;
    GWN_TRY(op_pack_supports(ws, &A, &lda_b, &lda, B, 1, V, precision, st, &os));
    if (dx) {
      TcSupports tcs = op_tc_supports(os, B, 1, false, precision);
      SupportView sv = support_bwd(A, lda, 1);
      const float* X[1] = {dy};
      float* Y[1] = {dx};
      GWN_TRY(node_gemm(&sv, 1, false, X, Y, nullptr, nullptr, B, L, 0, V, C, st, &tcs));
    }
    if (dA) {
      TcScratch ts{ws + op_support_floats(B, 1, V, precision), op_part_floats(V), precision == GWN_PREC_FP32X3 ? 1 : 0};
      for (int b = 0; b < B; ++b) {
        const float* Xp[1] = {x + b * slab};
        const float* Yp[1] = {dy + b * slab};
        GWN_TRY(support_grad_gemm(Xp, Yp, 1, dA + (i64)b * ldda_b, ldda, 1, L, V, C, st, &ts));
      }
    }
    return 0;
  }
  for (int b = 0; b < B; ++b) {
    if (dx) {
      SupportView sv = support_bwd(A + (i64)b * lda_b, lda, 1);
      const float* X[1] = {dy + b * slab};
      float* Y[1] = {dx + b * slab};
      GWN_TRY(node_gemm(&sv, 1, false, X, Y, nullptr, nullptr, 1, L, 0, V, C, st));
    }
    if (dA) {
      const float* Xp[1] = {x + b * slab};
      const float* Yp[1] = {dy + b * slab};
      GWN_TRY(support_grad_gemm(Xp, Yp, 1, dA + (i64)b * ldda_b, ldda, 1, L, V, C, st));
    }
  }
  return 0;
}

int gwn_gcn2_fwd(const gwn_gcn_desc* d, const float* x, const float* const* supports, const int64_t* lds_b, const int64_t* lds,
                 const float* W, const float* bias, const uint8_t* keep_mask, float* hops, float* y, void* workspace,
                 void* stream) {
  GWN_TRY(require_device());
  GWN_TRY(gcn_check(d));
  GWN_CHECK_ARG(x && supports && lds && lds_b && W && bias && hops && y, "gcn2_fwd: null pointer");
  GWN_CHECK_ARG(d->dropout_mode != GWN_DROPOUT_MASK || keep_mask, "gcn2_fwd: GWN_DROPOUT_MASK without keep_mask");
  return gcn_op_fwd(d, d->B, x, supports, lds_b, lds, W, bias, keep_mask, hops, y, workspace, (cudaStream_t)stream);
}

int gwn_gcn2_bwd(const gwn_gcn_desc* d, const float* dy, const float* x, const float* const* supports, const int64_t* lds_b,
                 const int64_t* lds, const float* W, const uint8_t* keep_mask, const float* hops, float* dx, float* dW,
                 float* dbias, float* const* dsupports, const int64_t* ldds_b, const int64_t* ldds, float* scratch,
                 void* stream) {
  GWN_TRY(require_device());
  GWN_TRY(gcn_check(d));
  GWN_CHECK_ARG(dy && x && supports && lds && lds_b && W && hops && dx && scratch, "gcn2_bwd: null pointer");
  GWN_CHECK_ARG((dW == nullptr) == (dbias == nullptr), "gcn2_bwd: dW and dbias must be given together");
  GWN_CHECK_ARG(!dsupports || (ldds_b && ldds), "gcn2_bwd: support-gradient strides missing");
  return gcn_op_bwd(d, d->B, dy, x, supports, lds_b, lds, W, keep_mask, hops, dx, dW, dbias, dsupports, ldds_b, ldds, scratch,
                    (cudaStream_t)stream);
}

int gwn_plan_create(const gwn_config* cfg, gwn_plan** out) {
  GWN_CHECK_ARG(cfg && out, "plan_create: null argument");
  gwn_plan* p = new gwn_plan();
  p->c = *cfg;
  int st = build_plan(p);
  if (st != 0) {
    delete p;
    return st;
  }
  *out = p;
  return 0;
}

void gwn_plan_destroy(gwn_plan* p) {
  if (!p) return;
#if !GWN_EMU
  for (int k = 0; k < 4; ++k)
    if (p->side_ev[k]) cudaEventDestroy(p->side_ev[k]);
  if (p->side) cudaStreamDestroy(p->side);
#endif
  delete p;
}

int gwn_plan_workspace_bytes(const gwn_plan* p, size_t* fwd, size_t* bwd) {
  GWN_CHECK_ARG(p, "null plan");
  if (fwd) *fwd = (size_t)p->fwd_floats * sizeof(float);
  if (bwd) *bwd = (size_t)p->bwd_floats * sizeof(float);
  return 0;
}

int gwn_plan_param_count(const gwn_plan* p, int* n_entries, int64_t* grad_floats) {
  GWN_CHECK_ARG(p, "null plan");
  if (n_entries) *n_entries = (int)p->entries.size();
  if (grad_floats) *grad_floats = p->grad_floats;
  return 0;
}

int gwn_plan_param_info(const gwn_plan* p, int i, char* name, int name_len, int64_t* grad_offset, int64_t* numel) {
  GWN_CHECK_ARG(p && i >= 0 && i < (int)p->entries.size(), "param_info: bad index");
  if (name && name_len > 0) snprintf(name, name_len, "%s", p->entries[i].name.c_str());
  if (grad_offset) *grad_offset = p->entries[i].grad_off;
  if (numel) *numel = p->entries[i].numel;
  return 0;
}

int gwn_plan_out_len(const gwn_plan* p, int* t_out, int* receptive_field) {
  GWN_CHECK_ARG(p, "null plan");
  if (t_out) *t_out = p->T_out;
  if (receptive_field) *receptive_field = p->RF;
  return 0;
}

int gwn_plan_debug_layout(const gwn_plan* p, char* buf, int len) {
  GWN_CHECK_ARG(p && buf && len > 0, "debug_layout: bad argument");
  std::string s;
  char t[160];
  auto put = [&](const char* space, const char* name, int idx, i64 off, i64 n) {
    snprintf(t, sizeof(t), "%s %s%d %lld %lld\n", space, name, idx, (long long)off, (long long)n);
    s += t;
  };
  const gwn_config& c = p->c;
  const i64 N = c.num_nodes;
  const int C = c.residual_channels, D = c.dilation_channels;
  put("fwd", "sup", 0, p->o_sup, (i64)std::max(p->S, 1) * N * p->ld);
  put("fwd", "supT", 0, p->o_supT, (i64)std::max(p->S, 1) * N * p->ld);
  put("fwd", "x0_", 0, p->o_x0, p->P0() * C);
  for (int i = 0; i < p->nL; ++i) {
    for (int q = 0; q < p->nseg; ++q) put("fwd", q == 0 ? "g" : (q == 1 ? "hopA_" : (q == 2 ? "hopB_" : "hopN_")), i, p->o_g[i] + (i64)q * p->P(i) * D, p->P(i) * D);
    put("fwd", "u", i, p->o_u[i], p->P(i) * C);
    put("fwd", "ac", i, p->o_ac[i], 2 * C);
    put("fwd", "mr", i, p->o_mr[i], 2 * C);
  }
  put("fwd", "skip", 0, p->o_skip, p->PT() * c.skip_channels);
  put("fwd", "e1_", 0, p->o_e1, p->PT() * c.end_channels);
  put("bwd", "dgh", 0, p->o_dgh, (i64)p->nL * p->PT() * D);
  put("bwd", "dout", 0, p->o_dout, p->PT() * p->ldo);
  put("bwd", "de1_", 0, p->o_de1, p->PT() * c.end_channels);
  put("bwd", "dskip", 0, p->o_dskip, p->PT() * c.skip_channels);
  put("bwd", "dA", 0, p->o_dA, N * p->ld);
  put("bwd", "dR", 0, p->o_dR, N * p->ld);
  put("bwd", "buf", 0, p->o_buf0, p->P0() * C);
  put("bwd", "buf", 1, p->o_buf1, p->P0() * C);
  put("bwd", "dsegs", 0, p->o_dsegs[0], p->P(0) * D * p->nseg);
  put("bwd", "dg", 0, p->o_dg, p->P(0) * D);
  put("bwd", "dpre", 0, p->o_dpre, p->P(0) * 2 * D);
  snprintf(buf, len, "%s", s.c_str());
  return 0;
}

int gwn_plan_forward(gwn_plan* p, const gwn_forward_args* a) {
  GWN_CHECK_ARG(p && a, "plan_forward: null argument");
  return plan_forward(p, a);
}

int gwn_plan_backward(gwn_plan* p, const gwn_backward_args* a) {
  GWN_CHECK_ARG(p && a, "plan_backward: null argument");
  return plan_backward(p, a);
}

size_t gwn_train_ctrl_bytes(void) { return sizeof(TrainCtrl); }

int gwn_train_ctrl_init(void* ctrl, uint64_t seed, int64_t step) {
  GWN_CHECK_ARG(ctrl, "train_ctrl_init: null control block");
  TrainCtrl c;
  memset(&c, 0, sizeof(c));
  c.seed = seed;
  c.step = step;
#if GWN_EMU
  memcpy(ctrl, &c, sizeof(c));
#else
  GWN_TRY(require_device());
  GWN_CUDA(cudaMemcpy(ctrl, &c, sizeof(c), cudaMemcpyHostToDevice));
#endif
  return 0;
}

int gwn_train_ctrl_read(const void* ctrl, uint64_t* seed, int64_t* step) {
  GWN_CHECK_ARG(ctrl, "train_ctrl_read: null control block");
  TrainCtrl c;
#if GWN_EMU
  memcpy(&c, ctrl, sizeof(c));
#else
  GWN_TRY(require_device());
  GWN_CUDA(cudaMemcpy(&c, ctrl, sizeof(c), cudaMemcpyDeviceToHost));
#endif
  if (seed) *seed = c.seed;
  if (step) *step = c.step;
  return 0;
}

int gwn_plan_train_fwd_bwd(gwn_plan* p, const gwn_train_args* a) {
  GWN_CHECK_ARG(p && a, "train_fwd_bwd: null argument");
  GWN_CHECK_ARG(a->ctrl && a->metrics && a->target && a->scratch && a->grad_flat && a->fwd.output, "train_fwd_bwd: null buffer");
  GWN_TRY(require_device());
  const gwn_config& c = p->c;
  cudaStream_t st = (cudaStream_t)a->fwd.stream;
  TrainCtrl* ctrl = reinterpret_cast<TrainCtrl*>(a->ctrl);
  GWN_LAUNCH_1D(train_begin_kernel, 1, st, ctrl);
  gwn_forward_args fa = a->fwd;
  fa.training = 1;
  fa.seed = 0;
  fa.seed_device = reinterpret_cast<const uint64_t*>(&ctrl->seed);
  GWN_TRY(plan_forward(p, &fa));
  float* sc = reinterpret_cast<float*>(a->scratch);
  {
    const i64 n = (i64)c.batch * c.out_dim * c.num_nodes * p->T_out;
    ProfScope prof("loss_metrics", st, 4.0 * n * 3.0 + 4.0 * p->PT() * p->ldo, 0.0);
    GWN_LAUNCH_1D(loss_reduce_kernel, n, st, (const float*)a->fwd.output, a->target, (i64)a->target_strides[0],
                  (i64)a->target_strides[1], (i64)a->target_strides[2], a->scaler_mean, a->scaler_std, c.batch, c.out_dim,
                  c.num_nodes, p->T_out, ctrl);
    GWN_LAUNCH_1D(loss_grad_kernel, p->PT() * p->ldo, st, (const float*)a->fwd.output, a->target, (i64)a->target_strides[0],
                  (i64)a->target_strides[1], (i64)a->target_strides[2], a->scaler_mean, a->scaler_std, c.batch, c.out_dim,
                  c.num_nodes, p->T_out, p->ldo, (const TrainCtrl*)ctrl, sc + p->o_dout, a->metrics);
  }
  gwn_backward_args ba;
  memset(&ba, 0, sizeof(ba));
  ba.params = fa.params; ba.supports = fa.supports; ba.support_strides = fa.support_strides; ba.input = fa.input;
  for (int k = 0; k < 4; ++k) ba.input_strides[k] = fa.input_strides[k];
  ba.grad_output = nullptr; ba.workspace = fa.workspace; ba.scratch = a->scratch; ba.grad_flat = a->grad_flat;
  ba.grad_input = nullptr; ba.training = 1; ba.dropout_mode = fa.dropout_mode; ba.keep_masks = fa.keep_masks;
  ba.seed = 0; ba.stream = fa.stream; ba.seed_device = fa.seed_device;
  return plan_backward(p, &ba, true);
}

int gwn_plan_eval_metrics(gwn_plan* p, const gwn_train_args* a) {
  GWN_CHECK_ARG(p && a, "eval_metrics: null argument");
  GWN_CHECK_ARG(a->ctrl && a->metrics && a->target && a->fwd.output, "eval_metrics: null buffer");
  GWN_TRY(require_device());
  const gwn_config& c = p->c;
  cudaStream_t st = (cudaStream_t)a->fwd.stream;
  TrainCtrl* ctrl = reinterpret_cast<TrainCtrl*>(a->ctrl);
  GWN_TRY(dev_memset(&ctrl->acc[0], 0, sizeof(double) * 4, st));
  gwn_forward_args fa = a->fwd;
  fa.training = 0;
  fa.dropout_mode = GWN_DROPOUT_NONE;
  GWN_TRY(plan_forward(p, &fa));
  const i64 n = (i64)c.batch * c.out_dim * c.num_nodes * p->T_out;
  ProfScope prof("loss_metrics", st, 4.0 * n * 2.0, 0.0);
  GWN_LAUNCH_1D(loss_reduce_kernel, n, st, (const float*)a->fwd.output, a->target, (i64)a->target_strides[0],
                (i64)a->target_strides[1], (i64)a->target_strides[2], a->scaler_mean, a->scaler_std, c.batch, c.out_dim,
                c.num_nodes, p->T_out, ctrl);
  GWN_LAUNCH_1D(metrics_finalize_kernel, 1, st, (const TrainCtrl*)ctrl, p->T_out, a->metrics);
  return 0;
}

int gwn_adam_step(const gwn_adam_args* a) {
  GWN_CHECK_ARG(a && a->param_flat && a->grad_flat && a->exp_avg && a->exp_avg_sq && a->live4 && a->hyper && a->ctrl,
                "adam_step: null argument");
  GWN_CHECK_ARG(a->n > 0 && a->n % 4 == 0, "adam_step: n must be a positive multiple of 4");
  GWN_TRY(require_device());
  cudaStream_t st = (cudaStream_t)a->stream;
  TrainCtrl* ctrl = reinterpret_cast<TrainCtrl*>(a->ctrl);
  ProfScope prof("clip_adam", st, 4.0 * a->n * 8.0, 0.0);
  GWN_TRY(dev_memset(&ctrl->acc[4], 0, sizeof(double), st));
  GWN_LAUNCH_1D(gradnorm_kernel, a->n / 4, st, (const float*)a->grad_flat, a->live4, a->n / 4, ctrl);
  GWN_LAUNCH_1D(adam_kernel, a->n / 4, st, a->param_flat, a->grad_flat, a->exp_avg, a->exp_avg_sq, a->live4, a->n / 4,
                reinterpret_cast<const AdamHyper*>(a->hyper), (const TrainCtrl*)ctrl, a->metrics, (const float*)nullptr);
  return 0;
}

// ---- data-parallel step tail over NVLink peer memory (p2p_allreduce.cuh)
size_t gwn_p2p_header_bytes(void) {
#if GWN_EMU
  return 4096;
#else
  return P2P_FLAG_BYTES;
#endif
}

int gwn_p2p_alloc(size_t bytes, void** base, unsigned char* handle64) {
  GWN_CHECK_ARG(base && handle64 && bytes > 0, "p2p_alloc: bad argument");
#if GWN_EMU
  set_error("p2p: not part of the host emulation");
  return GWN_ERR_UNSUPPORTED;
#else
  GWN_TRY(require_device());
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  GWN_CUDA(cudaMalloc(&p, bytes));
  GWN_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return GWN_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  *base = p;
  return 0;
#endif
}

int gwn_p2p_open(const unsigned char* handle64, void** base) {
  GWN_CHECK_ARG(base && handle64, "p2p_open: bad argument");
#if GWN_EMU
  set_error("p2p: not part of the host emulation");
  return GWN_ERR_UNSUPPORTED;
#else
  GWN_TRY(require_device());
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return GWN_ERR_CUDA;
  }
  *base = p;
  return 0;
#endif
}

int gwn_p2p_close(void* base) {
#if !GWN_EMU
  if (base) cudaIpcCloseMemHandle(base);
#else
  (void)base;
#endif
  return 0;
}

int gwn_p2p_free(void* base) {
#if !GWN_EMU
  if (base) cudaFree(base);
#else
  (void)base;
#endif
  return 0;
}

int gwn_allreduce_adam_step(const gwn_adam_args* a, const gwn_p2p_args* p) {
  GWN_CHECK_ARG(a && p && a->param_flat && a->grad_flat && a->exp_avg && a->exp_avg_sq && a->live4 && a->hyper && a->ctrl,
                "allreduce_adam_step: null argument");
  GWN_CHECK_ARG(a->n > 0 && a->n % 4 == 0, "allreduce_adam_step: n must be a positive multiple of 4");
#if GWN_EMU
  set_error("p2p: not part of the host emulation");
  return GWN_ERR_UNSUPPORTED;
#else
  GWN_CHECK_ARG(p->world >= 2 && p->world <= P2P_MAXRANKS && p->rank >= 0 && p->rank < p->world && p->sum_out,
                "allreduce_adam_step: world must be in [2,%d]", P2P_MAXRANKS);
  GWN_TRY(require_device());
  cudaStream_t st = (cudaStream_t)a->stream;
  TrainCtrl* ctrl = reinterpret_cast<TrainCtrl*>(a->ctrl);
  P2PArgs k;
  memset(&k, 0, sizeof(k));
  for (int q = 0; q < p->world; ++q) {
    GWN_CHECK_ARG(p->base[q] != nullptr, "allreduce_adam_step: rank %d's buffer is not mapped", q);
    k.flags[q] = reinterpret_cast<unsigned*>(p->base[q]);
    k.grad[q] = reinterpret_cast<const float*>(reinterpret_cast<const char*>(p->base[q]) + P2P_FLAG_BYTES);
  }
  GWN_CHECK_ARG(k.grad[p->rank] == a->grad_flat, "allreduce_adam_step: grad_flat must be this rank's p2p gradient buffer");
  k.out = p->sum_out; k.live4 = a->live4; k.n4 = a->n / 4; k.c = ctrl; k.rank = p->rank; k.world = p->world;
  ProfScope prof("p2p_allreduce_clip_adam", st, 4.0 * a->n * (8.0 + p->world), 0.0);
  GWN_TRY(dev_memset(&ctrl->acc[4], 0, sizeof(double), st));
  const i64 want = (k.n4 + P2P_THREADS - 1) / P2P_THREADS;
  const int grid = (int)std::min<i64>(want, 148);      // all blocks co-resident: they wait for one another's peers
  GWN_CUDA(launch_kernel(p2p_allreduce_gradnorm_kernel, dim3((unsigned)grid), dim3(P2P_THREADS), 0, st, k));
  count_launch();
  GWN_LAUNCH_1D(adam_kernel, a->n / 4, st, a->param_flat, a->grad_flat, a->exp_avg, a->exp_avg_sq, a->live4, a->n / 4,
                reinterpret_cast<const AdamHyper*>(a->hyper), (const TrainCtrl*)ctrl, a->metrics, (const float*)p->sum_out);
  return 0;
#endif
}

}  // extern "C"
