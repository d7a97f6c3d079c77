// Host-side launchers of the fused operators (fp32 tier) shared by the op-level C ABI
// (gwn_nconv_*, gwn_linear_*, gwn_gcn_*) and the whole-network plan (gwn_plan_*).
#pragma once
#include "functors.cuh"
#include "elementwise.cuh"
#include "nconv_tc.cuh"
#include "posgemm.cuh"
#include "tcpos.cuh"
#include "tcred.cuh"

namespace gwn {

typedef Tile<128, 128, 8, 8> TBig;    // node contraction, wide position GEMMs (256 threads)
typedef Tile<128, 64, 8, 4> TPos64;   // gated conv: N = 2*D = 64 (256 threads)
typedef Tile<128, 32, 8, 4> TPos32;   // N <= 32 outputs per position (128 threads)
typedef Tile<32, 128, 4, 4, 32> TW32; // weight gradients with <= 32 output rows (256 threads), BK = 32: more loads in flight
typedef Tile<64, 64, 4, 4, 32> TW64;  // small square reductions: gated-conv wgrad, dA (256 threads), BK = 32

typedef Tile<256, 32, 8, 4> TPG;      // epilogue mapping of posgemm.cuh: 8 warps x 32 rows, 32 columns per pass

constexpr int kTargetBlocks = 148 * 2;   // split-K reductions: one resident wave (2 blocks of 256 threads per SM)

// One support as the node-contraction A operand: op(k, m) = p[k*ks + m*ms].
//   forward  y[w] = sum_v A[v,w] x[v]:  k = v, m = w  -> (ks, ms) = (row stride, col stride) of A
//   backward dx[v] = sum_w A[v,w] dy[w]: k = w, m = v -> (ks, ms) = (col stride, row stride) of A, or the
//   row-major strides of a materialised A^T.
struct SupportView {
  const float* p;
  i64 ks, ms;
  int mlim;  // readable extent along m when ms == 1 (zero-padded ld), or 0 if only [0, M) is readable
};
inline SupportView support_fwd(const float* A, i64 rs, i64 cs) { return SupportView{A, rs, cs, 0}; }
inline SupportView support_bwd(const float* A, i64 rs, i64 cs) { return SupportView{A, cs, rs, 0}; }
inline SupportView support_padded(const float* Ap, int ld) { return SupportView{Ap, ld, 1, ld}; }

inline void fill_support(LdSupport& l, int idx, const SupportView& s, int M) {
  l.p[idx] = s.p;
  l.rs[idx] = s.ks;
  l.cs[idx] = s.ms;
  bool aligned = ((reinterpret_cast<uintptr_t>(s.p) & 15) == 0) && (s.ks % 4 == 0) && (s.ms == 1);
  l.vec[idx] = aligned ? 1 : 0;
  l.xlim[idx] = (s.mlim > 0) ? s.mlim : M;
}

// Y_s[m] = sum_k op_s(k,m) X_s[k] (+ add_s) for s < nsup (batched), or, when kcat,
// Y_0[m] = sum_s sum_k op_s(k,m) X_s[k] (+ add_0 + window(add2)).
// `tcs` (nullable): the same supports as K-contiguous padded buffers S[m][k] (ld = tc_ld) for the tcgen05 tier.
struct TcSupports {
  const float* S[MAXSUP];
  const float* Slo[MAXSUP];   // fp32x3 tier: remainders S - tf32_trunc(S)
  int ld;
  int precision;   // gwn_precision
  int per_sample;  // 1: S / Slo point at sample 0 of a per-sample support set, samples batch_stride floats apart
  i64 batch_stride;
};
inline int node_gemm(const SupportView* sup, int nsup, bool kcat, const float* const* X,
                     float* const* Y, const float* const* add, const float* add2, int B, int L, int T_out, int V, int C,
                     cudaStream_t stream, const TcSupports* tcs = nullptr) {
  GWN_CHECK_ARG(nsup >= 1 && nsup <= MAXSUP, "node_gemm: %d supports (max %d)", nsup, MAXSUP);
  GWN_CHECK_ARG(C % 4 == 0, "node_gemm: channels (%d) must be a multiple of 4", C);
  const bool x3 = tcs && tcs->precision == GWN_PREC_FP32X3 && tcs->Slo[0] && C == 32 && nsup <= TC_MAXSUP;
  if (tcs && (tcs->precision == GWN_PREC_TF32 || x3)) {
    GWN_CHECK_ARG(C == 32, "node_gemm: the tcgen05 tier needs 32 channels per slab row (got %d)", C);
    GWN_CHECK_ARG(nsup <= TC_MAXSUP, "node_gemm: the tcgen05 tier takes at most %d supports", TC_MAXSUP);
    NodeTcArgs t;
    memset(&t, 0, sizeof(t));
    for (int s = 0; s < nsup; ++s) { t.X[s] = X[s]; t.S[s] = tcs->S[s]; t.Slo[s] = x3 ? tcs->Slo[s] : nullptr; }
    const int nout = kcat ? 1 : nsup;
    for (int s = 0; s < nout; ++s) { t.Y[s] = Y[s]; t.add[s] = add ? add[s] : nullptr; }
    t.ld = tcs->ld; t.nsup = nsup; t.kcat = kcat ? 1 : 0; t.add2 = add2; t.B = B; t.L = L; t.T_out = T_out; t.V = V;
    t.per_sample = tcs->per_sample; t.s_batch_stride = tcs->batch_stride;
    return node_gemm_tc(t, stream);
  }
  GWN_CHECK_ARG(!(tcs && tcs->per_sample), "node_gemm: per-sample supports need the tcgen05 tiers (the caller loops over samples otherwise)");
  LdSupport a;
  LdSlab b;
  EpSlab e;
  memset(&a, 0, sizeof(a));
  memset(&b, 0, sizeof(b));
  memset(&e, 0, sizeof(e));
  for (int s = 0; s < nsup; ++s) {
    fill_support(a, s, sup[s], V);
    b.p[s] = X[s];
  }
  a.kper = V; a.kcat = kcat ? 1 : 0;
  b.V = V; b.C = C; b.kper = V; b.kcat = kcat ? 1 : 0;
  int nout = kcat ? 1 : nsup;
  for (int s = 0; s < nout; ++s) {
    e.y[s] = Y[s];
    e.add[s] = add ? add[s] : nullptr;
  }
  e.add2 = add2;
  e.V = V; e.C = C; e.L = L; e.T_out = T_out;
  GemmShape sh{(i64)V, (int)((i64)B * L * C), kcat ? nsup * V : V, 1, nout};
  GWN_CHECK_ARG((i64)B * L * C < 2147483647LL, "node_gemm: B*L*C too large");
  return launch_gemm<TBig>(a, b, e, sh, stream);
}

// Partial-result scratch of the tcgen05 reductions (tcred.cuh); null = not available (op-level calls).
struct TcScratch {
  float* partial;
  i64 floats;
  int x3;   // 1: fp32x3 tier (3xTF32 split inside the reduction)
};

// tcgen05 support gradient over any number of (X, T) pairs with their own slab counts (all layers of a backward pass
// in ONE launch): dA[v,w] += sum over pairs, slabs, c of X[(slab,v),c] * T[(slab,w),c].  -1 = not eligible.
inline int support_grad_tc(const float* const* Xp, const float* const* Yp, const int* slabs, int npairs, float* dA, i64 ldda,
                           int V, int C, const TcScratch& ts, cudaStream_t stream) {
#if GWN_EMU
  (void)Xp; (void)Yp; (void)slabs; (void)npairs; (void)dA; (void)ldda; (void)V; (void)C; (void)ts; (void)stream;
  return -1;
#else
  if (C != 32 || npairs < 1 || npairs > TR_MAXSRC || !ts.partial || ldda > 2147483647LL) return -1;
  TcRedArgs t;
  memset(&t, 0, sizeof(t));
  t.mode = 1; t.na = npairs; t.x3 = ts.x3;
  for (int i = 0; i < npairs; ++i) {
    t.a[i] = TcRedSrc{Xp[i], V, 32, 0, 0, slabs[i]};
    t.b[i] = TcRedSrc{Yp[i], V, 32, 0, 0, slabs[i]};
  }
  t.rows = V; t.partial = ts.partial; t.partial_floats = ts.floats;
  TcRedResult r;
  int st = launch_tcred(t, stream, &r);
  if (st != 0) return st;
  tc::SlotSupOut f{dA, (int)ldda, V, r.N, r.n_nt, r.mtiles * 128, r.n_mg};
  return launch_slot_reduce(ts.partial, r, (i64)V * V, f, stream);
#endif
}

// dA[v,w] += sum over pairs, slabs, c of Xp[(slab,v),c] * Yp[(slab,w),c]
inline int support_grad_gemm(const float* const* Xp, const float* const* Yp, int npairs, float* dA, i64 ldda, int B, int L,
                             int V, int C, cudaStream_t stream, const TcScratch* ts = nullptr) {
  GWN_CHECK_ARG(npairs >= 1 && npairs <= MAXSUP, "support_grad: bad pair count %d", npairs);
  if (current_math() != 0 && ts) {   // tensor-core tiers: tcgen05 + TMA
    int slabs[MAXSUP];
    for (int i = 0; i < npairs; ++i) slabs[i] = B * L;
    int st = support_grad_tc(Xp, Yp, slabs, npairs, dA, ldda, V, C, *ts, stream);
    if (st >= 0) return st;
  }
  LdSlabK a, b;
  memset(&a, 0, sizeof(a));
  memset(&b, 0, sizeof(b));
  for (int i = 0; i < npairs; ++i) { a.p[i] = Xp[i]; b.p[i] = Yp[i]; }
  a.V = b.V = V; a.C = b.C = C;
  a.cdiv = b.cdiv = make_divw(C);
  i64 K = (i64)B * L * C * npairs;
  GWN_CHECK_ARG(K < 2147483647LL, "support_grad: K too large");
  a.kper = b.kper = (int)((i64)B * L * C);
  EpAtomicMat e{dA, ldda};
  GemmShape sh{(i64)V, V, (int)K, pick_ksplit(V, V, K, TW64::BM, TW64::BN, kTargetBlocks), 1};
  return launch_gemm<TW64>(a, b, e, sh, stream);
}

struct MlpFwdArgs {
  const float* const* segs;  // nseg tensors [P, D]
  int nseg;
  i64 P;
  int D, C_out;
  const float* W;            // [C_out, nseg*D]
  const float* W_lo;         // nullable: 3xTF32 remainders of W (fp32x3 tier on the tcgen05 path)
  const float* bias;
  DropoutSrc drop;
  const float* res;          // nullable residual source
  Remap rrm;
  const float* rac;
  double* stats;             // nullable
  float* y;
  int tf32_tc;               // 1: tf32 tier -> tcgen05/TMA kernel when the shape allows
  // per-sample geometry of the residual (tcgen05 path): P = nb * rows_per_sample, residual row of output row r of a
  // sample = row r + res_rshift of a source with res_rows_src rows per sample.  nb == 0: flat positions, no residual.
  int nb, rows_per_sample, res_rows_src, res_rshift;
};
inline int mlp_forward(const MlpFwdArgs& m, cudaStream_t stream) {
  GWN_CHECK_ARG(m.nseg >= 1 && m.nseg <= MAXSEG, "mlp: %d segments (max %d)", m.nseg, MAXSEG);
  GWN_CHECK_ARG(m.D % 4 == 0, "mlp: c_in per segment (%d) must be a multiple of 4", m.D);
  GWN_CHECK_ARG((reinterpret_cast<uintptr_t>(m.W) & 15) == 0, "mlp: weight pointer must be 16-byte aligned");
  ProfScope prof("gcn_mlp_fwd", stream, 4.0 * m.P * ((double)m.nseg * m.D + (m.res ? 2.0 : 1.0) * m.C_out),
                 2.0 * m.P * m.nseg * m.D * m.C_out);
  LdWK b;
  memset(&b, 0, sizeof(b));
  b.p[0] = m.W; b.set_wd(m.nseg * m.D); b.ldw = m.nseg * m.D;
  if (m.tf32_tc && m.D == 32 && m.C_out == 32 && m.nseg <= TP_MAXSEG && m.P < 2147483647LL &&
      (!m.res || (m.nb > 0 && (i64)m.nb * m.rows_per_sample == m.P))) {   // tcgen05 + TMA path
    TcPosArgs t;
    memset(&t, 0, sizeof(t));
    const int nb = m.nb > 0 ? m.nb : 1, rows = m.nb > 0 ? m.rows_per_sample : (int)m.P;
    for (int q = 0; q < m.nseg; ++q) t.seg[q] = TcPosSeg{m.segs[q], rows, 32, 0, 0};
    t.nseg = m.nseg; t.nb = nb; t.rows_out = rows; t.Wp = m.W; t.N = 32;
    t.out = m.y; t.out_width = 32; t.out_nblk = 1; t.Wp_lo = m.W_lo;
    if (m.res) t.addend[0] = TcPosSeg{m.res, m.res_rows_src, 32, 0, m.res_rshift};
    RowMlp eg;
    memset(&eg, 0, sizeof(eg));
    eg.y = m.y; eg.bias = m.bias; eg.drop = m.drop; eg.res = m.res; eg.rrm = m.rrm; eg.rac = m.rac; eg.stats = m.stats;
    {
      static const bool nostats = getenv("GWNET_B200_DIAG_NOSTATS") != nullptr;   // timing diagnostics only (wrong results)
      if (nostats) eg.stats = nullptr;
    }
    int st = launch_tcpos<32>(t, eg, stream);
    if (st >= 0) return st;
  }
  if (current_math() != 0 && m.D == PG_WD && m.nseg <= PG_MAXSEG && m.C_out == 32) {   // direct-fragment tensor-core path
    ARows ar;
    memset(&ar, 0, sizeof(ar));
    for (int q = 0; q < m.nseg; ++q) ar.P[q] = m.segs[q];
    ar.rm[0] = ar.rm[1] = rowmap_identity();
    ar.nseg = m.nseg; ar.rs = m.D;
    EpMlp<TPG> eg;
    memset(&eg, 0, sizeof(eg));
    eg.y = m.y; eg.bias = m.bias; eg.C = m.C_out; eg.drop = m.drop; eg.res = m.res; eg.rrm = m.rrm; eg.rac = m.rac; eg.stats = m.stats;
    int st = launch_posgemm<TPG, 32>(ar, b, eg, m.P, m.C_out, stream);
    if (st >= 0) return st;
  }
  LdRows a;
  memset(&a, 0, sizeof(a));
  for (int q = 0; q < m.nseg; ++q) a.p[q] = m.segs[q];
  a.set_wd(m.D);
  EpMlp<TPos32> e;
  memset(&e, 0, sizeof(e));
  e.y = m.y; e.bias = m.bias; e.C = m.C_out; e.drop = m.drop; e.res = m.res; e.rrm = m.rrm; e.rac = m.rac; e.stats = m.stats;
  GemmShape sh{m.P, m.C_out, m.nseg * m.D, 1, 1};
  return launch_gemm<TPos32>(a, b, e, sh, stream);
}

struct MlpBwdArgs {
  const float* dh;            // [P, C_out] gradient wrt the (post-dropout) mlp output
  DropoutSrc drop;
  const float* const* segs;   // nseg saved inputs [P, D]
  int nseg;
  i64 P;
  int D, C_out;
  const float* W;
  float* dsegs;               // [nseg][P][D] (written); nullable to skip the data gradient
  float* dW;                  // accumulated (atomic); nullable
  float* dbias;               // accumulated; nullable iff dW is
  const float* WT;            // nullable: W transposed [nseg*D][C_out] -> tcgen05/TMA input-gradient kernel
  const float* WT_lo;         // nullable: its 3xTF32 remainders (fp32x3 tier)
  TcScratch ts;               // partial-result scratch of the tcgen05 weight-gradient reduction (null: not available)
};
inline int mlp_backward(const MlpBwdArgs& m, cudaStream_t stream) {
  GWN_CHECK_ARG(m.nseg >= 1 && m.nseg <= MAXSEG && m.D % 4 == 0 && m.C_out % 4 == 0,
                "mlp bwd: unsupported shape (nseg=%d D=%d C_out=%d)", m.nseg, m.D, m.C_out);
  const int Ktot = m.nseg * m.D;
  if (m.dsegs) {
    ProfScope prof("gcn_mlp_dgrad", stream, 4.0 * m.P * ((double)Ktot + m.C_out), 2.0 * m.P * Ktot * m.C_out);
    LdWN b;
    memset(&b, 0, sizeof(b));
    b.p[0] = m.W; b.set_wd(Ktot); b.ldw = Ktot;
    EpRows e;
    memset(&e, 0, sizeof(e));
    e.y = m.dsegs; e.M = m.P; e.set_seg(m.D);
    int st = -1;
    if (m.WT && m.C_out == 32 && m.drop.mode == GWN_DROPOUT_NONE && Ktot <= 256 && Ktot % 16 == 0 && m.P < 2147483647LL) {
      TcPosArgs t;
      memset(&t, 0, sizeof(t));
      t.seg[0] = TcPosSeg{m.dh, (int)m.P, 32, 0, 0};
      t.nseg = 1; t.nb = 1; t.rows_out = (int)m.P; t.Wp = m.WT; t.N = Ktot;
      t.out = m.dsegs; t.out_width = 32; t.out_nblk = m.nseg; t.out_blk_dim2 = 1; t.Wp_lo = m.WT_lo;
      RowSeg rs;
      memset(&rs, 0, sizeof(rs));
      rs.out = m.dsegs; rs.M = m.P;
      st = launch_tcpos<0>(t, rs, stream);
      if (st > 0) return st;
    }
    if (st < 0 && current_math() != 0 && m.C_out == PG_WD && m.drop.mode == GWN_DROPOUT_NONE && Ktot <= 256) {
      ARows ar;
      memset(&ar, 0, sizeof(ar));
      ar.P[0] = m.dh;
      ar.rm[0] = ar.rm[1] = rowmap_identity();
      ar.nseg = 1; ar.rs = m.C_out;
      st = launch_posgemm<TPG, 32>(ar, b, e, m.P, Ktot, stream);
      if (st > 0) return st;
    }
    if (st < 0) {
      LdRows a;
      memset(&a, 0, sizeof(a));
      a.p[0] = m.dh; a.set_wd(m.C_out); a.drop = m.drop;
      GemmShape sh{m.P, Ktot, m.C_out, 1, 1};
      GWN_TRY((launch_gemm<TBig>(a, b, e, sh, stream)));
    }
  }
  ProfScope prof_w("gcn_mlp_wgrad", stream, m.dW ? 4.0 * m.P * ((double)Ktot + m.C_out) : 0.0,
                   m.dW ? 2.0 * m.P * (Ktot + 1.0) * m.C_out : 0.0);
  if (m.dW && current_math() != 0 && m.D == 32 && m.C_out == 32 && m.nseg <= 7 && m.drop.mode == GWN_DROPOUT_NONE &&
      m.P < 2147483647LL && m.ts.partial) {   // tf32 tier: tcgen05 + TMA reduction (weights and, through the all-ones block, the bias)
#if !GWN_EMU
    TcRedArgs t;
    memset(&t, 0, sizeof(t));
    t.mode = 0; t.na = m.nseg; t.x3 = m.ts.x3;
    for (int q = 0; q < m.nseg; ++q) t.a[q] = TcRedSrc{m.segs[q], (int)m.P, 32, 0, 0, 0};
    t.b[0] = TcRedSrc{m.dh, (int)m.P, 32, 0, 0, 0};
    t.N = 32; t.nb = 1; t.rows = (int)m.P; t.partial = m.ts.partial; t.partial_floats = m.ts.floats;
    TcRedResult r;
    int st = launch_tcred(t, stream, &r);
    if (st > 0) return st;
    if (st == 0) {
      tc::SlotMlpOut f{m.dW, m.dbias, Ktot, m.nseg};
      return launch_slot_reduce(m.ts.partial, r, (i64)m.nseg * 32 * 32 + 32, f, stream);
    }
#endif
  }
  if (m.dW) {
    LdCols a;
    memset(&a, 0, sizeof(a));
    a.p[0] = m.dh; a.set_wd(m.C_out); a.nseg = 1; a.drop = m.drop;
    LdCols b;
    memset(&b, 0, sizeof(b));
    for (int q = 0; q < m.nseg; ++q) b.p[q] = m.segs[q];
    b.set_wd(m.D); b.nseg = m.nseg; b.ones = 1;
    EpWgrad e;
    memset(&e, 0, sizeof(e));
    e.dw[0] = m.dW; e.db[0] = m.dbias; e.set_wd(Ktot); e.nseg = 1; e.ldw = Ktot; e.nbias = 1;
    GWN_CHECK_ARG(m.P < 2147483647LL, "mlp bwd: too many positions");
    GemmShape sh{(i64)m.C_out, Ktot + 1, (int)m.P, pick_ksplit(m.C_out, Ktot + 1, m.P, TW32::BM, TW32::BN, kTargetBlocks), 1};
    GWN_TRY((launch_gemm<TW32>(a, b, e, sh, stream)));
  }
  return 0;
}

// ---------------------------------------------------------------------------------- gcn (model.py:41-55)
struct GcnShape {
  int B, L, V, D, C_out, S, order;
};
inline int hop_index(const GcnShape& g, int s, int k) { return 1 + s * g.order + (k - 1); }  // cat order (model.py:42-52)

// hops[q-1] (q >= 1) receives the q-th concatenated tensor; x is segment 0.
// hop_stride (floats between consecutive hop tensors; 0 = B*L*V*D) lets a caller run one sample of a larger batch.
inline int gcn_hops_forward(const GcnShape& g, const float* x, const SupportView* sup_fwd, float* hops,
                            cudaStream_t stream, const TcSupports* tcs = nullptr, i64 hop_stride = 0) {
  const i64 PD = hop_stride > 0 ? hop_stride : (i64)g.B * g.L * g.V * g.D;
  ProfScope prof("nconv_fwd", stream, 4.0 * PD * (1 + g.S + 2.0 * g.S * (g.order - 1)), 2.0 * PD * g.V * g.S * g.order);
  if (tcs && tcs->precision == GWN_PREC_TF32 && !tcs->per_sample && g.order == 2 && g.D == 32 && hop_stride == 0 && g.S <= TC_MAXSUP) {
    // tf32 tier, small graph: both hops of all supports in ONE launch (support resident in shared memory, hop 2 from TMEM)
    float* Y1[MAXSUP];
    float* Y2[MAXSUP];
    for (int s = 0; s < g.S; ++s) {
      Y1[s] = hops + (i64)(hop_index(g, s, 1) - 1) * PD;
      Y2[s] = hops + (i64)(hop_index(g, s, 2) - 1) * PD;
    }
    const int st = gcn_hops_fused_tc(x, tcs->S, g.S, tcs->ld, Y1, Y2, g.B, g.L, g.V, stream);
    if (st >= 0) return st;
  }
  for (int k = 1; k <= g.order; ++k) {
    const float* X[MAXSUP];
    float* Y[MAXSUP];
    for (int s = 0; s < g.S; ++s) {
      X[s] = (k == 1) ? x : hops + (i64)(hop_index(g, s, k - 1) - 1) * PD;
      Y[s] = hops + (i64)(hop_index(g, s, k) - 1) * PD;
    }
    GWN_TRY(node_gemm(sup_fwd, g.S, false, X, Y, nullptr, nullptr, g.B, g.L, 0, g.V, g.D, stream, tcs));
  }
  return 0;
}

// Backward through the diffusion chain.  dsegs: [1+S*order][P][D] gradient wrt each concatenated segment
// (overwritten in place by the chained t tensors).  dx = dseg_0 + sum_s A_s t_{s,1} (+ window(add2)).
// dsup[s] (nullable) += sum_k hop_{s,k-1}^T t_{s,k}.
inline int gcn_hops_backward(const GcnShape& g, const float* x, const float* hops, const SupportView* sup_bwd, float* dsegs,
                             float* dx, const float* add2, int T_out, float* const* dsup, const i64* ldds,
                             cudaStream_t stream, const TcSupports* tcs = nullptr, const TcScratch* ts = nullptr,
                             i64 hop_stride = 0) {
  const i64 PD = hop_stride > 0 ? hop_stride : (i64)g.B * g.L * g.V * g.D;
  {
    ProfScope prof("nconv_bwd_dx_hops", stream, 4.0 * PD * 3.0 * g.S * (g.order - 1), 2.0 * PD * g.V * g.S * (g.order - 1));
    for (int k = g.order; k >= 2; --k) {
      const float* X[MAXSUP];
      float* Y[MAXSUP];
      const float* A[MAXSUP];
      for (int s = 0; s < g.S; ++s) {
        X[s] = dsegs + (i64)hop_index(g, s, k) * PD;
        Y[s] = dsegs + (i64)hop_index(g, s, k - 1) * PD;
        A[s] = Y[s];
      }
      GWN_TRY(node_gemm(sup_bwd, g.S, false, X, Y, A, nullptr, g.B, g.L, 0, g.V, g.D, stream, tcs));
    }
  }
  for (int s = 0; s < g.S; ++s) {
    if (!dsup || !dsup[s]) continue;
    ProfScope prof("nconv_bwd_dA", stream, 4.0 * PD * 2.0 * g.order, 2.0 * PD * g.V * g.order);
    const float* Xp[MAXSUP];
    const float* Yp[MAXSUP];
    GWN_CHECK_ARG(g.order <= MAXSUP, "gcn bwd: order too large");
    for (int k = 1; k <= g.order; ++k) {
      Xp[k - 1] = (k == 1) ? x : hops + (i64)(hop_index(g, s, k - 1) - 1) * PD;
      Yp[k - 1] = dsegs + (i64)hop_index(g, s, k) * PD;
    }
    GWN_TRY(support_grad_gemm(Xp, Yp, g.order, dsup[s], ldds[s], g.B, g.L, g.V, g.D, stream, ts));
  }
  {
    ProfScope prof("nconv_bwd_dx_sum", stream, 4.0 * PD * (g.S + 2.0) + 4.0 * PD / g.L * T_out, 2.0 * PD * g.V * g.S);
    const float* X[MAXSUP];
    for (int s = 0; s < g.S; ++s) X[s] = dsegs + (i64)hop_index(g, s, 1) * PD;
    float* Y[1] = {dx};
    const float* A[1] = {dsegs};
    GWN_TRY(node_gemm(sup_bwd, g.S, true, X, Y, A, add2, g.B, g.L, T_out, g.V, g.D, stream, tcs));
  }
  return 0;
}

}  // namespace gwn
