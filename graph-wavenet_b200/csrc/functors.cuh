// Operand loaders and epilogues plugged into gemm.cuh for every contraction on the hot path.
// Tensor layout is BLNC fp32 (see include/gwnet_b200.h): an activation is a matrix of
// P = B*L*N position rows by C channels; the node contraction views it as B*L slabs of [N x C].
#pragma once
#include "gemm.cuh"

namespace gwn {

constexpr int MAXSUP = 8;    // supports per gcn
constexpr int MAXSEG = 32;   // K / N segments (hops of a gcn, layers of the skip sum)

GWN_HD void zero4(float (&v)[4]) { v[0] = v[1] = v[2] = v[3] = 0.0f; }
GWN_HD void get4(float (&v)[4], float4 f) { v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }

// ====================================================================== node contraction operands
// Support matrix as the A operand, K-outer: op(k, m) = S[k*rs + m*cs].
// nconv forward (model.py:13): k = v, m = w, S = A.   dX = A.dY: k = w, m = v -> pass swapped strides.
// `kcat`: K is the concatenation over `nsup` supports of kper rows each (sum over supports in one GEMM).
struct LdSupport {
  static constexpr bool kInner = false;
  const float* p[MAXSUP];
  i64 rs[MAXSUP], cs[MAXSUP];
  int xlim[MAXSUP];  // readable extent along m when cs == 1 (padded buffers: ld; else M)
  int vec[MAXSUP];   // cs == 1, rs % 4 == 0, base 16-byte aligned
  int kper, kcat;
  int bz_;
  GWN_HD void init(int bz) { bz_ = kcat ? 0 : bz; }
  GWN_HD void load4(float (&v)[4], int k, i64 m, int K, i64 M) const {
    zero4(v);
    if (k >= K || m >= M) return;
    int s = bz_, kk = k;
    if (kcat) { s = k / kper; kk = k - s * kper; }
    const float* q = p[s] + (i64)kk * rs[s];
    if (vec[s] && m + 3 < xlim[s]) {
      get4(v, ld4(q + m));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (m + i < M) v[i] = q[(m + i) * cs[s]];
    }
  }
};

// Slab tensor as the B operand, K-outer: op(k = node, n = (slab, c)) = X[(slab*V + node)*C + c].
struct LdSlab {
  static constexpr bool kInner = false;
  const float* p[MAXSUP];
  int V, C, kper, kcat;
  int bz_;
  GWN_HD void init(int bz) { bz_ = kcat ? 0 : bz; }
  GWN_HD void load4(float (&v)[4], int k, i64 n, int K, i64 N) const {
    zero4(v);
    if (k >= K || n >= N) return;
    int s = bz_, kk = k;
    if (kcat) { s = k / kper; kk = k - s * kper; }
    i64 slab = n / C;
    int c = (int)(n - slab * C);
    get4(v, ld4(p[s] + (slab * V + kk) * C + c));
  }
};

// Store of a node-contraction result: y[(slab*V + m)*C + c] = v (+ add) (+ head-window add2).
struct EpSlab {
  static constexpr bool kHasFinish = false;
  float* y[MAXSUP];
  const float* add[MAXSUP];  // same layout as y, nullable
  const float* add2;         // nullable: [B][T_out][V][C] added where l >= L - T_out
  int V, C, L, T_out;
  int bz_;
  GWN_HD void init(int bz) { bz_ = bz; }
  GWN_HD void store4(i64 m, int n, const float (&v)[4], int, int) const {
    int slab = n / C, c = n - slab * C;
    i64 idx = ((i64)slab * V + m) * C + c;
    float4 r = make_float4(v[0], v[1], v[2], v[3]);
    if (add[bz_]) {
      float4 a = ld4(add[bz_] + idx);
      r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
    }
    if (add2) {
      int b = slab / L, l = slab - b * L;
      if (l >= L - T_out) {
        float4 a = ld4(add2 + (((i64)b * T_out + (l - (L - T_out))) * V + m) * C + c);
        r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
      }
    }
    st4(y[bz_] + idx, r);
  }
  template <int MATH>
  GWN_DEV void finish(float*, int) const {}
};

// ====================================================================== position-row operands
// K-inner rows: op(k = (q, ci), m = p) = seg[q][row_q(p)*wd + ci]   (optionally * keep-mask, * a + c).
struct LdRows {
  static constexpr bool kInner = true;
  const float* p[MAXSEG];
  Remap rm[MAXSEG];
  int wd, use_remap;
  DivW wdiv;         // set by set_wd()
  const float* ac;   // nullable BN fold: x*ac[ci] + ac[wd+ci]
  DropoutSrc drop;   // single-segment only
  void set_wd(int w) { wd = w; wdiv = make_divw(w); }
  GWN_HD void init(int) {}
  GWN_HD void load4(float (&v)[4], int k, i64 m, int K, i64 M) const {
    zero4(v);
    if (k >= K || m >= M) return;
    int q = wdiv.div(k), ci = k - q * wd;
    i64 row = use_remap ? rm[q](m) : m;
    i64 e = row * wd + ci;
    get4(v, ld4(p[q] + e));
    if (ac) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaf(v[i], ac[ci + i], ac[wd + ci + i]);
    }
    if (drop.mode != GWN_DROPOUT_NONE) {
      float kp[4];
      drop.keep4(e, kp);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] *= kp[i];
    }
  }
};

// K-outer columns: op(k = p, x = (q, ci)) = seg[q][row_q(p)*wd + ci]; with `ones`, column nseg*wd reads 1
// (bias-gradient trick: the weight-gradient GEMM then also yields the column sums).
struct LdCols {
  static constexpr bool kInner = false;
  const float* p[MAXSEG];
  Remap rm[MAXSEG];
  int wd, nseg, use_remap, ones;
  DivW wdiv;
  const float* ac;
  DropoutSrc drop;
  void set_wd(int w) { wd = w; wdiv = make_divw(w); }
  GWN_HD void init(int) {}
  GWN_HD void load4(float (&v)[4], int k, i64 x, int K, i64 X) const {
    zero4(v);
    if (k >= K || x >= X) return;
    if (x >= (i64)nseg * wd) {
      if (ones) v[0] = 1.0f;
      return;
    }
    int q = wdiv.div((int)x), ci = (int)x - q * wd;
    i64 row = use_remap ? rm[q]((i64)k) : (i64)k;
    i64 e = row * wd + ci;
    get4(v, ld4(p[q] + e));
    if (ac) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaf(v[i], ac[ci + i], ac[wd + ci + i]);
    }
    if (drop.mode != GWN_DROPOUT_NONE) {
      float kp[4];
      drop.keep4(e, kp);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] *= kp[i];
    }
  }
};

// Gradient of the (1,2) dilated conv wrt its input, A operand (K-inner):
// op(k = (tap, j), m = p_in=(b,t',n)) = dpre[(b, t' - tap*d, n)*W2 + j] when 0 <= t' - tap*d < L_out else 0.
struct LdDpreTaps {
  static constexpr bool kInner = true;
  const float* dpre;
  int W2, N, L_in, L_out, d;
  GWN_HD void init(int) {}
  GWN_HD void load4(float (&v)[4], int k, i64 m, int K, i64 M) const {
    zero4(v);
    if (k >= K || m >= M) return;
    int tap = k / W2, j = k - tap * W2;
    i64 lon = (i64)L_in * N;
    i64 b = m / lon;
    int r = (int)(m - b * lon);
    int tp = r / N, n = r - tp * N;
    int t = tp - tap * d;
    if (t < 0 || t >= L_out) return;
    get4(v, ld4(dpre + ((b * L_out + t) * N + n) * W2 + j));
  }
};

// Network input [B,F,N,T] (any strides) as K-outer columns of the zero-padded time axis:
// op(k = p=(b,t,n) over L0 = T + pad, x = f) = t >= pad ? in[b,f,n,t-pad] : 0 ; column F reads 1 (bias).
struct LdInputCols {
  static constexpr bool kInner = false;
  const float* in;
  i64 sb, sf, sn, st;
  int F, N, L0, pad, ones;
  GWN_HD void init(int) {}
  GWN_HD void load4(float (&v)[4], int k, i64 x, int K, i64 X) const {
    zero4(v);
    if (k >= K || x >= X) return;
    i64 lon = (i64)L0 * N;
    i64 b = k / lon;
    int r = (int)(k - b * lon);
    int t = r / N, n = r - t * N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      i64 f = x + i;
      if (f < F) {
        if (t >= pad) v[i] = in[b * sb + f * sf + n * sn + (t - pad) * st];
      } else if (f == F && ones) {
        v[i] = 1.0f;
      }
    }
  }
};

// ====================================================================== weight operands
// K-inner weight: op(k = (q, kk), n) = W[q][n*ldw + kk]   (row-major [n_out, k] per segment).
struct LdWK {
  static constexpr bool kInner = true;
  const float* p[MAXSEG];
  int wd, ldw;
  DivW wdiv;
  void set_wd(int w) { wd = w; wdiv = make_divw(w); }
  GWN_HD void init(int) {}
  GWN_HD void load4(float (&v)[4], int k, i64 n, int K, i64 N) const {
    zero4(v);
    if (k >= K || n >= N) return;
    int q = wdiv.div(k), kk = k - q * wd;
    const float* w = p[q] + n * ldw + kk;
    if (k + 3 < K) {
      get4(v, ld4(w));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (k + i < K) v[i] = w[i];
    }
  }
};

// K-outer weight: op(k, n = (q, nn)) = W[q][k*ldw + nn].
struct LdWN {
  static constexpr bool kInner = false;
  const float* p[MAXSEG];
  int wd, ldw;
  DivW wdiv;
  void set_wd(int w) { wd = w; wdiv = make_divw(w); }
  GWN_HD void init(int) {}
  GWN_HD void load4(float (&v)[4], int k, i64 n, int K, i64 N) const {
    zero4(v);
    if (k >= K || n >= N) return;
    int q = wdiv.div((int)n), nn = (int)n - q * wd;
    const float* w = p[q] + (i64)k * ldw + nn;
    if (n + 3 < N) {
      get4(v, ld4(w));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n + i < N) v[i] = w[i];
    }
  }
};

// Filter/gate weights [D,C,1,2] (model.py:135-141) as the B operand of the fused gated conv:
// op(k = (tap, ci), n = 2*ch + g) = (g ? Wg : Wf)[(ch*C + ci)*2 + tap].
struct LdWTcn {
  static constexpr bool kInner = true;
  const float* wf;
  const float* wg;
  int C;
  GWN_HD void init(int) {}
  GWN_HD void load4(float (&v)[4], int k, i64 n, int K, i64 N) const {
    zero4(v);
    if (k >= K || n >= N) return;
    int tap = k / C, ci = k - tap * C;
    int ch = (int)(n >> 1);
    const float* w = ((n & 1) ? wg : wf) + ((i64)ch * C + ci) * 2 + tap;
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = w[2 * i];
  }
};
// Same weights, transposed role (input gradient): op(k = (tap, j = 2*ch+g), n = ci) = W_g[(ch*C + ci)*2 + tap].
struct LdWTcnT {
  static constexpr bool kInner = false;
  const float* wf;
  const float* wg;
  int C, W2;
  GWN_HD void init(int) {}
  GWN_HD void load4(float (&v)[4], int k, i64 n, int K, i64 N) const {
    zero4(v);
    if (k >= K || n >= N) return;
    int tap = k / W2, j = k - tap * W2;
    int ch = j >> 1;
    const float* w = ((j & 1) ? wg : wf) + ((i64)ch * C + n) * 2 + tap;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n + i < N) v[i] = w[2 * i];
  }
};

// ====================================================================== epilogues
// Row-major store with optional summed biases, relu, and relu-backward gate; or segmented store
// y[(q*M + m)*seg_wd + nn] (n = (q, nn)) when seg_wd > 0.
struct EpRows {
  static constexpr bool kHasFinish = false;
  float* y;
  i64 ldy, M;
  const float* bias[MAXSEG];
  int nbias, relu, seg_wd;
  DivW segdiv;
  void set_seg(int w) { seg_wd = w; segdiv = make_divw(w); }
  const float* gate;  // nullable: multiply by (gate[m*ldy+n] > 0)
  GWN_HD void init(int) {}
  GWN_HD void store4(i64 m, int n, const float (&v)[4], int nvalid, int) const {
    float r[4] = {v[0], v[1], v[2], v[3]};
    for (int q = 0; q < nbias; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < nvalid) r[i] += bias[q][n + i];
    if (relu)
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = fmaxf(r[i], 0.0f);
    i64 idx;
    if (seg_wd > 0) {
      int q = segdiv.div(n);
      idx = ((i64)q * M + m) * seg_wd + (n - q * seg_wd);
    } else {
      idx = m * ldy + n;
    }
    if (gate)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < nvalid) r[i] = gate[idx + i] > 0.0f ? r[i] : 0.0f;
    if (nvalid == 4) {
      st4(y + idx, make_float4(r[0], r[1], r[2], r[3]));
    } else {
      for (int i = 0; i < nvalid; ++i) y[idx + i] = r[i];
    }
  }
  template <int MATH>
  GWN_DEV void finish(float*, int) const {}
};

// end_conv_2 (model.py:240): bias + store straight into the reference's NCHW output [B,O,N,T_out].
struct EpNCHW {
  static constexpr bool kHasFinish = false;
  float* y;
  const float* bias;
  int N, T;          // nodes, T_out
  i64 sb, so, sn, st;
  GWN_HD void init(int) {}
  GWN_HD void store4(i64 m, int n, const float (&v)[4], int nvalid, int) const {
    i64 lon = (i64)T * N;
    i64 b = m / lon;
    int r = (int)(m - b * lon);
    int t = r / N, node = r - t * N;
    for (int i = 0; i < nvalid; ++i) y[b * sb + (n + i) * so + node * sn + t * st] = v[i] + bias[n + i];
  }
  template <int MATH>
  GWN_DEV void finish(float*, int) const {}
};

// Gated activation (model.py:208-212): columns come interleaved (f0,g0,f1,g1,...).
struct EpGate {
  static constexpr bool kHasFinish = false;
  float* y;  // [P, D]
  const float* bf;
  const float* bg;
  int D;
  GWN_HD void init(int) {}
  GWN_HD void store4(i64 m, int n, const float (&v)[4], int nvalid, int) const {
    int ch = n >> 1;
    float o0 = tanhf(v[0] + bf[ch]) * sigmoidf_(v[1] + bg[ch]);
    y[m * D + ch] = o0;
    if (nvalid == 4) y[m * D + ch + 1] = tanhf(v[2] + bf[ch + 1]) * sigmoidf_(v[3] + bg[ch + 1]);
  }
  template <int MATH>
  GWN_DEV void finish(float*, int) const {}
};

// Backward of the gate: from recomputed pre-activations and dg, write dpre (interleaved, [P, 2D]).
struct EpGateBwd {
  static constexpr bool kHasFinish = false;
  float* dpre;
  const float* dg;  // [P, D]
  const float* bf;
  const float* bg;
  int D;
  GWN_HD void init(int) {}
  GWN_HD void one(i64 m, int ch, float pf, float pg, float* out) const {
    float f = tanhf(pf + bf[ch]), s = sigmoidf_(pg + bg[ch]);
    float g = dg[m * D + ch];
    out[0] = g * s * (1.0f - f * f);
    out[1] = g * f * s * (1.0f - s);
  }
  GWN_HD void store4(i64 m, int n, const float (&v)[4], int nvalid, int) const {
    int ch = n >> 1;
    float* o = dpre + m * (2 * D) + n;
    one(m, ch, v[0], v[1], o);
    if (nvalid == 4) one(m, ch + 1, v[2], v[3], o + 2);
  }
  template <int MATH>
  GWN_DEV void finish(float*, int) const {}
};

// Column partial sums (two per column) reduced over the block, then added to global doubles.
template <class T>
struct ColStats {
  float s1[T::SLOTS * 4], s2[T::SLOTS * 4];   // indexed by (slot, i): the 4-wide column groups this thread owns
  GWN_HD void reset() {
#pragma unroll
    for (int i = 0; i < T::SLOTS * 4; ++i) s1[i] = s2[i] = 0.0f;
  }
  // Row-owner variant (tcpos.cuh): each of the 128 epilogue threads owns whole rows, i.e. holds a partial for every
  // one of the (<= 32) columns, indexed directly by column.  `etid` = thread index within the 4 epilogue warps.
  GWN_DEV void reduce_rows(float* smem, int etid, double* g1, double* g2, int ncols) {
#if !GWN_EMU
    if (etid < 64) smem[etid] = 0.0f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
    for (int col = 0; col < T::SLOTS * 4; ++col) {
      float a = s1[col], b = s2[col];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if ((etid & 31) == 0 && col < 32) {
        atomicAdd(smem + col, a);
        atomicAdd(smem + 32 + col, b);
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (etid < 32 && etid < ncols) {
      atomicAdd(g1 + etid, (double)smem[etid]);
      atomicAdd(g2 + etid, (double)smem[32 + etid]);
    }
#endif
  }
  template <int MATH>
  GWN_DEV void reduce(float* smem, int tid, double* g1, double* g2, int ncols) {
#if !GWN_EMU
    float* c1 = smem;
    float* c2 = smem + T::BN;
    for (int i = tid; i < 2 * T::BN; i += T::NT) smem[i] = 0.0f;
    __syncthreads();
    // Lanes of a warp that own the same columns are summed with shuffles first; a float atomicAdd on shared
    // memory is a CAS loop, and 16 lanes hitting one address made it 40 % of this kernel's instructions.
    const int lane = tid & 31;
    if (MATH == 0) {
      const int tx = tid % T::TX;
#pragma unroll
      for (int gn = 0; gn < T::GN; ++gn)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = s1[gn * 4 + j], b = s2[gn * 4 + j];
#pragma unroll
          for (int o = T::TX; o < 32; o <<= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
          }
          if (T::TX >= 32 || lane < T::TX) {
            int col = gn * (T::BN / T::GN) + tx * 4 + j;
            atomicAdd(c1 + col, a);
            atomicAdd(c2 + col, b);
          }
        }
    } else {   // tensor mode: slot = n8 tile of the warp; lanes with equal (lane >> 1) & 1 own the same 4 columns
      const int wn = (tid >> 5) / T::WM, t = tid & 3;
#pragma unroll
      for (int sl = 0; sl < T::NTL; ++sl)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = s1[sl * 4 + j], b = s2[sl * 4 + j];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            if (o == 2) continue;
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
          }
          if ((lane & 29) == 0) {
            int col = wn * T::WTN + sl * 8 + (t >> 1) * 4 + j;
            atomicAdd(c1 + col, a);
            atomicAdd(c2 + col, b);
          }
        }
    }
    __syncthreads();
    const int n0 = blockIdx.y * T::BN;
    for (int i = tid; i < T::BN; i += T::NT) {
      if (n0 + i < ncols) {
        atomicAdd(g1 + n0 + i, (double)c1[i]);
        atomicAdd(g2 + n0 + i, (double)c2[i]);
      }
    }
#endif
  }
};

// gcn tail (model.py:53-54) fused with the residual add and BatchNorm statistics (model.py:234-236):
//   h = (acc + bias) * keep ;  u = h + fold(res[row(p)]) ;  store u ;  stats += (u, u^2)
template <class T>
struct EpMlp {
  static constexpr bool kHasFinish = true;
  float* y;            // [P, C]
  const float* bias;
  int C;
  DropoutSrc drop;
  const float* res;    // nullable residual source (pre-BN tensor of the previous layer)
  Remap rrm;
  const float* rac;    // nullable fold of the residual
  double* stats;       // nullable: [2*C] sum, sum of squares
  ColStats<T> cs;
  GWN_HD void init(int) { cs.reset(); }
  GWN_HD void store4(i64 m, int n, const float (&v)[4], int nvalid, int gn) {
    float r[4] = {v[0], v[1], v[2], v[3]};
    i64 e = m * C + n;
    float kp[4];
    drop.keep4(e, kp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < nvalid) r[i] = (r[i] + bias[n + i]) * kp[i];
    if (res) {
      i64 ridx = rrm(m) * C + n;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < nvalid) {
          float x = res[ridx + i];
          if (rac) x = fmaf(x, rac[n + i], rac[C + n + i]);
          r[i] += x;
        }
    }
    for (int i = 0; i < nvalid; ++i) y[e + i] = r[i];
    if (stats) {
#if GWN_EMU
      for (int i = 0; i < nvalid; ++i) { stats[n + i] += r[i]; stats[C + n + i] += (double)r[i] * r[i]; }
#else
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < nvalid) { cs.s1[gn * 4 + i] += r[i]; cs.s2[gn * 4 + i] += r[i] * r[i]; }
#endif
    }
  }
  template <int MATH>
  GWN_DEV void finish(float* smem, int tid) {
    if (stats) cs.template reduce<MATH>(smem, tid, stats, stats + C, C);
  }
  GWN_DEV void finish_rows(float* smem, int etid) {
    if (stats) cs.reduce_rows(smem, etid, stats, stats + C, C);
  }
};

// Input gradient of the gated conv + residual path + BatchNorm-backward statistics of the layer below:
//   dx = acc + (t' >= L_in - L_out ? du[(b, t' - (L_in-L_out), n)] : 0) ; store dx ;
//   bsum += (dx, dx * xhat) with xhat = (u_prev - mean) * rstd.
template <class T>
struct EpTcnDgrad {
  static constexpr bool kHasFinish = true;
  float* dx;           // [P_in, C]
  const float* du;     // nullable [P_out, C]
  int C, N, L_in, L_out;
  const float* uprev;  // nullable (layer 0 has no BN below)
  const float* mr;     // mean[C], rstd[C]
  double* bsum;        // [2*C]
  ColStats<T> cs;
  GWN_HD void init(int) { cs.reset(); }
  GWN_HD void store4(i64 m, int n, const float (&v)[4], int nvalid, int gn) {
    float r[4] = {v[0], v[1], v[2], v[3]};
    if (du) {
      i64 lon = (i64)L_in * N;
      i64 b = m / lon;
      int rr = (int)(m - b * lon);
      int tp = rr / N, node = rr - tp * N;
      int t = tp - (L_in - L_out);
      if (t >= 0) {
        const float* s = du + ((b * L_out + t) * N + node) * C + n;
        for (int i = 0; i < nvalid; ++i) r[i] += s[i];
      }
    }
    i64 e = m * C + n;
    for (int i = 0; i < nvalid; ++i) dx[e + i] = r[i];
    if (uprev) {
      for (int i = 0; i < nvalid; ++i) {
        float xh = (uprev[e + i] - mr[n + i]) * mr[C + n + i];
#if GWN_EMU
        bsum[n + i] += r[i];
        bsum[C + n + i] += (double)r[i] * xh;
#else
        cs.s1[gn * 4 + i] += r[i];
        cs.s2[gn * 4 + i] += r[i] * xh;
#endif
      }
    }
  }
  template <int MATH>
  GWN_DEV void finish(float* smem, int tid) {
    if (uprev) cs.template reduce<MATH>(smem, tid, bsum, bsum + C, C);
  }
  GWN_DEV void finish_rows(float* smem, int etid) {
    if (uprev) cs.reduce_rows(smem, etid, bsum, bsum + C, C);
  }
};

// Accumulating (atomic) weight-gradient epilogue: dW[q][m*ldw + nn] += v for n = (q, nn) < nseg*wd;
// the extra column n == nseg*wd carries the bias gradient, added to every db[q][m].
struct EpWgrad {
  static constexpr bool kHasFinish = false;
  float* dw[MAXSEG];
  float* db[MAXSEG];
  int wd, nseg, ldw, nbias;
  DivW wdiv;
  void set_wd(int w) { wd = w; wdiv = make_divw(w); }
  GWN_HD void init(int) {}
  GWN_HD void store4(i64 m, int n, const float (&v)[4], int nvalid, int) const {
    for (int i = 0; i < nvalid; ++i) {
      int nn = n + i;
      if (nn < nseg * wd) {
        int q = wdiv.div(nn);
        atomic_add_f(dw[q] + m * ldw + (nn - q * wd), v[i]);
      } else {
        for (int q = 0; q < nbias; ++q) atomic_add_f(db[q] + m, v[i]);
      }
    }
  }
  template <int MATH>
  GWN_DEV void finish(float*, int) const {}
};

// Weight gradient of the fused gated conv: m = j = 2*ch+g, n = (tap, ci) -> dW_g[(ch*C+ci)*2+tap]; n == 2C -> db_g[ch].
struct EpWgradTcn {
  static constexpr bool kHasFinish = false;
  float* dwf;
  float* dwg;
  float* dbf;
  float* dbg;
  int C;
  GWN_HD void init(int) {}
  GWN_HD void store4(i64 m, int n, const float (&v)[4], int nvalid, int) const {
    int ch = (int)(m >> 1), g = (int)(m & 1);
    for (int i = 0; i < nvalid; ++i) {
      int nn = n + i;
      if (nn < 2 * C) {
        int tap = nn / C, ci = nn - tap * C;
        atomic_add_f((g ? dwg : dwf) + ((i64)ch * C + ci) * 2 + tap, v[i]);
      } else {
        atomic_add_f((g ? dbg : dbf) + ch, v[i]);
      }
    }
  }
  template <int MATH>
  GWN_DEV void finish(float*, int) const {}
};

// dA[m*ld + n] += v  (adaptive-adjacency gradient)
struct EpAtomicMat {
  static constexpr bool kHasFinish = false;
  float* y;
  i64 ld;
  GWN_HD void init(int) {}
  GWN_HD void store4(i64 m, int n, const float (&v)[4], int nvalid, int) const {
    for (int i = 0; i < nvalid; ++i) atomic_add_f(y + m * ld + n + i, v[i]);
  }
  template <int MATH>
  GWN_DEV void finish(float*, int) const {}
};

// K-inner slab rows for dA: op(k = (pair, slab, c), m = node) = X[pair][(slab*V + node)*C + c].
struct LdSlabK {
  static constexpr bool kInner = true;
  const float* p[MAXSUP];
  int V, C;
  int kper;   // slabs*C per pair (< 2^31, checked on the host)
  DivW cdiv;
  GWN_HD void init(int) {}
  GWN_HD void load4(float (&v)[4], int k, i64 m, int K, i64 M) const {
    zero4(v);
    if (k >= K || m >= M) return;
    int pr = 0, kk = k;
    while (kk >= kper) { kk -= kper; ++pr; }   // a handful of pairs at most
    int slab = cdiv.div(kk);
    int c = kk - slab * C;
    get4(v, ld4(p[pr] + ((i64)slab * V + (int)m) * C + c));
  }
};

}  // namespace gwn
