// Row-owner epilogues of the tcgen05 position GEMM (tcpos.cuh): each of the 128 epilogue threads owns one whole
// position row of the accumulator tile.  Contract:
//     static constexpr int kAccPerBlock;                   // accumulator columns that make one 32-wide output block
//     void init();
//     static constexpr int kAddends;                       // addend tiles TMA stages per row tile (0..2)
//     static constexpr int kGroups;                        // epilogue groups of 4 warps (2 unless registers forbid)
//     void load_addends(const AddendRows& a, bool valid);  // copy the row's addends from the staged tiles to registers
//     static constexpr bool kLazyAddends;                  // true: consume16 reads the staged tiles itself (fewer live
//                                                          // registers); the tiles are handed back after the last block
//     void consume16(i64 m, int c0, int n0, const float (&v)[16], RowSink& out);   // accumulator columns c0..c0+15 of
//                                                          // row m in this tile; n0 = the tile's first output column
//     void finish_rows(float* red, int etid);              // if kHasFinish: block-level reduce of column statistics
// Addends (residual rows, upstream gradients, saved activations) arrive like the GEMM operands: the producer warp
// TMA-loads them as 128-row tiles one tile ahead, so their HBM latency never sits on the epilogue's critical path
// (ncu r01c: with per-thread global loads the single epilogue warp per scheduler was latency-bound, "no eligible
// warp" 81-88 % of cycles).  Outputs go to a 128B-swizzled shared-memory staging tile that one thread hands to TMA
// (cp.async.bulk.tensor store): a row-owner thread storing its row straight to global writes 16 bytes per 128-byte
// line per instruction, and ncu (r01c) showed those kernels bound by L2 sector traffic (L2 72 %, DRAM 22 %).
// Same arithmetic as the functors of functors.cuh.
#pragma once
#include "functors.cuh"

namespace gwn {

#if !GWN_EMU
__device__ __forceinline__ void ld_row32(float (&r)[32], const float* p) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(p) + q);
    r[4 * q] = f.x; r[4 * q + 1] = f.y; r[4 * q + 2] = f.z; r[4 * q + 3] = f.w;
  }
}
// The calling thread's row in up to two staged addend tiles (SWIZZLE_128B, 128 bytes per row).
struct AddendRows {
  const uint8_t* row[2];   // tile + r * 128
  uint32_t x;              // r & 7
  template <int NF>   // NF floats (multiple of 4) starting at column col (multiple of 4)
  __device__ __forceinline__ void load(int which, int col, float (&r)[NF]) const {
#pragma unroll
    for (int q = 0; q < NF / 4; ++q) {
      const float4 f = *reinterpret_cast<const float4*>(row[which] + ((((uint32_t)col >> 2) + q) ^ x) * 16);
      r[4 * q] = f.x; r[4 * q + 1] = f.y; r[4 * q + 2] = f.z; r[4 * q + 3] = f.w;
    }
  }
};

// One row (128 bytes = 32 floats) of the staging tile, SWIZZLE_128B: 16-byte chunk j of row r lives at chunk j ^ (r & 7).
struct RowSink {
  uint8_t* row;   // staging tile + r * 128
  uint32_t x;     // r & 7
  __device__ __forceinline__ void put4(int col, float a, float b, float c, float d) const {   // col: 0..28, multiple of 4
    *reinterpret_cast<float4*>(row + ((((uint32_t)col >> 2) ^ x) << 4)) = make_float4(a, b, c, d);
  }
  __device__ __forceinline__ void put16(int col, const float (&o)[16]) const {
#pragma unroll
    for (int q = 0; q < 4; ++q) put4(col + 4 * q, o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  }
};

// Per-thread column statistics over 32 columns, reduced across the 128 epilogue threads at the end.
struct RowStats32 {
  float s1[32], s2[32];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < 32; ++i) s1[i] = s2[i] = 0.0f;
  }
  // Lane-transposing butterfly: after the 5 exchange steps lane L holds the warp total of column L (31 shuffles per
  // array instead of 160), the four warps meet in shared memory with one conflict-free atomic per lane, and 64
  // threads add the CTA totals to this CTA's replica of the global sums (GWN_STAT_REPL).
  __device__ __forceinline__ static float lane_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int j = 0; j < o; ++j) {
        const float send = up ? v[j] : v[j + o];
        const float keep = up ? v[j + o] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    return v[0];
  }
  __device__ __forceinline__ void reduce(float* smem, int etid, double* g1, double* g2, int ncols, int barid) {
    if (etid < 64) smem[etid] = 0.0f;
    asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
    const int lane = etid & 31;
    const float t1 = lane_transpose_sum(s1, lane), t2 = lane_transpose_sum(s2, lane);
    atomicAdd(smem + lane, t1);
    atomicAdd(smem + 32 + lane, t2);
    asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
    const size_t rep = (size_t)(blockIdx.x % GWN_STAT_REPL) * 64;   // replicas of [2*C] doubles, C == 32 here
    if (etid < 32 && etid < ncols) atomicAdd(g1 + rep + etid, (double)smem[etid]);
    else if (etid >= 32 && etid < 64 && etid - 32 < ncols) atomicAdd(g2 + rep + (etid - 32), (double)smem[etid]);
  }
};

// tanh / sigmoid of the gate on the tf32 tier: ex2.approx-based (2 MUFU + 3 FMA-class instructions each, ~1e-7
// absolute error) instead of the ~35-instruction tanhf / expf sequences -- with one epilogue warp per scheduler the
// accurate versions made the gate epilogue, not HBM, the limiter of the gated-conv kernels (ncu r01b).
// ex2 / rcp are issued as the bare .ftz MUFU forms: __expf and __fdividef wrap the same two instructions in range
// fix-ups (FSETP + predicated FMULs for denormal inputs / huge denominators) that cannot trigger here -- the gate
// saturates: ex2 -> +inf gives rcp(inf) = 0, ex2 -> 0 gives rcp(1) = 1 -- and the gate epilogue is instruction-bound
// (ncu r02x: its 8 warps 85 % busy, the issuer waiting for a free accumulator 55 % of the time).
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kSigmScale = -1.4426950408889634f, kTanhScale = 2.8853900817779268f;   // -log2(e), 2 log2(e)
__device__ __forceinline__ float gate_sigmoid(float x) { return mufu_rcp(1.0f + mufu_ex2(kSigmScale * x)); }
__device__ __forceinline__ float gate_tanh(float x) { return fmaf(-2.0f, mufu_rcp(mufu_ex2(kTanhScale * x) + 1.0f), 1.0f); }
// the same with the bias folded into the exponent: sb = kSigmScale * bias, tb = kTanhScale * bias
__device__ __forceinline__ float gate_sigmoid_b(float v, float sb) { return mufu_rcp(1.0f + mufu_ex2(fmaf(v, kSigmScale, sb))); }
__device__ __forceinline__ float gate_tanh_b(float v, float tb) { return fmaf(-2.0f, mufu_rcp(mufu_ex2(fmaf(v, kTanhScale, tb)) + 1.0f), 1.0f); }
// Two gates with ONE reciprocal (the gate epilogues are MUFU-bound: 4 quarter-rate instructions per gated output with
// the helpers above, 2.5 here).  With ea = e^{2a}, eb = e^{-b}: tanh(a) = (ea - 1) / (1 + ea), sigmoid(b) = 1 / (1 + eb),
// so every factor of a pair is a product of  r = 1 / ((1+ea0)(1+eb0)(1+ea1)(1+eb1))  with the other denominators.
// The exponents are clamped at 30: 2^30 leaves tanh within 2^-29 of 1 and sigmoid within 2^-30 of 0 (below fp32
// resolution of the results) and keeps the four-factor product under 2^124, so r stays a normal number.
struct GatePair { float num[2], q1[2], p1[2], den[2], r; };   // num = ea - 1, q1 = 1 + ea, p1 = 1 + eb, den = q1 * p1
__device__ __forceinline__ GatePair gate_pair(float vf0, float vg0, float vf1, float vg1, float tb0, float sb0, float tb1,
                                              float sb1) {
  GatePair g;
  const float ea0 = mufu_ex2(fminf(fmaf(vf0, kTanhScale, tb0), 30.0f)), eb0 = mufu_ex2(fminf(fmaf(vg0, kSigmScale, sb0), 30.0f));
  const float ea1 = mufu_ex2(fminf(fmaf(vf1, kTanhScale, tb1), 30.0f)), eb1 = mufu_ex2(fminf(fmaf(vg1, kSigmScale, sb1), 30.0f));
  g.num[0] = ea0 - 1.0f; g.q1[0] = ea0 + 1.0f; g.p1[0] = eb0 + 1.0f; g.den[0] = g.q1[0] * g.p1[0];
  g.num[1] = ea1 - 1.0f; g.q1[1] = ea1 + 1.0f; g.p1[1] = eb1 + 1.0f; g.den[1] = g.q1[1] * g.p1[1];
  g.r = mufu_rcp(g.den[0] * g.den[1]);
  return g;
}

// model.py:208-212 -- accumulator columns interleaved (f0,g0,f1,g1,...), N = 64 -> 32 gated outputs.
struct RowGate {
  static constexpr bool kLazyAddends = false;
  static constexpr bool kHasFinish = false;
  static constexpr int kAccPerBlock = 64;
  static constexpr int kAddends = 0;
  static constexpr int kGroups = 2;
  static constexpr bool kDirectStore = false;
  float* y;          // [P, 32] (written through the kernel's output tensor map)
  const float* bf;
  const float* bg;
  float tb[32], sb[32];   // biases in registers, pre-scaled for the exponentials (64 LDGs per tile and thread otherwise)
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < 32; ++j) { tb[j] = kTanhScale * __ldg(bf + j); sb[j] = kSigmScale * __ldg(bg + j); }
  }
  __device__ __forceinline__ void load_addends(const AddendRows&, bool) {}
  __device__ __forceinline__ void consume16(i64, int c0, int, const float (&v)[16], const RowSink& out) {
    const int ch0 = c0 >> 1;   // compile-time after unrolling (NCT > 0): tb / sb stay in registers
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {   // tanh(a) sigmoid(b) = (ea - 1) / den
      const GatePair g = gate_pair(v[2 * j], v[2 * j + 1], v[2 * j + 2], v[2 * j + 3], tb[ch0 + j], sb[ch0 + j], tb[ch0 + j + 1],
                                   sb[ch0 + j + 1]);
      o[j] = g.num[0] * (g.r * g.den[1]);
      o[j + 1] = g.num[1] * (g.r * g.den[0]);
    }
    out.put4(ch0, o[0], o[1], o[2], o[3]);
    out.put4(ch0 + 4, o[4], o[5], o[6], o[7]);
  }
  __device__ __forceinline__ void finish_rows(float*, int, int) {}
};

// Gate backward from recomputed pre-activations: dpre[m][2ch+{0,1}] (interleaved, 64 wide).
struct RowGateBwd {
  static constexpr bool kLazyAddends = false;
  static constexpr bool kHasFinish = false;
  static constexpr int kAccPerBlock = 32;
  static constexpr int kAddends = 1;   // dg rows
  static constexpr int kGroups = 2;
  static constexpr bool kDirectStore = false;
  float* dpre;       // [P, 64]
  const float* dg;   // [P, 32] (addend tile 0)
  const float* bf;
  const float* bg;
  float g[32];
  float tb[32], sb[32];   // pre-scaled biases in registers (see RowGate)
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < 32; ++j) { tb[j] = kTanhScale * __ldg(bf + j); sb[j] = kSigmScale * __ldg(bg + j); }
  }
  __device__ __forceinline__ void load_addends(const AddendRows& a, bool valid) {
    if (valid) a.load<32>(0, 0, g);
  }
  __device__ __forceinline__ void consume16(i64, int c0, int, const float (&v)[16], const RowSink& out) {
    const int ch0 = c0 >> 1;
    float o[16];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const GatePair gp = gate_pair(v[2 * j], v[2 * j + 1], v[2 * j + 2], v[2 * j + 3], tb[ch0 + j], sb[ch0 + j], tb[ch0 + j + 1],
                                    sb[ch0 + j + 1]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float rd = gp.r * gp.den[1 - h];   // 1 / den of this gate
        const float s = rd * gp.q1[h], f = gp.num[h] * (rd * gp.p1[h]);
        const float gg = g[ch0 + j + h];
        o[2 * (j + h)] = gg * s * (1.0f - f * f);
        o[2 * (j + h) + 1] = gg * f * s * (1.0f - s);
      }
    }
    out.put16(c0 & 31, o);
  }
  __device__ __forceinline__ void finish_rows(float*, int, int) {}
};

// gcn tail + residual + BatchNorm statistics (model.py:53-54, 234-236), N = 32.
struct RowMlp {
  static constexpr bool kLazyAddends = false;
  static constexpr bool kHasFinish = true;
  static constexpr int kAccPerBlock = 32;
  static constexpr int kAddends = 1;   // residual rows (tile 0; staged only when res != nullptr)
  static constexpr int kGroups = 2;
  static constexpr bool kDirectStore = false;
  float* y;            // [P, 32]
  const float* bias;
  DropoutSrc drop;
  const float* res;    // nullable
  Remap rrm;
  const float* rac;    // nullable fold of the residual
  double* stats;       // nullable
  float rrow[32];
  RowStats32 cs;
  __device__ __forceinline__ void init() {
    cs.reset();
    vec = ((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(rac)) & 15) == 0;
    if (drop.seed_dev) {   // the step's Philox key: read once per kernel instead of once per 8 elements
      drop.seed = (uint64_t)*drop.seed_dev;
      drop.seed_dev = nullptr;
    }
  }
  __device__ __forceinline__ void load_addends(const AddendRows& a, bool valid) {
    if (res && valid) a.load<32>(0, 0, rrow);
  }
  // 16 per-channel constants as four 128-bit loads (the address is warp-uniform); the flat parameter buffer and the
  // workspace keep every tensor 16-byte aligned, checked once in init()
  bool vec;
  __device__ __forceinline__ void load16(const float* p, float (&r)[16]) const {
    if (vec) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(p) + q);
        r[4 * q] = f.x; r[4 * q + 1] = f.y; r[4 * q + 2] = f.z; r[4 * q + 3] = f.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) r[c] = __ldg(p + c);
    }
  }
  __device__ __forceinline__ void consume16(i64 m, int c0, int, const float (&v)[16], const RowSink& out) {
    float o[16], bs[16];
    const i64 e = m * 32 + c0;
    load16(bias + c0, bs);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float kp[8];
      drop.keep8(e + 8 * q, kp);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[8 * q + i] = (v[8 * q + i] + bs[8 * q + i]) * kp[i];
    }
    // the null checks sit OUTSIDE the unrolled loops (inside, the compiler kept one branch per element: 64 per row)
    if (res) {
      if (rac) {
        float a[16], b[16];
        load16(rac + c0, a);
        load16(rac + 32 + c0, b);
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] += fmaf(rrow[c0 + c], a[c], b[c]);
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] += rrow[c0 + c];
      }
    }
    out.put16(c0, o);
    if (stats) {
#pragma unroll
      for (int c = 0; c < 16; ++c) { cs.s1[c0 + c] += o[c]; cs.s2[c0 + c] += o[c] * o[c]; }
    }
  }
  __device__ __forceinline__ void finish_rows(float* red, int etid, int barid) {
    if (stats) cs.reduce(red, etid, stats, stats + 32, 32, barid);
  }
};

// mlp input gradient: N = nseg*32 columns scattered to the per-segment tensors out[(q*M + m)*32 + nn].
struct RowSeg {
  static constexpr bool kLazyAddends = false;
  static constexpr bool kHasFinish = false;
  static constexpr int kAccPerBlock = 32;
  static constexpr int kAddends = 0;
  static constexpr int kGroups = 2;
  static constexpr bool kDirectStore = false;
  float* out;
  i64 M;
  __device__ __forceinline__ void init() {}
  __device__ __forceinline__ void load_addends(const AddendRows&, bool) {}
  __device__ __forceinline__ void consume16(i64, int c0, int, const float (&v)[16], const RowSink& o) { o.put16(c0 & 31, v); }
  __device__ __forceinline__ void finish_rows(float*, int, int) {}
};

// Gated-conv input gradient + residual path + BatchNorm-backward statistics of the layer below, N = 32.
struct RowTcnDgrad {
  static constexpr bool kLazyAddends = true;   // du / u_prev rows are read from the staged tiles 16 columns at a time:
  static constexpr bool kHasFinish = true;     // with them in registers (64) next to the 64 statistics registers only
  static constexpr int kAccPerBlock = 32;      // one epilogue group fitted and the kernel was epilogue-bound
  static constexpr int kAddends = 2;
  static constexpr int kGroups = 2;
  static constexpr bool kDirectStore = false;
  float* dx;           // [P_in, 32]
  const float* du;     // nullable [P_out, 32]
  int N, L_in, L_out;
  const float* uprev;  // nullable
  const float* mr;     // mean[32], rstd[32]
  double* bsum;
  AddendRows ar;
  RowStats32 cs;
  __device__ __forceinline__ void init() { cs.reset(); }
  // addend tile 0: du rows shifted by (L_in - L_out) time steps (rows before the first step arrive as zeros: TMA
  // out-of-bounds fill); tile 1: the layer-below pre-BN rows (only staged when uprev != nullptr)
  __device__ __forceinline__ void load_addends(const AddendRows& a, bool) { ar = a; }
  __device__ __forceinline__ void consume16(i64, int c0, int, const float (&v)[16], const RowSink& out) {
    float o[16];
    if (du) {
      float d[16];
      ar.load<16>(0, c0, d);
#pragma unroll
      for (int c = 0; c < 16; ++c) o[c] = v[c] + d[c];
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) o[c] = v[c];
    }
    out.put16(c0, o);
    if (uprev) {
      float ur[16];
      ar.load<16>(1, c0, ur);
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // mean / rstd of 4 channels per 128-bit load (workspace rows are 16-byte aligned)
        const float4 mu = __ldg(reinterpret_cast<const float4*>(mr + c0) + q), rs = __ldg(reinterpret_cast<const float4*>(mr + 32 + c0) + q);
        const float mu_[4] = {mu.x, mu.y, mu.z, mu.w}, rs_[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 4 * q + j;
          const float xh = (ur[c] - mu_[j]) * rs_[j];
          cs.s1[c0 + c] += o[c];
          cs.s2[c0 + c] += o[c] * xh;
        }
      }
    }
  }
  __device__ __forceinline__ void finish_rows(float* red, int etid, int barid) {
    if (uprev) cs.reduce(red, etid, bsum, bsum + 32, 32, barid);
  }
};
// Dense head layers (model.py:216-222, 238-240 and their input gradients): out = act(acc + bias[col]) [* (gate > 0)].
struct RowDense {
  static constexpr bool kLazyAddends = false;
  static constexpr bool kHasFinish = false;
  static constexpr int kAccPerBlock = 32;
  static constexpr int kAddends = 0;
  static constexpr int kGroups = 2;
  static constexpr bool kDirectStore = false;
  const float* bias;   // nullable, indexed by the absolute output column
  const float* gate;   // nullable [M][ldg]: ReLU mask of the forward activation (backward)
  int ldg, relu;
  bool vec;   // bias 16-byte aligned: 128-bit loads of the (warp-uniform) bias values
  __device__ __forceinline__ void init() { vec = (reinterpret_cast<uintptr_t>(bias) & 15) == 0; }
  __device__ __forceinline__ void load_addends(const AddendRows&, bool) {}
  __device__ __forceinline__ void consume16(i64 m, int c0, int n0, const float (&v)[16], const RowSink& out) {
    float o[16];
    const int ca = n0 + c0;   // absolute output column (multiple of 16)
    // the null / mode checks sit outside the element loops (see RowMlp)
    if (bias) {
      if (vec) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + ca) + q);
          o[4 * q] = v[4 * q] + b4.x; o[4 * q + 1] = v[4 * q + 1] + b4.y; o[4 * q + 2] = v[4 * q + 2] + b4.z; o[4 * q + 3] = v[4 * q + 3] + b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = v[j] + __ldg(bias + ca + j);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = v[j] + 0.0f;
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.0f);
    }
    if (gate) {
      const float4* gp = reinterpret_cast<const float4*>(gate + m * ldg + ca);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 g = __ldg(gp + q);
        if (!(g.x > 0.0f)) o[4 * q] = 0.0f;
        if (!(g.y > 0.0f)) o[4 * q + 1] = 0.0f;
        if (!(g.z > 0.0f)) o[4 * q + 2] = 0.0f;
        if (!(g.w > 0.0f)) o[4 * q + 3] = 0.0f;
      }
    }
    out.put16(c0 & 31, o);
  }
  __device__ __forceinline__ void finish_rows(float*, int, int) {}
};

// Last head layer (model.py:240): the network output in the reference's NCHW layout [B, O, N, T]; rows are BLNC
// positions m = (b*T + t)*N + n.  O <= 16 columns, written straight from registers (consecutive threads = consecutive
// nodes = consecutive addresses for T == 1).
struct RowNCHW {
  static constexpr bool kLazyAddends = false;
  static constexpr bool kHasFinish = false;
  static constexpr int kAccPerBlock = 16;
  static constexpr int kAddends = 0;
  static constexpr int kGroups = 2;
  static constexpr bool kDirectStore = true;
  float* y;
  const float* bias;
  int O, N, T;
  __device__ __forceinline__ void init() {}
  __device__ __forceinline__ void load_addends(const AddendRows&, bool) {}
  __device__ __forceinline__ void consume16(i64 m, int c0, int, const float (&v)[16], const RowSink&) {
    const unsigned nt = (unsigned)N * (unsigned)T;
    const unsigned b = (unsigned)m / nt, r = (unsigned)m - b * nt;
    const unsigned t = r / (unsigned)N, n = r - t * (unsigned)N;
    float* base = y + ((size_t)b * O * N + n) * T + t;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < O) base[(size_t)(c0 + j) * N * T] = v[j] + __ldg(bias + c0 + j);
  }
  __device__ __forceinline__ void finish_rows(float*, int, int) {}
};
#else
struct RowDense { const float* bias; const float* gate; int ldg, relu; };
struct RowNCHW { float* y; const float* bias; int O, N, T; };
struct RowGate { float* y; const float* bf; const float* bg; };
struct RowGateBwd { float* dpre; const float* dg; const float* bf; const float* bg; };
struct RowMlp { float* y; const float* bias; DropoutSrc drop; const float* res; Remap rrm; const float* rac; double* stats; };
struct RowSeg { float* out; i64 M; };
struct RowTcnDgrad { float* dx; const float* du; int N, L_in, L_out; const float* uprev; const float* mr; double* bsum; };
#endif

}  // namespace gwn
