// tcgen05 / TMEM / TMA node contraction -- see nconv_tc.cuh for the GEMM mapping.
#pragma once
#include "nconv_tc.cuh"
#include "tc_common.cuh"

#if !GWN_EMU

namespace gwn {
namespace tc {

constexpr int BLOCK_M = 128;   // 4 slabs x 32 channels
constexpr int SLABS = 4;
constexpr int CH = 32;
// Nodes per pipeline stage, BK = 32 (one 128-byte swizzle row of K for the support tile) or 16 (64-byte rows, SWIZZLE_64B):
// in the 3xTF32 mode a stage holds four planes [X | X_lo | S | S_lo] and only TWO 32-node stages fit beside each other
// at V ~ 200 -- the MMA issuer then waited for data 40-50 % of the time (ncu r02h: producer blocked on `empty`, splitter
// on `full`, tensor pipe 53 % active).  Half-size stages put five in the same shared memory: same bytes in flight, but
// the loads run four stages ahead of the MMAs instead of one.
constexpr int UMMA_K = 8;      // kind::tf32: 32 bytes of K per instruction
constexpr int NUM_THREADS = 384;                          // warps 0-3: control roles, warps 4-11: two epilogue sets (nconv_tc2.cuh)
constexpr int ACC_COLS = 256;                             // TMEM columns per accumulator buffer (2 buffers)

struct Maps {
  CUtensorMap x[TC_MAXSUP];
  CUtensorMap s[TC_MAXSUP];
  CUtensorMap slo[TC_MAXSUP];
};

struct Params {
  float* Y[TC_MAXSUP];
  const float* add[TC_MAXSUP];
  const float* add2;
  int nsup, kcat, V, L, T_out, nslabs;
  int n_tile, n_wt, n_jt, nkb, stages, total_tiles;
  int per_sample, tps;   // per-sample supports: tiles per sample = ceil(L / 4); X is addressed (c, v, l, b), S (k, m, b)
  int mode;     // debug: 0 normal, 1 = TMEM st/ld self-test (no MMA), 2 = A operand := support tile (K-major)
  float* dbg;   // debug: dump of the first pipeline stage (X tile then support tile) by block 0
};

// ------------------------------------------------------------------------------------------ kernel
// X3 = 3xTF32 mode (fp32-grade): D = X.S + X.S_lo + X_lo.S with S_lo precomputed in global memory and X_lo produced
// in shared memory by warps 2 and 3 from the tile TMA just landed (see tcpos.cuh; kind::tf32 truncates, so the fp32
// tiles themselves are the high parts).  Stage layout: [X | X_lo | S | S_lo].
template <bool X3, int BLOCK_K>
__global__ void __launch_bounds__(NUM_THREADS, 1) nconv_tc_kernel(const __grid_constant__ Maps maps, const Params p) {
  constexpr int X_STAGE_BYTES = SLABS * BLOCK_K * CH * 4;   // 16 KB (BK = 32) / 8 KB (BK = 16)
  constexpr int S_ROW_BYTES = BLOCK_K * 4;                  // one support row of a stage: 128 B (SWIZZLE_128B) / 64 B (SWIZZLE_64B)
  constexpr uint32_t S_LAYOUT = BLOCK_K == 32 ? 2u : 4u;    // descriptor layout type of the K-major support tile
  constexpr uint32_t S_SBO = 8 * S_ROW_BYTES;               // 8-row groups
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;           // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  constexpr int NPL = X3 ? 2 : 1;
  constexpr int XB = NPL * X_STAGE_BYTES;                 // X planes of one stage
  const int s_tile = p.n_tile * S_ROW_BYTES;              // one support plane of one stage
  const int stage_bytes = XB + NPL * s_tile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  const uint32_t bar0 = base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (p.stages + s); };
  auto split_bar = [&](int s) { return bar0 + 8u * (2 * p.stages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * p.stages + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * p.stages + 4);

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsup; ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x[s]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.s[s]) : "memory");
      if (X3) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.slo[s]) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(split_bar(s), 64);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * ACC_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  GWN_PDL_ENTRY();   // prologue above (barriers, TMEM, tensor-map prefetch) overlapped the previous kernel's tail
  const int per_out = p.n_jt * p.n_wt;

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp loops, one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.total_tiles && ok; tile += gridDim.x) {
      const int o = tile / per_out, rem = tile - o * per_out;
      const int wt = rem / p.n_jt, jt = rem - wt * p.n_jt;
      const int s0 = p.kcat ? 0 : o, s1 = p.kcat ? p.nsup : o + 1;
      for (int s = s0; s < s1 && ok; ++s) {
        for (int kb = 0; kb < p.nkb; ++kb) {
          if (!mbar_wait_warp(empty_bar(stage), phase ^ 1u, 1)) { ok = false; break; }
          const uint32_t dst = base + stage * stage_bytes;
          if (elect_one()) {
          mbar_expect_tx(full_bar(stage), (uint32_t)(X_STAGE_BYTES + NPL * s_tile));
          if (!p.per_sample) {
            tma_load_3d(dst, &maps.x[s], full_bar(stage), 0, kb * BLOCK_K, jt * SLABS);
            tma_load_2d(dst + XB, &maps.s[s], full_bar(stage), kb * BLOCK_K, wt * p.n_tile);
            if (X3) tma_load_2d(dst + XB + s_tile, &maps.slo[s], full_bar(stage), kb * BLOCK_K, wt * p.n_tile);
          } else {   // the tile's 4 time steps of ONE sample (steps past L arrive as zeros) against that sample's support
            const int sb = jt / p.tps, lt = jt - sb * p.tps;
            tma_load_4d(dst, &maps.x[s], full_bar(stage), 0, kb * BLOCK_K, lt * SLABS, sb);
            tma_load_3d(dst + XB, &maps.s[s], full_bar(stage), kb * BLOCK_K, wt * p.n_tile, sb);
            if (X3) tma_load_3d(dst + XB + s_tile, &maps.slo[s], full_bar(stage), kb * BLOCK_K, wt * p.n_tile, sb);
          }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (whole warp loops, one elected lane issues)
    // instruction descriptor: D=f32, A=B=tf32, A MN-major, B K-major, N = n_tile, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(p.n_tile >> 3) << 17) |
                           ((uint32_t)(BLOCK_M >> 4) << 24);
    int stage = 0, acc = 0;
    uint32_t phase = 0, accphase = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.total_tiles && ok; tile += gridDim.x) {
      if (!mbar_wait_warp(tempty_bar(acc), accphase ^ 1u, 2)) break;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
      const int nk_total = (p.kcat ? p.nsup : 1) * p.nkb;
      for (int it = 0; it < nk_total; ++it) {
        if (!mbar_wait_warp(full_bar(stage), phase, 3)) { ok = false; break; }
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && tile == 0 && it == 0 && lane == 0) {
          const float* sm = reinterpret_cast<const float*>(smem + (size_t)stage * stage_bytes);
          for (int i = 0; i < stage_bytes / 4; ++i) p.dbg[i] = sm[i];
        }
        const uint32_t xs = base + stage * stage_bytes;
        const uint32_t bs = xs + XB;
        if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
          // A (X^T, MN-major, 32-byte-atom swizzle): atoms of 4 k-rows x 128 B (SBO = 512 B between them);
          // the next 32 rows of M (next slab) lie BLOCK_K*128 B further (LBO); this k-step starts 8 rows in.
          uint64_t adesc = make_desc(xs + kk * (UMMA_K * 128), BLOCK_K * 128, 4 * 128, 1);
          uint32_t idesc_k = idesc;
          if (p.mode == 2) {   // debug: A := first 128 rows of the support tile, K-major
            adesc = make_desc(bs + kk * (UMMA_K * 4), 16, S_SBO, S_LAYOUT);
            idesc_k = idesc & ~(1u << 15);
          }
          // B (support, K-major): rows of BLOCK_K floats, 8-row groups S_SBO apart; this k-step starts 32 B further in
          const uint64_t bdesc = make_desc(bs + kk * (UMMA_K * 4), 16, S_SBO, S_LAYOUT);
          if (p.mode != 1) tc_mma_tf32(d_tmem, adesc, bdesc, idesc_k, (it > 0 || kk > 0) ? 1u : 0u);
          if (X3) tc_mma_tf32(d_tmem, adesc, make_desc(bs + s_tile + kk * (UMMA_K * 4), 16, S_SBO, S_LAYOUT), idesc_k, 1u);
        }
        }
        __syncwarp();
        if (X3) {   // the X_lo term last: the split of this stage ran while the eight MMAs above were issued
          if (!mbar_wait_warp(split_bar(stage), phase, 5)) { ok = false; break; }
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
            tc_mma_tf32(d_tmem, make_desc(xs + X_STAGE_BYTES + kk * (UMMA_K * 128), BLOCK_K * 128, 4 * 128, 1),
                        make_desc(bs + kk * (UMMA_K * 4), 16, S_SBO, S_LAYOUT), idesc, 1u);
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(empty_bar(stage));     // frees the smem stage once these MMAs have read it
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (!ok) break;
      if (elect_one()) tc_commit(tfull_bar(acc));         // accumulator complete -> epilogue
      __syncwarp();
      acc ^= 1;
      if (acc == 0) accphase ^= 1u;
    }
  } else if (X3 && (warp == 2 || warp == 3)) {
    // ===================================================== splitter: X_lo = X - tf32_trunc(X) (element-wise: layout-preserving)
    const int t64 = threadIdx.x - 64;
    int stage = 0;
    uint32_t phase = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.total_tiles && ok; tile += gridDim.x) {
      const int nk_total = (p.kcat ? p.nsup : 1) * p.nkb;
      for (int it = 0; it < nk_total; ++it) {
        if (!mbar_wait(full_bar(stage), phase, 6)) { ok = false; break; }
        const float4* src = reinterpret_cast<const float4*>(smem + (size_t)stage * stage_bytes);
        float4* dst = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes + X_STAGE_BYTES);
        {   // 16 (8) float4 per thread: all loads in flight before the first use (with 4 at a time the two splitter
            // warps were busy ~100 % of the time, stalled on LDS results: ncu source page of the reduction kernel)
          constexpr int NV = X_STAGE_BYTES / 16 / 64;
          static_assert(NV * 64 * 16 == X_STAGE_BYTES, "splitter: whole float4s per thread");
          float4 v[NV];
#pragma unroll
          for (int u = 0; u < NV; ++u) v[u] = src[t64 + 64 * u];
#pragma unroll
          for (int u = 0; u < NV; ++u)
            dst[t64 + 64 * u] = make_float4(tf32_lo(v[u].x), tf32_lo(v[u].y), tf32_lo(v[u].z), tf32_lo(v[u].w));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(split_bar(stage));
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: TMEM -> registers -> global
    const int ew = (warp - 4) & 3, eh = (warp - 4) >> 2;   // ew == warp % 4: TMEM lanes 32*ew .. 32*ew+31 = slab ew of the tile, lane = channel; eh: column-group half
    int acc = 0;
    uint32_t accphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int o = tile / per_out, rem = tile - o * per_out;
      const int wt = rem / p.n_jt, jt = rem - wt * p.n_jt;
      if (!mbar_wait(tfull_bar(acc), accphase, 4)) break;
      tc_fence_after();
      int slab = jt * SLABS + ew;
      bool slab_ok = slab < p.nslabs;
      if (p.per_sample) {
        const int sb = jt / p.tps, l = (jt - sb * p.tps) * SLABS + ew;
        slab = sb * p.L + l;
        slab_ok = l < p.L;
        if (!slab_ok) slab = 0;
      }
      float* y = p.Y[o] + (size_t)slab * p.V * CH + lane;
      const float* ad = p.add[o] ? p.add[o] + (size_t)slab * p.V * CH + lane : nullptr;
      const float* ad2 = nullptr;
      if (p.add2 && slab_ok) {
        const int b = slab / p.L, l = slab - b * p.L;
        if (l >= p.L - p.T_out) ad2 = p.add2 + ((size_t)(b * p.T_out + (l - (p.L - p.T_out))) * p.V) * CH + lane;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * ew) << 16) + (uint32_t)(acc * ACC_COLS);
      const int w_base = wt * p.n_tile;
      if (p.mode == 1) {   // debug: write lane*1000 + column into TMEM, read it back below
        for (int c0 = 0; c0 < p.n_tile; ++c0) {
          const uint32_t v = __float_as_uint((float)((32 * ew + lane) * 1000 + c0));
          asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c0), "r"(v) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      const bool has_ad = slab_ok && ad != nullptr, has_ad2 = slab_ok && ad2 != nullptr;   // warp-uniform
      // Columns in groups of 64: all addend loads of a group are issued back to back (unconditional, clamped
      // addresses: per-element predication made each load its own reconvergence block and serialised their
      // latencies), then the group's four accumulator reads, then the stores.  With 16-column groups the backward
      // launches paid one exposed L2/HBM round trip per 16 columns -- 13 per tile -- and ran 1.4-2x longer than the
      // addend-free forward launches of the same MMA work (ncu launch lists r01c/r01f).
      constexpr int GC = 64;
      const int ngrp = (p.n_tile + GC - 1) / GC, gsplit = ((ngrp + 1) >> 1) * GC;
      for (int g0 = eh ? gsplit : 0; g0 < (eh ? p.n_tile : gsplit); g0 += GC) {
        float av[GC];
#pragma unroll
        for (int j = 0; j < GC; ++j) av[j] = 0.0f;
        if (has_ad) {
#pragma unroll
          for (int j = 0; j < GC; ++j) av[j] = __ldg(ad + (size_t)min(w_base + g0 + j, p.V - 1) * CH);
        }
        if (has_ad2) {
#pragma unroll
          for (int j = 0; j < GC; ++j) av[j] += __ldg(ad2 + (size_t)min(w_base + g0 + j, p.V - 1) * CH);
        }
#pragma unroll
        for (int cc = 0; cc < GC; cc += 16) {
          const int c0 = g0 + cc;
          if (c0 < p.n_tile) {   // warp-uniform
            uint32_t r[16];
            tc_ld16(taddr + c0, r);
            tc_wait_ld();
            if (slab_ok) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int w = w_base + c0 + j;
                if (w < p.V) y[(size_t)w * CH] = __uint_as_float(r[j]) + av[cc + j];
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) accphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * ACC_COLS));
  }
}

}  // namespace tc

static float* g_dbg_ptr = nullptr;
static int g_dbg_mode = 0;
void tc_set_debug_buffer(float* p) { g_dbg_ptr = p; }
void tc_set_debug_mode(int m) { g_dbg_mode = m; }

int tc_error_flag(int reset) {
  int v = 0;
  cudaMemcpyFromSymbol(&v, tc::g_tc_err, sizeof(int));
  if (reset) {
    int z = 0;
    cudaMemcpyToSymbol(tc::g_tc_err, &z, sizeof(int));
  }
  return v;
}

static int tc_num_sms() {
  static const int n = [] {
    int dev = 0, v = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

template <bool X3, int BLOCK_K>
static int node_gemm_tc_impl(const NodeTcArgs& a, cudaStream_t stream) {
  using namespace tc;
  constexpr int X_STAGE_BYTES = SLABS * BLOCK_K * CH * 4;
  if (a.nsup < 1 || a.nsup > TC_MAXSUP) {
    set_error("node_gemm_tc: %d supports (max %d)", a.nsup, TC_MAXSUP);
    return GWN_ERR_UNSUPPORTED;
  }
  if (a.ld % 4 != 0 || a.ld < a.V) {
    set_error("node_gemm_tc: support leading dimension %d must be a multiple of 4 and >= V", a.ld);
    return GWN_ERR_UNSUPPORTED;
  }
  const long long nslabs = (long long)a.B * a.L;
  if (nslabs <= 0 || a.V <= 0) return 0;
  Maps maps;
  Params p;
  memset(&p, 0, sizeof(p));
  p.nsup = a.nsup; p.kcat = a.kcat; p.V = a.V; p.L = a.L; p.T_out = a.T_out; p.nslabs = (int)nslabs;
  p.per_sample = a.per_sample ? 1 : 0;
  p.tps = (a.L + SLABS - 1) / SLABS;
  p.n_jt = a.per_sample ? a.B * p.tps : (int)((nslabs + SLABS - 1) / SLABS);
  p.nkb = (a.V + BLOCK_K - 1) / BLOCK_K;
  const int nout = a.kcat ? 1 : a.nsup;
  p.n_tile = a.V > 256 ? 256 : round_up(a.V, 16);
  if (a.V <= 256) {
    // Small graphs: the 4-slab row tiles alone are coarse against 148 SMs (METR-LA, L = 12, K-concatenated sum:
    // 192 tiles = 1.3 waves, i.e. two rounds at 65 % occupancy; L = 4: 64 tiles on 148 SMs).  Splitting the output
    // columns into n_wt tiles multiplies the tile count, shrinks the stage (more pipeline stages fit) and costs only
    // the re-read of the X k-block from L2.  Pick the split with the smallest modelled makespan:
    //   rounds x (k-steps x (columns + per-k-step fixed cost) + per-tile fill/drain).
    static const int forced = [] {
      const char* e = getenv("GWNET_B200_NCONV_NWT");
      return e ? atoi(e) : 0;
    }();
    const int nk_total = (a.kcat ? a.nsup : 1) * p.nkb;
    long long best = -1;
    for (int nw = 1; nw <= 4; ++nw) {
      if (forced > 0 && nw != forced) continue;
      const int nt = round_up((a.V + nw - 1) / nw, 16);
      const int nwe = (a.V + nt - 1) / nt;
      const long long tiles_ = (long long)p.n_jt * nwe * nout;
      const long long rounds = (tiles_ + tc_num_sms() - 1) / tc_num_sms();
      const long long cost = rounds * ((long long)nk_total * ((nt < 64 ? 64 : nt) + 16) + 300);
      if (best < 0 || cost < best) { best = cost; p.n_tile = nt; }
    }
  }
  p.n_wt = (a.V + p.n_tile - 1) / p.n_tile;
  const int stage_bytes = (X3 ? 2 : 1) * (X_STAGE_BYTES + p.n_tile * BLOCK_K * 4);
  p.stages = (SMEM_LIMIT - 2048) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) {
    set_error("node_gemm_tc: tile does not fit shared memory");
    return GWN_ERR_UNSUPPORTED;
  }
  const long long tiles = (long long)p.n_jt * p.n_wt * nout;
  if (tiles > 2147483647LL) {
    set_error("node_gemm_tc: too many tiles");
    return GWN_ERR_UNSUPPORTED;
  }
  p.total_tiles = (int)tiles;
  for (int s = 0; s < a.nsup; ++s) {
    if ((reinterpret_cast<uintptr_t>(a.X[s]) & 15) || (reinterpret_cast<uintptr_t>(a.S[s]) & 15)) {
      set_error("node_gemm_tc: operands must be 16-byte aligned");
      return GWN_ERR_UNSUPPORTED;
    }
    cuuint64_t sd[3] = {(cuuint64_t)a.V, (cuuint64_t)a.V, (cuuint64_t)a.B};
    cuuint64_t ss[2] = {(cuuint64_t)a.ld * 4, (cuuint64_t)a.s_batch_stride * 4};
    cuuint32_t sb[3] = {BLOCK_K, (cuuint32_t)p.n_tile, 1};
    const int srank = a.per_sample ? 3 : 2;
    if (a.per_sample) {
      if (a.s_batch_stride % 4 != 0) {
        set_error("node_gemm_tc: per-sample support stride must be a multiple of 4 floats");
        return GWN_ERR_UNSUPPORTED;
      }
      cuuint64_t xd[4] = {(cuuint64_t)CH, (cuuint64_t)a.V, (cuuint64_t)a.L, (cuuint64_t)a.B};
      cuuint64_t xs[3] = {(cuuint64_t)CH * 4, (cuuint64_t)a.V * CH * 4, (cuuint64_t)a.L * a.V * CH * 4};
      cuuint32_t xb[4] = {CH, BLOCK_K, SLABS, 1};
      GWN_TRY(encode(&maps.x[s], a.X[s], 4, xd, xs, xb, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    } else {
      cuuint64_t xd[3] = {(cuuint64_t)CH, (cuuint64_t)a.V, (cuuint64_t)nslabs};
      cuuint64_t xs[2] = {(cuuint64_t)CH * 4, (cuuint64_t)a.V * CH * 4};
      cuuint32_t xb[3] = {CH, BLOCK_K, SLABS};
      GWN_TRY(encode(&maps.x[s], a.X[s], 3, xd, xs, xb, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    }
    const CUtensorMapSwizzle s_swz = BLOCK_K == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    GWN_TRY(encode(&maps.s[s], a.S[s], srank, sd, ss, sb, s_swz));
    if (X3) {
      if (!a.Slo[s] || (reinterpret_cast<uintptr_t>(a.Slo[s]) & 15)) {
        set_error("node_gemm_tc: 3xTF32 mode needs 16-byte aligned support remainders");
        return GWN_ERR_UNSUPPORTED;
      }
      GWN_TRY(encode(&maps.slo[s], a.Slo[s], srank, sd, ss, sb, s_swz));
    } else {
      maps.slo[s] = maps.s[s];
    }
  }
  for (int s = a.nsup; s < TC_MAXSUP; ++s) { maps.x[s] = maps.x[0]; maps.s[s] = maps.s[0]; maps.slo[s] = maps.slo[0]; }
  for (int o = 0; o < nout; ++o) {
    p.Y[o] = a.Y[o];
    p.add[o] = a.add[o];
  }
  p.add2 = a.add2;
  p.dbg = g_dbg_ptr;
  p.mode = g_dbg_mode;

  const int smem_bytes = p.stages * stage_bytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(nconv_tc_kernel<X3, BLOCK_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  });
  const int num_sms = tc_num_sms();
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(attr_err));
    return GWN_ERR_CUDA;
  }
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  GWN_CUDA(launch_kernel(nconv_tc_kernel<X3, BLOCK_K>, dim3(grid), dim3(NUM_THREADS), smem_bytes, stream, maps, p));
  count_launch();
  return 0;
}

}  // namespace gwn

#include "nconv_tc2.cuh"

namespace gwn {

int node_gemm_tc(const NodeTcArgs& a, cudaStream_t stream) {
  // large graphs: CTA pairs (cta_group::2, nconv_tc2.cuh); GWNET_B200_NCONV_2CTA=0 keeps the one-CTA kernel (A/B runs)
  static const bool two_cta = [] {
    const char* e = getenv("GWNET_B200_NCONV_2CTA");
    return !(e && e[0] == '0');
  }();
  // Which graphs take the CTA-pair kernel (GWNET_B200_NCONV_2CTA_MINV overrides both thresholds for A/B runs):
  //   V > 256: always -- the one-CTA kernel is bound by its L2 -> SM operand fill there (544 vs 774 TFLOP/s at N = 2048);
  //   V <= 256, 3xTF32: pairs with half-size stages (seven 29 KB stages per CTA): METR-LA step 2.71 -> 2.65 ms (r02l);
  //   V <= 256, tf32: the one-CTA kernel (both are HBM-bound there: 5.7 TB/s at V = 207, and it is 1 % faster).
  static const int min_v_env = [] {
    const char* e = getenv("GWNET_B200_NCONV_2CTA_MINV");
    return e ? atoi(e) : -1;
  }();
  const int min_v = min_v_env >= 0 ? min_v_env : (a.Slo[0] ? 16 : 257);
  if (two_cta && !a.per_sample && a.V >= min_v && a.V >= 16 && a.nsup >= 1 && a.nsup <= TC_MAXSUP && a.ld % 4 == 0 && a.ld >= a.V &&
      (long long)a.B * a.L > 0) {
    static const int bk2 = [] {   // GWNET_B200_NCONV2_BK = 32: 32-node stages also on small graphs (A/B runs)
      const char* e = getenv("GWNET_B200_NCONV2_BK");
      return (e && atoi(e) == 32) ? 32 : 16;
    }();
    int st;
    if (bk2 == 16 && a.V <= 256) st = a.Slo[0] ? node_gemm_tc2_impl<true, 16>(a, stream) : node_gemm_tc2_impl<false, 16>(a, stream);
    else st = a.Slo[0] ? node_gemm_tc2_impl<true, 32>(a, stream) : node_gemm_tc2_impl<false, 32>(a, stream);
    if (st >= 0) return st;
  }
  // 3xTF32 on small graphs: half-size stages (see BLOCK_K above); GWNET_B200_NCONV_BK=32 restores the 32-node stages
  static const int bk_x3 = [] {
    const char* e = getenv("GWNET_B200_NCONV_BK");
    return (e && atoi(e) == 32) ? 32 : 16;
  }();
  if (a.Slo[0]) return (bk_x3 == 16 && a.V <= 256) ? node_gemm_tc_impl<true, 16>(a, stream) : node_gemm_tc_impl<true, 32>(a, stream);
  return node_gemm_tc_impl<false, 32>(a, stream);
}

}  // namespace gwn

#include "gcn_hops_fused.cuh"

#else   // GWN_EMU: the tensor-core tier exists only on the GPU
namespace gwn {
void tc_set_debug_buffer(float*) {}
void tc_set_debug_mode(int) {}
int tc_error_flag(int) { return 0; }
int node_gemm_tc(const NodeTcArgs&, cudaStream_t) {
  set_error("the tcgen05 tier is not part of the host emulation");
  return GWN_ERR_UNSUPPORTED;
}
int gcn_hops_fused_tc(const float*, const float* const*, int, int, float* const*, float* const*, int, int, int, cudaStream_t) { return -1; }
}  // namespace gwn
#endif
