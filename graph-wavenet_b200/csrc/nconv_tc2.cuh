// Node contraction on CTA pairs: tcgen05.mma.cta_group::2 (model.py:13 and its autograd).  Written for large graphs
// (V > 256, column tiles of 256 nodes); small graphs (V <= 256) use it with 1-4 narrower column tiles.
//
//   D[j, m] = sum_s sum_k X_s[k, j] * S_s[m, k]          j = (slab, channel) row, m = output node
//
// Same GEMM mapping as nconv_tc_impl.cuh, but one tile is 256 rows (8 slabs x 32 channels) x 256 output nodes computed by
// a cluster of two CTAs with ONE instruction stream (the leader CTA issues M = 256 MMAs for the pair):
//   * each CTA stages only ITS 4 slabs of X and HALF (128 rows) of the 256-row support tile per k-block -- the tensor
//     core reads the other half from the peer CTA's shared memory.  Per k-step a CTA moves 1024 B from L2 and reads
//     8 KB of operands from shared memory instead of 1536 B / 12 KB for a 128 x 256 tile on one CTA; the one-CTA kernel
//     was paced by exactly that traffic (243 cycles per 128x256x8 MMA against a floor of 128: the L2 -> SM fill of
//     48 KB per k-block ran at ~50 B/cycle/SM, the chip-wide L2 cap is ~43).
//   * tiles are ordered column-tile-fastest, so the clusters running at any moment share a few 8-slab X groups (each X
//     k-block is fetched from HBM once and served to the other column tiles from L2) while the support stays L2-resident.
// Barrier protocol (s = pipeline stage, a = accumulator buffer; L = barrier lives in the leader CTA, E = in each CTA):
//   full[s]   L  1 arrival + tx bytes: the leader's producer arms it for BOTH CTAs' bytes, every TMA of either CTA
//                (cp.async.bulk.tensor ... .cta_group::2) completes on it
//   xfull[s]  E  3xTF32 mode only: the CTA's own X tile landed (its splitter warps wait on it)
//   split[s]  L  3xTF32 mode only: 2 arrivals = one per CTA, after its 64 splitter threads have written and fenced X_lo
//   empty[s]  E  tcgen05.commit.cta_group::2 ... multicast -> both CTAs: the MMAs reading stage s have completed
//   tfull[a]  E  commit multicast: accumulator a is complete in both CTAs' tensor memory
//   tempty[a] L  8 arrivals = the 4 epilogue warps of each CTA have drained accumulator a
#pragma once
#include "nconv_tc.cuh"
#include "tc_common.cuh"

#if !GWN_EMU

namespace gwn {
namespace tc2 {

using tc::smem_u32; using tc::mbar_init; using tc::mbar_expect_tx; using tc::mbar_wait; using tc::mbar_wait_warp;
using tc::elect_one; using tc::uniform_warp_id; using tc::tma_load_3d; using tc::tc_fence_before; using tc::tc_fence_after;
using tc::tc_ld16; using tc::tc_wait_ld; using tc::make_desc;

constexpr int SLABS = 4, CH = 32, UMMA_K = 8;
constexpr int ACC_COLS = 256;                       // TMEM columns per accumulator buffer (2 buffers)
constexpr int NUM_THREADS = 384;                     // producer, MMA issuer, 2 splitter warps, 2 x 4 epilogue warps
constexpr int MAXSTAGES = 8;

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
  int n_wt, n_jt, nkb, stages, total_tiles;
  int n_tile;    // output nodes per tile (MMA N, multiple of 16, <= 256); each CTA stages n_tile / 2 support rows
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {   // shared::cluster address of rank's copy
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// No ordering of the caller's earlier global stores is needed (the consumer only re-uses tensor memory that was read
// with tcgen05.ld + wait::ld before): a release at cluster scope would wait for those stores to drain.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");   // default: release at CTA scope
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads of a CTA pair: the bytes land in THIS CTA's shared memory, the transaction completes on the barrier at
// `bar_cluster` (a shared::cluster address: the leader's barrier)
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc2_commit(uint32_t bar_local) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_local),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc2_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// X3 = 3xTF32 mode: D = X.S + X.S_lo + X_lo.S; stage layout per CTA [X | X_lo | S half | S_lo half].
// BLOCK_K = nodes per pipeline stage: 32 (support rows of 128 B, SWIZZLE_128B) or 16 (64-byte rows, SWIZZLE_64B: twice
// the stages in the same shared memory, see nconv_tc_impl.cuh).
template <bool X3, int BLOCK_K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) nconv_tc2_kernel(const __grid_constant__ Maps maps, const Params p) {
  constexpr int X_BYTES = SLABS * BLOCK_K * CH * 4;   // this CTA's 128 rows of the A operand, one k-block
  constexpr uint32_t S_LAYOUT = BLOCK_K == 32 ? 2u : 4u;
  constexpr uint32_t S_SBO = 8 * BLOCK_K * 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;   // same offset in both CTAs: the dynamic shared window starts at the same address
  uint8_t* smem = smem_raw + (base - raw);
  constexpr int NPL = X3 ? 2 : 1;
  constexpr int XB = NPL * X_BYTES;
  const int N_TILE = p.n_tile, S_ROWS = N_TILE >> 1;   // this CTA's half of the B tile
  const int S_TX = S_ROWS * BLOCK_K * 4;                 // bytes one support TMA delivers
  const int S_BYTES = (S_TX + 1023) & ~1023;             // its plane in the stage: the next X tile must stay 1024-byte aligned
  const int STAGE = NPL * (X_BYTES + S_BYTES);
  const int stages = p.stages;
  const uint32_t bar0 = base + stages * STAGE;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAXSTAGES + s); };
  auto xfull_bar = [&](int s) { return bar0 + 8u * (2 * MAXSTAGES + s); };
  auto split_bar = [&](int s) { return bar0 + 8u * (3 * MAXSTAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (4 * MAXSTAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (4 * MAXSTAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)stages * STAGE + 8 * (4 * MAXSTAGES + 4));

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();   // 0 = leader of the pair
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nsup; ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x[s]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.s[s]) : "memory");
      if (X3) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.slo[s]) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(xfull_bar(s), 1);
      mbar_init(split_bar(s), 2);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  cluster_sync_all();        // both CTAs' barriers are initialised and both halves of the tensor memory are allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  GWN_PDL_ENTRY();

  const int cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int per_out = p.n_jt * p.n_wt;

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    bool ok = true;
    for (int tile = cl; tile < p.total_tiles && ok; tile += ncl) {
      const int o = tile / per_out, rem = tile - o * per_out;
      const int jt = rem / p.n_wt, wt = rem - jt * p.n_wt;   // column tile fastest
      const int s0 = p.kcat ? 0 : o, s1 = p.kcat ? p.nsup : o + 1;
      for (int s = s0; s < s1 && ok; ++s) {
        for (int kb = 0; kb < p.nkb; ++kb) {
          if (!mbar_wait_warp(empty_bar(stage), phase ^ 1u, 1)) { ok = false; break; }
          const uint32_t dst = base + stage * STAGE;
          if (elect_one()) {
            const uint32_t lfull = map_to_cta(full_bar(stage), 0);
            const int slab0 = (jt * 2 + (int)rank) * SLABS, row0 = wt * N_TILE + (int)rank * S_ROWS;
            if (X3) {
              mbar_expect_tx(xfull_bar(stage), (uint32_t)X_BYTES);
              tma_load_3d(dst, &maps.x[s], xfull_bar(stage), 0, kb * BLOCK_K, slab0);
              if (leader) mbar_expect_tx(full_bar(stage), (uint32_t)(4 * S_TX));
              tma2_load_2d(dst + XB, &maps.s[s], lfull, kb * BLOCK_K, row0);
              tma2_load_2d(dst + XB + S_BYTES, &maps.slo[s], lfull, kb * BLOCK_K, row0);
            } else {
              if (leader) mbar_expect_tx(full_bar(stage), (uint32_t)(2 * (X_BYTES + S_TX)));
              tma2_load_3d(dst, &maps.x[s], lfull, 0, kb * BLOCK_K, slab0);
              tma2_load_2d(dst + XB, &maps.s[s], lfull, kb * BLOCK_K, row0);
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================================================== MMA issuer (leader CTA only): M = 256 over the pair, N = 256
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(N_TILE >> 3) << 17) |
                           ((uint32_t)(256 >> 4) << 24);   // M = 256: 128 rows in each CTA's tensor memory
    int stage = 0, acc = 0;
    uint32_t phase = 0, accphase = 0;
    bool ok = true;
    for (int tile = cl; tile < p.total_tiles && ok; tile += ncl) {
      if (!mbar_wait_warp(tempty_bar(acc), accphase ^ 1u, 2)) break;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
      const int nk_total = (p.kcat ? p.nsup : 1) * p.nkb;
      for (int it = 0; it < nk_total; ++it) {
        if (!mbar_wait_warp(full_bar(stage), phase, 3)) { ok = false; break; }
        if (X3 && !mbar_wait_warp(split_bar(stage), phase, 5)) { ok = false; break; }   // both X tiles landed and are split
        tc_fence_after();
        const uint32_t xs = base + stage * STAGE;
        const uint32_t bs = xs + XB;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            const uint64_t adesc = make_desc(xs + kk * (UMMA_K * 128), BLOCK_K * 128, 4 * 128, 1);
            const uint64_t bdesc = make_desc(bs + kk * (UMMA_K * 4), 16, S_SBO, S_LAYOUT);
            tc2_mma_tf32(d_tmem, adesc, bdesc, idesc, (it > 0 || kk > 0) ? 1u : 0u);
            if (X3) {
              tc2_mma_tf32(d_tmem, adesc, make_desc(bs + S_BYTES + kk * (UMMA_K * 4), 16, S_SBO, S_LAYOUT), idesc, 1u);
              tc2_mma_tf32(d_tmem, make_desc(xs + X_BYTES + kk * (UMMA_K * 128), BLOCK_K * 128, 4 * 128, 1), bdesc, idesc, 1u);
            }
          }
          tc2_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      if (!ok) break;
      if (elect_one()) tc2_commit(tfull_bar(acc));
      __syncwarp();
      acc ^= 1;
      if (acc == 0) accphase ^= 1u;
    }
  } else if (X3 && (warp == 2 || warp == 3)) {
    // ===================================================== splitter (both CTAs): X_lo = X - tf32_trunc(X) of the CTA's own tile
    const int t64 = threadIdx.x - 64;
    int stage = 0;
    uint32_t phase = 0;
    bool ok = true;
    for (int tile = cl; tile < p.total_tiles && ok; tile += ncl) {
      const int nk_total = (p.kcat ? p.nsup : 1) * p.nkb;
      for (int it = 0; it < nk_total; ++it) {
        if (!mbar_wait(xfull_bar(stage), phase, 6)) { ok = false; break; }
        const float4* src = reinterpret_cast<const float4*>(smem + (size_t)stage * STAGE);
        float4* dst = reinterpret_cast<float4*>(smem + (size_t)stage * STAGE + X_BYTES);
        constexpr int NV = X_BYTES / 16 / 64;
        static_assert(NV * 64 * 16 == X_BYTES, "splitter: whole float4s per thread");
        float4 v[NV];
#pragma unroll
        for (int u = 0; u < NV; ++u) v[u] = src[t64 + 64 * u];
#pragma unroll
        for (int u = 0; u < NV; ++u)
          dst[t64 + 64 * u] = make_float4(tf32_lo(v[u].x), tf32_lo(v[u].y), tf32_lo(v[u].z), tf32_lo(v[u].w));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 64;" ::: "memory");     // the 64 splitter threads; then ONE arrive per CTA.  (A
        // release at CLUSTER scope here cost ~2000 cycles per stage: the 3xTF32 pair kernel ran at half the speed of the
        // one-CTA kernel.  The consumer of X_lo is this CTA's own tensor core -- each SM of a pair reads its own A rows --
        // so the proxy fence above plus the default CTA-scope release is what orders the writes before the MMA.)
        if (t64 == 0) mbar_arrive_cluster_relaxed(map_to_cta(split_bar(stage), 0));
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (both CTAs): own 128 TMEM lanes = own 4 slabs
    // Two sets of four warps drain every accumulator, each set its own half of the column groups: most launches of the
    // model give a cluster one to four tiles, so the last tile's epilogue -- one warp per scheduler, ~100 KB of adds
    // and stores -- is not hidden behind anything; splitting it halves that tail.
    const int ew = (warp - 4) & 3, eh = (warp - 4) >> 2;
    int acc = 0;
    uint32_t accphase = 0;
    for (int tile = cl; tile < p.total_tiles; tile += ncl) {
      const int o = tile / per_out, rem = tile - o * per_out;
      const int jt = rem / p.n_wt, wt = rem - jt * p.n_wt;
      if (!mbar_wait(tfull_bar(acc), accphase, 4)) break;
      tc_fence_after();
      const int slab = (jt * 2 + (int)rank) * SLABS + ew;
      const bool slab_ok = slab < p.nslabs;
      const size_t srow = (size_t)(slab_ok ? slab : 0) * p.V * CH + lane;
      float* y = p.Y[o] + srow;
      const float* ad = p.add[o] ? p.add[o] + srow : nullptr;
      const float* ad2 = nullptr;
      if (p.add2 && slab_ok) {
        const int b = slab / p.L, l = slab - b * p.L;
        if (l >= p.L - p.T_out) ad2 = p.add2 + ((size_t)(b * p.T_out + (l - (p.L - p.T_out))) * p.V) * CH + lane;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * ew) << 16) + (uint32_t)(acc * ACC_COLS);
      const int w_base = wt * N_TILE;
      const bool has_ad = slab_ok && ad != nullptr, has_ad2 = slab_ok && ad2 != nullptr;   // warp-uniform
      constexpr int GC = 64;
      const int ngrp = (N_TILE + GC - 1) / GC, gsplit = ((ngrp + 1) >> 1) * GC;   // set 0: groups [0, ceil(ngrp/2)), set 1: the rest
      for (int g0 = eh ? gsplit : 0; g0 < (eh ? N_TILE : gsplit); g0 += GC) {
        if (w_base + g0 >= p.V) break;   // warp-uniform: nothing left in this column tile
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
          if (g0 + cc < N_TILE) {   // warp-uniform
            uint32_t r[16];
            tc_ld16(taddr + g0 + cc, r);
            tc_wait_ld();
            if (slab_ok) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int w = w_base + g0 + cc + j;
                if (w < p.V) y[(size_t)w * CH] = __uint_as_float(r[j]) + av[cc + j];
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(map_to_cta(tempty_bar(acc), 0));
      acc ^= 1;
      if (acc == 0) accphase ^= 1u;
    }
  }

  tc_fence_before();
  cluster_sync_all();        // neither CTA may leave (or free tensor memory) while its peer's MMAs / arrivals can still touch it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace tc2

// Eligible: batched or K-concatenated supports shared by all samples.  Returns -1 when not eligible.
template <bool X3, int BLOCK_K>
static int node_gemm_tc2_impl(const NodeTcArgs& a, cudaStream_t stream) {
  using namespace tc2;
  constexpr int X_BYTES = SLABS * BLOCK_K * CH * 4;
  const long long nslabs = (long long)a.B * a.L;
  Maps maps;
  Params p;
  memset(&p, 0, sizeof(p));
  p.nsup = a.nsup; p.kcat = a.kcat; p.V = a.V; p.L = a.L; p.T_out = a.T_out; p.nslabs = (int)nslabs;
  p.n_jt = (int)((nslabs + 2 * SLABS - 1) / (2 * SLABS));
  p.nkb = (a.V + BLOCK_K - 1) / BLOCK_K;
  const int nout = a.kcat ? 1 : a.nsup;
  const int n_clusters = (tc_num_sms() & ~1) / 2;
  {   // V > 256: the fewest column tiles of <= 256 nodes, BALANCED (V = 325: 2 x 176 instead of 256 + a 69-node tile padded to 256)
    const int nw = (a.V + 255) / 256;
    p.n_tile = round_up((a.V + nw - 1) / nw, 16);
  }
  if (a.V <= 256) {
    // Small graphs: 8-slab row tiles alone are coarse against 74 clusters (METR-LA, K-concatenated dX sum at L = 12:
    // 96 tiles = 1.3 rounds).  Splitting the output columns multiplies the tile count at the price of narrower (per
    // column more expensive) MMAs; pick the split with the smallest modelled makespan, as the one-CTA kernel does.
    static const int forced = [] {
      const char* e = getenv("GWNET_B200_NCONV2_NWT");
      return e ? atoi(e) : 0;
    }();
    const int nk_total = (a.kcat ? a.nsup : 1) * p.nkb;
    long long best = -1;
    for (int nw = 1; nw <= 4; ++nw) {
      if (forced > 0 && nw != forced) continue;
      const int nt = round_up((a.V + nw - 1) / nw, 16);
      const int nwe = (a.V + nt - 1) / nt;
      const long long tiles_ = (long long)p.n_jt * nwe * nout;
      const long long rounds = (tiles_ + n_clusters - 1) / n_clusters;
      const long long cost = rounds * ((long long)nk_total * ((nt < 64 ? 64 : nt) + 48) + 400);
      if (best < 0 || cost < best) { best = cost; p.n_tile = nt; }
    }
  }
  p.n_wt = (a.V + p.n_tile - 1) / p.n_tile;
  const int S_BYTES = ((p.n_tile / 2) * BLOCK_K * 4 + 1023) & ~1023;
  const int stage_bytes = (X3 ? 2 : 1) * (X_BYTES + S_BYTES);
  p.stages = (tc::SMEM_LIMIT - 2048) / stage_bytes;
  if (p.stages > MAXSTAGES) p.stages = MAXSTAGES;
  const long long tiles = (long long)p.n_jt * p.n_wt * nout;
  if (tiles > 2147483647LL) return -1;
  p.total_tiles = (int)tiles;
  for (int s = 0; s < a.nsup; ++s) {
    if ((reinterpret_cast<uintptr_t>(a.X[s]) & 15) || (reinterpret_cast<uintptr_t>(a.S[s]) & 15)) {
      set_error("node_gemm_tc2: operands must be 16-byte aligned");
      return GWN_ERR_UNSUPPORTED;
    }
    cuuint64_t xd[3] = {(cuuint64_t)CH, (cuuint64_t)a.V, (cuuint64_t)nslabs};
    cuuint64_t xs[2] = {(cuuint64_t)CH * 4, (cuuint64_t)a.V * CH * 4};
    cuuint32_t xb[3] = {CH, BLOCK_K, SLABS};
    GWN_TRY(tc::encode(&maps.x[s], a.X[s], 3, xd, xs, xb, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    cuuint64_t sd[2] = {(cuuint64_t)a.V, (cuuint64_t)a.V};
    cuuint64_t ss[1] = {(cuuint64_t)a.ld * 4};
    cuuint32_t sb[2] = {BLOCK_K, (cuuint32_t)(p.n_tile / 2)};
    const CUtensorMapSwizzle s_swz = BLOCK_K == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    GWN_TRY(tc::encode(&maps.s[s], a.S[s], 2, sd, ss, sb, s_swz));
    if (X3) {
      if (!a.Slo[s] || (reinterpret_cast<uintptr_t>(a.Slo[s]) & 15)) {
        set_error("node_gemm_tc2: 3xTF32 mode needs 16-byte aligned support remainders");
        return GWN_ERR_UNSUPPORTED;
      }
      GWN_TRY(tc::encode(&maps.slo[s], a.Slo[s], 2, sd, ss, sb, s_swz));
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
  const int smem_bytes = p.stages * stage_bytes + 1024 /*alignment slack*/ + 8 * (4 * MAXSTAGES + 4) + 16;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(nconv_tc2_kernel<X3, BLOCK_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT);
  });
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(attr_err));
    return GWN_ERR_CUDA;
  }
  int num_sms = tc_num_sms() & ~1;
  long long want = 2 * tiles;
  const int grid = (int)(want < num_sms ? want : num_sms);
  GWN_CUDA(launch_kernel(nconv_tc2_kernel<X3, BLOCK_K>, dim3(grid), dim3(NUM_THREADS), smem_bytes, stream, maps, p));
  count_launch();
  return 0;
}

}  // namespace gwn
#endif
