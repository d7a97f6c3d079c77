// The gcn hop chain of model.py:41-50 for small graphs (V <= 256, order 2) as ONE kernel on CTA pairs -- the kernel shape
// BASELINE.json's north star describes: the support stays RESIDENT in shared memory and the hops are chained in-tile.
//
//   per support s and 256-row tile (8 slabs x 32 channels):   D1 = X . S_s     (hop 1: x1 = nconv(x, a))
//                                                              D2 = D1 . S_s    (hop 2: x2 = nconv(x1, a))
//
//   * a cluster of two CTAs owns ONE support for the whole launch: each CTA loads its half of the support's rows (all k)
//     once -- 93 KB at V = 207 -- and keeps it in shared memory; per tile only X is streamed (cp.async.bulk.tensor,
//     8 stages of 16 KB per CTA).  Clusters are dealt round-robin to the supports, row tiles round-robin to a support's
//     clusters.
//   * hop 1 is the nconv_tc2 main loop (tcgen05.mma.cta_group::2, A = X tile from shared memory, MN-major).  Its
//     accumulator D1 -- 128 TMEM lanes per CTA x n_tile fp32 columns -- IS hop 2's A operand: `[lanes = (slab, c) rows,
//     columns = node w]` is exactly the K-major A layout tcgen05.mma takes from tensor memory, and kind::tf32 reads the
//     fp32 bits as tf32.  Hop 2 is n_tile / 8 MMAs with A = [d1 + 8 k], B = the resident support: no shared-memory A
//     traffic, no HBM or L2 round trip of hop 1.  The epilogue warps store D1 (x1 is an mlp input and needed by the
//     backward pass) while the tensor core already runs hop 2, then store D2.
//   * single-pass TF32 tier only: in the 3xTF32 tier the support's two planes (2 x 93 KB per CTA) do not fit beside the X
//     stages and hop 2 would need the hop-1 remainder as a third TMEM region (3 x 208 columns > 512).
// HBM traffic of a layer's hop chain: x read once per support (L2 serves the repeats) + 2 S hop tensors written, instead
// of additionally re-reading the S hop-1 tensors in a second launch.
#pragma once
#include "nconv_tc2.cuh"

#if !GWN_EMU

namespace gwn {
namespace hopf {

using tc::smem_u32; using tc::mbar_init; using tc::mbar_expect_tx; using tc::mbar_wait; using tc::mbar_wait_warp;
using tc::elect_one; using tc::uniform_warp_id; using tc::tc_fence_before; using tc::tc_fence_after;
using tc::tc_ld16; using tc::tc_wait_ld; using tc::make_desc;
using tc2::cluster_rank; using tc2::map_to_cta; using tc2::mbar_arrive_cluster_relaxed; using tc2::cluster_sync_all;
using tc2::tma2_load_3d; using tc2::tma2_load_2d; using tc2::tc2_commit; using tc2::tc2_mma_tf32;

constexpr int SLABS = 4, CH = 32, BLOCK_K = 32, UMMA_K = 8;
constexpr int X_BYTES = SLABS * BLOCK_K * CH * 4;   // 16 KB
constexpr int NUM_THREADS = 256;
constexpr int MAXSTAGES = 8;
constexpr int D2_COL = 256;                         // TMEM: D1 in columns [0, n_tile), D2 in [256, 256 + n_tile)

struct Maps {
  CUtensorMap x;
  CUtensorMap s[TC_MAXSUP];
};
struct Params {
  float* Y1[TC_MAXSUP];
  float* Y2[TC_MAXSUP];
  int nsup, V, nslabs, n_jt, nkb, stages, n_tile;
};

// A operand from tensor memory (K-major by construction), B from shared memory
__device__ __forceinline__ void tc2_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) gcn_hops_fused_kernel(const __grid_constant__ Maps maps, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int N_TILE = p.n_tile, S_ROWS = N_TILE >> 1;
  const int S_BLK = S_ROWS * BLOCK_K * 4;                // one k-block of this CTA's support half: [S_ROWS][128 B], SWIZZLE_128B
  const int S_BLKP = (S_BLK + 1023) & ~1023;             // 1024-byte aligned pitch between k-blocks
  const int sres_bytes = p.nkb * S_BLKP;
  const uint32_t sres0 = base;                           // resident support half
  const uint32_t x0 = base + sres_bytes;                 // X stages
  const int stages = p.stages;
  const uint32_t bar0 = x0 + stages * X_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAXSTAGES + s); };
  const uint32_t sres_bar = bar0 + 8u * (2 * MAXSTAGES);
  const uint32_t d1full = bar0 + 8u * (2 * MAXSTAGES + 1), d2full = bar0 + 8u * (2 * MAXSTAGES + 2);
  const uint32_t d1empty = bar0 + 8u * (2 * MAXSTAGES + 3), d2empty = bar0 + 8u * (2 * MAXSTAGES + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)sres_bytes + (size_t)stages * X_BYTES + 8 * (2 * MAXSTAGES + 5));

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int cl = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int sup = cl % p.nsup;                           // this cluster's support
  const int cls = cl / p.nsup, ncls = (ncl - sup + p.nsup - 1) / p.nsup;   // its index / count among that support's clusters

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.s[sup]) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(sres_bar, 1);
    mbar_init(d1full, 1);
    mbar_init(d2full, 1);
    mbar_init(d1empty, 8);
    mbar_init(d2empty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  GWN_PDL_ENTRY();

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs): resident support half, then X tiles
    if (elect_one()) {
      const uint32_t lsres = map_to_cta(sres_bar, 0);
      if (leader) mbar_expect_tx(sres_bar, (uint32_t)(2 * p.nkb * S_BLK));
      for (int kb = 0; kb < p.nkb; ++kb)
        tma2_load_2d(sres0 + kb * S_BLKP, &maps.s[sup], lsres, kb * BLOCK_K, (int)rank * S_ROWS);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    bool ok = true;
    for (int jt = cls; jt < p.n_jt && ok; jt += ncls) {
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (!mbar_wait_warp(empty_bar(stage), phase ^ 1u, 1)) { ok = false; break; }
        if (elect_one()) {
          const uint32_t lfull = map_to_cta(full_bar(stage), 0);
          if (leader) mbar_expect_tx(full_bar(stage), (uint32_t)(2 * X_BYTES));
          tma2_load_3d(x0 + stage * X_BYTES, &maps.x, lfull, 0, kb * BLOCK_K, (jt * 2 + (int)rank) * SLABS);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================================================== MMA issuer (leader): hop 1 from shared memory, hop 2 from TMEM
    const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t idesc2 = idesc1 & ~(1u << 15);        // A from tensor memory: K-major
    int stage = 0;
    uint32_t phase = 0, tphase = 0;
    bool ok = mbar_wait_warp(sres_bar, 0, 7);
    tc_fence_after();
    const uint32_t d1 = tmem_base, d2 = tmem_base + D2_COL;
    for (int jt = cls; jt < p.n_jt && ok; jt += ncls) {
      if (!mbar_wait_warp(d1empty, tphase ^ 1u, 2)) break;          // both CTAs' epilogues have stored the previous D1
      tc_fence_after();
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (!mbar_wait_warp(full_bar(stage), phase, 3)) { ok = false; break; }
        tc_fence_after();
        const uint32_t xs = x0 + stage * X_BYTES, bs = sres0 + kb * S_BLKP;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
            tc2_mma_tf32(d1, make_desc(xs + kk * (UMMA_K * 128), BLOCK_K * 128, 4 * 128, 1), make_desc(bs + kk * (UMMA_K * 4), 16, 1024), idesc1,
                         (kb > 0 || kk > 0) ? 1u : 0u);
          tc2_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      if (!ok) break;
      if (elect_one()) tc2_commit(d1full);                           // hop 1 complete: the epilogue may store it ...
      __syncwarp();
      if (!mbar_wait_warp(d2empty, tphase ^ 1u, 4)) break;           // ... and, once the previous D2 is drained
      if (!mbar_wait_warp(d1full, tphase, 6)) break;                 // and D1 is final in both CTAs' tensor memory,
      tc_fence_after();
      if (elect_one()) {                                             // hop 2 reads it straight from tensor memory
        const int nk2 = N_TILE / UMMA_K;                             // K = the tile's node columns (padding columns are zero)
        for (int k2 = 0; k2 < nk2; ++k2) {
          const int kb = k2 >> 2, kk = k2 & 3;
          tc2_mma_tf32_ts(d2, d1 + (uint32_t)(k2 * UMMA_K), make_desc(sres0 + kb * S_BLKP + kk * (UMMA_K * 4), 16, 1024), idesc2,
                          k2 > 0 ? 1u : 0u);
        }
        tc2_commit(d2full);
      }
      __syncwarp();
      tphase ^= 1u;
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (both CTAs): D1 -> hop-1 tensor, D2 -> hop-2 tensor
    const int ew = warp - 4;
    uint32_t tphase = 0;
    for (int jt = cls; jt < p.n_jt; jt += ncls) {
      const int slab = (jt * 2 + (int)rank) * SLABS + ew;
      const bool slab_ok = slab < p.nslabs;
      const size_t srow = (size_t)(slab_ok ? slab : 0) * p.V * CH + lane;
      bool ok = true;
      for (int hop = 0; hop < 2 && ok; ++hop) {
        if (!mbar_wait(hop == 0 ? d1full : d2full, tphase, 5)) { ok = false; break; }
        tc_fence_after();
        float* y = (hop == 0 ? p.Y1[sup] : p.Y2[sup]) + srow;
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * ew) << 16) + (uint32_t)(hop * D2_COL);
        for (int c0 = 0; c0 < N_TILE; c0 += 16) {
          if (c0 >= p.V) break;   // warp-uniform
          uint32_t r[16];
          tc_ld16(taddr + c0, r);
          tc_wait_ld();
          if (slab_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.V) y[(size_t)(c0 + j) * CH] = __uint_as_float(r[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(map_to_cta(hop == 0 ? d1empty : d2empty, 0));
      }
      if (!ok) break;
      tphase ^= 1u;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace hopf

// Both hops of all supports in one launch.  Returns -1 when the shape is not eligible (the caller runs the two
// node-contraction launches instead).
int gcn_hops_fused_tc(const float* x, const float* const* S, int nsup, int ld, float* const* Y1, float* const* Y2, int B, int L, int V,
                      cudaStream_t stream) {
  using namespace hopf;
  static const bool on = [] {   // GWNET_B200_FUSED_HOPS=0: the two-launch hop chain (A/B runs)
    const char* e = getenv("GWNET_B200_FUSED_HOPS");
    return !(e && e[0] == '0');
  }();
  if (!on || V > 256 || V < 16 || nsup < 1 || nsup > TC_MAXSUP || ld % 4 != 0 || ld < V) return -1;
  const long long nslabs = (long long)B * L;
  if (nslabs <= 0 || nslabs > 2147483647LL) return -1;
  Maps maps;
  Params p;
  memset(&p, 0, sizeof(p));
  p.nsup = nsup; p.V = V; p.nslabs = (int)nslabs;
  p.n_jt = (int)((nslabs + 2 * SLABS - 1) / (2 * SLABS));
  p.nkb = (V + BLOCK_K - 1) / BLOCK_K;
  p.n_tile = round_up(V, 16);
  const int s_blkp = ((p.n_tile / 2) * BLOCK_K * 4 + 1023) & ~1023;
  const int sres_bytes = p.nkb * s_blkp;
  p.stages = (tc::SMEM_LIMIT - 2048 - sres_bytes) / X_BYTES;
  if (p.stages > MAXSTAGES) p.stages = MAXSTAGES;
  if (p.stages < 3) return -1;
  if ((reinterpret_cast<uintptr_t>(x) & 15)) return -1;
  {
    cuuint64_t xd[3] = {(cuuint64_t)CH, (cuuint64_t)V, (cuuint64_t)nslabs};
    cuuint64_t xs[2] = {(cuuint64_t)CH * 4, (cuuint64_t)V * CH * 4};
    cuuint32_t xb[3] = {CH, BLOCK_K, SLABS};
    GWN_TRY(tc::encode(&maps.x, x, 3, xd, xs, xb, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  }
  for (int s = 0; s < nsup; ++s) {
    if ((reinterpret_cast<uintptr_t>(S[s]) & 15)) return -1;
    cuuint64_t sd[2] = {(cuuint64_t)V, (cuuint64_t)V};
    cuuint64_t ss[1] = {(cuuint64_t)ld * 4};
    cuuint32_t sb[2] = {BLOCK_K, (cuuint32_t)(p.n_tile / 2)};
    GWN_TRY(tc::encode(&maps.s[s], S[s], 2, sd, ss, sb, CU_TENSOR_MAP_SWIZZLE_128B));
    p.Y1[s] = Y1[s];
    p.Y2[s] = Y2[s];
  }
  for (int s = nsup; s < TC_MAXSUP; ++s) maps.s[s] = maps.s[0];
  const int smem_bytes = sres_bytes + p.stages * X_BYTES + 1024 + 8 * (2 * MAXSTAGES + 5) + 16;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gcn_hops_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT);
  });
  if (attr_err != cudaSuccess) {
    set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(attr_err));
    return GWN_ERR_CUDA;
  }
  const long long tiles = (long long)p.n_jt * nsup;
  const int num_sms = tc_num_sms() & ~1;
  const int grid = (int)(2 * tiles < num_sms ? 2 * tiles : num_sms);
  GWN_CUDA(launch_kernel(gcn_hops_fused_kernel, dim3(grid), dim3(NUM_THREADS), smem_bytes, stream, maps, p));
  count_launch();
  return 0;
}

}  // namespace gwn
#endif
