// Position GEMM on tcgen05 / TMEM fed by TMA (tf32 tier):
//
//     out[m, n] = epilogue( sum_seg sum_k  A_seg[row_seg(m)][col0_seg + k] * Wp[n][seg*32 + k] )
//
//   m : BLNC position rows, tiled 128 at a time inside each of `nb` samples (so that a time-shifted source --
//       the second tap of the dilated conv, or dpre in its input gradient -- is one 3-D TMA box per segment,
//       out-of-range rows arriving as zeros);  K = nseg x 32;  N = output channels (multiple of 16, <= 256).
//   A : K-major SWIZZLE_128B tiles [128 rows][32 floats] straight from the fp32 activations (kind::tf32 reads fp32 bits);
//   Wp: packed weights [N][K] row-major, loaded once per persistent CTA and kept resident in shared memory;
//   D : fp32 in TMEM, double-buffered; the 4 epilogue warps own one position row per thread and run the same
//       row epilogues of rowepi.cuh (whole-row 128-bit loads of the addends, 128-bit stores).
//
// Used for the gated (1,2) conv and its backward recompute, the gcn mlp forward and input gradient, and the
// gated conv's input gradient -- the HBM-bound convolutions of SURVEY.md §8(a) rows a3, a4, a7, a9.
#pragma once
#include "functors.cuh"
#include "rowepi.cuh"
#include "tc_common.cuh"

namespace gwn {

constexpr int TP_MAXSEG = 16;

struct TcPosSeg {
  const float* src;   // [nb][rows_src][row_width] fp32
  int rows_src;       // rows per sample in the source
  int row_width;      // floats per source row (32 or 64)
  int col0;           // first column of this K segment inside the source row
  int rshift;         // source row = output row + rshift (may be negative; out-of-range rows read as 0)
};
struct TcPosArgs {
  TcPosSeg seg[TP_MAXSEG];
  int nseg;
  int nb;             // samples
  int rows_out;       // output rows per sample
  const float* Wp;    // packed weights [N_total][nseg*32]
  int N;              // output columns per tile (multiple of 16, <= 256)
  int wstream;        // 0: the whole [N][K] weight matrix stays resident in shared memory (N_total == N);
                      // 1: weights too large for that (head GEMMs): the k-block of the tile's N rows is streamed with A
  int N_total;        // wstream: total output columns (multiple of N); 0 = N
  int w_k, w_rows;    // actual extent of the weight matrix when smaller than nseg*32 x N_total (rest reads as zero); 0 = full
  const float* Wp_lo; // non-null: 3xTF32 (fp32-grade) mode -- Wp holds the tf32-exact high parts, Wp_lo the remainders
  // output tensor, written by TMA from the swizzled staging tile: out_nblk 32-column blocks per row tile;
  //   out_blk_dim2 == 0: [nb][rows_out][out_width], block k = columns 32k..32k+31;
  //   out_blk_dim2 == 1: [out_nblk][rows_out][32] (one tensor per block, nb == 1).
  float* out;
  int out_width, out_nblk, out_blk_dim2;
  // epilogue addend tiles (rows of 32 floats, same row tiling as the output, rshift like a segment); addend[k].src may
  // be null (slot unused); the epilogue functor knows which it reads
  TcPosSeg addend[2];
};

// Dummy "tile" for the column-statistics state of row-owner epilogues: 8 slots of 4 columns = 32 columns.
struct TRow {
  static constexpr int SLOTS = 8, NTL = 8, BN = 32, NT = 128, TX = 8, GN = 1, WM = 4, WTN = 32, BM = 128;
};

#if !GWN_EMU
namespace tc {

constexpr int TP_A_BYTES = 128 * 128;   // one A tile: 128 rows x 32 floats

struct TpMaps {
  CUtensorMap a[TP_MAXSEG];
  CUtensorMap w;
  CUtensorMap wlo;
  CUtensorMap out;
  CUtensorMap add[2];
};
struct TpParams {
  int nseg, nb, rows_out, tiles_per_sample, total_tiles, N, stages, out_blk_dim2, wstream, n_nt;
  int add_on[2], add_rshift[2], nadd;
  int col0[TP_MAXSEG], rshift[TP_MAXSEG];
};

// X3 = 3xTF32 mode (fp32-grade results on the tf32 tensor pipe): every operand is split into a tf32-exact high part
// and a remainder, D = A_hi.W_hi + A_hi.W_lo + A_lo.W_hi (the dropped A_lo.W_lo term is ~2^-22 relative).  The weight
// planes are precomputed; the A remainder is produced in shared memory by two otherwise idle warps (2, 3) from the
// tile TMA just landed: the element-wise split preserves the swizzled layout.  kind::tf32 ignores the low 13
// mantissa bits of its operands (probed: tests/tools/tf32_rounding_probe.py), so the fp32 A tile itself serves as A_hi.
// BK = channels of a K segment per pipeline stage: 32 (default: one stage per segment, rows of 128 B, SWIZZLE_128B) or 16
// (opt-in experiment: two half-segment stages, 64-byte rows, SWIZZLE_64B; see launch_tcpos).  Resident weights only.
template <class EP, int NCT, bool X3, int BK = 32>   // NCT > 0: compile-time column count (keeps per-slot epilogue state in registers)
__global__ void __launch_bounds__(128 + 128 * EP::kGroups, 1) tcpos_kernel(const __grid_constant__ TpMaps maps, const TpParams p, const EP ep_in) {
  constexpr int A_BYTES = 128 * BK * 4;                   // one stage of the A operand: 128 rows x BK floats
  constexpr int KH = 32 / BK;                             // stages per 32-channel K segment
  constexpr uint32_t A_SBO = 8 * BK * 4, A_LAYOUT = BK == 32 ? 2u : 4u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  constexpr int NPL = X3 ? 2 : 1;                         // operand planes (hi, lo)
  // N-concatenated remainder product (3xTF32, resident weights, compile-time N <= 64): W_lo's k-block sits right behind
  // W's, so A.[W | W_lo] is ONE instruction of twice the width into accumulator columns [0,N) and [N,2N) that the
  // epilogue adds; with the measured 66 + 0.75 N cycles per tcgen05.mma two N = 32 instructions (180 cycles) become
  // one N = 64 instruction (114).  Per k-step: 2 instructions instead of 3.
  // Streamed weights (head layers, N <= 128 per tile) keep W_lo's k-block behind W's inside every stage: same trick.
  const bool NCAT = X3 && !EP::kDirectStore && (NCT > 0 ? NCT <= 64 : (p.wstream != 0 && p.N <= 128));
  const int w_blk = p.N * 128;                            // one k-block of weights: [N rows][128 B]
  // Raw stages [A | (W blk | W_lo blk)] are what TMA writes -- the bytes in flight; the remainders A_lo live in their OWN
  // ring of LQ buffers (as in tcred.cuh): with A_lo inside every stage a CTA had three stages beside the mlp's resident
  // weights and staging tiles, 48 KB in flight per SM, and these kernels are bound by exactly that.  A remainder buffer
  // is busy from the split until its MMA retires (commit -> loempty); a raw stage from the TMA until the same commit --
  // NOT earlier: the splitter reads the raw tile, so the stage may only be handed back once split[l] has been seen.
  constexpr int LQ = X3 ? 2 : 0;
  const int STG = A_BYTES + (p.wstream ? NPL * w_blk : 0);
  const int w_plane = p.nseg * w_blk;                     // resident weights: [plane][seg][N rows][128 B]
  const int w_bytes = p.wstream ? 0 : NPL * w_plane;
  const uint32_t a0 = base + w_bytes;                     // raw stages
  const uint32_t lo0 = a0 + p.stages * STG;               // remainder ring
  constexpr int NSO = (EP::kGroups == 1) ? 2 : 2 * ((NCT > 0 && NCT / EP::kAccPerBlock == 1) ? 1 : 2);   // output staging tiles
  const uint32_t so0 = lo0 + LQ * A_BYTES;                // [128 rows][128 B] each, SWIZZLE_128B
  uint8_t* so_ptr = smem + w_bytes + (size_t)p.stages * STG + (size_t)LQ * A_BYTES;
  constexpr int NADD = EP::kAddends;                      // addend tiles: 2 buffers (tile parity) x NADD
  const uint32_t ad0 = so0 + NSO * TP_A_BYTES;
  const uint8_t* ad_ptr = so_ptr + NSO * TP_A_BYTES;
  const uint32_t bar0 = ad0 + 2 * NADD * TP_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(so_ptr + (NSO + 2 * NADD) * TP_A_BYTES);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (p.stages + s); };
  const int nb2 = 2 * p.stages;
  auto split_bar = [&](int l) { return bar0 + 8u * (nb2 + l); };          // remainder buffer l filled (l < 2)
  auto loempty_bar = [&](int l) { return bar0 + 8u * (nb2 + 2 + l); };    // ... and consumed
  auto tfull_bar = [&](int a) { return bar0 + 8u * (nb2 + 4 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (nb2 + 6 + a); };
  const uint32_t w_bar = bar0 + 8u * (nb2 + 8);
  auto efull_bar = [&](int e) { return bar0 + 8u * (nb2 + 9 + e); };
  auto eempty_bar = [&](int e) { return bar0 + 8u * (nb2 + 11 + e); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + nb2 + 13);
  float* red = reinterpret_cast<float*>(bars + nb2 + 14);   // 2 x 64 floats for the statistics reduces

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[s]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.out) : "memory");
    for (int k = 0; k < NADD; ++k)
      if (p.add_on[k]) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.add[k]) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int l = 0; l < 2; ++l) {
      mbar_init(split_bar(l), 64);
      mbar_init(loempty_bar(l), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    mbar_init(w_bar, 1);
    for (int e = 0; e < 2; ++e) {
      mbar_init(efull_bar(e), 1);
      mbar_init(eempty_bar(e), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  GWN_PDL_ENTRY();   // prologue above (barriers, TMEM, tensor-map prefetch) overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp loops, one elected lane issues)
    if (!p.wstream) {
      if (elect_one()) {
        mbar_expect_tx(w_bar, (uint32_t)w_bytes);
        if (NCAT) {   // [seg][W blk | W_lo blk]
          for (int s = 0; s < p.nseg; ++s) {
            tma_load_2d(base + s * 2 * w_blk, &maps.w, w_bar, s * 32, 0);
            tma_load_2d(base + s * 2 * w_blk + w_blk, &maps.wlo, w_bar, s * 32, 0);
          }
        } else {
          for (int s = 0; s < p.nseg; ++s) tma_load_2d(base + s * w_blk, &maps.w, w_bar, s * 32, 0);
          if (X3)
            for (int s = 0; s < p.nseg; ++s) tma_load_2d(base + w_plane + s * w_blk, &maps.wlo, w_bar, s * 32, 0);
        }
      }
      __syncwarp();
    }
    int stage = 0, eb = 0;
    uint32_t phase = 0, ephase = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.total_tiles && ok; tile += gridDim.x) {
      const int nt = tile % p.n_nt, rtile = tile / p.n_nt;
      const int b = rtile / p.tiles_per_sample, rt = rtile - b * p.tiles_per_sample;
      if (NADD > 0 && p.nadd > 0) {   // the epilogue's addend tiles of this row tile
        if (!mbar_wait_warp(eempty_bar(eb), ephase ^ 1u, 18)) { ok = false; break; }
        if (elect_one()) {
          mbar_expect_tx(efull_bar(eb), (uint32_t)p.nadd * TP_A_BYTES);
          for (int k = 0; k < NADD; ++k)
            if (p.add_on[k])
              tma_load_3d(ad0 + (eb * NADD + k) * TP_A_BYTES, &maps.add[k], efull_bar(eb), 0, rt * 128 + p.add_rshift[k], b);
        }
        __syncwarp();
        eb ^= 1;
        if (eb == 0) ephase ^= 1u;
      }
      for (int sh = 0; sh < p.nseg * KH; ++sh) {
        const int s = sh / KH, h = sh - s * KH;
        if (!mbar_wait_warp(empty_bar(stage), phase ^ 1u, 11)) { ok = false; break; }
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), A_BYTES + (p.wstream ? NPL * w_blk : 0));
          tma_load_3d(a0 + stage * STG, &maps.a[s], full_bar(stage), p.col0[s] + h * BK, rt * 128 + p.rshift[s], b);
          if (p.wstream) {
            const uint32_t wdst = a0 + stage * STG + A_BYTES;
            tma_load_2d(wdst, &maps.w, full_bar(stage), s * 32, nt * p.N);
            if (X3) tma_load_2d(wdst + w_blk, &maps.wlo, full_bar(stage), s * 32, nt * p.N);
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer: D[128 x N] += A[128 x 32] . Wseg[N x 32]^T
    // (whole warp loops, one elected lane issues)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int stage = 0, acc = 0, lq = 0;
    uint32_t phase = 0, accphase = 0, lphase = 0;
    bool ok = p.wstream ? true : mbar_wait_warp(w_bar, 0, 12);
    tc_fence_after();
    for (int tile = blockIdx.x; tile < p.total_tiles && ok; tile += gridDim.x) {
      if (!mbar_wait_warp(tempty_bar(acc), accphase ^ 1u, 13)) break;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int sh = 0; sh < p.nseg * KH; ++sh) {
        const int s = sh / KH, h = sh - s * KH;
        if (!mbar_wait_warp(full_bar(stage), phase, 14)) { ok = false; break; }
        tc_fence_after();
        const uint32_t as = a0 + stage * STG;
        // this stage's k-block of the weights: the segment's [N rows][128 B] tile, h * BK floats into every row
        const uint32_t ws = (p.wstream ? as + A_BYTES : (NCAT ? base + s * 2 * w_blk : base + s * w_blk)) + h * (BK * 4);
        const uint32_t wlo_off = p.wstream ? (uint32_t)w_blk : (uint32_t)w_plane;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            // both operands K-major: A rows of BK floats (8-row groups A_SBO apart), W rows of 128 B (SWIZZLE_128B,
            // 8-row groups 1024 B apart); this k-step 32 B further in
            const uint64_t ad = make_desc(as + kk * 32, 16, A_SBO, A_LAYOUT), wd = make_desc(ws + kk * 32, 16, 1024);
            if (NCAT) {   // A.[W | W_lo]: 2N columns
              const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * p.N) >> 3) << 17);
              tc_mma_tf32(d_tmem, ad, wd, idesc2, (sh > 0 || kk > 0) ? 1u : 0u);
            } else {
              tc_mma_tf32(d_tmem, ad, wd, idesc, (sh > 0 || kk > 0) ? 1u : 0u);
              if (X3) tc_mma_tf32(d_tmem, ad, make_desc(ws + wlo_off + kk * 32, 16, 1024), idesc, 1u);
            }
          }
        }
        __syncwarp();
        if (X3) {   // the A_lo term last: the split of this stage overlaps the MMAs above
          if (!mbar_wait_warp(split_bar(lq), lphase, 16)) { ok = false; break; }
          tc_fence_after();
          const uint32_t al = lo0 + (uint32_t)(lq * A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk)
              tc_mma_tf32(d_tmem, make_desc(al + kk * 32, 16, A_SBO, A_LAYOUT), make_desc(ws + kk * 32, 16, 1024), idesc, 1u);
          }
          __syncwarp();
        }
        if (elect_one()) {
          tc_commit(empty_bar(stage));             // after split[lq]: the splitter has read the raw tile
          if (X3) tc_commit(loempty_bar(lq));
        }
        __syncwarp();
        if (X3) {
          if (++lq == LQ) { lq = 0; lphase ^= 1u; }
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (!ok) break;
      if (elect_one()) tc_commit(tfull_bar(acc));
      __syncwarp();
      acc ^= 1;
      if (acc == 0) accphase ^= 1u;
    }
  } else if (X3 && (warp == 2 || warp == 3)) {
    // ===================================================== splitter: A_lo = A - tf32_trunc(A), 64 threads, 16 float4 each
    const int t64 = threadIdx.x - 64;
    int stage = 0, lq = 0;
    uint32_t phase = 0, lphase = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.total_tiles && ok; tile += gridDim.x) {
      for (int sh = 0; sh < p.nseg * KH; ++sh) {
        if (!mbar_wait(loempty_bar(lq), lphase ^ 1u, 20)) { ok = false; break; }   // the MMA that read this buffer has retired
        if (!mbar_wait(full_bar(stage), phase, 17)) { ok = false; break; }
        const float4* src = reinterpret_cast<const float4*>(smem + w_bytes + (size_t)stage * STG);
        float4* dst = reinterpret_cast<float4*>(smem + w_bytes + (size_t)p.stages * STG + (size_t)lq * A_BYTES);
        {   // 16 (8) float4 per thread: all loads in flight before the first use (with 4 at a time the two splitter
            // warps were busy ~100 % of the time, stalled on LDS results: ncu source page of the reduction kernel)
          constexpr int NV = A_BYTES / 16 / 64;
          static_assert(NV * 64 * 16 == A_BYTES, "splitter: whole float4s per thread");
          float4 v[NV];
#pragma unroll
          for (int u = 0; u < NV; ++u) v[u] = src[t64 + 64 * u];
#pragma unroll
          for (int u = 0; u < NV; ++u)
            dst[t64 + 64 * u] = make_float4(tf32_lo(v[u].x), tf32_lo(v[u].y), tf32_lo(v[u].z), tf32_lo(v[u].w));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(split_bar(lq));
        if (++lq == LQ) { lq = 0; lphase ^= 1u; }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: two groups of 4 warps, one position row per thread.
    // Group g owns accumulator buffer g, i.e. every other tile of this CTA, so two tiles' epilogues run concurrently:
    // with a single group the kernels were bound by the latency of ONE warp's instruction stream per scheduler
    // (ncu r01d: RowMlp ~1100 dependent instructions per tile at ~8 cycles each = 4.5 us per tile vs 2.9 us of HBM time).
    EP ep = ep_in;
    ep.init();
    constexpr int NG = EP::kGroups;
    const int grp = (warp - 4) >> 2, ew = warp & 3;
    const int barid = 1 + grp;
    const bool elected = (threadIdx.x & 127) == 0;
    constexpr int APB = EP::kAccPerBlock;
    constexpr int RING = (NG == 1 || (NCT > 0 && NCT / APB == 1)) ? (NG == 1 ? 2 : 1) : 2;      // staging tiles per group
    uint8_t* so_grp = so_ptr + grp * RING * TP_A_BYTES;
    const uint32_t so_grp32 = so0 + grp * RING * TP_A_BYTES;
    const int r = ew * 32 + lane;
    uint32_t ring = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if (NG == 2 && (it & 1) != grp) continue;
      const int buf = it & 1;                               // accumulator / addend buffer of this tile
      const uint32_t bphase = (uint32_t)(it >> 1) & 1u;     // ... and how often it has been used: its barrier parity
      const int nt = tile % p.n_nt, rtile = tile / p.n_nt;
      const int b = rtile / p.tiles_per_sample, rt = rtile - b * p.tiles_per_sample;
      const int n0 = nt * p.N;                              // first output column of this tile
      const int rl = rt * 128 + r;
      const bool valid = rl < p.rows_out;
      const i64 m = (i64)b * p.rows_out + rl;
      const bool has_add = NADD > 0 && p.nadd > 0;
      if (has_add) {        // this row's addends: staged tile -> registers, then hand the buffer back to the producer
        if (!mbar_wait(efull_bar(buf), bphase, 19)) break;
        AddendRows ar;
        ar.row[0] = ad_ptr + (size_t)(buf * NADD) * TP_A_BYTES + r * 128;
        ar.row[1] = ad_ptr + (size_t)(buf * NADD + (NADD > 1 ? 1 : 0)) * TP_A_BYTES + r * 128;
        ar.x = (uint32_t)(r & 7);
        ep.load_addends(ar, valid);
        if constexpr (!EP::kLazyAddends) {
          // Registers hold the rows from here on, so the buffer can go back to the producer -- but its next use is a TMA
          // write (async proxy) and these were generic-proxy reads: without the proxy fence the arrive was observed
          // while LDS data were still outstanding, and once the producer could run two tiles ahead (4 raw stages in
          // RowGateBwd) the next tile's rows landed in the buffer first: dpre rows off by ~1e-2.
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(eempty_bar(buf));
        }
      }
      if (!mbar_wait(tfull_bar(buf), bphase, 15)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * ew) << 16) + (uint32_t)(buf * 256);
      const int ncols = NCT > 0 ? NCT : p.N;
      auto do_block = [&](const int blk) {
        if constexpr (EP::kDirectStore) {   // tiny strided outputs (the network's NCHW result): straight from registers
          const RowSink sink{nullptr, 0u};
#pragma unroll
          for (int cc = 0; cc < APB; cc += 16) {
            const int c0 = blk * APB + cc;
            if (c0 < ncols) {
              uint32_t rr[16];
              tc_ld16(taddr + c0, rr);
              tc_wait_ld();
              if (valid) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
                ep.consume16(m, c0, n0, v, sink);
              }
            }
          }
        } else {
        // staging ring: the TMA store that last read this tile must have finished reading it
        if (elected) {
          if (RING == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
        const uint32_t slot = RING == 2 ? (ring & 1u) : 0u;
        const RowSink sink{so_grp + slot * TP_A_BYTES + r * 128, (uint32_t)(r & 7)};
#pragma unroll
        for (int cc = 0; cc < APB; cc += 16) {
          const int c0 = blk * APB + cc;
          uint32_t rr[16];
          tc_ld16(taddr + c0, rr);
          if (NCAT) {   // + the A.W_lo half of the accumulator
            uint32_t r2[16];
            tc_ld16(taddr + ncols + c0, r2);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) rr[j] = __float_as_uint(__uint_as_float(rr[j]) + __uint_as_float(r2[j]));
          } else {
            tc_wait_ld();
          }
          if (valid) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
            ep.consume16(m, c0, n0, v, sink);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staging writes -> visible to the TMA engine
        asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
        if (elected) {
          const int oblk = n0 / 32 + blk;                   // output block index across column tiles
          const int c0o = p.out_blk_dim2 ? 0 : 32 * oblk, c2o = p.out_blk_dim2 ? oblk : b;
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&maps.out),
                       "r"(so_grp32 + slot * TP_A_BYTES), "r"(c0o), "r"(rt * 128), "r"(c2o)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++ring;
        }
      };
      if constexpr (NCT > 0) {   // compile-time block count: the epilogue's per-column register arrays stay in registers
#pragma unroll
        for (int blk = 0; blk < NCT / APB; ++blk) do_block(blk);
      } else {
#pragma unroll 1
        for (int blk = 0; blk * APB < ncols; ++blk) do_block(blk);
      }
      if constexpr (EP::kLazyAddends) {   // consume16 read the staged addend tiles: hand them back only now
        if (has_add) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(eempty_bar(buf));
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
    }
    if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output tiles have landed
    if constexpr (EP::kHasFinish) {
      asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
      ep.finish_rows(red + 64 * grp, threadIdx.x & 127, barid);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace tc
#endif

// Returns 0 after launching; -1 when the shape is not eligible (caller falls back); > 0 on error.
template <int NCT, bool X3, int BK, class EP>
int launch_tcpos_impl(const TcPosArgs& a, const EP& ep, cudaStream_t stream) {
#if GWN_EMU
  (void)a; (void)ep; (void)stream;
  return -1;
#else
  using namespace tc;
  if (NCT > 0 && a.N != NCT) return -1;
  if (a.nseg < 1 || a.nseg > TP_MAXSEG || a.N < 16 || a.N > 256 || a.N % 16 != 0 || a.nb < 1 || a.rows_out < 1) return -1;
  const int N_total = (a.wstream && a.N_total > 0) ? a.N_total : a.N;
  if (N_total % a.N != 0) return -1;
  if (reinterpret_cast<uintptr_t>(a.Wp) & 15) return -1;
  TpMaps maps;
  TpParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = a.nseg; p.nb = a.nb; p.rows_out = a.rows_out; p.N = a.N;
  p.wstream = a.wstream ? 1 : 0;
  p.n_nt = N_total / a.N;
  p.tiles_per_sample = (a.rows_out + 127) / 128;
  const long long tiles = (long long)p.tiles_per_sample * a.nb * p.n_nt;
  if (tiles > 2147483647LL) return -1;
  p.total_tiles = (int)tiles;
  const int w_bytes = a.wstream ? 0 : (X3 ? 2 : 1) * a.nseg * a.N * 128;
  if (BK != 32 && a.wstream) return -1;                   // half-size stages: resident weights only
  const int LO_BYTES = X3 ? 2 * 128 * BK * 4 : 0;      // the remainder ring (2 buffers)
  const int STG = 128 * BK * 4 + (a.wstream ? (X3 ? 2 : 1) * a.N * 128 : 0);
  constexpr int NSO = (EP::kGroups == 1) ? 2 : 2 * ((NCT > 0 && NCT / EP::kAccPerBlock == 1) ? 1 : 2);
  constexpr int FIXED = (NSO + 2 * EP::kAddends) * TP_A_BYTES;   // output staging tiles + addend tiles
  p.stages = (SMEM_LIMIT - 2048 - w_bytes - FIXED - LO_BYTES) / STG;
  for (int k = 0; k < 2; ++k) {
    const TcPosSeg& g = a.addend[k];
    if (!g.src || k >= EP::kAddends) continue;
    if ((reinterpret_cast<uintptr_t>(g.src) & 15) || g.row_width != 32) return -1;
    cuuint64_t d[3] = {32, (cuuint64_t)g.rows_src, (cuuint64_t)a.nb};
    cuuint64_t st[2] = {128, (cuuint64_t)g.rows_src * 128};
    cuuint32_t box[3] = {32, 128, 1};
    GWN_TRY(encode(&maps.add[k], g.src, 3, d, st, box, CU_TENSOR_MAP_SWIZZLE_128B));
    p.add_on[k] = 1;
    p.add_rshift[k] = g.rshift;
    p.nadd += 1;
  }
  p.out_blk_dim2 = a.out_blk_dim2;
  if (!EP::kDirectStore &&
      (!a.out || (reinterpret_cast<uintptr_t>(a.out) & 15) || a.out_width % 4 != 0 || a.out_nblk < 1 ||
       a.out_nblk * EP::kAccPerBlock != N_total || (a.out_blk_dim2 ? (a.out_width != 32 || a.nb != 1) : (a.out_width < 32 * a.out_nblk))))
    return -1;
  if (EP::kDirectStore) {
    maps.out = maps.a[0];
  } else {
    cuuint64_t d[3] = {(cuuint64_t)a.out_width, (cuuint64_t)a.rows_out, (cuuint64_t)(a.out_blk_dim2 ? a.out_nblk : a.nb)};
    cuuint64_t st[2] = {(cuuint64_t)a.out_width * 4, (cuuint64_t)a.rows_out * a.out_width * 4};
    cuuint32_t box[3] = {32, 128, 1};
    GWN_TRY(encode(&maps.out, a.out, 3, d, st, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  if (p.stages > 8) p.stages = 8;
  {
    static const int cap = [] {   // diagnostics: cap the pipeline depth (GWNET_B200_TCPOS_STAGES)
      const char* e = getenv("GWNET_B200_TCPOS_STAGES");
      return e ? atoi(e) : 0;
    }();
    if (cap >= 2 && p.stages > cap) p.stages = cap;
  }
  if (p.stages < 2) return -1;
  for (int s = 0; s < a.nseg; ++s) {
    const TcPosSeg& g = a.seg[s];
    if ((reinterpret_cast<uintptr_t>(g.src) & 15) || g.row_width % 4 != 0 || g.col0 >= g.row_width) return -1;   // a box past the row end reads zeros
    cuuint64_t d[3] = {(cuuint64_t)g.row_width, (cuuint64_t)g.rows_src, (cuuint64_t)a.nb};
    cuuint64_t st[2] = {(cuuint64_t)g.row_width * 4, (cuuint64_t)g.rows_src * g.row_width * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, 128, 1};
    GWN_TRY(encode(&maps.a[s], g.src, 3, d, st, box, BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B));
    p.col0[s] = g.col0;
    p.rshift[s] = g.rshift;
  }
  for (int s = a.nseg; s < TP_MAXSEG; ++s) maps.a[s] = maps.a[0];
  {
    const int K = a.w_k > 0 ? a.w_k : a.nseg * 32;   // row length of the weight matrix (short rows are zero-filled)
    cuuint64_t d[2] = {(cuuint64_t)K, (cuuint64_t)(a.w_rows > 0 ? a.w_rows : N_total)};
    cuuint64_t st[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)a.N};
    GWN_TRY(encode(&maps.w, a.Wp, 2, d, st, box, CU_TENSOR_MAP_SWIZZLE_128B));
    if (X3) {
      if (reinterpret_cast<uintptr_t>(a.Wp_lo) & 15) return -1;
      GWN_TRY(encode(&maps.wlo, a.Wp_lo, 2, d, st, box, CU_TENSOR_MAP_SWIZZLE_128B));
    } else {
      maps.wlo = maps.w;
    }
  }
  for (int k = 0; k < 2; ++k)
    if (!p.add_on[k]) maps.add[k] = maps.out;
  const int smem_bytes = w_bytes + p.stages * STG + LO_BYTES + FIXED + 1024 + 1024;
  static cudaError_t attr = cudaFuncSetAttribute(tcpos_kernel<EP, NCT, X3, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  if (attr != cudaSuccess) {
    set_error("tcpos: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr));
    return GWN_ERR_CUDA;
  }
  static int num_sms = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  GWN_CUDA(launch_kernel(tcpos_kernel<EP, NCT, X3, BK>, dim3(grid), dim3(128 + 128 * EP::kGroups), smem_bytes, stream, maps, p, ep));
  count_launch();
  return 0;
#endif
}

template <int NCT, class EP>
int launch_tcpos(const TcPosArgs& a, const EP& ep, cudaStream_t stream) {
  if (!a.Wp_lo) return launch_tcpos_impl<NCT, false, 32>(a, ep, stream);
  // 3xTF32 with resident weights: GWNET_B200_TCPOS_BK=16 selects half-size stages (A/B runs).  Measured (r02o): no gain --
  // METR-LA step 2.685 vs 2.644 ms, mlp forward 316 vs 300 us: unlike the node contraction these kernels are bound by
  // the BYTES in flight per SM (three 16 KB raw planes beside 57 KB of weights and 64 KB of staging tiles), and six
  // half-size stages hold the same bytes as three full ones.
  static const bool half = [] {
    const char* e = getenv("GWNET_B200_TCPOS_BK");
    return e && atoi(e) == 16;
  }();
  if (half && !a.wstream) {
    const int st = launch_tcpos_impl<NCT, true, 16>(a, ep, stream);
    if (st >= 0) return st;
  }
  return launch_tcpos_impl<NCT, true, 32>(a, ep, stream);
}

}  // namespace gwn
