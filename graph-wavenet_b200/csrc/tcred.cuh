// Reduction GEMMs on tcgen05 / TMEM fed by TMA (tf32 tier): the contractions whose K dimension is the huge one
// (all positions of the batch) and whose output is a small matrix accumulated over the whole batch:
//
//   MODE_MN  weight gradients      D[(blk, i), n] = sum_p  A_blk[p][i] * B[p][n]
//              A: up to 8 blocks of 32 columns taken from position-row tensors (the 7 concatenated gcn inputs,
//                 or the two taps of the gated conv's input) + one synthetic all-ones block whose rows deliver the
//                 column sums of B, i.e. the bias gradient;  B: the output gradient rows (dh or dpre), N = 32 or 64.
//              Both operands are MN-major (positions are K): 32-byte-atom 128B swizzle, like X in nconv_tc.
//   MODE_K   support gradient      D[v, w] = sum_{slab, c} X[slab][v][c] * T[slab][w][c]      
//              Both operands K-major (channels are K within a slab), plain 128B swizzle, like tcpos.
//
// Each persistent CTA takes a contiguous range of K chunks, accumulates its partial D in TMEM over the whole
// range (no intermediate epilogues), and adds it to the global result with coalesced red.global.add at the end.
#pragma once
#include "functors.cuh"
#include "tc_common.cuh"

namespace gwn {

constexpr int TR_MAXSRC = 16;

struct TcRedSrc {
  const float* src;   // [nb][rows_src][row_width]
  int rows_src, row_width, col0, rshift;
  int nb;             // MODE_K: slabs of this pair (0 = TcRedArgs::nb)
};
struct TcRedArgs {
  int mode;           // 0 = MODE_MN, 1 = MODE_K
  int x3;             // 1 = 3xTF32 (fp32-grade) mode
  TcRedSrc a[TR_MAXSRC];
  int na;             // MODE_MN: real A blocks (<= 7; the ones block is appended); MODE_K: number of (X, T) pairs
  TcRedSrc b[TR_MAXSRC];   // MODE_MN: b[0] only; MODE_K: one per pair
  int N;              // MODE_MN: B columns per CTA tile (multiple of 32, <= 256); MODE_K: ignored (derived from rows)
  int ab;             // MODE_MN: real A blocks per CTA tile (0 = all `na`, at most 7); the output is tiled
                      //   ceil(na/ab) block groups x (N_total/N) column groups, each with its own K-split CTAs
  int N_total;        // MODE_MN: B columns in total (0 = N)
  int nb, rows;       // MODE_MN: samples and B rows per sample; MODE_K: default slabs per pair, rows = V
  float* partial;     // scratch for the per-CTA partial results
  i64 partial_floats; // its capacity
};
// Several MODE_MN reductions of the same shape class in ONE launch (the weight gradients of all layers, deferred to
// the end of the backward pass): job q reduces over its own positions,  D_q[(blk, i), n] = sum_p A_q,blk[p][i] * B_q[p][n].
// The A blocks of a job are 32-column boxes of ONE 4-D tensor [nseg_src][nb][a_rows_src][32] (block j = segment seg[j],
// rows shifted by rshift[j]); the CTAs are divided among the jobs in proportion to their K extents.
constexpr int TR_MAXJOBS = 8;
struct TcRedJob {
  const float* a_src;
  int a_rows_src, nseg_src;
  i64 a_seg_stride;     // floats between consecutive segments of a_src
  const float* b_src;   // [nb][rows][b_width]
  int b_width;
  int nb, rows;         // samples, B rows per sample
  int rshift[8];        // per A block: source row = B row + rshift
};
struct TcRedJobsArgs {
  int njobs;
  TcRedJob job[TR_MAXJOBS];
  int na;               // real A blocks per job (<= 7; the all-ones block is appended)
  int seg[8];           // per A block: segment of a_src
  int N;                // B columns (32 or 64 ... <= 256, multiple of 32)
  int x3;
  float* partial;
  i64 partial_floats;
};
// What the launch produced: `nslots` partial results of `slot_floats` floats each.
//   MODE_MN: slot layout [mtiles*128 rows = (blk, i)][N]; MODE_K: [mtiles*128 rows = v][Ntile], slot s covers the
//   output columns (s % n_nt)*Ntile .. +Ntile.
struct TcRedResult {
  int nslots, mtiles, N, n_nt, n_mg;   // MODE_K: output tiled n_mg (row groups of mtiles*128) x n_nt (column tiles of N)
  i64 slot_floats;
  int job_cta0[TR_MAXJOBS + 1];        // multi-job launches: job q wrote the slots job_cta0[q] .. job_cta0[q+1]-1
};

#if !GWN_EMU
namespace tc {

struct TrMaps {
  CUtensorMap a[TR_MAXSRC];
  CUtensorMap b[TR_MAXSRC];
};
struct TrParams {
  int mode, na, nblk, mtiles, N, nbn, nb, chunks_per_sample, total_chunks, stages, a_bytes, b_bytes, tx_bytes, n_nt, n_mg;
  int acol0[TR_MAXSRC], arshift[TR_MAXSRC], bcol0[TR_MAXSRC];
  int pair_end[TR_MAXSRC];   // MODE_K: cumulative chunk (slab) counts per pair
  int ab;                    // MODE_MN: real A blocks per block group
  float* partial;
  i64 slot_floats;
  // multi-job MODE_MN (njobs > 1): CTA ranges, K extents and A-block coordinates per job
  int njobs;
  int job_cta0[TR_MAXJOBS + 1], job_chunks[TR_MAXJOBS], job_cps[TR_MAXJOBS];
  int jseg[8], jrshift[TR_MAXJOBS][8];
  // drain > 0 (only when mtiles * N <= 64): the issuer hands the accumulator to the epilogue warps every `drain`
  // chunks (two TMEM buffers, alternating); they add it to fp32 registers with round-to-nearest and write the slot at
  // the end.  Keeps every tensor-core accumulation chain short: the MMA's own accumulate step truncates, so a chain of
  // ~150 chunks cost 1.4e-4 of relative error on the first layer's filter gradient.
  int drain;
  int lo_stages;   // 3xTF32: buffers of the remainder ring (0 otherwise)
  int interleave;  // K chunks dealt round-robin to the CTAs of an output tile instead of in contiguous ranges
  int abox;        // multi-job: the job's A blocks are the consecutive segments 0..na-1, unshifted: ONE 4-D TMA box
  int ncat;        // 3xTF32, drain mode (N = 32 or 64): A.[B | B_lo] as ONE instruction of twice the width (the splitter puts the
                   // remainder blocks right behind B's) into accumulator columns [0,N) | [N,2N) that the drain adds; with the
                   // measured 66 + 0.75 N cycles per tcgen05.mma a k-step of N = 32 costs 114 + 90 cycles instead of 3 x 90
};

// X3 = 3xTF32 mode (fp32-grade): both operands are activations, so warps 2 and 3 split BOTH tiles of a stage into
// their remainders and the issuer runs A.B + A.B_lo + A_lo.B (see tcpos.cuh).  The remainder tiles live in their OWN
// ring of p.lo_stages (2) buffers [A_lo | B_lo], separate from the p.stages TMA-written stages [A | B]: only the raw
// stages hold bytes in flight, and with remainders inside every stage a 227 KB CTA had just two or three of them --
// 64 KB in flight per SM, ~3 TB/s on the whole GPU whatever the MMA cost.  A remainder buffer is busy from the split
// until its MMAs retire (commit -> loempty), a raw stage from the TMA until the same commit (-> empty).
template <bool X3, bool NCAT = false>   // NCAT: the N-concatenated remainder product (TrParams::ncat), its own instantiation
__global__ void __launch_bounds__(256, 1) tcred_kernel(const __grid_constant__ TrMaps maps, const TrParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int plane_bytes = p.a_bytes + p.b_bytes;
  // raw stage [A | B]; remainder buffer [A_lo | B_lo], same size.  ncat (see TrParams): the splitter writes B_lo right
  // behind B INSIDE the raw stage -- [A | B | B_lo], so that [B | B_lo] is one operand of twice the width -- and the
  // remainder buffers hold A_lo only (a raw stage is handed back by the same commit that frees the remainder buffer).
  constexpr bool kNcat = X3 && NCAT;
  const int stage_bytes = plane_bytes + (kNcat ? p.b_bytes : 0);
  const int lo_bytes = kNcat ? p.a_bytes : plane_bytes;
  const int LQ = X3 ? p.lo_stages : 0;
  const uint32_t st0 = base;
  const uint32_t lo0 = st0 + p.stages * stage_bytes;   // remainder ring
  uint8_t* lo_ptr = smem + (size_t)p.stages * stage_bytes;
  const uint32_t bar0 = lo0 + LQ * lo_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lo_ptr + (size_t)LQ * lo_bytes);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (p.stages + s); };
  const int nb2 = 2 * p.stages;
  auto split_bar = [&](int l) { return bar0 + 8u * (nb2 + l); };          // remainder buffer l filled
  auto loempty_bar = [&](int l) { return bar0 + 8u * (nb2 + 2 + l); };    // ... and consumed (LQ <= 2)
  const uint32_t done_bar = bar0 + 8u * (nb2 + 4);
  auto tfull_bar = [&](int a) { return bar0 + 8u * (nb2 + 5 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (nb2 + 7 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + nb2 + 9);

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  // output tile of this CTA: (k split, row/block group mg, column group ntile)
  const int n_ot = p.n_nt * p.n_mg;
  int ntile, mg, kslot, nk, job = 0, total_chunks = p.total_chunks, cps = p.chunks_per_sample;
  if (p.njobs > 1) {   // this CTA's job: one output tile per job, its CTAs split K
    while (job + 1 < p.njobs && (int)blockIdx.x >= p.job_cta0[job + 1]) ++job;
    ntile = 0; mg = 0;
    kslot = (int)blockIdx.x - p.job_cta0[job];
    nk = p.job_cta0[job + 1] - p.job_cta0[job];
    total_chunks = p.job_chunks[job];
    cps = p.job_cps[job];
  } else {
    const int ot = (int)(blockIdx.x % n_ot);
    ntile = ot % p.n_nt; mg = ot / p.n_nt;
    kslot = (int)(blockIdx.x / n_ot); nk = (int)(gridDim.x / n_ot);
  }
  const int na_loc = p.mode == 0 ? min(p.ab, p.na - mg * p.ab) : 0;   // MODE_MN: real A blocks of this block group
  if (p.mode == 0) {
    // A blocks that TMA never writes: block `na_loc` of every stage is the all-ones block (its accumulator rows become
    // the column sums of B = the bias gradient); blocks beyond it are zero.
    for (int s = 0; s < p.stages; ++s) {
      float* blk = reinterpret_cast<float*>(smem + (size_t)s * stage_bytes + (size_t)na_loc * 4096);
      const int nfill = (p.mtiles * 4 - na_loc) * 1024;
      for (int i = threadIdx.x; i < nfill; i += 256) blk[i] = i < 1024 ? 1.0f : 0.0f;
    }
    for (int l = 0; l < LQ; ++l) {   // their remainders are zero
      float* lo = reinterpret_cast<float*>(lo_ptr + (size_t)l * lo_bytes + (size_t)na_loc * 4096);
      const int nfill = (p.mtiles * 4 - na_loc) * 1024;
      for (int i = threadIdx.x; i < nfill; i += 256) lo[i] = 0.0f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
  }
  if (warp == 0 && lane == 0) {
    if (p.njobs > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[job]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[job]) : "memory");
    } else {
      for (int s = 0; s < (p.na < TR_MAXSRC ? p.na : TR_MAXSRC); ++s) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[s]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[0]) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int l = 0; l < 2; ++l) {
      mbar_init(split_bar(l), 192);
      mbar_init(loempty_bar(l), 1);
    }
    mbar_init(done_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
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

  // contiguous chunk range of this CTA (MODE_K: CTAs are split over n_nt column tiles of the output)
  // This CTA's K chunks: kslot, kslot + nk, kslot + 2 nk, ... -- interleaved with the other CTAs of the same output
  // tile, so that at any moment they read ADJACENT 4 KB pieces of every operand stream (contiguous ranges per CTA
  // made ~1000 far-apart DRAM streams: 2.9 TB/s with neither the MMAs nor the split nor the TMA issue rate binding).
  // c_lin = index in this CTA's sequence, c = c_beg + c_lin * c_step the global chunk.
  const int c_step = p.interleave ? nk : 1;
  const int c_beg = p.interleave ? kslot : (int)((long long)total_chunks * kslot / nk);
  const int n_my = p.interleave ? (kslot < total_chunks ? (total_chunks - kslot + nk - 1) / nk : 0)
                                : (int)((long long)total_chunks * (kslot + 1) / nk) - c_beg;
  const int c_end = c_beg + n_my * c_step;   // exclusive bound of the strided sequence

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp loops, one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    int pair = 0;
    for (int c = c_beg; c < c_end; c += c_step) {
      if (!mbar_wait_warp(empty_bar(stage), phase ^ 1u, 21)) break;
      const uint32_t sa = st0 + stage * stage_bytes, sb = sa + p.a_bytes;
      if (p.mode != 0)
        while (c >= p.pair_end[pair]) ++pair;
      if (elect_one()) {
      mbar_expect_tx(full_bar(stage), p.mode == 0 ? (uint32_t)((na_loc + p.nbn) * 4096) : (uint32_t)p.tx_bytes);
      if (p.mode == 0) {
        const int b = c / cps, r0 = (c - b * cps) * 32;
        if (p.njobs > 1) {
          if (p.abox) {   // all segments in one instruction (8 TMA issues per 32-position chunk kept the producer busy)
            tma_load_4d(sa, &maps.a[job], full_bar(stage), 0, r0, b, 0);
          } else {
            for (int j = 0; j < na_loc; ++j)
              tma_load_4d(sa + j * 4096, &maps.a[job], full_bar(stage), 0, r0 + p.jrshift[job][j], b, p.jseg[j]);
          }
          for (int j = 0; j < p.nbn; ++j) tma_load_3d(sb + j * 4096, &maps.b[job], full_bar(stage), 32 * j, r0, b);
        } else {
          for (int j = 0; j < na_loc; ++j) {
            const int jj = mg * p.ab + j;
            tma_load_3d(sa + j * 4096, &maps.a[jj], full_bar(stage), p.acol0[jj], r0 + p.arshift[jj], b);
          }
          for (int j = 0; j < p.nbn; ++j) tma_load_3d(sb + j * 4096, &maps.b[0], full_bar(stage), p.bcol0[0] + ntile * p.N + 32 * j, r0, b);
        }
      } else {
        const int slab = c - (pair ? p.pair_end[pair - 1] : 0);
        for (int t = 0; t < p.mtiles; ++t)
          tma_load_3d(sa + t * 16384, &maps.a[pair], full_bar(stage), 0, (mg * p.mtiles + t) * 128, slab);
        tma_load_3d(sb, &maps.b[pair], full_bar(stage), 0, ntile * p.N, slab);
      }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (whole warp loops, one elected lane issues)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (p.mode == 0 ? ((1u << 15) | (1u << 16)) : 0u) |
                           ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int stage = 0, lq = 0;
    uint32_t phase = 0, lphase = 0;
    bool first = true;
    int acc = 0, in_acc = 0;                 // drain mode: current TMEM buffer, chunks accumulated in it
    uint32_t accphase[2] = {0u, 0u};
    constexpr bool ncat = X3 && NCAT;
    const uint32_t accN = (uint32_t)(ncat ? 2 * p.N : p.N);   // accumulator columns per row tile
    const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * p.N) >> 3) << 17);
    const uint32_t acc_cols = (uint32_t)p.mtiles * accN;
    uint32_t tm0 = tmem_base;
    bool ok = true;
    for (int c = c_beg; c < c_end; c += c_step) {
      if (p.drain > 0 && in_acc == 0) {      // a fresh buffer: wait until the epilogue has drained its previous use
        if (!mbar_wait_warp(tempty_bar(acc), accphase[acc] ^ 1u, 26)) { ok = false; break; }
        tc_fence_after();
        tm0 = tmem_base + (uint32_t)acc * acc_cols;
        first = true;
      }
      if (!mbar_wait_warp(full_bar(stage), phase, 22)) { ok = false; break; }
      tc_fence_after();
      const uint32_t sa = st0 + stage * stage_bytes, sb = sa + p.a_bytes;
      auto descs = [&](int t, int kk, uint64_t& adesc, uint64_t& bdesc) {
        if (NCAT || p.mode == 0) {
          // MN-major, 32-byte-atom swizzle: atoms of 4 k-rows x 128 B (SBO 512 B), next 32-wide block 4096 B on (LBO)
          adesc = make_desc(sa + t * 16384 + kk * 1024, 4096, 512, 1);
          bdesc = make_desc(sb + kk * 1024, 4096, 512, 1);
        } else {
          adesc = make_desc(sa + t * 16384 + kk * 32, 16, 1024, 2);
          bdesc = make_desc(sb + kk * 32, 16, 1024, 2);
        }
      };
      if (!ncat) {
      if (elect_one()) {
#pragma unroll 1
      for (int t = 0; t < p.mtiles; ++t) {
        const uint32_t d_tmem = tm0 + (uint32_t)t * accN;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint64_t adesc, bdesc;
          descs(t, kk, adesc, bdesc);
          tc_mma_tf32(d_tmem, adesc, bdesc, idesc, (!first || kk > 0) ? 1u : 0u);
        }
      }
      }
      __syncwarp();
      }
      if (X3) {   // remainder terms after the split of this stage (it ran while the MMAs above were issued); the remainder
                  // planes sit plane_bytes further, same layout
        if (!mbar_wait_warp(split_bar(lq), lphase, 24)) { ok = false; break; }
        tc_fence_after();
        // remainder buffer lq relative to raw stage `stage` (descriptor start addresses are in 16-byte units)
        const uint64_t off = (uint64_t)(((lo0 + (uint32_t)lq * lo_bytes) - sa) >> 4);
        // (two copies of the loop: the issuing thread paces these kernels, a branch per MMA is measurable)
        if (ncat) {   // [B | B_lo]: the remainder blocks follow B's in the raw stage, same 4096-byte block stride
          if (elect_one()) {
#pragma unroll 1
          for (int t = 0; t < p.mtiles; ++t) {
            const uint32_t d_tmem = tm0 + (uint32_t)t * accN;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              uint64_t adesc, bdesc;
              descs(t, kk, adesc, bdesc);
              tc_mma_tf32(d_tmem, adesc, bdesc, idesc2, (!first || kk > 0) ? 1u : 0u);
              tc_mma_tf32(d_tmem, adesc + off, bdesc, idesc, 1u);
            }
          }
          }
        } else {
          if (elect_one()) {
#pragma unroll 1
          for (int t = 0; t < p.mtiles; ++t) {
            const uint32_t d_tmem = tm0 + (uint32_t)t * accN;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              uint64_t adesc, bdesc;
              descs(t, kk, adesc, bdesc);
              tc_mma_tf32(d_tmem, adesc, bdesc + off, idesc, 1u);
              tc_mma_tf32(d_tmem, adesc + off, bdesc, idesc, 1u);
            }
          }
          }
        }
        __syncwarp();
      }
      first = false;
      const bool hand_over = p.drain > 0 && (in_acc + 1 == p.drain || c + c_step >= c_end);
      if (elect_one()) {
        tc_commit(empty_bar(stage));
        if (X3) tc_commit(loempty_bar(lq));
        if (hand_over) tc_commit(tfull_bar(acc));   // hand this accumulator buffer to the epilogue warps
      }
      __syncwarp();
      if (X3) {
        if (++lq == LQ) { lq = 0; lphase ^= 1u; }
      }
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      if (p.drain > 0) {
        if (hand_over) {
          accphase[acc] ^= 1u;
          acc ^= 1;
          in_acc = 0;
        } else {
          ++in_acc;
        }
      }
    }
    if (p.drain == 0 && ok) {
      if (elect_one()) tc_commit(done_bar);
      __syncwarp();
    }
  }
  // ===================================================== splitter: remainders of the TMA-written A blocks and of B.
  // Warps 2-3 and the four epilogue warps (192 threads).  Two splitter warps alone bound these kernels: the support
  // gradient splits 58 KB per chunk (200 us -> 112 us with the split disabled, against 182 us with the MMAs disabled),
  // and in the drain mode of the weight gradients -- where the epilogue warps used to sit on `tfull` between their
  // periodic drains -- ncu had the two splitter warps 93 % busy, the producer blocked on `empty` 75 % and the tensor
  // pipe 12 % active (r02w).  The epilogue warps therefore split every chunk too and fit their drains in between.
  constexpr int nsplit = 192;
  const int t64 = threadIdx.x - 64;
  const int a_live = (p.mode == 0 ? na_loc * 4096 : p.a_bytes) / 16, b_live = p.b_bytes / 16;
  auto split_chunk = [&](int stage, int lq, uint32_t phase, uint32_t lphase) -> bool {
    if (!mbar_wait(loempty_bar(lq), lphase ^ 1u, 28)) return false;   // the MMAs that read this remainder buffer have retired
    if (!mbar_wait(full_bar(stage), phase, 25)) return false;
    uint8_t* sp = smem + (size_t)stage * stage_bytes;
    uint8_t* lp = lo_ptr + (size_t)lq * lo_bytes;
    const float4* a_src = reinterpret_cast<const float4*>(sp);
    float4* a_dst = reinterpret_cast<float4*>(lp);
#pragma unroll 4
    for (int i = t64; i < a_live; i += nsplit) {
      const float4 v = a_src[i];
      a_dst[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
    }
    const float4* b_src = reinterpret_cast<const float4*>(sp + p.a_bytes);
    float4* b_dst = reinterpret_cast<float4*>(kNcat ? sp + plane_bytes : lp + p.a_bytes);
#pragma unroll 4
    for (int i = t64; i < b_live; i += nsplit) {
      const float4 v = b_src[i];
      b_dst[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(split_bar(lq));
    return true;
  };
  if (X3 && (warp == 2 || warp == 3 || (warp >= 4 && p.drain == 0))) {
    int stage = 0, lq = 0;
    uint32_t phase = 0, lphase = 0;
    for (int c = c_beg; c < c_end; c += c_step) {
      if (!split_chunk(stage, lq, phase, lphase)) break;
      if (++lq == LQ) { lq = 0; lphase ^= 1u; }
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  }
  if (warp >= 4) {
    // ===================================================== epilogue: this CTA's partial result -> its private slot
    // (plain 128-bit stores; a small follow-up kernel sums the slots -- float atomics from 148 CTAs onto the same
    // few thousand addresses cost ~35-85 us per launch, ncu r01b)
    if (p.drain > 0) {
      const int ew = warp - 4;
      const int row = ew * 32 + lane;
      float* slot = p.partial + (size_t)blockIdx.x * p.slot_floats;
      const int ngrp = p.mtiles * p.N / 16;          // 16-column groups of the accumulator (<= 4), g -> (row tile, column)
      const int gpt = p.N / 16;
      float sum[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) sum[i] = 0.0f;
      const int ndrain = (n_my + p.drain - 1) / p.drain;
      bool ok = true;
      auto drain_one = [&](int d) -> bool {     // add accumulator buffer d & 1 (hand-over number d) to the registers
        const int buf = d & 1;
        if (!mbar_wait(tfull_bar(buf), (uint32_t)(d >> 1) & 1u, 27)) return false;
        tc_fence_after();
        if (kNcat) {   // N-concatenated product: the A.B_lo columns sit p.N columns behind the main ones
          const uint32_t taddr = tmem_base + ((uint32_t)(32 * ew) << 16) + (uint32_t)(buf * p.mtiles * 2 * p.N);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < ngrp) {
              uint32_t r[16], r2[16];
              const uint32_t col = (uint32_t)((g / gpt) * 2 * p.N + (g % gpt) * 16);
              tc_ld16(taddr + col, r);
              tc_ld16(taddr + col + (uint32_t)p.N, r2);
              tc_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j) sum[g * 16 + j] += __uint_as_float(r[j]) + __uint_as_float(r2[j]);
            }
          }
        } else {
          const uint32_t taddr = tmem_base + ((uint32_t)(32 * ew) << 16) + (uint32_t)(buf * p.mtiles * p.N);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < ngrp) {
              uint32_t r[16];
              tc_ld16(taddr + (uint32_t)((g / gpt) * p.N + (g % gpt) * 16), r);
              tc_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j) sum[g * 16 + j] += __uint_as_float(r[j]);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(buf));
        return true;
      };
      if (X3) {
        // split every chunk with warps 2-3; hand-over d (after chunk (d + 1) * drain - 1, or the last chunk) is drained
        // one chunk later, AFTER this warp's share of that chunk's split: the issuer needs the split of chunk i before it
        // can issue anything, while the MMAs of hand-over d were issued a chunk ago and complete on their own.
        int stage = 0, lq = 0, i = 0;
        uint32_t phase = 0, lphase = 0;
        for (int c = c_beg; c < c_end && ok; c += c_step, ++i) {
          if (!split_chunk(stage, lq, phase, lphase)) { ok = false; break; }
          if (++lq == LQ) { lq = 0; lphase ^= 1u; }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          if (i > 0 && i % p.drain == 0) ok = drain_one(i / p.drain - 1);
        }
        if (ok && ndrain > 0) ok = drain_one(ndrain - 1);
      } else {
        for (int d = 0; d < ndrain && ok; ++d) ok = drain_one(d);
      }
      if (ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g < ngrp) {
            float* orow = slot + (size_t)((g / gpt) * 128 + row) * p.N + (g % gpt) * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(orow + 4 * q) = make_float4(sum[g * 16 + 4 * q], sum[g * 16 + 4 * q + 1], sum[g * 16 + 4 * q + 2],
                                                                      sum[g * 16 + 4 * q + 3]);
          }
        }
      }
    } else if (c_end > c_beg && mbar_wait(done_bar, 0, 23)) {
      tc_fence_after();
      const int ew = warp - 4;
      const int row = ew * 32 + lane;
      float* slot = p.partial + (size_t)blockIdx.x * p.slot_floats;
      for (int t = 0; t < p.mtiles; ++t) {
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * ew) << 16) + (uint32_t)(t * p.N);
        float* orow = slot + (size_t)(t * 128 + row) * p.N;
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          uint32_t r[16];
          tc_ld16(taddr + c0, r);
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(orow + c0 + 4 * q) = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                                         __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------- slot reduction
// out element i = sum over the slots {slot0, slot0 + step, ...} of partial[slot][off(i)] (and optionally a second
// offset), handed to the functor's store().  Block = 32 outputs x 8 slot groups.
template <class F>
__global__ void __launch_bounds__(256) slot_reduce_kernel(const float* __restrict__ partial, int nslots, i64 slot_floats, i64 nout, F f) {
  __shared__ float sm[2][8][32];
  GWN_PDL_ENTRY();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const i64 i = (i64)blockIdx.x * 32 + tx;
  float s0 = 0.0f, s1 = 0.0f;
  if (i < nout) {
    i64 off, off2;
    int slot0, step, send = nslots;
    f.where(i, off, off2, slot0, step, send);
    for (int slot = slot0 + ty * step; slot < send; slot += 8 * step) {
      const float* q = partial + (size_t)slot * slot_floats;
      s0 += __ldg(q + off);
      if (off2 >= 0) s1 += __ldg(q + off2);
    }
  }
  sm[0][ty][tx] = s0;
  sm[1][ty][tx] = s1;
  __syncthreads();
  if (ty == 0 && i < nout) {
#pragma unroll
    for (int y = 1; y < 8; ++y) { s0 += sm[0][y][tx]; s1 += sm[1][y][tx]; }
    f.store(i, s0, s1);
  }
}

// Four consecutive outputs per thread (one 128-bit load per slot) for functors whose outputs i .. i+3 (i % 4 == 0) lie at
// consecutive slot offsets with the same slot set -- SlotGridOut (head weight gradients: 144 slots x 16-32 K floats
// took 13-20 us element-wise).  Block = 32 output quads x 8 slot groups.
template <class F>
__global__ void __launch_bounds__(256) slot_reduce4_kernel(const float* __restrict__ partial, int nslots, i64 slot_floats, i64 nout, F f) {
  __shared__ float4 sm[8][32];
  GWN_PDL_ENTRY();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const i64 i = ((i64)blockIdx.x * 32 + tx) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < nout) {
    i64 off, off2;
    int slot0, step, send = nslots;
    f.where(i, off, off2, slot0, step, send);
    for (int slot = slot0 + ty * step; slot < send; slot += 8 * step) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (size_t)slot * slot_floats + off));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < nout) {
#pragma unroll
    for (int y = 1; y < 8; ++y) { const float4 v = sm[y][tx]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
    f.store(i, s.x, 0.f);
    f.store(i + 1, s.y, 0.f);
    f.store(i + 2, s.z, 0.f);
    f.store(i + 3, s.w, 0.f);
  }
}

// gcn mlp weight / bias gradient: slot rows (blk*32 + ci) x 32 columns (n = co); the ones block is row nblk*32.
struct SlotMlpOut {
  float* dW; float* db; int ldw, nblk;
  __device__ __forceinline__ void where(i64 i, i64& off, i64& off2, int& slot0, int& step, int&) const {
    const i64 nw = (i64)nblk * 32 * 32;
    off = i < nw ? i : (i64)nblk * 32 * 32 + (i - nw);
    off2 = -1; slot0 = 0; step = 1;
  }
  __device__ __forceinline__ void store(i64 i, float s0, float) const {
    const i64 nw = (i64)nblk * 32 * 32;
    if (i < nw) {
      const int row = (int)(i >> 5), n = (int)(i & 31);
      dW[(size_t)n * ldw + row] += s0;
    } else {
      db[i - nw] += s0;
    }
  }
};
// gated-conv weight / bias gradient with the BatchNorm affine of the layer below folded in (x = a*u + c):
// slot rows (tap*32 + ci) x 64 columns j = 2*ch + gate; ones block = row 64.   dW = a[ci]*R + c[ci]*S[j], db = S[j].
struct SlotTcnOut {
  const float* ac; float* dwf; float* dwg; float* dbf; float* dbg; int C, D;
  __device__ __forceinline__ void where(i64 i, i64& off, i64& off2, int& slot0, int& step, int&) const {
    const i64 nw = (i64)2 * C * 2 * D;
    if (i < nw) { off = i; off2 = ac ? (i64)2 * C * 2 * D + (i % (2 * D)) : -1; }
    else { off = (i64)2 * C * 2 * D + (i - nw); off2 = -1; }
    slot0 = 0; step = 1;
  }
  __device__ __forceinline__ void store(i64 i, float s0, float s1) const {
    const i64 nw = (i64)2 * C * 2 * D;
    if (i < nw) {
      const int k = (int)(i / (2 * D)), j = (int)(i - (i64)k * 2 * D);
      const int tap = k / C, ci = k - tap * C;
      float v = s0;
      if (ac) v = ac[ci] * s0 + ac[C + ci] * s1;
      ((j & 1) ? dwg : dwf)[((size_t)(j >> 1) * C + ci) * 2 + tap] += v;
    } else {
      const int j = (int)(i - nw);
      ((j & 1) ? dbg : dbf)[j >> 1] += s0;
    }
  }
};
// Tiled MODE_MN result (head weight gradients): element (j, i, n) of global A block j -> blk[j][n*sn + i*si]; the ones
// block of block group 0 delivers the column sums of B to every bias[q][n].  Slot = (k split, block group, column group).
struct SlotGridOut {
  float* blk[TR_MAXSRC];
  float* bias[TR_MAXSRC];
  int nbias, na, ab, n_ag, n_bg, Ntile, Ntot, n_valid;   // n_valid: B columns that exist (the rest were zero-filled)
  i64 sn, si;
  __device__ __forceinline__ void where(i64 i, i64& off, i64& off2, int& slot0, int& step, int&) const {
    const i64 nw = (i64)na * 32 * Ntot;
    int lb, ii, n, ga;
    if (i < nw) {
      const int j = (int)(i / (32 * Ntot)), r = (int)(i - (i64)j * 32 * Ntot);
      ii = r / Ntot; n = r - ii * Ntot; ga = j / ab; lb = j - ga * ab;
    } else {
      n = (int)(i - nw); ii = 0; ga = 0; lb = ab < na ? ab : na;   // the ones block follows the real blocks of group 0
    }
    const int gb = n / Ntile, ln = n - gb * Ntile;
    off = (i64)(lb * 32 + ii) * Ntile + ln;
    off2 = -1; slot0 = ga * n_bg + gb; step = n_ag * n_bg;
  }
  __device__ __forceinline__ void store(i64 i, float s0, float) const {
    const i64 nw = (i64)na * 32 * Ntot;
    if (i < nw) {
      const int j = (int)(i / (32 * Ntot)), r = (int)(i - (i64)j * 32 * Ntot);
      const int ii = r / Ntot, n = r - ii * Ntot;
      if (n < n_valid) blk[j][(size_t)n * sn + (size_t)ii * si] += s0;
    } else {
      const int n = (int)(i - nw);
      if (n < n_valid)
        for (int q = 0; q < nbias; ++q) bias[q][n] += s0;
    }
  }
};
// support gradient: a slot holds output rows (mg*MR .. +MR) x columns (nt*Ntile .. +Ntile); slot index =
// (k split * n_mg + mg) * n_nt + nt
struct SlotSupOut {
  float* dA; int ld, V, Ntile, n_nt, MR, n_mg;
  __device__ __forceinline__ void where(i64 i, i64& off, i64& off2, int& slot0, int& step, int&) const {
    const int v = (int)(i / V), w = (int)(i - (i64)v * V);
    const int nt = w / Ntile, mg = v / MR;
    off = (i64)(v - mg * MR) * Ntile + (w - nt * Ntile);
    off2 = -1; slot0 = mg * n_nt + nt; step = n_nt * n_mg;
  }
  __device__ __forceinline__ void store(i64 i, float s0, float) const {
    const int v = (int)(i / V), w = (int)(i - (i64)v * V);
    dA[(size_t)v * ld + w] += s0;
  }
};

// The same functor per job of a multi-job launch: output i = (job, element); the job's slots are its CTA range.
template <class F>
struct SlotJobs {
  F f[TR_MAXJOBS];
  int cta0[TR_MAXJOBS + 1];
  i64 nout_job;
  __device__ __forceinline__ void where(i64 i, i64& off, i64& off2, int& slot0, int& step, int& send) const {
    const int q = (int)(i / nout_job);
    int dummy = 0;
    f[q].where(i - (i64)q * nout_job, off, off2, slot0, step, dummy);
    slot0 += cta0[q];
    send = cta0[q + 1];
  }
  __device__ __forceinline__ void store(i64 i, float s0, float s1) const {
    const int q = (int)(i / nout_job);
    f[q].store(i - (i64)q * nout_job, s0, s1);
  }
};

}  // namespace tc

// SlotGridOut with Ntot, Ntile multiples of 4 (and 16-byte aligned slots): the 4-wide kernel
inline int launch_slot_reduce4(const float* partial, const TcRedResult& r, i64 nout, const tc::SlotGridOut& f, cudaStream_t stream) {
  if (nout <= 0) return 0;
  GWN_CUDA(launch_kernel(tc::slot_reduce4_kernel<tc::SlotGridOut>, dim3((unsigned)((nout / 4 + 31) / 32)), dim3(256), 0, stream, partial,
                         r.nslots, r.slot_floats, nout, f));
  count_launch();
  return 0;
}

template <class F>
inline int launch_slot_reduce(const float* partial, const TcRedResult& r, i64 nout, const F& f, cudaStream_t stream) {
  if (nout <= 0) return 0;
  GWN_CUDA(launch_kernel(tc::slot_reduce_kernel<F>, dim3((unsigned)((nout + 31) / 32)), dim3(256), 0, stream, partial, r.nslots,
                         r.slot_floats, nout, f));
  count_launch();
  return 0;
}
#endif

// GWNET_B200_TCRED_NCAT=0 turns the N-concatenated remainder product (TrParams::ncat) off for A/B runs.
inline int tcred_ncat_enabled() {
  static const int on = [] {
    const char* e = getenv("GWNET_B200_TCRED_NCAT");
    return e ? atoi(e) : 1;
  }();
  return on;
}
// 0 = launched, -1 = not eligible (caller falls back), > 0 = error.
inline int launch_tcred(const TcRedArgs& a, cudaStream_t stream, TcRedResult* res) {
#if GWN_EMU
  (void)a; (void)stream; (void)res;
  return -1;
#else
  using namespace tc;
  TrMaps maps;
  TrParams p;
  memset(&p, 0, sizeof(p));
  p.mode = a.mode; p.na = a.na; p.nb = a.nb; p.n_nt = 1; p.n_mg = 1;
  if (a.na < 1 || a.na > TR_MAXSRC || a.rows < 1 || !a.partial || !res) return -1;
  static int num_sms = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  if (a.mode == 0) {
    const int ab = a.ab > 0 ? a.ab : a.na;
    const int Ntot = a.N_total > 0 ? a.N_total : a.N;
    if (ab > 7 || ab < 1 || a.N % 32 != 0 || a.N < 32 || a.N > 256 || Ntot % a.N != 0 || a.nb < 1) return -1;
    p.N = a.N;
    p.ab = ab;
    p.n_mg = (a.na + ab - 1) / ab;             // block groups
    p.n_nt = Ntot / a.N;                       // column groups
    p.nblk = ab + 1;                           // + the all-ones block
    p.mtiles = (p.nblk + 3) / 4;
    if (p.mtiles * a.N > 512) return -1;       // TMEM columns
    p.nbn = a.N / 32;
    p.a_bytes = p.mtiles * 16384;
    p.b_bytes = p.nbn * 4096;
    p.tx_bytes = (ab + p.nbn) * 4096;
    p.chunks_per_sample = (a.rows + 31) / 32;
    const long long tot = (long long)p.chunks_per_sample * a.nb;
    if (tot > 2147483647LL) return -1;
    p.total_chunks = (int)tot;
    for (int j = 0; j < a.na; ++j) {
      const TcRedSrc& g = a.a[j];
      if ((reinterpret_cast<uintptr_t>(g.src) & 15) || g.row_width % 4 || g.col0 + 32 > g.row_width) return -1;
      cuuint64_t d[3] = {(cuuint64_t)g.row_width, (cuuint64_t)g.rows_src, (cuuint64_t)a.nb};
      cuuint64_t st[2] = {(cuuint64_t)g.row_width * 4, (cuuint64_t)g.rows_src * g.row_width * 4};
      cuuint32_t box[3] = {32, 32, 1};
      GWN_TRY(encode(&maps.a[j], g.src, 3, d, st, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
      p.acol0[j] = g.col0;
      p.arshift[j] = g.rshift;
    }
    {
      const TcRedSrc& g = a.b[0];
      if ((reinterpret_cast<uintptr_t>(g.src) & 15) || g.row_width % 4 || g.col0 >= g.row_width) return -1;   // short rows read zeros
      cuuint64_t d[3] = {(cuuint64_t)g.row_width, (cuuint64_t)g.rows_src, (cuuint64_t)a.nb};
      cuuint64_t st[2] = {(cuuint64_t)g.row_width * 4, (cuuint64_t)g.rows_src * g.row_width * 4};
      cuuint32_t box[3] = {32, 32, 1};
      GWN_TRY(encode(&maps.b[0], g.src, 3, d, st, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
      for (int j = 0; j < p.nbn; ++j) p.bcol0[j] = g.col0 + 32 * j;
    }
    for (int j = a.na; j < TR_MAXSRC; ++j) maps.a[j] = maps.a[0];
    for (int j = 1; j < TR_MAXSRC; ++j) maps.b[j] = maps.b[0];
  } else {
    // D[v, w] over V x V: mtiles row tiles of 128 live in TMEM side by side, Ntile columns each (mtiles*Ntile <= 512);
    // n_nt column tiles are spread over the CTAs
    p.nblk = 0;
    // Tiling of the V x V output over CTAs: mt row tiles of 128 side by side in TMEM per CTA (n_mg row groups), n_nt
    // column tiles of `ntile`, the rest of the SMs split K.  One tcgen05.mma of 128 x N x 8 (tf32, both operands
    // from shared memory) costs about 66 + 0.75 N cycles on B200 (measured on the node contraction: N = 112 -> 150,
    // 208 -> 222, 256 -> 258), i.e. wide instructions are cheaper per column, so the modelled time of a K chunk is
    // mt x (66 + 0.75 ntile) and a candidate's cost that time divided by its K-split factor.  METR-LA (V = 207):
    // one row tile x 208 columns per CTA, 2 row groups x 74 K splits (was 2 x 112 columns x 74: 26 % more MMA time).
    const int rowtiles = (a.rows + 127) / 128;
    int best_mt = 0, best_nt = 0;
    double best_cost = 0.0;
    for (int mt = 1; mt <= 4 && mt <= rowtiles; ++mt) {
      const int n_mg = (rowtiles + mt - 1) / mt;
      for (int cnt = 1; cnt <= 64; ++cnt) {
        const int nt = round_up((a.rows + cnt - 1) / cnt, 16);
        if (nt > 256 || mt * nt > 512) continue;
        const int n_nt = (a.rows + nt - 1) / nt;
        if ((SMEM_LIMIT - 2048) / ((a.x3 ? 2 : 1) * (mt * 16384 + nt * 128)) < 2) continue;   // two pipeline stages
        int nk = num_sms / (n_nt * n_mg);
        if (nk < 1) nk = 1;
        const double rounds = (double)((n_nt * n_mg + num_sms - 1) / num_sms);   // > 1 only when there are more tiles than SMs
        const double cost = rounds * mt * (66.0 + 0.75 * nt) / nk;
        if (best_mt == 0 || cost < best_cost * 0.999 || (cost < best_cost * 1.001 && mt > best_mt)) {
          best_mt = mt; best_nt = nt; best_cost = cost;
        }
        if (nt <= 16) break;
      }
    }
    if (best_mt == 0) return -1;
    p.mtiles = best_mt;
    p.n_mg = (rowtiles + best_mt - 1) / best_mt;
    const int ntile = best_nt;
    p.N = ntile;
    p.n_nt = (a.rows + ntile - 1) / ntile;
    p.a_bytes = p.mtiles * 16384;
    p.b_bytes = ntile * 128;
    p.tx_bytes = p.a_bytes + p.b_bytes;
    long long tot = 0;
    for (int j = 0; j < a.na; ++j) {
      const int nbj = a.a[j].nb > 0 ? a.a[j].nb : a.nb;
      if (nbj < 1) return -1;
      tot += nbj;
      if (tot > 2147483647LL) return -1;
      p.pair_end[j] = (int)tot;
      const TcRedSrc* gs[2] = {&a.a[j], &a.b[j]};
      for (int w = 0; w < 2; ++w) {
        const TcRedSrc& g = *gs[w];
        if ((reinterpret_cast<uintptr_t>(g.src) & 15) || g.row_width != 32) return -1;
        cuuint64_t d[3] = {32, (cuuint64_t)a.rows, (cuuint64_t)nbj};
        cuuint64_t st[2] = {128, (cuuint64_t)a.rows * 128};
        cuuint32_t box[3] = {32, (cuuint32_t)(w == 0 ? 128 : ntile), 1};
        GWN_TRY(encode(w == 0 ? &maps.a[j] : &maps.b[j], g.src, 3, d, st, box, CU_TENSOR_MAP_SWIZZLE_128B));
      }
    }
    for (int j = a.na; j < TR_MAXSRC; ++j) { p.pair_end[j] = (int)tot; maps.a[j] = maps.a[0]; maps.b[j] = maps.b[0]; }
    p.total_chunks = (int)tot;
  }
  {
    static const int inter = [] {
      const char* e = getenv("GWNET_B200_TCRED_INTERLEAVE");
      return e ? atoi(e) : 1;
    }();
    p.interleave = inter;
  }
  if (a.mode == 0 && p.mtiles * p.N <= 64) p.drain = 16;
  p.ncat = (a.x3 && p.drain > 0 && tcred_ncat_enabled()) ? 1 : 0;
  // raw stage [A | B] (+ B_lo with ncat); the 3xTF32 remainder ring has 2 buffers [A_lo | B_lo] (A_lo only with ncat)
  const int stage_bytes = p.a_bytes + p.b_bytes * (p.ncat ? 2 : 1);
  const int lo_bytes = p.ncat ? p.a_bytes : p.a_bytes + p.b_bytes;
  p.lo_stages = a.x3 ? 2 : 0;
  p.stages = (SMEM_LIMIT - 2048 - p.lo_stages * lo_bytes) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) return -1;
  // grid: a multiple of the output tile count, every CTA gets at least one chunk
  const int n_ot = p.n_nt * p.n_mg;
  int nk = num_sms / n_ot;
  if (nk < 1) nk = 1;
  if (nk > p.total_chunks) nk = p.total_chunks;
  if (nk < 1) return -1;
  const int grid = nk * n_ot;
  p.slot_floats = (i64)p.mtiles * 128 * p.N;
  if ((i64)grid * p.slot_floats > a.partial_floats) return -1;
  p.partial = a.partial;
  const int smem_bytes = p.stages * stage_bytes + p.lo_stages * lo_bytes + 1024 + 256;
  static cudaError_t attr = cudaFuncSetAttribute(tcred_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  static cudaError_t attr3 = cudaFuncSetAttribute(tcred_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  static cudaError_t attr3c = cudaFuncSetAttribute(tcred_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  if (attr != cudaSuccess || attr3 != cudaSuccess || attr3c != cudaSuccess) {
    set_error("tcred: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr != cudaSuccess ? attr : attr3 != cudaSuccess ? attr3 : attr3c));
    return GWN_ERR_CUDA;
  }
  if (a.x3 && p.ncat) GWN_CUDA(launch_kernel(tcred_kernel<true, true>, dim3(grid), dim3(256), smem_bytes, stream, maps, p));
  else if (a.x3) GWN_CUDA(launch_kernel(tcred_kernel<true>, dim3(grid), dim3(256), smem_bytes, stream, maps, p));
  else GWN_CUDA(launch_kernel(tcred_kernel<false>, dim3(grid), dim3(256), smem_bytes, stream, maps, p));
  GWN_LAUNCH_CHECK();
  count_launch();
  res->nslots = grid; res->mtiles = p.mtiles; res->N = p.N; res->n_nt = p.n_nt; res->n_mg = p.n_mg; res->slot_floats = p.slot_floats;
  return 0;
#endif
}

// Several same-shaped MODE_MN reductions in one launch (see TcRedJobsArgs).  0 = launched, -1 = not eligible, > 0 = error.
inline int launch_tcred_jobs(const TcRedJobsArgs& a, cudaStream_t stream, TcRedResult* res) {
#if GWN_EMU
  (void)a; (void)stream; (void)res;
  return -1;
#else
  using namespace tc;
  if (a.njobs < 2 || a.njobs > TR_MAXJOBS || a.na < 1 || a.na > 7 || a.N % 32 != 0 || a.N < 32 || a.N > 256 || !a.partial || !res)
    return -1;
  static int num_sms = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  if (a.njobs > num_sms) return -1;
  TrMaps maps;
  TrParams p;
  memset(&p, 0, sizeof(p));
  p.mode = 0; p.na = a.na; p.n_nt = 1; p.n_mg = 1; p.N = a.N; p.ab = a.na;
  p.nblk = a.na + 1;
  p.mtiles = (p.nblk + 3) / 4;
  if (p.mtiles * a.N > 512) return -1;
  p.nbn = a.N / 32;
  p.a_bytes = p.mtiles * 16384;
  p.b_bytes = p.nbn * 4096;
  p.tx_bytes = (a.na + p.nbn) * 4096;
  p.njobs = a.njobs;
  for (int j = 0; j < a.na; ++j) p.jseg[j] = a.seg[j];
  p.abox = 1;
  for (int j = 0; j < a.na; ++j)
    if (a.seg[j] != j) p.abox = 0;
  for (int q = 0; q < a.njobs; ++q) {
    if (a.job[q].nseg_src < a.na) p.abox = 0;
    for (int j = 0; j < a.na; ++j)
      if (a.job[q].rshift[j] != 0) p.abox = 0;
  }
  long long tot = 0;
  for (int q = 0; q < a.njobs; ++q) {
    const TcRedJob& g = a.job[q];
    if (g.nb < 1 || g.rows < 1 || g.nseg_src < 1 || (reinterpret_cast<uintptr_t>(g.a_src) & 15) || (reinterpret_cast<uintptr_t>(g.b_src) & 15) ||
        g.b_width % 4 != 0 || g.b_width < a.N || g.a_seg_stride % 4 != 0)
      return -1;
    for (int j = 0; j < a.na; ++j) {
      if (a.seg[j] < 0 || a.seg[j] >= g.nseg_src) return -1;
      p.jrshift[q][j] = g.rshift[j];
    }
    p.job_cps[q] = (g.rows + 31) / 32;
    const long long ch = (long long)p.job_cps[q] * g.nb;
    if (ch > 2147483647LL) return -1;
    p.job_chunks[q] = (int)ch;
    tot += ch;
    {
      const i64 segs = g.nseg_src > 1 ? g.a_seg_stride : (i64)g.a_rows_src * g.nb * 32;
      cuuint64_t d[4] = {32, (cuuint64_t)g.a_rows_src, (cuuint64_t)g.nb, (cuuint64_t)g.nseg_src};
      cuuint64_t st[3] = {128, (cuuint64_t)g.a_rows_src * 128, (cuuint64_t)segs * 4};
      cuuint32_t box[4] = {32, 32, 1, (cuuint32_t)(p.abox ? a.na : 1)};
      GWN_TRY(encode(&maps.a[q], g.a_src, 4, d, st, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    }
    {
      cuuint64_t d[3] = {(cuuint64_t)g.b_width, (cuuint64_t)g.rows, (cuuint64_t)g.nb};
      cuuint64_t st[2] = {(cuuint64_t)g.b_width * 4, (cuuint64_t)g.rows * g.b_width * 4};
      cuuint32_t box[3] = {32, 32, 1};
      GWN_TRY(encode(&maps.b[q], g.b_src, 3, d, st, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    }
  }
  for (int q = a.njobs; q < TR_MAXSRC; ++q) { maps.a[q] = maps.a[0]; maps.b[q] = maps.b[0]; }
  // CTAs per job in proportion to its K extent (at least one, at most one per chunk); leftovers to the most loaded jobs.
  // A CTA accumulates its whole K range in ONE fp32 TMEM accumulator, and the tensor core's accumulate step truncates:
  // the error grows linearly with the chain (measured: 147 chunks per CTA -> 1.4e-4 on the first layer's filter
  // gradient, 34 chunks -> 3e-5).  Small accumulators (<= 64 columns) are drained into registers every 16 chunks
  // (TrParams::drain); otherwise the grid grows to whole extra rounds of CTAs until a range is <= 48 chunks.
  const int sms = num_sms;
  int T = sms;
  if (p.mtiles * p.N <= 64) {
    p.drain = 16;   // short chains by draining into registers: one round of CTAs is enough
    p.ncat = (a.x3 && tcred_ncat_enabled()) ? 1 : 0;
  } else {
    long long want = (tot + 47) / 48;
    want = (want + sms - 1) / sms * sms;
    const long long cap = a.partial_floats / ((i64)p.mtiles * 128 * p.N);
    if (want > cap) want = cap / sms * sms;
    if (want > sms) T = (int)want;
  }
  int nkq[TR_MAXJOBS], used = 0;
  for (int q = 0; q < a.njobs; ++q) {
    long long n = (long long)T * p.job_chunks[q] / tot;
    if (n < 1) n = 1;
    if (n > p.job_chunks[q]) n = p.job_chunks[q];
    nkq[q] = (int)n;
    used += nkq[q];
  }
  while (used > T) {   // the "at least one" floor overshot: take from the least loaded job that has more than one
    int w = -1;
    for (int q = 0; q < a.njobs; ++q)
      if (nkq[q] > 1 && (w < 0 || (double)p.job_chunks[q] / nkq[q] < (double)p.job_chunks[w] / nkq[w])) w = q;
    if (w < 0) return -1;
    --nkq[w]; --used;
  }
  while (used < T) {
    int w = -1;
    for (int q = 0; q < a.njobs; ++q)
      if (nkq[q] < p.job_chunks[q] && (w < 0 || (double)p.job_chunks[q] / nkq[q] > (double)p.job_chunks[w] / nkq[w])) w = q;
    if (w < 0) break;
    ++nkq[w]; ++used;
  }
  p.job_cta0[0] = 0;
  for (int q = 0; q < a.njobs; ++q) p.job_cta0[q + 1] = p.job_cta0[q] + nkq[q];
  const int grid = p.job_cta0[a.njobs];
  {
    static const int inter = [] {
      const char* e = getenv("GWNET_B200_TCRED_INTERLEAVE");
      return e ? atoi(e) : 1;
    }();
    p.interleave = inter;
  }
  // raw stage [A | B] (+ B_lo with ncat); the 3xTF32 remainder ring has 2 buffers [A_lo | B_lo] (A_lo only with ncat)
  const int stage_bytes = p.a_bytes + p.b_bytes * (p.ncat ? 2 : 1);
  const int lo_bytes = p.ncat ? p.a_bytes : p.a_bytes + p.b_bytes;
  p.lo_stages = a.x3 ? 2 : 0;
  p.stages = (SMEM_LIMIT - 2048 - p.lo_stages * lo_bytes) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) return -1;
  p.slot_floats = (i64)p.mtiles * 128 * p.N;
  if ((i64)grid * p.slot_floats > a.partial_floats) return -1;
  p.partial = a.partial;
  const int smem_bytes = p.stages * stage_bytes + p.lo_stages * lo_bytes + 1024 + 256;
  static cudaError_t attr = cudaFuncSetAttribute(tcred_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  static cudaError_t attr3 = cudaFuncSetAttribute(tcred_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  static cudaError_t attr3c = cudaFuncSetAttribute(tcred_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  if (attr != cudaSuccess || attr3 != cudaSuccess || attr3c != cudaSuccess) {
    set_error("tcred: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr != cudaSuccess ? attr : attr3 != cudaSuccess ? attr3 : attr3c));
    return GWN_ERR_CUDA;
  }
  if (a.x3 && p.ncat) GWN_CUDA(launch_kernel(tcred_kernel<true, true>, dim3(grid), dim3(256), smem_bytes, stream, maps, p));
  else if (a.x3) GWN_CUDA(launch_kernel(tcred_kernel<true>, dim3(grid), dim3(256), smem_bytes, stream, maps, p));
  else GWN_CUDA(launch_kernel(tcred_kernel<false>, dim3(grid), dim3(256), smem_bytes, stream, maps, p));
  count_launch();
  res->nslots = grid; res->mtiles = p.mtiles; res->N = p.N; res->n_nt = 1; res->n_mg = 1; res->slot_floats = p.slot_floats;
  for (int q = 0; q <= a.njobs; ++q) res->job_cta0[q] = p.job_cta0[q];
  return 0;
#endif
}

}  // namespace gwn
