// Tensor-core node contraction (nconv, model.py:13 and its autograd) for sm_100a:
// tcgen05.mma kind::tf32 with the accumulator in TMEM, operands staged by TMA (128-byte swizzle),
// warp-specialised (TMA producer / MMA issuer / 4 epilogue warps), persistent over tiles.
//
//   D[j, m] = sum_s sum_k X_s[k, j] * S_s[m, k]          j = (slab, channel) row, m = output node
//
// GEMM orientation: M = 128 rows of j (4 slabs x 32 channels: each slab's [node][32 ch] block is
// exactly one MN-major SWIZZLE_128B atom column), N = output nodes (<= 256 per tile), K = nodes.
// The support is read from a K-contiguous ("k-major") padded buffer S[m][k] (A^T for the forward
// contraction, A itself for dX = A.dY).
#pragma once
#include "common.cuh"

namespace gwn {

constexpr int TC_MAXSUP = 4;

struct NodeTcArgs {
  const float* X[TC_MAXSUP];       // slab tensors [nslabs][V][32]
  const float* S[TC_MAXSUP];       // k-major supports [V][ld]
  const float* Slo[TC_MAXSUP];     // 3xTF32 mode (all non-null): their remainders S - tf32_trunc(S), same layout
  int ld;
  int nsup;
  int kcat;                        // 1: one output, summed over supports; 0: nsup independent outputs
  float* Y[TC_MAXSUP];
  const float* add[TC_MAXSUP];     // nullable, same layout as Y
  const float* add2;               // nullable head window [B][T_out][V][32]
  int B, L, T_out, V;
  int per_sample;                  // 1: one support set per sample (gwnet_diff_G): S[s] points at sample 0's matrix,
  long long s_batch_stride;        //    consecutive samples s_batch_stride floats apart; tiles never straddle samples
};

// Returns GWN_ERR_UNSUPPORTED (with a message) when the shape cannot use this kernel.
int node_gemm_tc(const NodeTcArgs& a, cudaStream_t stream);
// Both hops of an order-2 gcn for all supports in ONE launch (gcn_hops_fused.cuh: support resident in shared memory, hop 2
// fed from the hop-1 accumulator in tensor memory; single-pass TF32 tier, V <= 256).  S[s]: K-contiguous supports [V][ld];
// Y1[s] / Y2[s]: the hop-1 / hop-2 tensors.  0 = launched, -1 = not eligible (run two node_gemm_tc launches), > 0 = error.
int gcn_hops_fused_tc(const float* x, const float* const* S, int nsup, int ld, float* const* Y1, float* const* Y2, int B, int L, int V,
                      cudaStream_t stream);
int tc_error_flag(int reset);
void tc_set_debug_buffer(float* p);
void tc_set_debug_mode(int m);

}  // namespace gwn
