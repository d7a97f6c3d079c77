// Position GEMM for the tensor-core tiers:   out[p, n] = epilogue( sum_k A(p, k) * W(k, n) )
//   p : position rows of BLNC activations (M = B*L*N, huge);  K = nseg * 32 <= 256;  N <= 256.
// This is the shape of every convolution on the hot path when residual = dilation channels = 32 (the
// reference default, engine.py:27-31): the gated (1,2) conv and its gate-backward recompute (K = 2 taps x 32),
// the gcn mlp (K = 7 hops x 32), the mlp / gated-conv input gradients.
//
// Design (ncu on the generic gemm.cuh kernel showed these GEMMs were bound by index arithmetic and shared-
// memory staging, not by math or HBM): NO shared-memory staging of A.  Each warp owns 32 rows; the m16n8k8
// A fragments are loaded straight from global memory (each quad of lanes reads one 16-byte piece of a row,
// a whole 32-float segment = 32 independent loads in flight per thread), all row arithmetic is done once
// per thread before the K loop, and only the small weight matrix W (<= 36 KB) is staged in shared memory
// once per block.  Epilogues are the functors of functors.cuh (same store4 / finish contract).
#pragma once
#include "functors.cuh"

namespace gwn {

// dst_row(p) = b*lon_dst + r + off  with  b = p / lon_src, r = p % lon_src;  valid iff 0 <= r + off < hi.
struct RowMap {
  int lon_src, lon_dst, off, hi;
};
inline RowMap rowmap_identity() { return RowMap{1 << 30, 1 << 30, 0, 1 << 30}; }
inline RowMap rowmap_shift(int L_src, int L_dst, int off_t, int N) {   // (b,t,n) -> (b,t+off_t,n), valid inside [0, L_dst)
  return RowMap{L_src * N, L_dst * N, off_t * N, L_dst * N};
}

constexpr int PG_WD = 32;       // segment width (channels)
constexpr int PG_MAXSEG = 16;

struct ARows {
  const float* P[PG_MAXSEG];        // per K segment: base pointer (column offset already applied)
  unsigned char rmap[PG_MAXSEG];    // per K segment: which row map
  RowMap rm[2];
  int nseg;
  int rs;                           // row stride of the source tensors (floats)
  const float* ac;                  // nullable BatchNorm fold a[32], c[32] (applied to valid rows only)
};

#if !GWN_EMU
template <class T, class WL, class EP, int MATH, int NC>
__global__ void __launch_bounds__(256, 1) posgemm_kernel(const ARows A, WL wl, EP ep, i64 M, int N) {
  static_assert(T::NT == 256 && T::WM == 8 && T::WN == 1, "posgemm: epilogue tile must be 8 warps x 32 rows");
  static_assert(NC % 8 == 0 && NC <= 64, "posgemm: NC");
  constexpr int NTL = NC / 8;
  extern __shared__ __align__(16) float smem[];
  const int K = A.nseg * PG_WD;
  const int WS = N + 8;                       // row stride == 8 (mod 32) for conflict-free B fragments when N % 32 == 0
  float* Ws = smem;                           // [K][WS]
  float* red = smem + (size_t)K * WS;         // 2 * 32 floats for the column-statistics reduce
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  wl.init(0);
  ep.init(0);

  // ---- stage W once.  Rows are stored K-permuted inside every group of 16 so that a thread's 128-bit A load
  // (4 consecutive k of one row) supplies the "k = t" and "k = t+4" fragment elements of two MMA steps:
  // physical k = 4t + 2e + u  ->  row  e*8 + u*4 + t  of its group  (e = MMA step, u = fragment half).
  auto srow = [](int k) { return (k & ~15) | (((k >> 1) & 1) << 3) | ((k & 1) << 2) | ((k >> 2) & 3); };
  if (WL::kInner) {
    for (int i = tid; i < (K / 4) * N; i += 256) {
      const int n = i / (K / 4), k4 = (i - n * (K / 4)) * 4;
      float v[4];
      wl.load4(v, k4, (i64)n, K, (i64)N);
#pragma unroll
      for (int q = 0; q < 4; ++q) Ws[srow(k4 + q) * WS + n] = v[q];
    }
  } else {
    const int n4s = (N + 3) / 4;
    for (int i = tid; i < K * n4s; i += 256) {
      const int k = i / n4s, n = (i - k * n4s) * 4;
      float v[4];
      wl.load4(v, k, (i64)n, K, (i64)N);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (n + q < N) Ws[srow(k) * WS + n + q] = v[q];
    }
  }

  // ---- per-thread rows: g, g+8, g+16, g+24 of the warp's 32
  const i64 p0 = (i64)blockIdx.x * 256 + warp * 32;
  i64 roff[2][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const i64 p = p0 + g + 8 * r;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const RowMap& rm = A.rm[mi];
      const unsigned b = (unsigned)p / (unsigned)rm.lon_src;
      const int rr = (int)((unsigned)p - b * (unsigned)rm.lon_src) + rm.off;
      const bool ok = p < M && rr >= 0 && rr < rm.hi;
      roff[mi][r] = ok ? ((i64)b * rm.lon_dst + rr) * A.rs : -1;
    }
  }
  float fa[2][4], fc[2][4];   // BatchNorm fold for this thread's physical k = h*16 + 4t + c
  if (A.ac) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 0; c < 4; ++c) { fa[h][c] = A.ac[h * 16 + 4 * t + c]; fc[h][c] = A.ac[PG_WD + h * 16 + 4 * t + c]; }
  }
  __syncthreads();

  for (int n0 = 0; n0 < N; n0 += NC) {
    float acc[2][NTL][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NTL; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.0f;

    // Software pipeline over the 32-wide K segments: the 8 x 128-bit loads of segment seg+1 are in flight while
    // segment seg is multiplied.  Invalid rows load row 0 (clamped, unconditional) and are zeroed afterwards.
    float4 cur[4][2], nxt[4][2];
    bool okc[4], okn[4];
    auto issue = [&](int seg, float4 (&dst)[4][2], bool (&ok)[4]) {
      const float* base = A.P[seg] + 4 * t;
      const int mi = A.rmap[seg];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const i64 o = mi ? roff[1][r] : roff[0][r];
        ok[r] = o >= 0;
        const float* q = base + (ok[r] ? o : 0);
        dst[r][0] = __ldg(reinterpret_cast<const float4*>(q));
        dst[r][1] = __ldg(reinterpret_cast<const float4*>(q + 16));
      }
    };
    issue(0, cur, okc);
#pragma unroll 1
    for (int seg = 0; seg < A.nseg; ++seg) {
      if (seg + 1 < A.nseg) issue(seg + 1, nxt, okn);
      float af[4][4][2];   // [MMA step s = 2h+e][row][fragment half u]
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[4] = {cur[r][h].x, cur[r][h].y, cur[r][h].z, cur[r][h].w};
          if (A.ac) {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = fmaf(v[c], fa[h][c], fc[h][c]);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) af[2 * h + (c >> 1)][r][c & 1] = okc[r] ? v[c] : 0.0f;
        }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        uint32_t ah[2][4], alo[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          // a0 (row g, k t)  a1 (row g+8, k t)  a2 (row g, k t+4)  a3 (row g+8, k t+4)
          const float f[4] = {af[s][2 * i][0], af[s][2 * i + 1][0], af[s][2 * i][1], af[s][2 * i + 1][1]};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ah[i][q] = to_tf32(f[q]);
            if (MATH == 3) alo[i][q] = to_tf32(f[q] - __uint_as_float(ah[i][q]));
          }
        }
        const float* wrow = Ws + (size_t)(seg * PG_WD + s * 8 + t) * WS + n0 + g;
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
          if (n0 + j * 8 < N) {
            const float f0 = wrow[j * 8], f1 = wrow[4 * WS + j * 8];
            uint32_t bh[2] = {to_tf32(f0), to_tf32(f1)}, blo[2];
            if (MATH == 3) {
              blo[0] = to_tf32(f0 - __uint_as_float(bh[0]));
              blo[1] = to_tf32(f1 - __uint_as_float(bh[1]));
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (MATH == 3) {
                mma_tf32(acc[i][j], alo[i], bh);
                mma_tf32(acc[i][j], ah[i], blo);
              }
              mma_tf32(acc[i][j], ah[i], bh);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        cur[r][0] = nxt[r][0]; cur[r][1] = nxt[r][1];
        okc[r] = okn[r];
      }
    }
    // ---- epilogue of this column pass (same lane-pair exchange as gemm.cuh)
    const bool odd = (lane & 1) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NTL; ++j) {
        const float* c = acc[i][j];
        const float s0 = odd ? c[0] : c[2], s1 = odd ? c[1] : c[3];
        const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
        float v[4];
        if (!odd) { v[0] = c[0]; v[1] = c[1]; v[2] = r0; v[3] = r1; }
        else      { v[0] = r0;   v[1] = r1;   v[2] = c[2]; v[3] = c[3]; }
        const i64 m = p0 + i * 16 + g + (odd ? 8 : 0);
        const int n = n0 + j * 8 + (t >> 1) * 4;
        if (m < M && n < N) ep.store4(m, n, v, min(4, N - n), j);
      }
  }
  if (EP::kHasFinish) {
    __syncthreads();
    ep.template finish<1>(red, tid);
  }
}
#endif

// Host launcher.  Returns 0 after launching, or -1 when the shape is not eligible (caller falls back to gemm.cuh).
template <class T, int NC, class WL, class EP>
int launch_posgemm(const ARows& A, const WL& wl, const EP& ep, i64 M, int N, cudaStream_t stream) {
#if GWN_EMU
  (void)A; (void)wl; (void)ep; (void)M; (void)N; (void)stream;
  return -1;
#else
  const int math = current_math();
  if (math == 0 || A.nseg < 1 || A.nseg > PG_MAXSEG || N < 1 || N > 256 || M <= 0) return -1;
  const int K = A.nseg * PG_WD;
  const size_t smem_bytes = ((size_t)K * (N + 8) + 64) * sizeof(float);
  if (smem_bytes > 200 * 1024) return -1;
  const i64 blocks = (M + 255) / 256;
  if (blocks > 2147483647LL) return -1;
  static cudaError_t attr3 = cudaFuncSetAttribute(posgemm_kernel<T, WL, EP, 3, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  static cudaError_t attr1 = cudaFuncSetAttribute(posgemm_kernel<T, WL, EP, 1, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaError_t e = math == 3 ? attr3 : attr1;
  if (e == cudaSuccess) {
    if (math == 3) posgemm_kernel<T, WL, EP, 3, NC><<<(unsigned)blocks, 256, smem_bytes, stream>>>(A, wl, ep, M, N);
    else posgemm_kernel<T, WL, EP, 1, NC><<<(unsigned)blocks, 256, smem_bytes, stream>>>(A, wl, ep, M, N);
  }
  if (e != cudaSuccess) {
    set_error("posgemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return GWN_ERR_CUDA;
  }
  GWN_LAUNCH_CHECK();
  count_launch();
  return 0;
#endif
}

}  // namespace gwn
