// Generic fp32 SIMT GEMM   C[m,n] (+)= sum_k A(k,m) * B(k,n)
// with operand access and the epilogue supplied as functors.  This is the fp32 parity tier
// (GWN_PREC_FP32) of every contraction on the Graph WaveNet hot path: the node contraction of
// nconv (model.py:13), the 1x1 / (1,2) convolutions and every weight gradient.
//
//   Loader concept
//     static constexpr bool kInner;
//     void init(int bz);                                // batch index (blockIdx.z / ksplit)
//     void load4(float (&v)[4], int k, int x, int K, int X) const;
//        kInner == false : v[i] = op(k, x+i)   (x % 4 == 0)   -- "K-outer", x contiguous
//        kInner == true  : v[i] = op(k+i, x)   (k % 4 == 0)   -- "K-inner", k contiguous
//        out-of-range elements (k >= K or x >= X) must read as 0.
//   Epilogue concept
//     static constexpr bool kHasFinish;
//     void init(int bz);
//     void store4(i64 m, int n, const float (&v)[4], int nvalid, int gn);   // m < M, n % 4 == 0, 1 <= nvalid <= 4
//     void finish(float* smem, int tid);                                     // block-wide (all threads call it)
//
// Tile: BM x BN x 16, (BM/TM)*(BN/TN) threads, TM x TN register micro-tile split in 4-wide groups that
// are strided across the tile so that shared-memory fragment reads are 128-bit and conflict-free.
#pragma once
#include "common.cuh"

namespace gwn {

template <int BM_, int BN_, int TM_, int TN_>
struct Tile {
  static constexpr int BM = BM_, BN = BN_, BK = 16, TM = TM_, TN = TN_;
  static constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
  static constexpr int GM = TM / 4, GN = TN / 4;
  static constexpr int AS = BM + 4, BS = BN + 4;  // padded smem row strides (floats, multiple of 4)
  static_assert(TM % 4 == 0 && TN % 4 == 0, "micro-tile must be made of 4-wide groups");
  static_assert(NT >= 64 && NT <= 1024, "bad thread count");
};

struct GemmShape {
  i64 M;
  int N, K;
  int ksplit;   // >= 1; > 1 requires an accumulating (atomic) epilogue
  int nbatch;   // >= 1
};

#if !GWN_EMU
template <class T, class L, int ROWLEN, int STRIDE>
__device__ __forceinline__ void tile_fetch(const L& ld, float (&r)[(T::BK * ROWLEN / 4 + T::NT - 1) / T::NT][4], int k0,
                                           i64 x0, int K, i64 X, int tid) {
  constexpr int G = T::BK * ROWLEN / 4;
  constexpr int PER = (G + T::NT - 1) / T::NT;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    int g = tid + j * T::NT;
    if (G % T::NT == 0 || g < G) {
      if (!L::kInner) {
        int k = g / (ROWLEN / 4), x = (g % (ROWLEN / 4)) * 4;
        ld.load4(r[j], k0 + k, x0 + x, K, X);
      } else {
        int x = g / (T::BK / 4), k = (g % (T::BK / 4)) * 4;
        ld.load4(r[j], k0 + k, x0 + x, K, X);
      }
    }
  }
}

template <class T, class L, int ROWLEN, int STRIDE>
__device__ __forceinline__ void tile_commit(float* s, const float (&r)[(T::BK * ROWLEN / 4 + T::NT - 1) / T::NT][4],
                                            int tid) {
  constexpr int G = T::BK * ROWLEN / 4;
  constexpr int PER = (G + T::NT - 1) / T::NT;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    int g = tid + j * T::NT;
    if (G % T::NT == 0 || g < G) {
      if (!L::kInner) {
        int k = g / (ROWLEN / 4), x = (g % (ROWLEN / 4)) * 4;
        *reinterpret_cast<float4*>(s + k * STRIDE + x) = make_float4(r[j][0], r[j][1], r[j][2], r[j][3]);
      } else {
        int x = g / (T::BK / 4), k = (g % (T::BK / 4)) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) s[(k + i) * STRIDE + x] = r[j][i];
      }
    }
  }
}

template <class T, class AL, class BL, class EP>
__global__ void __launch_bounds__(T::NT, 512 / T::NT) gemm_kernel(AL al, BL bl, EP ep, i64 M, int N, int K, int kchunk, int ksplit) {
  constexpr int BM = T::BM, BN = T::BN, BK = T::BK, TM = T::TM, TN = T::TN;
  constexpr int AS = T::AS, BS = T::BS;
  __shared__ __align__(16) float As[2][BK * AS];
  __shared__ __align__(16) float Bs[2][BK * BS];

  const int tid = threadIdx.x;
  const int tx = tid % T::TX, ty = tid / T::TX;
  const int bz = blockIdx.z / ksplit, kz = blockIdx.z % ksplit;
  al.init(bz);
  bl.init(bz);
  ep.init(bz);
  const i64 m0 = (i64)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int kbeg = kz * kchunk;
  const int kend = min(K, kbeg + kchunk);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  constexpr int APER = (BK * BM / 4 + T::NT - 1) / T::NT;
  constexpr int BPER = (BK * BN / 4 + T::NT - 1) / T::NT;
  float ra[APER][4], rb[BPER][4];

  if (kbeg < kend) {
    tile_fetch<T, AL, BM, AS>(al, ra, kbeg, m0, kend, M, tid);
    tile_fetch<T, BL, BN, BS>(bl, rb, kbeg, (i64)n0, kend, (i64)N, tid);
    tile_commit<T, AL, BM, AS>(As[0], ra, tid);
    tile_commit<T, BL, BN, BS>(Bs[0], rb, tid);
  }
  __syncthreads();

  int cur = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK) < kend;
    if (more) {
      tile_fetch<T, AL, BM, AS>(al, ra, k0 + BK, m0, kend, M, tid);
      tile_fetch<T, BL, BN, BS>(bl, rb, k0 + BK, (i64)n0, kend, (i64)N, tid);
    }
    const float* as = As[cur];
    const float* bs = Bs[cur];
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < T::GM; ++g) {
        float4 v = *reinterpret_cast<const float4*>(as + kk * AS + g * (BM / T::GM) + ty * 4);
        a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < T::GN; ++g) {
        float4 v = *reinterpret_cast<const float4*>(bs + kk * BS + g * (BN / T::GN) + tx * 4);
        b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      tile_commit<T, AL, BM, AS>(As[cur ^ 1], ra, tid);
      tile_commit<T, BL, BN, BS>(Bs[cur ^ 1], rb, tid);
    }
    __syncthreads();
    cur ^= 1;
  }

  if (kbeg < kend || kz == 0) {
#pragma unroll
    for (int gm = 0; gm < T::GM; ++gm)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const i64 m = m0 + gm * (BM / T::GM) + ty * 4 + i;
        if (m < M) {
#pragma unroll
          for (int gn = 0; gn < T::GN; ++gn) {
            const int n = n0 + gn * (BN / T::GN) + tx * 4;
            if (n < N) {
              float v[4] = {acc[gm * 4 + i][gn * 4 + 0], acc[gm * 4 + i][gn * 4 + 1], acc[gm * 4 + i][gn * 4 + 2],
                            acc[gm * 4 + i][gn * 4 + 3]};
              ep.store4(m, n, v, min(4, N - n), gn);
            }
          }
        }
      }
  }
  if (EP::kHasFinish) {
    __syncthreads();
    ep.finish(As[0], tid);
  }
}
#endif  // !GWN_EMU

// Launch (or, in the test-only host emulation, run serially) one GEMM.
template <class T, class AL, class BL, class EP>
int launch_gemm(const AL& al, const BL& bl, const EP& ep, const GemmShape& s, cudaStream_t stream) {
  if (s.M <= 0 || s.N <= 0) return 0;
  GWN_CHECK_ARG(s.ksplit >= 1 && s.nbatch >= 1, "gemm: bad ksplit/nbatch");
  int ksplit = s.ksplit;
  int kchunk = (s.K + ksplit - 1) / ksplit;
  kchunk = (kchunk + T::BK - 1) / T::BK * T::BK;
  if (kchunk < T::BK) kchunk = T::BK;
  ksplit = s.K > 0 ? (s.K + kchunk - 1) / kchunk : 1;
#if !GWN_EMU
  i64 gx = (s.M + T::BM - 1) / T::BM;
  i64 gy = (s.N + T::BN - 1) / T::BN;
  i64 gz = (i64)ksplit * s.nbatch;
  GWN_CHECK_ARG(gx <= 2147483647LL && gy <= 65535 && gz <= 65535, "gemm: grid too large (%lld,%lld,%lld)", gx, gy, gz);
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
  gemm_kernel<T, AL, BL, EP><<<grid, T::NT, 0, stream>>>(al, bl, ep, s.M, s.N, s.K, kchunk, ksplit);
  GWN_LAUNCH_CHECK();
  count_launch();
#else
  (void)stream;
  // Serial emulation with the kernel's access granularity: 4-wide loader groups, per-split epilogue calls.
  for (int bz = 0; bz < s.nbatch; ++bz)
    for (int kz = 0; kz < ksplit; ++kz) {
      AL a = al;
      BL b = bl;
      EP e = ep;
      a.init(bz);
      b.init(bz);
      e.init(bz);
      const int kbeg = kz * kchunk, kend = std::min(s.K, kbeg + kchunk);
      const int Kp = (kend - kbeg + 3) / 4 * 4;
      const i64 Mp = (s.M + 3) / 4 * 4;
      const int Np = (s.N + 3) / 4 * 4;
      std::vector<float> At((size_t)std::max(Kp, 4) * Mp, 0.f), Bt((size_t)std::max(Kp, 4) * Np, 0.f);
      auto fill = [&](auto& ld, std::vector<float>& t, i64 X, i64 Xp) {
        using LT = typename std::decay<decltype(ld)>::type;
        for (int k = 0; k < Kp; k += (LT::kInner ? 4 : 1))
          for (i64 x = 0; x < Xp; x += (LT::kInner ? 1 : 4)) {
            float v[4];
            ld.load4(v, kbeg + k, x, kend, X);
            for (int i = 0; i < 4; ++i) {
              if (LT::kInner) t[(size_t)(k + i) * Xp + x] = v[i];
              else t[(size_t)k * Xp + x + i] = v[i];
            }
          }
      };
      fill(a, At, s.M, Mp);
      fill(b, Bt, (i64)s.N, (i64)Np);
      for (i64 m = 0; m < s.M; ++m)
        for (int n = 0; n < s.N; n += 4) {
          float v[4] = {0, 0, 0, 0};
          for (int k = 0; k < kend - kbeg; ++k)
            for (int i = 0; i < 4; ++i) v[i] = fmaf(At[(size_t)k * Mp + m], Bt[(size_t)k * Np + n + i], v[i]);
          e.store4(m, n, v, std::min(4, s.N - n), (n / 4) % T::GN);
        }
      if (EP::kHasFinish) e.finish(nullptr, 0);
    }
#endif
  return 0;
}

// Split-K heuristic for reduction GEMMs with tiny outputs: enough blocks to fill the GPU ~2x.
inline int pick_ksplit(i64 M, int N, i64 K, int BM, int BN, int target_blocks) {
  i64 tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  i64 ks = (target_blocks + tiles - 1) / tiles;
  i64 maxk = (K + 63) / 64;  // at least 64 k per split
  if (ks > maxk) ks = maxk;
  if (ks < 1) ks = 1;
  if (ks > 4096) ks = 4096;
  return (int)ks;
}

}  // namespace gwn
