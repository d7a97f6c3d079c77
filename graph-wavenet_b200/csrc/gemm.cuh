// Generic GEMM   C[m,n] (+)= sum_k A(k,m) * B(k,n)
// with operand access and the epilogue supplied as functors.  Every contraction on the Graph WaveNet hot
// path that is NOT the node contraction of nconv (which has its own tcgen05 kernel, nconv_tc.cu) goes
// through here: the (1,2) gated convolutions, the 1x1 convolutions, and every weight gradient -- and the
// node contraction itself in the fp32 tiers.
//
// Math modes (template parameter MATH):
//   0  fp32 FMA                                   -- GWN_PREC_FP32, the bit-for-bit-stable parity tier
//   3  3xTF32 on the tensor cores (mma.sync):     -- GWN_PREC_FP32X3, fp32-grade (error ~1e-6)
//        a = a_hi + a_lo,  acc += a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
//   1  single-pass TF32 on the tensor cores       -- GWN_PREC_TF32 (what cuDNN does for the reference's convs)
//
//   Loader concept
//     static constexpr bool kInner;
//     void init(int bz);                                // batch index (blockIdx.z / ksplit)
//     void load4(float (&v)[4], int k, i64 x, int K, i64 X) const;
//        kInner == false : v[i] = op(k, x+i)   (x % 4 == 0)   -- "K-outer", x contiguous
//        kInner == true  : v[i] = op(k+i, x)   (k % 4 == 0)   -- "K-inner", k contiguous
//        out-of-range elements (k >= K or x >= X) must read as 0.
//   Epilogue concept
//     static constexpr bool kHasFinish;
//     void init(int bz);
//     void store4(i64 m, int n, const float (&v)[4], int nvalid, int slot);  // m < M, n % 4 == 0, 1 <= nvalid <= 4
//     template <int MATH> void finish(float* smem, int tid);                 // block-wide (all threads call it)
//
// Tile: BM x BN x 16.  FMA mode: (BM/TM)*(BN/TN) threads with a TM x TN register micro-tile split in 4-wide
// groups strided across the tile (128-bit conflict-free fragment reads).  Tensor mode: the same threads as
// NT/32 warps in a WM x WN grid, each owning a 32 x WTN sub-tile as 2 x (WTN/8) m16n8k8 MMAs per 8 of K.
#pragma once
#include "common.cuh"

namespace gwn {

template <int BM_, int BN_, int TM_, int TN_, int BK_ = 16>
struct Tile {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
  static constexpr int KIS = BK_ + 4;   // tensor mode: row stride of an untransposed K-inner tile (conflict-free fragments)
  static constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
  static constexpr int GM = TM / 4, GN = TN / 4;
  // tensor-core warp grid
  static constexpr int NW = NT / 32;
  static constexpr int WM = (BM / 32 < NW) ? BM / 32 : NW;
  static constexpr int WN = NW / WM;
  static constexpr int WTN = BN / WN;        // warp tile is 32 x WTN
  static constexpr int NTL = WTN / 8;        // n8 tiles per warp
  static constexpr int SLOTS = (GN > NTL) ? GN : NTL;   // 4-wide column groups a thread can own
  static_assert(TM % 4 == 0 && TN % 4 == 0, "micro-tile must be made of 4-wide groups");
  static_assert(NT >= 64 && NT <= 1024 && NT % 32 == 0, "bad thread count");
  static_assert(WM * WN == NW && BM == WM * 32 && WTN % 8 == 0 && WTN * WN == BN, "bad tensor-core warp grid");
};

struct GemmShape {
  i64 M;
  int N, K;
  int ksplit;   // >= 1; > 1 requires an accumulating (atomic) epilogue
  int nbatch;   // >= 1
};

// Math mode of the GEMMs launched by the current ABI call on this thread (set by MathScope).
int current_math();
void set_current_math(int m);
struct MathScope {
  int prev;
  explicit MathScope(int m) : prev(current_math()) { set_current_math(m); }
  ~MathScope() { set_current_math(prev); }
};

#if !GWN_EMU
template <class T, class L, int ROWLEN>
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

template <class T, class L, int ROWLEN, int STRIDE, bool NATIVE>
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
      } else if (NATIVE) {
        int x = g / (T::BK / 4), k = (g % (T::BK / 4)) * 4;
        *reinterpret_cast<float4*>(s + x * T::KIS + k) = make_float4(r[j][0], r[j][1], r[j][2], r[j][3]);
      } else {
        int x = g / (T::BK / 4), k = (g % (T::BK / 4)) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) s[(k + i) * STRIDE + x] = r[j][i];
      }
    }
  }
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <class T, class AL, class BL, class EP, int MATH>
__global__ void __launch_bounds__(T::NT, 512 / T::NT) gemm_kernel(AL al, BL bl, EP ep, i64 M, int N, int K, int kchunk, int ksplit) {
  constexpr int BM = T::BM, BN = T::BN, BK = T::BK, TM = T::TM, TN = T::TN;
  constexpr int PAD = MATH == 0 ? 4 : 8;   // tensor mode: row stride == 8 (mod 32) makes fragment reads conflict-free
  constexpr int AS = BM + PAD, BS = BN + PAD;
  constexpr bool ANAT = MATH != 0 && AL::kInner, BNAT = MATH != 0 && BL::kInner;   // untransposed K-inner tiles
  constexpr int KI_STRIDE = T::KIS;
  constexpr int ASZ = ANAT ? BM * KI_STRIDE : BK * AS, BSZ = BNAT ? BN * KI_STRIDE : BK * BS;
  __shared__ __align__(16) float As[2][ASZ];
  __shared__ __align__(16) float Bs[2][BSZ];
  GWN_PDL_ENTRY();

  const int tid = threadIdx.x;
  const int bz = blockIdx.z / ksplit, kz = blockIdx.z % ksplit;
  al.init(bz);
  bl.init(bz);
  ep.init(bz);
  const i64 m0 = (i64)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int kbeg = kz * kchunk;
  const int kend = min(K, kbeg + kchunk);

  constexpr int NACC = MATH == 0 ? TM * TN : 2 * T::NTL * 4;
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;

  constexpr int APER = (BK * BM / 4 + T::NT - 1) / T::NT;
  constexpr int BPER = (BK * BN / 4 + T::NT - 1) / T::NT;
  float ra[APER][4], rb[BPER][4];

  // FMA-mode thread coordinates
  const int tx = tid % T::TX, ty = tid / T::TX;
  // tensor-mode warp coordinates
  const int lane = tid & 31, wid = tid >> 5;
  const int wm = wid % T::WM, wn = wid / T::WM;
  const int g = lane >> 2, t = lane & 3;

  if (kbeg < kend) {
    tile_fetch<T, AL, BM>(al, ra, kbeg, m0, kend, M, tid);
    tile_fetch<T, BL, BN>(bl, rb, kbeg, (i64)n0, kend, (i64)N, tid);
    tile_commit<T, AL, BM, AS, ANAT>(As[0], ra, tid);
    tile_commit<T, BL, BN, BS, BNAT>(Bs[0], rb, tid);
  }
  __syncthreads();

  int cur = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK) < kend;
    if (more) {
      tile_fetch<T, AL, BM>(al, ra, k0 + BK, m0, kend, M, tid);
      tile_fetch<T, BL, BN>(bl, rb, k0 + BK, (i64)n0, kend, (i64)N, tid);
    }
    const float* as = As[cur];
    const float* bs = Bs[cur];
    if (MATH == 0) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int q = 0; q < T::GM; ++q) {
          float4 v = *reinterpret_cast<const float4*>(as + kk * AS + q * (BM / T::GM) + ty * 4);
          a[q * 4 + 0] = v.x; a[q * 4 + 1] = v.y; a[q * 4 + 2] = v.z; a[q * 4 + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < T::GN; ++q) {
          float4 v = *reinterpret_cast<const float4*>(bs + kk * BS + q * (BN / T::GN) + tx * 4);
          b[q * 4 + 0] = v.x; b[q * 4 + 1] = v.y; b[q * 4 + 2] = v.z; b[q * 4 + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i * TN + j] = fmaf(a[i], b[j], acc[i * TN + j]);
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < BK; kk += 8) {
        // A fragments of the warp's two m16 tiles: A[m][k] = as[k][m]
        uint32_t ah[2][4], alo[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float f[4];
          if (ANAT) {
            const float* pa = as + (wm * 32 + i * 16 + g) * KI_STRIDE + kk + t;
            f[0] = pa[0]; f[1] = pa[8 * KI_STRIDE]; f[2] = pa[4]; f[3] = pa[8 * KI_STRIDE + 4];
          } else {
            const float* pa = as + (kk + t) * AS + wm * 32 + i * 16 + g;
            f[0] = pa[0]; f[1] = pa[8]; f[2] = pa[4 * AS]; f[3] = pa[4 * AS + 8];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ah[i][q] = to_tf32(f[q]);
            if (MATH == 3) alo[i][q] = to_tf32(f[q] - __uint_as_float(ah[i][q]));
          }
        }
#pragma unroll
        for (int j = 0; j < T::NTL; ++j) {
          float f[2];
          if (BNAT) {
            const float* pb = bs + (wn * T::WTN + j * 8 + g) * KI_STRIDE + kk + t;
            f[0] = pb[0]; f[1] = pb[4];
          } else {
            const float* pb = bs + (kk + t) * BS + wn * T::WTN + j * 8 + g;
            f[0] = pb[0]; f[1] = pb[4 * BS];
          }
          uint32_t bh[2], blo[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            bh[q] = to_tf32(f[q]);
            if (MATH == 3) blo[q] = to_tf32(f[q] - __uint_as_float(bh[q]));
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            float(&c)[4] = *reinterpret_cast<float(*)[4]>(&acc[(i * T::NTL + j) * 4]);
            if (MATH == 3) {
              mma_tf32(c, alo[i], bh);
              mma_tf32(c, ah[i], blo);
            }
            mma_tf32(c, ah[i], bh);
          }
        }
      }
    }
    if (more) {
      tile_commit<T, AL, BM, AS, ANAT>(As[cur ^ 1], ra, tid);
      tile_commit<T, BL, BN, BS, BNAT>(Bs[cur ^ 1], rb, tid);
    }
    __syncthreads();
    cur ^= 1;
  }

  if (MATH == 0) {
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
              float v[4] = {acc[(gm * 4 + i) * TN + gn * 4 + 0], acc[(gm * 4 + i) * TN + gn * 4 + 1],
                            acc[(gm * 4 + i) * TN + gn * 4 + 2], acc[(gm * 4 + i) * TN + gn * 4 + 3]};
              ep.store4(m, n, v, min(4, N - n), gn);
            }
          }
        }
      }
  } else {
    // C fragment: c0,c1 = (row g, cols 2t,2t+1); c2,c3 = (row g+8, same cols).  Pair lanes (t even/odd) swap halves so
    // that each thread ends up with 4 consecutive columns of one row, the granularity the epilogues work at.
    const bool odd = (lane & 1) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < T::NTL; ++j) {
        const float* c = &acc[(i * T::NTL + j) * 4];
        const float s0 = odd ? c[0] : c[2], s1 = odd ? c[1] : c[3];
        const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
        float v[4];
        if (!odd) { v[0] = c[0]; v[1] = c[1]; v[2] = r0; v[3] = r1; }
        else      { v[0] = r0;   v[1] = r1;   v[2] = c[2]; v[3] = c[3]; }
        const i64 m = m0 + wm * 32 + i * 16 + g + (odd ? 8 : 0);
        const int n = n0 + wn * T::WTN + j * 8 + (t >> 1) * 4;
        if (m < M && n < N) ep.store4(m, n, v, min(4, N - n), j);
      }
  }
  if (EP::kHasFinish) {
    __syncthreads();
    ep.template finish<MATH>(As[0], tid);
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
  switch (current_math()) {
    case 1: GWN_CUDA(launch_kernel(gemm_kernel<T, AL, BL, EP, 1>, grid, dim3(T::NT), 0, stream, al, bl, ep, s.M, s.N, s.K, kchunk, ksplit)); break;
    case 3: GWN_CUDA(launch_kernel(gemm_kernel<T, AL, BL, EP, 3>, grid, dim3(T::NT), 0, stream, al, bl, ep, s.M, s.N, s.K, kchunk, ksplit)); break;
    default: GWN_CUDA(launch_kernel(gemm_kernel<T, AL, BL, EP, 0>, grid, dim3(T::NT), 0, stream, al, bl, ep, s.M, s.N, s.K, kchunk, ksplit)); break;
  }
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
          e.store4(m, n, v, std::min(4, s.N - n), (n / 4) % T::SLOTS);
        }
      if (EP::kHasFinish) e.template finish<0>(nullptr, 0);
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
