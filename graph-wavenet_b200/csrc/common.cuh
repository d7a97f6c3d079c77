// Common definitions for the gwnet_b200 kernels.
//
// The same sources build two ways:
//   * nvcc, sm_100a           -> libgwnet_b200.so, the product (CUDA only, no CPU fallback);
//   * g++ -DGWN_HOST_EMU      -> tests/_hostemu/libgwnet_hostemu.so, a TEST-ONLY emulation in which
//     every kernel launch is replaced by a serial host loop over the very same loader / epilogue
//     functors.  It exists so the index arithmetic and the plan orchestration can be checked against
//     the oracle in the GPU-less build container.  The python package never loads it.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cassert>
#include <string>
#include <vector>
#include <algorithm>
#include <type_traits>
#include <utility>

#include <cuda_runtime.h>

#include "../../include/gwnet_b200.h"

#ifdef GWN_HOST_EMU
#define GWN_HD inline
#define GWN_DEV inline
#define GWN_EMU 1
#else
#define GWN_HD __host__ __device__ __forceinline__
#define GWN_DEV __device__ __forceinline__
#define GWN_EMU 0
#endif

namespace gwn {

typedef long long i64;

void set_error(const char* fmt, ...);
void count_launch();   // bumps the counter behind gwn_launch_count()

// Per-op device timing for bench.py's roofline table (gwn_profile_begin / gwn_profile_end): when recording is on,
// a scope brackets the launches of one operator with CUDA events on the launching stream and carries the
// operator's ALGORITHMIC bytes and flops.  Off (the default): two predictable branches.
struct ProfScope {
  int idx;
  cudaStream_t st;
  ProfScope(const char* tag, cudaStream_t stream, double bytes, double flops);
  ~ProfScope();
};

#define GWN_CHECK_ARG(cond, ...)                  \
  do {                                            \
    if (!(cond)) {                                \
      ::gwn::set_error(__VA_ARGS__);              \
      return GWN_ERR_INVALID;                     \
    }                                             \
  } while (0)

#define GWN_TRY(expr)                 \
  do {                                \
    int _st = (expr);                 \
    if (_st != 0) return _st;         \
  } while (0)

#if !GWN_EMU
#define GWN_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::gwn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return GWN_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)
#define GWN_LAUNCH_CHECK() GWN_CUDA(cudaGetLastError())
#endif

// Programmatic dependent launch (PDL): every kernel of the library is launched with the programmatic-stream-
// serialization attribute and executes GWN_PDL_ENTRY() (or, in the tcgen05 kernels, pdl_wait() after the
// barrier/TMEM prologue) before it touches global memory.  `griddepcontrol.wait` returns once the preceding kernel
// of the stream has completed and flushed, so the memory ordering is that of a plain stream; what overlaps is the
// launch latency and the prologue of kernel n+1 with the tail of kernel n -- ~180 launches per training step, also
// inside the captured CUDA graph (programmatic edges).  The trigger comes AFTER the wait (and after the TMEM
// allocation in the tcgen05 kernels), so at most one dependent grid is resident and waiting and it can never hold
// tensor memory that its predecessor still has to allocate.  GWNET_B200_PDL=0 turns the attribute off.
#if GWN_EMU
#define GWN_PDL_ENTRY() do {} while (0)
#else
#define GWN_PDL_ENTRY()                                             \
  do {                                                              \
    asm volatile("griddepcontrol.wait;" ::: "memory");              \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)
bool pdl_enabled();
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  if (pdl_enabled()) {
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// ---------------------------------------------------------------- small device helpers
GWN_HD float4 ld4(const float* p) {
#if GWN_EMU
  assert((reinterpret_cast<uintptr_t>(p) & 15) == 0 && "misaligned float4 load");
#endif
  return *reinterpret_cast<const float4*>(p);
}
GWN_HD void st4(float* p, float4 v) {
#if GWN_EMU
  assert((reinterpret_cast<uintptr_t>(p) & 15) == 0 && "misaligned float4 store");
#endif
  *reinterpret_cast<float4*>(p) = v;
}

GWN_HD void atomic_add_f(float* p, float v) {
#if GWN_EMU
  *p += v;
#else
#ifdef __CUDA_ARCH__
  atomicAdd(p, v);
#else
  *p += v;
#endif
#endif
}
GWN_HD void atomic_add_d(double* p, double v) {
#if GWN_EMU
  *p += v;
#else
#ifdef __CUDA_ARCH__
  atomicAdd(p, v);
#else
  *p += v;
#endif
#endif
}

GWN_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Remainder of the 3xTF32 split: tcgen05 kind::tf32 TRUNCATES an fp32 operand to its top 19 bits (probed on B200,
// tests/tools/tf32_rounding_probe.py), so x itself serves as the high part and x - trunc(x) -- exact in fp32 -- is
// the low part.
GWN_HD float tf32_lo(float v) {
  unsigned u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  float h;
  memcpy(&h, &u, 4);
  return v - h;
}

// Row remap between two BLNC tensors that share (B, N) but differ in time length:
// position p = (b, t, n) of a tensor with lon_out = L_out*N rows per sample maps to
// row (b, t + off_t, n) of a tensor with lon_in = L_in*N rows per sample; off = off_t*N.
struct Remap {
  int lon_out, lon_in, off;
  GWN_HD i64 operator()(i64 p) const {   // positions are < 2^31 (checked on the host): 32-bit division
    unsigned b = (unsigned)p / (unsigned)lon_out;
    unsigned r = (unsigned)p - b * (unsigned)lon_out;
    return (i64)b * lon_in + off + r;
  }
};

// Division by a channel width that is almost always a power of two (32, 64, 256 ...).
struct DivW {
  int d, shift;   // shift >= 0 when d == 1 << shift
  GWN_HD int div(int k) const { return shift >= 0 ? (k >> shift) : k / d; }
};
inline DivW make_divw(int d) {
  DivW w;
  w.d = d;
  w.shift = -1;
  for (int s = 0; s < 31; ++s)
    if ((1 << s) == d) w.shift = s;
  return w;
}
inline Remap make_remap(int L_out, int L_in, int off_t, int N) { return Remap{L_out * N, L_in * N, off_t * N}; }
inline Remap identity_remap() { return Remap{1 << 30, 1 << 30, 0}; }

// ---------------------------------------------------------------- Philox4x32-10 (dropout)
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  GWN_HD static void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  // 4 uniform 32-bit words for (key, counter)
  GWN_HD static void gen(uint64_t key, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t (&out)[4]) {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t h0, l0, h1, l1;
      mulhilo(M0, c0, h0, l0);
      mulhilo(M1, c2, h1, l1);
      uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

// Dropout keep-scale for 4 consecutive elements starting at flat element index e (e % 4 == 0).
struct DropoutSrc {
  int mode;              // gwn_dropout_mode
  const uint8_t* mask;   // GWN_DROPOUT_MASK
  uint64_t seed, stream; // GWN_DROPOUT_PHILOX: key = seed, counter hi = stream (layer id)
  const unsigned long long* seed_dev;   // non-null: the key is read from device memory (CUDA-graph replays draw fresh masks)
  float p, scale;        // scale = 1/(1-p)
  // One Philox4x32-10 block (128 random bits) serves 8 consecutive elements, 16 bits each: element e takes half-word
  // (e & 7) of the block with counter e >> 3 and is kept iff  h >= p * 65536  (keep probability exact to 1.5e-5).
  GWN_HD void keep4(i64 e, float (&k)[4]) const {
    if (mode == GWN_DROPOUT_NONE) {
      k[0] = k[1] = k[2] = k[3] = 1.0f;
    } else if (mode == GWN_DROPOUT_MASK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) k[i] = mask[e + i] ? scale : 0.0f;
    } else {
      uint32_t r[4];
      Philox::gen(seed_dev ? (uint64_t)*seed_dev : seed, (uint64_t)(e >> 3), stream, r);
      const uint32_t thr = (uint32_t)(p * 65536.0f);
      const int w0 = (int)((e >> 2) & 1) * 2;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t h = (r[w0 + (i >> 1)] >> (16 * (i & 1))) & 0xFFFFu;
        k[i] = (h >= thr) ? scale : 0.0f;
      }
    }
  }
  // 8 consecutive elements starting at e (e % 8 == 0): one Philox block.
  GWN_HD void keep8(i64 e, float (&k)[8]) const {
    if (mode == GWN_DROPOUT_NONE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) k[i] = 1.0f;
    } else if (mode == GWN_DROPOUT_MASK) {
#pragma unroll
      for (int i = 0; i < 8; ++i) k[i] = mask[e + i] ? scale : 0.0f;
    } else {
      uint32_t r[4];
      Philox::gen(seed_dev ? (uint64_t)*seed_dev : seed, (uint64_t)(e >> 3), stream, r);
      const uint32_t thr = (uint32_t)(p * 65536.0f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t h = (r[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
        k[i] = (h >= thr) ? scale : 0.0f;
      }
    }
  }
};
inline DropoutSrc make_dropout(int mode, const uint8_t* mask, uint64_t seed, uint64_t stream, float p,
                               const unsigned long long* seed_dev = nullptr) {
  DropoutSrc d;
  d.seed_dev = seed_dev;
  d.mode = (p <= 0.0f) ? (int)GWN_DROPOUT_NONE : mode;
  d.mask = mask;
  d.seed = seed;
  d.stream = stream;
  d.p = p;
  d.scale = (p < 1.0f) ? 1.0f / (1.0f - p) : 0.0f;
  return d;
}

// BatchNorm column statistics (forward sums, backward sums) are accumulated with fp64 atomics by every CTA of the
// producing GEMM: 296 epilogue groups onto the same 64 addresses serialise in the L2 atomic unit.  The sums therefore
// live in GWN_STAT_REPL replicas of [2*C] doubles (a CTA adds to replica blockIdx % GWN_STAT_REPL); consumers add
// the replicas up.
constexpr int GWN_STAT_REPL = 8;
GWN_HD double stat_sum(const double* s, int idx, int C) {   // element idx (< 2*C) summed over the replicas
  double v = 0.0;
#pragma unroll
  for (int r = 0; r < GWN_STAT_REPL; ++r) v += s[(size_t)r * 2 * C + idx];
  return v;
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline i64 round_up64(i64 x, i64 m) { return (x + m - 1) / m * m; }

}  // namespace gwn
