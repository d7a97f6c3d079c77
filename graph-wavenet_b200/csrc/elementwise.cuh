// Non-GEMM kernels of the hot path: layout conversion, support packing, adaptive adjacency
// (model.py:187) forward/backward, start conv (model.py:181), BatchNorm finalize / backward apply
// (model.py:236), head-window add.  Each kernel body is written once; under GWN_HOST_EMU (tests only)
// the launch macro turns into a serial host loop over the same body.
#pragma once
#include "common.cuh"

namespace gwn {

#if GWN_EMU
#define GWN_GLOBAL static void
#define GWN_FOR_EACH(i, n) for (i64 i = 0; i < (i64)(n); ++i)
#define GWN_FOR_EACH_WARP_ROW(row, nrows, lane, WS) \
  const int WS = 1;                                 \
  const int lane = 0;                               \
  for (i64 row = 0; row < (i64)(nrows); ++row)
#define GWN_LAUNCH_1D(kernel, n, stream, ...) \
  do {                                        \
    (void)(stream);                           \
    kernel(__VA_ARGS__);                      \
  } while (0)
#define GWN_LAUNCH_WARP_ROWS(kernel, nrows, stream, ...) GWN_LAUNCH_1D(kernel, nrows, stream, __VA_ARGS__)
template <class T>
inline T warp_sum(T v) { return v; }
template <class T>
inline T warp_max(T v) { return v; }
#else
#define GWN_GLOBAL static __global__ void
#define GWN_FOR_EACH(i, n)                                                                  \
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < (i64)(n); i += (i64)gridDim.x * blockDim.x)
#define GWN_FOR_EACH_WARP_ROW(row, nrows, lane, WS)                                         \
  const int WS = 32;                                                                        \
  const int lane = threadIdx.x & 31;                                                        \
  for (i64 row = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < (i64)(nrows);     \
       row += ((i64)gridDim.x * blockDim.x) >> 5)
#define GWN_LAUNCH_1D(kernel, n, stream, ...)                                               \
  do {                                                                                      \
    i64 _n = (i64)(n);                                                                      \
    if (_n > 0) {                                                                           \
      i64 _b = (_n + 255) / 256;                                                            \
      if (_b > 148 * 32) _b = 148 * 32;                                                     \
      GWN_CUDA(::gwn::launch_kernel(kernel, dim3((unsigned)_b), dim3(256), 0, stream, __VA_ARGS__)); \
      ::gwn::count_launch();                                                                \
    }                                                                                       \
  } while (0)
#define GWN_LAUNCH_WARP_ROWS(kernel, nrows, stream, ...)                                    \
  do {                                                                                      \
    i64 _n = (i64)(nrows);                                                                  \
    if (_n > 0) {                                                                           \
      i64 _b = (_n + 7) / 8;                                                                \
      if (_b > 148 * 16) _b = 148 * 16;                                                     \
      GWN_CUDA(::gwn::launch_kernel(kernel, dim3((unsigned)_b), dim3(256), 0, stream, __VA_ARGS__)); \
      ::gwn::count_launch();                                                                \
    }                                                                                       \
  } while (0)
template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

struct Strides4 {
  i64 s[4];
};
struct Sizes4 {
  i64 n[4];
};

// dst[i0,i1,i2,i3] = src[i0,i1,i2,i3]; iteration is dst-major over `order` (a permutation of the dims
// sorted by decreasing dst stride) so writes coalesce.
GWN_GLOBAL permute4d_kernel(const float* src, Strides4 ss, float* dst, Strides4 ds, Sizes4 sz, Sizes4 ord, i64 total) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, total) {
    i64 r = i, idx[4];
#pragma unroll
    for (int d = 3; d >= 0; --d) {
      int dim = (int)ord.n[d];
      idx[dim] = r % sz.n[dim];
      r /= sz.n[dim];
    }
    i64 so = 0, dof = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      so += idx[d] * ss.s[d];
      dof += idx[d] * ds.s[d];
    }
    dst[dof] = src[so];
  }
}

// Pack a static support into zero-padded row-major A[v*ld + w] and its transpose AT[w*ld + v].
GWN_GLOBAL support_pack_kernel(const float* A, i64 rs, i64 cs, float* Ap, float* ATp, int N, int ld) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, (i64)N * ld) {
    int v = (int)(i / ld), w = (int)(i - (i64)v * ld);
    float val = (w < N) ? A[v * rs + w * cs] : 0.0f;
    Ap[(i64)v * ld + w] = val;
    if (w < N) ATp[(i64)w * ld + v] = val;
    else ATp[(i64)v * ld + w] = 0.0f;
  }
}

// adp = softmax(relu(E1 @ E2), dim=1)   (model.py:187); one warp per row; writes A and A^T (padded).
GWN_GLOBAL adp_fwd_kernel(const float* E1, const float* E2, int R, float* Ap, float* ATp, int N, int ld) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH_WARP_ROW(v, N, lane, WS) {
    float* row = Ap + v * ld;
    float mx = 0.0f;  // relu output is >= 0
    for (int w = lane; w < N; w += WS) {
      float acc = 0.0f;
      for (int k = 0; k < R; ++k) acc = fmaf(E1[v * R + k], E2[(i64)k * N + w], acc);
      acc = fmaxf(acc, 0.0f);
      row[w] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int w = lane; w < N; w += WS) {
      float e = expf(row[w] - mx);
      row[w] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    float inv = 1.0f / sum;
    for (int w = lane; w < ld; w += WS) {
      float p = (w < N) ? row[w] * inv : 0.0f;
      row[w] = p;
      if (w < N) ATp[(i64)w * ld + v] = p;
      else ATp[v * ld + w] = 0.0f;
    }
  }
}

// Backward of adp: dR = P*(dP - sum_w dP*P) masked by (E1@E2 > 0); dE1[v,k] = sum_w dR[v,w] E2[k,w].
GWN_GLOBAL adp_bwd_rows_kernel(const float* dA, const float* Ap, const float* E1, const float* E2, int R, float* dR,
                               float* dE1, int N, int ld) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH_WARP_ROW(v, N, lane, WS) {
    float s = 0.0f;
    for (int w = lane; w < N; w += WS) s = fmaf(dA[v * ld + w], Ap[v * ld + w], s);
    s = warp_sum(s);
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
    for (int w = lane; w < N; w += WS) {
      float pre = 0.0f;
      for (int k = 0; k < R; ++k) pre = fmaf(E1[v * R + k], E2[(i64)k * N + w], pre);
      float g = (pre > 0.0f) ? Ap[v * ld + w] * (dA[v * ld + w] - s) : 0.0f;
      dR[v * ld + w] = g;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < R) acc[k] = fmaf(g, E2[(i64)k * N + w], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k < R) {
        float t = warp_sum(acc[k]);
        if (lane == 0) dE1[v * R + k] = t;
      }
    }
  }
}
// dE2[k,w] = sum_v E1[v,k] dR[v,w]
GWN_GLOBAL adp_bwd_cols_kernel(const float* dR, const float* E1, int R, float* dE2, int N, int ld) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, (i64)R * N) {
    int k = (int)(i / N), w = (int)(i - (i64)k * N);
    float acc = 0.0f;
    for (int v = 0; v < N; ++v) acc = fmaf(E1[(i64)v * R + k], dR[(i64)v * ld + w], acc);
    dE2[i] = acc;
  }
}

// start_conv (model.py:181) fused with the receptive-field left pad (model.py:176-180):
// x0[b,t,n,c] = bias[c] + sum_f W[c,f] * (t >= pad ? in[b,f,n,t-pad] : 0)
GWN_GLOBAL start_fwd_kernel(const float* in, Strides4 is, const float* W, const float* bias, float* x0, int B, int F,
                            int N, int L0, int pad, int C) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, (i64)B * L0 * N * C) {
    int c = (int)(i % C);
    i64 p = i / C;
    int n = (int)(p % N);
    i64 bt = p / N;
    int t = (int)(bt % L0);
    i64 b = bt / L0;
    float acc = bias[c];
    if (t >= pad)
      for (int f = 0; f < F; ++f) acc = fmaf(W[c * F + f], in[b * is.s[0] + f * is.s[1] + n * is.s[2] + (t - pad) * is.s[3]], acc);
    x0[i] = acc;
  }
}
// grad wrt the network input, contiguous [B,F,N,T]: gi[b,f,n,t] = sum_c W[c,f] dx0[(b,t+pad,n), c]
GWN_GLOBAL start_dgrad_kernel(const float* dx0, const float* W, float* gi, int B, int F, int N, int T, int L0, int pad,
                              int C) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, (i64)B * F * N * T) {
    int t = (int)(i % T);
    i64 r = i / T;
    int n = (int)(r % N);
    r /= N;
    int f = (int)(r % F);
    i64 b = r / F;
    const float* d = dx0 + ((b * L0 + t + pad) * N + n) * C;
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) acc = fmaf(W[c * F + f], d[c], acc);
    gi[i] = acc;
  }
}

#if !GWN_EMU
// ---- warp-per-position versions of the start conv kernels for C == 32 (lane = channel) ------------------------------
// The element-per-thread kernels above pay three 64-bit divisions per output float and re-read the strided input
// C times: 36 us for 22 MB (0.6 TB/s) forward, and the weight gradient ran as an 82 us generic GEMM with K = 172k
// positions and 3 output columns.  Here a warp owns a position: (b, t, n) is decoded once per warp, the F input
// features are broadcast loads, the 128-byte channel row is one coalesced access.
constexpr int START_MAXF = 4;
__global__ void __launch_bounds__(256) start_fwd32_kernel(const float* __restrict__ in, Strides4 is, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ x0, int B, int F,
                                                          int N, int L0, int pad) {
  GWN_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  float w[START_MAXF];
#pragma unroll
  for (int f = 0; f < START_MAXF; ++f) w[f] = f < F ? W[lane * F + f] : 0.0f;
  const float bs = bias[lane];
  const unsigned P = (unsigned)B * (unsigned)L0 * (unsigned)N;
  const unsigned nw = (gridDim.x * blockDim.x) >> 5;
  // four positions per iteration: their (dependent) input loads are in flight together
  for (unsigned pb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pb < P; pb += 4 * nw) {
    float xin[4][START_MAXF];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned p0 = pb + u * nw;
      const unsigned pc = p0 < P ? p0 : 0u;
      const unsigned n = pc % (unsigned)N, bt = pc / (unsigned)N;
      const unsigned t = bt % (unsigned)L0, b = bt / (unsigned)L0;
      const bool live = p0 < P && (int)t >= pad;
      const float* src = in + (i64)b * is.s[0] + (i64)n * is.s[2] + (i64)(live ? (int)t - pad : 0) * is.s[3];
#pragma unroll
      for (int f = 0; f < START_MAXF; ++f) xin[u][f] = (live && f < F) ? __ldg(src + (i64)f * is.s[1]) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned p0 = pb + u * nw;
      float acc = bs;
#pragma unroll
      for (int f = 0; f < START_MAXF; ++f) acc = fmaf(w[f], xin[u][f], acc);
      if (p0 < P) x0[(size_t)p0 * 32 + lane] = acc;
    }
  }
}
// dW[c][f] += sum_p dx0[p][c] * in(p, f), db[c] += sum_p dx0[p][c]  (padding positions contribute to db only).
__global__ void __launch_bounds__(256) start_wgrad32_kernel(const float* __restrict__ dx0, const float* __restrict__ in, Strides4 is,
                                                            float* dW, float* db, int B, int F, int N, int L0, int pad) {
  __shared__ float red[8][32][START_MAXF + 1];
  GWN_PDL_ENTRY();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[START_MAXF + 1];
#pragma unroll
  for (int f = 0; f <= START_MAXF; ++f) acc[f] = 0.0f;
  const unsigned P = (unsigned)B * (unsigned)L0 * (unsigned)N;
  const unsigned nw = (gridDim.x * blockDim.x) >> 5;
  // four positions per iteration: 4 x 128 B of dx0 and their input features in flight per warp (one position at a
  // time left ~600 KB in flight on the whole GPU: 0.6 TB/s)
  for (unsigned pb = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pb < P; pb += 4 * nw) {
    float v[4], xin[4][START_MAXF];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned p0 = pb + u * nw;
      const unsigned pc = p0 < P ? p0 : 0u;
      const unsigned n = pc % (unsigned)N, bt = pc / (unsigned)N;
      const unsigned t = bt % (unsigned)L0, b = bt / (unsigned)L0;
      const bool live = p0 < P && (int)t >= pad;
      v[u] = p0 < P ? __ldg(dx0 + (size_t)p0 * 32 + lane) : 0.0f;
      const float* src = in + (i64)b * is.s[0] + (i64)n * is.s[2] + (i64)(live ? (int)t - pad : 0) * is.s[3];
#pragma unroll
      for (int f = 0; f < START_MAXF; ++f) xin[u][f] = (live && f < F) ? __ldg(src + (i64)f * is.s[1]) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc[START_MAXF] += v[u];
#pragma unroll
      for (int f = 0; f < START_MAXF; ++f) acc[f] = fmaf(v[u], xin[u][f], acc[f]);
    }
  }
#pragma unroll
  for (int f = 0; f <= START_MAXF; ++f) red[warp][lane][f] = acc[f];
  __syncthreads();
  for (int k = threadIdx.x; k < 32 * (START_MAXF + 1); k += blockDim.x) {
    const int c = k / (START_MAXF + 1), f = k - c * (START_MAXF + 1);
    float sum = 0.0f;
#pragma unroll
    for (int y = 0; y < 8; ++y) sum += red[y][c][f];
    if (f == START_MAXF) atomicAdd(db + c, sum);
    else if (f < F) atomicAdd(dW + c * F + f, sum);
  }
}
// Block-per-(b, t) versions: the warp-per-position kernels above decode (b, t, n) with four integer divisions per
// position and keep 32 lanes busy with one 128-byte row -- ~100 warp instructions per position, 28 / 25 us for tensors a
// streaming kernel moves in 5.  Here a block owns one (sample, time step) row of N positions, decoded ONCE; a thread owns
// 4 channels (one 128-bit access) of every 32nd node, its weights in registers.  Same FMA order as above.
__global__ void __launch_bounds__(256) start_fwd32b_kernel(const float* __restrict__ in, Strides4 is, const float* __restrict__ W,
                                                           const float* __restrict__ bias, float* __restrict__ x0, int F, int N,
                                                           int L0, int pad) {
  GWN_PDL_ENTRY();
  const int bt = blockIdx.x, b = bt / L0, t = bt - b * L0;
  const int c4 = (threadIdx.x & 7) * 4;
  float w[4][START_MAXF], bs[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bs[j] = bias[c4 + j];
#pragma unroll
    for (int f = 0; f < START_MAXF; ++f) w[j][f] = f < F ? W[(c4 + j) * F + f] : 0.0f;
  }
  const bool live = t >= pad;
  const float* src = in + (i64)b * is.s[0] + (i64)(live ? t - pad : 0) * is.s[3];
  float* dst = x0 + (size_t)bt * N * 32 + c4;
  for (int n = threadIdx.x >> 3; n < N; n += 32) {
    float xin[START_MAXF];
#pragma unroll
    for (int f = 0; f < START_MAXF; ++f) xin[f] = (live && f < F) ? __ldg(src + (i64)n * is.s[2] + (i64)f * is.s[1]) : 0.0f;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = bs[j];
#pragma unroll
      for (int f = 0; f < START_MAXF; ++f) acc = fmaf(w[j][f], xin[f], acc);
      o[j] = acc;
    }
    st4(dst + (size_t)n * 32, make_float4(o[0], o[1], o[2], o[3]));
  }
}
__global__ void __launch_bounds__(256) start_wgrad32b_kernel(const float* __restrict__ dx0, const float* __restrict__ in, Strides4 is,
                                                             float* dW, float* db, int F, int N, int L0, int pad, int nbt) {
  __shared__ float red[8][8][4 * (START_MAXF + 1)];   // [warp][channel quad][4 channels x (F features + bias)]
  GWN_PDL_ENTRY();
  const int q = threadIdx.x & 7, c4 = q * 4, warp = threadIdx.x >> 5;
  float acc[4][START_MAXF + 1];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int f = 0; f <= START_MAXF; ++f) acc[j][f] = 0.0f;
  for (int bt = blockIdx.x; bt < nbt; bt += gridDim.x) {   // a few rows per block: 160 atomics per BLOCK at the end
    const int b = bt / L0, t = bt - b * L0;
    const bool live = t >= pad;
    const float* src = in + (i64)b * is.s[0] + (i64)(live ? t - pad : 0) * is.s[3];
    const float* g = dx0 + (size_t)bt * N * 32 + c4;
#pragma unroll 4
    for (int n = threadIdx.x >> 3; n < N; n += 32) {
      const float4 v4 = ld4(g + (size_t)n * 32);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      float xin[START_MAXF];
#pragma unroll
      for (int f = 0; f < START_MAXF; ++f) xin[f] = (live && f < F) ? __ldg(src + (i64)n * is.s[2] + (i64)f * is.s[1]) : 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][START_MAXF] += v[j];
#pragma unroll
        for (int f = 0; f < START_MAXF; ++f) acc[j][f] = fmaf(v[j], xin[f], acc[j][f]);
      }
    }
  }
  // the 4 node groups of a warp (lanes q, q + 8, q + 16, q + 24), then the 8 warps through shared memory
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int f = 0; f <= START_MAXF; ++f) {
      float x = acc[j][f];
      x += __shfl_xor_sync(0xffffffffu, x, 8);
      x += __shfl_xor_sync(0xffffffffu, x, 16);
      if ((threadIdx.x & 31) < 8) red[warp][q][j * (START_MAXF + 1) + f] = x;
    }
  __syncthreads();
  constexpr int NV = 32 * (START_MAXF + 1);
  for (int k = threadIdx.x; k < NV; k += blockDim.x) {
    const int c = k / (START_MAXF + 1), f = k - c * (START_MAXF + 1);
    float sum = 0.0f;
#pragma unroll
    for (int y = 0; y < 8; ++y) sum += red[y][c >> 2][(c & 3) * (START_MAXF + 1) + f];
    if (f == START_MAXF) atomicAdd(db + c, sum);
    else if (f < F) atomicAdd(dW + c * F + f, sum);
  }
}
// dE2[k,w] += sum_{v in this block's slice} E1[v,k] dR[v,w]: thread = column w (coalesced dR rows, broadcast E1),
// blockIdx.y = slice of v; the element-per-output kernel above ran 207-long dependent chains on 9 blocks (38 us).
__global__ void __launch_bounds__(256) adp_bwd_cols_split_kernel(const float* __restrict__ dR, const float* __restrict__ E1, int R,
                                                                 float* dE2, int N, int ld, int vchunk) {
  GWN_PDL_ENTRY();
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= N) return;
  const int v0 = blockIdx.y * vchunk, v1 = min(N, v0 + vchunk);
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
  for (int v = v0; v < v1; ++v) {
    const float g = __ldg(dR + (size_t)v * ld + w);
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < R) acc[k] = fmaf(__ldg(E1 + (size_t)v * R + k), g, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (k < R) atomicAdd(dE2 + (size_t)k * N + w, acc[k]);
}
#endif

// BatchNorm2d training-mode finalize (model.py:236): batch mean / biased var -> fold constants
// a = gamma*rstd, c = beta - mean*a; running stats with momentum and unbiased var; num_batches_tracked += 1.
// batch statistics of channel c -> mean, biased variance, rstd and the fold constants a = gamma*rstd, cc = beta - mean*a
GWN_HD void bn_consts(const double* sums, double count, const float* gamma, const float* beta, float eps, int c, int C,
                      double& mean, double& var, float& rstd, float& a, float& cc) {
  mean = stat_sum(sums, c, C) / count;
  var = stat_sum(sums, C + c, C) / count - mean * mean;
  if (var < 0.0) var = 0.0;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
  a = gamma[c] * rstd;
  cc = beta[c] - (float)mean * a;
}
GWN_HD void bn_finalize_channel(const double* sums, double count, const float* gamma, const float* beta, float* rmean,
                                float* rvar, long long* nbt, float eps, float momentum, float* ac, float* mr, int C, int c) {
  double mean, var;
  float rstd, a, cc;
  bn_consts(sums, count, gamma, beta, eps, c, C, mean, var, rstd, a, cc);
  ac[c] = a;
  ac[C + c] = cc;
  mr[c] = (float)mean;
  mr[C + c] = rstd;
  double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
  rmean[c] = (1.0f - momentum) * rmean[c] + momentum * (float)mean;
  rvar[c] = (1.0f - momentum) * rvar[c] + momentum * (float)unbiased;
  if (c == 0) nbt[0] += 1;
}
GWN_GLOBAL bn_finalize_kernel(const double* sums, double count, const float* gamma, const float* beta, float* rmean,
                              float* rvar, long long* nbt, float eps, float momentum, float* ac, float* mr, int C) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(c, C) { bn_finalize_channel(sums, count, gamma, beta, rmean, rvar, nbt, eps, momentum, ac, mr, C, (int)c); }
}
// Eval mode: fold constants from the running statistics.
GWN_GLOBAL bn_eval_kernel(const float* gamma, const float* beta, const float* rmean, const float* rvar, float eps,
                          float* ac, float* mr, int C) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(c, C) {
    float rstd = 1.0f / sqrtf(rvar[c] + eps);
    float a = gamma[c] * rstd;
    ac[c] = a;
    ac[C + c] = beta[c] - rmean[c] * a;
    mr[c] = rmean[c];
    mr[C + c] = rstd;
  }
}
// BatchNorm backward apply, in place on dy:  du = a*(dy - S1/n - xhat*S2/n)  (train) | a*dy (eval);
// also emits dgamma = S2, dbeta = S1.
// `dh` (nullable) additionally receives du * dropout-keep, the gradient wrt the pre-dropout mlp output, so that the
// mlp gradient GEMMs need not regenerate the mask (i is visited in groups of 4 consecutive elements per thread).
GWN_GLOBAL bn_bwd_apply_kernel(float* dy, const float* u, const float* ac, const float* mr, const double* bsum,
                               double count, int training, float* dgamma, float* dbeta, i64 P, int C, float* dh,
                               DropoutSrc drop) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i4, P * C / 4) {
    float kp[4] = {1.f, 1.f, 1.f, 1.f};
    if (dh) drop.keep4(i4 * 4, kp);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
    const i64 i = i4 * 4 + j;
    int c = (int)(i % C);
    float a = ac[c];
    float g = dy[i];
    float r;
    if (training) {
      float xh = (u[i] - mr[c]) * mr[C + c];
      float m1 = (float)(stat_sum(bsum, c, C) / count), m2 = (float)(stat_sum(bsum, C + c, C) / count);
      r = a * (g - m1 - xh * m2);
    } else {
      r = a * g;
    }
    dy[i] = r;
    if (dh) dh[i] = r * kp[j];
    if (i < C) {
      dgamma[i] = (float)stat_sum(bsum, C + (int)i, C);
      dbeta[i] = (float)stat_sum(bsum, (int)i, C);
    }
    }
  }
}

#if !GWN_EMU
// The same, vectorised: 8 consecutive channels per thread (two 128-bit accesses per tensor, one Philox block), the
// per-channel constants hoisted out of the grid-stride loop (the stride, 2048 floats per block row, is a multiple of C
// for the power-of-two widths this path is used for).  The scalar kernel above paid two fp64 divisions and a 64-bit
// modulo per element: 29 % of the HBM roof in the r01c operator table.
__global__ void __launch_bounds__(256) bn_bwd_apply8_kernel(float* __restrict__ dy, const float* __restrict__ u,
                                                            const float* __restrict__ ac, const float* __restrict__ mr,
                                                            const double* __restrict__ bsum, double count, int training,
                                                            float* dgamma, float* dbeta, i64 n8, int C, float* __restrict__ dh,
                                                            DropoutSrc drop) {
  GWN_PDL_ENTRY();
  // the replicated backward sums, added up once per block
  __shared__ double sb[1024];   // [2*C], C <= 512 (checked by the launcher)
  for (int k = threadIdx.x; k < 2 * C; k += blockDim.x) sb[k] = stat_sum(bsum, k, C);
  __syncthreads();
  const i64 t0 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)((t0 * 8) & (i64)(C - 1));
  float a[8], mean[8], rstd[8], m1[8], m2[8];
  const double inv_count = 1.0 / count;   // ONE fp64 division per thread: 16 of them (x 256 threads x ~1200 blocks) were
                                          // a ~5 us prologue per launch on a GPU with 64 fp64 lanes per SM
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = ac[c0 + j];
    mean[j] = mr[c0 + j];
    rstd[j] = mr[C + c0 + j];
    m1[j] = training ? (float)(sb[c0 + j] * inv_count) : 0.0f;
    m2[j] = training ? (float)(sb[C + c0 + j] * inv_count) : 0.0f;
  }
  if (t0 * 8 < C) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dgamma[c0 + j] = (float)sb[C + c0 + j];
      dbeta[c0 + j] = (float)sb[c0 + j];
    }
  }
  for (i64 i8 = t0; i8 < n8; i8 += (i64)gridDim.x * blockDim.x) {
    float4 g0 = ld4(dy + i8 * 8), g1 = ld4(dy + i8 * 8 + 4);
    float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float r[8];
    if (training) {
      const float4 u0 = ld4(u + i8 * 8), u1 = ld4(u + i8 * 8 + 4);
      const float uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (uu[j] - mean[j]) * rstd[j];
        r[j] = a[j] * (g[j] - m1[j] - xh * m2[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = a[j] * g[j];
    }
    st4(dy + i8 * 8, make_float4(r[0], r[1], r[2], r[3]));
    st4(dy + i8 * 8 + 4, make_float4(r[4], r[5], r[6], r[7]));
    if (dh) {
      float kp[8];
      drop.keep8(i8 * 8, kp);
      st4(dh + i8 * 8, make_float4(r[0] * kp[0], r[1] * kp[1], r[2] * kp[2], r[3] * kp[3]));
      st4(dh + i8 * 8 + 4, make_float4(r[4] * kp[4], r[5] * kp[5], r[6] * kp[6], r[7] * kp[7]));
    }
  }
}
#endif

// ---- weight packing for the tcgen05 position GEMMs (tcpos.cuh): all K-major [N][K] fp32 ----------------------------
// Gated conv, BatchNorm affine of the layer below folded in:  Wp[2ch+g][tap*C+ci] = W_g[ch][ci][tap] * a[ci];
// bias_g'[ch] = b_g[ch] + sum_{tap,ci} W_g[ch][ci][tap] * c[ci]      (conv(W, a*u + c) = conv(W*diag(a), u) + W.c)
GWN_GLOBAL pack_tcn_fwd_kernel(const float* wf, const float* wg, const float* bf, const float* bg, const float* ac, float* Wp,
                               float* bfp, float* bgp, int D, int C, float* Wlo) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH_WARP_ROW(n, 2 * D, lane, WS) {   // one warp per packed output row n = 2*ch + gate
    const int ch = (int)(n >> 1), g = (int)(n & 1);
    const float* w = g ? wg : wf;
    float extra = 0.0f;
    for (int k = lane; k < 2 * C; k += WS) {
      const int tap = k / C, ci = k - tap * C;
      const float v = w[((i64)ch * C + ci) * 2 + tap];
      const float pv = ac ? v * ac[ci] : v;
      Wp[n * (2 * C) + k] = pv;
      if (Wlo) Wlo[n * (2 * C) + k] = tf32_lo(pv);
      if (ac) extra = fmaf(v, ac[C + ci], extra);
    }
    extra = warp_sum(extra);
    if (lane == 0) (g ? bgp : bfp)[ch] = (g ? bg : bf)[ch] + extra;
  }
}
// BatchNorm finalize of layer i AND the packed gated-conv weights of layer i+1 (which fold that BatchNorm) in one
// launch -- two tiny kernels less on the critical path of every layer.  Warp rows 0..2D-1 pack (each recomputes the
// fold constants of the channels it touches from the sums: same arithmetic as bn_consts, hence the same bits as
// ac[] below), warp row 2D does the finalize duties.
GWN_GLOBAL bn_finalize_pack_kernel(const double* sums, double count, const float* gamma, const float* beta, float* rmean,
                                   float* rvar, long long* nbt, float eps, float momentum, float* ac, float* mr, int C,
                                   const float* wf, const float* wg, const float* bf, const float* bg, float* Wp, float* bfp,
                                   float* bgp, int D, float* Wlo) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH_WARP_ROW(n, 2 * D + 1, lane, WS) {
    if (n == 2 * D) {
      for (int c = lane; c < C; c += WS) bn_finalize_channel(sums, count, gamma, beta, rmean, rvar, nbt, eps, momentum, ac, mr, C, c);
    } else {
      const int ch = (int)(n >> 1), g = (int)(n & 1);
      const float* w = g ? wg : wf;
      float extra = 0.0f;
      for (int k = lane; k < 2 * C; k += WS) {
        const int tap = k / C, ci = k - tap * C;
        double mean, var;
        float rstd, a, cc;
        bn_consts(sums, count, gamma, beta, eps, ci, C, mean, var, rstd, a, cc);
        const float v = w[((i64)ch * C + ci) * 2 + tap];
        const float pv = v * a;
        Wp[n * (2 * C) + k] = pv;
        if (Wlo) Wlo[n * (2 * C) + k] = tf32_lo(pv);
        extra = fmaf(v, cc, extra);
      }
      extra = warp_sum(extra);
      if (lane == 0) (g ? bgp : bfp)[ch] = (g ? bg : bf)[ch] + extra;
    }
  }
}
// Gated conv input gradient:  Wd[ci][tap*2D + j] = W_{j&1}[j>>1][ci][tap]
GWN_GLOBAL pack_tcn_dgrad_kernel(const float* wf, const float* wg, float* Wd, int D, int C, float* Wlo) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, (i64)C * 4 * D) {
    const int ci = (int)(i / (4 * D)), k = (int)(i - (i64)ci * 4 * D);
    const int tap = k / (2 * D), j = k - tap * 2 * D;
    const float v = ((j & 1) ? wg : wf)[((i64)(j >> 1) * C + ci) * 2 + tap];
    Wd[i] = v;
    if (Wlo) Wlo[i] = tf32_lo(v);
  }
}
// Gated-conv weight gradient from the raw tcgen05 reduction (tcred.cuh, kind 1): R[(tap*C+ci)*2D + j] = sum_p u[p+tap][ci] dpre[p][j],
// S[j] = sum_p dpre[p][j].  With the BatchNorm affine x = a*u + c of the layer below:  dW = a[ci]*R + c[ci]*S,  db = S.
GWN_GLOBAL tcn_wgrad_finalize_kernel(const float* R, const float* S, const float* ac, float* dwf, float* dwg, float* dbf,
                                     float* dbg, int D, int C) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, (i64)2 * D * 2 * C + 2 * D) {
    if (i < (i64)2 * D * 2 * C) {
      const int j = (int)(i / (2 * C)), k = (int)(i - (i64)j * 2 * C);
      const int tap = k / C, ci = k - tap * C;
      float v = R[(i64)(tap * C + ci) * (2 * D) + j];
      if (ac) v = ac[ci] * v + ac[C + ci] * S[j];
      ((j & 1) ? dwg : dwf)[((i64)(j >> 1) * C + ci) * 2 + tap] = v;
    } else {
      const int j = (int)(i - (i64)2 * D * 2 * C);
      ((j & 1) ? dbg : dbf)[j >> 1] = S[j];
    }
  }
}
// WT[c][r] = W[r][c]
GWN_GLOBAL transpose_kernel(const float* W, float* WT, int R, int Cc, float* WTlo) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, (i64)R * Cc) {
    const int c = (int)(i / R), r = (int)(i - (i64)c * R);
    const float v = W[(i64)r * Cc + c];
    WT[i] = v;
    if (WTlo) WTlo[i] = tf32_lo(v);
  }
}
// Head weights for the tcgen05 position GEMMs (nL skip convs W_i [Sk][D], biases b_i [Sk]):
//   Wcat[sk][i*D + c] = W_i[sk][c]      (skip sum over the live columns as ONE GEMM with K = nL*D)
//   WT[(i*D + c)][sk] = W_i[sk][c]      (its input gradient)
//   bsum[sk] = sum_i b_i[sk]
struct SkipWeights {
  const float* w[16];
  const float* b[16];
};
GWN_GLOBAL pack_skip_kernel(SkipWeights sw, int nL, int D, int Sk, float* Wcat, float* Wcat_lo, float* WT, float* WT_lo,
                            float* bsum) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, (i64)Sk * nL * D + Sk) {
    if (i < (i64)Sk * nL * D) {
      const int sk = (int)(i / (nL * D)), k = (int)(i - (i64)sk * nL * D);
      const int l = k / D, c = k - l * D;
      const float v = sw.w[l][(i64)sk * D + c];
      if (Wcat) { Wcat[i] = v; if (Wcat_lo) Wcat_lo[i] = tf32_lo(v); }
      if (WT) { WT[(i64)k * Sk + sk] = v; if (WT_lo) WT_lo[(i64)k * Sk + sk] = tf32_lo(v); }
    } else if (bsum) {
      const int sk = (int)(i - (i64)Sk * nL * D);
      float acc = 0.0f;
      for (int l = 0; l < nL; ++l) acc += sw.b[l][sk];
      bsum[sk] = acc;
    }
  }
}
// W2T[e*ldo + o] = W2[o*E + e] (zero for o >= O): the last head layer transposed, rows padded to ldo floats
GWN_GLOBAL pack_e2t_kernel(const float* W2, float* WT, float* WT_lo, int O, int E, int ldo) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, (i64)E * ldo) {
    const int e = (int)(i / ldo), o = (int)(i - (i64)e * ldo);
    const float v = o < O ? W2[(i64)o * E + e] : 0.0f;
    WT[i] = v;
    if (WT_lo) WT_lo[i] = tf32_lo(v);
  }
}
// lo[i] = w[i] - tf32_trunc(w[i])
// dst[e] = src[e] * dropout keep-scale of flat element e (n % 4 == 0): the gradient wrt a pre-dropout tensor
GWN_GLOBAL apply_dropout_kernel(const float* src, float* dst, i64 n, DropoutSrc drop) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i4, n / 4) {
    float kp[4];
    drop.keep4(i4 * 4, kp);
    const float4 v = ld4(src + i4 * 4);
    st4(dst + i4 * 4, make_float4(v.x * kp[0], v.y * kp[1], v.z * kp[2], v.w * kp[3]));
  }
}

GWN_GLOBAL split_lo_kernel(const float* w, float* lo, i64 n) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, n) { lo[i] = tf32_lo(w[i]); }
}

// dst[p, c] = (src ? src[p,c] : 0) + (t >= L - T_out ? win[(b, t-(L-T_out), n), c] : 0)
GWN_GLOBAL add_window_kernel(float* dst, const float* src, const float* win, int B, int L, int N, int C, int T_out) {
  GWN_PDL_ENTRY();
  GWN_FOR_EACH(i, (i64)B * L * N * C) {
    i64 p = i / C;
    int c = (int)(i - p * C);
    int n = (int)(p % N);
    i64 bt = p / N;
    int t = (int)(bt % L);
    i64 b = bt / L;
    float v = src ? src[i] : 0.0f;
    if (t >= L - T_out) v += win[((b * T_out + (t - (L - T_out))) * N + n) * C + c];
    dst[i] = v;
  }
}

}  // namespace gwn
