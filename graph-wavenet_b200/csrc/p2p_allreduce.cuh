// Data-parallel gradient exchange fused with clip_grad_norm_'s reduction (engine.py:53-54):
// ONE kernel per step reads every rank's flat gradient buffer over NVLink peer memory, sums the replicas in rank order
// (so every rank computes bit-identical sums), writes the sum to a local buffer and accumulates the squared norm of the
// live elements -- the all-reduce, the 1/world-independent norm pass and the Adam step counter in one launch that a
// CUDA graph can hold (no NCCL call, no second graph, no host involvement).
//
// Memory: every rank owns one allocation [flags | gradient] made by gwn_p2p_alloc (cudaMalloc + IPC handle) and maps
// its peers' allocations with gwn_p2p_open (cudaIpcOpenMemHandle, peer access enabled lazily by the driver).
// Protocol (epoch = Adam step count + 1, identical on all ranks; flags only ever grow):
//   1. block 0 writes READY[rank] = epoch into every peer's flag block -- this kernel starts after the backward pass
//      of its own stream, so the local gradient is complete;
//   2. every block waits until READY[q] >= epoch for all q in its LOCAL flag block, then sums its slice of all W
//      buffers (peer loads bypass L1: ld.volatile) and writes the result to `out`;
//   3. the last block of the grid to finish (local atomic counter) writes DONE[rank] = epoch to every peer, waits for
//      DONE[q] >= epoch from all of them and only then lets the kernel complete: the next kernel of the stream (Adam,
//      and a millisecond later the next backward pass) may overwrite the gradient buffer, so no peer may still read it.
// Every wait is clock-bounded (a lost peer must not hang the GPU): on time-out the error flag of the tcgen05 kernels is
// latched (gwn_tc_error_flag) and the kernel continues.
#pragma once
#include "train_tail.cuh"

#if !GWN_EMU
#include "tc_common.cuh"

namespace gwn {

constexpr int P2P_MAXRANKS = 8;
constexpr int P2P_FLAG_BYTES = 4096;        // header of every rank's allocation: READY[8], DONE[8], block counter
constexpr int P2P_THREADS = 512;

struct P2PArgs {
  const float* grad[P2P_MAXRANKS];          // every rank's gradient buffer as mapped in THIS process (grad[rank] = local)
  unsigned* flags[P2P_MAXRANKS];            // every rank's flag block (flags[rank] = local)
  float* out;                               // local: sum over ranks
  const uint8_t* live4;
  i64 n4;
  TrainCtrl* c;
  int rank, world;
};

__device__ __forceinline__ void p2p_store_flag(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned p2p_load_flag(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 p2p_load4(const float* p) {   // never served from a stale L1 line of an earlier step
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// waits until flags[base + q] >= epoch for all q < world (one thread); false on time-out
__device__ __forceinline__ bool p2p_wait_all(const unsigned* flags, int base, int world, unsigned epoch) {
  const long long t0 = clock64();
  for (int q = 0; q < world; ++q) {
    unsigned spins = 0;
    while ((int)(p2p_load_flag(flags + base + q) - epoch) < 0) {
      if ((++spins & 63u) == 0 && clock64() - t0 > 4000000000LL) {
        atomicCAS(&tc::g_tc_err, 0, 90 + base / P2P_MAXRANKS);
        return false;
      }
    }
  }
  return true;
}

__global__ void __launch_bounds__(P2P_THREADS) p2p_allreduce_gradnorm_kernel(const P2PArgs a) {
  GWN_PDL_ENTRY();
  __shared__ double red[P2P_THREADS / 32];
  __shared__ int last_block;
  unsigned* mine = a.flags[a.rank];
  const unsigned epoch = (unsigned)(a.c->step + 1);          // stable until the LAST block bumps it below
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    p2p_store_flag(a.flags[threadIdx.x] + a.rank, epoch);                       // READY[rank] on peer threadIdx.x
  }
  if (threadIdx.x == 0) p2p_wait_all(mine, 0, a.world, epoch);
  __syncthreads();
  double s = 0.0;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += (i64)gridDim.x * blockDim.x) {
    float4 v[P2P_MAXRANKS];       // all peers' loads in flight before the first use (one NVLink round trip, not `world`)
#pragma unroll
    for (int q = 0; q < P2P_MAXRANKS; ++q)
      if (q < a.world) v[q] = p2p_load4(a.grad[q] + 4 * i);
    float4 acc = v[0];
#pragma unroll
    for (int q = 1; q < P2P_MAXRANKS; ++q)
      if (q < a.world) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }   // rank order on every rank
    st4(a.out + 4 * i, acc);
    if (a.live4[i]) s += (double)acc.x * acc.x + (double)acc.y * acc.y + (double)acc.z * acc.z + (double)acc.w * acc.w;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < P2P_THREADS / 32; ++w) t += red[w];
    if (t != 0.0) atomicAdd(&a.c->acc[4], t);
    __threadfence();
    unsigned* counter = mine + 2 * P2P_MAXRANKS;
    last_block = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
    if (last_block) *counter = 0u;
  }
  __syncthreads();
  if (last_block) {       // every block of this rank has finished reading the peers
    if (threadIdx.x < a.world) {
      __threadfence_system();
      p2p_store_flag(a.flags[threadIdx.x] + P2P_MAXRANKS + a.rank, epoch);      // DONE[rank] on peer threadIdx.x
    }
    if (threadIdx.x == 0) {
      a.c->step += 1;                                                           // what gradnorm_kernel does in the one-GPU step
      p2p_wait_all(mine, P2P_MAXRANKS, a.world, epoch);
    }
  }
}

}  // namespace gwn
#endif
