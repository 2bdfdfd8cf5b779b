// SPDX-License-Identifier: Apache-2.0
// All-reduce (scaled sum, fp32, in place) of the weight gradients over NVLink / NVSwitch PEER MEMORY: the
// only collective of the sparse-conv path. The wgrad kernel reduces its tiles
// straight into a buffer that every rank of the box has mapped (symmetric memory); this kernel
// then finishes the job without NCCL:
//
//   barrier A (every rank's local sums are complete)
//   two-shot reduce: rank r owns slice r of the buffer — it reads that slice from ALL peers over
//     NVLink, adds, and stores the total back into slice r of EVERY peer's buffer
//   barrier B (every slice of every buffer holds the total)
//
// For the 1.77 MB dW of one 128 -> 128 layer that is 2 x 7/8 x 1.77 MB = 3.1 MB per rank over links
// of 900 GB/s per direction: the cost is the two barriers (flag writes into the peers' memory), not
// bandwidth — which is why it beats a ring / tree collective whose latency grows with its step
// count. A few small CTAs only (default 32 x 128 threads), so it runs next to the dgrad kernel it overlaps with.
//
// Barriers: flags[cta][source rank] live in each rank's symmetric flag buffer; a rank signals by
// storing the CTA's epoch (a device-side counter, so a captured CUDA graph can be replayed) into
// every peer's flags with st.release.sys and waits until all of its own flags reach that epoch.
// Every CTA index forms its own channel across the ranks: no grid-wide synchronisation inside a GPU.
//
// The reference has no collective code at all (users wrap DDP).
#include "common.cuh"

namespace wcn {

constexpr int kArMaxRanks = 16;
constexpr int kArMaxCtas = 128;
// 128 threads (ONE warp per SM sub-partition) x <= 128 registers, no shared memory: a CTA of this
// kernel fits on an SM NEXT TO a resident CTA of the gather-GEMM kernel — 9 warps x 128 registers,
// i.e. 3 warps = 12K of the 16K registers of one sub-partition, ~200 KB of shared memory — so the
// reduction really runs under dgrad instead of waiting for its persistent CTAs to retire (measured,
// profiles/r2t_peer_allreduce.md: with 256 threads x 104 registers the CTAs did NOT co-reside and
// the dgrad CTAs queued behind them). A library collective with its own shared-memory and register
// footprint has to wait.
constexpr int kArThreads = 128;
// flag buffer of one rank (uint32): [kArMaxCtas][kArMaxRanks] flags, [kArMaxCtas] epochs, then one
// word counting barrier waits that gave up (a peer that never arrives must not hang this GPU)
constexpr int kArFlagWords = kArMaxCtas * kArMaxRanks + kArMaxCtas + 1;
constexpr int kArTimeoutWord = kArMaxCtas * kArMaxRanks + kArMaxCtas;
constexpr unsigned long long kArTimeoutNs = 10ull * 1000 * 1000 * 1000;

struct PeerAllReduceParams {
  float* bufs[kArMaxRanks];        // the same buffer on every rank (peer-mapped pointers)
  unsigned* flags[kArMaxRanks];    // the flag buffer of every rank
  long long n;                     // floats, multiple of 4
  float scale;                     // result = scale * sum (1 / world: the mean)
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 v;  // relaxed system-scope load: never served from a stale local cache line
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_v4(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// all ranks meet: CTA channel `b`, epoch value `e`
__device__ __forceinline__ void peer_barrier(const PeerAllReduceParams& p, int b, unsigned e) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    st_release_sys(p.flags[peer] + b * kArMaxRanks + p.rank, e);
    const unsigned* mine = p.flags[p.rank] + b * kArMaxRanks + peer;
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      if ((++spins & 0xfffu) == 0) {  // every 4096 polls: give up after 10 s, and say so
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (now - t0 > kArTimeoutNs) {
          atomicAdd(p.flags[p.rank] + kArTimeoutWord, 1u);
          break;
        }
      }
    }
  }
  __syncthreads();
}

// W = compile-time world size (all W x U loads of a thread are in flight before the first add:
// one NVLink round trip per pass, not one per peer), 0 = any world size
template <int W, int U>
__global__ void __maxnreg__(128)
peer_allreduce_kernel(const __grid_constant__ PeerAllReduceParams p) {
  pdl_begin();
  const int b = blockIdx.x, nb = gridDim.x;
  const int world = W > 0 ? W : p.world;
  unsigned* epoch_word = p.flags[p.rank] + kArMaxCtas * kArMaxRanks + b;
  const unsigned e0 = *epoch_word;  // advanced by 2 per call, identically on every rank
  peer_barrier(p, b, e0 + 1u);
  // slice of this rank, sub-slice of this CTA (float4 granularity)
  const long long n4 = p.n / 4;
  const long long r0 = n4 * p.rank / world, r1 = n4 * (p.rank + 1) / world;
  const long long c0 = r0 + (r1 - r0) * b / nb, c1 = r0 + (r1 - r0) * (b + 1) / nb;
  if constexpr (W > 0) {
    for (long long base = c0; base < c1; base += (long long)kArThreads * U) {
      float4 v[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = base + u * kArThreads + threadIdx.x;
#pragma unroll
        for (int k = 0; k < W; ++k)
          if (i < c1) v[u][k] = ld_peer_v4(p.bufs[(p.rank + k) % W] + 4 * i);  // home first
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = base + u * kArThreads + threadIdx.x;
        if (i >= c1) continue;
        float4 acc = v[u][0];
#pragma unroll
        for (int k = 1; k < W; ++k) {
          acc.x += v[u][k].x; acc.y += v[u][k].y; acc.z += v[u][k].z; acc.w += v[u][k].w;
        }
        acc.x *= p.scale; acc.y *= p.scale; acc.z *= p.scale; acc.w *= p.scale;
#pragma unroll
        for (int k = 0; k < W; ++k) st_peer_v4(p.bufs[(p.rank + k) % W] + 4 * i, acc);
      }
    }
  } else {
    for (long long i = c0 + threadIdx.x; i < c1; i += kArThreads) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < world; ++k) {
        const float4 v = ld_peer_v4(p.bufs[(p.rank + k) % world] + 4 * i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x *= p.scale; acc.y *= p.scale; acc.z *= p.scale; acc.w *= p.scale;
      for (int k = 0; k < world; ++k) st_peer_v4(p.bufs[(p.rank + k) % world] + 4 * i, acc);
    }
  }
  peer_barrier(p, b, e0 + 2u);
  if (threadIdx.x == 0) *epoch_word = e0 + 2u;
}

int peer_allreduce_f32(void* const* bufs, void* const* flags, int rank, int world, long long n,
                       float scale, int n_ctas, cudaStream_t s) {
  if (world < 1 || world > kArMaxRanks || rank < 0 || rank >= world || n < 0 || (n & 3) != 0)
    return kErrInvalidArg;
  if (n == 0 || world == 1) return kOk;
  if (n_ctas < 1) n_ctas = 32;
  if (n_ctas > kArMaxCtas) n_ctas = kArMaxCtas;
  PeerAllReduceParams p;
  for (int i = 0; i < world; ++i) {
    if (bufs[i] == nullptr || flags[i] == nullptr) return kErrInvalidArg;
    if (reinterpret_cast<uintptr_t>(bufs[i]) & 15) return kErrAlignment;
    p.bufs[i] = static_cast<float*>(bufs[i]);
    p.flags[i] = static_cast<unsigned*>(flags[i]);
  }
  p.n = n;
  p.scale = scale;
  p.rank = rank;
  p.world = world;
  const dim3 grid(n_ctas), block(kArThreads);
  switch (world) {
    case 2: wcn_launch(peer_allreduce_kernel<2, 8>, grid, block, 0, s, p); break;
    case 4: wcn_launch(peer_allreduce_kernel<4, 4>, grid, block, 0, s, p); break;
    case 8: wcn_launch(peer_allreduce_kernel<8, 2>, grid, block, 0, s, p); break;
    default: wcn_launch(peer_allreduce_kernel<0, 1>, grid, block, 0, s, p); break;
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
}

int peer_allreduce_flag_words() { return kArFlagWords; }
int peer_allreduce_timeout_word() { return kArTimeoutWord; }

}  // namespace wcn
