// SPDX-License-Identifier: Apache-2.0
// Micro-benchmark (bring-up only): is the L2->SM ingress limit seen by the forward kernel a per-SM
// port limit or a chip-wide one, and does TMA multicast of the weight slice escape it?
//   G  : 4 warps gather 256-byte rows with LDGSTS (as conv_fwd.cu's producers)
//   B  : one thread streams 32 KB weight slices with cp.async.bulk
//   GB : both at once (the forward kernel's traffic mix)
//   M<c>: like B but the slice is multicast across a cluster of c CTAs (each CTA issues 1/c)
//   GM<c>: G + M<c>
// Each variant runs on 148 / 74 / 32 CTAs (1 CTA per SM) to separate per-SM from chip-wide caps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2sm_bench l2sm_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../warpconvnet_b200/csrc/common.cuh"
namespace wcn { void count_launch() {} }
using namespace wcn;

constexpr int kStages = 3;
constexpr int kABytes = 32768;  // 128 rows x 256 B
constexpr int kBBytes = 32768;  // one 128x128 bf16 weight slice

struct Ctrl { uint64_t fullA[kStages]; uint64_t fullB[kStages]; };

__device__ __forceinline__ void bulk_copy_mc(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}

// GATHER: 0/1, BULK: 0 none, 1 unicast, c>1 multicast over a cluster of c
template <int GATHER, int BULK>
__global__ void __launch_bounds__(160, 1)
l2sm_kernel(const uint8_t* __restrict__ feats, long long ld_bytes, const int* __restrict__ idx,
            const uint8_t* __restrict__ wimg, int n_slices, int iters,
            long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                       kStages * (kABytes + kBBytes));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&ctrl->fullA[s]), 128);
      mbar_init(smem_u32(&ctrl->fullB[s]), 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (BULK > 1) cluster_sync();
  const long long t0 = clock64();
  if (warp < 4) {
    if (GATHER) {
      const int* my_idx = idx + (size_t)blockIdx.x * iters * 128;
      const int half = lane >> 4, u = lane & 15, ch = u >> 3, c8 = u & 7;
      for (int it = 0; it < iters + kStages - 1; ++it) {
        if (it < iters) {
          const int stage = it % kStages;
          const uint32_t a_smem = smem_base + stage * (kABytes + kBBytes);
          const uint32_t bar = smem_u32(&ctrl->fullA[stage]);
          const int my = __ldg(my_idx + (size_t)it * 128 + warp * 32 + lane);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int r = __shfl_sync(0xffffffffu, my, 2 * q + half);
            const uint32_t row = warp * 32 + 2 * q + half;
            const uint8_t* src = feats + (long long)r * ld_bytes + u * 16;
            cp_async_16(a_smem + ch * 16384 + sw128_offset(row, c8), src, 16);
          }
          cp_async_mbar_arrive_noinc(bar);
        }
        const int wit = it - (kStages - 1);
        if (wit >= 0) mbar_wait(smem_u32(&ctrl->fullA[wit % kStages]), (wit / kStages) & 1);
      }
    }
  } else if (lane == 0) {
    if (BULK) {
      const uint32_t rank = BULK > 1 ? cluster_rank() : 0;
      const int cid = BULK > 1 ? blockIdx.x / BULK : blockIdx.x;
      for (int it = 0; it < iters + kStages - 1; ++it) {
        if (it < iters) {
          const int stage = it % kStages;
          const uint32_t b_smem = smem_base + stage * (kABytes + kBBytes) + kABytes;
          const uint32_t bar = smem_u32(&ctrl->fullB[stage]);
          const uint8_t* src = wimg + (size_t)((it * 7 + cid * 3) % n_slices) * kBBytes;
          mbar_arrive_expect_tx(bar, kBBytes);
          if (BULK == 1) {
            bulk_copy_g2s(b_smem, src, kBBytes, bar);
          } else {
            constexpr int part = kBBytes / (BULK > 0 ? BULK : 1);
            bulk_copy_mc(b_smem + rank * part, src + rank * part, part, bar,
                         (uint16_t)((1u << BULK) - 1));
          }
        }
        const int wit = it - (kStages - 1);
        if (wit >= 0) mbar_wait(smem_u32(&ctrl->fullB[wit % kStages]), (wit / kStages) & 1);
      }
    }
  }
  __syncthreads();
  if (BULK > 1) cluster_sync();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int GATHER, int BULK>
void run(const char* name, int ctas, const uint8_t* feats, long long ld, const int* idx,
         const uint8_t* wimg, int iters, long long* d_cycles) {
  const size_t smem = kStages * (kABytes + kBBytes) + sizeof(Ctrl) + 1024;
  auto kern = l2sm_kernel<GATHER, BULK>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (BULK > 1) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  cudaError_t le = cudaSuccess;
  for (int rep = 0; rep < 4; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(160); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = BULK > 1 ? BULK : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEventRecord(a);
    le = cudaLaunchKernelEx(&cfg, kern, feats, ld, idx, wimg, 27, iters, d_cycles);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(ctas); cudaMemcpy(h.data(), d_cycles, ctas * 8, cudaMemcpyDeviceToHost);
  double mc = 0; for (auto v : h) mc += v; mc /= ctas;
  const double per_cta = (double)iters * ((GATHER ? kABytes : 0) + (BULK ? kBBytes : 0));
  printf("%-10s ctas=%3d: %8.1f us  %6.2f TB/s into SMs  %6.1f B/cyc/SM  (clk %.2f GHz) %s %s\n", name,
         ctas, best * 1e3, per_cta * ctas / (best * 1e-3) / 1e12, per_cta / mc,
         mc / (best * 1e-3) / 1e9, le == cudaSuccess ? "" : cudaGetErrorString(le),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const int n_rows = 200704, C = 128, iters = 300;
  const long long ld = C * 2;
  uint8_t* feats; cudaMalloc(&feats, (size_t)n_rows * ld); cudaMemset(feats, 1, (size_t)n_rows * ld);
  uint8_t* wimg; cudaMalloc(&wimg, (size_t)27 * kBBytes); cudaMemset(wimg, 1, (size_t)27 * kBBytes);
  long long* d_cycles; cudaMalloc(&d_cycles, 148 * 8);
  int* d_idx; cudaMalloc(&d_idx, (size_t)148 * iters * 128 * 4);
  { void* big; cudaMalloc(&big, 1 << 30); for (int i = 0; i < 100; ++i) cudaMemset(big, i, 1 << 30); cudaDeviceSynchronize(); cudaFree(big); }
  for (int order = 0; order < 2; ++order) {
    std::vector<int> h((size_t)148 * iters * 128);
    srand(1);
    for (size_t i = 0; i < h.size(); ++i)
      h[i] = order == 0 ? (int)(((long long)rand() * 7919 + rand()) % n_rows) : (int)(i % n_rows);
    cudaMemcpy(d_idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    printf("--- row order: %s\n", order == 0 ? "random" : "sequential (coalesced stream)");
    for (int ctas : {148, 72, 32}) {
      run<1, 0>("G", ctas, feats, ld, d_idx, wimg, iters, d_cycles);
      if (order == 0) {
        run<0, 1>("B", ctas, feats, ld, d_idx, wimg, iters, d_cycles);
        run<0, 2>("M2", ctas, feats, ld, d_idx, wimg, iters, d_cycles);
        run<0, 4>("M4", ctas, feats, ld, d_idx, wimg, iters, d_cycles);
        run<0, 8>("M8", ctas / 8 * 8, feats, ld, d_idx, wimg, iters, d_cycles);
      }
      run<1, 1>("GB", ctas, feats, ld, d_idx, wimg, iters, d_cycles);
      run<1, 2>("GM2", ctas, feats, ld, d_idx, wimg, iters, d_cycles);
      run<1, 4>("GM4", ctas, feats, ld, d_idx, wimg, iters, d_cycles);
    }
  }
  return 0;
}
