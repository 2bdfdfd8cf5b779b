// SPDX-License-Identifier: Apache-2.0
// Micro-benchmark (bring-up only): does the per-SM rate of 256-byte row gathers (cp.async 16 B per
// lane into 128B-swizzled shared memory, as the producers of conv_fwd.cu / conv_wgrad.cu issue them)
// depend on HOW MANY WARPS issue them? NW = 4 / 8 / 16 gather warps share every 128-row stage;
// 4 stages of 32 KB, random rows of a 51 MB matrix, 148 CTAs (1 per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_warps_bench gather_warps_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../warpconvnet_b200/csrc/common.cuh"
namespace wcn { void count_launch() {} }
using namespace wcn;

constexpr int kStages = 4;
constexpr int kABytes = 32768;  // 128 rows x 256 B

template <int NW, int ROWS_PER_STAGE>
__global__ void __launch_bounds__(NW * 32, 1)
gw_kernel(const uint8_t* __restrict__ feats, long long ld_bytes, const int* __restrict__ idx,
          int iters, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int kStageBytes = ROWS_PER_STAGE * 256;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                               kStages * kStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&full[s]), NW * 32);
    fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  constexpr int kRowsPerWarp = ROWS_PER_STAGE / NW;   // rows of a stage per warp
  constexpr int kInstr = kRowsPerWarp / 2;            // 2 rows (2 x 16 lanes) per instruction
  const int* my_idx = idx + (size_t)blockIdx.x * iters * ROWS_PER_STAGE;
  const int half = lane >> 4, u = lane & 15, ch = u >> 3, c8 = u & 7;
  for (int it = 0; it < iters + kStages - 1; ++it) {
    if (it < iters) {
      const int stage = it % kStages;
      const uint32_t a_smem = smem_base + stage * kStageBytes;
      const uint32_t bar = smem_u32(&full[stage]);
      const int my = lane < kRowsPerWarp
                         ? __ldg(my_idx + (size_t)it * ROWS_PER_STAGE + warp * kRowsPerWarp + lane) : 0;
#pragma unroll
      for (int q = 0; q < kInstr; ++q) {
        const int r = __shfl_sync(0xffffffffu, my, 2 * q + half);
        const uint32_t row = warp * kRowsPerWarp + 2 * q + half;
        const uint8_t* src = feats + (long long)r * ld_bytes + u * 16;
        cp_async_16(a_smem + ch * (ROWS_PER_STAGE * 128) + sw128_offset(row, c8), src, 16);
      }
      cp_async_mbar_arrive_noinc(bar);
    }
    const int wit = it - (kStages - 1);
    if (wit >= 0) mbar_wait(smem_u32(&full[wit % kStages]), (wit / kStages) & 1);
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int NW, int ROWS>
void run(const uint8_t* feats, long long ld, const int* idx, int iters, long long* d_cycles) {
  const size_t smem = kStages * ROWS * 256 + 64 + 1024;
  auto kern = gw_kernel<NW, ROWS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a);
    kern<<<148, NW * 32, smem>>>(feats, ld, idx, iters, d_cycles);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148); cudaMemcpy(h.data(), d_cycles, 148 * 8, cudaMemcpyDeviceToHost);
  double mc = 0; for (auto v : h) mc += v; mc /= 148;
  const double per_cta = (double)iters * ROWS * 256;
  printf("warps=%2d rows/stage=%3d (%d stages, %3d KB in flight): %8.1f us  %6.2f TB/s  %6.1f B/cyc/SM %s\n",
         NW, ROWS, kStages, (kStages - 1) * ROWS * 256 / 1024, best * 1e3,
         per_cta * 148 / (best * 1e-3) / 1e12, per_cta / mc, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const int n_rows = 200704, iters = 300;
  const long long ld = 256;
  uint8_t* feats; cudaMalloc(&feats, (size_t)n_rows * ld); cudaMemset(feats, 1, (size_t)n_rows * ld);
  long long* d_cycles; cudaMalloc(&d_cycles, 148 * 8);
  int* d_idx; cudaMalloc(&d_idx, (size_t)148 * iters * 256 * 4);
  std::vector<int> h((size_t)148 * iters * 256);
  srand(1);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (int)(((long long)rand() * 7919 + rand()) % n_rows);
  cudaMemcpy(d_idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  run<4, 128>(feats, ld, d_idx, iters, d_cycles);
  run<8, 128>(feats, ld, d_idx, iters, d_cycles);
  run<16, 128>(feats, ld, d_idx, iters, d_cycles);
  run<4, 256>(feats, ld, d_idx, iters / 2, d_cycles);
  run<8, 256>(feats, ld, d_idx, iters / 2, d_cycles);
  run<16, 256>(feats, ld, d_idx, iters / 2, d_cycles);
  return 0;
}
