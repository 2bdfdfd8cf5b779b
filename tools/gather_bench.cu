// SPDX-License-Identifier: Apache-2.0
// Micro-benchmark (bring-up only): how fast can one SM pull gathered 128-byte row segments from
// L2 into 128B-swizzled shared memory? Variants: LDGSTS (cp.async 16 B), LDG.128 + STS.128,
// TMA gather4, with 4 / 8 producer warps. No consumer: every stage is just waited for.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../warpconvnet_b200/csrc/common.cuh"
namespace wcn { void count_launch() {} }
using namespace wcn;

constexpr int kStages = 4;
constexpr int kStageBytes = 16384;  // 128 rows x 128 B (modes >= 4 use two such slabs per stage)

struct Ctrl { uint64_t full[kStages]; };

// mode 0: LDGSTS cg, 1: LDGSTS ca, 2: LDG+STS, 3: TMA gather4,
// 4/5: LDGSTS cg/ca whole 256-byte rows (both chunks per stage), 6: TMA gather4 both chunks
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
gather_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ feats,
              long long ld_bytes, int n_rows, const int* __restrict__ idx, int iters,
              long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int SB = (MODE >= 4) ? 2 * kStageBytes : kStageBytes;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem_raw + (smem_base - smem_u32(smem_raw)) + kStages * SB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int T = WARPS * 32;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s)
      mbar_init(smem_u32(&ctrl->full[s]), (MODE == 3 || MODE == 6) ? 1 : T);
    fence_mbar_init();
  }
  __syncthreads();
  const int* my_idx = idx + (size_t)blockIdx.x * iters * 128;
  const int c16 = lane & 7, sub = lane >> 3;
  constexpr int RPW = 128 / WARPS;   // rows per warp per stage
  constexpr int Q = RPW / 4;         // LDGSTS per thread per stage
  const long long t0 = clock64();
  int stage = 0; uint32_t phase = 0;
  // software pipeline: issue stage i, wait for stage i - (kStages-1)
  for (int it = 0; it < iters + kStages - 1; ++it) {
    if (it < iters) {
      const uint32_t a_smem = smem_base + stage * SB;
      const uint32_t bar = smem_u32(&ctrl->full[stage]);
      const int chunk = it & 1;
      if (MODE == 6) {
        if (warp == 0 && lane == 0) mbar_arrive_expect_tx(bar, 2 * kStageBytes);
        if (lane < RPW / 4) {
          const int4 r = *reinterpret_cast<const int4*>(my_idx + (size_t)it * 128 + warp * RPW + lane * 4);
          tma_gather4(a_smem + (warp * RPW + lane * 4) * 128, &tmap, 0, r.x, r.y, r.z, r.w, bar);
          tma_gather4(a_smem + kStageBytes + (warp * RPW + lane * 4) * 128, &tmap, 64, r.x, r.y, r.z, r.w, bar);
        }
      } else if (MODE == 4 || MODE == 5) {
        const int my = (lane < RPW) ? __ldg(my_idx + (size_t)it * 128 + warp * RPW + lane) : 0;
        const int half = lane >> 4, u = lane & 15, ch = u >> 3, c8 = u & 7;
#pragma unroll
        for (int q = 0; q < RPW / 2; ++q) {
          const int r = __shfl_sync(0xffffffffu, my, 2 * q + half);
          const uint32_t row = warp * RPW + 2 * q + half;
          const uint8_t* src = feats + (long long)r * ld_bytes + u * 16;
          const uint32_t dst = a_smem + ch * kStageBytes + sw128_offset(row, c8);
          if (MODE == 4) cp_async_16(dst, src, 16); else cp_async_16_ca(dst, src, 16);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
      } else if (MODE == 3) {
        // each warp covers RPW rows = RPW/4 gather4 requests issued by lanes < RPW/4
        if (warp == 0 && lane == 0) mbar_arrive_expect_tx(bar, kStageBytes);
        if (lane < RPW / 4) {
          const int4 r = *reinterpret_cast<const int4*>(my_idx + (size_t)it * 128 + warp * RPW + lane * 4);
          tma_gather4(a_smem + (warp * RPW + lane * 4) * 128, &tmap, chunk * 64, r.x, r.y, r.z, r.w, bar);
        }
      } else {
        const int my = (lane < RPW) ? __ldg(my_idx + (size_t)it * 128 + warp * RPW + lane) : 0;
        if (MODE == 2) {
          uint4 v[Q];
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const int r = __shfl_sync(0xffffffffu, my, 4 * q + sub);
            v[q] = __ldg(reinterpret_cast<const uint4*>(feats + (long long)r * ld_bytes + chunk * 128 + c16 * 16));
          }
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const uint32_t row = warp * RPW + 4 * q + sub;
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_smem + sw128_offset(row, c16)),
                         "r"(v[q].x), "r"(v[q].y), "r"(v[q].z), "r"(v[q].w) : "memory");
          }
          mbar_arrive(bar);
        } else {
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const int r = __shfl_sync(0xffffffffu, my, 4 * q + sub);
            const uint32_t row = warp * RPW + 4 * q + sub;
            const uint8_t* src = feats + (long long)r * ld_bytes + chunk * 128 + c16 * 16;
            if (MODE == 0) cp_async_16(a_smem + sw128_offset(row, c16), src, 16);
            else cp_async_16_ca(a_smem + sw128_offset(row, c16), src, 16);
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
        }
      }
    }
    // wait for the stage issued kStages-1 iterations ago (so kStages-1 stages are in flight)
    const int wit = it - (kStages - 1);
    if (wit >= 0) {
      const int ws = wit % kStages;
      const uint32_t wphase = (wit / kStages) & 1;
      mbar_wait(smem_u32(&ctrl->full[ws]), wphase);
    }
    if (++stage == kStages) { stage = 0; phase ^= 1u; }
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int WARPS>
void run(const char* name, const CUtensorMap& tmap, const uint8_t* feats, long long ld, int n_rows,
         const int* idx, int iters, long long* d_cycles, int order) {
  const size_t smem = kStages * (MODE >= 4 ? 2 : 1) * kStageBytes + sizeof(Ctrl) + 1024;
  cudaFuncSetAttribute(gather_kernel<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    gather_kernel<MODE, WARPS><<<148, WARPS * 32, smem>>>(tmap, feats, ld, n_rows, idx, iters, d_cycles);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148); cudaMemcpy(h.data(), d_cycles, 148 * 8, cudaMemcpyDeviceToHost);
  double mc = 0; for (auto v : h) mc += v; mc /= 148;
  const double sbytes = (MODE >= 4 ? 2.0 : 1.0) * kStageBytes;
  const double bytes = 148.0 * iters * sbytes;
  printf("%-28s order=%d: %8.1f us  %7.2f TB/s  %6.1f B/cyc/SM  (%.0f cyc/stage, clk %.2f GHz) %s\n", name, order,
         best * 1e3, bytes / (best * 1e-3) / 1e12, (double)iters * sbytes / mc, mc / iters,
         mc / (best * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const int n_rows = 200704, C = 128, iters = 400;
  const long long ld = C * 2;
  uint8_t* feats; cudaMalloc(&feats, (size_t)n_rows * ld); cudaMemset(feats, 1, (size_t)n_rows * ld);
  long long* d_cycles; cudaMalloc(&d_cycles, 148 * 8);
  int* d_idx; cudaMalloc(&d_idx, (size_t)148 * iters * 128 * 4);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)n_rows}; cuuint64_t gstr[1] = {(cuuint64_t)ld};
  cuuint32_t box[2] = {64, 1}, estr[2] = {1, 1};
  ((EncodeFn)sym)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, feats, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  // warm the clocks
  { void* big; cudaMalloc(&big, 1 << 30); for (int i = 0; i < 200; ++i) cudaMemset(big, i, 1 << 30); cudaDeviceSynchronize(); cudaFree(big); }
  for (int order = 2; order < 5; ++order) {
    // order 0: random rows; order 1: locally coherent rows (neighbouring indices, like a surface)
    std::vector<int> h((size_t)148 * iters * 128);
    srand(1);
    for (size_t i = 0; i < h.size(); ++i) {
      if (order == 0) h[i] = (int)(((long long)rand() * 7919 + rand()) % n_rows);
      else if (order == 1) h[i] = (int)((i * 3 + (rand() % 900)) % n_rows);
      else { const int run = order == 2 ? 2 : (order == 3 ? 4 : 16); static int base = 0; if (i % run == 0) base = (int)(((long long)rand() * 7919 + rand()) % (n_rows - run)); h[i] = base + (int)(i % run); }
    }
    cudaMemcpy(d_idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    run<0, 4>("LDGSTS.cg 4 warps", tmap, feats, ld, n_rows, d_idx, iters, d_cycles, order);
    run<4, 4>("LDGSTS.cg 256B-row 4 warps", tmap, feats, ld, n_rows, d_idx, iters, d_cycles, order);
    run<4, 8>("LDGSTS.cg 256B-row 8 warps", tmap, feats, ld, n_rows, d_idx, iters, d_cycles, order);
    run<5, 4>("LDGSTS.ca 256B-row 4 warps", tmap, feats, ld, n_rows, d_idx, iters, d_cycles, order);
    run<6, 16>("TMA gather4x2 16 warps", tmap, feats, ld, n_rows, d_idx, iters, d_cycles, order);
  }
  return 0;
}
