// SPDX-License-Identifier: Apache-2.0
// Micro-benchmark (bring-up only): can the per-SM row-gather rate be raised above the ~27 B/cycle
// LDGSTS limit by splitting every stage between the LSU path (LDGSTS) and the TMA unit
// (tile::gather4), which tools/l2sm_bench.cu showed to be independent ingress paths?
//   tma_rows = rows of the 128-row stage fetched by TMA gather4 (0 = all LDGSTS, 128 = all TMA)
//   TW       = warps that issue the gather4 requests (one request per lane)
//   nosw     = gather4 with a 256-byte no-swizzle box (per-request vs per-byte limit; layout
//              unusable by the MMA, measurement only)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hybrid_bench hybrid_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../warpconvnet_b200/csrc/common.cuh"
namespace wcn { void count_launch() {} }
using namespace wcn;

constexpr int kStages = 4;
constexpr int kStageBytes = 32768;  // 128 rows x 256 B as two 16 KB 128B-swizzled slabs
struct Ctrl { uint64_t full[kStages]; };

// warps 0-3: LDGSTS rows [0, nl) ; warps 4..4+TW-1: gather4 rows [nl, 128)
template <int TW, int NOSW>
__global__ void __launch_bounds__(128 + 32 * TW, 1)
hybrid_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ feats,
              long long ld_bytes, const int* __restrict__ idx, int iters, int frac8,
              long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                       kStages * kStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = frac8 * 16;  // rows by TMA
  const int nl = 128 - nt;    // rows by LDGSTS
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s)
      mbar_init(smem_u32(&ctrl->full[s]), (nl > 0 ? 128 : 0) + (nt > 0 ? 1 : 0));
    fence_mbar_init();
  }
  __syncthreads();
  const int* my_idx = idx + (size_t)blockIdx.x * iters * 128;
  // indices are prefetched one iteration ahead with coalesced loads (LDGSTS warps: row = the
  // lane-th row this warp serves; TMA warps: one int4 = one gather4 request per lane)
  const int half = lane >> 4, u = lane & 15, ch = u >> 3, c8 = u & 7;
  const int n_instr = nl / 2;                   // 2-row LDGSTS instructions per stage
  const int per_warp = (n_instr + 3) / 4;       // instructions of this warp (rows 2*per_warp)
  const int row0 = warp * 2 * per_warp;         // first LDGSTS row of this warp
  const int n_req = nt / 4;
  const int tw = warp - 4;
  const int req = tw * 32 + lane;               // this lane's request (TMA warps), valid < n_req
  int nxt = 0;
  int4 nxt4 = make_int4(0, 0, 0, 0);
  auto prefetch = [&](int it) {
    if (it >= iters) return;
    if (warp < 4) { if (lane < 2 * per_warp && row0 + lane < nl) nxt = __ldg(my_idx + (size_t)it * 128 + row0 + lane); }
    else if (req < n_req) nxt4 = *reinterpret_cast<const int4*>(my_idx + (size_t)it * 128 + nl + 4 * req);
  };
  prefetch(0);
  const long long t0 = clock64();
  for (int it = 0; it < iters + kStages - 1; ++it) {
    if (it < iters) {
      const int stage = it % kStages;
      const uint32_t a_smem = smem_base + stage * kStageBytes;
      const uint32_t bar = smem_u32(&ctrl->full[stage]);
      const int my = nxt;
      const int4 r = nxt4;
      prefetch(it + 1);
      if (warp < 4) {
        if (nl > 0) {
          for (int i = 0; i < per_warp; ++i) {
            const int rr = __shfl_sync(0xffffffffu, my, 2 * i + half);
            const int row = row0 + 2 * i + half;
            const uint8_t* src = feats + (long long)rr * ld_bytes + u * 16;
            if (row < nl) cp_async_16(a_smem + ch * 16384 + sw128_offset(row, c8), src, 16);
          }
          cp_async_mbar_arrive_noinc(bar);
        }
      } else if (nt > 0) {
        if (tw == 0 && lane == 0) mbar_arrive_expect_tx(bar, nt * 256);
        __syncwarp();
        if (req < n_req) {
          const int row = nl + 4 * req;
          if (NOSW) {
            tma_gather4(a_smem + row * 256, &tmap, 0, r.x, r.y, r.z, r.w, bar);
          } else {
            tma_gather4(a_smem + row * 128, &tmap, 0, r.x, r.y, r.z, r.w, bar);
            tma_gather4(a_smem + 16384 + row * 128, &tmap, 64, r.x, r.y, r.z, r.w, bar);
          }
        }
      }
    }
    const int wit = it - (kStages - 1);
    if (wit >= 0) mbar_wait(smem_u32(&ctrl->full[wit % kStages]), (wit / kStages) & 1);
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                             const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int TW, int NOSW>
void run(const CUtensorMap& tmap, const uint8_t* feats, long long ld, const int* idx, int iters,
         int frac8, long long* d_cycles) {
  const size_t smem = kStages * kStageBytes + sizeof(Ctrl) + 1024;
  auto kern = hybrid_kernel<TW, NOSW>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a);
    kern<<<148, 128 + 32 * TW, smem>>>(tmap, feats, ld, idx, iters, frac8, d_cycles);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148); cudaMemcpy(h.data(), d_cycles, 148 * 8, cudaMemcpyDeviceToHost);
  double mc = 0; for (auto v : h) mc += v; mc /= 148;
  const double bytes = 148.0 * iters * kStageBytes;
  printf("tma_rows=%3d/128 TW=%d nosw=%d: %8.1f us  %6.2f TB/s  %6.1f B/cyc/SM %s\n", frac8 * 16, TW,
         NOSW, best * 1e3, bytes / (best * 1e-3) / 1e12, (double)iters * kStageBytes / mc,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const int n_rows = 200704, C = 128, iters = 400;
  const long long ld = C * 2;
  uint8_t* feats; cudaMalloc(&feats, (size_t)n_rows * ld); cudaMemset(feats, 1, (size_t)n_rows * ld);
  long long* d_cycles; cudaMalloc(&d_cycles, 148 * 8);
  int* d_idx; cudaMalloc(&d_idx, (size_t)148 * iters * 128 * 4);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  CUtensorMap tmap, tmap_nosw;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)n_rows}; cuuint64_t gstr[1] = {(cuuint64_t)ld};
  cuuint32_t box[2] = {64, 1}, box2[2] = {128, 1}, estr[2] = {1, 1};
  CUresult r1 = ((EncodeFn)sym)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, feats, gdim, gstr, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = ((EncodeFn)sym)(&tmap_nosw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, feats, gdim, gstr,
                                box2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("tensor maps: %d %d\n", (int)r1, (int)r2);
  { void* big; cudaMalloc(&big, 1 << 30); for (int i = 0; i < 100; ++i) cudaMemset(big, i, 1 << 30); cudaDeviceSynchronize(); cudaFree(big); }
  std::vector<int> h((size_t)148 * iters * 128);
  srand(1);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (int)(((long long)rand() * 7919 + rand()) % n_rows);
  cudaMemcpy(d_idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  for (int f = 0; f <= 8; ++f) run<1, 0>(tmap, feats, ld, d_idx, iters, f, d_cycles);
  for (int f : {2, 3, 4, 8}) run<2, 0>(tmap, feats, ld, d_idx, iters, f, d_cycles);
  for (int f : {3, 4, 8}) run<4, 0>(tmap, feats, ld, d_idx, iters, f, d_cycles);
  for (int f : {4, 8}) run<1, 1>(tmap_nosw, feats, ld, d_idx, iters, f, d_cycles);
  for (int f : {4, 8}) run<4, 1>(tmap_nosw, feats, ld, d_idx, iters, f, d_cycles);
  return 0;
}
