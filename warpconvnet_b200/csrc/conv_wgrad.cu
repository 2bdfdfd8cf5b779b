// SPDX-License-Identifier: Apache-2.0
// wgrad AtB_gather_gather on tcgen05: dW_k[ci, co] += sum over pairs (i,o) of offset k of
// X[i, ci] * dY[o, co].
//
// Replaces (semantics, not code): warpconvnet/nn/functional/sparse_conv/detail/explicit.py:95-97
// and the production kernels csrc/mask_gemm/include/MaskGemm_wgrad_*.h (32-deep mma.sync
// contraction tiles + fp32 atomics) / detail/cute_grouped.py:282-432.
//
// The contraction runs over the pair list, so both operands are gathered rows that land in shared
// memory exactly as they sit in HBM (channels contiguous) = "MN-major" operands for the tensor
// core: A = X rows (M = cin), B = dY rows (N = cout), K = pairs. The CSR pair lists are exact, so
// no padded work is issued. A work unit is a fixed-size slice of one offset's pair list; its fp32
// 128 x cout result is reduced into dW with vectorised red.global.add.v4.f32.
//
// CTA layout identical to the forward kernel: warps 0-3 gather, warp 4 issues MMAs, warps 5-8
// drain TMEM (double buffered).
#include "common.cuh"
#include "conv_gemm.cuh"

namespace wcn {

constexpr int kWgThreads = 288;
constexpr int kWgMaxStages = 8;
constexpr int kWgTmemCols = 512;
constexpr int kWgAccStride = 256;
constexpr int kWgMaxK = 1024;  // offsets per kernel map supported by the in-kernel unit table

struct WgSmemCtrl {
  uint64_t full[kWgMaxStages];
  uint64_t empty[kWgMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  int total_units;
  int unit_start[kWgMaxK + 1];
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

template <typename T>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  constexpr int kElem = (int)sizeof(T);
  constexpr int kBlkElems = 128 / kElem;          // channels per 128-byte MN block
  constexpr int kPairs = (kElem == 2) ? 64 : 32;  // pairs (K extent) per pipeline stage
  constexpr int kKPerMma = 32 / kElem;            // 16 (bf16/f16) or 8 (tf32) pairs per MMA
  constexpr int kAStage = (128 / kBlkElems) * kPairs * 128;  // A always spans M = 128 channels
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int ys = blockIdx.y;  // Cin slab / group slab
  const int zs = blockIdx.z;  // Cout slab
  const int cin_here = (ys == (int)gridDim.y - 1) ? p.cin_last : p.cin;
  const int n_blk_a = (cin_here + kBlkElems - 1) / kBlkElems;
  const int n_blk_b = (p.cout + kBlkElems - 1) / kBlkElems;
  const int b_stage = n_blk_b * kPairs * 128;
  const int stage_bytes = kAStage + ((b_stage + 1023) & ~1023);
  const int stages = p.stages;
  WgSmemCtrl* ctrl = reinterpret_cast<WgSmemCtrl*>(smem_gen + (size_t)stages * stage_bytes);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&ctrl->full[s]), 128);
      mbar_init(smem_u32(&ctrl->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&ctrl->acc_full[a]), 1);
      mbar_init(smem_u32(&ctrl->acc_empty[a]), 128);
    }
    fence_mbar_init();
    int acc = 0;
    for (int k = 0; k < p.K; ++k) {
      ctrl->unit_start[k] = acc;
      const int len = p.offsets[k + 1] - p.offsets[k];
      acc += (len + p.unit_pairs - 1) / p.unit_pairs;
    }
    ctrl->unit_start[p.K] = acc;
    ctrl->total_units = acc;
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&ctrl->tmem_base), kWgTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  const int total_units = ctrl->total_units;

  // unit -> (offset k, first pair, number of pairs); every role evaluates it identically
  auto locate = [&](int u, int& k, int& first, int& count) {
    int lo = 0, hi = p.K;  // largest k with unit_start[k] <= u (empty offsets have no units)
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (ctrl->unit_start[mid] <= u) lo = mid; else hi = mid;
    }
    k = lo;
    const int beg = p.offsets[k], end = p.offsets[k + 1];
    first = beg + (u - ctrl->unit_start[k]) * p.unit_pairs;
    count = min(p.unit_pairs, end - first);
  };

  if (warp < 4) {
    // ===================================== gather producers =====================================
    const uint8_t* xs = reinterpret_cast<const uint8_t*>(p.feats);
    const uint8_t* gs = reinterpret_cast<const uint8_t*>(p.gout);
    const long long in_ld_bytes = p.in_ld * kElem;
    const long long out_ld_bytes = p.out_ld * kElem;
    const int c16 = tid & 7;
    const int seg0 = tid >> 3;  // 0..15: row segment handled per pass (16 segments per pass)
    const long long a_col0 = (long long)(p.in_coff + ys * p.in_y_stride) * kElem + c16 * 16;
    const long long b_col0 =
        (long long)(p.out_coff + ys * p.out_y_stride + zs * p.out_z_stride) * kElem + c16 * 16;
    const int cin_bytes = cin_here * kElem;
    const int cout_bytes = p.cout * kElem;
    constexpr int kPass = kPairs / 16;
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      int k, first, count;
      locate(u, k, first, count);
      int pi_n[kPass], po_n[kPass];
#pragma unroll
      for (int t = 0; t < kPass; ++t) {
        const int pr = seg0 + 16 * t;
        pi_n[t] = pr < count ? __ldg(p.in_maps + first + pr) : -1;
        po_n[t] = pr < count ? __ldg(p.out_maps + first + pr) : -1;
      }
      for (int done = 0; done < count; done += kPairs) {
        int pi[kPass], po[kPass];
#pragma unroll
        for (int t = 0; t < kPass; ++t) { pi[t] = pi_n[t]; po[t] = po_n[t]; }
        if (done + kPairs < count) {  // prefetch the next stage's pair indices
#pragma unroll
          for (int t = 0; t < kPass; ++t) {
            const int pr = done + kPairs + seg0 + 16 * t;
            pi_n[t] = pr < count ? __ldg(p.in_maps + first + pr) : -1;
            po_n[t] = pr < count ? __ldg(p.out_maps + first + pr) : -1;
          }
        }
        mbar_wait(smem_u32(&ctrl->empty[stage]), phase ^ 1u);
        const uint32_t a_smem = smem_base + stage * stage_bytes;
        const uint32_t b_smem = a_smem + kAStage;
#pragma unroll
        for (int t = 0; t < kPass; ++t) {
          const int pr = seg0 + 16 * t;
          const bool valid = pi[t] >= 0;
          const uint8_t* xrow = xs + (long long)(valid ? pi[t] : 0) * in_ld_bytes + a_col0;
          const uint8_t* grow = gs + (long long)(valid ? po[t] : 0) * out_ld_bytes + b_col0;
          const uint32_t sz = valid ? 16u : 0u;
          const uint32_t soff = sw128_offset(pr, c16);
          for (int b = 0; b < n_blk_a; ++b) {
            if (b * 128 + c16 * 16 < cin_bytes)
              cp_async_16(a_smem + b * (kPairs * 128) + soff, xrow + b * 128, sz);
          }
          for (int b = 0; b < n_blk_b; ++b) {
            if (b * 128 + c16 * 16 < cout_bytes)
              cp_async_16(b_smem + b * (kPairs * 128) + soff, grow + b * 128, sz);
          }
        }
        cp_async_mbar_arrive_noinc(smem_u32(&ctrl->full[stage]));
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 4) {
    // ======================================= MMA issuer =========================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(ElemTraits<T>::kFmt, 128, p.cout, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t use = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        int k, first, count;
        locate(u, k, first, count);
        const uint32_t acc = use & 1u;
        mbar_wait(smem_u32(&ctrl->acc_empty[acc]), ((use >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kWgAccStride;
        uint32_t accumulate = 0;
        for (int done = 0; done < count; done += kPairs) {
          mbar_wait(smem_u32(&ctrl->full[stage]), phase);
          tc_fence_after();
          const uint32_t a_smem = smem_base + stage * stage_bytes;
          const uint32_t b_smem = a_smem + kAStage;
#pragma unroll
          for (int j = 0; j < kPairs / kKPerMma; ++j) {
            const uint32_t koff = j * kKPerMma * 128;  // kKPerMma pair rows of 128 bytes
            const uint64_t adesc = make_smem_desc_sw128(a_smem + koff, kPairs * 128, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(b_smem + koff, kPairs * 128, 1024);
            umma_ss<ElemTraits<T>::kTF32>(tmem_d, adesc, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(smem_u32(&ctrl->empty[stage]));
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&ctrl->acc_full[acc]));
        ++use;
      }
    }
  } else {
    // ======================================== epilogue ==========================================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // input-channel row of the slab handled by this thread
    const int gl = (p.gps > 1) ? r / p.cin_g : 0;
    uint32_t use = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      int k, first, count;
      locate(u, k, first, count);
      const uint32_t acc = use & 1u;
      mbar_wait(smem_u32(&ctrl->acc_full[acc]), (use >> 1) & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kWgAccStride;
      float* base = p.dw + (long long)k * p.dw_k_stride + ys * p.dw_y_stride + zs * p.dw_z_stride;
      if (p.gps == 1) {
        float* dst = base + (long long)r * p.dw_ld;
        for (int col = 0; col < p.cout; col += 16) {
          uint32_t v[16];
          tmem_ld_x16(trow + col, v);
          tmem_ld_wait();
          if (r < cin_here) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              red_add_v4(dst + col + i, p.alpha * __uint_as_float(v[i]),
                         p.alpha * __uint_as_float(v[i + 1]), p.alpha * __uint_as_float(v[i + 2]),
                         p.alpha * __uint_as_float(v[i + 3]));
          }
        }
      } else {
        // densified group conv: keep only the diagonal block of each row's group. tcgen05.ld is
        // warp-collective with one address, so walk the groups this warp's 32 rows belong to.
        const int g_lo = (q * 32) / p.cin_g;
        const int g_hi = min(p.gps - 1, (q * 32 + 31) / p.cin_g);
        for (int gg = g_lo; gg <= g_hi; ++gg) {
          float* dst = base + (long long)gg * p.dw_g_stride + (long long)(r - gg * p.cin_g) * p.dw_ld;
          for (int c = 0; c < p.cout_g; c += 8) {
            uint32_t v[8];
            tmem_ld_x8(trow + gg * p.cout_g + c, v);
            tmem_ld_wait();
            if (gl == gg && r < cin_here) {
              red_add_v4(dst + c, p.alpha * __uint_as_float(v[0]), p.alpha * __uint_as_float(v[1]),
                         p.alpha * __uint_as_float(v[2]), p.alpha * __uint_as_float(v[3]));
              red_add_v4(dst + c + 4, p.alpha * __uint_as_float(v[4]),
                         p.alpha * __uint_as_float(v[5]), p.alpha * __uint_as_float(v[6]),
                         p.alpha * __uint_as_float(v[7]));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&ctrl->acc_empty[acc]));
      ++use;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kWgTmemCols);
  }
}

template <typename T>
static int launch_wgrad_t(WgradParams p, int cin_slabs, int cout_slabs, int max_ctas,
                          cudaStream_t stream) {
  constexpr int kElem = (int)sizeof(T);
  constexpr int kBlkElems = 128 / kElem;
  constexpr int kPairs = (kElem == 2) ? 64 : 32;
  const int a_stage = (128 / kBlkElems) * kPairs * 128;
  const int n_blk_b = (p.cout + kBlkElems - 1) / kBlkElems;
  const int stage_bytes = a_stage + ((n_blk_b * kPairs * 128 + 1023) & ~1023);
  if (p.stages <= 0) {
    p.stages = (int)((227 * 1024 - sizeof(WgSmemCtrl) - 1024) / stage_bytes);
    if (p.stages > kWgMaxStages) p.stages = kWgMaxStages;
  }
  if (p.stages < 2) return kErrUnsupportedShape;
  if (p.unit_pairs <= 0) p.unit_pairs = 4096;
  p.unit_pairs = ((p.unit_pairs + kPairs - 1) / kPairs) * kPairs;
  const size_t smem = (size_t)p.stages * stage_bytes + sizeof(WgSmemCtrl) + 1024;
  static int configured_smem = 0;
  if ((int)smem > configured_smem) {
    if (cudaFuncSetAttribute(wgrad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return kErrCuda;
    configured_smem = (int)smem;
  }
  int per_slab = max_ctas / (cin_slabs * cout_slabs);
  if (per_slab < 1) per_slab = 1;
  dim3 grid(per_slab, cin_slabs, cout_slabs);
  wgrad_kernel<T><<<grid, kWgThreads, smem, stream>>>(p);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
}

int launch_wgrad(const WgradParams& p, int dtype, int cin_slabs, int cout_slabs, int max_ctas,
                 cudaStream_t stream) {
  const int es = dtype_size(dtype);
  if (p.K > kWgMaxK || p.K < 1) return kErrUnsupportedShape;
  if (p.cout < 16 || p.cout > 256 || p.cout % 16 != 0) return kErrUnsupportedShape;
  if (p.cin < 1 || p.cin > 128 || (p.cin * es) % 16 != 0) return kErrUnsupportedShape;
  if ((p.in_ld * es) % 16 != 0 || (p.in_coff * es) % 16 != 0) return kErrAlignment;
  if ((p.out_ld * es) % 16 != 0 || (p.out_coff * es) % 16 != 0) return kErrAlignment;
  if (p.dw_ld % 4 != 0 || p.dw_k_stride % 4 != 0 || (reinterpret_cast<uintptr_t>(p.dw) & 15))
    return kErrAlignment;
  if ((reinterpret_cast<uintptr_t>(p.feats) & 15) || (reinterpret_cast<uintptr_t>(p.gout) & 15))
    return kErrAlignment;
  switch (dtype) {
    case kBF16: return launch_wgrad_t<__nv_bfloat16>(p, cin_slabs, cout_slabs, max_ctas, stream);
    case kF16: return launch_wgrad_t<__half>(p, cin_slabs, cout_slabs, max_ctas, stream);
    // fp32 rows as MN-major tf32 operands returned zeros on B200 (bring-up log round 1); the host
    // side splits fp32 into bf16 hi/lo parts instead (detail/unified.py:_wgrad_call).
    default: return kErrUnsupportedDtype;
  }
}

}  // namespace wcn
