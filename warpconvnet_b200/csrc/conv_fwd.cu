// SPDX-License-Identifier: Apache-2.0
// Forward AB_gather_scatter / dgrad ABt_gather_scatter as ONE warp-specialised tcgen05 kernel.
//
// Replaces (semantics, not code): warpconvnet/nn/functional/sparse_conv/detail/explicit.py:22-57
// (Y[out] += X[in] @ W_k over all offsets) and the production kernel family
// warpconvnet/csrc/mask_gemm/include/MaskGemm_forward_*.h (output-stationary, mma.sync).
//
// CTA layout (288 threads, 1 CTA / SM, persistent over tiles):
//   warps 0-3  gather producers: 8 lanes fetch one 128-byte row segment with cp.async (16 B each,
//              zero-fill for missing neighbours) straight into the 128B-swizzled K-major A tile
//   warp  4    lane 0 issues tcgen05.mma (M=128, N=bn, K=32 B per instruction), accumulators in
//              TMEM, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 5-8  epilogue: tcgen05.ld -> (+bias, ReLU) -> bf16/fp16/fp32 -> 128-bit global stores
// The per-offset weight slice is a pre-swizzled image in global memory and is pulled by the TMA
// unit with one cp.async.bulk per pipeline stage.
#include "common.cuh"
#include "conv_gemm.cuh"

namespace wcn {

constexpr int kTileM = 128;
constexpr int kProducerThreads = 128;
constexpr int kMmaWarp = 4;
constexpr int kGemmThreads = 288;
constexpr int kAStageBytes = kTileM * 128;  // 16 KB
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;             // TMEM column offset of the second accumulator
constexpr int kMaxStages = 8;

struct GemmSmemCtrl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

template <typename T>
__global__ void __launch_bounds__(kGemmThreads, 1)
gather_gemm_kernel(const __grid_constant__ GatherGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int slab = blockIdx.y;

  const int b_stage_bytes = p.bn * 128;
  const int stage_bytes = kAStageBytes + ((b_stage_bytes + 1023) & ~1023);
  const int stages = p.stages;
  GemmSmemCtrl* ctrl = reinterpret_cast<GemmSmemCtrl*>(smem_gen + (size_t)stages * stage_bytes);

  constexpr int kElem = (int)sizeof(T);
  constexpr int kChunkElems = 128 / kElem;  // channels per 128-byte row segment
  const int n_chunks = (p.cin + kChunkElems - 1) / kChunkElems;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&ctrl->full[s]), kProducerThreads + 1);
      mbar_init(smem_u32(&ctrl->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&ctrl->acc_full[a]), 1);
      mbar_init(smem_u32(&ctrl->acc_empty[a]), 128);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(smem_u32(&ctrl->tmem_base), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  if (warp < 4) {
    // ===================================== gather producers =====================================
    const uint8_t* feats = reinterpret_cast<const uint8_t*>(p.feats);
    const long long in_ld_bytes = p.in_ld * kElem;
    const uint8_t* wimg = reinterpret_cast<const uint8_t*>(p.wimg) +
                          (size_t)slab * p.K * n_chunks * b_stage_bytes;
    const int c16 = lane & 7;
    const int sub = lane >> 3;  // which of the 4 rows an instruction covers
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nk = p.tile_nk[tile];
      const uint16_t* ks = p.tile_ks + (size_t)tile * p.k_stride;
      const int pos = tile * kTileM + warp * 32 + lane;
      int idx_next = (nk > 0) ? __ldg(p.nbr + (size_t)ks[0] * p.m_pad + pos) : -1;
      for (int ki = 0; ki < nk; ++ki) {
        const int k = ks[ki];
        const int idx_own = idx_next;
        if (ki + 1 < nk) idx_next = __ldg(p.nbr + (size_t)ks[ki + 1] * p.m_pad + pos);
        const int wk = p.kflip ? (p.K - 1 - k) : k;
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(smem_u32(&ctrl->empty[stage]), phase ^ 1u);
          const uint32_t a_smem = smem_base + stage * stage_bytes;
          const uint32_t full_bar = smem_u32(&ctrl->full[stage]);
          if (tid == 0) {
            mbar_arrive_expect_tx(full_bar, (uint32_t)b_stage_bytes);
            bulk_copy_g2s(a_smem + kAStageBytes,
                          wimg + ((size_t)wk * n_chunks + c) * b_stage_bytes,
                          (uint32_t)b_stage_bytes, full_bar);
          }
          const int chunk_bytes = min(128, (p.cin - c * kChunkElems) * kElem);
          const long long col_bytes =
              (long long)(p.in_coff + slab * p.in_slab_stride + c * kChunkElems) * kElem + c16 * 16;
          const bool lane_active = c16 * 16 < chunk_bytes;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rloc = 4 * j + sub;
            const int src_idx = __shfl_sync(0xffffffffu, idx_own, rloc);  // warp-uniform call
            const uint32_t row = warp * 32 + rloc;
            const uint8_t* src =
                feats + (long long)(src_idx >= 0 ? src_idx : 0) * in_ld_bytes + col_bytes;
            if (lane_active)
              cp_async_16(a_smem + sw128_offset(row, c16), src, src_idx >= 0 ? 16u : 0u);
          }
          cp_async_mbar_arrive_noinc(full_bar);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ======================================= MMA issuer =========================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(ElemTraits<T>::kFmt, kTileM, p.bn, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t use = 0;  // number of accumulator uses so far
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int nk = p.tile_nk[tile];
        if (nk == 0) continue;
        const uint32_t acc = use & 1u;
        mbar_wait(smem_u32(&ctrl->acc_empty[acc]), ((use >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        uint32_t accumulate = 0;
        for (int ki = 0; ki < nk; ++ki) {
          for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(smem_u32(&ctrl->full[stage]), phase);
            tc_fence_after();
            const uint32_t a_smem = smem_base + stage * stage_bytes;
            const uint32_t b_smem = a_smem + kAStageBytes;
            const int chunk_bytes = min(128, (p.cin - c * kChunkElems) * kElem);
            const int n_mma = chunk_bytes >> 5;  // 32 bytes of K per instruction
            for (int j = 0; j < n_mma; ++j) {
              const uint64_t adesc = make_smem_desc_sw128(a_smem + j * 32, 16, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(b_smem + j * 32, 16, 1024);
              umma_ss<ElemTraits<T>::kTF32>(tmem_d, adesc, bdesc, idesc, accumulate);
              accumulate = 1;
            }
            umma_commit(smem_u32(&ctrl->empty[stage]));
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
        umma_commit(smem_u32(&ctrl->acc_full[acc]));
        ++use;
      }
    }
  } else {
    // ======================================== epilogue ==========================================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    uint8_t* out = reinterpret_cast<uint8_t*>(p.out);
    const long long out_ld_bytes = p.out_ld * kElem;
    const int col_base = p.out_coff + slab * p.bn;
    uint32_t use = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nk = p.tile_nk[tile];
      const int out_row = __ldg(p.rows + tile * kTileM + r);
      uint32_t acc = 0;
      if (nk > 0) {
        acc = use & 1u;
        mbar_wait(smem_u32(&ctrl->acc_full[acc]), (use >> 1) & 1u);
        tc_fence_after();
      }
      uint8_t* out_ptr = out + (long long)(out_row >= 0 ? out_row : 0) * out_ld_bytes +
                         (long long)col_base * kElem;
      for (int col = 0; col < p.bn; col += 16) {
        uint32_t v[16];
        if (nk > 0) {
          tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride + col, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            v[i] = __float_as_uint(__uint_as_float(v[i]) + __ldg(p.bias + col_base + col + i));
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(fmaxf(__uint_as_float(v[i]), 0.f));
        }
        if (out_row >= 0) {
          if constexpr (sizeof(T) == 2) {
            uint4* dst = reinterpret_cast<uint4*>(out_ptr + col * 2);
            dst[0] = pack8<T>(v);
            dst[1] = pack8<T>(v + 8);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(out_ptr + col * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
      }
      if (nk > 0) {
        tc_fence_before();
        mbar_arrive(smem_u32(&ctrl->acc_empty[acc]));
        ++use;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

static size_t gemm_smem_bytes(int bn, int stages) {
  const int stage_bytes = kAStageBytes + ((bn * 128 + 1023) & ~1023);
  return (size_t)stages * stage_bytes + sizeof(GemmSmemCtrl) + 1024;
}

int pick_gemm_stages(int bn) {
  const int stage_bytes = kAStageBytes + ((bn * 128 + 1023) & ~1023);
  int s = (int)((227 * 1024 - sizeof(GemmSmemCtrl) - 1024) / stage_bytes);
  if (s > kMaxStages) s = kMaxStages;
  return s;
}

template <typename T>
static int launch_gather_gemm_t(GatherGemmParams p, int n_slabs, int max_ctas, cudaStream_t stream) {
  if (p.stages <= 0) p.stages = pick_gemm_stages(p.bn);
  if (p.stages < 2) return kErrUnsupportedShape;
  const size_t smem = gemm_smem_bytes(p.bn, p.stages);
  static int configured_smem = 0;  // per instantiation
  if ((int)smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(gather_gemm_kernel<T>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return kErrCuda;
    configured_smem = (int)smem;
  }
  int ctas = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  if (ctas < 1) return kOk;
  dim3 grid(ctas, n_slabs, 1);
  gather_gemm_kernel<T><<<grid, kGemmThreads, smem, stream>>>(p);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
}

int launch_gather_gemm(const GatherGemmParams& p, int dtype, int n_slabs, int max_ctas,
                       cudaStream_t stream) {
  const int es = dtype_size(dtype);
  if (p.bn < 16 || p.bn > 256 || (p.bn % 16) != 0) return kErrUnsupportedShape;
  if (p.cin <= 0 || (p.cin * es) % 32 != 0) return kErrUnsupportedShape;
  if ((p.in_ld * es) % 16 != 0 || (p.in_coff * es) % 16 != 0 || (p.in_slab_stride * es) % 16 != 0)
    return kErrAlignment;
  if ((p.out_ld * es) % 16 != 0 || (p.out_coff * es) % 16 != 0) return kErrAlignment;
  if ((reinterpret_cast<uintptr_t>(p.feats) & 15) || (reinterpret_cast<uintptr_t>(p.out) & 15) ||
      (reinterpret_cast<uintptr_t>(p.wimg) & 15))
    return kErrAlignment;
  if (p.m_pad % kTileM != 0 || p.num_tiles * kTileM > p.m_pad) return kErrInvalidArg;
  switch (dtype) {
    case kBF16: return launch_gather_gemm_t<__nv_bfloat16>(p, n_slabs, max_ctas, stream);
    case kF16: return launch_gather_gemm_t<__half>(p, n_slabs, max_ctas, stream);
    case kF32: return launch_gather_gemm_t<float>(p, n_slabs, max_ctas, stream);
    default: return kErrUnsupportedDtype;
  }
}

}  // namespace wcn
