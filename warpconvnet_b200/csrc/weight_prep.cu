// SPDX-License-Identifier: Apache-2.0
// Weight image builder: turns W[K,(G,)Cin/G,Cout/G] into the exact shared-memory byte image the
// gather-GEMM kernel consumes (K-major "B" tiles, 128-byte rows, 128B XOR swizzle baked in), so
// the kernel can pull one [bn x 128 B] slice per pipeline stage with a single cp.async.bulk.
//
// Replaces: the per-call `weight.transpose(1,2).contiguous()` of the reference's dgrad
// (warpconvnet/nn/functional/sparse_conv/detail/unified.py:654-671) and the in-kernel B-tile
// cp.async of MaskGemm_forward_*.h.
//
// image layout: [n_slabs][K][n_chunks][bn][128 B]
//   row n of slab s  <-> output channel (fwd) / input channel (dgrad)  s*bn + n
//   chunk c, byte b  <-> contraction channel c*(128/es) + b/es (zero beyond the real length)
// Group convolutions are densified block-diagonally inside a slab (gps groups per slab).
#include "common.cuh"
#include "conv_gemm.cuh"

namespace wcn {

// blockIdx.y selects the parameter block: one launch builds the forward image and the transposed
// (dgrad) image of the same weights, optionally converting fp32 master weights to the 16-bit
// compute type on the way (replaces weight.to(bf16) + two image launches per layer and step).
__global__ void weight_image_kernel(const WeightPrepParams pa, const WeightPrepParams pb) {
  pdl_begin();
  const WeightPrepParams& p = blockIdx.y == 0 ? pa : pb;
  const int bn = p.gps * p.rg;
  const int cdim = p.gps * p.cg;
  const long long total = (long long)p.n_slabs * p.K * p.n_chunks * bn * 8;
  const int epu = 16 / p.es;  // elements per 16-byte unit
  const int ce = 128 / p.es;  // elements per chunk
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c16 = (int)(t & 7);
    long long u = t >> 3;
    const int n = (int)(u % bn); u /= bn;
    const int c = (int)(u % p.n_chunks); u /= p.n_chunks;
    const int k = (int)(u % p.K);
    const int s = (int)(u / p.K);
    const int gl_row = n / p.rg;
    const int r = n - gl_row * p.rg;
    uint4 val = make_uint4(0, 0, 0, 0);
    uint8_t* vb = reinterpret_cast<uint8_t*>(&val);
    for (int e = 0; e < epu; ++e) {
      const int j = c * ce + c16 * epu + e;  // contraction index inside the slab
      if (j < cdim && j / p.cg == gl_row) {
        const int cc = j - gl_row * p.cg;
        const long long src = (long long)k * p.w_k_stride +
                              (long long)(s * p.gps + gl_row) * p.w_g_stride +
                              (long long)r * p.w_r_stride + (long long)cc * p.w_c_stride;
        if (p.cvt == 1) {
          const __nv_bfloat16 h = __float2bfloat16_rn(reinterpret_cast<const float*>(p.w)[src]);
          reinterpret_cast<uint16_t*>(vb)[e] = *reinterpret_cast<const uint16_t*>(&h);
        } else if (p.cvt == 2) {
          const __half h = __float2half_rn(reinterpret_cast<const float*>(p.w)[src]);
          reinterpret_cast<uint16_t*>(vb)[e] = *reinterpret_cast<const uint16_t*>(&h);
        } else if (p.es == 2) {
          reinterpret_cast<uint16_t*>(vb)[e] = reinterpret_cast<const uint16_t*>(p.w)[src];
        } else {
          reinterpret_cast<uint32_t*>(vb)[e] = reinterpret_cast<const uint32_t*>(p.w)[src];
        }
      }
    }
    const long long tile = (((long long)s * p.K + k) * p.n_chunks + c) * bn;
    uint8_t* dst = reinterpret_cast<uint8_t*>(p.img) + (tile + n) * 128 + ((c16 ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = val;
  }
}

static long long image_units(const WeightPrepParams& p) {
  return (long long)p.n_slabs * p.K * p.n_chunks * (p.gps * p.rg) * 8;
}

// second == nullptr: one image
int launch_weight_image(const WeightPrepParams& p, const WeightPrepParams* second,
                        cudaStream_t stream) {
  for (const WeightPrepParams* q : {&p, second}) {
    if (q == nullptr) continue;
    const int bn = q->gps * q->rg;
    if (bn % 16 != 0 || bn > 256 || bn < 16) return kErrUnsupportedShape;
  }
  long long total = image_units(p);
  if (second != nullptr && image_units(*second) > total) total = image_units(*second);
  if (total == 0) return kOk;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  wcn_launch(weight_image_kernel, dim3(dim3(blocks, second ? 2 : 1)), dim3(256), 0, stream, p, second ? *second : p);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
}

}  // namespace wcn
