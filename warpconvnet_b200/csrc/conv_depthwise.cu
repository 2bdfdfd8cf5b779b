// SPDX-License-Identifier: Apache-2.0
// Depthwise sparse convolution (weight [K, C]): y[out, c] = sum_k x[nbr_k(out), c] * w[k, c].
// No contraction over channels, so this is an HBM/L2-bound gather-FMA, not tensor-core work
// (SURVEY.md §8 f3). Output-stationary over the [K, M] neighbour table the kernel map already
// holds: one thread owns a 16-byte channel vector of one output row, walks the K offsets with the
// neighbour indices loaded in independent batches, accumulates in fp32 registers and writes the
// row once (no atomics, no zero-fill). dgrad is the same kernel on the reverse table (for
// submanifold maps: the forward table with the offset index flipped). wgrad reduces
// x[nbr_k(out), c] * dy[out, c] per (k, c) in shared memory and leaves with one fp32 atomic per
// (k, c) and block.
//
// Replaces (semantics, not code): warpconvnet/nn/functional/sparse_conv_depth.py:227-420
// (explicit: one index_select + multiply + index_add per offset; implicit: csrc/implicit_fma_kernel.cu,
// csrc/implicit_reduction.cu).
#include "common.cuh"

namespace wcn {

struct DepthwiseParams {
  const void* x;        // gathered operand [n_src, ld_x]
  const void* dy;       // wgrad: the row-aligned operand [M, ld_dy]
  void* y;              // fwd / dgrad output [M, ld_y]
  const float* w;       // [K, C] fp32
  float* dw;            // wgrad target [K, C] fp32, accumulated into
  const float* bias;    // optional [C]
  const int* table;     // [K, M] neighbour rows (-1 = none)
  // mask-sorted tile plan (wcn_build_tiles), forward / dgrad only: compact step lists
  const int* step_nbr;  // [num_tiles][K][tile_rows]
  const int* step_k;    // [num_tiles][K]
  const int* rows;      // [num_tiles * tile_rows] output row of each sorted position, -1 = padding
  const int* tile_nk;   // [num_tiles]
  int num_tiles, tile_rows;
  long long ld_x, ld_dy, ld_y;
  int M, K, C;
  int kflip;            // weight row K-1-k for table row k
  int relu;
};

constexpr int kDwThreads = 256;
constexpr int kDwBatch = 9;  // neighbour indices in flight per thread

template <typename T, int V>
__device__ __forceinline__ void dw_load(const T* p, float (&f)[V]) {
  if constexpr (V == 1) {
    f[0] = (float)p[0];
  } else {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
    if constexpr (sizeof(T) == 4) {
      f[0] = __uint_as_float(raw.x); f[1] = __uint_as_float(raw.y);
      f[2] = __uint_as_float(raw.z); f[3] = __uint_as_float(raw.w);
    } else {
      const T* h = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] = (float)h[i];
    }
  }
}

template <typename T, int V>
__device__ __forceinline__ void dw_store(T* p, const float (&f)[V]) {
  if constexpr (V == 1) {
    p[0] = (T)f[0];
  } else if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<uint4*>(p) = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]),
                                              __float_as_uint(f[2]), __float_as_uint(f[3]));
  } else {
    uint4 raw;
    T* h = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int i = 0; i < V; ++i) h[i] = (T)f[i];
    *reinterpret_cast<uint4*>(p) = raw;
  }
}

// Raw gathered vector: kept packed (4 registers per 16-byte load) until it is consumed, so a
// thread can keep kDwBatch loads in flight; `on == false` zeroes it before the unpack (NaN-safe:
// a missing neighbour reads row 0).
template <typename T, int V>
struct DwRaw {
  uint4 q;
  float s;
};
template <typename T, int V>
__device__ __forceinline__ DwRaw<T, V> dw_load_raw(const T* p) {
  DwRaw<T, V> r;
  if constexpr (V == 1) {
    r.s = (float)p[0];
    r.q = make_uint4(0, 0, 0, 0);
  } else {
    r.q = __ldg(reinterpret_cast<const uint4*>(p));
    r.s = 0.f;
  }
  return r;
}
template <typename T, int V>
__device__ __forceinline__ void dw_unpack(DwRaw<T, V> r, bool on, float (&f)[V]) {
  if constexpr (V == 1) {
    f[0] = on ? r.s : 0.f;
  } else {
    const uint32_t w[4] = {on ? r.q.x : 0u, on ? r.q.y : 0u, on ? r.q.z : 0u, on ? r.q.w : 0u};
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) f[i] = __uint_as_float(w[i]);
    } else if constexpr (ElemTraits<T>::kFmt == 1) {  // bf16: the high half of an fp32
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        f[2 * i] = v.x;
        f[2 * i + 1] = v.y;
      }
    }
  }
}

// forward / dgrad
template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads) depthwise_fwd_kernel(const DepthwiseParams p) {
  extern __shared__ float s_w[];  // [K][C]
  for (int i = threadIdx.x; i < p.K * p.C; i += kDwThreads) s_w[i] = __ldg(p.w + i);
  __syncthreads();
  const int vecs = (p.C + V - 1) / V;
  const int rpb = kDwThreads / vecs;
  const int vc = threadIdx.x % vecs, rl = threadIdx.x / vecs;
  if (rl >= rpb) return;
  const T* x = reinterpret_cast<const T*>(p.x) + vc * V;
  T* y = reinterpret_cast<T*>(p.y) + vc * V;
  float bias[V];
#pragma unroll
  for (int i = 0; i < V; ++i)
    bias[i] = (p.bias != nullptr && vc * V + i < p.C) ? __ldg(p.bias + vc * V + i) : 0.f;
  for (long long r = (long long)blockIdx.x * rpb + rl; r < p.M; r += (long long)gridDim.x * rpb) {
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = bias[i];
    for (int k0 = 0; k0 < p.K; k0 += kDwBatch) {
      int idx[kDwBatch];
#pragma unroll
      for (int u = 0; u < kDwBatch; ++u)
        idx[u] = (k0 + u < p.K) ? __ldg(p.table + (size_t)(k0 + u) * p.M + r) : -1;
      // unconditional gathers (a missing neighbour reads row 0, which stays in L1, and its value
      // is replaced by 0): all kDwBatch loads of a thread are in flight together; the branchy form
      // serialised one L2 round trip per valid neighbour (302 us on C3-S at 128 channels)
      DwRaw<T, V> raw[kDwBatch];
#pragma unroll
      for (int u = 0; u < kDwBatch; ++u)
        raw[u] = dw_load_raw<T, V>(x + (long long)max(idx[u], 0) * p.ld_x);
#pragma unroll
      for (int u = 0; u < kDwBatch; ++u) {
        // the loads above are unconditional (all in flight together); the arithmetic of a
        // missing neighbour (2 of 3 on surface data) is skipped
        if (idx[u] < 0) continue;
        const int k = k0 + u;
        const float* wk = s_w + (p.kflip ? p.K - 1 - k : k) * p.C + vc * V;
        float f[V];
        dw_unpack<T, V>(raw[u], true, f);
        if constexpr (V >= 4) {
#pragma unroll
          for (int i = 0; i < V; i += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wk + i);
            acc[i] = fmaf(f[i], w4.x, acc[i]);
            acc[i + 1] = fmaf(f[i + 1], w4.y, acc[i + 1]);
            acc[i + 2] = fmaf(f[i + 2], w4.z, acc[i + 2]);
            acc[i + 3] = fmaf(f[i + 3], w4.w, acc[i + 3]);
          }
        } else {
          acc[0] = fmaf(f[0], wk[0], acc[0]);
        }
      }
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = fmaxf(acc[i], 0.f);
    }
    dw_store<T, V>(y + r * p.ld_y, acc);
  }
}

// forward / dgrad on the mask-sorted tile plan the tensor-core kernels use: a tile's rows share one
// compact list of active offsets (9.2 of 27 on surface data), so the control flow is uniform, no
// index is loaded for an inactive offset, the step's neighbour indices are contiguous (coalesced)
// and its weight row is a shared-memory broadcast. One thread = one channel vector of one tile
// row; work unit = (tile, pass of 256 / vecs rows), grid-stride.
constexpr int kDwPlanBatch = 5;
template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads) depthwise_plan_kernel(const DepthwiseParams p) {
  extern __shared__ float s_w[];  // [K][C]
  for (int i = threadIdx.x; i < p.K * p.C; i += kDwThreads) s_w[i] = __ldg(p.w + i);
  __syncthreads();
  const int vecs = (p.C + V - 1) / V;
  const int rpb = kDwThreads / vecs;
  const int vc = threadIdx.x % vecs, rl = threadIdx.x / vecs;
  const int passes = (p.tile_rows + rpb - 1) / rpb;
  const T* x = reinterpret_cast<const T*>(p.x) + vc * V;
  T* y = reinterpret_cast<T*>(p.y) + vc * V;
  float bias[V];
#pragma unroll
  for (int i = 0; i < V; ++i)
    bias[i] = (p.bias != nullptr && vc * V + i < p.C) ? __ldg(p.bias + vc * V + i) : 0.f;
  const long long units = (long long)p.num_tiles * passes;
  for (long long q = blockIdx.x; q < units; q += gridDim.x) {
    const int tile = (int)(q / passes);
    const int row_in_tile = (int)(q - (long long)tile * passes) * rpb + rl;
    if (rl >= rpb || row_in_tile >= p.tile_rows) continue;
    const int out_row = __ldg(p.rows + (size_t)tile * p.tile_rows + row_in_tile);
    if (out_row < 0) continue;
    const int nk = __ldg(p.tile_nk + tile);
    const int* nbr = p.step_nbr + (size_t)tile * p.K * p.tile_rows + row_in_tile;
    const int* sk = p.step_k + (size_t)tile * p.K;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = bias[i];
    for (int i0 = 0; i0 < nk; i0 += kDwPlanBatch) {
      int idx[kDwPlanBatch], kk[kDwPlanBatch];
#pragma unroll
      for (int u = 0; u < kDwPlanBatch; ++u) {
        const bool live = i0 + u < nk;  // block-uniform
        idx[u] = live ? __ldg(nbr + (size_t)(i0 + u) * p.tile_rows) : -1;
        kk[u] = live ? __ldg(sk + i0 + u) : 0;
      }
      DwRaw<T, V> raw[kDwPlanBatch];
#pragma unroll
      for (int u = 0; u < kDwPlanBatch; ++u)
        raw[u] = dw_load_raw<T, V>(x + (long long)max(idx[u], 0) * p.ld_x);
#pragma unroll
      for (int u = 0; u < kDwPlanBatch; ++u) {
        if (i0 + u >= nk) break;  // block-uniform
        const float* wk = s_w + (p.kflip ? p.K - 1 - kk[u] : kk[u]) * p.C + vc * V;
        float f[V];
        dw_unpack<T, V>(raw[u], idx[u] >= 0, f);
        if constexpr (V >= 4) {
#pragma unroll
          for (int i = 0; i < V; i += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wk + i);
            acc[i] = fmaf(f[i], w4.x, acc[i]);
            acc[i + 1] = fmaf(f[i + 1], w4.y, acc[i + 1]);
            acc[i + 2] = fmaf(f[i + 2], w4.z, acc[i + 2]);
            acc[i + 3] = fmaf(f[i + 3], w4.w, acc[i + 3]);
          }
        } else {
          acc[0] = fmaf(f[0], wk[0], acc[0]);
        }
      }
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = fmaxf(acc[i], 0.f);
    }
    dw_store<T, V>(y + (long long)out_row * p.ld_y, acc);
  }
}

// wgrad: dw[k][c] += sum_r x[table[k][r]][c] * dy[r][c]. blockIdx.y selects a chunk of kDwBatch
// offsets; a thread owns one channel vector, walks rows with a grid stride and keeps the chunk's
// kDwBatch x V partial sums in registers (no atomics inside the row loop). Partials are merged per
// block in shared memory and leave as one fp32 atomic per (k, c) and block.
template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads) depthwise_wgrad_kernel(const DepthwiseParams p) {
  extern __shared__ float s_dw[];  // [kDwBatch][C]
  for (int i = threadIdx.x; i < kDwBatch * p.C; i += kDwThreads) s_dw[i] = 0.f;
  __syncthreads();
  const int vecs = (p.C + V - 1) / V;
  const int rpb = kDwThreads / vecs;
  const int vc = threadIdx.x % vecs, rl = threadIdx.x / vecs;
  const int k0 = blockIdx.y * kDwBatch;
  if (rl < rpb) {
    const T* x = reinterpret_cast<const T*>(p.x) + vc * V;
    const T* dy = reinterpret_cast<const T*>(p.dy) + vc * V;
    float acc[kDwBatch][V];
#pragma unroll
    for (int u = 0; u < kDwBatch; ++u)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[u][i] = 0.f;
    for (long long r = (long long)blockIdx.x * rpb + rl; r < p.M;
         r += (long long)gridDim.x * rpb) {
      int idx[kDwBatch];
#pragma unroll
      for (int u = 0; u < kDwBatch; ++u)
        idx[u] = (k0 + u < p.K) ? __ldg(p.table + (size_t)(k0 + u) * p.M + r) : -1;
      float g[V];
      dw_load<T, V>(dy + r * p.ld_dy, g);
      DwRaw<T, V> raw[kDwBatch];
#pragma unroll
      for (int u = 0; u < kDwBatch; ++u)
        raw[u] = dw_load_raw<T, V>(x + (long long)max(idx[u], 0) * p.ld_x);
#pragma unroll
      for (int u = 0; u < kDwBatch; ++u) {
        // branch-free here: skipping the arithmetic of missing neighbours measured slower in this
        // kernel (662 vs 374 us at 128 channels on C3-S), unlike the forward kernel
        float f[V];
        dw_unpack<T, V>(raw[u], idx[u] >= 0, f);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[u][i] = fmaf(f[i], g[i], acc[u][i]);
      }
    }
#pragma unroll
    for (int u = 0; u < kDwBatch; ++u)
#pragma unroll
      for (int i = 0; i < V; ++i)
        if (vc * V + i < p.C && acc[u][i] != 0.f) atomicAdd(s_dw + u * p.C + vc * V + i, acc[u][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kDwBatch * p.C; i += kDwThreads) {
    const int u = i / p.C, ch = i - u * p.C;
    const float v = s_dw[i];
    if (k0 + u < p.K && v != 0.f) atomicAdd(p.dw + (size_t)(k0 + u) * p.C + ch, v);
  }
}

// wgrad on the tile plan: dw[k][c] += sum over the tile rows r of x[nbr_i(r)][c] * dy[row(r)][c] for
// every step i (offset k = step_k[i]) of every tile. A block walks tiles with a grid stride; a
// thread owns one channel vector and the tile rows rl, rl + rpb, ...; kDwPlanBatch steps at a time
// are accumulated in registers over those rows, merged into the block's [K][C] shared-memory
// accumulator with one atomic per (step, channel, thread) and flushed to dw once per block.
template <typename T, int V>
__global__ void __launch_bounds__(kDwThreads) depthwise_plan_wgrad_kernel(const DepthwiseParams p) {
  extern __shared__ float s_dw[];  // [K][C]
  for (int i = threadIdx.x; i < p.K * p.C; i += kDwThreads) s_dw[i] = 0.f;
  __syncthreads();
  const int vecs = (p.C + V - 1) / V;
  const int rpb = kDwThreads / vecs;
  const int vc = threadIdx.x % vecs, rl = threadIdx.x / vecs;
  if (rl < rpb) {
    const T* x = reinterpret_cast<const T*>(p.x) + vc * V;
    const T* dy = reinterpret_cast<const T*>(p.dy) + vc * V;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nk = __ldg(p.tile_nk + tile);
      const int* sk = p.step_k + (size_t)tile * p.K;
      for (int i0 = 0; i0 < nk; i0 += kDwPlanBatch) {
        float acc[kDwPlanBatch][V];
#pragma unroll
        for (int u = 0; u < kDwPlanBatch; ++u)
#pragma unroll
          for (int i = 0; i < V; ++i) acc[u][i] = 0.f;
        for (int row = rl; row < p.tile_rows; row += rpb) {
          const int out_row = __ldg(p.rows + (size_t)tile * p.tile_rows + row);
          if (out_row < 0) continue;
          const int* nbr = p.step_nbr + ((size_t)tile * p.K + i0) * p.tile_rows + row;
          int idx[kDwPlanBatch];
#pragma unroll
          for (int u = 0; u < kDwPlanBatch; ++u)
            idx[u] = (i0 + u < nk) ? __ldg(nbr + (size_t)u * p.tile_rows) : -1;
          float g[V];
          dw_load<T, V>(dy + (long long)out_row * p.ld_dy, g);
          DwRaw<T, V> raw[kDwPlanBatch];
#pragma unroll
          for (int u = 0; u < kDwPlanBatch; ++u)
            raw[u] = dw_load_raw<T, V>(x + (long long)max(idx[u], 0) * p.ld_x);
#pragma unroll
          for (int u = 0; u < kDwPlanBatch; ++u) {
            float f[V];
            dw_unpack<T, V>(raw[u], idx[u] >= 0, f);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[u][i] = fmaf(f[i], g[i], acc[u][i]);
          }
        }
#pragma unroll
        for (int u = 0; u < kDwPlanBatch; ++u) {
          if (i0 + u >= nk) break;  // block-uniform
          float* dst = s_dw + __ldg(sk + i0 + u) * p.C + vc * V;
#pragma unroll
          for (int i = 0; i < V; ++i)
            if (vc * V + i < p.C) atomicAdd(dst + i, acc[u][i]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.K * p.C; i += kDwThreads) {
    const float v = s_dw[i];
    if (v != 0.f) atomicAdd(p.dw + i, v);
  }
}

template <typename T, int V>
static int dw_launch_tv(int mode, const DepthwiseParams& p, cudaStream_t s) {
  // mode 0: forward / dgrad on the dense table, 1: wgrad on the table, 2: forward / dgrad on the
  // tile plan, 3: wgrad on the tile plan
  const bool wgrad = mode == 1;
  const size_t sh = (size_t)(wgrad ? kDwBatch : p.K) * p.C * sizeof(float);
  const int vecs = (p.C + V - 1) / V;
  const int rpb = kDwThreads / vecs;
  static int occ[4] = {0, 0, 0, 0};
  static size_t configured_dev[kMaxDevices][4] = {};  // per instantiation and device
  size_t* configured = configured_dev[current_device_slot()];
  auto set_attr = [&](int bytes) {
    switch (mode) {
      case 0: return cudaFuncSetAttribute(depthwise_fwd_kernel<T, V>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      case 1: return cudaFuncSetAttribute(depthwise_wgrad_kernel<T, V>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      case 2: return cudaFuncSetAttribute(depthwise_plan_kernel<T, V>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      default: return cudaFuncSetAttribute(depthwise_plan_wgrad_kernel<T, V>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    }
  };
  auto get_occ = [&](int* o) {
    switch (mode) {
      case 0: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(o, depthwise_fwd_kernel<T, V>,
                                                                   kDwThreads, sh);
      case 1: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(o, depthwise_wgrad_kernel<T, V>,
                                                                   kDwThreads, sh);
      case 2: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(o, depthwise_plan_kernel<T, V>,
                                                                   kDwThreads, sh);
      default: return cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          o, depthwise_plan_wgrad_kernel<T, V>, kDwThreads, sh);
    }
  };
  if (sh > 48 * 1024 && sh > configured[mode]) {
    if (set_attr((int)sh) != cudaSuccess) return kErrCuda;
    configured[mode] = sh;
    occ[mode] = 0;
  }
  if (occ[mode] == 0) {
    int o = 0;
    occ[mode] = (get_occ(&o) == cudaSuccess && o > 0) ? o : 1;
  }
  const long long cap = (long long)kNumSMsB200 * occ[mode];
  if (mode == 3) {
    long long blocks = p.num_tiles < cap ? p.num_tiles : cap;
    if (blocks < 1) blocks = 1;
    depthwise_plan_wgrad_kernel<T, V><<<(int)blocks, kDwThreads, sh, s>>>(p);
  } else if (mode == 2) {
    const long long units = (long long)p.num_tiles * ((p.tile_rows + rpb - 1) / rpb);
    long long blocks = units < cap ? units : cap;
    if (blocks < 1) blocks = 1;
    depthwise_plan_kernel<T, V><<<(int)blocks, kDwThreads, sh, s>>>(p);
  } else {
    long long blocks = ((long long)p.M + rpb - 1) / rpb;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (wgrad) {
      // y = offset chunk; the resident wave is shared between the chunks
      const int chunks = (p.K + kDwBatch - 1) / kDwBatch;
      long long bx = (cap + chunks - 1) / chunks;
      if (bx > blocks) bx = blocks;
      if (bx < 1) bx = 1;
      depthwise_wgrad_kernel<T, V><<<dim3((unsigned)bx, (unsigned)chunks), kDwThreads, sh, s>>>(p);
    } else {
      depthwise_fwd_kernel<T, V><<<(int)blocks, kDwThreads, sh, s>>>(p);
    }
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
}

template <typename T>
static int dw_launch_t(int mode, const DepthwiseParams& p, bool vec, cudaStream_t s) {
  constexpr int V = 16 / (int)sizeof(T);
  return vec ? dw_launch_tv<T, V>(mode, p, s) : dw_launch_tv<T, 1>(mode, p, s);
}

static bool dw_aligned(const void* ptr, long long ld, int es) {
  return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * es) % 16 == 0);
}

static int dw_dispatch(int mode, DepthwiseParams& p, int dtype, cudaStream_t s) {
  if ((size_t)p.K * p.C * sizeof(float) > 200 * 1024) return kErrUnsupportedShape;
  const int es = dtype_size(dtype);
  const int v = 16 / es;
  const bool vec = (p.C % v == 0) && p.C / v <= kDwThreads && dw_aligned(p.x, p.ld_x, es) &&
                   dw_aligned(p.dy, p.ld_dy, es) && dw_aligned(p.y, p.ld_y, es);
  if (!vec && p.C > kDwThreads) return kErrUnsupportedShape;
  switch (dtype) {
    case kBF16: return dw_launch_t<__nv_bfloat16>(mode, p, vec, s);
    case kF16: return dw_launch_t<__half>(mode, p, vec, s);
    case kF32: return dw_launch_t<float>(mode, p, vec, s);
    default: return kErrUnsupportedDtype;
  }
}

int depthwise_launch(bool wgrad, const void* x, long long ld_x, const void* dy, long long ld_dy,
                     void* y, long long ld_y, const float* w, float* dw, const float* bias,
                     const int* table, int M, int K, int C, int kflip, int relu, int dtype,
                     cudaStream_t s) {
  if (M < 0 || K < 1 || C < 1) return kErrInvalidArg;
  if (M == 0) return kOk;
  DepthwiseParams p{};
  p.x = x; p.dy = dy; p.y = y; p.w = w; p.dw = dw; p.bias = bias; p.table = table;
  p.ld_x = ld_x; p.ld_dy = ld_dy; p.ld_y = ld_y;
  p.M = M; p.K = K; p.C = C; p.kflip = kflip; p.relu = relu;
  return dw_dispatch(wgrad ? 1 : 0, p, dtype, s);
}

static int depthwise_plan_any(const void* x, long long ld_x, void* y, long long ld_y, const float* w,
                              const float* bias, const int* step_nbr, const int* step_k,
                              const int* rows, const int* tile_nk, int num_tiles, int tile_rows,
                              int K, int C, int kflip, int relu, int dtype, cudaStream_t s,
                              const void* dy, long long ld_dy, float* dw) {
  if (num_tiles < 0 || K < 1 || C < 1 || tile_rows < 1) return kErrInvalidArg;
  if (num_tiles == 0) return kOk;
  DepthwiseParams p{};
  p.x = x; p.y = y; p.w = w; p.bias = bias; p.dy = dy; p.ld_dy = ld_dy; p.dw = dw;
  p.step_nbr = step_nbr; p.step_k = step_k; p.rows = rows; p.tile_nk = tile_nk;
  p.num_tiles = num_tiles; p.tile_rows = tile_rows;
  p.ld_x = ld_x; p.ld_y = ld_y;
  p.K = K; p.C = C; p.kflip = kflip; p.relu = relu;
  return dw_dispatch(dw != nullptr ? 3 : 2, p, dtype, s);
}

int depthwise_plan_launch(const void* x, long long ld_x, void* y, long long ld_y, const float* w,
                          const float* bias, const int* step_nbr, const int* step_k,
                          const int* rows, const int* tile_nk, int num_tiles, int tile_rows, int K,
                          int C, int kflip, int relu, int dtype, cudaStream_t s) {
  return depthwise_plan_any(x, ld_x, y, ld_y, w, bias, step_nbr, step_k, rows, tile_nk, num_tiles,
                            tile_rows, K, C, kflip, relu, dtype, s, nullptr, 0, nullptr);
}

int depthwise_plan_wgrad_launch(const void* x, long long ld_x, const void* dy, long long ld_dy,
                                float* dw, const int* step_nbr, const int* step_k, const int* rows,
                                const int* tile_nk, int num_tiles, int tile_rows, int K, int C,
                                int dtype, cudaStream_t s) {
  if (dw == nullptr || dy == nullptr) return kErrInvalidArg;
  return depthwise_plan_any(x, ld_x, nullptr, 0, nullptr, nullptr, step_nbr, step_k, rows, tile_nk,
                            num_tiles, tile_rows, K, C, 0, 0, dtype, s, dy, ld_dy, dw);
}

}  // namespace wcn
