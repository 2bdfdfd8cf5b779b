// SPDX-License-Identifier: Apache-2.0
// Coordinate-set operations of the sparse-conv path, fully on the device:
// the output coordinates of strided convs (floor division + unique), of generative convs
// (union of coord + offset_k over all kernel offsets + unique) and Voxels.unique(). One chain for
// all three:
//   key generation (floor-div, + offset, range check, packed sortable 64-bit key)
//   -> radix sort of (key, source index) pairs  [CUB DeviceRadixSort: library sort]
//   -> head flags / block counts -> block scan -> compaction + per-batch offsets
// Nothing here synchronises with the host: the result rows go into an upper-bound sized buffer and
// (per-batch offsets, total, status) into a small device array the caller reads back when it needs
// the size. Output rows are sorted by (batch, x, y, z): deterministic, batch-contiguous.
//
// Behaviour follows (re-implemented, not copied):
//   stride_coords   warpconvnet/geometry/coords/ops/stride.py:18-56 (floor division, unique, batch sort)
//   expand_coords   warpconvnet/geometry/coords/ops/expand.py:17-75
//   unique          warpconvnet/geometry/types/voxels.py:271-278, coords/ops/voxel.py:112-151
// The reference dedups with a racy hash insert and argsorts the batch column (row order
// unspecified); here the order is the sorted key order and the representative of duplicate rows
// is the smallest source index.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace wcn {

constexpr int kCoordBias = 1 << 17;  // maps [-131072, 131071] to [0, 262143]
constexpr int kUniqBlock = 256;
constexpr int kUniqItems = 4;  // keys per thread in the flag / compaction passes
constexpr int kUniqTile = kUniqBlock * kUniqItems;

__device__ __forceinline__ int floor_div(int a, int s) {
  return a >= 0 ? a / s : -((-a + s - 1) / s);
}

__device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
  return (unsigned)b <= 511u && (unsigned)(x + kCoordBias) < 262144u &&
         (unsigned)(y + kCoordBias) < 262144u && (unsigned)(z + kCoordBias) < 262144u;
}

__device__ __forceinline__ unsigned long long parked_key(int n_batches) {
  return ((unsigned long long)n_batches << 54) | ((1ull << 54) - 1ull);
}

__device__ __forceinline__ unsigned long long sortable_key(int b, int x, int y, int z) {
  return ((unsigned long long)b << 54) | ((unsigned long long)(x + kCoordBias) << 36) |
         ((unsigned long long)(y + kCoordBias) << 18) | (unsigned long long)(z + kCoordBias);
}

// keys[k * n + i] = key(floor(c_i / stride) + off_k), idx[...] = i; meta[status] |= 2 when a
// generated coordinate leaves the packed range (same contract as the hash insert kernel)
__global__ void __launch_bounds__(256)
coords_keys_kernel(const int4* __restrict__ bcoords, int n, int sx, int sy, int sz,
                   const int* __restrict__ offs, int K, int n_batches,
                   unsigned long long* __restrict__ keys, int* __restrict__ idx,
                   int* __restrict__ status) {
  pdl_begin();
  const long long total = (long long)n * K;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int k = (int)(t / n);
  const int i = (int)(t - (long long)k * n);
  const int4 c = __ldg(bcoords + i);
  int x = floor_div(c.y, sx), y = floor_div(c.z, sy), z = floor_div(c.w, sz);
  if (offs != nullptr) {
    x += __ldg(offs + 3 * k);
    y += __ldg(offs + 3 * k + 1);
    z += __ldg(offs + 3 * k + 2);
  }
  if (!coord_in_range(c.x, x, y, z) || c.x >= n_batches) {
    atomicOr(status, 2);
    // park the row behind the last batch item (batch field = n_batches) so it cannot alias a
    // valid cell and sorts last inside the 54 + batch_bits sorted bits
    keys[t] = parked_key(n_batches);
  } else {
    keys[t] = sortable_key(c.x, x, y, z);
  }
  if (idx != nullptr) idx[t] = i;
}

// block_counts[b] = number of distinct keys that START inside tile b of the sorted key array
__global__ void __launch_bounds__(kUniqBlock)
coords_heads_kernel(const unsigned long long* __restrict__ keys, long long n, int n_batches,
                    int* __restrict__ block_counts) {
  pdl_begin();
  const unsigned long long parked = parked_key(n_batches);
  __shared__ int s_warp[kUniqBlock / 32];
  const long long base = (long long)blockIdx.x * kUniqTile + (long long)threadIdx.x * kUniqItems;
  int cnt = 0;
  unsigned long long prev = (base > 0 && base <= n) ? __ldg(keys + base - 1) : 0ull;
#pragma unroll
  for (int j = 0; j < kUniqItems; ++j) {
    const long long t = base + j;
    if (t < n) {
      const unsigned long long k = __ldg(keys + t);
      if (k != parked && (t == 0 || k != prev)) ++cnt;
      prev = k;
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kUniqBlock / 32; ++w) t += s_warp[w];
    block_counts[blockIdx.x] = t;
  }
}

// exclusive scan of block_counts[0..nb) in place (single block); total -> meta[n_batches] and
// meta[n_batches + 1] (the "total" word); the per-batch offsets are filled by the compaction pass
__global__ void __launch_bounds__(1024)
coords_scan_kernel(int* __restrict__ block_counts, int nb, int* __restrict__ meta, int n_batches) {
  pdl_begin();
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? block_counts[i] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int warp_off = 0;
    for (int w = 0; w < warp; ++w) warp_off += s_warp[w];
    const int carry = s_carry;
    if (i < nb) block_counts[i] = carry + warp_off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + warp_off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    meta[n_batches] = s_carry;      // offsets[n_batches] = number of unique rows
    meta[n_batches + 1] = s_carry;  // total
  }
  // batch items with no rows at all keep offsets consistent: pre-fill every offset with the
  // total, the compaction pass overwrites offsets[b] for every b <= (last batch seen)
  for (int b = threadIdx.x; b < n_batches; b += blockDim.x) meta[b] = 0x7fffffff;
}

__device__ __forceinline__ int4 unpack_sortable(unsigned long long k) {
  return make_int4((int)(k >> 54), (int)((k >> 36) & 0x3FFFF) - kCoordBias,
                   (int)((k >> 18) & 0x3FFFF) - kCoordBias, (int)(k & 0x3FFFF) - kCoordBias);
}

// Writes the unique rows (ascending key order), the source index of each one's first occurrence
// and offsets[b] = first output row of batch item b (atomicMin over the heads of b and of every
// later batch item, so empty items inherit the next item's start).
__global__ void __launch_bounds__(kUniqBlock)
coords_compact_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ idx,
                      long long n, const int* __restrict__ block_prefix,
                      int4* __restrict__ out_coords, int* __restrict__ first_index,
                      int* __restrict__ meta, int n_batches) {
  pdl_begin();
  __shared__ int s_warp[kUniqBlock / 32];
  const unsigned long long parked = parked_key(n_batches);
  const long long base = (long long)blockIdx.x * kUniqTile + (long long)threadIdx.x * kUniqItems;
  unsigned long long k[kUniqItems];
  bool head[kUniqItems];
  int cnt = 0;
  unsigned long long prev = (base > 0 && base <= n) ? __ldg(keys + base - 1) : 0ull;
#pragma unroll
  for (int j = 0; j < kUniqItems; ++j) {
    const long long t = base + j;
    head[j] = false;
    k[j] = parked;
    if (t < n) {
      k[j] = __ldg(keys + t);
      head[j] = k[j] != parked && (t == 0 || k[j] != prev);
      prev = k[j];
    }
    cnt += head[j] ? 1 : 0;
  }
  // exclusive scan of cnt over the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int pos = block_prefix[blockIdx.x] + incl - cnt;
  for (int w = 0; w < warp; ++w) pos += s_warp[w];
#pragma unroll
  for (int j = 0; j < kUniqItems; ++j) {
    if (head[j]) {
      const int4 c = unpack_sortable(k[j]);
      out_coords[pos] = c;
      if (first_index != nullptr) first_index[pos] = __ldg(idx + base + j);
      // first row of a batch item: the previous key belongs to another item (or there is none)
      const long long t = base + j;
      const int pb = t > 0 ? (int)(__ldg(keys + t - 1) >> 54) : -1;
      if (pb != c.x)
        for (int b = pb + 1; b <= c.x && b < n_batches; ++b) atomicMin(meta + b, pos);
      ++pos;
    }
  }
}

// offsets of batch items after the last one that has rows: the total
__global__ void coords_offsets_tail_kernel(int* __restrict__ meta, int n_batches) {
  pdl_begin();
  const int total = meta[n_batches];
  for (int b = threadIdx.x; b < n_batches; b += blockDim.x)
    if (meta[b] == 0x7fffffff) meta[b] = total;
}

static inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cuda_ok2() { return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda; }

static size_t cub_pairs_bytes(long long n) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr,
                                  n > 0 ? n : 1, 0, 64, (cudaStream_t)0);
  return temp;
}

// workspace: [keys_in n*8][keys_out n*8][idx_in n*4][idx_out n*4][block counts][cub temp]
size_t coords_unique_workspace_bytes(long long n_keys) {
  const size_t n = (size_t)(n_keys > 0 ? n_keys : 1);
  const size_t nb = (n + kUniqTile - 1) / kUniqTile;
  return 2 * align_up_sz(n * 8, 256) + 2 * align_up_sz(n * 4, 256) + align_up_sz(nb * 4, 256) +
         align_up_sz(cub_pairs_bytes((long long)n), 256) + 256;
}

int coords_unique(const int* bcoords, int n, int sx, int sy, int sz, const int* offsets3, int K,
                  int n_batches, int* out_coords, int* first_index, int* meta, void* workspace,
                  size_t ws_bytes, cudaStream_t s) {
  if (n < 0 || K < 1 || sx < 1 || sy < 1 || sz < 1 || n_batches < 1 || n_batches > 512)
    return kErrInvalidArg;
  const long long total = (long long)n * K;
  if (total > 0x7fffffffll) return kErrInvalidArg;
  if (cudaMemsetAsync(meta, 0, (size_t)(n_batches + 3) * 4, s) != cudaSuccess) return kErrCuda;
  if (total == 0) return kOk;
  if (ws_bytes < coords_unique_workspace_bytes(total)) return kErrWorkspace;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const size_t kb = align_up_sz((size_t)total * 8, 256), ib = align_up_sz((size_t)total * 4, 256);
  const int nb = (int)((total + kUniqTile - 1) / kUniqTile);
  unsigned long long* keys_in = reinterpret_cast<unsigned long long*>(ws);
  unsigned long long* keys_out = reinterpret_cast<unsigned long long*>(ws + kb);
  int* idx_in = reinterpret_cast<int*>(ws + 2 * kb);
  int* idx_out = reinterpret_cast<int*>(ws + 2 * kb + ib);
  int* block_counts = reinterpret_cast<int*>(ws + 2 * kb + 2 * ib);
  void* temp = ws + 2 * kb + 2 * ib + align_up_sz((size_t)nb * 4, 256);
  size_t temp_bytes = ws_bytes - (2 * kb + 2 * ib + align_up_sz((size_t)nb * 4, 256));
  int* status = meta + n_batches + 2;
  const bool want_idx = first_index != nullptr;
  wcn_launch(coords_keys_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, 
      reinterpret_cast<const int4*>(bcoords), n, sx, sy, sz, offsets3, K, n_batches, keys_in,
      want_idx ? idx_in : nullptr, status);
  count_launch();
  // the batch field holds 0..n_batches (n_batches = parked rows): sort only the bits in use
  int batch_bits = 1;
  while ((1 << batch_bits) <= n_batches) ++batch_bits;
  const int end_bit = 54 + batch_bits;
  cudaError_t e;
  if (want_idx)
    e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, idx_out,
                                        total, 0, end_bit, s);
  else
    e = cub::DeviceRadixSort::SortKeys(temp, temp_bytes, keys_in, keys_out, total, 0, end_bit, s);
  if (e != cudaSuccess) return kErrCuda;
  wcn_launch(coords_heads_kernel, dim3(nb), dim3(kUniqBlock), 0, s, keys_out, total, n_batches, block_counts);
  count_launch();
  wcn_launch(coords_scan_kernel, dim3(1), dim3(1024), 0, s, block_counts, nb, meta, n_batches);
  count_launch();
  wcn_launch(coords_compact_kernel, dim3(nb), dim3(kUniqBlock), 0, s, keys_out, want_idx ? idx_out : nullptr, total,
                                                  block_counts, reinterpret_cast<int4*>(out_coords),
                                                  first_index, meta, n_batches);
  count_launch();
  wcn_launch(coords_offsets_tail_kernel, dim3(1), dim3(256), 0, s, meta, n_batches);
  count_launch();
  return cuda_ok2();
}

}  // namespace wcn
