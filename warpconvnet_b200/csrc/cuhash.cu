// SPDX-License-Identifier: Apache-2.0
// Kernel-map construction for sparse convolution: packed-coordinate hash table, per-offset
// neighbour probe, deterministic CSR compaction, reverse table, and the mask-sorted tile tables
// the tensor-core kernels consume. All integer work, HBM/L2-bound; results are bit-exact against
// oracle/kernel_map.py.
//
// Behaviour follows (re-implemented, not copied):
//   key packing / Splitmix64 / linear probing  warpconvnet/csrc/include/cuhash/hash_functions.cuh:29-84,
//                                              hash_table.cuh:36-108
//   offset enumeration k -> (i,j,l)            warpconvnet/csrc/include/cuhash/kernel_map.cuh:34-54
//   found[K,M] -> counts -> CSR                warpconvnet/csrc/cuhash_kernel_map.cu:93-134,508-599
//   pair mask / argsort / reverse table        warpconvnet/csrc/mask_data_kernels.cu:23-220
// Differences by design: one thread owns a query and walks all K offsets (the query row is
// loaded and packed once instead of K times), per-offset counts are warp-aggregated, and the CSR
// is emitted in ascending output-row order by a two-pass block scan (the reference reserves slots
// with atomics, so its row order is non-deterministic).
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace wcn {

constexpr int kMapBlock = 256;  // queries per block in the search / scatter passes
constexpr uint64_t kValidBit = 1ull << 63;
constexpr uint64_t kEmptyKey = 0ull;

__host__ __device__ __forceinline__ uint64_t pack_key(int b, int x, int y, int z) {
  return kValidBit | ((uint64_t)(b & 0x1FF) << 54) | ((uint64_t)(x & 0x3FFFF) << 36) |
         ((uint64_t)(y & 0x3FFFF) << 18) | (uint64_t)(z & 0x3FFFF);
}

// true when (b, x, y, z) fits the packed key: a probe outside the range is a MISS (masking it
// would alias a cell 2^18 voxels away by wrap-around)
__device__ __forceinline__ bool key_in_range(int b, int x, int y, int z) {
  return (unsigned)b <= 511u && (unsigned)(x + 131072) < 262144u &&
         (unsigned)(y + 131072) < 262144u && (unsigned)(z + 131072) < 262144u;
}

__device__ __forceinline__ uint32_t splitmix_slot(uint64_t key, uint32_t mask) {
  key ^= key >> 30;
  key *= 0xBF58476D1CE4E5B9ull;
  key ^= key >> 27;
  key *= 0x94D049BB133111EBull;
  key ^= key >> 31;
  return (uint32_t)key & mask;
}

__device__ __forceinline__ int table_lookup(const uint64_t* __restrict__ keys,
                                            const int* __restrict__ values, uint64_t key,
                                            uint32_t mask) {
  uint32_t slot = splitmix_slot(key, mask);
  for (uint32_t attempts = 0; attempts <= mask; ++attempts) {
    const uint64_t k = __ldg(keys + slot);
    if (k == key) return __ldg(values + slot);
    if (k == kEmptyKey) return -1;
    slot = (slot + 1) & mask;
  }
  return -1;
}

// --------------------------------------------------------------------------------------------
// hash table build
// --------------------------------------------------------------------------------------------
__global__ void hash_insert_kernel(uint64_t* __restrict__ keys, int* __restrict__ values,
                                   const int4* __restrict__ coords, int n, uint32_t mask,
                                   int* __restrict__ status) {
  pdl_begin();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = __ldg(coords + i);  // (batch, x, y, z): one coalesced 16-byte load
  if (c.x < 0 || c.x > 511 || c.y < -131072 || c.y > 131071 || c.z < -131072 || c.z > 131071 ||
      c.w < -131072 || c.w > 131071) {
    atomicOr(status, 2);  // out-of-range coordinate (the reference raises ValueError on the host)
    return;
  }
  const uint64_t key = pack_key(c.x, c.y, c.z, c.w);
  uint32_t slot = splitmix_slot(key, mask);
  for (uint32_t attempts = 0; attempts <= mask; ++attempts) {
    const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(keys + slot),
                                              (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (prev == key) atomicOr(status, 4);  // duplicate coordinate (informational, see below)
    if (prev == kEmptyKey || prev == key) {
      // value = insertion index; duplicates keep the SMALLEST index, so the table is
      // deterministic (the reference keeps whichever thread wins the race). Empty value slots
      // are pre-filled with 0x7f7f7f7f by hash_prepare, so atomicMin works for both cases.
      atomicMin(values + slot, i);
      return;
    }
    slot = (slot + 1) & mask;
  }
  atomicOr(status, 1);  // table full
}

__global__ void hash_search_kernel(const uint64_t* __restrict__ keys,
                                   const int* __restrict__ values,
                                   const int4* __restrict__ queries, int* __restrict__ results,
                                   int n, uint32_t mask) {
  pdl_begin();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = __ldg(queries + i);
  results[i] = key_in_range(c.x, c.y, c.z, c.w)
                   ? table_lookup(keys, values, pack_key(c.x, c.y, c.z, c.w), mask) : -1;
}

// --------------------------------------------------------------------------------------------
// kernel map: probe all K offsets of every output coordinate
// --------------------------------------------------------------------------------------------
// pair_table[k][m] = index of input voxel at out[m]*stride + offset[k], or -1
// block_counts[k][blockIdx.x] = number of hits of offset k among this block's 256 queries
// mask_keys[m] = bit k set iff offset k hit (bits folded modulo 64 when K > 64)
__global__ void __launch_bounds__(kMapBlock)
kernel_map_search_kernel(const uint64_t* __restrict__ keys, const int* __restrict__ values,
                         uint32_t mask, const int4* __restrict__ out_coords, int M,
                         const int* __restrict__ offs, int K, int sx, int sy, int sz,
                         int* __restrict__ pair_table, int* __restrict__ block_counts,
                         unsigned long long* __restrict__ mask_keys) {
  pdl_begin();
  extern __shared__ int s_mem[];  // [K] counts, then [3K] offsets
  int* s_cnt = s_mem;
  int* s_off = s_mem + K;
  for (int i = threadIdx.x; i < K; i += blockDim.x) s_cnt[i] = 0;
  for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) s_off[i] = offs[i];
  __syncthreads();

  const int m = blockIdx.x * kMapBlock + threadIdx.x;
  const bool live = m < M;
  int4 c = make_int4(0, 0, 0, 0);
  if (live) c = __ldg(out_coords + m);
  const int bx = c.y * sx, by = c.z * sy, bz = c.w * sz;
  unsigned long long bits = 0ull;
  const int lane = threadIdx.x & 31;

  for (int k0 = 0; k0 < K; k0 += 4) {
    // four independent probes in flight per thread
    int res[4];
    uint32_t slot[4];
    uint64_t key[4];
    uint64_t got[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u;
      res[u] = -1;
      const int qx = k < K ? bx + s_off[3 * k] : 0, qy = k < K ? by + s_off[3 * k + 1] : 0,
                qz = k < K ? bz + s_off[3 * k + 2] : 0;
      if (k < K && live && key_in_range(c.x, qx, qy, qz)) {
        key[u] = pack_key(c.x, qx, qy, qz);
        slot[u] = splitmix_slot(key[u], mask);
        got[u] = __ldg(keys + slot[u]);
      } else {
        key[u] = 1;  // never matches: stored keys have the valid bit set
        slot[u] = 0;
        got[u] = kEmptyKey;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t attempts = 0;
      uint64_t g = got[u];
      uint32_t s = slot[u];
      while (g != kEmptyKey && attempts <= mask) {
        if (g == key[u]) {
          res[u] = __ldg(values + s);
          break;
        }
        s = (s + 1) & mask;
        g = __ldg(keys + s);
        ++attempts;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u;
      if (k < K) {  // warp-uniform
        const bool hit = res[u] >= 0;
        if (live) pair_table[(size_t)k * M + m] = res[u];
        if (hit) bits ^= 1ull << (k & 63);
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (lane == 0 && ballot) atomicAdd(&s_cnt[k], __popc(ballot));  // warp-aggregated
      }
    }
  }
  if (live && mask_keys != nullptr) mask_keys[m] = bits;
  __syncthreads();
  if (block_counts != nullptr) {
    for (int i = threadIdx.x; i < K; i += blockDim.x)
      block_counts[(size_t)i * gridDim.x + blockIdx.x] = s_cnt[i];
  }
}

// Submanifold variant (in == out coordinates, odd kernel, stride 1): offset K-1-k is the mirror
// of offset k, so a hit "voxel v sits at m + off_k" also says "voxel m sits at v + off_{K-1-k}".
// Only the first K/2 offsets (+ the centre) are probed and every hit writes both entries. Own
// probes are written hit or miss (coalesced stores), so only the mirrored rows k > K/2 of the
// table need the -1 pre-fill. Requires unique coordinates: when the insert kernel flagged a
// duplicate (status bit 2) every offset is probed instead (no mirrors).
// (Producing the block counts and row masks in this pass with atomics on the mirror rows was
// measured SLOWER than the separate coalesced pass of kernel_map_stats_kernel: 55.7 us vs
// 32.8 + 13.2 us on C3, profiles/r2_kernel_map_chain.md.)
__global__ void __launch_bounds__(kMapBlock)
kernel_map_search_sym_kernel(const uint64_t* __restrict__ keys, const int* __restrict__ values,
                             uint32_t mask, const int4* __restrict__ coords, int M,
                             const int* __restrict__ offs, int K, const int* __restrict__ status,
                             int* __restrict__ pair_table) {
  pdl_begin();
  extern __shared__ int s_off[];  // [3K]
  for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) s_off[i] = offs[i];
  __syncthreads();
  const int m = blockIdx.x * kMapBlock + threadIdx.x;
  if (m >= M) return;
  const bool has_dups = (__ldg(status) & 4) != 0;
  const int k_end = has_dups ? K : K / 2 + 1;
  const int4 c = __ldg(coords + m);
  // kSymBatch probes in flight per thread: first the key slots, then the values of the hits
  // (two rounds of independent loads instead of a dependent chain per offset)
  constexpr int kSymBatch = 4;  // 8 measured slower (41.9 vs 36.5 us on C3: registers, occupancy)
  for (int k0 = 0; k0 < k_end; k0 += kSymBatch) {
    uint32_t slot[kSymBatch];
    uint64_t key[kSymBatch], got[kSymBatch];
#pragma unroll
    for (int u = 0; u < kSymBatch; ++u) {
      const int k = k0 + u;
      const int qx = k < k_end ? c.y + s_off[3 * k] : 0, qy = k < k_end ? c.z + s_off[3 * k + 1] : 0,
                qz = k < k_end ? c.w + s_off[3 * k + 2] : 0;
      if (k < k_end && key_in_range(c.x, qx, qy, qz)) {
        key[u] = pack_key(c.x, qx, qy, qz);
        slot[u] = splitmix_slot(key[u], mask);
        got[u] = __ldg(keys + slot[u]);
      } else {
        key[u] = 1;
        slot[u] = 0;
        got[u] = kEmptyKey;
      }
    }
    // resolve collisions (rare at load factor <= 0.5), leaving slot[u] = matching slot or ~0u
#pragma unroll
    for (int u = 0; u < kSymBatch; ++u) {
      uint32_t attempts = 0;
      uint64_t g = got[u];
      uint32_t sl = slot[u];
      while (g != kEmptyKey && g != key[u] && attempts <= mask) {
        sl = (sl + 1) & mask;
        g = __ldg(keys + sl);
        ++attempts;
      }
      slot[u] = (g == key[u]) ? sl : 0xFFFFFFFFu;
    }
    int val[kSymBatch];
#pragma unroll
    for (int u = 0; u < kSymBatch; ++u)
      val[u] = slot[u] != 0xFFFFFFFFu ? __ldg(values + slot[u]) : -1;
#pragma unroll
    for (int u = 0; u < kSymBatch; ++u) {
      const int k = k0 + u;
      if (k < k_end) {
        pair_table[(size_t)k * M + m] = val[u];
        if (val[u] >= 0 && !has_dups && k < K / 2) pair_table[(size_t)(K - 1 - k) * M + val[u]] = m;
      }
    }
  }
}

// block_counts[k][block] = hits of offset k among the block's 256 rows; mask_keys[m] = offset
// bitmask of row m. One coalesced pass over the pair table, 32 offsets at a time: the 32 loads of
// a thread are independent (all in flight together) and the block synchronises twice per chunk
// instead of twice per offset (the per-offset form chained K dependent L2 round trips: 20 us for
// the 21.6 MB table of C3).
__global__ void __launch_bounds__(kMapBlock)
kernel_map_stats_kernel(const int* __restrict__ pair_table, int K, int M,
                        int* __restrict__ block_counts,
                        unsigned long long* __restrict__ mask_keys) {
  pdl_begin();
  __shared__ int s_cnt[32][kMapBlock / 32];
  const int m = blockIdx.x * kMapBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long bits = 0ull;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int kc = min(32, K - k0);
    unsigned w = 0u;
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const bool hit = kk < kc && m < M && __ldg(pair_table + (size_t)(k0 + kk) * M + m) >= 0;
      w |= (hit ? 1u : 0u) << kk;
    }
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const unsigned ballot = __ballot_sync(0xffffffffu, (w >> kk) & 1u);
      if (lane == 0) s_cnt[kk][warp] = __popc(ballot);
    }
    __syncthreads();
    if (threadIdx.x < kc) {
      int t = 0;
#pragma unroll
      for (int ww = 0; ww < kMapBlock / 32; ++ww) t += s_cnt[threadIdx.x][ww];
      block_counts[(size_t)(k0 + threadIdx.x) * gridDim.x + blockIdx.x] = t;
    }
    __syncthreads();
    bits ^= (unsigned long long)w << (k0 & 63);  // bit (k & 63) of offset k, folded for K > 64
  }
  if (m < M && mask_keys != nullptr) mask_keys[m] = bits;
}

// exclusive scan of block_counts[k][0..nb) in place, total into counts[k]; one block per offset
__global__ void __launch_bounds__(256)
block_scan_kernel(int* __restrict__ block_counts, int nb, int* __restrict__ counts) {
  pdl_begin();
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  int* row = block_counts + (size_t)blockIdx.x * nb;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += 256) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? row[i] : 0;
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
    if (i < nb) row[i] = carry + warp_off + incl - v;
    __syncthreads();
    if (threadIdx.x == 255) s_carry = carry + warp_off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[blockIdx.x] = s_carry;
}

// offsets[0..K] = exclusive scan of counts; single small block
__global__ void offsets_kernel(const int* __restrict__ counts, int K, int* __restrict__ offsets) {
  pdl_begin();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k < K; ++k) {
      offsets[k] = acc;
      acc += counts[k];
    }
    offsets[K] = acc;
  }
}

// CSR emission in ascending output-row order inside every offset (deterministic); 32 offsets per
// chunk with independent loads, two block synchronisations per chunk (see the stats kernel)
__global__ void __launch_bounds__(kMapBlock)
kernel_map_scatter_kernel(const int* __restrict__ pair_table, const int* __restrict__ block_prefix,
                          const int* __restrict__ offsets, int* __restrict__ in_maps,
                          int* __restrict__ out_maps, int K, int M) {
  pdl_begin();
  __shared__ int s_cnt[32][kMapBlock / 32];
  __shared__ int s_base[32];
  const int m = blockIdx.x * kMapBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lane_mask = (1u << lane) - 1u;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int kc = min(32, K - k0);
    int v[32];
#pragma unroll
    for (int kk = 0; kk < 32; ++kk)
      v[kk] = (kk < kc && m < M) ? __ldg(pair_table + (size_t)(k0 + kk) * M + m) : -1;
    if (threadIdx.x < kc)
      s_base[threadIdx.x] = offsets[k0 + threadIdx.x] +
                            block_prefix[(size_t)(k0 + threadIdx.x) * gridDim.x + blockIdx.x];
    int rank[32];
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const unsigned ballot = __ballot_sync(0xffffffffu, v[kk] >= 0);
      if (lane == 0) s_cnt[kk][warp] = __popc(ballot);
      rank[kk] = __popc(ballot & lane_mask);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      if (v[kk] >= 0) {
        int r = rank[kk];
        for (int ww = 0; ww < warp; ++ww) r += s_cnt[kk][ww];
        const int pos = s_base[kk] + r;
        in_maps[pos] = v[kk];
        out_maps[pos] = m;
      }
    }
    __syncthreads();
  }
}

// rev[k][in] = out  for every valid pair (rev must be pre-filled with -1)
__global__ void reverse_table_kernel(const int* __restrict__ pair_table, int K, int M,
                                     int* __restrict__ rev, int n_in) {
  pdl_begin();
  const long long total = (long long)K * M;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int v = __ldg(pair_table + t);
    if (v >= 0) {
      const int k = (int)(t / M);
      const int m = (int)(t - (long long)k * M);
      rev[(size_t)k * n_in + v] = m;
    }
  }
}

// table[k][row_maps[j]] = val_maps[j] for every CSR pair j of offset k (table pre-filled with -1)
__global__ void csr_to_table_kernel(const int* __restrict__ val_maps,
                                    const int* __restrict__ row_maps,
                                    const int* __restrict__ offsets, int K, int n_rows,
                                    int* __restrict__ table) {
  pdl_begin();
  extern __shared__ int s_offs[];  // [K+1]
  for (int i = threadIdx.x; i <= K; i += blockDim.x) s_offs[i] = offsets[i];
  __syncthreads();
  const int L = s_offs[K];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) {
    int lo = 0, hi = K;  // largest k with offsets[k] <= j
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_offs[mid] <= j) lo = mid; else hi = mid;
    }
    table[(size_t)lo * n_rows + __ldg(row_maps + j)] = __ldg(val_maps + j);
  }
}

__global__ void mask_keys_kernel(const int* __restrict__ table, int K, int M,
                                 unsigned long long* __restrict__ keys) {
  pdl_begin();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  unsigned long long bits = 0ull;
  for (int k = 0; k < K; ++k)
    if (__ldg(table + (size_t)k * M + m) >= 0) bits ^= 1ull << (k & 63);
  keys[m] = bits;
}

__global__ void iota_kernel(int* __restrict__ v, int n) {
  pdl_begin();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// One block per tile (tile_rows = 128 or 256 rows) of the mask-sorted order: compacts the offsets
// that are active anywhere in the tile into the tile's step list and stores each step's
// neighbour rows contiguously, in the order the GEMM kernel consumes them.
__global__ void __launch_bounds__(256)
build_tiles_kernel(const int* __restrict__ table, int K, int M, const int* __restrict__ sorted_rows,
                   int tile_rows, int* __restrict__ step_nbr, int* __restrict__ step_k,
                   int* __restrict__ rows_padded, int* __restrict__ tile_nk,
                   const unsigned long long* __restrict__ masks) {
  pdl_begin();
  // 32 offsets per chunk: the 32 table loads of a thread are independent (one round trip instead
  // of a dependent chain with a block barrier per offset: 20.4 us on C3), the tile's union mask of
  // the chunk is ONE shared-memory OR, and every thread then writes its neighbours of the active
  // offsets at their compacted step positions.
  __shared__ unsigned s_union;
  const int tile = blockIdx.x;
  const int pos = tile * tile_rows + threadIdx.x;  // blockDim.x == tile_rows
  const int row = pos < M ? __ldg(sorted_rows + pos) : -1;
  rows_padded[pos] = row;
  // masks (optional, K <= 64): bit k of masks[row] <=> table[k][row] >= 0. A row then only loads
  // the entries it has — the table reads are scattered 4-byte accesses, one 32-byte sector each,
  // and a surface voxel has ~9 of the 27 (C3-S: 5.4 M -> 1.8 M sector reads, 16 -> 9 us)
  unsigned long long mybits = ~0ull;
  if (masks != nullptr) mybits = row >= 0 ? __ldg(masks + row) : 0ull;
  int n = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int kc = min(32, K - k0);
    if (threadIdx.x == 0) s_union = 0u;
    __syncthreads();
    int v[32];
    unsigned hits = 0u;
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      v[kk] = (kk < kc && row >= 0 && ((mybits >> ((k0 + kk) & 63)) & 1ull))
                  ? __ldg(table + (size_t)(k0 + kk) * M + row) : -1;
      hits |= (v[kk] >= 0 ? 1u : 0u) << kk;
    }
    const unsigned warp_or = __reduce_or_sync(0xffffffffu, hits);
    if ((threadIdx.x & 31) == 0 && warp_or) atomicOr(&s_union, warp_or);
    __syncthreads();
    const unsigned uni = s_union;
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      if ((uni >> kk) & 1u) {  // block-uniform
        const size_t step = (size_t)tile * K + n + __popc(uni & ((1u << kk) - 1u));
        step_nbr[step * tile_rows + threadIdx.x] = v[kk];
        if (threadIdx.x == 0) step_k[step] = k0 + kk;
      }
    }
    n += __popc(uni);
    __syncthreads();  // s_union is reset by the next chunk
  }
  if (threadIdx.x == 0) tile_nk[tile] = n;
}

// tile_cum[0..n] = exclusive prefix sum of tile_nk[0..n); single block
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const int* __restrict__ tile_nk, int n, int* __restrict__ tile_cum, int U,
                 int n_ctas, int* __restrict__ cta_units) {
  pdl_begin();
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? tile_nk[i] : 0;
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
    if (i < n) tile_cum[i] = carry + warp_off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + warp_off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_cum[n] = s_carry;
  // Optional: the gather-GEMM kernel's per-CTA unit ranges (unit = 128 rows, U per tile; cost of
  // a unit = its tile's step count), so its prologue reads two words instead of running a
  // 3-round search over tile_cum. Same rule as conv_fwd.cu: first unit whose start cost reaches
  // S*b/G, moved one down when that boundary is nearer.
  if (cta_units == nullptr || n_ctas < 1) return;
  __syncthreads();  // tile_cum written by this block is visible to all its threads
  const long long S = (long long)U * s_carry;
  const int total_units = U * n;
  auto cost = [&](int u) {
    const int t = u / U, r = u - t * U;
    const int c0 = tile_cum[t];
    return (long long)U * c0 + (r ? (long long)r * (tile_cum[t + 1] - c0) : 0ll);
  };
  for (int b = threadIdx.x; b <= n_ctas; b += blockDim.x) {
    int u = total_units;
    if (b < n_ctas) {
      const long long target = S * b / n_ctas;
      int lo = 0, hi = total_units;  // first u in [0, total] with cost(u) >= target
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cost(mid) >= target) hi = mid; else lo = mid + 1;
      }
      u = lo;
      if (b > 0 && u > 0 && target - cost(u - 1) < cost(u) - target) --u;
    }
    cta_units[b] = u;
  }
}

// --------------------------------------------------------------------------------------------
// host launchers
// --------------------------------------------------------------------------------------------
static inline int cuda_ok() { return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda; }
static inline bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }

// keys = empty (0), values = 0x7f7f7f7f (+inf for the atomicMin of the insert kernel): one launch
// instead of two memset nodes
__global__ void hash_prepare_kernel(uint4* __restrict__ keys16, uint4* __restrict__ values16,
                                    int n_keys16, int n_values16) {
  pdl_begin();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_keys16; i += stride)
    keys16[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_values16; i += stride)
    values16[i] = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
}

int hash_prepare(uint64_t* keys, int* values, int capacity, cudaStream_t s) {
  if (!is_pow2(capacity)) return kErrInvalidArg;
  if (capacity >= 4 && (reinterpret_cast<uintptr_t>(keys) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(values) & 15) == 0) {
    const int nk = capacity / 2, nv = capacity / 4;  // 16-byte words
    int blocks = (nk + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    wcn_launch(hash_prepare_kernel, dim3(blocks), dim3(256), 0, s, reinterpret_cast<uint4*>(keys),
                                               reinterpret_cast<uint4*>(values), nk, nv);
    count_launch();
    return cuda_ok();
  }
  if (cudaMemsetAsync(keys, 0, (size_t)capacity * 8, s) != cudaSuccess) return kErrCuda;
  if (cudaMemsetAsync(values, 0x7F, (size_t)capacity * 4, s) != cudaSuccess) return kErrCuda;
  return kOk;
}

int hash_insert(uint64_t* keys, int* values, const int* coords, int n, int capacity, int* status,
                cudaStream_t s) {
  if (!is_pow2(capacity) || n < 0) return kErrInvalidArg;
  if (n == 0) return kOk;
  wcn_launch(hash_insert_kernel, dim3((n + 255) / 256), dim3(256), 0, s, keys, values,
                                                    reinterpret_cast<const int4*>(coords), n,
                                                    (uint32_t)(capacity - 1), status);
  count_launch();
  return cuda_ok();
}

int hash_search(const uint64_t* keys, const int* values, const int* queries, int* results, int n,
                int capacity, cudaStream_t s) {
  if (!is_pow2(capacity) || n < 0) return kErrInvalidArg;
  if (n == 0) return kOk;
  wcn_launch(hash_search_kernel, dim3((n + 255) / 256), dim3(256), 0, s, 
      keys, values, reinterpret_cast<const int4*>(queries), results, n, (uint32_t)(capacity - 1));
  count_launch();
  return cuda_ok();
}

int kernel_map_num_blocks(int M) { return (M + kMapBlock - 1) / kMapBlock; }

int kernel_map_search(const uint64_t* keys, const int* values, int capacity, const int* out_coords,
                      int M, const int* offsets3, int K, int sx, int sy, int sz, int* pair_table,
                      int* block_counts, unsigned long long* mask_keys, cudaStream_t s) {
  if (!is_pow2(capacity) || M < 0 || K < 1 || K > 4096) return kErrInvalidArg;
  if (M == 0) return kOk;
  const int nb = kernel_map_num_blocks(M);
  wcn_launch(kernel_map_search_kernel, dim3(nb), dim3(kMapBlock), (size_t)4 * K * sizeof(int), s, 
      keys, values, (uint32_t)(capacity - 1), reinterpret_cast<const int4*>(out_coords), M,
      offsets3, K, sx, sy, sz, pair_table, block_counts, mask_keys);
  count_launch();
  return cuda_ok();
}

int kernel_map_search_sym(const uint64_t* keys, const int* values, int capacity, const int* coords,
                          int M, const int* offsets3, int K, const int* status, int* pair_table,
                          cudaStream_t s) {
  if (!is_pow2(capacity) || M < 0 || K < 1 || K > 4096 || (K & 1) == 0) return kErrInvalidArg;
  if (M == 0) return kOk;
  // rows 0 .. K/2 are written in full by the kernel; rows K/2+1 .. K-1 receive mirrored hits only
  if (K > 1 && cudaMemsetAsync(pair_table + (size_t)(K / 2 + 1) * M, 0xFF,
                               (size_t)(K - K / 2 - 1) * M * 4, s) != cudaSuccess)
    return kErrCuda;
  wcn_launch(kernel_map_search_sym_kernel, dim3(kernel_map_num_blocks(M)), dim3(kMapBlock), (size_t)3 * K * sizeof(int), s, keys, values, (uint32_t)(capacity - 1),
                                      reinterpret_cast<const int4*>(coords), M, offsets3, K, status,
                                      pair_table);
  count_launch();
  return cuda_ok();
}

int kernel_map_stats(const int* pair_table, int K, int M, int* block_counts,
                     unsigned long long* mask_keys, cudaStream_t s) {
  if (K < 1 || M < 0) return kErrInvalidArg;
  if (M == 0) return kOk;
  wcn_launch(kernel_map_stats_kernel, dim3(kernel_map_num_blocks(M)), dim3(kMapBlock), 0, s, pair_table, K, M,
                                                                        block_counts, mask_keys);
  count_launch();
  return cuda_ok();
}

int kernel_map_count(int* block_counts, int K, int nb, int* counts, int* offsets, cudaStream_t s) {
  if (K < 1) return kErrInvalidArg;
  if (nb > 0) {
    wcn_launch(block_scan_kernel, dim3(K), dim3(256), 0, s, block_counts, nb, counts);
  count_launch();
  } else {
    if (cudaMemsetAsync(counts, 0, (size_t)K * 4, s) != cudaSuccess) return kErrCuda;
  }
  wcn_launch(offsets_kernel, dim3(1), dim3(32), 0, s, counts, K, offsets);
  count_launch();
  return cuda_ok();
}

int kernel_map_scatter(const int* pair_table, const int* block_prefix, const int* offsets,
                       int* in_maps, int* out_maps, int K, int M, cudaStream_t s) {
  if (M == 0) return kOk;
  wcn_launch(kernel_map_scatter_kernel, dim3(kernel_map_num_blocks(M)), dim3(kMapBlock), 0, s, 
      pair_table, block_prefix, offsets, in_maps, out_maps, K, M);
  count_launch();
  return cuda_ok();
}

int reverse_pair_table(const int* pair_table, int K, int M, int* rev, int n_in, cudaStream_t s) {
  if (cudaMemsetAsync(rev, 0xFF, (size_t)K * n_in * 4, s) != cudaSuccess) return kErrCuda;
  const long long total = (long long)K * M;
  if (total == 0) return kOk;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  wcn_launch(reverse_table_kernel, dim3(blocks), dim3(256), 0, s, pair_table, K, M, rev, n_in);
  count_launch();
  return cuda_ok();
}

int csr_to_table(const int* val_maps, const int* row_maps, const int* offsets, int K, int n_rows,
                 int L_upper, int* table, cudaStream_t s) {
  if (cudaMemsetAsync(table, 0xFF, (size_t)K * n_rows * 4, s) != cudaSuccess) return kErrCuda;
  if (L_upper <= 0) return kOk;
  int blocks = (L_upper + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  wcn_launch(csr_to_table_kernel, dim3(blocks), dim3(256), (size_t)(K + 1) * sizeof(int), s, val_maps, row_maps,
                                                                        offsets, K, n_rows, table);
  count_launch();
  return cuda_ok();
}

int mask_keys_from_table(const int* table, int K, int M, unsigned long long* keys, cudaStream_t s) {
  if (M == 0) return kOk;
  wcn_launch(mask_keys_kernel, dim3((M + 255) / 256), dim3(256), 0, s, table, K, M, keys);
  count_launch();
  return cuda_ok();
}

// workspace: [keys_out M*8][rows_in M*4][cub temp]
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr int kSortThreads = 512;     // single-kernel mask sort (mask_sort_kernel below)
constexpr int kSortMaxCtas = 148;
constexpr int kSortMaxRows = 1 << 20;
constexpr int kSortKeysPerCta = 2048;  // measured on C3: 512 / 1024 keys 47 us, 2048 43 us, 4096 61 us

size_t sort_workspace_bytes(int M) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr,
                                  M > 0 ? M : 1, 0, 64, (cudaStream_t)0);
  const size_t cub_bytes =
      align_up((size_t)M * 8, 256) + align_up((size_t)M * 4, 256) + align_up(temp, 256) + 256;
  // single-kernel mask sort: 2 key + 2 value buffers, the [cta][256] count matrix, barrier words
  const size_t own_bytes = 4 * align_up((size_t)M * 4, 256) +
                           align_up((size_t)kSortMaxCtas * 256 * 4, 256) + 256;
  return cub_bytes > own_bytes ? cub_bytes : own_bytes;
}

// K <= 32: the offset masks fit 32 bits; sorting (u32 key, row) pairs moves 8 instead of 12 bytes
// per element and pass. Fused with the row iota.
// Masks of 25..32 bits are compressed to 24-bit sort keys, which saves the fourth 8-bit radix
// pass (12.7 us of C3's 45 us sort). Any row order gives a valid plan; what the order has to keep
// is (i) rows with equal masks adjacent — equal masks still get equal keys — and (ii) similar
// masks close. The compression: for an odd K the centre bit (the identity offset, set in every row
// of a submanifold map: dropped exactly) is removed, then the r lowest bits — the least
// significant for the order — are XORed into the bottom of the remaining high 24. On the surface
// scenes the tile plan needs 0.2 % more steps than with the full-width sort (7 248 vs 7 235 on
// C3-S; folding the three HIGH bits instead costs 1.7 %, plain truncation 5-8 %).
__global__ void narrow_keys_iota_kernel(const unsigned long long* __restrict__ keys,
                                        unsigned* __restrict__ keys32, int* __restrict__ rows,
                                        int M, int drop_bit, int fold_bits) {
  pdl_begin();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) {
    unsigned k = (unsigned)keys[i];
    if (drop_bit >= 0) k = ((k >> (drop_bit + 1)) << drop_bit) | (k & ((1u << drop_bit) - 1u));
    if (fold_bits > 0) k = (k >> fold_bits) ^ (k & ((1u << fold_bits) - 1u));
    keys32[i] = k;
    rows[i] = i;
  }
}

// --------------------------------------------------------------------------------------------
// Mask sort in ONE kernel (maps of up to 2^20 rows, K <= 32). The CUB onesweep sort of the 24-bit
// mask keys is latency-bound at this size (3 passes of 12.7 us with ~50 blocks each, plus a
// histogram, a scan and the key-narrowing kernel: 45 us of the C3 step). Here one persistent grid
// (<= 148 CTAs, all resident) runs every 8-bit LSD pass itself: per pass a shared-memory digit
// histogram of the CTA's contiguous chunk, a grid barrier, the CTA's global bases from the
// [cta][digit] count matrix (coalesced column reads + a block scan over the digits), and a STABLE
// scatter (warps take turns inside a 512-key round, lanes ranked with match.any). Pass 0 narrows
// the 64-bit masks on the fly (centre bit dropped, low bits folded) and creates the row iota, so
// the separate narrowing kernel is gone too. The barrier is a monotonic counter in the workspace,
// zeroed by the launcher.
// --------------------------------------------------------------------------------------------

struct MaskSortParams {
  const unsigned long long* keys64;  // per-row offset masks, or nullptr:
  const int* table;                  // ... derive them from the [K][M] neighbour table in pass 0
  int K;
  unsigned* k_a;   // ping
  unsigned* k_b;   // pong
  int* v_a;
  int* v_b;
  int* rows_out;   // values of the last pass
  int* counts;     // [gridDim.x][256]
  unsigned* bar;   // arrivals (monotonic counter, zero at launch)
  int M, per_cta, passes, drop_bit, fold_bits;
};

__device__ __forceinline__ unsigned narrow_mask_key(unsigned long long k64, int drop_bit,
                                                    int fold_bits) {
  unsigned k = (unsigned)k64;
  if (drop_bit >= 0) k = ((k >> (drop_bit + 1)) << drop_bit) | (k & ((1u << drop_bit) - 1u));
  if (fold_bits > 0) k = (k >> fold_bits) ^ (k & ((1u << fold_bits) - 1u));
  return k;
}

__device__ __forceinline__ void sort_grid_barrier(unsigned* bar, unsigned n_ctas, unsigned& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned target = (gen + 1u) * n_ctas;
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
    } while (seen < target);
    __threadfence();
  }
  ++gen;
  __syncthreads();
}

__global__ void __launch_bounds__(kSortThreads, 1) mask_sort_kernel(const MaskSortParams p) {
  pdl_begin();
  __shared__ int s_hist[256];      // histogram of the chunk, then running digit counters
  __shared__ int s_base[256];      // global position of the CTA's first key of every digit
  __shared__ int s_scan[256];
  __shared__ int s_part[4][256];   // per-half column sums (totals, then "before this CTA")
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, c = blockIdx.x;
  const int begin = min(p.M, c * p.per_cta), end = min(p.M, begin + p.per_cta);
  unsigned gen = 0;
  for (int pass = 0; pass < p.passes; ++pass) {
    const unsigned* k_src = (pass & 1) ? p.k_b : p.k_a;
    const int* v_src = (pass & 1) ? p.v_b : p.v_a;
    unsigned* k_dst = (pass & 1) ? p.k_a : p.k_b;
    int* v_dst = (pass == p.passes - 1) ? p.rows_out : ((pass & 1) ? p.v_a : p.v_b);
    const int shift = 8 * pass;
    // pass 0 of the table form: the row's offset mask from its K table entries (coalesced column
    // reads of the L2-resident table the search kernel has just written), narrowed, kept in k_a
    // (free during pass 0) so that the scatter phase does not read the table again
    const bool from_table = pass == 0 && p.table != nullptr;
    auto load_key = [&](int i) -> unsigned {
      // __ldcg: written by other CTAs in the previous pass (never trust a stale L1 line)
      if (pass != 0) return __ldcg(k_src + i);
      if (p.table != nullptr) return p.k_a[i];  // stored by this thread in phase 1
      return narrow_mask_key(__ldg(p.keys64 + i), p.drop_bit, p.fold_bits);
    };
    // ---- phase 1: digit histogram of this CTA's chunk ------------------------------------------
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    for (int i = begin + tid; i < end; i += kSortThreads) {
      unsigned key;
      if (from_table) {
        unsigned long long bits = 0ull;
#pragma unroll 9
        for (int k = 0; k < p.K; ++k)
          bits |= (unsigned long long)(__ldcg(p.table + (size_t)k * p.M + i) >= 0 ? 1u : 0u) << k;
        key = narrow_mask_key(bits, p.drop_bit, p.fold_bits);
        p.k_a[i] = key;
      } else {
        key = load_key(i);
      }
      atomicAdd(&s_hist[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (tid < 256) p.counts[c * 256 + tid] = s_hist[tid];
    sort_grid_barrier(p.bar, (unsigned)G, gen);
    // ---- phase 2: base[d] = (keys of smaller digits, all CTAs) + (digit d in earlier CTAs) ----
    // every thread sums one digit column over half of the CTAs: independent L2 loads, 16 in flight
    {
      const int d = tid & 255, h = tid >> 8;
      const int c0 = h ? (G + 1) / 2 : 0, c1 = h ? G : (G + 1) / 2;
      int before = 0, total = 0;
#pragma unroll 16
      for (int cc = c0; cc < c1; ++cc) {
        const int v = __ldcg(p.counts + cc * 256 + d);
        total += v;
        before += cc < c ? v : 0;
      }
      s_part[h][d] = total;
      s_part[2 + h][d] = before;
    }
    __syncthreads();
    if (tid < 256) {
      s_scan[tid] = s_part[0][tid] + s_part[1][tid];
      s_base[tid] = s_part[2][tid] + s_part[3][tid];
      s_hist[tid] = 0;  // now: keys of digit d already placed by this CTA
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 256 digit totals: 8 per lane
      int loc[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = s_scan[lane * 8 + j]; sum += loc[j]; }
      int incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      int run = incl - sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) { s_scan[lane * 8 + j] = run; run += loc[j]; }
    }
    __syncthreads();
    if (tid < 256) s_base[tid] += s_scan[tid];
    __syncthreads();
    // ---- phase 3: stable scatter, 512 keys per round, warps in order -------------------------
    for (int r0 = begin; r0 < end; r0 += kSortThreads) {
      const int i = r0 + tid;
      const bool live = i < end;
      unsigned key = 0u;
      int val = 0;
      if (live) {
        key = load_key(i);
        val = pass == 0 ? i : __ldcg(v_src + i);
      }
      const unsigned digit = (key >> shift) & 255u;
      // lanes of the warp with the same digit (dead lanes get a private value)
      const unsigned peers = __match_any_sync(0xffffffffu, live ? digit : (256u + lane));
      const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
      int pos = 0;
      for (int w = 0; w < kSortThreads / 32; ++w) {
        if (warp == w && live) {
          pos = s_base[digit] + s_hist[digit] + rank_in_warp;
        }
        __syncwarp();
        if (warp == w && live && rank_in_warp == 0) s_hist[digit] += __popc(peers);
        __syncthreads();
      }
      if (live) {
        if (pass != p.passes - 1) k_dst[pos] = key;
        v_dst[pos] = val;
      }
    }
    if (pass != p.passes - 1) sort_grid_barrier(p.bar, (unsigned)G, gen);
  }
}

// bring-up builds (WCN_BRINGUP=1 build.sh): WCN_FOLD_MASK_KEYS=0 in the environment keeps the
// full-width sort for A/B measurements (read once); the default build always compresses
static constexpr bool g_own_mask_sort = true;  // false: CUB onesweep for every size (A/B builds)
#ifdef WCN_BRINGUP
static const bool g_fold_mask_keys = [] {
  const char* v = getenv("WCN_FOLD_MASK_KEYS");
  return !(v && v[0] == '0');
}();
#else
static constexpr bool g_fold_mask_keys = true;
#endif

static int launch_mask_sort(const unsigned long long* keys, const int* table, int M, int K,
                            int* rows_out, uint8_t* ws, cudaStream_t s) {
  {
    int drop_bit = -1, fold_bits = 0, key_bits = K;
    if (K > 24 && g_fold_mask_keys) {
      if (K & 1) { drop_bit = K / 2; --key_bits; }
      fold_bits = key_bits - 24;
      key_bits = 24;
    }
    const size_t mb = align_up((size_t)M * 4, 256);
    MaskSortParams sp;
    sp.keys64 = keys;
    sp.table = table;
    sp.K = K;
    sp.k_a = reinterpret_cast<unsigned*>(ws);
    sp.k_b = reinterpret_cast<unsigned*>(ws + mb);
    sp.v_a = reinterpret_cast<int*>(ws + 2 * mb);
    sp.v_b = reinterpret_cast<int*>(ws + 3 * mb);
    sp.counts = reinterpret_cast<int*>(ws + 4 * mb);
    sp.bar = reinterpret_cast<unsigned*>(ws + 4 * mb + align_up((size_t)kSortMaxCtas * 256 * 4, 256));
    sp.rows_out = rows_out;
    sp.M = M;
    sp.passes = (key_bits + 7) / 8;
    sp.drop_bit = drop_bit;
    sp.fold_bits = fold_bits;
    int keys_per_cta = kSortKeysPerCta;
#ifdef WCN_BRINGUP
    if (const char* e = getenv("WCN_SORT_KEYS_PER_CTA")) keys_per_cta = atoi(e) > 0 ? atoi(e) : keys_per_cta;
#endif
    int ctas = (M + keys_per_cta - 1) / keys_per_cta;
    if (ctas > kSortMaxCtas) ctas = kSortMaxCtas;
    if (ctas < 1) ctas = 1;
    int sms = 0, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0 &&
        ctas > sms)
      ctas = sms;  // every CTA must be resident: the passes meet at a grid barrier
    sp.per_cta = ((M + ctas - 1) / ctas + kSortThreads - 1) / kSortThreads * kSortThreads;
    if (cudaMemsetAsync(sp.bar, 0, 8, s) != cudaSuccess) return kErrCuda;  // barrier counter
    wcn_launch(mask_sort_kernel, dim3(ctas), dim3(kSortThreads), 0, s, sp);
    count_launch();
    return cuda_ok();
  }
}

// Rows sorted by the offset mask of their table column, masks derived inside the sort kernel
// (no separate mask pass over the table). kErrUnsupportedShape: use wcn_mask_keys + sort_rows_by_key.
int sort_rows_by_table(const int* table, int K, int M, int* rows_out, void* workspace,
                       size_t ws_bytes, cudaStream_t s) {
  if (M == 0) return kOk;
  if (K < 1 || K > 32 || M > kSortMaxRows || !g_own_mask_sort) return kErrUnsupportedShape;
  if (ws_bytes < sort_workspace_bytes(M)) return kErrWorkspace;
  return launch_mask_sort(nullptr, table, M, K, rows_out, reinterpret_cast<uint8_t*>(workspace), s);
}

int sort_rows_by_key(const unsigned long long* keys, int M, int K, int* rows_out, void* workspace,
                     size_t ws_bytes, cudaStream_t s) {
  if (M == 0) return kOk;
  if (ws_bytes < sort_workspace_bytes(M)) return kErrWorkspace;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  unsigned long long* keys_out = reinterpret_cast<unsigned long long*>(ws);
  int* rows_in = reinterpret_cast<int*>(ws + align_up((size_t)M * 8, 256));
  void* temp = ws + align_up((size_t)M * 8, 256) + align_up((size_t)M * 4, 256);
  size_t temp_bytes = ws_bytes - (align_up((size_t)M * 8, 256) + align_up((size_t)M * 4, 256));
  if (K <= 32 && M <= kSortMaxRows && g_own_mask_sort)
    return launch_mask_sort(keys, nullptr, M, K, rows_out, ws, s);
  if (K <= 32) {
    unsigned* k32_in = reinterpret_cast<unsigned*>(ws);            // the u64 key_out region holds
    unsigned* k32_out = k32_in + align_up((size_t)M, 32);          // both u32 key buffers
    if (align_up((size_t)M, 32) * 8 <= align_up((size_t)M * 8, 256)) {
      int drop_bit = -1, fold_bits = 0, key_bits = K;
      if (K > 24 && g_fold_mask_keys) {
        if (K & 1) { drop_bit = K / 2; --key_bits; }
        fold_bits = key_bits - 24;
        key_bits = 24;
      }
      wcn_launch(narrow_keys_iota_kernel, dim3((M + 255) / 256), dim3(256), 0, s, keys, k32_in, rows_in, M, drop_bit,
                                                              fold_bits);
      count_launch();
      cudaError_t e32 = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k32_in, k32_out, rows_in,
                                                        rows_out, M, 0, key_bits, s);
      return e32 == cudaSuccess ? cuda_ok() : kErrCuda;
    }
  }
  wcn_launch(iota_kernel, dim3((M + 255) / 256), dim3(256), 0, s, rows_in, M);
  count_launch();
  const int end_bit = K < 64 ? K : 64;
  // LSD radix sort is stable: equal masks keep ascending row order (deterministic tiles)
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_out, rows_in,
                                                  rows_out, M, 0, end_bit, s);
  return e == cudaSuccess ? cuda_ok() : kErrCuda;
}

int build_tiles(const int* table, int K, int M, const int* sorted_rows, int tile_rows, int m_pad,
                int* step_nbr, int* step_k, int* rows_padded, int* tile_nk, int* tile_cum,
                int n_range_ctas, int* cta_units, const unsigned long long* masks, cudaStream_t s) {
  if (K > 64) masks = nullptr;  // the 64-bit row masks are exact up to 64 offsets only
  if (tile_rows != 128 && tile_rows != 256) return kErrInvalidArg;
  if (m_pad % tile_rows != 0 || m_pad < M || K < 1) return kErrInvalidArg;
  const int num_tiles = m_pad / tile_rows;
  if (num_tiles > 0) {
    wcn_launch(build_tiles_kernel, dim3(num_tiles), dim3(tile_rows), 0, s, table, K, M, sorted_rows, tile_rows,
                                                      step_nbr, step_k, rows_padded, tile_nk, masks);
    count_launch();
  }
  wcn_launch(tile_scan_kernel, dim3(1), dim3(1024), 0, s, tile_nk, num_tiles, tile_cum, tile_rows / 128, n_range_ctas,
                                      cta_units);
  count_launch();
  return cuda_ok();
}

}  // namespace wcn
