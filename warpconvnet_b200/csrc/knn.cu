// SPDX-License-Identifier: Apache-2.0
// Exact batched k-nearest-neighbour search on a uniform grid (the neighbour search behind
// `Points.neighbors` / `PointConv`).
//
// Replaces (behaviour, not code): warpconvnet/geometry/coords/search/knn.py:10-142 — a per-batch
// Python loop over chunked `torch.cdist` + `torch.topk`, O(M*N) distance tiles. Here the
// reference points are counting-sorted into a dense grid (one slab of dims^3 cells per batch
// item, cell size chosen by the host from N alone, so no host sync) and every query thread
// visits Chebyshev shells of cells around its own cell until its k-th best distance is provably
// final. All of it is integer / fp32 HBM-bound work: coalesced float4 point records, no GEMM.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace wcn {

struct KnnGrid {
  float origin[3];
  float inv_cs[3];
  float cs_min;
  int dims;  // cells per axis
};

__device__ __forceinline__ int float_flip(float f) {  // order-preserving float -> int
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float float_unflip(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF);
}

// bbox[0..2] = min, bbox[3..5] = max over ALL points (order-preserving ints, atomics)
__global__ void knn_bbox_kernel(const float* __restrict__ pts, int n, int* __restrict__ bbox) {
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int v = float_flip(__ldg(pts + 3 * (size_t)i + a));
      lo[a] = min(lo[a], v);
      hi[a] = max(hi[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int d = 16; d > 0; d >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], d));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], d));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(bbox + a, lo[a]);
      atomicMax(bbox + 3 + a, hi[a]);
    }
  }
}

__global__ void knn_params_kernel(const int* __restrict__ bbox, int dims, KnnGrid* __restrict__ g) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float cs_min = 3.0e38f;
  for (int a = 0; a < 3; ++a) {
    const float lo = float_unflip(bbox[a]), hi = float_unflip(bbox[3 + a]);
    float ext = hi - lo;
    if (!(ext > 0.f)) ext = 1.f;
    const float cs = ext / dims * 1.0001f;  // the max point still lands in the last cell
    g->origin[a] = lo;
    g->inv_cs[a] = 1.f / cs;
    cs_min = fminf(cs_min, cs);
  }
  g->cs_min = cs_min;
  g->dims = dims;
}

__device__ __forceinline__ int batch_of(const int* __restrict__ offsets, int nb, int i) {
  int lo = 0, hi = nb;  // largest b with offsets[b] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(offsets + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void cell_of(const KnnGrid& g, float x, float y, float z, int& cx,
                                        int& cy, int& cz) {
  cx = min(max((int)((x - g.origin[0]) * g.inv_cs[0]), 0), g.dims - 1);
  cy = min(max((int)((y - g.origin[1]) * g.inv_cs[1]), 0), g.dims - 1);
  cz = min(max((int)((z - g.origin[2]) * g.inv_cs[2]), 0), g.dims - 1);
}

__global__ void knn_count_kernel(const float* __restrict__ pts, int n,
                                 const int* __restrict__ offsets, int nb,
                                 const KnnGrid* __restrict__ gp, int* __restrict__ cell_id,
                                 int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const KnnGrid g = *gp;
  int cx, cy, cz;
  cell_of(g, __ldg(pts + 3 * (size_t)i), __ldg(pts + 3 * (size_t)i + 1),
          __ldg(pts + 3 * (size_t)i + 2), cx, cy, cz);
  const int b = batch_of(offsets, nb, i);
  const int cell = ((b * g.dims + cx) * g.dims + cy) * g.dims + cz;
  cell_id[i] = cell;
  atomicAdd(counts + cell, 1);
}

// point records (x, y, z, bits(global index)) grouped by cell
__global__ void knn_fill_kernel(const float* __restrict__ pts, int n,
                                const int* __restrict__ cell_id,
                                const int* __restrict__ cell_start, int* __restrict__ cursor,
                                float4* __restrict__ recs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cell = cell_id[i];
  const int pos = __ldg(cell_start + cell) + atomicAdd(cursor + cell, 1);
  recs[pos] = make_float4(__ldg(pts + 3 * (size_t)i), __ldg(pts + 3 * (size_t)i + 1),
                          __ldg(pts + 3 * (size_t)i + 2), __int_as_float(i));
}

// One thread per query; top-K kept sorted (ascending) in registers. Ties: the smaller index wins,
// so the result is deterministic although the order of points inside a cell is not.
template <int K>
__global__ void __launch_bounds__(128)
knn_query_kernel(const float* __restrict__ q, int m, const int* __restrict__ q_offsets, int nb,
                 const KnnGrid* __restrict__ gp, const int* __restrict__ cell_start,
                 const float4* __restrict__ recs, int k, long long* __restrict__ out_idx,
                 float* __restrict__ out_dist) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= m) return;
  const KnnGrid g = *gp;
  const float x = __ldg(q + 3 * (size_t)qi), y = __ldg(q + 3 * (size_t)qi + 1),
              z = __ldg(q + 3 * (size_t)qi + 2);
  const int b = batch_of(q_offsets, nb, qi);
  int cx, cy, cz;
  cell_of(g, x, y, z, cx, cy, cz);
  float bd[K];
  int bi[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { bd[j] = 3.0e38f; bi[j] = -1; }
  const int D = g.dims;
  for (int r = 0; r < D; ++r) {
    for (int dx = -r; dx <= r; ++dx) {
      const int ux = cx + dx;
      if (ux < 0 || ux >= D) continue;
      for (int dy = -r; dy <= r; ++dy) {
        const int uy = cy + dy;
        if (uy < 0 || uy >= D) continue;
        const bool face = (dx == -r || dx == r || dy == -r || dy == r);
        const int step = face ? 1 : 2 * r;  // interior columns only touch the two z caps
        for (int dz = -r; dz <= r; dz += (step > 0 ? step : 1)) {
          const int uz = cz + dz;
          if (uz < 0 || uz >= D) continue;
          const int cell = ((b * D + ux) * D + uy) * D + uz;
          const int s = __ldg(cell_start + cell), e = __ldg(cell_start + cell + 1);
          for (int pidx = s; pidx < e; ++pidx) {
            const float4 rec = __ldg(recs + pidx);
            const float ddx = rec.x - x, ddy = rec.y - y, ddz = rec.z - z;
            const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
            const int id = __float_as_int(rec.w);
            if (d2 < bd[K - 1] || (d2 == bd[K - 1] && id < bi[K - 1])) {
              // insertion into the sorted list (fully unrolled: stays in registers)
              float cd = d2;
              int ci = id;
#pragma unroll
              for (int j = 0; j < K; ++j) {
                const bool before = cd < bd[j] || (cd == bd[j] && ci < bi[j]);
                const float td = bd[j];
                const int ti = bi[j];
                if (before) { bd[j] = cd; bi[j] = ci; cd = td; ci = ti; }
              }
            }
          }
          if (r == 0) break;
        }
      }
    }
    // everything outside shell r is at least r * cs_min away
    const float reach = r * g.cs_min;
    float kth_d = bd[K - 1];
    int kth_i = bi[K - 1];
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (j == k - 1) { kth_d = bd[j]; kth_i = bi[j]; }
    if (kth_i >= 0 && kth_d <= reach * reach) break;
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (j < k) {
      out_idx[(size_t)qi * k + j] = bi[j];
      if (out_dist != nullptr) out_dist[(size_t)qi * k + j] = sqrtf(bd[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Radius search on the same grid (replaces warpconvnet/geometry/coords/search/radius.py:16-291 and
// csrc/radius_search_kernels.cu:17-133): pass 1 counts the reference points within `radius` of
// every query (same batch item), the caller scans the counts into row splits, pass 2 writes the
// CSR lists. FILL = false: count only.
// ---------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(128)
radius_query_kernel(const float* __restrict__ q, int m, const int* __restrict__ q_offsets, int nb,
                    const KnnGrid* __restrict__ gp, const int* __restrict__ cell_start,
                    const float4* __restrict__ recs, float radius, int* __restrict__ counts,
                    const long long* __restrict__ row_splits, int* __restrict__ out_idx,
                    float* __restrict__ out_dist) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= m) return;
  const KnnGrid g = *gp;
  const float x = __ldg(q + 3 * (size_t)qi), y = __ldg(q + 3 * (size_t)qi + 1),
              z = __ldg(q + 3 * (size_t)qi + 2);
  const int b = batch_of(q_offsets, nb, qi);
  const int D = g.dims;
  // cell range that can hold a point within `radius` (clamped to the grid; queries outside the
  // reference bounding box are handled by the clamp, the distance test decides)
  int lo[3], hi[3];
  const float p3[3] = {x, y, z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = min(max((int)floorf((p3[a] - radius - g.origin[a]) * g.inv_cs[a]), 0), D - 1);
    hi[a] = min(max((int)floorf((p3[a] + radius - g.origin[a]) * g.inv_cs[a]), 0), D - 1);
  }
  const float r2 = radius * radius;
  int n = 0;
  long long w = FILL ? row_splits[qi] : 0;
  for (int ux = lo[0]; ux <= hi[0]; ++ux)
    for (int uy = lo[1]; uy <= hi[1]; ++uy) {
      // cells (ux, uy, lo..hi) are consecutive in memory: one contiguous record range
      const int c0 = ((b * D + ux) * D + uy) * D + lo[2];
      const int s = __ldg(cell_start + c0), e = __ldg(cell_start + c0 + (hi[2] - lo[2]) + 1);
      for (int pidx = s; pidx < e; ++pidx) {
        const float4 rec = __ldg(recs + pidx);
        const float ddx = rec.x - x, ddy = rec.y - y, ddz = rec.z - z;
        const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
        if (d2 <= r2) {
          if (FILL) {
            out_idx[w] = __float_as_int(rec.w);
            if (out_dist != nullptr) out_dist[w] = sqrtf(d2);
            ++w;
          }
          ++n;
        }
      }
    }
  if (!FILL) counts[qi] = n;
}

static inline int cuda_ok2() { return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda; }

size_t knn_workspace_bytes(int n_ref, int n_batches, int dims) {
  const size_t cells = (size_t)n_batches * dims * dims * dims;
  size_t scan_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const int*)nullptr, (int*)nullptr,
                                (int)(cells + 1));
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  return al(256) /*bbox + grid*/ + al((cells + 1) * 4) * 3 /*counts, start, cursor*/ +
         al((size_t)n_ref * 4) /*cell ids*/ + al((size_t)n_ref * 16) /*records*/ + al(scan_tmp);
}

int knn_dims_for(int n_ref, int n_batches) {
  // ~4 reference points per cell on average, at most 96^3 cells per batch item
  const double per_batch = n_batches > 0 ? (double)n_ref / n_batches : 0.0;
  int d = (int)llround(cbrt(per_batch / 4.0));
  if (d < 1) d = 1;
  if (d > 96) d = 96;
  return d;
}

struct KnnWorkspace {
  int* bbox;
  KnnGrid* grid;
  int* counts;
  int* start;
  int* cursor;
  int* cell_id;
  float4* recs;
  void* scan_tmp;
  size_t scan_bytes;
};

static KnnWorkspace knn_carve(void* workspace, size_t ws_bytes, int n_ref, int n_batches, int dims) {
  const size_t cells = (size_t)n_batches * dims * dims * dims;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  KnnWorkspace k;
  k.bbox = reinterpret_cast<int*>(w);
  k.grid = reinterpret_cast<KnnGrid*>(w + 64);
  w += al(256);
  k.counts = reinterpret_cast<int*>(w); w += al((cells + 1) * 4);
  k.start = reinterpret_cast<int*>(w); w += al((cells + 1) * 4);
  k.cursor = reinterpret_cast<int*>(w); w += al((cells + 1) * 4);
  k.cell_id = reinterpret_cast<int*>(w); w += al((size_t)n_ref * 4);
  k.recs = reinterpret_cast<float4*>(w); w += al((size_t)n_ref * 16);
  k.scan_tmp = w;
  k.scan_bytes = ws_bytes - (size_t)(w - reinterpret_cast<uint8_t*>(workspace));
  return k;
}

// counting-sorts the reference points into the grid (bbox -> cell ids -> scan -> records)
static int knn_build_grid(const float* ref, int n_ref, const int* ref_offsets, int n_batches,
                          int dims, const KnnWorkspace& k, cudaStream_t s) {
  const size_t cells = (size_t)n_batches * dims * dims * dims;
  const int h_init[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  if (cudaMemcpyAsync(k.bbox, h_init, sizeof(h_init), cudaMemcpyHostToDevice, s) != cudaSuccess)
    return kErrCuda;
  if (cudaMemsetAsync(k.counts, 0, (cells + 1) * 4, s) != cudaSuccess) return kErrCuda;
  if (cudaMemsetAsync(k.cursor, 0, (cells + 1) * 4, s) != cudaSuccess) return kErrCuda;
  if (n_ref > 0) {
    int blocks = (n_ref + 255) / 256;
    knn_bbox_kernel<<<blocks < 592 ? blocks : 592, 256, 0, s>>>(ref, n_ref, k.bbox);
    count_launch();
  }
  knn_params_kernel<<<1, 32, 0, s>>>(k.bbox, dims, k.grid);
  count_launch();
  if (n_ref > 0) {
    knn_count_kernel<<<(n_ref + 255) / 256, 256, 0, s>>>(ref, n_ref, ref_offsets, n_batches, k.grid,
                                                        k.cell_id, k.counts);
    count_launch();
  }
  size_t scan_bytes = k.scan_bytes;
  if (cub::DeviceScan::ExclusiveSum(k.scan_tmp, scan_bytes, k.counts, k.start, (int)(cells + 1),
                                    s) != cudaSuccess)
    return kErrCuda;
  if (n_ref > 0) {
    knn_fill_kernel<<<(n_ref + 255) / 256, 256, 0, s>>>(ref, n_ref, k.cell_id, k.start, k.cursor,
                                                       k.recs);
    count_launch();
  }
  return kOk;
}

// pass 1 of the radius search: builds the grid in `workspace` and writes counts[n_query]
int radius_count(const float* ref, int n_ref, const int* ref_offsets, const float* query,
                 int n_query, const int* query_offsets, int n_batches, float radius, int* counts,
                 void* workspace, size_t ws_bytes, cudaStream_t s) {
  if (!(radius > 0.f) || n_batches < 1 || n_ref < 0 || n_query < 0) return kErrInvalidArg;
  const int dims = knn_dims_for(n_ref, n_batches);
  if (ws_bytes < knn_workspace_bytes(n_ref, n_batches, dims)) return kErrWorkspace;
  const KnnWorkspace k = knn_carve(workspace, ws_bytes, n_ref, n_batches, dims);
  const int st = knn_build_grid(ref, n_ref, ref_offsets, n_batches, dims, k, s);
  if (st != kOk) return st;
  if (n_query == 0) return kOk;
  radius_query_kernel<false><<<(n_query + 127) / 128, 128, 0, s>>>(
      query, n_query, query_offsets, n_batches, k.grid, k.start, k.recs, radius, counts, nullptr,
      nullptr, nullptr);
  count_launch();
  return cuda_ok2();
}

// pass 2: `workspace` must be the untouched workspace of radius_count for the same reference set
int radius_fill(int n_ref, const float* query, int n_query, const int* query_offsets,
                int n_batches, float radius, const long long* row_splits, int* out_idx,
                float* out_dist, void* workspace, size_t ws_bytes, cudaStream_t s) {
  if (!(radius > 0.f) || n_batches < 1 || n_query < 0) return kErrInvalidArg;
  if (n_query == 0) return kOk;
  const int dims = knn_dims_for(n_ref, n_batches);
  if (ws_bytes < knn_workspace_bytes(n_ref, n_batches, dims)) return kErrWorkspace;
  const KnnWorkspace k = knn_carve(workspace, ws_bytes, n_ref, n_batches, dims);
  radius_query_kernel<true><<<(n_query + 127) / 128, 128, 0, s>>>(
      query, n_query, query_offsets, n_batches, k.grid, k.start, k.recs, radius, nullptr,
      row_splits, out_idx, out_dist);
  count_launch();
  return cuda_ok2();
}

int knn_search(const float* ref, int n_ref, const int* ref_offsets, const float* query, int n_query,
               const int* query_offsets, int n_batches, int k, long long* out_idx, float* out_dist,
               void* workspace, size_t ws_bytes, cudaStream_t s) {
  if (k < 1 || k > 64 || n_batches < 1 || n_ref < 0 || n_query < 0) return kErrInvalidArg;
  if (n_query == 0) return kOk;
  const int dims = knn_dims_for(n_ref, n_batches);
  if (ws_bytes < knn_workspace_bytes(n_ref, n_batches, dims)) return kErrWorkspace;
  const KnnWorkspace kw = knn_carve(workspace, ws_bytes, n_ref, n_batches, dims);
  {
    const int st = knn_build_grid(ref, n_ref, ref_offsets, n_batches, dims, kw, s);
    if (st != kOk) return st;
  }
  KnnGrid* grid = kw.grid;
  int* start = kw.start;
  float4* recs = kw.recs;
  const int qb = (n_query + 127) / 128;
  if (k <= 8)
    knn_query_kernel<8><<<qb, 128, 0, s>>>(query, n_query, query_offsets, n_batches, grid, start,
                                           recs, k, out_idx, out_dist);
  else if (k <= 16)
    knn_query_kernel<16><<<qb, 128, 0, s>>>(query, n_query, query_offsets, n_batches, grid, start,
                                            recs, k, out_idx, out_dist);
  else if (k <= 32)
    knn_query_kernel<32><<<qb, 128, 0, s>>>(query, n_query, query_offsets, n_batches, grid, start,
                                            recs, k, out_idx, out_dist);
  else
    knn_query_kernel<64><<<qb, 128, 0, s>>>(query, n_query, query_offsets, n_batches, grid, start,
                                            recs, k, out_idx, out_dist);
  count_launch();
  return cuda_ok2();
}

}  // namespace wcn
