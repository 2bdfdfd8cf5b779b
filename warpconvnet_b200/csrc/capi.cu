// SPDX-License-Identifier: Apache-2.0
// extern "C" boundary of libwcn_b200.so — see include/wcn_b200.h for the contract and the
// reference bindings each entry point replaces.
#include "../../include/wcn_b200.h"

#include <cstdlib>

#include "common.cuh"
#include "conv_gemm.cuh"
#include "rownorm.cuh"

namespace wcn {
// cuhash.cu
int hash_prepare(uint64_t*, int*, int, cudaStream_t);
int hash_insert(uint64_t*, int*, const int*, int, int, int*, cudaStream_t);
int hash_search(const uint64_t*, const int*, const int*, int*, int, int, cudaStream_t);
int kernel_map_num_blocks(int);
int kernel_map_search(const uint64_t*, const int*, int, const int*, int, const int*, int, int, int,
                      int, int*, int*, unsigned long long*, cudaStream_t);
int kernel_map_count(int*, int, int, int*, int*, cudaStream_t);
int kernel_map_search_sym(const uint64_t*, const int*, int, const int*, int, const int*, int,
                          const int*, int*, cudaStream_t);
int kernel_map_stats(const int*, int, int, int*, unsigned long long*, cudaStream_t);
int kernel_map_scatter(const int*, const int*, const int*, int*, int*, int, int, cudaStream_t);
int reverse_pair_table(const int*, int, int, int*, int, cudaStream_t);
int mask_keys_from_table(const int*, int, int, unsigned long long*, cudaStream_t);
int csr_to_table(const int*, const int*, const int*, int, int, int, int*, cudaStream_t);
size_t sort_workspace_bytes(int);
int sort_rows_by_key(const unsigned long long*, int, int, int*, void*, size_t, cudaStream_t);
int sort_rows_by_table(const int*, int, int, int*, void*, size_t, cudaStream_t);
int build_tiles(const int*, int, int, const int*, int, int, int*, int*, int*, int*, int*, int, int*,
                const unsigned long long*, cudaStream_t);
// coords.cu
size_t coords_unique_workspace_bytes(long long);
int coords_unique(const int*, int, int, int, int, const int*, int, int, int*, int*, int*, void*,
                  size_t, cudaStream_t);
// peer_allreduce.cu
int peer_allreduce_f32(void* const*, void* const*, int, int, long long, float, int, cudaStream_t);
int peer_allreduce_flag_words();
int peer_allreduce_timeout_word();
// knn.cu
size_t knn_workspace_bytes(int, int, int);
int knn_dims_for(int, int);
int knn_search(const float*, int, const int*, const float*, int, const int*, int, int, long long*,
               float*, void*, size_t, cudaStream_t);
int radius_count(const float*, int, const int*, const float*, int, const int*, int, float, int*,
                 void*, size_t, cudaStream_t);
int radius_fill(int, const float*, int, const int*, int, float, const long long*, int*, float*,
                void*, size_t, cudaStream_t);
// weight_prep.cu
int launch_weight_image(const WeightPrepParams&, const WeightPrepParams*, cudaStream_t);
// rownorm.cu
int rownorm_launch(int which, const RowNormParams&, int dtype, cudaStream_t);
int bn_finalize(const double*, int, int, const float*, const float*, float, float, float*, float*,
                float*, float*, float*, cudaStream_t);
// conv_depthwise.cu
int depthwise_launch(bool wgrad, const void* x, long long ld_x, const void* dy, long long ld_dy,
                     void* y, long long ld_y, const float* w, float* dw, const float* bias,
                     const int* table, int M, int K, int C, int kflip, int relu, int dtype,
                     cudaStream_t s);
int depthwise_plan_launch(const void* x, long long ld_x, void* y, long long ld_y, const float* w,
                          const float* bias, const int* step_nbr, const int* step_k,
                          const int* rows, const int* tile_nk, int num_tiles, int tile_rows, int K,
                          int C, int kflip, int relu, int dtype, cudaStream_t s);
int depthwise_plan_wgrad_launch(const void* x, long long ld_x, const void* dy, long long ld_dy,
                                float* dw, const int* step_nbr, const int* step_k, const int* rows,
                                const int* tile_nk, int num_tiles, int tile_rows, int K, int C,
                                int dtype, cudaStream_t s);
// conv_fwd.cu / conv_wgrad.cu
int launch_gather_gemm(const GatherGemmParams&, int dtype, int n_slabs, int max_ctas,
                       int n_range_ctas, cudaStream_t);
int launch_wgrad(const WgradParams&, int dtype, int y_slabs, int z_slabs, int max_ctas,
                 long long n_in_rows, long long n_out_rows, cudaStream_t);

static long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }

static int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = kNumSMsB200;
  }
  return cached;
}

// Slab plan shared by the weight image, the gather-GEMM and wgrad.
//   rows = channels produced per group (rg), contraction = channels consumed per group (cg)
struct SlabPlan {
  int n_slabs;  // grid.y
  int gps;      // groups per slab (1 for dense)
  int bn;       // rows per slab
  int cdim;     // contraction channels per slab
};

static int plan_slabs(int groups, int rg, int cg, SlabPlan* plan) {
  if (groups < 1 || rg < 1 || cg < 1) return kErrInvalidArg;
  if (groups == 1) {
    int n = (rg + 255) / 256;
    while (n <= rg && (rg % n != 0 || (rg / n) % 16 != 0)) ++n;
    if (n > rg) return kErrUnsupportedShape;
    plan->n_slabs = n;
    plan->gps = 1;
    plan->bn = rg / n;
    plan->cdim = cg;
    return kOk;
  }
  // group conv: densify `gps` groups block-diagonally per slab
  int best = 0;
  for (int g = 1; g <= groups; ++g) {
    if (groups % g) continue;
    if (g * rg <= 128 && g * cg <= 128 && (g * rg) % 16 == 0) best = g;
  }
  if (best == 0) {
    // large groups: one group per slab if it fits a single tile
    if (rg <= 256 && rg % 16 == 0) best = 1; else return kErrUnsupportedShape;
  }
  plan->gps = best;
  plan->n_slabs = groups / best;
  plan->bn = best * rg;
  plan->cdim = best * cg;
  return kOk;
}

}  // namespace wcn

using namespace wcn;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

const char* wcn_version(void) { return "wcn_b200 0.1.0 (sm_100a, tcgen05)"; }
int wcn_built_for_sm100a(void) { return 1; }
long long wcn_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int wcn_hash_prepare(uint64_t* keys, int32_t* values, int capacity, void* stream) {
  if (!keys || !values) return kErrInvalidArg;
  return hash_prepare(keys, values, capacity, S(stream));
}
int wcn_hash_insert(uint64_t* keys, int32_t* values, const int32_t* coords, int n, int capacity,
                    int32_t* status, void* stream) {
  if (!keys || !values || !status || (n > 0 && !coords)) return kErrInvalidArg;
  if ((long long)n * 2 > (long long)capacity) return kErrInvalidArg;  // load factor <= 0.5
  return hash_insert(keys, values, coords, n, capacity, status, S(stream));
}
int wcn_hash_search(const uint64_t* keys, const int32_t* values, const int32_t* queries,
                    int32_t* results, int n, int capacity, void* stream) {
  if (!keys || !values || (n > 0 && (!queries || !results))) return kErrInvalidArg;
  return hash_search(keys, values, queries, results, n, capacity, S(stream));
}

int wcn_kernel_map_num_blocks(int M) { return kernel_map_num_blocks(M); }
int wcn_kernel_map_search(const uint64_t* keys, const int32_t* values, int capacity,
                          const int32_t* out_coords, int M, const int32_t* offsets3, int K,
                          int stride_x, int stride_y, int stride_z, int32_t* pair_table,
                          int32_t* block_counts, uint64_t* mask_keys, void* stream) {
  if (!keys || !values || !offsets3 || (M > 0 && (!out_coords || !pair_table)))
    return kErrInvalidArg;
  return kernel_map_search(keys, values, capacity, out_coords, M, offsets3, K, stride_x, stride_y,
                           stride_z, pair_table, block_counts,
                           reinterpret_cast<unsigned long long*>(mask_keys), S(stream));
}
int wcn_kernel_map_search_symmetric(const uint64_t* keys, const int32_t* values, int capacity,
                                    const int32_t* coords, int M, const int32_t* offsets3, int K,
                                    const int32_t* status, int32_t* pair_table, void* stream) {
  if (!keys || !values || !offsets3 || !status || (M > 0 && (!coords || !pair_table)))
    return kErrInvalidArg;
  return kernel_map_search_sym(keys, values, capacity, coords, M, offsets3, K, status, pair_table,
                               S(stream));
}
int wcn_kernel_map_stats(const int32_t* pair_table, int K, int M, int32_t* block_counts,
                         uint64_t* mask_keys, void* stream) {
  if (M > 0 && (!pair_table || !block_counts)) return kErrInvalidArg;
  return kernel_map_stats(pair_table, K, M, block_counts,
                          reinterpret_cast<unsigned long long*>(mask_keys), S(stream));
}
int wcn_kernel_map_count(int32_t* block_counts, int K, int num_blocks, int32_t* counts,
                         int32_t* offsets, void* stream) {
  if (!counts || !offsets || (num_blocks > 0 && !block_counts)) return kErrInvalidArg;
  return kernel_map_count(block_counts, K, num_blocks, counts, offsets, S(stream));
}
int wcn_kernel_map_scatter(const int32_t* pair_table, const int32_t* block_prefix,
                           const int32_t* offsets, int32_t* in_maps, int32_t* out_maps, int K,
                           int M, void* stream) {
  if (M > 0 && (!pair_table || !block_prefix || !offsets)) return kErrInvalidArg;
  return kernel_map_scatter(pair_table, block_prefix, offsets, in_maps, out_maps, K, M, S(stream));
}

int wcn_reverse_pair_table(const int32_t* pair_table, int K, int M, int32_t* rev, int n_in,
                           void* stream) {
  if (K < 1 || M < 0 || n_in < 0 || (n_in > 0 && !rev)) return kErrInvalidArg;
  return reverse_pair_table(pair_table, K, M, rev, n_in, S(stream));
}
int wcn_csr_to_pair_table(const int32_t* val_maps, const int32_t* row_maps, const int32_t* offsets,
                          int K, int n_rows, int num_pairs, int32_t* table, void* stream) {
  if (K < 1 || n_rows < 0 || (n_rows > 0 && !table) || !offsets) return kErrInvalidArg;
  if (num_pairs > 0 && (!val_maps || !row_maps)) return kErrInvalidArg;
  return csr_to_table(val_maps, row_maps, offsets, K, n_rows, num_pairs, table, S(stream));
}
int wcn_mask_keys(const int32_t* table, int K, int M, uint64_t* keys, void* stream) {
  if (M > 0 && (!table || !keys)) return kErrInvalidArg;
  return mask_keys_from_table(table, K, M, reinterpret_cast<unsigned long long*>(keys), S(stream));
}
size_t wcn_coords_unique_workspace_bytes(long long n_keys) {
  return coords_unique_workspace_bytes(n_keys);
}
int wcn_coords_unique(const int32_t* bcoords, int n, int stride_x, int stride_y, int stride_z,
                      const int32_t* offsets3, int K, int n_batches, int32_t* out_coords,
                      int32_t* first_index, int32_t* meta, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (!meta || (n > 0 && (!bcoords || !out_coords || !workspace))) return kErrInvalidArg;
  if (K > 1 && !offsets3) return kErrInvalidArg;
  return coords_unique(bcoords, n, stride_x, stride_y, stride_z, offsets3, K, n_batches,
                       out_coords, first_index, meta, workspace, workspace_bytes, S(stream));
}
size_t wcn_sort_workspace_bytes(int M) { return sort_workspace_bytes(M); }
int wcn_sort_rows_by_key(const uint64_t* keys, int M, int K, int32_t* rows_out, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (M > 0 && (!keys || !rows_out || !workspace)) return kErrInvalidArg;
  return sort_rows_by_key(reinterpret_cast<const unsigned long long*>(keys), M, K, rows_out,
                          workspace, workspace_bytes, S(stream));
}
int wcn_sort_rows_by_table(const int32_t* table, int K, int M, int32_t* rows_out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (M > 0 && (!table || !rows_out || !workspace)) return kErrInvalidArg;
  return sort_rows_by_table(table, K, M, rows_out, workspace, workspace_bytes, S(stream));
}
static int build_tiles_checked(const int32_t* table, int K, int M, const int32_t* sorted_rows,
                               int tile_rows, int m_pad, int32_t* step_nbr, int32_t* step_k,
                               int32_t* rows_padded, int32_t* tile_nk, int32_t* tile_cum,
                               int n_range_ctas, int32_t* cta_units,
                               const unsigned long long* row_masks, void* stream) {
  if (!tile_cum || !tile_nk) return kErrInvalidArg;
  if (m_pad > 0 && (!table || !sorted_rows || !step_nbr || !step_k || !rows_padded))
    return kErrInvalidArg;
  if (n_range_ctas < 0 || (n_range_ctas > 0 && !cta_units)) return kErrInvalidArg;
  return build_tiles(table, K, M, sorted_rows, tile_rows, m_pad, step_nbr, step_k, rows_padded,
                     tile_nk, tile_cum, n_range_ctas, cta_units, row_masks, S(stream));
}

int wcn_build_tiles(const int32_t* table, int K, int M, const int32_t* sorted_rows, int tile_rows,
                    int m_pad, int32_t* step_nbr, int32_t* step_k, int32_t* rows_padded,
                    int32_t* tile_nk, int32_t* tile_cum, int n_range_ctas, int32_t* cta_units,
                    void* stream) {
  return build_tiles_checked(table, K, M, sorted_rows, tile_rows, m_pad, step_nbr, step_k,
                             rows_padded, tile_nk, tile_cum, n_range_ctas, cta_units, nullptr,
                             stream);
}

int wcn_build_tiles_masked(const int32_t* table, int K, int M, const int32_t* sorted_rows,
                           int tile_rows, int m_pad, int32_t* step_nbr, int32_t* step_k,
                           int32_t* rows_padded, int32_t* tile_nk, int32_t* tile_cum,
                           int n_range_ctas, int32_t* cta_units,
                           const unsigned long long* row_masks, void* stream) {
  return build_tiles_checked(table, K, M, sorted_rows, tile_rows, m_pad, step_nbr, step_k,
                             rows_padded, tile_nk, tile_cum, n_range_ctas, cta_units, row_masks,
                             stream);
}

size_t wcn_knn_workspace_bytes(int n_ref, int n_batches) {
  return knn_workspace_bytes(n_ref, n_batches, knn_dims_for(n_ref, n_batches));
}
int wcn_knn_search(const float* ref, int n_ref, const int32_t* ref_offsets, const float* query,
                   int n_query, const int32_t* query_offsets, int n_batches, int k,
                   long long* out_idx, float* out_dist, void* workspace, size_t workspace_bytes,
                   void* stream) {
  if (!ref_offsets || !query_offsets || !workspace || (n_ref > 0 && !ref) ||
      (n_query > 0 && (!query || !out_idx)))
    return kErrInvalidArg;
  return knn_search(ref, n_ref, ref_offsets, query, n_query, query_offsets, n_batches, k, out_idx,
                    out_dist, workspace, workspace_bytes, S(stream));
}

size_t wcn_weight_image_bytes(int K, int groups, int cin_g, int cout_g, int dtype, int transpose_w,
                              int* n_slabs_out, int* gps_out) {
  const int rg = transpose_w ? cin_g : cout_g;
  const int cg = transpose_w ? cout_g : cin_g;
  SlabPlan plan;
  if (plan_slabs(groups, rg, cg, &plan) != kOk) return 0;
  if (n_slabs_out) *n_slabs_out = plan.n_slabs;
  if (gps_out) *gps_out = plan.gps;
  const int ce = 128 / dtype_size(dtype);
  const int n_chunks = (plan.cdim + ce - 1) / ce;
  return (size_t)plan.n_slabs * K * n_chunks * plan.bn * 128;
}

// fills the parameter block of one image; src_dtype may be fp32 with a 16-bit image dtype
static int weight_image_params(const void* weight, void* image, int K, int groups, int cin_g,
                               int cout_g, int dtype, int src_dtype, int transpose_w,
                               WeightPrepParams* out) {
  if (!weight || !image || K < 1) return kErrInvalidArg;
  if (dtype < 0 || dtype > 2 || src_dtype < 0 || src_dtype > 2) return kErrUnsupportedDtype;
  if (src_dtype != dtype && !(src_dtype == kF32 && dtype != kF32)) return kErrUnsupportedDtype;
  const int rg = transpose_w ? cin_g : cout_g;
  const int cg = transpose_w ? cout_g : cin_g;
  SlabPlan plan;
  int st = plan_slabs(groups, rg, cg, &plan);
  if (st != kOk) return st;
  WeightPrepParams p;
  p.w = weight;
  p.img = image;
  p.es = dtype_size(dtype);
  p.src_es = dtype_size(src_dtype);
  p.cvt = src_dtype == dtype ? 0 : (dtype == kBF16 ? 1 : 2);
  p.K = K;
  p.n_slabs = plan.n_slabs;
  p.gps = plan.gps;
  p.w_k_stride = (long long)groups * cin_g * cout_g;
  // element (k, g, ci, co) sits at k*K_stride + g*cin_g*cout_g + ci*cout_g + co
  const long long r_stride = transpose_w ? cout_g : 1;  // step of a row (rows = ci for dgrad)
  const long long c_stride = transpose_w ? 1 : cout_g;  // step of a contraction channel
  p.w_r_stride = r_stride;
  p.w_c_stride = c_stride;
  if (groups == 1) {
    // slabs are windows of rows of the single group
    p.rg = plan.bn;
    p.cg = cg;
    p.w_g_stride = (long long)plan.bn * r_stride;
  } else {
    p.rg = rg;
    p.cg = cg;
    p.w_g_stride = (long long)cin_g * cout_g;
  }
  const int ce = 128 / p.es;
  p.n_chunks = (plan.cdim + ce - 1) / ce;
  *out = p;
  return kOk;
}

int wcn_weight_image(const void* weight, void* image, int K, int groups, int cin_g, int cout_g,
                     int dtype, int transpose_w, void* stream) {
  WeightPrepParams p;
  const int st = weight_image_params(weight, image, K, groups, cin_g, cout_g, dtype, dtype,
                                     transpose_w, &p);
  if (st != kOk) return st;
  return launch_weight_image(p, nullptr, S(stream));
}

int wcn_weight_image_pair(const void* weight, int src_dtype, void* image_fwd, void* image_t, int K,
                          int groups, int cin_g, int cout_g, int dtype, void* stream) {
  WeightPrepParams a, b;
  int st = weight_image_params(weight, image_fwd, K, groups, cin_g, cout_g, dtype, src_dtype, 0, &a);
  if (st != kOk) return st;
  if (image_t == nullptr) return launch_weight_image(a, nullptr, S(stream));
  st = weight_image_params(weight, image_t, K, groups, cin_g, cout_g, dtype, src_dtype, 1, &b);
  if (st != kOk) return st;
  return launch_weight_image(a, &b, S(stream));
}

int wcn_gather_gemm(const void* feats, int n_in_rows, long long in_ld, const void* wimg, void* out,
                    long long out_ld, const int32_t* step_nbr, const int32_t* step_k,
                    const int32_t* rows, const int32_t* tile_nk, const int32_t* tile_cum,
                    int num_tiles, int tile_rows, int m_pad, int K, int groups, int cin_g,
                    int cout_g, int dtype, const float* bias, int relu, int kflip, int max_ctas,
                    const int32_t* cta_units, int n_range_ctas, double* stats, void* stream) {
  if (num_tiles > 0 && (!feats || !wimg || !out || !step_nbr || !step_k || !rows || !tile_nk ||
                        !tile_cum))
    return kErrInvalidArg;
  if (dtype < 0 || dtype > 2) return kErrUnsupportedDtype;
  if (num_tiles == 0) return kOk;
  SlabPlan plan;
  int st = plan_slabs(groups, cout_g, cin_g, &plan);
  if (st != kOk) return st;
  GatherGemmParams p;
  if (n_in_rows < 0 || (in_ld * dtype_size(dtype)) % 16 != 0 ||
      (reinterpret_cast<uintptr_t>(feats) & 15))
    return n_in_rows < 0 ? kErrInvalidArg : kErrAlignment;
  p.n_in_rows = n_in_rows;
  p.feats = feats;
  p.wimg = wimg;
  p.out = out;
  p.step_nbr = step_nbr;
  p.step_k = step_k;
  p.rows = rows;
  p.tile_nk = tile_nk;
  p.tile_cum = tile_cum;
  p.cta_units = (cta_units != nullptr && n_range_ctas > 0) ? cta_units : nullptr;
  p.tile_rows = tile_rows;
  p.bias = bias;
  p.in_ld = in_ld;
  p.out_ld = out_ld;
  p.in_coff = 0;
  p.in_slab_stride = (groups == 1) ? 0 : plan.cdim;
  p.out_coff = 0;
  p.cin = plan.cdim;
  p.bn = plan.bn;
  p.K = K;
  p.m_pad = m_pad;
  p.num_tiles = num_tiles;
  p.kflip = kflip;
  p.stages = 0;
  p.relu = relu;
  p.stats = stats;
  p.stats_c = groups * cout_g;
  p.debug = 0;
  p.dbg_out = nullptr;
#ifdef WCN_BRINGUP  // bring-up builds only (WCN_BRINGUP=1 build.sh): experiment switches of tools/exp_*.py
  {
    const char* e = getenv("WCN_DEBUG");
    p.debug = e ? atoi(e) : 0;
    const char* dp = getenv("WCN_DEBUG_PTR");
    p.dbg_out = dp ? reinterpret_cast<long long*>(strtoull(dp, nullptr, 10)) : nullptr;
    const char* st_env = getenv("WCN_STAGES");
    if (st_env) p.stages = atoi(st_env);
  }
#endif
  if (max_ctas <= 0) max_ctas = sm_count();
  return launch_gather_gemm(p, dtype, plan.n_slabs, max_ctas, n_range_ctas, S(stream));
}

int wcn_wgrad(const void* feats, long long in_ld, const void* gout, long long out_ld, float* dw,
              const int32_t* in_maps, const int32_t* out_maps, const int32_t* offsets, int K,
              int groups, int cin_g, int cout_g, int dtype, float alpha, int unit_pairs,
              int max_ctas, const int32_t* row_block_prefix, int n_row_blocks, int row_parts,
              int rounds, int identity_k, const int32_t* status, long long n_in_rows,
              long long n_out_rows, void* stream) {
  if (!feats || !gout || !dw || !offsets) return kErrInvalidArg;
  if (dtype < 0 || dtype > 2) return kErrUnsupportedDtype;
  if (groups < 1 || cin_g < 1 || cout_g < 1) return kErrInvalidArg;
  WgradParams p;
  p.feats = feats;
  p.gout = gout;
  p.dw = dw;
  p.in_maps = in_maps;
  p.out_maps = out_maps;
  p.offsets = offsets;
  p.in_ld = in_ld;
  p.out_ld = out_ld;
  p.in_coff = 0;
  p.out_coff = 0;
  p.K = K;
  p.unit_pairs = unit_pairs;
  p.stages = 0;
  p.alpha = alpha;
  // row-block-major unit order (optional): needs the kernel map's block prefix and a unit table
  // that fits shared memory; anything else silently keeps the offset-major order
  p.blk_prefix = nullptr;
  p.n_row_blocks = 0;
  p.row_parts = 1;
  p.rounds = 1;
  p.identity_k = (identity_k >= 0 && identity_k < K && n_in_rows > 0 && n_out_rows > 0) ? identity_k
                                                                                        : -1;
  p.status = status;
  if (row_block_prefix != nullptr && row_parts >= 1 && rounds >= 1 && n_row_blocks >= row_parts &&
      (long long)row_parts * K <= 1024 && (row_parts > 1 || rounds > 1)) {
    p.blk_prefix = row_block_prefix;
    p.n_row_blocks = n_row_blocks;
    p.row_parts = row_parts;
    p.rounds = rounds;
  }
  p.debug = 0;
  p.dbg_out = nullptr;
#ifdef WCN_BRINGUP  // bring-up builds only: experiment switches of tools/exp_wgrad*.py
  {
    const char* e = getenv("WCN_DEBUG");
    p.debug = e ? atoi(e) : 0;
    const char* dp = getenv("WCN_DEBUG_PTR");
    p.dbg_out = dp ? reinterpret_cast<long long*>(strtoull(dp, nullptr, 10)) : nullptr;
  }
#endif
  p.dw_k_stride = (long long)groups * cin_g * cout_g;
  p.dw_g_stride = (long long)cin_g * cout_g;
  p.dw_ld = cout_g;
  int y_slabs, z_slabs;
  if (groups == 1) {
    y_slabs = (cin_g + 127) / 128;
    z_slabs = (cout_g + 255) / 256;
    while (z_slabs <= cout_g && (cout_g % z_slabs != 0 || (cout_g / z_slabs) % 16 != 0)) ++z_slabs;
    if (z_slabs > cout_g) return kErrUnsupportedShape;
    p.cin = cin_g < 128 ? cin_g : 128;
    p.cin_last = cin_g - 128 * (y_slabs - 1);
    p.cout = cout_g / z_slabs;
    p.gps = 1;
    p.cin_g = cin_g;
    p.cout_g = cout_g;
    p.in_y_stride = 128;
    p.out_y_stride = 0;
    p.out_z_stride = p.cout;
    p.dw_y_stride = 128LL * cout_g;
    p.dw_z_stride = p.cout;
  } else {
    SlabPlan plan;
    int st = plan_slabs(groups, cout_g, cin_g, &plan);
    if (st != kOk) return st;
    if (plan.gps * cin_g > 128 || plan.gps * cout_g > 256 || cout_g % 8 != 0)
      return kErrUnsupportedShape;
    y_slabs = plan.n_slabs;
    z_slabs = 1;
    p.gps = plan.gps;
    p.cin = plan.gps * cin_g;
    p.cin_last = p.cin;
    p.cout = plan.gps * cout_g;
    p.cin_g = cin_g;
    p.cout_g = cout_g;
    p.in_y_stride = p.cin;
    p.out_y_stride = p.cout;
    p.out_z_stride = 0;
    p.dw_y_stride = (long long)plan.gps * p.dw_g_stride;
    p.dw_z_stride = 0;
    if (plan.gps == 1) {
      // one (large) group per slab behaves like a dense slab
      p.dw_ld = cout_g;
    }
  }
  if (max_ctas <= 0) max_ctas = sm_count();
  return launch_wgrad(p, dtype, y_slabs, z_slabs, max_ctas, n_in_rows, n_out_rows, S(stream));
}

/* ---- per-channel normalisation / activation passes over the feature matrix (rownorm.cu) ---- */
static RowNormParams rn_params(int n, int c) {
  RowNormParams p{};
  p.n = n;
  p.c = c;
  return p;
}

int wcn_bn_stats(const void* x, long long ld_x, int n, int c, int dtype, double* sums,
                 void* stream) {
  if (!x || !sums) return kErrInvalidArg;
  RowNormParams p = rn_params(n, c);
  p.x = x; p.ld_x = ld_x; p.sums = sums;
  return rownorm_launch(0, p, dtype, S(stream));
}

int wcn_bn_finalize(const double* sums, int n, int c, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    float* scale, float* shift, float* mean_rstd, void* stream) {
  if (!sums || !scale || !shift || !mean_rstd) return kErrInvalidArg;
  return bn_finalize(sums, n, c, gamma, beta, eps, momentum, running_mean, running_var, scale,
                     shift, mean_rstd, S(stream));
}

int wcn_scale_shift_act(const void* x, long long ld_x, const void* res, long long ld_res, void* y,
                        long long ld_y, int n, int c, int dtype, const float* scale,
                        const float* shift, int relu, void* stream) {
  if (!x || !y || !scale || !shift) return kErrInvalidArg;
  RowNormParams p = rn_params(n, c);
  p.x = x; p.ld_x = ld_x; p.res = res; p.ld_res = ld_res; p.y = y; p.ld_y = ld_y;
  p.scale = scale; p.shift = shift; p.relu = relu;
  return rownorm_launch(1, p, dtype, S(stream));
}

int wcn_bn_bwd_reduce(const void* dy, long long ld_dy, const void* x, long long ld_x,
                      const void* y, long long ld_y, int n, int c, int dtype,
                      const float* mean_rstd, const float* mask_scale, const float* mask_shift,
                      double* sums, void* stream) {
  if (!dy || !x || !mean_rstd || !sums) return kErrInvalidArg;
  RowNormParams p = rn_params(n, c);
  p.dy = dy; p.ld_dy = ld_dy; p.x = x; p.ld_x = ld_x; p.y_in = y; p.ld_yin = ld_y;
  p.mean_rstd = mean_rstd; p.sums = sums;
  if (mask_scale && mask_shift) { p.mask_scale = mask_scale; p.mask_shift = mask_shift; }
  return rownorm_launch(2, p, dtype, S(stream));
}

int wcn_bn_bwd_apply(const void* dy, long long ld_dy, const void* x, long long ld_x, const void* y,
                     long long ld_y, void* dx, long long ld_dx, void* dres, long long ld_dres,
                     int n, int c, int dtype, const float* gamma, const float* mean_rstd,
                     const double* sums, const float* mask_scale, const float* mask_shift,
                     int training, void* stream) {
  if (!dy || !dx || !gamma) return kErrInvalidArg;
  if (training && (!x || !mean_rstd || !sums)) return kErrInvalidArg;
  RowNormParams p = rn_params(n, c);
  p.dy = dy; p.ld_dy = ld_dy; p.x = x; p.ld_x = ld_x; p.y_in = y; p.ld_yin = ld_y;
  p.y = dx; p.ld_y = ld_dx; p.dres = dres; p.ld_dres = ld_dres;
  p.scale = gamma; p.mean_rstd = mean_rstd; p.sums = const_cast<double*>(sums);
  p.training = training;
  if (mask_scale && mask_shift) { p.mask_scale = mask_scale; p.mask_shift = mask_shift; }
  return rownorm_launch(3, p, dtype, S(stream));
}

/* One call per layer and direction: fewer host round trips for the host-paced small levels. */
int wcn_bn_forward(const void* x, long long ld_x, const void* res, long long ld_res, void* y,
                   long long ld_y, int n, int c, int dtype, const float* gamma, const float* beta,
                   float eps, float momentum, float* running_mean, float* running_var,
                   double* sums, float* scale_shift_mean_rstd, int relu, void* stream) {
  if (!x || !y || !sums || !scale_shift_mean_rstd || n < 1 || c < 1) return kErrInvalidArg;
  if (cudaMemsetAsync(sums, 0, (size_t)2 * c * sizeof(double), S(stream)) != cudaSuccess)
    return kErrCuda;
  float* scale = scale_shift_mean_rstd;
  float* shift = scale + c;
  float* mean_rstd = scale + 2 * c;
  int st = wcn_bn_stats(x, ld_x, n, c, dtype, sums, stream);
  if (st != kOk) return st;
  st = wcn_bn_finalize(sums, n, c, gamma, beta, eps, momentum, running_mean, running_var, scale,
                       shift, mean_rstd, stream);
  if (st != kOk) return st;
  return wcn_scale_shift_act(x, ld_x, res, ld_res, y, ld_y, n, c, dtype, scale, shift, relu, stream);
}

int wcn_bn_backward(const void* dy, long long ld_dy, const void* x, long long ld_x, const void* y,
                    long long ld_y, void* dx, long long ld_dx, void* dres, long long ld_dres, int n,
                    int c, int dtype, const float* gamma, const float* mean_rstd,
                    const float* mask_scale, const float* mask_shift, double* sums, void* stream) {
  if (!dy || !x || !dx || !gamma || !mean_rstd || !sums || n < 1 || c < 1) return kErrInvalidArg;
  if (cudaMemsetAsync(sums, 0, (size_t)2 * c * sizeof(double), S(stream)) != cudaSuccess)
    return kErrCuda;
  int st = wcn_bn_bwd_reduce(dy, ld_dy, x, ld_x, y, ld_y, n, c, dtype, mean_rstd, mask_scale,
                             mask_shift, sums, stream);
  if (st != kOk) return st;
  return wcn_bn_bwd_apply(dy, ld_dy, x, ld_x, y, ld_y, dx, ld_dx, dres, ld_dres, n, c, dtype, gamma,
                          mean_rstd, sums, mask_scale, mask_shift, 1, stream);
}

/* ---- radius search (knn.cu) ---- */
int wcn_radius_count(const float* ref, int n_ref, const int32_t* ref_offsets, const float* query,
                     int n_query, const int32_t* query_offsets, int n_batches, float radius,
                     int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  if (!ref_offsets || !query_offsets || !workspace || (n_query > 0 && (!query || !counts)) ||
      (n_ref > 0 && !ref))
    return kErrInvalidArg;
  return radius_count(ref, n_ref, ref_offsets, query, n_query, query_offsets, n_batches, radius,
                      counts, workspace, workspace_bytes, S(stream));
}

int wcn_radius_fill(int n_ref, const float* query, int n_query, const int32_t* query_offsets,
                    int n_batches, float radius, const long long* row_splits, int32_t* out_idx,
                    float* out_dist, void* workspace, size_t workspace_bytes, void* stream) {
  if (!query_offsets || !workspace || (n_query > 0 && (!query || !row_splits || !out_idx)))
    return kErrInvalidArg;
  return radius_fill(n_ref, query, n_query, query_offsets, n_batches, radius, row_splits, out_idx,
                     out_dist, workspace, workspace_bytes, S(stream));
}

/* ---- depthwise sparse convolution (conv_depthwise.cu) ---- */
int wcn_depthwise_conv(const void* feats, long long in_ld, void* out, long long out_ld,
                       const float* weight, const float* bias, const int32_t* table, int n_rows,
                       int K, int channels, int dtype, int kflip, int relu, void* stream) {
  if (!feats || !out || !weight || !table) return kErrInvalidArg;
  return depthwise_launch(false, feats, in_ld, nullptr, 0, out, out_ld, weight, nullptr, bias, table,
                          n_rows, K, channels, kflip, relu, dtype, S(stream));
}

int wcn_depthwise_conv_plan(const void* feats, long long in_ld, void* out, long long out_ld,
                            const float* weight, const float* bias, const int32_t* step_nbr,
                            const int32_t* step_k, const int32_t* rows, const int32_t* tile_nk,
                            int num_tiles, int tile_rows, int K, int channels, int dtype, int kflip,
                            int relu, void* stream) {
  if (!feats || !out || !weight || !step_nbr || !step_k || !rows || !tile_nk) return kErrInvalidArg;
  return depthwise_plan_launch(feats, in_ld, out, out_ld, weight, bias, step_nbr, step_k, rows,
                               tile_nk, num_tiles, tile_rows, K, channels, kflip, relu, dtype,
                               S(stream));
}

int wcn_depthwise_wgrad_plan(const void* feats, long long in_ld, const void* gout,
                             long long gout_ld, float* dw, const int32_t* step_nbr,
                             const int32_t* step_k, const int32_t* rows, const int32_t* tile_nk,
                             int num_tiles, int tile_rows, int K, int channels, int dtype,
                             void* stream) {
  if (!feats || !gout || !dw || !step_nbr || !step_k || !rows || !tile_nk) return kErrInvalidArg;
  return depthwise_plan_wgrad_launch(feats, in_ld, gout, gout_ld, dw, step_nbr, step_k, rows,
                                     tile_nk, num_tiles, tile_rows, K, channels, dtype, S(stream));
}

int wcn_depthwise_wgrad(const void* feats, long long in_ld, const void* gout, long long gout_ld,
                        float* dw, const int32_t* table, int n_rows, int K, int channels,
                        int dtype, void* stream) {
  if (!feats || !gout || !dw || !table) return kErrInvalidArg;
  return depthwise_launch(true, feats, in_ld, gout, gout_ld, nullptr, 0, nullptr, dw, nullptr, table,
                          n_rows, K, channels, 0, 0, dtype, S(stream));
}

int wcn_peer_allreduce_flag_words(void) { return peer_allreduce_flag_words(); }
int wcn_peer_allreduce_timeout_word(void) { return peer_allreduce_timeout_word(); }

int wcn_peer_allreduce_f32(void* const* bufs, void* const* flags, int rank, int world,
                           long long n, float scale, int n_ctas, void* stream) {
  if (!bufs || !flags) return kErrInvalidArg;
  return peer_allreduce_f32(bufs, flags, rank, world, n, scale, n_ctas, S(stream));
}

}  // extern "C"
