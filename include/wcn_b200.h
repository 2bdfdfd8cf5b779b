/* SPDX-License-Identifier: Apache-2.0
 *
 * wcn_b200 — C-ABI of the B200-native sparse-convolution hot path.
 *
 * This header is the drop-in boundary: every entry point replaces one pybind11 binding of the
 * reference's `warpconvnet._C` extension (cited per function as file:line under the reference
 * tree). Conventions shared by all functions:
 *   - plain device pointers + sizes, no framework types; the CALLER owns and allocates every
 *     buffer including outputs and scratch (same ownership rule as the reference, whose Python
 *     side allocates all tensors: detail/mask_gemm.py:723,874,934);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call only
 *     enqueues work on it and never synchronises;
 *   - return value: 0 = success, negative = error (same "int status" convention as
 *     csrc/include/gemm_error_codes.h:7-15): -1 invalid argument, -2 unsupported shape,
 *     -3 misaligned pointer/stride, -4 unsupported dtype, -5 CUDA launch error,
 *     -6 workspace too small;
 *   - dtype codes: 0 = bf16, 1 = fp16, 2 = fp32 (computed as TF32 on the tensor cores;
 *     wcn_wgrad takes 16-bit operands only and returns -4 for fp32 — the host splits fp32 into
 *     bf16 hi/lo parts);
 *     accumulation is always fp32;
 *   - device-side failures (hash table full, coordinate out of the packed-key range) are reported
 *     through a caller-provided int status word, like the reference's status tensor
 *     (csrc/include/cuhash/hash_table.cuh:59-62).
 */
#ifndef WCN_B200_H_
#define WCN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WCN_OK 0
#define WCN_ERR_INVALID_ARG (-1)
#define WCN_ERR_UNSUPPORTED_SHAPE (-2)
#define WCN_ERR_ALIGNMENT (-3)
#define WCN_ERR_UNSUPPORTED_DTYPE (-4)
#define WCN_ERR_CUDA (-5)
#define WCN_ERR_WORKSPACE (-6)

#define WCN_BF16 0
#define WCN_F16 1
#define WCN_F32 2

/* library / build identification (reference: csrc/warpconvnet_pybind.cpp:21-32 `__build_commit__`) */
const char* wcn_version(void);
/* 1 when the library was compiled for sm_100a (always, this build has no other target). */
int wcn_built_for_sm100a(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py's
 * `gpu_launches` evidence; library kernels such as the CUB radix sort are not counted). */
long long wcn_launch_count(void);

/* ------------------------------------------------------------------------------------------ */
/* Packed-coordinate hash table  (replaces _C.cuhash.packed_prepare / packed_insert /        */
/* packed_search: csrc/bindings/cuhash_bindings.cpp:241-330, csrc/cuhash_hash_table.cu:19-100) */
/* keys: uint64[capacity], values: int32[capacity], capacity a power of two >= 2*n.           */
/* Key layout 1|9b batch|18b x|18b y|18b z (csrc/include/cuhash/hash_functions.cuh:29-44).    */
/* ------------------------------------------------------------------------------------------ */
int wcn_hash_prepare(uint64_t* keys, int32_t* values, int capacity, void* stream);
/* coords: int32[n][4] = (batch, x, y, z). value = insertion index (smallest index wins for
 * duplicates). status (device int, caller-zeroed): bit2 = duplicate coordinates seen (not an
 * error), bit0 = table full, bit1 = coordinate outside
 * batch [0,511] / xyz [-131072,131071] (the reference checks the range on the host,
 * geometry/coords/search/packed_hashmap.py:66-82). */
int wcn_hash_insert(uint64_t* keys, int32_t* values, const int32_t* coords, int n, int capacity,
                    int32_t* status, void* stream);
/* results[i] = stored index of queries[i] or -1. */
int wcn_hash_search(const uint64_t* keys, const int32_t* values, const int32_t* queries,
                    int32_t* results, int n, int capacity, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Kernel map  (replaces _C.cuhash.packed_kernel_map_size / packed_kernel_map_offset /        */
/* postprocess_count / postprocess_scatter: csrc/cuhash_kernel_map.cu:93-134,508-599; call    */
/* sites geometry/coords/search/torch_discrete.py:154,254-287)                                */
/* ------------------------------------------------------------------------------------------ */
/* number of 256-query blocks the search/scatter passes use for M queries */
int wcn_kernel_map_num_blocks(int M);
/* pair_table[k*M + m] = index of the input voxel at out_coords[m]*stride + offsets3[k], or -1.
 * offsets3: device int32[K][3] (x,y,z offset of kernel element k, reference order
 * csrc/include/cuhash/kernel_map.cuh:34-54). block_counts: int32[K][num_blocks] (optional, may be
 * NULL) receives per-block hit counts; mask_keys: uint64[M] (optional) receives the per-row offset
 * bitmask (bit k, folded modulo 64 when K > 64). */
int wcn_kernel_map_search(const uint64_t* keys, const int32_t* values, int capacity,
                          const int32_t* out_coords, int M, const int32_t* offsets3, int K,
                          int stride_x, int stride_y, int stride_z, int32_t* pair_table,
                          int32_t* block_counts, uint64_t* mask_keys, void* stream);
/* Submanifold fast path (query coordinates == the table's coordinates, odd kernel, stride 1):
 * probes only the first K/2 offsets plus the centre and mirrors every hit into offset K-1-k
 * (same idea as the reference's skip_symmetric_kernel_map, torch_discrete.py:296-432). `status`
 * is the word wcn_hash_insert wrote; if it reports duplicate coordinates (bit 2) every offset is
 * probed instead, so the result always equals wcn_kernel_map_search. Fills the whole table. */
int wcn_kernel_map_search_symmetric(const uint64_t* keys, const int32_t* values, int capacity,
                                    const int32_t* coords, int M, const int32_t* offsets3, int K,
                                    const int32_t* status, int32_t* pair_table, void* stream);
/* block_counts[K][num_blocks] and mask_keys[M] (optional) from a finished pair table. */
int wcn_kernel_map_stats(const int32_t* pair_table, int K, int M, int32_t* block_counts,
                         uint64_t* mask_keys, void* stream);
/* In-place exclusive scan of block_counts per offset; counts[K], offsets[K+1] (exclusive scan of
 * counts, offsets[K] = total pairs L). */
int wcn_kernel_map_count(int32_t* block_counts, int K, int num_blocks, int32_t* counts,
                         int32_t* offsets, void* stream);
/* CSR emission, ascending output row inside every offset: in_maps/out_maps int32[L]. */
int wcn_kernel_map_scatter(const int32_t* pair_table, const int32_t* block_prefix,
                           const int32_t* offsets, int32_t* in_maps, int32_t* out_maps, int K,
                           int M, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Mask / tile preparation  (replaces _C.mask_gemm build_pair_mask / mask_argsort /           */
/* build_reverse_mask_data: csrc/mask_data_kernels.cu:23-220, detail/mask_gemm.py:127-377)    */
/* ------------------------------------------------------------------------------------------ */
/* rev[k*n_in + i] = m for every pair_table[k*M + m] = i >= 0, -1 elsewhere. */
int wcn_reverse_pair_table(const int32_t* pair_table, int K, int M, int32_t* rev, int n_in,
                           void* stream);
/* Dense table from the CSR lists (replaces csr_to_pair_table, csrc/mask_data_kernels.cu:55-82):
 * table[k*n_rows + row_maps[j]] = val_maps[j] for every pair j of offset k, -1 elsewhere.
 * (val=in_maps,row=out_maps) gives the forward pair table, swapped gives the reverse table. */
int wcn_csr_to_pair_table(const int32_t* val_maps, const int32_t* row_maps, const int32_t* offsets,
                          int K, int n_rows, int num_pairs, int32_t* table, void* stream);
int wcn_mask_keys(const int32_t* table, int K, int M, uint64_t* keys, void* stream);
/* ---- coordinate-set operations ----------------------------------------------
 * Replaces the torch glue of warpconvnet/geometry/coords/ops/stride.py:18-56 (stride_coords:
 * floor division + hash-unique + batch argsort), ops/expand.py:17-75 (expand_coords) and
 * geometry/types/voxels.py:271-278 (Voxels.unique) with one device chain and NO host sync:
 *   rows = unique{ (b, floor(x / sx) + ox_k, floor(y / sy) + oy_k, floor(z / sz) + oz_k) }
 * over all input rows and all K offsets (offsets3 = NULL with K = 1: no offsets), sorted by
 * (batch, x, y, z).
 *   out_coords  [n * K, 4] int32 upper-bound buffer; the first meta[n_batches] rows are valid
 *   first_index [n * K] int32 or NULL: smallest source row of every output row (K = 1 only)
 *   meta        [n_batches + 3] int32: offsets[0..n_batches] (per-batch row offsets), total,
 *               status (bit 1 = a coordinate left the packed range b 0..511, xyz -131072..131071,
 *               or b >= n_batches; such rows are dropped)
 *   workspace   wcn_coords_unique_workspace_bytes(n * K) bytes */
size_t wcn_coords_unique_workspace_bytes(long long n_keys);
int wcn_coords_unique(const int32_t* bcoords, int n, int stride_x, int stride_y, int stride_z,
                      const int32_t* offsets3, int K, int n_batches, int32_t* out_coords,
                      int32_t* first_index, int32_t* meta, void* workspace, size_t workspace_bytes,
                      void* stream);
size_t wcn_sort_workspace_bytes(int M);
/* rows_out = stable argsort of the row masks: rows with equal masks stay adjacent in ascending row
 * order, similar masks close. K <= 24 and K > 32: numeric order of the low min(K,64) bits;
 * 24 < K <= 32: order of a 24-bit compression of the mask (three radix passes instead of four; the
 * centre bit of an odd K is dropped, the lowest bits are folded in — see cuhash.cu; bring-up builds
 * keep the numeric order with WCN_FOLD_MASK_KEYS=0). Any order is a valid plan. */
int wcn_sort_rows_by_key(const uint64_t* keys, int M, int K, int32_t* rows_out, void* workspace,
                         size_t workspace_bytes, void* stream);
/* Same order, with the masks derived from the table inside the sort kernel's first pass (no
 * wcn_mask_keys / wcn_kernel_map_stats pass over the table on the critical path). Returns -2
 * (unsupported shape) for K > 32 or M > 2^20: use wcn_mask_keys + wcn_sort_rows_by_key then. */
int wcn_sort_rows_by_table(const int32_t* table, int K, int M, int32_t* rows_out, void* workspace,
                           size_t workspace_bytes, void* stream);
/* Tile plan in mask-sorted order. tile_rows is 128 or 256, m_pad = ceil(M/tile_rows)*tile_rows,
 * num_tiles = m_pad / tile_rows. For tile t (sorted positions t*tile_rows ...):
 *   rows_padded[p]                 = sorted_rows[p]  (-1 for padding)
 *   tile_nk[t]                     = number of kernel offsets active anywhere in the tile (steps)
 *   step_k[t*K + i]                = kernel offset of step i < tile_nk[t] (ascending)
 *   step_nbr[(t*K + i)*tile_rows + r] = table[step_k][sorted row r of the tile]  (-1 = none)
 *   tile_cum[0..num_tiles]         = exclusive prefix sum of tile_nk (work balancing)
 *   cta_units[0..n_range_ctas]     = (optional, n_range_ctas > 0) first 128-row unit of each of
 *                                    n_range_ctas CTAs of wcn_gather_gemm, balanced by step count;
 *                                    pass the same array and count to wcn_gather_gemm, whose
 *                                    prologue then skips its own search when its grid matches
 * step_nbr is sized for the upper bound K*m_pad ints, step_k for K*num_tiles ints. */
int wcn_build_tiles(const int32_t* table, int K, int M, const int32_t* sorted_rows, int tile_rows,
                    int m_pad, int32_t* step_nbr, int32_t* step_k, int32_t* rows_padded,
                    int32_t* tile_nk, int32_t* tile_cum, int n_range_ctas, int32_t* cta_units,
                    void* stream);
/* Same, with the per-row offset masks of wcn_mask_keys / wcn_kernel_map_stats (bit k of
 * row_masks[row] <=> table[k][row] >= 0; exact for K <= 64, ignored above; may be NULL): every row
 * then reads only the table entries it has instead of all K. */
int wcn_build_tiles_masked(const int32_t* table, int K, int M, const int32_t* sorted_rows,
                           int tile_rows, int m_pad, int32_t* step_nbr, int32_t* step_k,
                           int32_t* rows_padded, int32_t* tile_nk, int32_t* tile_cum,
                           int n_range_ctas, int32_t* cta_units,
                           const unsigned long long* row_masks, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Batched exact k-nearest-neighbour search for Points / PointConv                            */
/* (replaces geometry/coords/search/knn.py:10-142: per-batch chunked torch.cdist + topk)      */
/* ------------------------------------------------------------------------------------------ */
size_t wcn_knn_workspace_bytes(int n_ref, int n_batches);
/* ref float32[n_ref][3], query float32[n_query][3]; *_offsets: device int32[n_batches + 1] row
 * ranges of each batch item; every batch item must hold >= k reference points, 1 <= k <= 64.
 * out_idx int64[n_query][k]: GLOBAL reference rows, ascending distance (ties: smaller index);
 * out_dist (optional) float32[n_query][k] Euclidean distances. */
int wcn_knn_search(const float* ref, int n_ref, const int32_t* ref_offsets, const float* query,
                   int n_query, const int32_t* query_offsets, int n_batches, int k,
                   long long* out_idx, float* out_dist, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Radius search on the same grid (replaces geometry/coords/search/radius.py:16-291,
 * csrc/radius_search_kernels.cu:17-133). Two passes sharing one workspace of
 * wcn_knn_workspace_bytes(n_ref, n_batches) bytes that must stay untouched in between:
 *   wcn_radius_count builds the grid and writes counts[q] = reference points of q's batch item
 *     within `radius` (Euclidean, <=) of query q;
 *   the caller turns counts into int64 row_splits[n_query + 1] (exclusive scan) and allocates
 *     out_idx / out_dist with row_splits[n_query] entries;
 *   wcn_radius_fill writes the CSR lists: GLOBAL reference rows (order inside a row
 *     unspecified, as in the reference) and, optionally, their distances. */
int wcn_radius_count(const float* ref, int n_ref, const int32_t* ref_offsets, const float* query,
                     int n_query, const int32_t* query_offsets, int n_batches, float radius,
                     int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);
int wcn_radius_fill(int n_ref, const float* query, int n_query, const int32_t* query_offsets,
                    int n_batches, float radius, const long long* row_splits, int32_t* out_idx,
                    float* out_dist, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Weight image for the gather-GEMM kernel                                                    */
/* (replaces weight.transpose(1,2).contiguous(), detail/unified.py:654-671)                   */
/* ------------------------------------------------------------------------------------------ */
/* bytes of the image for the given problem */
size_t wcn_weight_image_bytes(int K, int groups, int cin_g, int cout_g, int dtype, int transpose_w,
                              int* n_slabs_out, int* gps_out);
/* weight: [K][groups][cin_g][cout_g] contiguous (groups = 1 for a dense conv).
 * transpose_w = 0: forward image (rows = output channels, contraction over input channels);
 * transpose_w = 1: dgrad image (rows = input channels, contraction over output channels). */
int wcn_weight_image(const void* weight, void* image, int K, int groups, int cin_g, int cout_g,
                     int dtype, int transpose_w, void* stream);
/* One launch for a layer's step: the forward image and (image_t != NULL) the dgrad image of the
 * same weights. src_dtype is the dtype of `weight`: equal to `dtype`, or fp32 master weights
 * converted (round to nearest even) to a 16-bit image dtype on the way
 * (replaces weight.to(compute_dtype) + the two calls above). */
int wcn_weight_image_pair(const void* weight, int src_dtype, void* image_fwd, void* image_t, int K,
                          int groups, int cin_g, int cout_g, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* The three sparse-conv GEMMs                                                                */
/* ------------------------------------------------------------------------------------------ */
/* forward AB_gather_scatter and dgrad ABt_gather_scatter
 * (replaces _C.mask_gemm.fwd / .dgrad: csrc/bindings/mask_gemm_bindings.cu:993-1750,2071-2116;
 *  semantics detail/explicit.py:22-57,60-101):
 *   out[rows[p], :] = sum over the steps i of p's tile of feats[step_nbr[i][p], :] @ W[step_k[i]]
 *   (rows not listed are untouched; every listed row is overwritten, so `out` needs no zero-fill)
 * feats [n_in_rows, in_ld] (16-byte aligned base and pitch), out [n_out, out_ld]; channels: cin_total = groups*cin_g gathered per row
 * (dgrad: pass cout/cin swapped and a transpose_w=1 image); bias (optional fp32[groups*cout_g]);
 * kflip=1 uses weight K-1-k for table row k (dgrad of a submanifold conv on the forward table).
 * The plan arrays come from wcn_build_tiles.  * stats (optional, fp64 [2][groups * cout_g], caller zero-fills): per-channel sum and sum of squares
 * of the output AS STORED (after bias / ReLU / rounding to the feature dtype), accumulated in the
 * epilogue — the statistics pass of the BatchNorm that follows the conv in the reference's ConvBlock
 * (models/mink_unet.py:31-53) without re-reading Y. NULL = off. */
int wcn_gather_gemm(const void* feats, int n_in_rows, long long in_ld, const void* wimg, void* out,
                    long long out_ld, const int32_t* step_nbr, const int32_t* step_k,
                    const int32_t* rows, const int32_t* tile_nk, const int32_t* tile_cum,
                    int num_tiles, int tile_rows, int m_pad, int K, int groups, int cin_g,
                    int cout_g, int dtype, const float* bias, int relu, int kflip, int max_ctas,
                    const int32_t* cta_units, int n_range_ctas, double* stats, void* stream);

/* wgrad AtB_gather_gather
 * (replaces _C.mask_gemm.wgrad: csrc/bindings/mask_gemm_bindings.cu:1755-2040 and
 *  _C.gemm.cutlass_gemm_trAB_gather: csrc/bindings/gemm_bindings.cpp:810-918;
 *  semantics detail/explicit.py:95-97):
 *   dw[k][g][ci][co] += alpha * sum over pairs j of offset k of feats[in_maps[j], g*cin_g+ci] *
 *                                                     gout[out_maps[j], g*cout_g+co]
 * dw is fp32 [K][groups][cin_g][cout_g], accumulated into (caller zero-fills);
 * offsets: device int32[K+1].
 * Optional L2-locality order: row_block_prefix = the [K][n_row_blocks] array wcn_kernel_map_count
 * leaves behind (pairs of offset k whose output row is < 256*b; requires out_maps ascending inside
 * every offset, which wcn_kernel_map_scatter guarantees). The pair lists are then walked in
 * `row_parts` row blocks x K offsets, every CTA taking `rounds` round-robin chunks, so a block of
 * rows sees all K offsets while L2-resident. NULL / (1, 1) = plain offset-major order.
 * Optional identity offset: identity_k >= 0 promises that offset identity_k of the map is the
 * identity (in_maps == out_maps == 0 .. n-1 in order: the centre offset of a submanifold map with
 * unique coordinates); its contiguous rows are then fetched as 2-D TMA tiles instead of row gathers.
 * `status` (optional) = the hash table's status word, whose duplicate bit (4) switches the
 * shortcut off on the device; n_in_rows / n_out_rows = rows of feats / gout (tensor-map bounds).
 * Pass -1 / NULL / 0 / 0 when unknown. */
int wcn_wgrad(const void* feats, long long in_ld, const void* gout, long long out_ld, float* dw,
              const int32_t* in_maps, const int32_t* out_maps, const int32_t* offsets, int K,
              int groups, int cin_g, int cout_g, int dtype, float alpha, int unit_pairs,
              int max_ctas, const int32_t* row_block_prefix, int n_row_blocks, int row_parts,
              int rounds, int identity_k, const int32_t* status, long long n_in_rows,
              long long n_out_rows, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Per-channel normalisation + activation + residual on the [n, c] feature matrix             */
/* (replaces the torch chain nn.BatchNorm1d -> ReLU (-> + identity -> ReLU)  */
/*  of the reference's ConvBlock / BasicBlock: models/mink_unet.py:31-53,104-140,             */
/*  nn/modules/normalizations.py:53-67, nn/modules/activations.py:36-53)                      */
/* All matrices are row-major with a row pitch in elements; dtype as for the GEMMs.           */
/* ------------------------------------------------------------------------------------------ */
/* sums[ch] += sum_r x[r][ch], sums[c + ch] += sum_r x[r][ch]^2 (fp64 [2c], caller zero-fills) */
int wcn_bn_stats(const void* x, long long ld_x, int n, int c, int dtype, double* sums,
                 void* stream);
/* From the sums: mean_rstd[2c] (biased variance, like nn.BatchNorm1d), scale = gamma * rstd,
 * shift = beta - mean * scale, and (optional) running statistics updated with `momentum`
 * (unbiased variance). gamma / beta may be NULL (1 / 0). */
int wcn_bn_finalize(const double* sums, int n, int c, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    float* scale, float* shift, float* mean_rstd, void* stream);
/* y = act(x * scale[ch] + shift[ch] (+ res)), act = ReLU when relu != 0; res optional. Also the
 * eval-mode BatchNorm and the plain bias / affine epilogue. */
int wcn_scale_shift_act(const void* x, long long ld_x, const void* res, long long ld_res, void* y,
                        long long ld_y, int n, int c, int dtype, const float* scale,
                        const float* shift, int relu, void* stream);
/* dz = dy * (y > 0) (y optional: no activation); sums[ch] += sum dz, sums[c+ch] += sum dz*xhat.
 * mask_scale / mask_shift (optional, both or neither): the forward scale / shift of a ReLU layer
 * WITHOUT residual; the mask is then recomputed as x * mask_scale + mask_shift > 0 and y is not
 * read (one matrix read less per backward pass). */
int wcn_bn_bwd_reduce(const void* dy, long long ld_dy, const void* x, long long ld_x,
                      const void* y, long long ld_y, int n, int c, int dtype,
                      const float* mean_rstd, const float* mask_scale, const float* mask_shift,
                      double* sums, void* stream);
/* training != 0: dx = gamma*rstd*(dz - sums[ch]/n - xhat*sums[c+ch]/n); training == 0:
 * dx = dz * gamma[ch] (pass gamma * running rstd). dres (optional) receives dz, the gradient
 * of the residual input. */
int wcn_bn_bwd_apply(const void* dy, long long ld_dy, const void* x, long long ld_x, const void* y,
                     long long ld_y, void* dx, long long ld_dx, void* dres, long long ld_dres,
                     int n, int c, int dtype, const float* gamma, const float* mean_rstd,
                     const double* sums, const float* mask_scale, const float* mask_shift,
                     int training, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Depthwise sparse convolution, weight [K][channels] fp32                  */
/* (replaces nn/functional/sparse_conv_depth.py:227-420, csrc/implicit_fma_kernel.cu,         */
/*  csrc/implicit_reduction.cu)                                                               */
/* ------------------------------------------------------------------------------------------ */
/* out[r][c] = bias[c] + sum_k feats[table[k][r]][c] * weight[kflip ? K-1-k : k][c] for every row
 * r < n_rows (each row written once, no zero-fill needed); table = int32 [K][n_rows] neighbour
 * rows, -1 = none (wcn_kernel_map_search / wcn_csr_to_pair_table). dgrad: pass the reverse table
 * (or the forward table with kflip = 1 for a submanifold map) and the upstream gradient as feats. */
int wcn_depthwise_conv(const void* feats, long long in_ld, void* out, long long out_ld,
                       const float* weight, const float* bias, const int32_t* table, int n_rows,
                       int K, int channels, int dtype, int kflip, int relu, void* stream);
/* Same result on the mask-sorted tile plan of wcn_build_tiles (the plan the tensor-core kernels
 * use): only the offsets active in a tile are visited and their neighbour indices are contiguous.
 * Rows listed in `rows` are written; the preferred entry point when the plan already exists. */
int wcn_depthwise_conv_plan(const void* feats, long long in_ld, void* out, long long out_ld,
                            const float* weight, const float* bias, const int32_t* step_nbr,
                            const int32_t* step_k, const int32_t* rows, const int32_t* tile_nk,
                            int num_tiles, int tile_rows, int K, int channels, int dtype, int kflip,
                            int relu, void* stream);
/* wgrad on the same plan: dw[step_k[i]][c] += sum over tile rows of feats[step_nbr[i][r]][c] *
 * gout[rows[r]][c]; dw fp32 [K][channels], caller zero-fills */
int wcn_depthwise_wgrad_plan(const void* feats, long long in_ld, const void* gout,
                             long long gout_ld, float* dw, const int32_t* step_nbr,
                             const int32_t* step_k, const int32_t* rows, const int32_t* tile_nk,
                             int num_tiles, int tile_rows, int K, int channels, int dtype,
                             void* stream);
/* dw[k][c] += sum_r feats[table[k][r]][c] * gout[r][c]; dw fp32 [K][channels], caller zero-fills */
int wcn_depthwise_wgrad(const void* feats, long long in_ld, const void* gout, long long gout_ld,
                        float* dw, const int32_t* table, int n_rows, int K, int channels,
                        int dtype, void* stream);

/* Training-mode BatchNorm (+ residual) (+ ReLU) of one layer in one call each way: zero-fill of
 * `sums` (fp64 [2c]), statistics, finalize and apply (forward); zero-fill, reduction and apply
 * (backward). scale_shift_mean_rstd is fp32 [4c]: scale, shift, mean, rstd (kept for backward). */
int wcn_bn_forward(const void* x, long long ld_x, const void* res, long long ld_res, void* y,
                   long long ld_y, int n, int c, int dtype, const float* gamma, const float* beta,
                   float eps, float momentum, float* running_mean, float* running_var,
                   double* sums, float* scale_shift_mean_rstd, int relu, void* stream);
int wcn_bn_backward(const void* dy, long long ld_dy, const void* x, long long ld_x, const void* y,
                    long long ld_y, void* dx, long long ld_dx, void* dres, long long ld_dres, int n,
                    int c, int dtype, const float* gamma, const float* mean_rstd,
                    const float* mask_scale, const float* mask_shift, double* sums, void* stream);

/* All-reduce (fp32, in place: every buffer ends as scale * sum over ranks) of the weight gradients over NVLink / NVSwitch peer memory —
 * the path's only collective (the reference has no collective code, users wrap
 * DDP). bufs[r] / flags[r] are HOST arrays of `world` device pointers: the same buffer of `n` floats
 * (n % 4 == 0, 16-byte aligned) and a flag buffer of wcn_peer_allreduce_flag_words() uint32,
 * ZEROED ONCE at allocation, as mapped on this rank for every rank r of the box (CUDA IPC / VMM
 * symmetric memory; torch.distributed._symmetric_memory provides both). Every rank calls it in
 * the same order; n_ctas (<= 128, default 32 when < 1) must be equal on all ranks. Graph-capturable:
 * the barrier epochs live in the flag buffer. A barrier wait gives up after 10 s (a peer that
 * never arrives must not hang the GPU) and counts that in word wcn_peer_allreduce_timeout_word()
 * of this rank's flag buffer: non-zero = the buffer contents are not a valid sum. */
int wcn_peer_allreduce_flag_words(void);
int wcn_peer_allreduce_timeout_word(void);
int wcn_peer_allreduce_f32(void* const* bufs, void* const* flags, int rank, int world,
                           long long n, float scale, int n_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WCN_B200_H_ */
