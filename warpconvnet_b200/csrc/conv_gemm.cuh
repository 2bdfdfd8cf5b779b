// SPDX-License-Identifier: Apache-2.0
// Parameter blocks shared by the sparse-conv GEMM kernels and their launchers.
#pragma once
#include <stdint.h>

namespace wcn {

// Output-stationary fused gather-GEMM (forward AB_gather_scatter and dgrad ABt_gather_scatter).
// One CTA owns a contiguous range of mask-sorted tiles; for every step (= kernel offset active
// in the tile) it gathers the tile's neighbour rows into 128B-swizzled shared memory, streams the
// offset's weight slice with a bulk copy and accumulates on the tensor cores into TMEM.
struct GatherGemmParams {
  const void* feats;        // [n_in_rows, in_ld] source features (X for fwd, dY for dgrad)
  const void* wimg;         // weight image [n_slabs][K][n_chunks][BN][128 B], see weight_prep.cu
  void* out;                // [n_out_rows, out_ld]
  const int* step_nbr;      // [num_tiles][K][tile_rows] neighbour rows of step i of tile t at
                            // (t*K + i)*tile_rows, -1 = none; only the first tile_nk[t] steps exist
  const int* step_k;        // [num_tiles][K] kernel offset of each step
  const int* rows;          // [m_pad] output row of each sorted position, -1 = padding
  const int* tile_nk;       // [num_tiles] number of steps (active offsets) of each tile
  const int* tile_cum;      // [num_tiles + 1] exclusive prefix sum of tile_nk
  const int* cta_units;     // optional [gridDim.x + 1] precomputed unit range of every CTA
  const float* bias;        // optional [cout_total] fp32, added in the epilogue
  // optional per-channel statistics of the OUTPUT as stored (after bias / ReLU / rounding to the
  // feature dtype): stats[ch] += sum_r y[r, ch], stats[stats_c + ch] += sum_r y[r, ch]^2, fp64,
  // caller zero-fills. What the BatchNorm that follows a conv needs (csrc/rownorm.cu bn_stats),
  // accumulated in the epilogue while the tile is in registers: saves one read pass over Y.
  double* stats;
  int stats_c;              // channel count of the statistics buffer (= total output channels)
  long long in_ld;          // row strides in elements
  long long out_ld;
  int n_in_rows;            // rows of feats (bounds documentation; neighbours are < n_in_rows)
  int in_coff;              // first input channel used by slab 0
  int in_slab_stride;       // input-channel step between slabs (group conv), 0 = shared input
  int out_coff;             // first output channel written by slab 0
  int cin;                  // contraction length (channels gathered per row)
  int bn;                   // output channels per slab (multiple of 16, <= 256)
  int K;                    // kernel volume
  int tile_rows;            // rows per tile of the plan: 128 or 256
  int m_pad;                // padded sorted length (multiple of tile_rows)
  int num_tiles;
  int kflip;                // weight index = K-1-k (dgrad of a submanifold conv on the fwd table)
  int stages;               // shared-memory pipeline depth
  int relu;                 // fused ReLU in the epilogue (0/1)
  long long* dbg_out;       // bring-up only: per-CTA wait-cycle counters (env WCN_DEBUG_PTR)
  int debug;                // bring-up only (env WCN_DEBUG): 16 force TM = 1
};

// wgrad AtB_gather_gather: per offset k, dW_k[cin, cout] += X[in_maps]^T * dY[out_maps].
// Work units are fixed-size slices of one offset's pair list; the fp32 result of a unit is
// reduced into dW with vectorised red.global. grid = (ctas, y slabs, z slabs):
//   y = 128-channel slab of Cin (dense conv) or a slab of `gps` groups (group conv, densified
//       block-diagonally: only the diagonal cin_g x cout_g blocks are written back)
//   z = slab of Cout (dense conv only)
struct WgradParams {
  const void* feats;   // X  [n_in_rows, in_ld]
  const void* gout;    // dY [n_out_rows, out_ld]
  float* dw;           // fp32 accumulation target
  const int* in_maps;  // [L]
  const int* out_maps; // [L]
  const int* offsets;  // [K+1] device copy of the CSR offsets
  // optional row-block-major unit order (L2 locality): blk_prefix[k][b] = pairs of offset k whose
  // output row is < 256*b (the kernel-map block scan, [K][n_row_blocks]); 0 / nullptr = off
  const int* blk_prefix;
  int n_row_blocks;
  int row_parts;       // P: row blocks of the virtual order
  int rounds;          // R: chunks per CTA (round-robin over the virtual list)
  // identity offset of a submanifold map (in_maps == out_maps == 0..n-1 for offset identity_k):
  // its rows are contiguous, so full stages are fetched by the TMA unit as 2-D tiles instead of
  // 2 x 64 row gathers. -1 = off. `status` = the hash table's status word (bit 2: duplicate
  // coordinates were seen, the identity property does not hold then).
  int identity_k;
  const int* status;
  long long in_ld;
  long long out_ld;
  long long dw_k_stride;  // elements between offsets
  long long dw_g_stride;  // elements between groups (group conv)
  long long dw_ld;        // elements between consecutive input channels
  long long dw_y_stride;  // dW element offset per y slab
  long long dw_z_stride;  // dW element offset per z slab
  int in_coff, in_y_stride;                 // X channel window: in_coff + y*in_y_stride
  int out_coff, out_y_stride, out_z_stride; // dY channel window
  int cin;             // M extent of a y slab (<= 128)
  int cin_last;        // M extent of the last y slab
  int cout;            // N extent of a slab (<= 256, multiple of 16)
  int gps;             // groups per slab (1 = dense)
  int cin_g, cout_g;   // per-group channels when gps > 1
  int K;
  int unit_pairs;      // pairs per work unit (multiple of the stage depth)
  int stages;
  float alpha;
  int debug;           // bring-up only (env WCN_DEBUG): 64 skip proxy fence, 128 128-pair stages
  long long* dbg_out;  // bring-up only: per-CTA cycle counters (env WCN_DEBUG_PTR)
};

// Weight image builder (weight_prep.cu): see that file for the image layout.
struct WeightPrepParams {
  const void* w;
  void* img;
  long long w_k_stride, w_g_stride, w_r_stride, w_c_stride;  // element strides
  int K, n_slabs, gps;  // gps = groups per slab
  int rg, cg;           // rows / contraction channels per group
  int n_chunks;
  int es;               // element size of the IMAGE in bytes
  int src_es;           // element size of the source weights (== es, or 4 with a 16-bit image)
  int cvt;              // 0: copy bits, 1: fp32 -> bf16, 2: fp32 -> fp16 (round to nearest even)
};

}  // namespace wcn
