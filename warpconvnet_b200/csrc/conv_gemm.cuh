// SPDX-License-Identifier: Apache-2.0
// Parameter blocks shared by the sparse-conv GEMM kernels and their launchers.
#pragma once
#include <stdint.h>

namespace wcn {

// Output-stationary fused gather-GEMM (forward AB_gather_scatter and dgrad ABt_gather_scatter).
// One CTA owns a tile of 128 mask-sorted output rows; for every kernel offset that is active in
// the tile it gathers the 128 neighbour rows into 128B-swizzled shared memory, streams the
// offset's weight slice with a bulk copy and accumulates on the tensor cores into TMEM.
struct GatherGemmParams {
  const void* feats;        // [n_in_rows, in_ld] source features (X for fwd, dY for dgrad)
  const void* wimg;         // weight image [n_slabs][K][n_chunks][BN][128 B], see weight_prep.cu
  void* out;                // [n_out_rows, out_ld]
  const int* nbr;           // [K][m_pad] neighbour row per (offset, sorted position), -1 = none
  const int* rows;          // [m_pad] output row of each sorted position, -1 = padding
  const uint16_t* tile_ks;  // [num_tiles][k_stride] active offsets of each tile
  const int* tile_nk;       // [num_tiles] number of active offsets
  const float* bias;        // optional [cout_total] fp32, added in the epilogue
  long long in_ld;          // row strides in elements
  long long out_ld;
  int in_coff;              // first input channel used by slab 0
  int in_slab_stride;       // input-channel step between slabs (group conv), 0 = shared input
  int out_coff;             // first output channel written by slab 0
  int cin;                  // contraction length (channels gathered per row)
  int bn;                   // output channels per slab (multiple of 16, <= 256)
  int K;                    // kernel volume
  int k_stride;             // row pitch of tile_ks
  int m_pad;                // padded sorted length (multiple of 128)
  int num_tiles;
  int kflip;                // weight index = K-1-k (dgrad of a submanifold conv on the fwd table)
  int stages;               // shared-memory pipeline depth
  int relu;                 // fused ReLU in the epilogue (0/1)
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
};

// Weight image builder (weight_prep.cu): see that file for the image layout.
struct WeightPrepParams {
  const void* w;
  void* img;
  long long w_k_stride, w_g_stride, w_r_stride, w_c_stride;  // element strides
  int K, n_slabs, gps;  // gps = groups per slab
  int rg, cg;           // rows / contraction channels per group
  int n_chunks;
  int es;               // element size in bytes
};

}  // namespace wcn
