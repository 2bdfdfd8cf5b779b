// SPDX-License-Identifier: Apache-2.0
// Forward AB_gather_scatter / dgrad ABt_gather_scatter as ONE warp-specialised tcgen05 kernel.
//
// Replaces (semantics, not code): warpconvnet/nn/functional/sparse_conv/detail/explicit.py:22-57
// (Y[out] += X[in] @ W_k over all offsets) and the production kernel family
// warpconvnet/csrc/mask_gemm/include/MaskGemm_forward_*.h (output-stationary, mma.sync).
//
// Work decomposition: output rows are mask-sorted into tiles of `tile_rows` (128 or 256) rows;
// a tile owns a compact STEP LIST = the kernel offsets active anywhere in the tile, each step
// with the `tile_rows` neighbour row indices stored contiguously (see wcn_build_tiles). A CTA
// takes a contiguous range of tiles balanced by step count, so every role streams through
// memory sequentially and index loads are prefetched several steps ahead (no dependent-load
// chain in the steady state).
//
// CTA layout (288 threads, 1 CTA / SM, persistent):
//   warps 0-3  gather producers: cp.async (16 B per lane, zero-fill for missing neighbours)
//              straight into the 128B-swizzled K-major A tile. A stage covers GC = 2 adjacent
//              128-byte channel chunks, so 16 lanes fetch 256 CONTIGUOUS bytes of one feature row:
//              measured on B200 (tools/l2sm_bench.cu, tools/hybrid_bench.cu) the LSU path serves
//              about one gathered row per 8-9 cycles per SM whatever its length up to 256 bytes
//              (16 B/cycle/SM for 128-byte rows, 27-35 for 256-byte rows; sequential = random), TMA
//              tile::gather4 is request-rate bound below that (7.8 B/cycle/SM) and does not add to
//              it, while large cp.async.bulk copies ride on a separate, much faster path. Shared-
//              memory offsets are per-thread constants. With TM = 2 a stage holds two 128-row A
//              sub-tiles that share one weight slice. Thread 0 also pulls the per-offset weight
//              slice (a pre-swizzled image in global memory) with one cp.async.bulk per stage.
//   warp  4    lane 0 issues tcgen05.mma (K = 32 B per instruction; swap-AB form: M = 128 output
//              channels x N = 256 tile rows, else M = 128 rows x N = bn), accumulators in TMEM,
//              double buffered so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 5-8  epilogue: tcgen05.ld -> (+bias, ReLU) -> bf16/fp16/fp32 -> swizzled smem staging ->
//              128-bit global stores, 16 lanes per 256 contiguous bytes of an output row
#include "common.cuh"
#include "conv_gemm.cuh"

namespace wcn {

constexpr int kTileM = 128;
constexpr int kProducerThreads = 128;
constexpr int kMmaWarp = 4;
constexpr int kGemmThreads = 288;
constexpr int kAStageBytes = kTileM * 128;  // 16 KB per 128-row sub-tile
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;             // TMEM column offset of the second accumulator set
constexpr int kMaxStages = 8;
constexpr int kPrefetch = 4;                // steps of index look-ahead in the producers
constexpr int kEpiStageWarp = 32 * 256;     // epilogue staging: 32 rows x 256 B per warp

struct GemmSmemCtrl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  int u_begin;  // this CTA's range of 128-row units (tile_rows / 128 units per tile)
  int u_end;
};

// Work is balanced at 128-row UNIT granularity: a tile of the plan has U = tile_rows / 128 units
// and unit u of tile t starts at cost U * cum[t] + (u - t*U) * nk[t] (cum = tile_cum).
// Returns the first unit index in [0, U*nt] whose start cost is >= target. Warp-cooperative
// 32-ary search (all 32 lanes must call it): 3 rounds of parallel loads cover 32^3 units, where a
// scalar binary search would chain ~2*log2(n) dependent L2 round trips in the kernel prologue.
__device__ __forceinline__ long long unit_cost(const int* __restrict__ cum,
                                               const int* __restrict__ nk, int U, int u) {
  // nk[t] == cum[t+1] - cum[t]: two INDEPENDENT loads of adjacent words instead of a second
  // array (cum has nt + 1 entries and u <= U*nt - 1 here, or u == U*nt with u - t*U == 0)
  const int t = u / U;
  const int r = u - t * U;
  const int c0 = __ldg(cum + t);
  const int c1 = r ? __ldg(cum + t + 1) : c0;
  (void)nk;
  return (long long)U * c0 + (long long)r * (c1 - c0);
}
__device__ __forceinline__ int lower_bound_unit(const int* __restrict__ cum,
                                                const int* __restrict__ nk, int nt, int U,
                                                long long target, int lane) {
  int lo = 0, hi = U * nt;  // answer in [lo, hi]; cost(hi) >= target by construction
  while (hi - lo > 32) {
    const int span = hi - lo;
    // probe points lo + span*(lane+1)/33 (strictly inside (lo, hi))
    const int u = lo + (int)(((long long)span * (lane + 1)) / 33);
    const bool ge = unit_cost(cum, nk, U, u) >= target;
    const unsigned b = __ballot_sync(0xffffffffu, ge);
    // first probe that is >= target bounds the answer from above, the previous one from below
    const int first = b ? __ffs(b) - 1 : 32;
    const int new_hi = first < 32 ? lo + (int)(((long long)span * (first + 1)) / 33) : hi;
    const int new_lo = first > 0 ? lo + (int)(((long long)span * first) / 33) + 1 : lo;
    lo = new_lo;
    hi = new_hi;
  }
  const int u = lo + lane;
  const bool ge = u <= hi && (u == hi || unit_cost(cum, nk, U, u) >= target);
  const unsigned b = __ballot_sync(0xffffffffu, ge);
  return b ? lo + __ffs(b) - 1 : hi;
}

// Sub-tile span [lo, hi) of `tile` that belongs to the unit range [ub, ue).
__device__ __forceinline__ void unit_span(int tile, int U, int ub, int ue, int& lo, int& hi) {
  lo = max(0, ub - tile * U);
  hi = min(U, ue - tile * U);
}

// Walks the (tile, pass, step) sequence of this CTA; all members are warp-uniform. With TM == U
// a tile is one pass over the active sub-tiles [lo, hi); with TM == 1 < U every sub-tile in
// [lo, hi) is its own pass h.
template <int TM>
struct StepCursor {
  int tile, h, h1, i, nk, lo, hi;
  __device__ __forceinline__ void advance_tile(const GatherGemmParams& p, int U, int ub, int ue,
                                               int t_end) {
    for (;;) {
      ++tile;
      if (tile >= t_end) return;
      nk = __ldg(p.tile_nk + tile);
      unit_span(tile, U, ub, ue, lo, hi);
      if (nk > 0 && lo < hi) break;
    }
    h = (TM == 1) ? lo : 0;
    h1 = (TM == 1) ? hi : 1;
    i = 0;
  }
  __device__ __forceinline__ void init(const GatherGemmParams& p, int U, int ub, int ue,
                                       int t_begin, int t_end) {
    tile = t_begin - 1;
    h = h1 = i = nk = lo = hi = 0;
    advance_tile(p, U, ub, ue, t_end);
  }
  __device__ __forceinline__ bool valid(int t_end) const { return tile < t_end; }
  __device__ __forceinline__ void next(const GatherGemmParams& p, int U, int ub, int ue,
                                       int t_end) {
    if (++i < nk) return;
    i = 0;
    if (++h < h1) return;
    advance_tile(p, U, ub, ue, t_end);
  }
};

// SWAP (bn <= 128, TM == 2): the operand roles are exchanged — the weight slice is the M = 128 (Cout)
// operand and the gathered rows are the N operand, so ONE N = 256 instruction covers both
// sub-tiles: D^T[cout, voxel] += W_k^T[cout, cin] * Xg^T[cin, voxel]. A 128x128x16 SS instruction
// reads 8 KB of shared memory for 64 tensor cycles (the shared-memory port, not the tensor pipe,
// paces it: measured ~150 cycles per instruction); the 128x256x16 form reads 12 KB for twice the
// work, i.e. the weight slice is fetched once per step instead of once per sub-tile. Both
// operands are K-major in the same canonical 128B-swizzled layout, so no image changes.
// STATS: the epilogue also accumulates the per-channel sum and sum of squares of the stored output
// (see GatherGemmParams::stats). A template flag, so the plain kernel is unchanged.
template <typename T, int TM, int GC, bool SWAP, bool STATS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gather_gemm_kernel(const __grid_constant__ GatherGemmParams p) {
  pdl_begin();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int slab = blockIdx.y;
  if (WCN_DBG_OUT(p) != nullptr && tid == 0) {  // bring-up: absolute ns at kernel entry
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    WCN_DBG_OUT(p)[blockIdx.x * 16 + 8] = (long long)ns;
  }

  constexpr int kASub = GC * kAStageBytes;  // one 128-row sub-tile: GC slabs of 128 rows x 128 B
  constexpr int kAStage = TM * kASub;
  // byte strides inside the A part of a stage: [sub-tile][chunk][128 rows] normally,
  // [chunk][sub-tile][128 rows] when SWAP (the N = 256 operand needs 256 contiguous rows per chunk)
  constexpr int kChunkStride = SWAP ? TM * kAStageBytes : kAStageBytes;
  constexpr int kSubStride = SWAP ? kAStageBytes : kASub;
  const int b_chunk_bytes = p.bn * 128;     // weight slab of one 128-byte channel chunk
  const int stage_bytes = kAStage + ((GC * b_chunk_bytes + 1023) & ~1023);
  const int stages = p.stages;
  GemmSmemCtrl* ctrl = reinterpret_cast<GemmSmemCtrl*>(smem_gen + (size_t)stages * stage_bytes +
                                                       4 * kEpiStageWarp);

  constexpr int kElem = (int)sizeof(T);
  constexpr int kChunkElems = 128 / kElem;  // channels per 128-byte row segment
  const int n_chunks = (p.cin + kChunkElems - 1) / kChunkElems;
  const int n_groups = (n_chunks + GC - 1) / GC;  // stages per step
  const int row_bytes = p.cin * kElem;            // gathered bytes per feature row
  constexpr int kUnitRows = kTileM * TM;    // rows a CTA processes per step

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
  if (warp < 2) {
    // contiguous range of 128-row units, balanced by step count (warp 0: begin, warp 1: end)
    const int nt = p.num_tiles;
    const int U = p.tile_rows / kTileM;
    const long long S = (long long)U * __ldg(p.tile_cum + nt);
    const int G = gridDim.x, b = blockIdx.x + warp;
    const long long target = S * b / G;
    int u;
    if (p.cta_units != nullptr) {
      u = __ldg(p.cta_units + b);  // computed with the plan (tile_scan_kernel), same rule
    } else {
      u = (b >= G) ? U * nt : lower_bound_unit(p.tile_cum, p.tile_nk, nt, U, target, lane);
      // round to the NEAREST unit boundary (the lower bound alone biases every range by up to a
      // whole unit = ~9 steps: 81 .. 114 unit-steps per CTA instead of 98 +- 5, tools/exp_dbg.py)
      if (b > 0 && b < G && u > 0) {
        const long long above = unit_cost(p.tile_cum, p.tile_nk, U, u) - target;
        const long long below = target - unit_cost(p.tile_cum, p.tile_nk, U, u - 1);
        if (below < above) --u;
      }
    }
    if (lane == 0) {
      if (warp == 0) ctrl->u_begin = u; else ctrl->u_end = u;
    }
  }
  if (warp == kMmaWarp) {
    tmem_alloc(smem_u32(&ctrl->tmem_base), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  const int U = p.tile_rows / kTileM;      // 128-row units per tile
  const int ub = ctrl->u_begin, ue = ctrl->u_end;
  const int t_begin = ub / U;
  const int t_end = (ue + U - 1) / U;

  if (warp < 4) {
    // ===================================== gather producers =====================================
    const uint8_t* wimg = reinterpret_cast<const uint8_t*>(p.wimg) +
                          (size_t)slab * p.K * n_chunks * b_chunk_bytes;
    const long long in_ld_bytes = p.in_ld * kElem;
    // lane -> (row of the instruction, 16-byte unit of the GC*128-byte row segment)
    constexpr int kLanesPerRow = 8 * GC;
    constexpr int kRowsPerInstr = 32 / kLanesPerRow;  // 4 (GC = 1) or 2 (GC = 2)
    constexpr int kInstr = 32 / kRowsPerInstr;        // LDGSTS per thread, sub-tile and stage
    constexpr int kPar = 8 / kRowsPerInstr;           // distinct (row & 7) phases of a thread
    const int u = lane % kLanesPerRow;                // 16-byte unit inside the row segment
    const int sub = lane / kLanesPerRow;              // row of the instruction this lane serves
    const int ch = u >> 3, c8 = u & 7;
    // row(q) = warp*32 + kRowsPerInstr*q + sub;  row & 7 only depends on q % kPar:
    //   off(q) = off_par[q % kPar] + (q / kPar) * 1024
    uint32_t off_par[kPar];
#pragma unroll
    for (int i = 0; i < kPar; ++i) {
      const int r7 = kRowsPerInstr * i + sub;
      off_par[i] = (uint32_t)ch * kChunkStride + (uint32_t)(warp * 32 + r7) * 128u +
                   (uint32_t)((c8 ^ r7) << 4);
    }
    const uint8_t* col_base = reinterpret_cast<const uint8_t*>(p.feats) +
                              (long long)(p.in_coff + slab * p.in_slab_stride) * kElem + u * 16;
    // row pitch in bytes fits 32 bits (checked by the launcher): the source address of a row is
    // ONE 32x32->64 multiply-add (IMAD.WIDE.U32) instead of an emulated 64-bit multiply
    const uint32_t pitch = (uint32_t)in_ld_bytes;

    StepCursor<TM> cur, pre;
    cur.init(p, U, ub, ue, t_begin, t_end);
    pre.init(p, U, ub, ue, t_begin, t_end);
    // ring of prefetched steps, slot 0 = the step issued next (shifted by register moves so the
    // loop body exists once: the role loops of this kernel must stay small in the instruction cache)
    int idx_ring[kPrefetch][TM];  // neighbour row of tile row warp*32 + lane, per sub-tile
    int k_ring[kPrefetch];
    // idx[j] < -1 marks a sub-tile that is not part of this CTA's range (nothing is gathered)
    auto load_step = [&](const StepCursor<TM>& c, int (&idx)[TM], int& k) {
      const size_t step = (size_t)c.tile * p.K + c.i;
      k = __ldg(p.step_k + step);
      const int* base = p.step_nbr + step * p.tile_rows + c.h * kUnitRows + warp * 32 + lane;
#pragma unroll
      for (int j = 0; j < TM; ++j)
        idx[j] = (TM == 1 || (j >= c.lo && j < c.hi)) ? __ldg(base + j * kTileM) : -2;
    };
#pragma unroll
    for (int d = 0; d < kPrefetch; ++d) {
      k_ring[d] = 0;
#pragma unroll
      for (int j = 0; j < TM; ++j) idx_ring[d][j] = -1;
      if (pre.valid(t_end)) {
        load_step(pre, idx_ring[d], k_ring[d]);
        pre.next(p, U, ub, ue, t_end);
      }
    }

    int stage = 0;
    uint32_t phase = 0;
    long long w_empty = 0;
    const long long t_start = WCN_CLOCK();
    unsigned long long ns_start = 0;
#ifdef WCN_BRINGUP
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_start));
#endif
#pragma unroll 1
    while (cur.valid(t_end)) {
      int idx_own[TM];
#pragma unroll
      for (int j = 0; j < TM; ++j) idx_own[j] = idx_ring[0][j];
      const int k = k_ring[0];
#pragma unroll
      for (int d = 0; d + 1 < kPrefetch; ++d) {
        k_ring[d] = k_ring[d + 1];
#pragma unroll
        for (int j = 0; j < TM; ++j) idx_ring[d][j] = idx_ring[d + 1][j];
      }
      if (pre.valid(t_end)) {  // refill the last slot with the step kPrefetch ahead
        load_step(pre, idx_ring[kPrefetch - 1], k_ring[kPrefetch - 1]);
        pre.next(p, U, ub, ue, t_end);
      }
      cur.next(p, U, ub, ue, t_end);
      const int wk = p.kflip ? (p.K - 1 - k) : k;
#pragma unroll 1
      for (int g = 0; g < n_groups; ++g) {
        {
          const long long t0 = WCN_CLOCK();
          mbar_wait(smem_u32(&ctrl->empty[stage]), phase ^ 1u);
          w_empty += WCN_CLOCK() - t0;
        }
        const uint32_t a_smem = smem_base + stage * stage_bytes;
        const uint32_t full_bar = smem_u32(&ctrl->full[stage]);
        if (tid == 0) {
          const int chunks_here = min(GC, n_chunks - g * GC);
          const uint32_t b_bytes = (uint32_t)(chunks_here * b_chunk_bytes);
          if WCN_DBG(p, 8) {  // bring-up: no weight traffic
            mbar_arrive(full_bar);
          } else {
            mbar_arrive_expect_tx(full_bar, b_bytes);
            bulk_copy_g2s(a_smem + kAStage,
                          wimg + ((size_t)wk * n_chunks + g * GC) * b_chunk_bytes, b_bytes,
                          full_bar);
          }
        }
        const int seg_off = g * GC * 128;  // byte offset of this chunk group inside the row
        const bool lane_active = seg_off + u * 16 < row_bytes && !WCN_DBG(p, 2);
        const uint8_t* seg_base = col_base + seg_off;
#pragma unroll
        for (int j = 0; j < TM; ++j) {
          if (TM > 1 && idx_own[j] < -1) continue;  // warp-uniform: sub-tile not ours
#pragma unroll
          for (int q = 0; q < kInstr; ++q) {
            // warp-uniform shuffle; only the copy itself is predicated
            const int src_idx = __shfl_sync(0xffffffffu, idx_own[j], kRowsPerInstr * q + sub);
            const uint8_t* src =
                seg_base + (unsigned long long)(uint32_t)max(src_idx, 0) * pitch;
            const uint32_t dst =
                a_smem + j * kSubStride + off_par[q % kPar] + (uint32_t)(q / kPar) * 1024u;
#ifdef WCN_ZERO_ROWS_SECOND_PASS
            if (lane_active && src_idx >= 0) cp_async_16(dst, src, 16u);
#else
            if (lane_active) cp_async_16(dst, src, src_idx >= 0 ? 16u : 0u);
#endif
          }
        }
#ifdef WCN_ZERO_ROWS_SECOND_PASS
        // EXPERIMENT (not compiled by default, not yet measured; profiles/r1i): a zero-size
        // cp.async still takes a slot of the global-load path, so missing neighbours are written
        // with st.shared in a second pass, after all copies of the stage are in flight. The
        // stores are generic-proxy writes like the cp.async data; they are performed before this
        // thread's arrival can fire, and the consumer's fence.proxy.async covers both.
        {
          bool stored_zero = false;
#pragma unroll
          for (int j = 0; j < TM; ++j) {
            if (TM > 1 && idx_own[j] < -1) continue;
#pragma unroll
            for (int q = 0; q < kInstr; ++q) {
              const int src_idx = __shfl_sync(0xffffffffu, idx_own[j], kRowsPerInstr * q + sub);
              const uint32_t dst =
                  a_smem + j * kSubStride + off_par[q % kPar] + (uint32_t)(q / kPar) * 1024u;
              if (lane_active && src_idx == -1) {
                st_shared_zero_16(dst);
                stored_zero = true;
              }
            }
          }
          if (stored_zero) fence_acq_rel_cta();
        }
#endif
        cp_async_mbar_arrive_noinc(full_bar);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
    if (WCN_DBG_OUT(p) != nullptr && tid == 0) {
      WCN_DBG_OUT(p)[blockIdx.x * 16 + 0] = WCN_CLOCK() - t_start;
      WCN_DBG_OUT(p)[blockIdx.x * 16 + 1] = w_empty;
      unsigned long long ns_end;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_end));
      WCN_DBG_OUT(p)[blockIdx.x * 16 + 7] = (long long)(ns_end - ns_start);
      WCN_DBG_OUT(p)[blockIdx.x * 16 + 9] = (long long)ns_start;
      WCN_DBG_OUT(p)[blockIdx.x * 16 + 10] = (long long)ns_end;
    }
  } else if (warp == kMmaWarp) {
    // ======================================= MMA issuer =========================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(ElemTraits<T>::kFmt, kTileM, p.bn, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t use = 0;  // number of accumulator-set uses so far
      int nk_next = (t_begin < t_end) ? __ldg(p.tile_nk + t_begin) : 0;
      long long w_full = 0, w_acc = 0;
      const long long t_start = WCN_CLOCK();
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int nk = nk_next;
        if (tile + 1 < t_end) nk_next = __ldg(p.tile_nk + tile + 1);
        int lo, hi;
        unit_span(tile, U, ub, ue, lo, hi);
        if (nk == 0 || lo >= hi) continue;
        const int h0 = (TM == 1) ? lo : 0, h1 = (TM == 1) ? hi : 1;
        for (int h = h0; h < h1; ++h) {
          const uint32_t acc = use & 1u;
          {
            const long long t0 = WCN_CLOCK();
            mbar_wait(smem_u32(&ctrl->acc_empty[acc]), ((use >> 1) & 1u) ^ 1u);
            w_acc += WCN_CLOCK() - t0;
          }
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kAccStride;
          uint32_t accumulate = 0;
          for (int ki = 0; ki < nk; ++ki) {
            for (int g = 0; g < n_groups; ++g) {
              {
                const long long t0 = WCN_CLOCK();
                mbar_wait(smem_u32(&ctrl->full[stage]), phase);
                w_full += WCN_CLOCK() - t0;
              }
              fence_proxy_async_smem();  // cp.async wrote A through the generic proxy
              tc_fence_after();
              const uint32_t a_smem = smem_base + stage * stage_bytes;
              const uint32_t b_smem = a_smem + kAStage;
#pragma unroll
              for (int cc = 0; cc < GC; ++cc) {
                const int chunk_bytes = min(128, row_bytes - (g * GC + cc) * 128);
                const int n_mma = chunk_bytes > 0 ? (chunk_bytes >> 5) : 0;  // 32 B of K per MMA
                if constexpr (SWAP) {
                  // M = Cout (weight slice), N = the CTA's share of the tile's 256 rows
                  const uint32_t idesc_sw =
                      make_idesc(ElemTraits<T>::kFmt, kTileM, (hi - lo) * kTileM, 0, 0);
                  for (int m = 0; m < n_mma; ++m) {
                    if WCN_DBG(p, 4) continue;  // bring-up: no tensor work
                    const uint64_t wdesc =
                        make_smem_desc_sw128(b_smem + cc * b_chunk_bytes + m * 32, 16, 1024);
                    const uint64_t xdesc = make_smem_desc_sw128(
                        a_smem + cc * kChunkStride + lo * kAStageBytes + m * 32, 16, 1024);
                    umma_ss<ElemTraits<T>::kTF32>(tmem_d + lo * kTileM, wdesc, xdesc, idesc_sw,
                                                  (accumulate | (uint32_t)(m + cc)) ? 1u : 0u);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < TM; ++j) {
                    if (TM > 1 && (j < lo || j >= hi)) continue;  // sub-tile not ours
                    for (int m = 0; m < n_mma; ++m) {
                      if WCN_DBG(p, 4) continue;  // bring-up: no tensor work
                      const uint64_t adesc = make_smem_desc_sw128(
                          a_smem + j * kASub + cc * kAStageBytes + m * 32, 16, 1024);
                      const uint64_t bdesc =
                          make_smem_desc_sw128(b_smem + cc * b_chunk_bytes + m * 32, 16, 1024);
                      umma_ss<ElemTraits<T>::kTF32>(tmem_d + j * p.bn, adesc, bdesc, idesc,
                                                    (accumulate | (uint32_t)(m + cc)) ? 1u : 0u);
                    }
                  }
                }
              }
              accumulate = 1;
              umma_commit(smem_u32(&ctrl->empty[stage]));
              if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
          }
          umma_commit(smem_u32(&ctrl->acc_full[acc]));
          ++use;
        }
      }
      if (WCN_DBG_OUT(p) != nullptr) {
        WCN_DBG_OUT(p)[blockIdx.x * 16 + 2] = WCN_CLOCK() - t_start;
        WCN_DBG_OUT(p)[blockIdx.x * 16 + 3] = w_full;
        WCN_DBG_OUT(p)[blockIdx.x * 16 + 4] = w_acc;
      }
    }
  } else {
    // ======================================== epilogue ==========================================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    uint8_t* out = reinterpret_cast<uint8_t*>(p.out);
    const long long out_ld_bytes = p.out_ld * kElem;
    const int col_base = p.out_coff + slab * p.bn;
    constexpr int kChunkCols = 256 / kElem;  // output columns per 256-byte staging chunk
    const uint32_t stage_warp = smem_base + stages * stage_bytes + (uint32_t)q * kEpiStageWarp;
    const uint32_t stage_row = stage_warp + lane * 256;
    uint32_t use = 0;
    long long w_accf = 0;
    const long long t_start = WCN_CLOCK();
    // STATS accumulators. SWAP: this thread owns ONE output channel (q*32 + lane). Row-major:
    // the store loop hands this thread 16-byte piece (lane & 15) of every second staged row, i.e.
    // kPieceElems fixed columns per 256-byte column chunk.
    constexpr int kPieceElems = 16 / kElem;
    constexpr int kStatChunks = SWAP ? 1 : (256 * kElem) / 256;  // column chunks of a 256-col slab
    float st_s[kStatChunks][SWAP ? 1 : kPieceElems], st_q[kStatChunks][SWAP ? 1 : kPieceElems];
#pragma unroll
    for (int a = 0; a < kStatChunks; ++a)
#pragma unroll
      for (int b = 0; b < (SWAP ? 1 : kPieceElems); ++b) st_s[a][b] = st_q[a][b] = 0.f;
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int nk = __ldg(p.tile_nk + tile);
      int lo, hi;
      unit_span(tile, U, ub, ue, lo, hi);
      if (lo >= hi) continue;
      const int h0 = (TM == 1) ? lo : 0, h1 = (TM == 1) ? hi : 1;
      for (int h = h0; h < h1; ++h) {
        uint32_t acc = 0;
        if (nk > 0) {
          acc = use & 1u;
          const long long t0 = WCN_CLOCK();
          mbar_wait(smem_u32(&ctrl->acc_full[acc]), (use >> 1) & 1u);
          w_accf += WCN_CLOCK() - t0;
          tc_fence_after();
        }
        if constexpr (SWAP) {
          // TMEM lane = output channel, TMEM column = tile row: every thread owns one channel of
          // its warp's 32. Batches of 32 rows are transposed through a warp-private staging block
          // ([row][32 channels]) and leave as 16-byte pieces, 32 * sizeof(T) contiguous bytes of
          // an output row per warp.
          constexpr int kRowB = 32 * kElem;           // staging row: this warp's 32 channels
          constexpr int kPieces = kRowB / 16;         // 4 (16-bit) or 8 (fp32)
          constexpr int kRowsPerSt = 32 / kPieces;    // rows per store instruction
          // bn < 128: the M = 128 instruction read rows bn..127 of the weight operand from
          // whatever follows the slice in shared memory; those accumulator lanes are garbage
          // and are never stored (warps / pieces past bn are skipped)
          const bool warp_valid = q * 32 < p.bn;
          const float bias_c = (p.bias != nullptr && q * 32 + lane < p.bn)
                                   ? __ldg(p.bias + col_base + q * 32 + lane) : 0.f;
          const int s_row = lane / kPieces, s_piece = lane % kPieces;
          const bool piece_valid = q * 32 + (s_piece + 1) * (16 / kElem) <= p.bn;
          for (int j = lo; warp_valid && j < hi; ++j) {
            for (int vb = 0; vb < kTileM / 32; ++vb) {
              const int out_row =
                  __ldg(p.rows + (size_t)tile * p.tile_rows + j * kTileM + vb * 32 + lane);
              uint32_t v[32];
              if (nk > 0) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride +
                                       j * kTileM + vb * 32;
                tmem_ld_x16(taddr, v);
                tmem_ld_x16(taddr + 16, v + 16);
                tmem_ld_wait();
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0u;
              }
              unsigned row_ok = 0u;
              if constexpr (STATS) row_ok = __ballot_sync(0xffffffffu, out_row >= 0);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float f = __uint_as_float(v[i]) + bias_c;
                if (p.relu) f = fmaxf(f, 0.f);
                st_shared_elem<T>(stage_warp + i * kRowB + lane * kElem, f);
                if constexpr (STATS) {
                  const float fr = round_to<T>(f);  // the value as stored
                  if ((row_ok >> i) & 1u) {
                    st_s[0][0] += fr;
                    st_q[0][0] = fmaf(fr, fr, st_q[0][0]);
                  }
                }
              }
              __syncwarp();
#pragma unroll
              for (int it = 0; it < kPieces; ++it) {
                const int row = it * kRowsPerSt + s_row;
                const int orow = __shfl_sync(0xffffffffu, out_row, row);
                const uint4 val = ld_shared_v4(stage_warp + row * kRowB + s_piece * 16);
                if (orow >= 0 && piece_valid && !WCN_DBG(p, 1))
                  *reinterpret_cast<uint4*>(out + (long long)orow * out_ld_bytes +
                                            (long long)(col_base + q * 32) * kElem +
                                            s_piece * 16) = val;
              }
              __syncwarp();
            }
          }
        } else {
#pragma unroll
        for (int j = 0; j < TM; ++j) {
          if (TM > 1 && (j < lo || j >= hi)) continue;  // sub-tile not ours
          const int out_row =
              __ldg(p.rows + (size_t)tile * p.tile_rows + h * kUnitRows + j * kTileM + r);
          // 256-byte column chunks: TMEM -> registers -> (bias, ReLU, convert) -> XOR-swizzled
          // staging rows in shared memory -> 16 lanes write one output row's 256 contiguous
          // bytes (two rows per store instruction instead of 32 scattered 16-byte pieces)
          for (int c0 = 0, ci = 0; c0 < p.bn; c0 += kChunkCols, ++ci) {
            const int c1 = min(p.bn, c0 + kChunkCols);
            for (int col = c0; col < c1; col += 16) {
              uint32_t v[16];
              if (nk > 0) {
                tmem_ld_x16(
                    tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride + j * p.bn + col, v);
                tmem_ld_wait();
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0u;
              }
              if (p.bias != nullptr) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  v[i] = __float_as_uint(__uint_as_float(v[i]) +
                                         __ldg(p.bias + col_base + col + i));
              }
              if (p.relu) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  v[i] = __float_as_uint(fmaxf(__uint_as_float(v[i]), 0.f));
              }
              const int piece0 = (col - c0) * kElem / 16;  // first 16-byte piece of these columns
              if constexpr (sizeof(T) == 2) {
                st_shared_v4(stage_row + (((piece0 + 0) ^ (lane & 15)) << 4), pack8<T>(v));
                st_shared_v4(stage_row + (((piece0 + 1) ^ (lane & 15)) << 4), pack8<T>(v + 8));
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  st_shared_v4(stage_row + (((piece0 + i) ^ (lane & 15)) << 4),
                               make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
              }
            }
            __syncwarp();
            const int piece = lane & 15;
            const bool piece_ok = c0 * kElem + piece * 16 < p.bn * kElem;
            // Branch-free store loop: the shared-memory read of an iteration never sits behind a
            // branch, so the unrolled iterations overlap their shuffle / LDS / STG latencies. The
            // statistics of this column chunk are accumulated from a zero-selected piece (rows
            // past the end of the plan and pieces past bn contribute +0) and folded into the
            // per-chunk registers once per chunk. (With a branch per row the STATS epilogue cost
            // ~10 k cycles per 256-row tile against ~5 k without, profiles/r2v_epilogue.md.)
            float cs[STATS ? kPieceElems : 1], cq[STATS ? kPieceElems : 1];
            if constexpr (STATS) {
#pragma unroll
              for (int e = 0; e < kPieceElems; ++e) cs[e] = cq[e] = 0.f;
            }
#pragma unroll 4
            for (int rr = 0; rr < 16; ++rr) {
              const int row = 2 * rr + (lane >> 4);
              const int orow = __shfl_sync(0xffffffffu, out_row, row);
              const uint4 val =
                  ld_shared_v4(stage_warp + row * 256 + ((piece ^ (row & 15)) << 4));
              const bool ok = orow >= 0 && piece_ok;
              if (ok && !WCN_DBG(p, 1))  // debug 1: skip stores (bring-up)
                *reinterpret_cast<uint4*>(out + (long long)orow * out_ld_bytes +
                                          (long long)(col_base + c0) * kElem + piece * 16) = val;
              if constexpr (STATS && !SWAP) {
                float fe[kPieceElems];
                unpack_piece<T>(ok ? val : make_uint4(0u, 0u, 0u, 0u), fe);  // +0.0 in every type
#pragma unroll
                for (int e = 0; e < kPieceElems; ++e) {
                  cs[e] += fe[e];
                  cq[e] = fmaf(fe[e], fe[e], cq[e]);
                }
              }
            }
            if constexpr (STATS && !SWAP) {
#pragma unroll
              for (int cc = 0; cc < kStatChunks; ++cc) {  // static register indexing
                if (cc == ci) {
#pragma unroll
                  for (int e = 0; e < kPieceElems; ++e) {
                    st_s[cc][e] += cs[e];
                    st_q[cc][e] += cq[e];
                  }
                }
              }
            }
            __syncwarp();
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
    if constexpr (STATS) {
      if constexpr (SWAP) {
        const int ch = q * 32 + lane;
        if (ch < p.bn) {
          atomicAdd(p.stats + col_base + ch, (double)st_s[0][0]);
          atomicAdd(p.stats + p.stats_c + col_base + ch, (double)st_q[0][0]);
        }
      } else {
        // merge the four epilogue warps (and the two half-warps that share a piece) in shared
        // memory — the staging blocks are free now — then one fp64 atomic per column and CTA
        float* sh = reinterpret_cast<float*>(smem_gen + (size_t)stages * stage_bytes);  // [2][256]
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = tid - 5 * 32; i < 2 * 256; i += 128) sh[i] = 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int piece = lane & 15;
#pragma unroll
        for (int ci = 0; ci < (256 * kElem) / 256; ++ci) {
#pragma unroll
          for (int e = 0; e < kPieceElems; ++e) {
            const int col = ci * kChunkCols + piece * kPieceElems + e;
            if (col < p.bn) {
              atomicAdd(sh + col, st_s[ci][e]);
              atomicAdd(sh + 256 + col, st_q[ci][e]);
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = tid - 5 * 32; i < p.bn; i += 128) {
          atomicAdd(p.stats + col_base + i, (double)sh[i]);
          atomicAdd(p.stats + p.stats_c + col_base + i, (double)sh[256 + i]);
        }
      }
    }
    if (WCN_DBG_OUT(p) != nullptr && tid == 5 * 32) {
      WCN_DBG_OUT(p)[blockIdx.x * 16 + 5] = WCN_CLOCK() - t_start;
      WCN_DBG_OUT(p)[blockIdx.x * 16 + 6] = w_accf;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (WCN_DBG_OUT(p) != nullptr && tid == 0) {  // bring-up: absolute ns when every role is done
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    WCN_DBG_OUT(p)[blockIdx.x * 16 + 11] = (long long)ns;
  }
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

static int stage_bytes_of(int bn, int tm, int gc) {
  return tm * gc * kAStageBytes + ((gc * bn * 128 + 1023) & ~1023);
}

static size_t gemm_smem_bytes(int bn, int tm, int gc, int stages) {
  return (size_t)stages * stage_bytes_of(bn, tm, gc) + 4 * kEpiStageWarp + sizeof(GemmSmemCtrl) +
         1024;
}

static int pick_gemm_stages(int bn, int tm, int gc) {
  int s = (int)((227 * 1024 - 4 * kEpiStageWarp - sizeof(GemmSmemCtrl) - 1024) /
                stage_bytes_of(bn, tm, gc));
  if (s > kMaxStages) s = kMaxStages;
  return s;
}

template <typename T, int TM, int GC, bool SWAP, bool STATS>
static int launch_gather_gemm_ts(GatherGemmParams p, int n_slabs, int max_ctas, int n_range_ctas,
                                 cudaStream_t stream) {
  const int max_stages = pick_gemm_stages(p.bn, TM, GC);
  if (p.stages <= 0 || p.stages > max_stages) p.stages = max_stages;
  if (p.stages < 2) return kErrUnsupportedShape;
  const size_t smem = gemm_smem_bytes(p.bn, TM, GC, p.stages);
  static int configured[kMaxDevices] = {};  // per instantiation and device
  int& configured_smem = configured[current_device_slot()];
  if ((int)smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(gather_gemm_kernel<T, TM, GC, SWAP, STATS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return kErrCuda;
    configured_smem = (int)smem;
  }
  const int units = p.num_tiles * (p.tile_rows / kTileM);
  int ctas = units < max_ctas ? units : max_ctas;
  if (ctas < 1) return kOk;
  if (p.cta_units != nullptr && ctas != n_range_ctas) p.cta_units = nullptr;  // other grid: search
  dim3 grid(ctas, n_slabs, 1);
  wcn_launch(gather_gemm_kernel<T, TM, GC, SWAP, STATS>, dim3(grid), dim3(kGemmThreads), smem, stream, p);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
}

template <typename T, int TM, int GC, bool SWAP>
static int launch_gather_gemm_t(const GatherGemmParams& p, int n_slabs, int max_ctas,
                                int n_range_ctas, cudaStream_t stream) {
  return p.stats != nullptr
             ? launch_gather_gemm_ts<T, TM, GC, SWAP, true>(p, n_slabs, max_ctas, n_range_ctas, stream)
             : launch_gather_gemm_ts<T, TM, GC, SWAP, false>(p, n_slabs, max_ctas, n_range_ctas, stream);
}

template <typename T>
static int launch_gather_gemm_tm(const GatherGemmParams& p, int n_slabs, int max_ctas,
                                 int n_range_ctas, cudaStream_t stream) {
  // TM = 2: two 128-row sub-tiles share each weight slice when the plan has 256-row tiles and
  // both accumulators (double buffered) fit the 512 TMEM columns.
  // GC = 2: a stage covers two 128-byte channel chunks so gathers are 256-byte requests.
  const bool tm2 = p.tile_rows == 2 * kTileM && p.bn <= 128 && !WCN_DBG(p, 16);  // 16: bring-up
  const bool gc2 = p.cin * (int)sizeof(T) > 128 && !WCN_DBG(p, 32);             // 32: bring-up
  // SWAP: Cout is exactly one M = 128 operand and the tile's 256 rows form the N operand
  // (bn < 128 pads M to 128 with don't-care rows: correct for every bn, but measured slower than
  // the row-major form at bn <= 96 on the MinkUNet-14 layers — 422 vs 402 us at 96 channels, 140 vs
  // 117 us at 32/64 — so it is only taken where the padding is small)
  const bool swap = tm2 && p.bn > 96 && p.bn <= kTileM && !WCN_DBG(p, 64);       // 64: bring-up
  if (swap) {
    return gc2 ? launch_gather_gemm_t<T, 2, 2, true>(p, n_slabs, max_ctas, n_range_ctas, stream)
               : launch_gather_gemm_t<T, 2, 1, true>(p, n_slabs, max_ctas, n_range_ctas, stream);
  }
  if (tm2) {
    return gc2 ? launch_gather_gemm_t<T, 2, 2, false>(p, n_slabs, max_ctas, n_range_ctas, stream)
               : launch_gather_gemm_t<T, 2, 1, false>(p, n_slabs, max_ctas, n_range_ctas, stream);
  }
  return gc2 ? launch_gather_gemm_t<T, 1, 2, false>(p, n_slabs, max_ctas, n_range_ctas, stream)
             : launch_gather_gemm_t<T, 1, 1, false>(p, n_slabs, max_ctas, n_range_ctas, stream);
}

int launch_gather_gemm(const GatherGemmParams& p, int dtype, int n_slabs, int max_ctas,
                       int n_range_ctas, cudaStream_t stream) {
  const int es = dtype_size(dtype);
  if (p.bn < 16 || p.bn > 256 || (p.bn % 16) != 0) return kErrUnsupportedShape;
  if (p.cin <= 0 || (p.cin * es) % 32 != 0) return kErrUnsupportedShape;
  if (p.in_ld * es >= (1ll << 31)) return kErrUnsupportedShape;  // 32-bit row pitch in the gather
  if ((p.in_ld * es) % 16 != 0 || (p.in_coff * es) % 16 != 0 || (p.in_slab_stride * es) % 16 != 0)
    return kErrAlignment;
  if ((p.out_ld * es) % 16 != 0 || (p.out_coff * es) % 16 != 0) return kErrAlignment;
  if ((reinterpret_cast<uintptr_t>(p.feats) & 15) || (reinterpret_cast<uintptr_t>(p.out) & 15) ||
      (reinterpret_cast<uintptr_t>(p.wimg) & 15))
    return kErrAlignment;
  if (p.tile_rows != kTileM && p.tile_rows != 2 * kTileM) return kErrInvalidArg;
  if (p.m_pad % p.tile_rows != 0 || (long long)p.num_tiles * p.tile_rows > p.m_pad)
    return kErrInvalidArg;
  switch (dtype) {
    case kBF16: return launch_gather_gemm_tm<__nv_bfloat16>(p, n_slabs, max_ctas, n_range_ctas, stream);
    case kF16: return launch_gather_gemm_tm<__half>(p, n_slabs, max_ctas, n_range_ctas, stream);
    case kF32: return launch_gather_gemm_tm<float>(p, n_slabs, max_ctas, n_range_ctas, stream);
    default: return kErrUnsupportedDtype;
  }
}

}  // namespace wcn
