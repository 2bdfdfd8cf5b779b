// SPDX-License-Identifier: Apache-2.0
// wgrad AtB_gather_gather on tcgen05: dW_k[ci, co] += sum over pairs (i,o) of offset k of
// X[i, ci] * dY[o, co].
//
// Replaces (semantics, not code): warpconvnet/nn/functional/sparse_conv/detail/explicit.py:95-97
// and the production kernels csrc/mask_gemm/include/MaskGemm_wgrad_*.h (32-deep mma.sync
// contraction tiles + fp32 atomics) / detail/cute_grouped.py:282-432.
//
// The contraction runs over the pair list, so both operands are gathered rows that land in shared
// memory exactly as they sit in HBM (channels contiguous) = "MN-major" operands for the tensor
// core: A = X rows (M = cin), B = dY rows (N = cout), K = pairs. The CSR pair lists are exact, so
// no padded work is issued. A work unit is a fixed-size slice of one offset's pair list; its fp32
// 128 x cout result is reduced into dW with vectorised red.global.add.v4.f32.
//
// CTA layout identical to the forward kernel: warps 0-3 gather, warp 4 issues MMAs, warps 5-8
// drain TMEM (double buffered).
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "conv_gemm.cuh"

namespace wcn {

constexpr int kWgThreads = 288;
constexpr int kWgMaxStages = 8;
constexpr int kWgTmemCols = 512;
constexpr int kWgAccStride = 256;
constexpr int kWgMaxK = 65535;
constexpr int kWgMaxUnits = 1024;  // (row parts) x K units of the row-block-major order
constexpr int kWgSegTableBytes = (2 * kWgMaxUnits + 2) * 4;

struct WgSmemCtrl {
  uint64_t full[kWgMaxStages];
  uint64_t empty[kWgMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  int pair_begin;  // this CTA's slice of the concatenated pair lists
  int pair_end;
  int k_first;     // offset that contains pair_begin
};

// Walks this CTA's work segments (offset k, first pair, pair count); every role runs it
// identically (all state is warp-uniform).
//  * default order: the CTA owns one contiguous slice of the concatenated (offset-major) pair
//    lists; a segment is slice ∩ offset.
//  * row-block-major order (tab != nullptr): the pair lists are re-cut into units (row part p,
//    offset k) = pairs of offset k whose OUTPUT row lies in part p, laid out p-major as one virtual
//    list (vstart = prefix sums, ustart = start of each unit in the real lists). The virtual list is
//    split into G*R equal chunks and CTA b takes chunks b, b+G, b+2G, ...: at any time all CTAs
//    work inside a window of 1/R of the rows, which sees all K offsets while its X / dY rows are
//    L2-resident (the offset-major order swept the whole 102 MB working set 27 times: 343 MB of
//    DRAM reads for 126 MB of algorithmic bytes on C3).
struct WgSegCursor {
  const int* offsets;
  const int* vstart;  // shared memory [U + 1] or nullptr
  const int* ustart;  // shared memory [U]
  int K, U, G, R, b;
  int pair_begin, pair_end, k;  // default order
  int r, u;                     // row-block-major order
  long long v0, v1;

  __device__ __forceinline__ bool next(int& k_out, int& first, int& count) {
    if (vstart == nullptr) {
      for (; k < K; ++k) {
        const int ob = __ldg(offsets + k), oe = __ldg(offsets + k + 1);
        if (ob >= pair_end) return false;
        first = max(ob, pair_begin);
        count = min(oe, pair_end) - first;
        if (count > 0) {
          k_out = k++;
          return true;
        }
      }
      return false;
    }
    for (;;) {
      if (r >= R) return false;
      if (u < 0) {  // open chunk r
        const long long Lv = vstart[U];
        const long long j = (long long)b + (long long)r * G;
        v0 = Lv * j / ((long long)G * R);
        v1 = Lv * (j + 1) / ((long long)G * R);
        if (v0 >= v1) { ++r; continue; }
        int lo = 0, hi = U;  // largest u with vstart[u] <= v0
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (vstart[mid] <= v0) lo = mid; else hi = mid;
        }
        u = lo;
      }
      if (u >= U || vstart[u] >= v1) { ++r; u = -1; continue; }
      const long long s = max(v0, (long long)vstart[u]);
      const long long e = min(v1, (long long)vstart[u + 1]);
      const int cu = u++;
      if (e <= s) continue;
      k_out = cu % K;
      first = ustart[cu] + (int)(s - vstart[cu]);
      count = (int)(e - s);
      return true;
    }
  }
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

template <typename T, int PAIRS, int NSEGB>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ WgradParams p, const __grid_constant__ CUtensorMap tmap_x,
             const __grid_constant__ CUtensorMap tmap_dy) {
  pdl_begin();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  constexpr int kElem = (int)sizeof(T);
  constexpr int kBlkElems = 128 / kElem;          // channels per 128-byte MN block
  constexpr int kPairs = PAIRS;                   // pairs (K extent) per pipeline stage
  constexpr int kKPerMma = 32 / kElem;            // 16 (bf16/f16) or 8 (tf32) pairs per MMA
  constexpr int kAStage = (128 / kBlkElems) * kPairs * 128;  // A always spans M = 128 channels
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int ys = blockIdx.y;  // Cin slab / group slab
  const int zs = blockIdx.z;  // Cout slab
  const int cin_here = (ys == (int)gridDim.y - 1) ? p.cin_last : p.cin;
  const int n_blk_a = (cin_here + kBlkElems - 1) / kBlkElems;
  const int n_blk_b = (p.cout + kBlkElems - 1) / kBlkElems;
  const int b_stage = 2 * NSEGB * kPairs * 128;  // whole 256-byte segments (zero-filled tails)
  const int stage_bytes = kAStage + ((b_stage + 1023) & ~1023);
  const int stages = p.stages;
  WgSmemCtrl* ctrl = reinterpret_cast<WgSmemCtrl*>(smem_gen + (size_t)stages * stage_bytes);
  int* seg_vstart = reinterpret_cast<int*>(ctrl + 1);  // [U + 1]
  int* seg_ustart = seg_vstart + kWgMaxUnits + 1;       // [U]
  const bool blocked = p.blk_prefix != nullptr;
  const int n_units = blocked ? p.row_parts * p.K : 0;
  if (blocked) {
    // unit (part, k): pairs of offset k with output row in [256*blk(part), 256*blk(part+1))
    const int nb = p.n_row_blocks, P = p.row_parts;
    for (int un = tid; un < n_units; un += kWgThreads) {
      const int part = un / p.K, k = un - part * p.K;
      const int ob = __ldg(p.offsets + k);
      const int b0 = (int)((long long)part * nb / P), b1 = (int)((long long)(part + 1) * nb / P);
      const int lo = ob + (part == 0 ? 0 : __ldg(p.blk_prefix + (size_t)k * nb + b0));
      const int hi = (part == P - 1) ? __ldg(p.offsets + k + 1)
                                     : ob + __ldg(p.blk_prefix + (size_t)k * nb + b1);
      seg_ustart[un] = lo;
      seg_vstart[un + 1] = hi - lo;  // length; turned into a prefix sum below
    }
    __syncthreads();
    if (warp == 0) {
      const int per = (n_units + 31) / 32;
      const int base = lane * per;
      int sum = 0;
      for (int i = 0; i < per; ++i)
        if (base + i < n_units) sum += seg_vstart[base + i + 1];
      int incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      int run = incl - sum;
      for (int i = 0; i < per; ++i)
        if (base + i < n_units) {
          run += seg_vstart[base + i + 1];
          seg_vstart[base + i + 1] = run;
        }
      if (lane == 0) seg_vstart[0] = 0;
    }
  }

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&ctrl->full[s]), 128);
      mbar_init(smem_u32(&ctrl->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&ctrl->acc_full[a]), 1);
      mbar_init(smem_u32(&ctrl->acc_empty[a]), 128);
    }
    fence_mbar_init();
    // Balanced split of the concatenated pair lists: CTA b owns pairs [L*b/G, L*(b+1)/G). A work
    // unit is the intersection of that slice with one offset's list, so every CTA contracts the
    // same number of pairs (+-1) and issues at most (offsets touched) reductions into dW.
    const long long L = p.offsets[p.K];
    const int G = gridDim.x, b = blockIdx.x;
    const int pb = (int)(L * b / G), pe = (int)(L * (b + 1) / G);
    int lo = 0, hi = p.K;  // largest k with offsets[k] <= pb
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (p.offsets[mid] <= pb) lo = mid; else hi = mid;
    }
    ctrl->pair_begin = pb;
    ctrl->pair_end = pe;
    ctrl->k_first = lo;
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&ctrl->tmem_base), kWgTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  const int pair_begin = ctrl->pair_begin;
  const int pair_end = ctrl->pair_end;
  const int k_first = ctrl->k_first;

  auto make_cursor = [&]() {
    WgSegCursor c;
    c.offsets = p.offsets;
    c.vstart = blocked ? seg_vstart : nullptr;
    c.ustart = seg_ustart;
    c.K = p.K; c.U = n_units; c.G = gridDim.x; c.R = p.rounds; c.b = blockIdx.x;
    c.pair_begin = pair_begin; c.pair_end = pair_end; c.k = k_first;
    c.r = 0; c.u = -1; c.v0 = c.v1 = 0;
    return c;
  };

  if (warp < 4) {
    // ===================================== gather producers =====================================
    // 16 lanes fetch 256 contiguous bytes of one gathered row (two 128-byte MN blocks): the
    // L2->SM path serves about one gather request per 8 cycles per SM whatever its size
    // (tools/gather_bench.cu), so whole-row requests double the gather rate of 128-byte ones.
    const uint8_t* xs = reinterpret_cast<const uint8_t*>(p.feats);
    const uint8_t* gs = reinterpret_cast<const uint8_t*>(p.gout);
    // row pitches fit 32 bits (checked by the launcher): one IMAD.WIDE.U32 per source address
    const uint32_t in_ld_bytes = (uint32_t)(p.in_ld * kElem);
    const uint32_t out_ld_bytes = (uint32_t)(p.out_ld * kElem);
    const int u = lane & 15;       // 16-byte unit inside a 256-byte row segment
    const int blk = u >> 3;        // 128-byte MN block inside the segment
    const int c16 = u & 7;
    const int sub = lane >> 4;     // which of the 2 pair rows an instruction covers
    constexpr int kRowsPerWarp = kPairs / 4;   // pair rows a warp gathers per stage (16 or 32)
    constexpr int kInstr = kRowsPerWarp / 2;   // cp.async per thread, operand and segment
    constexpr int kDepth = 4;                  // stages of pair-index look-ahead
    static_assert(kRowsPerWarp <= 32, "one index per lane");
    const uint8_t* a_base = xs + (long long)(p.in_coff + ys * p.in_y_stride) * kElem + u * 16;
    const uint8_t* b_base =
        gs + (long long)(p.out_coff + ys * p.out_y_stride + zs * p.out_z_stride) * kElem + u * 16;
    const int cin_bytes = cin_here * kElem;
    const int cout_bytes = p.cout * kElem;
    // pair row of instruction q: pr(q) = warp*kRowsPerWarp + 2q + sub; (pr & 7) = (2q + sub) & 7
    uint32_t off_par[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r7 = 2 * i + sub;
      off_par[i] = (uint32_t)blk * (kPairs * 128) +
                   (uint32_t)(warp * kRowsPerWarp + r7) * 128u + (uint32_t)((c16 ^ r7) << 4);
    }
    int stage = 0;
    uint32_t phase = 0;
    long long w_empty = 0, n_stage = 0, t_issue = 0, t_arrive = 0;
    const long long t_start = WCN_CLOCK();
    // identity offset through the TMA unit (dense slab, one 128-byte-block pair per operand)
    const bool tma_ok = NSEGB == 1 && kPairs == 64 && p.identity_k >= 0 &&
                        (p.status == nullptr || (__ldg(p.status) & 4) == 0);
    WgSegCursor seg = make_cursor();
    int k, first, count;
    while (seg.next(k, first, count)) {
      const int n_st = (count + kPairs - 1) / kPairs;
      const bool ident = tma_ok && k == p.identity_k;
      const int ident_row0 = ident ? first - __ldg(p.offsets + k) : 0;
      // lane l < kRowsPerWarp holds the input row, lane 16 + ... the output row of pair
      // warp*kRowsPerWarp + l of the stage (kRowsPerWarp = 16: both in one register)
      auto load_idx = [&](int st, int& vi, int& vo) {
        const int pr = st * kPairs + warp * kRowsPerWarp + (lane % kRowsPerWarp);
        const bool ok = pr < count;
        if (kRowsPerWarp == 16) {
          const int* src = (lane < 16) ? p.in_maps : p.out_maps;
          vi = ok ? __ldg(src + first + pr) : -1;
          vo = vi;
        } else {
          vi = ok ? __ldg(p.in_maps + first + pr) : -1;
          vo = ok ? __ldg(p.out_maps + first + pr) : -1;
        }
      };
      int vi_r[kDepth], vo_r[kDepth];
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
        vi_r[d] = -1;
        vo_r[d] = -1;
        if (d < n_st) load_idx(d, vi_r[d], vo_r[d]);
      }
      for (int st0 = 0; st0 < n_st; st0 += kDepth) {
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
          const int st = st0 + d;
          if (st >= n_st) break;
          const int vi = vi_r[d], vo = vo_r[d];
          if (st + kDepth < n_st) load_idx(st + kDepth, vi_r[d], vo_r[d]);
          {
            const long long t0 = WCN_CLOCK();
            mbar_wait(smem_u32(&ctrl->empty[stage]), phase ^ 1u);
            w_empty += WCN_CLOCK() - t0;
            ++n_stage;
          }
          const uint32_t a_smem = smem_base + stage * stage_bytes;
          const uint32_t b_smem = a_smem + kAStage;
          const long long t_i0 = WCN_CLOCK();
          if (ident && (st + 1) * kPairs <= count) {
            // pairs (r, r) for r = ident_row0 + st*64 ...: four 64-row x 128-byte boxes, swizzled
            // by the TMA unit exactly like the gathered layout ([block][pair row][128 B])
            if (tid == 0) {
              const uint32_t bar = smem_u32(&ctrl->full[stage]);
              const int r = ident_row0 + st * kPairs;
              mbar_expect_tx(bar, 4u * kPairs * 128u);
              tma_load_2d(a_smem, &tmap_x, 0, r, bar);
              tma_load_2d(a_smem + kPairs * 128, &tmap_x, kBlkElems, r, bar);
              tma_load_2d(b_smem, &tmap_dy, 0, r, bar);
              tma_load_2d(b_smem + kPairs * 128, &tmap_dy, kBlkElems, r, bar);
            }
            cp_async_mbar_arrive_noinc(smem_u32(&ctrl->full[stage]));
            if (++stage == stages) { stage = 0; phase ^= 1u; }
            continue;
          }
#pragma unroll
          for (int q = 0; q < kInstr; ++q) {
            const int row = 2 * q + sub;  // pair row inside this warp's share of the stage
            int pi, po;
            if (kRowsPerWarp == 16) {
              pi = __shfl_sync(0xffffffffu, vi, row);
              po = __shfl_sync(0xffffffffu, vi, 16 + row);
            } else {
              pi = __shfl_sync(0xffffffffu, vi, row);
              po = __shfl_sync(0xffffffffu, vo, row);
            }
            const bool valid = pi >= 0;
            const uint8_t* xrow = a_base + (unsigned long long)(uint32_t)max(pi, 0) * in_ld_bytes;
            const uint8_t* grow = b_base + (unsigned long long)(uint32_t)max(po, 0) * out_ld_bytes;
            const uint32_t sz = valid ? 16u : 0u;
            const uint32_t soff = off_par[q & 3] + (uint32_t)(q >> 2) * 1024u;
            // branch-free: out-of-range lanes copy 0 source bytes (= zero fill)
            cp_async_16(a_smem + soff, xrow, (u * 16 < cin_bytes) ? sz : 0u);
#pragma unroll
            for (int sgi = 0; sgi < NSEGB; ++sgi)
              cp_async_16(b_smem + sgi * 2 * (kPairs * 128) + soff, grow + sgi * 256,
                          (sgi * 256 + u * 16 < cout_bytes) ? sz : 0u);
          }
          const long long t_i1 = WCN_CLOCK();
          cp_async_mbar_arrive_noinc(smem_u32(&ctrl->full[stage]));
          const long long t_i2 = WCN_CLOCK();
          t_issue += t_i1 - t_i0;
          t_arrive += t_i2 - t_i1;
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    if (WCN_DBG_OUT(p) != nullptr && tid == 0 && WCN_DBG(p, 1024)) {
      WCN_DBG_OUT(p)[blockIdx.x * 8 + 3] = t_issue;
      WCN_DBG_OUT(p)[blockIdx.x * 8 + 4] = t_arrive;
    }
    if (WCN_DBG_OUT(p) != nullptr && tid == 0) {
      WCN_DBG_OUT(p)[blockIdx.x * 8 + 0] = WCN_CLOCK() - t_start;
      WCN_DBG_OUT(p)[blockIdx.x * 8 + 1] = w_empty;
      WCN_DBG_OUT(p)[blockIdx.x * 8 + 7] = n_stage;
    }
  } else if (warp == 4) {
    // ======================================= MMA issuer =========================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(ElemTraits<T>::kFmt, 128, p.cout, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t use = 0;
      long long w_full = 0, w_acc = 0;
      const long long t_start = WCN_CLOCK();
      WgSegCursor seg = make_cursor();
      int k, first, count;
      while (seg.next(k, first, count)) {
        const uint32_t acc = use & 1u;
        {
          const long long t0 = WCN_CLOCK();
          mbar_wait(smem_u32(&ctrl->acc_empty[acc]), ((use >> 1) & 1u) ^ 1u);
          w_acc += WCN_CLOCK() - t0;
        }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kWgAccStride;
        uint32_t accumulate = 0;
        for (int done = 0; done < count; done += kPairs) {
          {
            const long long t0 = WCN_CLOCK();
            mbar_wait(smem_u32(&ctrl->full[stage]), phase);
            w_full += WCN_CLOCK() - t0;
          }
          if (!WCN_DBG(p, 64)) fence_proxy_async_smem();  // cp.async = generic-proxy writes
          tc_fence_after();
          const uint32_t a_smem = smem_base + stage * stage_bytes;
          const uint32_t b_smem = a_smem + kAStage;
#pragma unroll
          for (int j = 0; j < kPairs / kKPerMma && !WCN_DBG(p, 256); ++j) {
            const uint32_t koff = j * kKPerMma * 128;  // kKPerMma pair rows of 128 bytes
            const uint64_t adesc = make_smem_desc_sw128(a_smem + koff, kPairs * 128, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(b_smem + koff, kPairs * 128, 1024);
            umma_ss<ElemTraits<T>::kTF32>(tmem_d, adesc, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(smem_u32(&ctrl->empty[stage]));
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&ctrl->acc_full[acc]));
        ++use;
      }
      if (WCN_DBG_OUT(p) != nullptr && !WCN_DBG(p, 1024)) {
        WCN_DBG_OUT(p)[blockIdx.x * 8 + 2] = WCN_CLOCK() - t_start;
        WCN_DBG_OUT(p)[blockIdx.x * 8 + 3] = w_full;
        WCN_DBG_OUT(p)[blockIdx.x * 8 + 4] = w_acc;
      }
    }
  } else {
    // ======================================== epilogue ==========================================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // input-channel row of the slab handled by this thread
    const int gl = (p.gps > 1) ? r / p.cin_g : 0;
    uint32_t use = 0;
    long long w_accf = 0;
    const long long t_start = WCN_CLOCK();
    WgSegCursor seg = make_cursor();
    int k, first, count;
    while (seg.next(k, first, count)) {
      const uint32_t acc = use & 1u;
      {
        const long long t0 = WCN_CLOCK();
        mbar_wait(smem_u32(&ctrl->acc_full[acc]), (use >> 1) & 1u);
        w_accf += WCN_CLOCK() - t0;
      }
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kWgAccStride;
      float* base = p.dw + (long long)k * p.dw_k_stride + ys * p.dw_y_stride + zs * p.dw_z_stride;
      if (p.gps == 1) {
        float* dst = base + (long long)r * p.dw_ld;
        for (int col = 0; col < p.cout; col += 16) {
          uint32_t v[16];
          tmem_ld_x16(trow + col, v);
          tmem_ld_wait();
          if (r < cin_here) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              red_add_v4(dst + col + i, p.alpha * __uint_as_float(v[i]),
                         p.alpha * __uint_as_float(v[i + 1]), p.alpha * __uint_as_float(v[i + 2]),
                         p.alpha * __uint_as_float(v[i + 3]));
          }
        }
      } else {
        // densified group conv: keep only the diagonal block of each row's group. tcgen05.ld is
        // warp-collective with one address, so walk the groups this warp's 32 rows belong to.
        const int g_lo = (q * 32) / p.cin_g;
        const int g_hi = min(p.gps - 1, (q * 32 + 31) / p.cin_g);
        for (int gg = g_lo; gg <= g_hi; ++gg) {
          float* dst = base + (long long)gg * p.dw_g_stride + (long long)(r - gg * p.cin_g) * p.dw_ld;
          for (int c = 0; c < p.cout_g; c += 8) {
            uint32_t v[8];
            tmem_ld_x8(trow + gg * p.cout_g + c, v);
            tmem_ld_wait();
            if (gl == gg && r < cin_here) {
              red_add_v4(dst + c, p.alpha * __uint_as_float(v[0]), p.alpha * __uint_as_float(v[1]),
                         p.alpha * __uint_as_float(v[2]), p.alpha * __uint_as_float(v[3]));
              red_add_v4(dst + c + 4, p.alpha * __uint_as_float(v[4]),
                         p.alpha * __uint_as_float(v[5]), p.alpha * __uint_as_float(v[6]),
                         p.alpha * __uint_as_float(v[7]));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&ctrl->acc_empty[acc]));
      ++use;
    }
    if (WCN_DBG_OUT(p) != nullptr && tid == 5 * 32) {
      WCN_DBG_OUT(p)[blockIdx.x * 8 + 5] = WCN_CLOCK() - t_start;
      WCN_DBG_OUT(p)[blockIdx.x * 8 + 6] = w_accf;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kWgTmemCols);
  }
}

typedef CUresult (*WcnEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static WcnEncodeTiledFn encode_tiled_fn() {
  static WcnEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WcnEncodeTiledFn>(sym);
  }
  return fn;
}

// [n_rows, channels] row-major matrix as a 2-D tensor map with a (64-row x 128-byte) box and the
// 128-byte swizzle of the MN-major operand layout. Returns false when the matrix cannot be mapped.
static bool make_row_tile_map(CUtensorMap* map, const void* base, long long ld_elems, int channels,
                              long long n_rows, int es, int box_rows, bool is_half) {
  WcnEncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr || n_rows < 1 || (ld_elems * es) % 16 != 0 ||
      (reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)channels, (cuuint64_t)n_rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)(ld_elems * es)};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (is_half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                            : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  return fn(map, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T, int PAIRS, int NSEGB>
static int launch_wgrad_t(WgradParams p, int cin_slabs, int cout_slabs, int max_ctas,
                          long long n_in_rows, long long n_out_rows, cudaStream_t stream) {
  constexpr int kElem = (int)sizeof(T);
  constexpr int kBlkElems = 128 / kElem;
  constexpr int kPairs = PAIRS;
  const int a_stage = (128 / kBlkElems) * kPairs * 128;
  const int n_blk_b = (p.cout + kBlkElems - 1) / kBlkElems;
  const int stage_bytes = a_stage + 2 * NSEGB * kPairs * 128;
  if (p.stages <= 0) {
    p.stages = (int)((227 * 1024 - sizeof(WgSmemCtrl) - kWgSegTableBytes - 1024) / stage_bytes);
    if (p.stages > kWgMaxStages) p.stages = kWgMaxStages;
  }
  if (p.stages < 2) return kErrUnsupportedShape;
  (void)n_blk_b;
  // TMA tiles for the identity offset: single dense slab, 16-bit operands, both matrices mappable
  CUtensorMap tmap_x, tmap_dy;
  memset(&tmap_x, 0, sizeof(tmap_x));
  memset(&tmap_dy, 0, sizeof(tmap_dy));
  if (p.identity_k >= 0) {
    const bool ok = NSEGB == 1 && kPairs == 64 && kElem == 2 && cin_slabs == 1 && cout_slabs == 1 &&
                    p.gps == 1 && p.in_coff == 0 && p.out_coff == 0 && p.identity_k < p.K &&
                    make_row_tile_map(&tmap_x, p.feats, p.in_ld, p.cin, n_in_rows, kElem, kPairs,
                                      ElemTraits<T>::kFmt == 0) &&
                    make_row_tile_map(&tmap_dy, p.gout, p.out_ld, p.cout, n_out_rows, kElem, kPairs,
                                      ElemTraits<T>::kFmt == 0);
    if WCN_DBG(p, 2048)  // bring-up: WCN_DEBUG=2048 reports whether the TMA identity path is on
      fprintf(stderr, "wcn wgrad: identity_k=%d tma_tiles=%d (encode fn %p)\n", p.identity_k, (int)ok,
              (void*)encode_tiled_fn());
    if (!ok) p.identity_k = -1;
  }
  const size_t smem = (size_t)p.stages * stage_bytes + sizeof(WgSmemCtrl) + kWgSegTableBytes + 1024;
  static int configured[kMaxDevices] = {};  // per instantiation and device
  int& configured_smem = configured[current_device_slot()];
  if ((int)smem > configured_smem) {
    if (cudaFuncSetAttribute(wgrad_kernel<T, PAIRS, NSEGB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return kErrCuda;
    configured_smem = (int)smem;
  }
  int per_slab = max_ctas / (cin_slabs * cout_slabs);
  if (per_slab < 1) per_slab = 1;
  dim3 grid(per_slab, cin_slabs, cout_slabs);
  wcn_launch(wgrad_kernel<T, PAIRS, NSEGB>, dim3(grid), dim3(kWgThreads), smem, stream, p, tmap_x, tmap_dy);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
}

int launch_wgrad(const WgradParams& p, int dtype, int cin_slabs, int cout_slabs, int max_ctas,
                 long long n_in_rows, long long n_out_rows, cudaStream_t stream) {
  const int es = dtype_size(dtype);
  if (p.K > kWgMaxK || p.K < 1) return kErrUnsupportedShape;
  if (p.cout < 16 || p.cout > 256 || p.cout % 16 != 0) return kErrUnsupportedShape;
  if (p.cin < 1 || p.cin > 128 || (p.cin * es) % 16 != 0) return kErrUnsupportedShape;
  if (p.in_ld * es >= (1ll << 31) || p.out_ld * es >= (1ll << 31)) return kErrUnsupportedShape;
  if ((p.in_ld * es) % 16 != 0 || (p.in_coff * es) % 16 != 0) return kErrAlignment;
  if ((p.out_ld * es) % 16 != 0 || (p.out_coff * es) % 16 != 0) return kErrAlignment;
  if (p.dw_ld % 4 != 0 || p.dw_k_stride % 4 != 0 || (reinterpret_cast<uintptr_t>(p.dw) & 15))
    return kErrAlignment;
  if ((reinterpret_cast<uintptr_t>(p.feats) & 15) || (reinterpret_cast<uintptr_t>(p.gout) & 15))
    return kErrAlignment;
  switch (dtype) {
    case kBF16:
      return p.cout * es > 256 ? launch_wgrad_t<__nv_bfloat16, 64, 2>(p, cin_slabs, cout_slabs, max_ctas, n_in_rows, n_out_rows, stream)
                               : launch_wgrad_t<__nv_bfloat16, 64, 1>(p, cin_slabs, cout_slabs, max_ctas, n_in_rows, n_out_rows, stream);
    case kF16:
      return p.cout * es > 256 ? launch_wgrad_t<__half, 64, 2>(p, cin_slabs, cout_slabs, max_ctas, n_in_rows, n_out_rows, stream)
                               : launch_wgrad_t<__half, 64, 1>(p, cin_slabs, cout_slabs, max_ctas, n_in_rows, n_out_rows, stream);
    // fp32 rows as MN-major tf32 operands returned zeros on B200 (bring-up log round 1); the host
    // side splits fp32 into bf16 hi/lo parts instead (detail/unified.py:_wgrad_call).
    default: return kErrUnsupportedDtype;
  }
}

}  // namespace wcn
