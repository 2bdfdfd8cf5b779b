// SPDX-License-Identifier: Apache-2.0
// Per-channel normalisation + activation + residual passes over the [N, C] feature matrix that
// sits between two sparse convolutions. HBM-bound row streaming: every thread
// owns one 16-byte channel vector and walks rows with a grid stride, so a warp reads whole
// contiguous rows; per-channel reductions stay in registers, are merged per block through shared
// memory and leave as one fp64 atomic per channel and block.
//
// Replaces (semantics, not code): the feature-matrix chain of the reference's ConvBlock /
// BasicBlock — nn.BatchNorm1d -> ReLU (-> + identity -> ReLU), warpconvnet/models/mink_unet.py:31-53,
// 104-140, warpconvnet/nn/modules/normalizations.py:53-67 — which torch runs as 4-6 separate
// passes (collect statistics, transform, ReLU, add) forward and as many backward.
#include "common.cuh"
#include "rownorm.cuh"

namespace wcn {


template <typename T>
struct Vec16 {
  static constexpr int kElems = 16 / (int)sizeof(T);
};

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* p, float (&f)[V]) {
  if constexpr (V == 1) {
    f[0] = (float)p[0];
  } else {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
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
__device__ __forceinline__ void store_vec(T* p, const float (&f)[V]) {
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

// Raw 16-byte vectors: a kernel first issues ALL loads of an iteration as raw vectors (4 registers
// each instead of 8 converted floats) and converts when it consumes them — the passes are pure
// streaming, bound by the bytes the resident threads keep in flight (Little's law against ~1 us of
// loaded HBM latency: >= 64 KB per SM), so rows in flight per thread x resident threads is the
// quantity to maximise.
template <typename T, int V>
struct RawOf { using type = uint4; };
template <typename T>
struct RawOf<T, 1> { using type = T; };

template <typename T, int V>
__device__ __forceinline__ typename RawOf<T, V>::type load_raw(const T* p) {
  if constexpr (V == 1) return p[0];
  else return *reinterpret_cast<const uint4*>(p);
}

template <typename T, int V>
__device__ __forceinline__ void unpack(const typename RawOf<T, V>::type& raw, float (&f)[V]) {
  if constexpr (V == 1) {
    f[0] = (float)raw;
  } else if constexpr (sizeof(T) == 4) {
    f[0] = __uint_as_float(raw.x); f[1] = __uint_as_float(raw.y);
    f[2] = __uint_as_float(raw.z); f[3] = __uint_as_float(raw.w);
  } else {
    const T* h = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = (float)h[i];
  }
}

// Measured on 2.4 M x 96 bf16 rows (tools/exp_norm.py, profiles/r2u_norm_kernels.md): the
// one-tensor forward apply gains from 4 rows in flight (177 -> 156 us, 5.9 TB/s of the 6.35 TB/s a
// torch copy reaches); with a residual (two tensors in flight) and in the backward kernels two rows
// at three resident blocks are as good or better than four rows at two blocks.
#ifndef WCN_RN_UF
#define WCN_RN_UF 4   // rows in flight per thread, forward apply without a residual
#endif
#ifndef WCN_RN_UB
#define WCN_RN_UB 2   // rows in flight per thread, backward reduce / apply
#endif
#ifndef WCN_RN_BB
#define WCN_RN_BB 3   // resident blocks per SM the backward kernels are compiled for
#endif
#ifndef WCN_RN_BF
#define WCN_RN_BF 3   // resident blocks per SM the forward apply kernel is compiled for
#endif

constexpr int kRnThreads = 256;

// Thread -> (vector column vc, row lane rl): vecs = ceil(c / V) vectors per row, rpb = threads /
// vecs rows per block iteration. Threads beyond rpb * vecs idle.
struct RnMap {
  int vc, rl, rpb, vecs;
  bool active;
};
template <int V>
__device__ __forceinline__ RnMap rn_map(int c) {
  RnMap m;
  m.vecs = (c + V - 1) / V;
  m.rpb = kRnThreads / m.vecs;
  m.vc = threadIdx.x % m.vecs;
  m.rl = threadIdx.x / m.vecs;
  m.active = m.rl < m.rpb;
  return m;
}

// Block-level merge of per-thread partial sums a[V], b[V] into sums[0..c) and sums[c..2c): every
// thread parks its partials in shared memory ([rpb][2c] floats, <= 16 KB), 2c threads add the
// columns and leave one fp64 atomic each (the second half multiplied by half2_scale[c + ch] when
// given). No shared-memory atomics: fp32 atomicAdd on shared memory is a compare-and-swap loop,
// and rpb = 21-64 threads per address made it a fixed ~25 us tail of every reduction kernel
// (profiles/r2u_norm_kernels.md).
template <int V>
__device__ __forceinline__ void rn_merge(const RnMap& m, int c, const float (&a)[V],
                                         const float (&b)[V], double* sums, float* sh,
                                         const float* half2_scale = nullptr) {
  if (m.active) {
    float* row = sh + (size_t)m.rl * 2 * c;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int ch = m.vc * V + i;
      if (ch < c) {
        row[ch] = a[i];
        row[c + ch] = b[i];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * c; i += kRnThreads) {
    float t = 0.f;
    for (int r = 0; r < m.rpb; ++r) t += sh[(size_t)r * 2 * c + i];
    const float f = (half2_scale != nullptr && i >= c) ? __ldg(half2_scale + i) : 1.f;
    atomicAdd(sums + i, (double)t * (double)f);
  }
}

// ---- forward statistics: sums[ch] += sum_r x[r,ch], sums[c+ch] += sum_r x[r,ch]^2 -------------
template <typename T, int V>
__global__ void __launch_bounds__(kRnThreads, 3) bn_stats_kernel(const RowNormParams p) {
  pdl_begin();
  extern __shared__ float sh[];
  const RnMap m = rn_map<V>(p.c);
  float s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = q[i] = 0.f;
  if (m.active) {
    // eight independent raw 16-byte loads in flight per thread (the pass is pure HBM streaming:
    // 3 resident blocks x 256 threads x 8 x 16 B = 96 KB per SM)
    constexpr int U = 8;
    using Raw = typename RawOf<T, V>::type;
    const long long stride = (long long)gridDim.x * m.rpb;
    const T* xp = reinterpret_cast<const T*>(p.x) + m.vc * V +
                  ((long long)blockIdx.x * m.rpb + m.rl) * p.ld_x;
    const long long step = stride * p.ld_x;  // elements between two rows of this thread
    for (long long r0 = (long long)blockIdx.x * m.rpb + m.rl; r0 < p.n; r0 += U * stride) {
      Raw raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r0 + u * stride < p.n) raw[u] = load_raw<T, V>(xp + u * step);
      xp += U * step;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r0 + u * stride >= p.n) break;
        float f[V];
        unpack<T, V>(raw[u], f);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          s[i] += f[i];
          q[i] = fmaf(f[i], f[i], q[i]);
        }
      }
    }
  }
  rn_merge<V>(m, p.c, s, q, p.sums, sh);
}

// ---- y = act(x * scale + shift (+ res)) -----------------------------------------------------------
template <typename T, int V, bool RES>
__global__ void __launch_bounds__(kRnThreads, WCN_RN_BF) scale_shift_act_kernel(const RowNormParams p) {
  pdl_begin();
  const RnMap m = rn_map<V>(p.c);
  if (!m.active) return;
  float sc[V], sf[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int ch = min(m.vc * V + i, p.c - 1);
    sc[i] = __ldg(p.scale + ch);
    sf[i] = __ldg(p.shift + ch);
  }
  const T* x = reinterpret_cast<const T*>(p.x) + m.vc * V;
  const T* res = RES ? reinterpret_cast<const T*>(p.res) + m.vc * V : nullptr;
  T* y = reinterpret_cast<T*>(p.y) + m.vc * V;
  constexpr int U = RES ? 2 : WCN_RN_UF;  // rows in flight per thread
  using Raw = typename RawOf<T, V>::type;
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rl; r0 < p.n; r0 += U * stride) {
    Raw rx[U], rr[RES ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r < p.n) {
        rx[u] = load_raw<T, V>(x + r * p.ld_x);
        if constexpr (RES) rr[u] = load_raw<T, V>(res + r * p.ld_res);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r >= p.n) break;
      float f[V], g[V];
      unpack<T, V>(rx[u], f);
      if constexpr (RES) unpack<T, V>(rr[u], g);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float v = fmaf(f[i], sc[i], sf[i]);
        if constexpr (RES) v += g[i];
        f[i] = p.relu ? fmaxf(v, 0.f) : v;
      }
      store_vec<T, V>(y + r * p.ld_y, f);
    }
  }
}

// ---- backward reduce: dz = dy * (y > 0); sums += (sum dz, sum dz * xhat) ---------------------------
// Register diet (round 2): the kernel streams two tensors and is bound by the bytes its resident
// threads keep in flight, so occupancy matters more than instruction count. Per channel only the
// mean and the two mask constants stay in registers; sum dz * (x - mean) is accumulated and the
// factor rstd is applied once per block when the partial sums leave (4 resident blocks per SM
// instead of 2: 2.2 -> TB/s figures in profiles/r2_norm_kernels.md).
// MASK: 0 = no ReLU behind the norm, 1 = mask from the saved output y_in, 2 = mask recomputed from
// x (x * mask_scale + mask_shift > 0). A template parameter: the unused paths cost registers.
template <typename T, int V, int MASK>
__global__ void __launch_bounds__(kRnThreads, WCN_RN_BB) bn_bwd_reduce_kernel(const RowNormParams p) {
  pdl_begin();
  extern __shared__ float sh[];
  const RnMap m = rn_map<V>(p.c);
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
  if (m.active) {
    float mu[V], msc[MASK == 2 ? V : 1], msh[MASK == 2 ? V : 1];
    constexpr bool mask_x = MASK == 2;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int ch = min(m.vc * V + i, p.c - 1);
      mu[i] = __ldg(p.mean_rstd + ch);
      if constexpr (MASK == 2) {
        msc[i] = __ldg(p.mask_scale + ch);
        msh[i] = __ldg(p.mask_shift + ch);
      }
    }
    const T* x = reinterpret_cast<const T*>(p.x) + m.vc * V;
    const T* dy = reinterpret_cast<const T*>(p.dy) + m.vc * V;
    const T* yin = MASK == 1 ? reinterpret_cast<const T*>(p.y_in) + m.vc * V : nullptr;
    constexpr int U = WCN_RN_UB;  // rows in flight per thread
    using Raw = typename RawOf<T, V>::type;
    const long long stride = (long long)gridDim.x * m.rpb;
    for (long long r0 = (long long)blockIdx.x * m.rpb + m.rl; r0 < p.n; r0 += U * stride) {
      Raw rx[U], rd[U], ry[MASK == 1 ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * stride;
        if (r < p.n) {
          rx[u] = load_raw<T, V>(x + r * p.ld_x);
          rd[u] = load_raw<T, V>(dy + r * p.ld_dy);
          if constexpr (MASK == 1) ry[u] = load_raw<T, V>(yin + r * p.ld_yin);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r0 + u * stride >= p.n) break;
        float fx[V], fd[V], fy[V];
        unpack<T, V>(rx[u], fx);
        unpack<T, V>(rd[u], fd);
        if constexpr (MASK == 1) unpack<T, V>(ry[u], fy);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          bool off = false;
          if constexpr (MASK == 1) off = !(fy[i] > 0.f);
          if constexpr (mask_x) off = !(fmaf(fx[i], msc[i], msh[i]) > 0.f);
          const float d = off ? 0.f : fd[i];
          s1[i] += d;
          s2[i] = fmaf(d, fx[i] - mu[i], s2[i]);
        }
      }
    }
  }
  // block merge; the second half leaves multiplied by rstd (mean_rstd[c + ch]):
  // sum dz * xhat = rstd * sum dz * (x - mean)
  rn_merge<V>(m, p.c, s1, s2, p.sums, sh, p.mean_rstd);
}

// ---- backward apply: dx = gamma * rstd * (dz - s1/n - xhat * s2/n); dres = dz ----------------------
// folded per channel into dx = A * dz + B * x + C (A = gamma * rstd, B = -A * rstd * s2/n,
// C = -A * s1/n - B * mean): three constants + the two mask constants per channel in registers.
template <typename T, int V, int MASK, bool TRAIN>
__global__ void __launch_bounds__(kRnThreads, WCN_RN_BB) bn_bwd_apply_kernel(const RowNormParams p) {
  pdl_begin();
  const RnMap m = rn_map<V>(p.c);
  if (!m.active) return;
  float ca[V], cb[TRAIN ? V : 1], cc[TRAIN ? V : 1], msc[MASK == 2 ? V : 1], msh[MASK == 2 ? V : 1];
  constexpr bool mask_x = MASK == 2;  // (the launcher only picks it in training mode: x is read)
  const float inv_n = 1.f / (float)p.n;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int ch = min(m.vc * V + i, p.c - 1);
    const float g = __ldg(p.scale + ch);  // gamma (training) or gamma * rstd_running (eval)
    if constexpr (MASK == 2) {
      msc[i] = __ldg(p.mask_scale + ch);
      msh[i] = __ldg(p.mask_shift + ch);
    }
    if constexpr (TRAIN) {
      const float mu = __ldg(p.mean_rstd + ch);
      const float rs = __ldg(p.mean_rstd + p.c + ch);
      const float m1 = (float)(p.sums[ch] * (double)inv_n);
      const float m2 = (float)(p.sums[p.c + ch] * (double)inv_n);
      ca[i] = g * rs;
      cb[i] = -ca[i] * rs * m2;
      cc[i] = -ca[i] * m1 - cb[i] * mu;
    } else {
      ca[i] = g;
    }
  }
  const T* x = reinterpret_cast<const T*>(p.x) + m.vc * V;
  const T* dy = reinterpret_cast<const T*>(p.dy) + m.vc * V;
  const T* yin = MASK == 1 ? reinterpret_cast<const T*>(p.y_in) + m.vc * V : nullptr;
  T* dx = reinterpret_cast<T*>(p.y) + m.vc * V;
  T* dres = p.dres ? reinterpret_cast<T*>(p.dres) + m.vc * V : nullptr;
  constexpr int U = WCN_RN_UB;  // rows in flight per thread
  using Raw = typename RawOf<T, V>::type;
  const long long stride = (long long)gridDim.x * m.rpb;
  for (long long r0 = (long long)blockIdx.x * m.rpb + m.rl; r0 < p.n; r0 += U * stride) {
    Raw rx[TRAIN ? U : 1], rd[U], ry[MASK == 1 ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r < p.n) {
        rd[u] = load_raw<T, V>(dy + r * p.ld_dy);
        if constexpr (MASK == 1) ry[u] = load_raw<T, V>(yin + r * p.ld_yin);
        if constexpr (TRAIN) rx[u] = load_raw<T, V>(x + r * p.ld_x);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      if (r >= p.n) break;
      float fx[V], fd[V], fy[V];
      unpack<T, V>(rd[u], fd);
      if constexpr (TRAIN) unpack<T, V>(rx[u], fx);
      if constexpr (MASK == 1) {
        unpack<T, V>(ry[u], fy);
#pragma unroll
        for (int i = 0; i < V; ++i) fd[i] = fy[i] > 0.f ? fd[i] : 0.f;
      } else if constexpr (mask_x) {
#pragma unroll
        for (int i = 0; i < V; ++i) fd[i] = fmaf(fx[i], msc[i], msh[i]) > 0.f ? fd[i] : 0.f;
      }
      if (dres) store_vec<T, V>(dres + r * p.ld_dres, fd);
      if constexpr (TRAIN) {
#pragma unroll
        for (int i = 0; i < V; ++i) fd[i] = fmaf(ca[i], fd[i], fmaf(cb[i], fx[i], cc[i]));
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) fd[i] *= ca[i];
      }
      store_vec<T, V>(dx + r * p.ld_y, fd);
    }
  }
}

// ---- finalize: mean / rstd / scale / shift / running statistics from the fp64 sums ----------------
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int n, int c,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, float* running_mean,
                                   float* running_var, float* scale, float* shift,
                                   float* mean_rstd) {
  pdl_begin();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const double mean = sums[ch] / (double)n;
  double var = sums[c + ch] / (double)n - mean * mean;  // biased (normalisation)
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[ch] : 1.f;
  const float b = beta ? beta[ch] : 0.f;
  scale[ch] = g * rstd;
  shift[ch] = b - (float)mean * g * rstd;
  mean_rstd[ch] = (float)mean;
  mean_rstd[c + ch] = rstd;
  if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
enum RnKernel { kStats = 0, kApply = 1, kBwdReduce = 2, kBwdApply = 3 };

// One full wave of resident blocks (grid-stride rows): blocks = SMs x occupancy of the kernel,
// never more than the rows need. Reductions: every block gets >= 8 row iterations (the per-block
// merge is paid once per block).
static int rn_grid(int n, int c, int v, int resident_per_sm, bool reduction) {
  const int vecs = (c + v - 1) / v;
  const int rpb = kRnThreads / vecs;
  const long long rows_per_block = (long long)rpb * (reduction ? 8 : 1);
  long long blocks = ((long long)n + rows_per_block - 1) / rows_per_block;
  const long long cap = (long long)kNumSMsB200 * resident_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename K>
static int rn_occupancy(K kernel, size_t smem) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kRnThreads, smem) != cudaSuccess ||
      occ < 1)
    occ = 2;
  return occ;
}

template <typename T, int V>
static int rn_launch_tv(int which, const RowNormParams& p, cudaStream_t s) {
  const size_t sh = (size_t)(kRnThreads / ((p.c + V - 1) / V)) * 2 * p.c * sizeof(float);  // [rpb][2c]
  static int occ[4] = {0, 0, 0, 0};  // per instantiation and kernel
  if (occ[which] == 0) {
    switch (which) {
      case kStats: occ[which] = rn_occupancy(bn_stats_kernel<T, V>, sh); break;
      case kApply: occ[which] = rn_occupancy(scale_shift_act_kernel<T, V, true>, 0); break;
      case kBwdReduce: occ[which] = rn_occupancy(bn_bwd_reduce_kernel<T, V, 2>, sh); break;
      default: occ[which] = rn_occupancy(bn_bwd_apply_kernel<T, V, 2, true>, 0); break;
    }
  }
  const int grid = rn_grid(p.n, p.c, V, occ[which], which == kStats || which == kBwdReduce);
  switch (which) {
    case kStats: wcn_launch(bn_stats_kernel<T, V>, dim3(grid), dim3(kRnThreads), sh, s, p); break;
    case kApply:
      if (p.res != nullptr) wcn_launch(scale_shift_act_kernel<T, V, true>, dim3(grid), dim3(kRnThreads), 0, s, p);
      else wcn_launch(scale_shift_act_kernel<T, V, false>, dim3(grid), dim3(kRnThreads), 0, s, p);
      break;
    case kBwdReduce: {
      const int mask = p.mask_scale != nullptr ? 2 : (p.y_in != nullptr ? 1 : 0);
      if (mask == 2) wcn_launch(bn_bwd_reduce_kernel<T, V, 2>, dim3(grid), dim3(kRnThreads), sh, s, p);
      else if (mask == 1) wcn_launch(bn_bwd_reduce_kernel<T, V, 1>, dim3(grid), dim3(kRnThreads), sh, s, p);
      else wcn_launch(bn_bwd_reduce_kernel<T, V, 0>, dim3(grid), dim3(kRnThreads), sh, s, p);
      break;
    }
    default: {
      // the mask is recomputed from x only in training mode (eval mode does not read x)
      const int mask = (p.mask_scale != nullptr && p.training) ? 2 : (p.y_in != nullptr ? 1 : 0);
#define WCN_BN_APPLY(MK, TR) \
  wcn_launch(bn_bwd_apply_kernel<T, V, MK, TR>, dim3(grid), dim3(kRnThreads), 0, s, p)
      if (p.training) {
        if (mask == 2) WCN_BN_APPLY(2, true); else if (mask == 1) WCN_BN_APPLY(1, true); else WCN_BN_APPLY(0, true);
      } else {
        if (mask == 1) WCN_BN_APPLY(1, false); else WCN_BN_APPLY(0, false);
      }
#undef WCN_BN_APPLY
      break;
    }
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
}

template <typename T>
static int rn_launch_t(int which, const RowNormParams& p, bool vec, cudaStream_t s) {
  constexpr int V = Vec16<T>::kElems;
  return vec ? rn_launch_tv<T, V>(which, p, s) : rn_launch_tv<T, 1>(which, p, s);
}

static bool aligned16(const void* ptr, long long ld, int es) {
  return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * es) % 16 == 0);
}

int rownorm_launch(int which, const RowNormParams& p, int dtype, cudaStream_t s) {
  if (p.n < 0 || p.c < 1 || p.c > 4096) return kErrInvalidArg;
  if (p.n == 0) return kOk;
  const int es = dtype_size(dtype);
  const int v = 16 / es;
  // 16-byte channel vectors when every operand allows it, else one element per thread
  const bool vec = (p.c % v == 0) && p.c / v <= kRnThreads && aligned16(p.x, p.ld_x, es) &&
                   aligned16(p.res, p.ld_res, es) && aligned16(p.y_in, p.ld_yin, es) &&
                   aligned16(p.dy, p.ld_dy, es) && aligned16(p.y, p.ld_y, es) &&
                   aligned16(p.dres, p.ld_dres, es);
  if (!vec && p.c > kRnThreads) return kErrUnsupportedShape;
  switch (dtype) {
    case kBF16: return rn_launch_t<__nv_bfloat16>(which, p, vec, s);
    case kF16: return rn_launch_t<__half>(which, p, vec, s);
    case kF32: return rn_launch_t<float>(which, p, vec, s);
    default: return kErrUnsupportedDtype;
  }
}

int bn_finalize(const double* sums, int n, int c, const float* gamma, const float* beta, float eps,
                float momentum, float* running_mean, float* running_var, float* scale,
                float* shift, float* mean_rstd, cudaStream_t s) {
  if (c < 1 || n < 1) return kErrInvalidArg;
  wcn_launch(bn_finalize_kernel, dim3((c + 127) / 128), dim3(128), 0, s, sums, n, c, gamma, beta, eps, momentum,
                                                     running_mean, running_var, scale, shift,
                                                     mean_rstd);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? kOk : kErrCuda;
}

}  // namespace wcn
