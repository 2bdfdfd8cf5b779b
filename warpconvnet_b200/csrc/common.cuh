// SPDX-License-Identifier: Apache-2.0
// Thin inline-PTX layer for sm_100a: mbarrier, cp.async, bulk copy (TMA unit), tcgen05/TMEM.
// Everything here is hand-written for B200; nothing is borrowed from a template library.
#pragma once
#include <cstdlib>

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <stdint.h>

namespace wcn {

// ---------------------------------------------------------------------------------------------
// error codes of the C-ABI (mirrors the reference's "0 = ok, negative = unsupported" convention,
// reference: warpconvnet/csrc/include/gemm_error_codes.h:7-15)
// ---------------------------------------------------------------------------------------------
enum Status : int {
  kOk = 0,
  kErrInvalidArg = -1,
  kErrUnsupportedShape = -2,
  kErrAlignment = -3,
  kErrUnsupportedDtype = -4,
  kErrCuda = -5,
  kErrWorkspace = -6,
};

enum DType : int { kBF16 = 0, kF16 = 1, kF32 = 2 };

__host__ __device__ inline int dtype_size(int dt) { return dt == kF32 ? 4 : 2; }

constexpr int kNumSMsB200 = 148;

// Per-role wait-cycle counters of the GEMM kernels (tools/exp_dbg.py, tools/exp_wgrad.py) are a
// bring-up aid: the clock reads are compiled in only with -DWCN_KERNEL_COUNTERS
// (WCN_KERNEL_COUNTERS=1 csrc/build.sh); the shipped library carries none of them.
// Programmatic dependent launch (PDL): every kernel of this library starts with pdl_begin()
// (griddepcontrol.launch_dependents + griddepcontrol.wait: nothing is read or written before the
// wait, so semantics are unchanged) and goes through wcn_launch(). With the launch attribute set
// (-DWCN_ENABLE_PDL) the next kernel of a stream is resident before its predecessor has drained,
// which removes the ~1-2 us gap between dependent kernels: C3 step 0.442 -> 0.436 ms. It is OFF in
// the default build: together with the grid-barrier sort kernel and a second stream (the CSR
// branch of the kernel-map build) back-to-back eager steps showed sporadic ~11 ms stalls
// (profiles/r2_pdl_and_barrier_kernel.md); without the attribute pdl_begin() is a no-op.
__device__ __forceinline__ void pdl_begin() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t wcn_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
#ifdef WCN_ENABLE_PDL
  cfg.numAttrs = 1;
#else
  cfg.numAttrs = 0;
#endif
#ifdef WCN_BRINGUP  // bring-up builds: WCN_PDL_OFF=0 / 1 overrides the build default (A/B)
  static const int pdl_env = [] { const char* e = getenv("WCN_PDL_OFF"); return e ? (e[0] == '1' ? 0 : 1) : -1; }();
  if (pdl_env >= 0) cfg.numAttrs = pdl_env;
#endif
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// Experiment switches (GemmParams::debug bits, per-CTA counter dumps through dbg_out) exist only
// in bring-up builds (WCN_BRINGUP=1 csrc/build.sh): in the shipped library WCN_DBG is the
// constant false and WCN_DBG_OUT the constant nullptr, so the predicates, the counter stores and
// the %globaltimer reads behind them are removed by the compiler.
#ifdef WCN_BRINGUP
#define WCN_DBG(params, bits) ((((params).debug) & (bits)) != 0)
#define WCN_DBG_OUT(params) ((params).dbg_out)
#else
#define WCN_DBG(params, bits) (false)
#define WCN_DBG_OUT(params) (static_cast<long long*>(nullptr))
#endif
#ifdef WCN_KERNEL_COUNTERS
#define WCN_CLOCK() clock64()
#else
#define WCN_CLOCK() 0ll
#endif

// number of kernels this library has launched in this process (wcn_launch_count in the C-ABI)
void count_launch();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: caches of "already configured"
// sizes are kept per device ordinal so a process that drives several GPUs stays correct.
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev % kMaxDevices;
}

// ---------------------------------------------------------------------------------------------
// shared-memory address helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// try_wait with a suspend-time hint: the waiting thread is parked by the hardware (it does not
// compete for issue slots) until the phase completes or the hint (ns) expires. Without the hint
// the default time limit is short and a polling loop of high-numbered warps (the scheduler
// prefers the highest warp id) starves the producer warps of the same SM sub-partition.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(1000000u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (CUDA error on the host), never
// as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      printf("wcn_b200: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cp.async (LDGSTS) 16-byte copies with zero fill, completion tracked by an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void st_shared_zero_16(uint32_t dst) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cta() {
  asm volatile("fence.acq_rel.cta;" ::: "memory");
}
__device__ __forceinline__ void cp_async_16_ca(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src),
               "r"(src_bytes)
               : "memory");
}
// The arrive fires when all cp.async issued so far by this thread have landed. ".noinc": the
// arrival is one of the barrier's expected arrivals (counted in mbar_init).
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------------------------------------
// bulk asynchronous copy global -> shared through the TMA unit (no tensor map needed)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                              uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// Orders generic-proxy writes to shared memory (cp.async / st.shared, made visible through an
// mbarrier) before async-proxy reads (tcgen05.mma operand fetch, TMA).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}

// one element of T (converted from fp32, round to nearest) to shared memory
template <typename T>
__device__ __forceinline__ void st_shared_elem(uint32_t addr, float f);
template <>
__device__ __forceinline__ void st_shared_elem<float>(uint32_t addr, float f) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(__float_as_uint(f)) : "memory");
}
template <>
__device__ __forceinline__ void st_shared_elem<__nv_bfloat16>(uint32_t addr, float f) {
  const __nv_bfloat16 h = __float2bfloat16_rn(f);
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const uint16_t*>(&h))
               : "memory");
}
template <>
__device__ __forceinline__ void st_shared_elem<__half>(uint32_t addr, float f) {
  const __half h = __float2half_rn(f);
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const uint16_t*>(&h))
               : "memory");
}

// fp32 -> T (round to nearest) -> fp32: the value a stored element of T holds
template <typename T>
__device__ __forceinline__ float round_to(float f);
template <>
__device__ __forceinline__ float round_to<float>(float f) { return f; }
template <>
__device__ __forceinline__ float round_to<__nv_bfloat16>(float f) {
  return __bfloat162float(__float2bfloat16_rn(f));
}
template <>
__device__ __forceinline__ float round_to<__half>(float f) {
  return __half2float(__float2half_rn(f));
}

// the 16 / sizeof(T) elements of one 16-byte piece as fp32
template <typename T>
__device__ __forceinline__ void unpack_piece(const uint4& v, float (&f)[16 / sizeof(T)]);
template <>
__device__ __forceinline__ void unpack_piece<float>(const uint4& v, float (&f)[4]) {
  f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
  f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
template <>
__device__ __forceinline__ void unpack_piece<__nv_bfloat16>(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <>
__device__ __forceinline__ void unpack_piece<__half>(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// adds to the pending transaction count of the current phase without arriving
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}

// TMA tiled 2-D load: box (c0 = innermost coordinate, c1 = row) of the tensor map lands densely
// (with the map's swizzle) at dst; out-of-range elements are zero-filled and still counted.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// TMA gather4: four rows (r0..r3, arbitrary) x one box width of a 2-D tensor map land as four
// consecutive box rows at dst; rows outside the tensor are zero-filled.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const void* tmap, int col, int r0, int r1,
                                            int r2, int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
      ".cta_group::1 [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / tensor memory
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread for the whole CTA.
template <int KIND_TF32>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
  if (KIND_TF32) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// Arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptor for tcgen05.mma (bit layout as documented for sm_100:
// [0,14) addr>>4, [16,30) leading-dim byte offset>>4, [32,46) stride-dim byte offset>>4,
// [46,48) version=1, [61,64) swizzle mode (2 = 128-byte swizzle)).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16 / kind::tf32 with fp32 accumulation.
// fmt: 0 = f16, 1 = bf16, 2 = tf32; a_mn/b_mn: 1 = operand is MN-major in shared memory.
__host__ __device__ inline uint32_t make_idesc(int fmt, int M, int N, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                          // accumulator format: f32
  d |= (uint32_t)fmt << 7;               // A format
  d |= (uint32_t)fmt << 10;              // B format
  d |= (uint32_t)(a_mn & 1) << 15;       // A major
  d |= (uint32_t)(b_mn & 1) << 16;       // B major
  d |= (uint32_t)(N >> 3) << 17;         // N / 8
  d |= (uint32_t)(M >> 4) << 24;         // M / 16
  return d;
}

// 128-byte-swizzle position of 16-byte unit `c16` (0..7) of 128-byte row `row`
// (rows are stored back to back, 8-row / 1024-byte atoms, atom base 1024-byte aligned).
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t c16) {
  return row * 128u + ((c16 ^ (row & 7u)) << 4);
}

template <typename T>
struct ElemTraits;
template <>
struct ElemTraits<__nv_bfloat16> {
  static constexpr int kFmt = 1;
  static constexpr int kTF32 = 0;
  static constexpr int kDType = kBF16;
};
template <>
struct ElemTraits<__half> {
  static constexpr int kFmt = 0;
  static constexpr int kTF32 = 0;
  static constexpr int kDType = kF16;
};
template <>
struct ElemTraits<float> {
  static constexpr int kFmt = 2;
  static constexpr int kTF32 = 1;
  static constexpr int kDType = kF32;
};

// pack 8 fp32 accumulators (as raw bits) into one 16-byte store of T
template <typename T>
__device__ __forceinline__ uint4 pack8(const uint32_t* v);
template <>
__device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const uint32_t* v) {
  uint4 r;
  __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(v[0]), __uint_as_float(v[1]));
  __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[2]), __uint_as_float(v[3]));
  __nv_bfloat162 c = __floats2bfloat162_rn(__uint_as_float(v[4]), __uint_as_float(v[5]));
  __nv_bfloat162 d = __floats2bfloat162_rn(__uint_as_float(v[6]), __uint_as_float(v[7]));
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c);
  r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}
template <>
__device__ __forceinline__ uint4 pack8<__half>(const uint32_t* v) {
  uint4 r;
  __half2 a = __floats2half2_rn(__uint_as_float(v[0]), __uint_as_float(v[1]));
  __half2 b = __floats2half2_rn(__uint_as_float(v[2]), __uint_as_float(v[3]));
  __half2 c = __floats2half2_rn(__uint_as_float(v[4]), __uint_as_float(v[5]));
  __half2 d = __floats2half2_rn(__uint_as_float(v[6]), __uint_as_float(v[7]));
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c);
  r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}

}  // namespace wcn
