// Shared device/host helpers for the sm_100a kernels of libpegainfer_kernels_b200.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../include/pegainfer_kernels.h"

namespace pk {

using bf16 = __nv_bfloat16;

constexpr int kWarp = 32;

// ---- launch bookkeeping (per rank thread, like the reference's cuBLAS handles) ----
struct ThreadState {
  int64_t launches = 0;
  bool pdl = false;
  // scratch created by cublas_init()
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  int sm_count = 0;
  int device = -1;
  // split-K workspace of the skinny tensor-core GEMM (gemm.cu): fp32 partial tiles + per-tile tickets
  float* gemm_part = nullptr;
  unsigned int* gemm_cnt = nullptr;
  size_t gemm_part_bytes = 0;
};
ThreadState& tls();
int sm_count();

// Launch with optional programmatic dependent launch (PDL).  Kernels that opt in call
// pdl_wait() before touching data produced by the previous kernel in the stream.
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          cudaStream_t stream, bool allow_pdl, Args... args) {
  ThreadState& ts = tls();
  ts.launches++;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (allow_pdl && ts.pdl) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// The reference's void entry points have no error channel: its Rust wrappers assert the shapes and panic.  A void
// entry here that is handed a shape its kernel does not cover fails the same way -- loudly -- instead of returning
// with the output untouched.
[[noreturn]] inline void unsupported(const char* fn, const char* why) {
  fprintf(stderr, "pegainfer_kernels_b200: %s: unsupported arguments (%s)\n", fn, why);
  abort();
}

// ---- device helpers ----
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;");
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16(v); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
// round-trip through bf16 (a reference rounding point kept in a fused kernel)
__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16(v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` is >= 33 floats of shared memory; all threads get the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` reuse
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// 16-byte read-only streaming load (weights / KV: read once, keep out of L1)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes,
                                         uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// fire-and-forget L2 prefetch of a contiguous global range; bytes % 16 == 0, address 16-B aligned
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// ---- tensor-parallel peer-memory communicator (tp_allreduce.cu, gemv.cu) ----
constexpr int kTpMaxWorld = 8;
constexpr int kTpMaxCtas = 64;
constexpr int kTpThreads = 512;

struct TpDev {
  uint8_t* stage[kTpMaxWorld];   // stage[p]: rank p's staging base as mapped in THIS process
  uint32_t* flags[kTpMaxWorld];  // flags[p]: rank p's flag array  [2][kTpMaxCtas][world] + ctl
  int rank, world;
  int64_t slot_bytes;            // bytes per (slot, src) region
  int64_t raw_bytes;             // leading part of a region used for plain rows; the tail holds the two LL areas
  int64_t ll_off, ll_bytes;      // LL lines of the standalone all-reduce kernel (PK_TP_PROTO=ll)
  int64_t gll_off, gll_bytes;    // 8-byte LL lines of the GEMV-fused all-reduce (gemv.cu epi 3)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_v2(void* p, uint32_t x, uint32_t y) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_v2(const void* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace pk

struct pk_tp_comm {
  pk::TpDev d;
  int64_t staging_bytes;
};
