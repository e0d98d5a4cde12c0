// RMSNorm and fused residual-add + RMSNorm.  Replaces csrc/flashinfer_norm.cu:49-105 (FlashInfer
// norm.cuh:36-111,386-477) of the reference with the same rounding points:
//   rms_norm        : out = bf16(x * rsqrt(mean(x^2)+eps) * w), fp32 throughout, ONE rounding
//   fused_add_norm  : x = f32(hidden)+f32(residual); hidden = bf16(x);
//                     out = bf16(x * rsqrt(mean(x^2)+eps) * w) on the UNROUNDED x
// Unlike the reference there is no residual->out staging memcpy.  One CTA per token; the row is
// read from HBM once (16-B vectors) and parked in shared memory as fp32 between the two passes.
#include "common.cuh"

namespace pk {

template <bool kFusedAdd>
__global__ void rms_norm_kernel(const bf16* __restrict__ x, bf16* __restrict__ hidden,
                                const bf16* __restrict__ residual, const bf16* __restrict__ w,
                                bf16* __restrict__ out, int dim, float eps, int vec_ok) {
  extern __shared__ float srow[];  // dim floats + 33 for the reduction
  float* red = srow + dim;
  const size_t base = (size_t)blockIdx.x * dim;
  const bf16* in = kFusedAdd ? hidden + base : x + base;
  pdl_wait();
  float ss = 0.f;
  if (vec_ok) {
    const int nv = dim >> 3;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      const uint4 a = reinterpret_cast<const uint4*>(in)[i];
      float v[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y),
                    bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
      if (kFusedAdd) {
        const uint4 r = reinterpret_cast<const uint4*>(residual + base)[i];
        v[0] += bf16_lo(r.x); v[1] += bf16_hi(r.x); v[2] += bf16_lo(r.y); v[3] += bf16_hi(r.y);
        v[4] += bf16_lo(r.z); v[5] += bf16_hi(r.z); v[6] += bf16_lo(r.w); v[7] += bf16_hi(r.w);
        uint4 h;
        h.x = pack_bf16(v[0], v[1]); h.y = pack_bf16(v[2], v[3]);
        h.z = pack_bf16(v[4], v[5]); h.w = pack_bf16(v[6], v[7]);
        reinterpret_cast<uint4*>(hidden + base)[i] = h;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ss = fmaf(v[j], v[j], ss);
        srow[i * 8 + j] = v[j];
      }
    }
  } else {
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
      float v = bf2f(in[i]);
      if (kFusedAdd) {
        v += bf2f(residual[base + i]);
        hidden[base + i] = f2bf(v);
      }
      ss = fmaf(v, v, ss);
      srow[i] = v;
    }
  }
  const float total = block_sum(ss, red);
  const float r = rsqrtf(total / (float)dim + eps);
  if (vec_ok) {
    const int nv = dim >> 3;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      const uint4 g = reinterpret_cast<const uint4*>(w)[i];
      const float* v = srow + i * 8;
      uint4 o;
      o.x = pack_bf16(v[0] * r * bf16_lo(g.x), v[1] * r * bf16_hi(g.x));
      o.y = pack_bf16(v[2] * r * bf16_lo(g.y), v[3] * r * bf16_hi(g.y));
      o.z = pack_bf16(v[4] * r * bf16_lo(g.z), v[5] * r * bf16_hi(g.z));
      o.w = pack_bf16(v[6] * r * bf16_lo(g.w), v[7] * r * bf16_hi(g.w));
      reinterpret_cast<uint4*>(out + base)[i] = o;
    }
  } else {
    for (int i = threadIdx.x; i < dim; i += blockDim.x) out[base + i] = f2bf(srow[i] * r * bf2f(w[i]));
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <bool kFusedAdd>
static void launch_norm(const pk_bf16* x, pk_bf16* hidden, const pk_bf16* residual,
                        const pk_bf16* w, pk_bf16* out, int dim, int rows, float eps,
                        pk_stream stream) {
  if (dim <= 0 || rows <= 0) return;
  const bool vec = dim % 8 == 0 && al16(x) && al16(hidden) && al16(residual) && al16(w) && al16(out);
  int threads = vec ? (dim / 8) : dim;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  const size_t smem = sizeof(float) * ((size_t)dim + 40);
  auto kern = rms_norm_kernel<kFusedAdd>;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch(kern, dim3(rows), dim3(threads), smem, stream, true, (const bf16*)x, (bf16*)hidden,
         (const bf16*)residual, (const bf16*)w, (bf16*)out, dim, eps, vec ? 1 : 0);
}

}  // namespace pk

extern "C" {

void rms_norm_cuda(const pk_bf16* x, const pk_bf16* weight, pk_bf16* out, int n, float eps,
                   pk_stream stream) {
  pk::launch_norm<false>(x, nullptr, nullptr, weight, out, n, 1, eps, stream);
}
void rms_norm_batched_cuda(const pk_bf16* x, const pk_bf16* weight, pk_bf16* out, int hidden_dim,
                           int seq_len, float eps, pk_stream stream) {
  pk::launch_norm<false>(x, nullptr, nullptr, weight, out, hidden_dim, seq_len, eps, stream);
}
void fused_add_rms_norm_cuda(pk_bf16* hidden, const pk_bf16* residual, const pk_bf16* weight,
                             pk_bf16* out, int n, float eps, pk_stream stream) {
  pk::launch_norm<true>(nullptr, hidden, residual, weight, out, n, 1, eps, stream);
}
void fused_add_rms_norm_batched_cuda(pk_bf16* hidden, const pk_bf16* residual,
                                     const pk_bf16* weight, pk_bf16* out, int hidden_dim,
                                     int batch_size, float eps, pk_stream stream) {
  pk::launch_norm<true>(nullptr, hidden, residual, weight, out, hidden_dim, batch_size, eps,
                        stream);
}

}  // extern "C"
