// Element-wise ops and embedding gathers.  Replaces csrc/elementwise.cu and the SiLU kernel of
// csrc/fused_proj.cu of the reference; same arithmetic (fp32 math, rounding points cited).
// HBM-bound: 16-byte vector accesses when shapes allow, grid sized to the SM count.
#include "common.cuh"

namespace pk {

__device__ __forceinline__ float silu(float g) { return g / (1.0f + expf(-g)); }

// out = bf16(f32(a) + f32(b))                       (csrc/elementwise.cu:8-20)
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                           bf16* __restrict__ out, int n, int vec_ok) {
  pdl_wait();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  if (vec_ok) {
    const int nv = n >> 3;
    for (int i = tid; i < nv; i += nthr) {
      uint4 x = reinterpret_cast<const uint4*>(a)[i], y = reinterpret_cast<const uint4*>(b)[i], r;
      r.x = pack_bf16(bf16_lo(x.x) + bf16_lo(y.x), bf16_hi(x.x) + bf16_hi(y.x));
      r.y = pack_bf16(bf16_lo(x.y) + bf16_lo(y.y), bf16_hi(x.y) + bf16_hi(y.y));
      r.z = pack_bf16(bf16_lo(x.z) + bf16_lo(y.z), bf16_hi(x.z) + bf16_hi(y.z));
      r.w = pack_bf16(bf16_lo(x.w) + bf16_lo(y.w), bf16_hi(x.w) + bf16_hi(y.w));
      reinterpret_cast<uint4*>(out)[i] = r;
    }
    for (int i = (nv << 3) + tid; i < n; i += nthr) out[i] = f2bf(bf2f(a[i]) + bf2f(b[i]));
  } else {
    for (int i = tid; i < n; i += nthr) out[i] = f2bf(bf2f(a[i]) + bf2f(b[i]));
  }
}

// out = bf16(f32(bf16(silu(g))) * u)                 (csrc/elementwise.cu:27-42)
__global__ void silu_mul_kernel(const bf16* __restrict__ gate, const bf16* __restrict__ up,
                                bf16* __restrict__ out, int n) {
  pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = f2bf(round_bf16(silu(bf2f(gate[i]))) * bf2f(up[i]));
}

// out[t, i] = bf16(silu(gate_up[t, i]) * gate_up[t, I + i])   (csrc/fused_proj.cu:44-63)
__global__ void silu_mul_fused_kernel(const bf16* __restrict__ gate_up, bf16* __restrict__ out,
                                      int inter, int bs, int vec_ok) {
  pdl_wait();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  if (vec_ok) {
    const int per = inter >> 3, total = per * bs;
    for (int i = tid; i < total; i += nthr) {
      const int t = i / per, c = i - t * per;
      const uint4 g = reinterpret_cast<const uint4*>(gate_up + (size_t)t * 2 * inter)[c];
      const uint4 u = reinterpret_cast<const uint4*>(gate_up + (size_t)t * 2 * inter + inter)[c];
      uint4 r;
      r.x = pack_bf16(silu(bf16_lo(g.x)) * bf16_lo(u.x), silu(bf16_hi(g.x)) * bf16_hi(u.x));
      r.y = pack_bf16(silu(bf16_lo(g.y)) * bf16_lo(u.y), silu(bf16_hi(g.y)) * bf16_hi(u.y));
      r.z = pack_bf16(silu(bf16_lo(g.z)) * bf16_lo(u.z), silu(bf16_hi(g.z)) * bf16_hi(u.z));
      r.w = pack_bf16(silu(bf16_lo(g.w)) * bf16_lo(u.w), silu(bf16_hi(g.w)) * bf16_hi(u.w));
      reinterpret_cast<uint4*>(out + (size_t)t * inter)[c] = r;
    }
  } else {
    const int total = inter * bs;
    for (int i = tid; i < total; i += nthr) {
      const int t = i / inter, c = i - t * inter;
      const float g = bf2f(gate_up[(size_t)t * 2 * inter + c]);
      const float u = bf2f(gate_up[(size_t)t * 2 * inter + inter + c]);
      out[i] = f2bf(silu(g) * u);
    }
  }
}

// out[t, :] = embed[ids[t] - vocab_start, :] (zeros outside the shard)
// (csrc/elementwise.cu:49-112).  One warp-group of 16-B copies per row chunk.
__global__ void embedding_kernel(const bf16* __restrict__ embed, const uint32_t* __restrict__ ids,
                                 bf16* __restrict__ out, int hidden, int seq_len,
                                 uint32_t vocab_start, uint32_t part_vocab, int vec_ok) {
  pdl_wait();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  if (vec_ok) {
    const int per = hidden >> 3, total = per * seq_len;
    for (int i = tid; i < total; i += nthr) {
      const int t = i / per, c = i - t * per;
      const uint32_t id = __ldg(ids + t);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (id >= vocab_start && id - vocab_start < part_vocab)
        v = reinterpret_cast<const uint4*>(embed + (size_t)(id - vocab_start) * hidden)[c];
      reinterpret_cast<uint4*>(out + (size_t)t * hidden)[c] = v;
    }
  } else {
    const int total = hidden * seq_len;
    for (int i = tid; i < total; i += nthr) {
      const int t = i / hidden, c = i - t * hidden;
      const uint32_t id = __ldg(ids + t);
      bf16 v = f2bf(0.f);
      if (id >= vocab_start && id - vocab_start < part_vocab)
        v = embed[(size_t)(id - vocab_start) * hidden + c];
      out[i] = v;
    }
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int grid_for(int64_t work_items, int block) {
  int64_t g = (work_items + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pk

using namespace pk;

extern "C" {

pk_curesult add_cuda(const pk_bf16* a, const pk_bf16* b, pk_bf16* out, int n, pk_stream stream) {
  if (n <= 0) return (pk_curesult)cudaGetLastError();
  const int vec = (aligned16(a) && aligned16(b) && aligned16(out)) ? 1 : 0;
  launch(add_kernel, dim3(grid_for(vec ? (n + 7) / 8 : n, 256)), dim3(256), 0, stream, true,
         (const bf16*)a, (const bf16*)b, (bf16*)out, n, vec);
  return (pk_curesult)cudaGetLastError();
}

pk_curesult silu_mul_triton_aot_cuda(const pk_bf16* gate, const pk_bf16* up, pk_bf16* out, int n,
                                     pk_stream stream) {
  if (n <= 0) return (pk_curesult)cudaGetLastError();
  launch(silu_mul_kernel, dim3(grid_for(n, 256)), dim3(256), 0, stream, true, (const bf16*)gate,
         (const bf16*)up, (bf16*)out, n);
  return (pk_curesult)cudaGetLastError();
}

void silu_mul_fused_cuda(const pk_bf16* gate_up, pk_bf16* out, int intermediate_size, int bs,
                         pk_stream stream) {
  if (intermediate_size <= 0 || bs <= 0) return;
  const int vec = (intermediate_size % 8 == 0 && aligned16(gate_up) && aligned16(out)) ? 1 : 0;
  const int64_t items = (int64_t)intermediate_size * bs / (vec ? 8 : 1);
  launch(silu_mul_fused_kernel, dim3(grid_for(items, 256)), dim3(256), 0, stream, true,
         (const bf16*)gate_up, (bf16*)out, intermediate_size, bs, vec);
}

static pk_curesult embedding_impl(const pk_bf16* embed, const uint32_t* ids, pk_bf16* out,
                                  int hidden, int seq_len, uint32_t vstart, uint32_t vpart,
                                  pk_stream stream) {
  if (hidden <= 0 || seq_len <= 0) return (pk_curesult)cudaGetLastError();
  const int vec = (hidden % 8 == 0 && aligned16(embed) && aligned16(out)) ? 1 : 0;
  const int64_t items = (int64_t)hidden * seq_len / (vec ? 8 : 1);
  launch(embedding_kernel, dim3(grid_for(items, 256)), dim3(256), 0, stream, true,
         (const bf16*)embed, ids, (bf16*)out, hidden, seq_len, vstart, vpart, vec);
  return (pk_curesult)cudaGetLastError();
}

pk_curesult embedding_batched_cuda(const pk_bf16* embed, const uint32_t* token_ids, pk_bf16* out,
                                   int hidden_size, int seq_len, pk_stream stream) {
  return embedding_impl(embed, token_ids, out, hidden_size, seq_len, 0u, 0xffffffffu, stream);
}

pk_curesult embedding_decode_cuda(const pk_bf16* embed, const uint32_t* token_id, pk_bf16* out,
                                  int hidden_size, pk_stream stream) {
  return embedding_impl(embed, token_id, out, hidden_size, 1, 0u, 0xffffffffu, stream);
}

pk_curesult embedding_batched_vocab_shard_cuda(const pk_bf16* embed, const uint32_t* token_ids,
                                               pk_bf16* out, int hidden_size, int seq_len,
                                               uint32_t vocab_start, uint32_t part_vocab_size,
                                               pk_stream stream) {
  return embedding_impl(embed, token_ids, out, hidden_size, seq_len, vocab_start, part_vocab_size,
                        stream);
}

}  // extern "C"
