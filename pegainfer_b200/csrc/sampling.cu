// Greedy sampling: arg-max over bf16 logits.  Replaces csrc/argmax.cu and
// csrc/flashinfer_top1.cu (FlashInfer RadixTopKMultiCTA for vocab 151,936) of the reference.
// Tie rule: LOWEST index wins (csrc/argmax.cu:18); the reference's radix path leaves it undefined.
//
// flashinfer_top1_cuda is a two-stage multi-CTA reduction that fits the 300 KB logits row:
// stage 1: every CTA reduces a 16-B-vectorised slice to (value, index) and publishes it in
// row_states_scratch; the last CTA to finish (atomic ticket) reduces the partials, writes the
// token id (+ bf16 value) and resets the ticket so the call is CUDA-graph replayable.
#include "common.cuh"

namespace pk {

struct MaxIdx {
  float v;
  int i;
};
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) {
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ MaxIdx warp_best(MaxIdx m) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxIdx t;
    t.v = __shfl_xor_sync(0xffffffffu, m.v, o);
    t.i = __shfl_xor_sync(0xffffffffu, m.i, o);
    m = better(m, t);
  }
  return m;
}
__device__ __forceinline__ MaxIdx block_best(MaxIdx m, MaxIdx* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  m = warp_best(m);
  if (lane == 0) sm[w] = m;
  __syncthreads();
  if (w == 0) {
    MaxIdx t = lane < nw ? sm[lane] : MaxIdx{-INFINITY, 0x7fffffff};
    t = warp_best(t);
    if (lane == 0) sm[0] = t;
  }
  __syncthreads();
  return sm[0];
}
__device__ __forceinline__ MaxIdx scan_range(const bf16* __restrict__ x, int lo, int hi, int tid,
                                             int nthr) {
  MaxIdx m{-INFINITY, 0x7fffffff};
  // NaNs never win (comparisons false), like the reference's `val > local_max`.
  const int lo_al = min(hi, (lo + 7) & ~7);
  for (int i = lo + tid; i < lo_al; i += nthr) m = better(m, MaxIdx{bf2f(x[i]), i});
  const int nv = (hi - lo_al) >> 3;
  for (int v = tid; v < nv; v += nthr) {
    const int i0 = lo_al + v * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(x + i0);
    const float f[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y),
                        bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) m = better(m, MaxIdx{f[j], i0 + j});
  }
  for (int i = lo_al + nv * 8 + tid; i < hi; i += nthr) m = better(m, MaxIdx{bf2f(x[i]), i});
  return m;
}

__global__ void argmax_kernel(const bf16* __restrict__ x, int* __restrict__ out, int n) {
  __shared__ MaxIdx sm[32];
  pdl_wait();
  const bool al = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  MaxIdx m{-INFINITY, 0x7fffffff};
  if (al) {
    m = scan_range(x, 0, n, threadIdx.x, blockDim.x);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = better(m, MaxIdx{bf2f(x[i]), i});
  }
  m = block_best(m, sm);
  if (threadIdx.x == 0) out[0] = m.i == 0x7fffffff ? 0 : m.i;
}

struct Top1State {  // lives in row_states_scratch
  unsigned int ticket;
  unsigned int pad[3];
  MaxIdx part[256];
};

__global__ void top1_kernel(const bf16* __restrict__ x, bf16* __restrict__ top_val,
                            Top1State* __restrict__ st, int* __restrict__ out, int n) {
  __shared__ MaxIdx sm[32];
  __shared__ bool last;
  pdl_wait();
  const int per = (((n + gridDim.x - 1) / gridDim.x) + 7) & ~7;
  const int lo = min(n, (int)blockIdx.x * per), hi = min(n, lo + per);
  MaxIdx m = scan_range(x, lo, hi, threadIdx.x, blockDim.x);
  m = block_best(m, sm);
  if (threadIdx.x == 0) {
    st->part[blockIdx.x] = m;
    __threadfence();
    const unsigned t = atomicAdd(&st->ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  MaxIdx r = threadIdx.x < gridDim.x ? st->part[threadIdx.x] : MaxIdx{-INFINITY, 0x7fffffff};
  __syncthreads();
  r = block_best(r, sm);
  if (threadIdx.x == 0) {
    out[0] = r.i == 0x7fffffff ? 0 : r.i;
    if (top_val) top_val[0] = f2bf(r.v);
    st->ticket = 0;  // replayable
  }
}

}  // namespace pk

extern "C" {

void argmax_cuda(const pk_bf16* x, int* out, int n, pk_stream stream) {
  if (n <= 0) return;
  pk::launch(pk::argmax_kernel, dim3(1), dim3(1024), 0, stream, true, (const pk::bf16*)x, out, n);
}

void flashinfer_top1_cuda(const pk_bf16* logits, pk_bf16* top1_value_scratch,
                          uint8_t* row_states_scratch, int* output, int vocab_size,
                          pk_stream stream) {
  if (vocab_size <= 0) return;
  const bool al = (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  if (!al || row_states_scratch == nullptr || vocab_size < 8192) {
    argmax_cuda(logits, output, vocab_size, stream);
    return;
  }
  int grid = (vocab_size + 2047) / 2048;
  if (grid > 128) grid = 128;
  pk::launch(pk::top1_kernel, dim3(grid), dim3(256), 0, stream, true, (const pk::bf16*)logits,
             (pk::bf16*)top1_value_scratch, reinterpret_cast<pk::Top1State*>(row_states_scratch),
             output, vocab_size);
}

}  // extern "C"
