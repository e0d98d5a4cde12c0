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

}  // extern "C"
namespace pk {
// small-row path of flashinfer_top1_cuda: the single-CTA arg-max writes only the index; the winner's value is part of
// the contract too (the vocab-sharded TP exchange compares it across ranks)
__global__ void top1_value_gather_kernel(const bf16* logits, const int* idx, bf16* value) {
  pdl_wait();
  if (threadIdx.x == 0) value[0] = logits[idx[0]];
}
}  // namespace pk
extern "C" {

void flashinfer_top1_cuda(const pk_bf16* logits, pk_bf16* top1_value_scratch,
                          uint8_t* row_states_scratch, int* output, int vocab_size,
                          pk_stream stream) {
  if (vocab_size <= 0) return;
  const bool al = (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  if (!al || row_states_scratch == nullptr || vocab_size < 8192) {
    argmax_cuda(logits, output, vocab_size, stream);
    if (top1_value_scratch)
      pk::launch(pk::top1_value_gather_kernel, dim3(1), dim3(32), 0, stream, true, (const pk::bf16*)logits, (const int*)output,
                 (pk::bf16*)top1_value_scratch);
    return;
  }
  int grid = (vocab_size + 2047) / 2048;
  if (grid > 128) grid = 128;
  pk::launch(pk::top1_kernel, dim3(grid), dim3(256), 0, stream, true, (const pk::bf16*)logits,
             (pk::bf16*)top1_value_scratch, reinterpret_cast<pk::Top1State*>(row_states_scratch),
             output, vocab_size);
}

}  // extern "C"

// =====================================================================================================
// Non-greedy sampling: temperature softmax + top-k / top-p filtering + multinomial draw.
// Replaces csrc/flashinfer_sampling.cu:13-110 (logits_to_probs_kernel + FlashInfer
// {TopKTopP,TopK,TopP,}SamplingFromProb).  Distribution semantics are the reference's: probs =
// softmax(bf16 logits * inv_temperature) in fp32; a token is eligible iff it is among the top_k most probable
// (top_k > 0) AND inside the smallest prefix of the probability-sorted vocabulary whose mass reaches top_p
// (top_p < 1), both evaluated on the unfiltered probs ("joint" filtering); the draw is from the renormalised
// eligible set.  The RANDOM STREAM is not FlashInfer's Philox: u = splitmix64(seed) (documented in DESIGN.md);
// the reference's own test only pins greedy-equivalent cases and "token in range" (ops/tests.rs:228-305).
// One CTA; thresholds by bisection on the fp32 bit pattern (probabilities are >= 0, so bits are monotone).
namespace pk {

constexpr int kSampleThreads = 1024;

__device__ __forceinline__ float block_reduce_sum_f(float v, float* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  float t = threadIdx.x < 32 ? sm[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (lane == 0) sm[32] = t;
  }
  __syncthreads();
  return sm[32];
}
__device__ __forceinline__ float block_reduce_max_f(float v, float* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  float t = threadIdx.x < 32 ? sm[threadIdx.x] : -INFINITY;
  if (w == 0) {
    t = warp_max(t);
    if (lane == 0) sm[32] = t;
  }
  __syncthreads();
  return sm[32];
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(kSampleThreads) sample_kernel(const bf16* __restrict__ logits, float* __restrict__ probs,
                                                                uint8_t* __restrict__ valid, int* __restrict__ out, int n,
                                                                float inv_temperature, int top_k, float top_p,
                                                                uint64_t seed) {
  __shared__ float sm[40];
  __shared__ float s_scan[kSampleThreads];
  __shared__ int s_pick;
  const int tid = threadIdx.x;
  pdl_wait();
  // ---- softmax(logits * inv_temperature), fp32 (logits_to_probs_kernel) ----
  float mx = -INFINITY;
  for (int i = tid; i < n; i += kSampleThreads) {
    const float v = bf2f(logits[i]) * inv_temperature;
    probs[i] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_reduce_max_f(mx, sm);
  float sum = 0.f;
  for (int i = tid; i < n; i += kSampleThreads) {
    const float v = expf(probs[i] - mx);
    probs[i] = v;
    sum += v;
  }
  sum = block_reduce_sum_f(sum, sm);
  const float inv = 1.0f / sum;
  for (int i = tid; i < n; i += kSampleThreads) probs[i] *= inv;
  __syncthreads();
  // ---- eligibility threshold: largest t with count(p >= t) >= top_k and mass(p >= t) >= top_p ----
  uint32_t thr_bits = 0;  // p >= +0 : everything
  if (top_k > 0 && top_k < n) {
    uint32_t lo = 0, hi = 0x3f800001u;  // invariant: count(p >= lo) >= k, count(p >= hi) < k
    while (hi - lo > 1) {
      const uint32_t mid = lo + (hi - lo) / 2;
      float c = 0.f;
      for (int i = tid; i < n; i += kSampleThreads) c += (__float_as_uint(probs[i]) >= mid) ? 1.f : 0.f;
      c = block_reduce_sum_f(c, sm);
      if (c >= (float)top_k) lo = mid; else hi = mid;
    }
    thr_bits = lo;
  }
  if (top_p < 1.0f) {
    uint32_t lo = 0, hi = 0x3f800001u;  // invariant: mass(p >= lo) >= top_p, mass(p >= hi) < top_p
    while (hi - lo > 1) {
      const uint32_t mid = lo + (hi - lo) / 2;
      float msum = 0.f;
      for (int i = tid; i < n; i += kSampleThreads) {
        const float pv = probs[i];
        msum += (__float_as_uint(pv) >= mid) ? pv : 0.f;
      }
      msum = block_reduce_sum_f(msum, sm);
      if (msum >= top_p) lo = mid; else hi = mid;
    }
    thr_bits = max(thr_bits, lo);
  }
  // ---- draw from the renormalised eligible set by inverse CDF in index order ----
  float total = 0.f;
  for (int i = tid; i < n; i += kSampleThreads) {
    const float pv = probs[i];
    total += (__float_as_uint(pv) >= thr_bits) ? pv : 0.f;
  }
  total = block_reduce_sum_f(total, sm);
  const float u = (float)(splitmix64(seed) >> 40) * (1.0f / 16777216.0f) * total;
  if (tid == 0) s_pick = 0x7fffffff;
  __syncthreads();
  float running = 0.f;
  int last_ok = -1;
  for (int base = 0; base < n; base += kSampleThreads) {
    const int i = base + tid;
    const float pv = (i < n && __float_as_uint(probs[i]) >= thr_bits) ? probs[i] : 0.f;
    if (pv > 0.f) last_ok = i;
    // inclusive block scan (Hillis-Steele over shared memory)
    s_scan[tid] = pv;
    __syncthreads();
    for (int off = 1; off < kSampleThreads; off <<= 1) {
      const float add = tid >= off ? s_scan[tid - off] : 0.f;
      __syncthreads();
      s_scan[tid] += add;
      __syncthreads();
    }
    const float incl = running + s_scan[tid];
    // first eligible index whose inclusive prefix exceeds u: prefixes are non-decreasing in i, so the minimum
    // over all candidates of the first block that has one is the inverse-CDF pick (deterministic)
    if (pv > 0.f && incl > u) atomicMin(&s_pick, i);
    running += s_scan[kSampleThreads - 1];
    __syncthreads();
    if (s_pick != 0x7fffffff) break;
  }
  // rounding at the very end of the CDF: fall back to the last eligible token
  int lo_all = last_ok;
  for (int o = 16; o > 0; o >>= 1) lo_all = max(lo_all, __shfl_xor_sync(0xffffffffu, lo_all, o));
  __shared__ int s_last[32];
  if ((tid & 31) == 0) s_last[tid >> 5] = lo_all;
  __syncthreads();
  if (tid == 0) {
    int pick = s_pick == 0x7fffffff ? -1 : s_pick;
    if (pick < 0) {
      for (int w = 0; w < kSampleThreads / 32; ++w) pick = max(pick, s_last[w]);
      if (pick < 0) pick = 0;
    }
    out[0] = pick;
    if (valid) valid[0] = 1;
  }
}

}  // namespace pk

extern "C" void gpu_sample_flashinfer_cuda(const pk_bf16* logits, float* probs_scratch, uint8_t* valid_scratch,
                                           int* output, int vocab_size, float inv_temperature, int top_k,
                                           float top_p, uint64_t seed, pk_stream stream) {
  if (vocab_size <= 0) return;
  pk::launch(pk::sample_kernel, dim3(1), dim3(pk::kSampleThreads), 0, stream, true, (const pk::bf16*)logits,
             probs_scratch, valid_scratch, output, vocab_size, inv_temperature, top_k, top_p, seed);
}
