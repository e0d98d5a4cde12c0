// GQA decode attention over the paged KV cache (one query token per request).
// Replaces csrc/paged_attention.cu:77-230 (FlashInfer BatchDecodeWithPagedKVCache + the
// VariableLengthMergeStates merge kernel) of the reference, and adds the fused
// QK-norm + RoPE + KV-append + split-KV + merge single-launch variant.
//
// Arithmetic follows flashinfer/attention/decode.cuh:62-145 and state.cuh:30-80: q, k, v cast to
// fp32; s = (q.k) * sm_scale * log2(e); running max / denominator with ex2.approx; o = sum p*v in
// fp32; o/d rounded ONCE to bf16.  The ABI split-KV entry keeps the reference's rounding of the
// per-chunk NORMALISED partial to bf16 (tmp_v) and base-2 LSE (tmp_s); the internal split used by
// the non-partition / fused entries keeps fp32 partials (documented in DESIGN.md).
//
// HBM-bound: K+V bytes per (token, kv head) = 512; algorithmic bytes per call = 4096 * kv_len.
// Layout: one CTA per (kv chunk, kv head, request); the 4 warps stride over the chunk's tokens,
// each half-warp owning one token row per step (16 lanes x 16 B = one 256-B K row, fully
// coalesced), the GQA group's q heads held in registers so each K/V row is read once for all of
// them; 8 rows of K and 8 of V are in flight per lane-group per iteration.  The grid is sized so
// chunks x kv heads x requests covers the 148 SMs; the last CTA of a (request, kv head) merges.
#include <cstdlib>
#include <cstring>

#include "decode_attention_cluster.cuh"

namespace pk {

constexpr int HD = 128;
constexpr int kAttWarps = 4;
constexpr int kAttThreads = kAttWarps * 32;
constexpr int kTokPerStep = kAttWarps * 2;
constexpr int kUnroll = 8;   // 64 tokens (4 pages) per CTA round; all K/V rows of a round in flight
constexpr int kPartStride = HD + 2;  // o[128], m, d

struct DecodeAttnArgs {
  const bf16* q;
  bf16* out;
  bf16* kv;
  int64_t k_off, v_off;
  const int* page_indices;
  const int* page_indptr;
  const int* last_page_len;
  int mode;  // 0: implicit chunking + fp32 partials + ticket merge; 1: explicit slot plan (ABI split)
  const int* request_indices;
  const int* kv_tile_indices;
  const int* kv_chunk_size;
  const uint8_t* valid_mask;
  bf16* tmp_v;
  float* tmp_s;
  float* partial;
  int* counters;
  int min_chunk, max_chunks;
  int nq, nkv, page_size;
  int64_t stride_page;
  float sm_scale_log2;
  // fused new-token path
  int fused;
  const bf16* k_new;
  const bf16* v_new;
  const int* positions;
  const bf16* qw;
  const bf16* kw;
  const bf16* cosc;
  const bf16* sinc;
  float eps;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void unpack8(const uint4& a, float* f) {
  f[0] = bf16_lo(a.x); f[1] = bf16_hi(a.x); f[2] = bf16_lo(a.y); f[3] = bf16_hi(a.y);
  f[4] = bf16_lo(a.z); f[5] = bf16_hi(a.z); f[6] = bf16_lo(a.w); f[7] = bf16_hi(a.w);
}

// Warp-level QK RMSNorm + RoPE of one 128-wide head (same math as qk_norm_rope.cu); lane owns
// elements 4*lane..4*lane+3; result (bf16-rounded) written to dst[128] in shared memory.
__device__ __forceinline__ void warp_norm_rope(const bf16* __restrict__ src,
                                               const bf16* __restrict__ w,
                                               const bf16* __restrict__ cosc,
                                               const bf16* __restrict__ sinc, int pos, float eps,
                                               bf16* dst, int lane) {
  const uint2 raw = reinterpret_cast<const uint2*>(src)[lane];
  const float v[4] = {bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y)};
  float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)HD + eps);
  const uint2 wr = reinterpret_cast<const uint2*>(w)[lane];
  const float wv[4] = {bf16_lo(wr.x), bf16_hi(wr.x), bf16_lo(wr.y), bf16_hi(wr.y)};
  const int ci = (lane & 15) * 4;
  const uint2 cr = reinterpret_cast<const uint2*>(cosc + (size_t)pos * HD + ci)[0];
  const uint2 sr = reinterpret_cast<const uint2*>(sinc + (size_t)pos * HD + ci)[0];
  const float c[4] = {bf16_lo(cr.x), bf16_hi(cr.x), bf16_lo(cr.y), bf16_hi(cr.y)};
  const float s[4] = {bf16_lo(sr.x), bf16_hi(sr.x), bf16_lo(sr.y), bf16_hi(sr.y)};
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float t = round_bf16(round_bf16(v[j] * inv) * wv[j]);
    const float other = __shfl_xor_sync(0xffffffffu, t, 16);
    o[j] = lane < 16 ? t * c[j] - other * s[j] : other * s[j] + t * c[j];
  }
  uint2 res;
  res.x = pack_bf16(o[0], o[1]);
  res.y = pack_bf16(o[2], o[3]);
  reinterpret_cast<uint2*>(dst)[lane] = res;
}

template <int GROUP>
__global__ void __launch_bounds__(kAttThreads)
decode_attention_kernel(const DecodeAttnArgs a) {
  __shared__ __align__(16) float st_o[kTokPerStep + 1][GROUP][HD];
  __shared__ float st_m[kTokPerStep + 1][GROUP], st_d[kTokPerStep + 1][GROUP];
  __shared__ __align__(16) bf16 q_s[GROUP][HD];
  __shared__ __align__(16) bf16 k_s[HD], v_s[HD];
  __shared__ int s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  const int kvh = blockIdx.y;
  int b, chunk_idx, out_slot;
  if (a.mode == 1) {
    out_slot = blockIdx.x;
    if (a.valid_mask && !a.valid_mask[out_slot]) return;
    b = a.request_indices[out_slot];
    chunk_idx = a.kv_tile_indices[out_slot];
  } else {
    out_slot = blockIdx.z;
    b = a.request_indices ? a.request_indices[out_slot] : out_slot;
    chunk_idx = blockIdx.x;
  }
  const int npages = a.page_indptr[b + 1] - a.page_indptr[b];
  const int len = npages <= 0 ? 0 : (npages - 1) * a.page_size + a.last_page_len[b];
  int chunk, nchunks;
  if (a.mode == 1) {
    chunk = a.kv_chunk_size[0];
    nchunks = 0;  // unused
  } else {
    chunk = max(a.min_chunk, (len + a.max_chunks - 1) / a.max_chunks);
    chunk = (chunk + 15) & ~15;
    nchunks = max(1, (len + chunk - 1) / chunk);
    if (chunk_idx >= nchunks) return;
  }
  const int lo = chunk_idx * chunk;
  const int hi = min(len, lo + chunk);
  const int* pages = a.page_indices + a.page_indptr[b];

  // Let the next kernel (the o_proj GEMV) become resident and prefetch its weights while we run.
  pdl_launch_dependents();

  // Fused path: the cached K/V rows of round 0 do not depend on the previous kernel (the step's own
  // token is injected from shared memory, never read from the cache), so request them BEFORE
  // griddepcontrol.wait: their HBM latency overlaps the tail of the qkv GEMV.
  constexpr int kRoundTokens = kTokPerStep * kUnroll;
  uint4 kr0[kUnroll], vr0[kUnroll];
  bool ok0[kUnroll];
  const bool early = a.fused != 0 && a.page_size == 16;
  if (early) {
    const int pos_e = a.positions[b];
    const int64_t kvh_off = (int64_t)kvh * HD + l16 * 8;
    int pg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pg[i] = (lo + 16 * i < hi) ? __ldg(pages + (lo >> 4) + i) : 0;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int t = lo + warp * 2 + half + u * kTokPerStep;
      ok0[u] = t < hi && t != pos_e;
      kr0[u] = make_uint4(0, 0, 0, 0);
      vr0[u] = make_uint4(0, 0, 0, 0);
      if (ok0[u]) {
        const int64_t off = (int64_t)pg[u >> 1] * a.stride_page + (int64_t)(t & 15) * a.nkv * HD + kvh_off;
        kr0[u] = ldg_stream(a.kv + a.k_off + off);
        vr0[u] = ldg_stream(a.kv + a.v_off + off);
      }
    }
  }

  pdl_wait();

  // ---- query heads of this kv head -> registers (fp32) ----
  float qf[GROUP][8];
  bool inject_new = false;  // fused: this CTA attends to the step's new token from smem
  if (a.fused) {
    const int pos = a.positions[b];
    for (int h = warp; h < GROUP; h += kAttWarps)
      warp_norm_rope(a.q + ((size_t)b * a.nq + kvh * GROUP + h) * HD, a.qw, a.cosc, a.sinc, pos,
                     a.eps, q_s[h], lane);
    inject_new = (pos >= lo && pos < hi);
    if (inject_new && warp == 0) {
      warp_norm_rope(a.k_new + ((size_t)b * a.nkv + kvh) * HD, a.kw, a.cosc, a.sinc, pos, a.eps,
                     k_s, lane);
      reinterpret_cast<uint2*>(v_s)[lane] =
          reinterpret_cast<const uint2*>(a.v_new + ((size_t)b * a.nkv + kvh) * HD)[lane];
      __syncwarp();
      // append to the cache (read back by later steps, never by this launch)
      const int page = pages[pos / a.page_size], slot = pos % a.page_size;
      const int64_t dst = (int64_t)page * a.stride_page + ((int64_t)slot * a.nkv + kvh) * HD;
      reinterpret_cast<uint2*>(a.kv + a.k_off + dst)[lane] = reinterpret_cast<uint2*>(k_s)[lane];
      reinterpret_cast<uint2*>(a.kv + a.v_off + dst)[lane] = reinterpret_cast<uint2*>(v_s)[lane];
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < GROUP; ++h) unpack8(reinterpret_cast<const uint4*>(q_s[h])[l16], qf[h]);
  } else {
#pragma unroll
    for (int h = 0; h < GROUP; ++h)
      unpack8(reinterpret_cast<const uint4*>(a.q + ((size_t)b * a.nq + kvh * GROUP + h) * HD)[l16],
              qf[h]);
  }

  float m[GROUP], d[GROUP], o[GROUP][8];
#pragma unroll
  for (int h = 0; h < GROUP; ++h) {
    m[h] = -INFINITY;
    d[h] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[h][j] = 0.f;
  }

  const int new_pos = inject_new ? a.positions[b] : -1;
  const bf16* kbase = a.kv + a.k_off + (int64_t)kvh * HD + l16 * 8;
  const bf16* vbase = a.kv + a.v_off + (int64_t)kvh * HD + l16 * 8;
  // NOTE: the trip count must be warp-uniform (full-mask shuffles inside): iterate on the CTA's round
  // base and let each half-warp mask its own tokens.  A round is 64 tokens = 4 pages: the 4 page ids
  // are fetched first (one latency), then all 8 K and 8 V rows of the half-warp (one latency).
  for (int round = lo; round < hi; round += kRoundTokens) {
    const int base = round + warp * 2 + half;
    uint4 kr[kUnroll], vr[kUnroll];
    bool ok[kUnroll];
    if (early && round == lo) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        kr[u] = kr0[u];
        vr[u] = vr0[u];
        ok[u] = ok0[u];
      }
    } else {
      int pg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pg[i] = (round + 16 * i < hi) ? __ldg(pages + (round >> 4) + i) : 0;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int t = base + u * kTokPerStep;
        ok[u] = t < hi && t != new_pos;
        kr[u] = make_uint4(0, 0, 0, 0);
        vr[u] = make_uint4(0, 0, 0, 0);
        if (ok[u]) {
          // (t - round) / 16 == u / 2 because warp*2+half < 8; page_size is 16 on this path
          const int page = a.page_size == 16 ? pg[u >> 1] : __ldg(pages + t / a.page_size);
          const int slot = t % a.page_size;
          const int64_t off = (int64_t)page * a.stride_page + (int64_t)slot * a.nkv * HD;
          kr[u] = ldg_stream(kbase + off);
          vr[u] = ldg_stream(vbase + off);
        }
      }
    }
    float s[GROUP][kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float kf[8];
      unpack8(kr[u], kf);
#pragma unroll
      for (int h = 0; h < GROUP; ++h) {
        float p = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) p = fmaf(qf[h][j], kf[j], p);
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
        s[h][u] = ok[u] ? p * a.sm_scale_log2 : -INFINITY;
      }
    }
#pragma unroll
    for (int h = 0; h < GROUP; ++h) {
      float mn = m[h];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) mn = fmaxf(mn, s[h][u]);
      if (mn == -INFINITY) {  // nothing valid yet: p = 0 (never let -inf reach the PV FMAs)
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) s[h][u] = 0.f;
        continue;
      }
      const float sc = ex2(m[h] - mn);  // m = -inf -> 0
      m[h] = mn;
      d[h] *= sc;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[h][j] *= sc;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) s[h][u] = ex2(s[h][u] - mn);  // -inf -> 0
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) d[h] += s[h][u];
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float vf[8];
      unpack8(vr[u], vf);
#pragma unroll
      for (int h = 0; h < GROUP; ++h)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[h][j] = fmaf(s[h][u], vf[j], o[h][j]);
    }
  }

  // ---- CTA merge of the 8 half-warp states (+ the injected new token) ----
  const int sid = warp * 2 + half;
#pragma unroll
  for (int h = 0; h < GROUP; ++h) {
    if (l16 == 0) {
      st_m[sid][h] = m[h];
      st_d[sid][h] = d[h];
    }
    float4* dst = reinterpret_cast<float4*>(&st_o[sid][h][l16 * 8]);
    dst[0] = make_float4(o[h][0], o[h][1], o[h][2], o[h][3]);
    dst[1] = make_float4(o[h][4], o[h][5], o[h][6], o[h][7]);
  }
  if (inject_new) {
    for (int h = warp; h < GROUP; h += kAttWarps) {
      // s_new = (q . k_new) * scale; lane owns 4 elements
      const uint2 qr = reinterpret_cast<const uint2*>(q_s[h])[lane];
      const uint2 kr2 = reinterpret_cast<const uint2*>(k_s)[lane];
      float p = bf16_lo(qr.x) * bf16_lo(kr2.x);
      p = fmaf(bf16_hi(qr.x), bf16_hi(kr2.x), p);
      p = fmaf(bf16_lo(qr.y), bf16_lo(kr2.y), p);
      p = fmaf(bf16_hi(qr.y), bf16_hi(kr2.y), p);
      p = warp_sum(p);
      if (lane == 0) {
        st_m[kTokPerStep][h] = p * a.sm_scale_log2;
        st_d[kTokPerStep][h] = 1.f;
      }
      const uint2 vr2 = reinterpret_cast<const uint2*>(v_s)[lane];
      *reinterpret_cast<float4*>(&st_o[kTokPerStep][h][lane * 4]) =
          make_float4(bf16_lo(vr2.x), bf16_hi(vr2.x), bf16_lo(vr2.y), bf16_hi(vr2.y));
    }
  }
  __syncthreads();
  const int nstates = inject_new ? kTokPerStep + 1 : kTokPerStep;
  const int t = threadIdx.x;  // output dim
  float M[GROUP], D[GROUP], O[GROUP];
#pragma unroll
  for (int h = 0; h < GROUP; ++h) {
    float mx = -INFINITY;
    for (int i = 0; i < nstates; ++i) mx = fmaxf(mx, st_m[i][h]);
    float dd = 0.f, oo = 0.f;
    if (mx != -INFINITY) {
      for (int i = 0; i < nstates; ++i) {
        const float w = ex2(st_m[i][h] - mx);
        dd = fmaf(st_d[i][h], w, dd);
        oo = fmaf(st_o[i][h][t], w, oo);
      }
    }
    M[h] = mx; D[h] = dd; O[h] = oo;
  }

  if (a.mode == 1) {
#pragma unroll
    for (int h = 0; h < GROUP; ++h) {
      const int head = kvh * GROUP + h;
      a.tmp_v[((size_t)out_slot * a.nq + head) * HD + t] = f2bf(__fdividef(O[h], D[h]));
      if (t == 0) a.tmp_s[(size_t)out_slot * a.nq + head] = M[h] + lg2(D[h]);
    }
    return;
  }
  if (nchunks == 1) {
#pragma unroll
    for (int h = 0; h < GROUP; ++h)
      a.out[((size_t)out_slot * a.nq + kvh * GROUP + h) * HD + t] = f2bf(__fdividef(O[h], D[h]));
    return;
  }
  // ---- cross-CTA: publish fp32 partial, last arrival merges ----
#pragma unroll
  for (int h = 0; h < GROUP; ++h) {
    float* p = a.partial +
               (((size_t)out_slot * a.max_chunks + chunk_idx) * a.nq + kvh * GROUP + h) * kPartStride;
    p[t] = O[h];
    if (t == 0) {
      p[HD] = M[h];
      p[HD + 1] = D[h];
    }
  }
  __threadfence();
  __syncthreads();
  if (t == 0) {
    const int ticket = atomicAdd(a.counters + out_slot * a.nkv + kvh, 1);
    s_last = (ticket == nchunks - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last CTA: weights of all chunks first (warp-parallel over chunks), then one unrolled pass ----
  float* sw = &st_o[0][0][0];  // reuse: [nchunks][GROUP] weights, then [GROUP] denominators
  float* sD = sw + 64 * GROUP;
  __syncthreads();
  for (int h = warp; h < GROUP; h += kAttWarps) {
    const float* p0 =
        a.partial + (((size_t)out_slot * a.max_chunks) * a.nq + kvh * GROUP + h) * kPartStride;
    const size_t cstride = (size_t)a.nq * kPartStride;
    const int c0 = lane, c1 = lane + 32;
    const float m0 = c0 < nchunks ? __ldcg(p0 + c0 * cstride + HD) : -INFINITY;
    const float m1 = c1 < nchunks ? __ldcg(p0 + c1 * cstride + HD) : -INFINITY;
    const float d0 = c0 < nchunks ? __ldcg(p0 + c0 * cstride + HD + 1) : 0.f;
    const float d1 = c1 < nchunks ? __ldcg(p0 + c1 * cstride + HD + 1) : 0.f;
    const float mx = warp_max(fmaxf(m0, m1));
    const float w0 = ex2(m0 - mx), w1 = ex2(m1 - mx);
    sw[c0 * GROUP + h] = w0;
    sw[c1 * GROUP + h] = w1;
    const float dd = warp_sum(d0 * w0 + d1 * w1);
    if (lane == 0) sD[h] = dd;
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < GROUP; ++h) {
    const float* p0 =
        a.partial + (((size_t)out_slot * a.max_chunks) * a.nq + kvh * GROUP + h) * kPartStride + t;
    const size_t cstride = (size_t)a.nq * kPartStride;
    float oo0 = 0.f, oo1 = 0.f, oo2 = 0.f, oo3 = 0.f;
    int c = 0;
    for (; c + 4 <= nchunks; c += 4) {
      const float v0 = __ldcg(p0 + (c + 0) * cstride), v1 = __ldcg(p0 + (c + 1) * cstride);
      const float v2 = __ldcg(p0 + (c + 2) * cstride), v3 = __ldcg(p0 + (c + 3) * cstride);
      oo0 = fmaf(v0, sw[(c + 0) * GROUP + h], oo0);
      oo1 = fmaf(v1, sw[(c + 1) * GROUP + h], oo1);
      oo2 = fmaf(v2, sw[(c + 2) * GROUP + h], oo2);
      oo3 = fmaf(v3, sw[(c + 3) * GROUP + h], oo3);
    }
    for (; c < nchunks; ++c) oo0 = fmaf(__ldcg(p0 + c * cstride), sw[c * GROUP + h], oo0);
    a.out[((size_t)out_slot * a.nq + kvh * GROUP + h) * HD + t] =
        f2bf(__fdividef((oo0 + oo1) + (oo2 + oo3), sD[h]));
  }
  if (t == 0) a.counters[out_slot * a.nkv + kvh] = 0;  // graph-replayable
}

// Merge of the ABI split-KV partials (cascade.cuh VariableLengthMergeStates semantics):
// bf16 normalised partials + base-2 LSE -> fp32 merge -> bf16.
__global__ void merge_states_kernel(const bf16* __restrict__ tmp_v, const float* __restrict__ tmp_s,
                                    const int* __restrict__ o_indptr, bf16* __restrict__ out,
                                    int nq) {
  const int b = blockIdx.x, head = blockIdx.y, t = threadIdx.x;
  pdl_wait();
  const int s0 = o_indptr[b], s1 = o_indptr[b + 1];
  if (s1 <= s0) return;
  float mx = -INFINITY;
  for (int s = s0; s < s1; ++s) mx = fmaxf(mx, tmp_s[(size_t)s * nq + head]);
  float dd = 0.f, oo = 0.f;
  for (int s = s0; s < s1; ++s) {
    const float w = ex2(tmp_s[(size_t)s * nq + head] - mx);
    dd += w;
    oo = fmaf(bf2f(tmp_v[((size_t)s * nq + head) * HD + t]), w, oo);
  }
  out[((size_t)b * nq + head) * HD + t] = f2bf(__fdividef(oo, dd));
}

static cudaError_t launch_decode_attn(const DecodeAttnArgs& a, dim3 grid, cudaStream_t stream) {
  const int group = a.nq / a.nkv;
  switch (group) {
    case 1: return launch(decode_attention_kernel<1>, grid, dim3(kAttThreads), 0, stream, true, a);
    case 2: return launch(decode_attention_kernel<2>, grid, dim3(kAttThreads), 0, stream, true, a);
    case 4: return launch(decode_attention_kernel<4>, grid, dim3(kAttThreads), 0, stream, true, a);
    case 8: return launch(decode_attention_kernel<8>, grid, dim3(kAttThreads), 0, stream, true, a);
    default: return cudaErrorInvalidValue;
  }
}

// choose chunks so that chunks * nkv * bs ~ 2 CTAs per SM, bounded by the scratch size
static int pick_max_chunks(int bs, int nkv, int nq, size_t scratch_floats) {
  int mc = (2 * sm_count() + bs * nkv - 1) / (bs * nkv);
  if (mc > 64) mc = 64;
  if (mc < 1) mc = 1;
  while (mc > 1 && (size_t)bs * mc * nq * kPartStride > scratch_floats) --mc;
  return mc;
}


// PK_ATTN = tma (default: decode_attention_tma.cu) | cluster (round-1 LDG cluster kernel) | ticket (global-memory merge)
static int attention_impl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PK_ATTN");
    v = (e && strcmp(e, "ticket") == 0) ? 0 : ((e && strcmp(e, "cluster") == 0) ? 1 : 2);
  }
  return v;
}
static bool use_cluster_attention() { return attention_impl() >= 1; }
static int launch_cluster_or_tma(const ClusterAttnArgs& c, int nkv, int bs, cudaStream_t stream) {
  if (attention_impl() == 2) {  // (the L2-prefetch spans are an experiment of the round-1 cluster kernel only)
    const int rc = launch_decode_attention_tma(c, nkv, bs, stream);
    if (rc != -2) return rc;
  }
  return (int)launch_decode_attention_cluster(c, nkv, bs, stream);
}

// Extra cluster rows of the fused launch that only issue L2 prefetches (PK_PF_Y, default 8 -> 64 CTAs).
static int prefetch_cluster_rows() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PK_PF_Y");
    v = e ? atoi(e) : 8;
    if (v < 1) v = 1;
    if (v > 32) v = 32;
  }
  return v;
}

}  // namespace pk

using namespace pk;

extern "C" {

int paged_attention_decode_cuda(const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data,
                                int64_t k_offset_elems, int64_t v_offset_elems,
                                const int* page_indices, const int* page_indptr,
                                const int* last_page_len_d, const int* request_indices,
                                const int* kv_tile_indices, const int* kv_chunk_size_ptr,
                                int num_qo_heads, int num_kv_heads, int head_dim, int page_size,
                                int batch_size, int64_t stride_page, float sm_scale,
                                pk_stream stream) {
  (void)kv_tile_indices;
  (void)kv_chunk_size_ptr;
  if (head_dim != HD || num_kv_heads <= 0 || num_qo_heads % num_kv_heads != 0) return -1;
  if (batch_size <= 0) return 0;
  if (num_qo_heads == 4 * num_kv_heads && page_size == 16 && use_cluster_attention()) {
    // 8-CTA cluster per (request, kv head) with the merge over distributed shared memory: fp32 throughout, as the
    // reference's non-partition kernel
    ClusterAttnArgs c{};
    c.q = (const bf16*)q; c.out = (bf16*)output; c.kv = (bf16*)kv_data;
    c.k_off = k_offset_elems; c.v_off = v_offset_elems; c.stride_page = stride_page;
    c.page_indices = page_indices; c.page_indptr = page_indptr; c.last_page_len = last_page_len_d;
    c.request_indices = request_indices;
    c.sm_scale_log2 = sm_scale * 1.44269504088896340736f;
    c.nq = num_qo_heads; c.nkv = num_kv_heads;
    return launch_cluster_or_tma(c, num_kv_heads, batch_size, stream);
  }
  ThreadState& ts = tls();
  // scratch layout: [counters: bs*nkv ints, padded to 4 KB][fp32 partials]
  const size_t counter_bytes = 4096 + (((size_t)batch_size * num_kv_heads * 4 + 4095) & ~(size_t)4095);
  DecodeAttnArgs a{};
  a.q = (const bf16*)q; a.out = (bf16*)output; a.kv = (bf16*)kv_data;
  a.k_off = k_offset_elems; a.v_off = v_offset_elems;
  a.page_indices = page_indices; a.page_indptr = page_indptr; a.last_page_len = last_page_len_d;
  a.mode = 0; a.request_indices = request_indices;
  a.nq = num_qo_heads; a.nkv = num_kv_heads; a.page_size = page_size; a.stride_page = stride_page;
  a.sm_scale_log2 = sm_scale * 1.44269504088896340736f;
  a.min_chunk = 64;
  if (ts.scratch != nullptr && ts.scratch_bytes > counter_bytes) {
    a.counters = reinterpret_cast<int*>(ts.scratch);
    a.partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ts.scratch) + counter_bytes);
    a.max_chunks = pick_max_chunks(batch_size, num_kv_heads, num_qo_heads,
                                   (ts.scratch_bytes - counter_bytes) / 4);
  } else {
    a.max_chunks = 1;  // cublas_init() not called on this thread: no split (reference behaviour)
  }
  return (int)launch_decode_attn(a, dim3(a.max_chunks, num_kv_heads, batch_size), stream);
}

int paged_attention_decode_split_kv_cuda(
    const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data, int64_t k_offset_elems,
    int64_t v_offset_elems, const int* page_indices, const int* page_indptr,
    const int* last_page_len_d, const int* request_indices, const int* kv_tile_indices,
    const int* kv_chunk_size_ptr, const int* o_indptr, const uint8_t* block_valid_mask,
    pk_bf16* tmp_v, float* tmp_s, int num_qo_heads, int num_kv_heads, int head_dim, int page_size,
    int batch_size, int padded_batch_size, int64_t stride_page, float sm_scale, pk_stream stream) {
  if (head_dim != HD || num_kv_heads <= 0 || num_qo_heads % num_kv_heads != 0) return -1;
  if (batch_size <= 0 || padded_batch_size <= 0) return 0;
  DecodeAttnArgs a{};
  a.q = (const bf16*)q; a.out = (bf16*)output; a.kv = (bf16*)kv_data;
  a.k_off = k_offset_elems; a.v_off = v_offset_elems;
  a.page_indices = page_indices; a.page_indptr = page_indptr; a.last_page_len = last_page_len_d;
  a.mode = 1; a.request_indices = request_indices; a.kv_tile_indices = kv_tile_indices;
  a.kv_chunk_size = kv_chunk_size_ptr; a.valid_mask = block_valid_mask;
  a.tmp_v = (bf16*)tmp_v; a.tmp_s = tmp_s;
  a.nq = num_qo_heads; a.nkv = num_kv_heads; a.page_size = page_size; a.stride_page = stride_page;
  a.sm_scale_log2 = sm_scale * 1.44269504088896340736f;
  cudaError_t e = launch_decode_attn(a, dim3(padded_batch_size, num_kv_heads, 1), stream);
  if (e != cudaSuccess) return (int)e;
  return (int)launch(merge_states_kernel, dim3(batch_size, num_qo_heads), dim3(HD), 0, stream, true,
                     (const bf16*)tmp_v, (const float*)tmp_s, o_indptr, (bf16*)output,
                     num_qo_heads);
}

static int decode_attention_fused_impl(
    const pk_bf16* q, const pk_bf16* k, const pk_bf16* v, pk_bf16* output, pk_bf16* kv_data,
    int64_t k_offset_elems, int64_t v_offset_elems, const int* page_indices,
    const int* page_indptr, const int* last_page_len_d, const int* positions,
    const pk_bf16* q_norm_weight, const pk_bf16* k_norm_weight, const pk_bf16* cos_cache,
    const pk_bf16* sin_cache, float rms_eps, float* partial_scratch, int* counters,
    int chunk_tokens, int max_chunks, int num_qo_heads, int num_kv_heads, int head_dim,
    int page_size, int batch_size, int64_t stride_page, float sm_scale,
    const pk_b200_prefetch_span* spans, int num_spans, pk_stream stream) {
  if (head_dim != HD || num_kv_heads <= 0 || num_qo_heads % num_kv_heads != 0) return -1;
  if (batch_size <= 0) return 0;
  if (num_qo_heads == 4 * num_kv_heads && page_size == 16 && use_cluster_attention()) {
    // default: one 8-CTA cluster per (request, kv head), merge over distributed shared memory
    ClusterAttnArgs c{};
    c.q = (const bf16*)q; c.k_new = (const bf16*)k; c.v_new = (const bf16*)v; c.out = (bf16*)output;
    c.kv = (bf16*)kv_data; c.k_off = k_offset_elems; c.v_off = v_offset_elems; c.stride_page = stride_page;
    c.page_indices = page_indices; c.page_indptr = page_indptr; c.last_page_len = last_page_len_d;
    c.positions = positions;
    c.qw = (const bf16*)q_norm_weight; c.kw = (const bf16*)k_norm_weight;
    c.cosc = (const bf16*)cos_cache; c.sinc = (const bf16*)sin_cache;
    c.eps = rms_eps; c.sm_scale_log2 = sm_scale * 1.44269504088896340736f;
    c.nq = num_qo_heads; c.nkv = num_kv_heads;
    for (int s = 0; s < num_spans && c.npf < kMaxPfSpans; ++s) {
      const pk_b200_prefetch_span& sp = spans[s];
      if (!sp.base || sp.rows <= 0 || sp.slices <= 0 || sp.prefetch_rows <= 0) continue;
      if (sp.row_bytes <= 0 || sp.row_bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(sp.base) & 15) != 0) return -1;
      c.pf[c.npf++] = PfSpan{static_cast<const uint8_t*>(sp.base), sp.rows, sp.row_bytes, sp.slices, sp.prefetch_rows};
    }
    c.pf_y = prefetch_cluster_rows();
    return launch_cluster_or_tma(c, num_kv_heads, batch_size, stream);
  }
  DecodeAttnArgs a{};
  a.q = (const bf16*)q; a.out = (bf16*)output; a.kv = (bf16*)kv_data;
  a.k_off = k_offset_elems; a.v_off = v_offset_elems;
  a.page_indices = page_indices; a.page_indptr = page_indptr; a.last_page_len = last_page_len_d;
  a.mode = 0;
  a.partial = partial_scratch; a.counters = counters;
  a.min_chunk = chunk_tokens > 0 ? chunk_tokens : 64;
  a.max_chunks = max_chunks > 0 ? max_chunks : 1;
  a.nq = num_qo_heads; a.nkv = num_kv_heads; a.page_size = page_size; a.stride_page = stride_page;
  a.sm_scale_log2 = sm_scale * 1.44269504088896340736f;
  a.fused = 1;
  a.k_new = (const bf16*)k; a.v_new = (const bf16*)v; a.positions = positions;
  a.qw = (const bf16*)q_norm_weight; a.kw = (const bf16*)k_norm_weight;
  a.cosc = (const bf16*)cos_cache; a.sinc = (const bf16*)sin_cache; a.eps = rms_eps;
  return (int)launch_decode_attn(a, dim3(a.max_chunks, num_kv_heads, batch_size), stream);
}

int pk_b200_decode_attention_fused(
    const pk_bf16* q, const pk_bf16* k, const pk_bf16* v, pk_bf16* output, pk_bf16* kv_data,
    int64_t k_offset_elems, int64_t v_offset_elems, const int* page_indices,
    const int* page_indptr, const int* last_page_len_d, const int* positions,
    const pk_bf16* q_norm_weight, const pk_bf16* k_norm_weight, const pk_bf16* cos_cache,
    const pk_bf16* sin_cache, float rms_eps, float* partial_scratch, int* counters,
    int chunk_tokens, int max_chunks, int num_qo_heads, int num_kv_heads, int head_dim,
    int page_size, int batch_size, int64_t stride_page, float sm_scale, pk_stream stream) {
  return decode_attention_fused_impl(q, k, v, output, kv_data, k_offset_elems, v_offset_elems, page_indices,
                                     page_indptr, last_page_len_d, positions, q_norm_weight, k_norm_weight,
                                     cos_cache, sin_cache, rms_eps, partial_scratch, counters, chunk_tokens,
                                     max_chunks, num_qo_heads, num_kv_heads, head_dim, page_size, batch_size,
                                     stride_page, sm_scale, nullptr, 0, stream);
}

int pk_b200_decode_attention_fused_prefetch(
    const pk_bf16* q, const pk_bf16* k, const pk_bf16* v, pk_bf16* output, pk_bf16* kv_data,
    int64_t k_offset_elems, int64_t v_offset_elems, const int* page_indices,
    const int* page_indptr, const int* last_page_len_d, const int* positions,
    const pk_bf16* q_norm_weight, const pk_bf16* k_norm_weight, const pk_bf16* cos_cache,
    const pk_bf16* sin_cache, float rms_eps, float* partial_scratch, int* counters,
    int chunk_tokens, int max_chunks, int num_qo_heads, int num_kv_heads, int head_dim,
    int page_size, int batch_size, int64_t stride_page, float sm_scale,
    const pk_b200_prefetch_span* spans, int num_spans, pk_stream stream) {
  if (num_spans < 0 || (num_spans > 0 && !spans)) return -1;
  return decode_attention_fused_impl(q, k, v, output, kv_data, k_offset_elems, v_offset_elems, page_indices,
                                     page_indptr, last_page_len_d, positions, q_norm_weight, k_norm_weight,
                                     cos_cache, sin_cache, rms_eps, partial_scratch, counters, chunk_tokens,
                                     max_chunks, num_qo_heads, num_kv_heads, head_dim, page_size, batch_size,
                                     stride_page, sm_scale, spans, num_spans, stream);
}

}  // extern "C"
