// Fused decode attention, thread-block-cluster version (the default for the fused decode path).
//
// One cluster of 8 CTAs per (request, kv head): CTA r owns the r-th eighth of the context.  Each CTA does the
// QK-norm + RoPE of its GQA group's 4 query heads, streams its K/V rows (64-token rounds, the next round's
// 16-byte rows requested before the current round is reduced), keeps an online-softmax state per half-warp,
// merges inside the CTA through shared memory, then writes its (max, denominator, 4 x 128 numerators) straight
// into the LEADER CTA's shared memory over DSMEM (st.shared::cluster), one cluster barrier, and the leader
// produces the bf16 output.  Compared with the ticket version in decode_attention.cu (profiles/README.md,
// round 1 v3: 22 us per layer, of which the partial-store -> __threadfence -> atomic ticket -> L2 re-read chain
// is 3-4 dependent global round trips) there is no global-memory synchronisation at all.
// The CTA that owns the step's position also appends the normed/roped K and the V row to the paged cache and
// injects the new token from shared memory.  Arithmetic and rounding points: see decode_attention.cu.
#include "decode_attention_cluster.cuh"

namespace pk {

constexpr int CHD = 128;
constexpr int C_WARPS = 8;
constexpr int C_THREADS = C_WARPS * 32;
constexpr int C_STEP = C_WARPS * 2;  // tokens per step: one per half-warp
constexpr int C_U = 4;               // steps per round -> 64 tokens
constexpr int C_ROUND = C_STEP * C_U;
constexpr int C_CLUSTER = 8;
constexpr int C_GROUP = 4;

__device__ __forceinline__ float cex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cunpack8(const uint4& a, float* f) {
  f[0] = bf16_lo(a.x); f[1] = bf16_hi(a.x); f[2] = bf16_lo(a.y); f[3] = bf16_hi(a.y);
  f[4] = bf16_lo(a.z); f[5] = bf16_hi(a.z); f[6] = bf16_lo(a.w); f[7] = bf16_hi(a.w);
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_rank(const void* smem_ptr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// split cluster barrier: arrive at entry, wait before the first remote store (every CTA of the cluster has started)
__device__ __forceinline__ void cluster_arrive_open() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_open() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// warp-level QK RMSNorm + RoPE of one 128-wide head (qk_norm_rope.cu arithmetic), result to smem
__device__ __forceinline__ void c_norm_rope(const bf16* __restrict__ src, const bf16* __restrict__ w,
                                            const bf16* __restrict__ cosc, const bf16* __restrict__ sinc, int pos,
                                            float eps, bf16* dst, int lane) {
  const uint2 raw = reinterpret_cast<const uint2*>(src)[lane];
  const float v[4] = {bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y)};
  float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)CHD + eps);
  const uint2 wr = reinterpret_cast<const uint2*>(w)[lane];
  const float wv[4] = {bf16_lo(wr.x), bf16_hi(wr.x), bf16_lo(wr.y), bf16_hi(wr.y)};
  const int ci = (lane & 15) * 4;
  const uint2 cr = reinterpret_cast<const uint2*>(cosc + (size_t)pos * CHD + ci)[0];
  const uint2 sr = reinterpret_cast<const uint2*>(sinc + (size_t)pos * CHD + ci)[0];
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

#ifndef PK_ATTN_MIN_CTAS
#define PK_ATTN_MIN_CTAS 1
#endif
__global__ void __cluster_dims__(C_CLUSTER, 1, 1) __launch_bounds__(C_THREADS, PK_ATTN_MIN_CTAS)
decode_attention_cluster_kernel(const ClusterAttnArgs a) {
  __shared__ __align__(16) float st_o[C_WARPS + 1][C_GROUP][CHD];
  __shared__ float st_m[C_WARPS + 1][C_GROUP], st_d[C_WARPS + 1][C_GROUP];
  __shared__ __align__(16) bf16 q_s[C_GROUP][CHD];
  __shared__ __align__(16) bf16 k_s[CHD], v_s[CHD];
  // leader-side landing zone for the 8 CTAs' states (every CTA has it; only the leader's is used)
  __shared__ __align__(16) float c_o[C_CLUSTER][C_GROUP][CHD];
  __shared__ float c_m[C_CLUSTER][C_GROUP], c_d[C_CLUSTER][C_GROUP];

  if ((int)blockIdx.y >= a.nkv) {
    // ---- prefetch clusters: decode attention is latency-bound and leaves HBM almost idle (a few MB of K/V per
    // layer against ~10 us of dependent steps), so these CTAs spend the window pulling the head of the NEXT
    // GEMVs' weight slices into L2 (fire-and-forget bulk prefetches).  They wait for the previous kernel first:
    // issuing earlier would only compete with the qkv GEMV's own stream.
    pdl_launch_dependents();
    pdl_wait();
    const int ncta = a.pf_y * C_CLUSTER * (int)gridDim.z;
    const int cta = (((int)blockIdx.z * a.pf_y) + ((int)blockIdx.y - a.nkv)) * C_CLUSTER + (int)blockIdx.x;
    const int nthr = ncta * C_THREADS;
    int unit0 = 0;
    for (int s = 0; s < a.npf; ++s) {
      const PfSpan sp = a.pf[s];
      const int units = sp.slices * sp.pf_rows;
      // round-robin over all prefetch threads, continuing where the previous span stopped
      int u = cta * C_THREADS + (int)threadIdx.x - unit0 % nthr;
      if (u < 0) u += nthr;
      for (; u < units; u += nthr) {
        const int slice = u / sp.pf_rows, j = u - slice * sp.pf_rows;
        const int r0 = (int)(((int64_t)slice * sp.rows) / sp.slices);
        const int r1 = (int)(((int64_t)(slice + 1) * sp.rows) / sp.slices);
        if (r0 + j < r1) bulk_prefetch_l2(sp.base + (size_t)(r0 + j) * sp.row_bytes, (uint32_t)sp.row_bytes);
      }
      unit0 += units;
    }
    return;
  }
  cluster_arrive_open();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const int rank = (int)cluster_rank();
  // plain mode (the ABI's paged_attention_decode_cuda): q arrives normed + roped, the step's K/V row was appended
  // by an earlier launch -> no norm, no injection, and no K/V request before griddepcontrol.wait
  const bool plain = a.k_new == nullptr;
  const int kvh = blockIdx.y;
  const int b = (plain && a.request_indices) ? a.request_indices[blockIdx.z] : (int)blockIdx.z;
  const int npages = a.page_indptr[b + 1] - a.page_indptr[b];
  const int len = npages <= 0 ? 0 : (npages - 1) * 16 + a.last_page_len[b];
  int chunk = (len + C_CLUSTER - 1) / C_CLUSTER;
  chunk = (chunk + C_ROUND - 1) / C_ROUND * C_ROUND;
  const int lo = min(len, rank * chunk), hi = min(len, lo + chunk);
  const int* pages = a.page_indices + a.page_indptr[b];
  const int pos = plain ? -1 : a.positions[b];
  const bool inject = pos >= lo && pos < hi;
  const bf16* kbase = a.kv + a.k_off + (int64_t)kvh * CHD + l16 * 8;
  const bf16* vbase = a.kv + a.v_off + (int64_t)kvh * CHD + l16 * 8;

  pdl_launch_dependents();  // the o_proj GEMV may start prefetching its weights

  // K/V rows of one round for this half-warp: token t = round + warp*2 + half + 16*u lives in page
  // (round >> 4) + u.  Cached rows never depend on the previous kernel -> round 0 is requested before
  // griddepcontrol.wait.
  uint4 kr[C_U], vr[C_U];
  bool ok[C_U];
  auto load_round = [&](int round, uint4* kk, uint4* vv, bool* okk) {
#pragma unroll
    for (int u = 0; u < C_U; ++u) {
      const int t = round + warp * 2 + half + C_STEP * u;
      okk[u] = t < hi && t != pos;
      kk[u] = make_uint4(0, 0, 0, 0);
      vv[u] = make_uint4(0, 0, 0, 0);
      if (okk[u]) {
        const int page = __ldg(pages + (t >> 4));
        const int64_t off = (int64_t)page * a.stride_page + (int64_t)(t & 15) * a.nkv * CHD;
        kk[u] = ldg_stream(kbase + off);
        vv[u] = ldg_stream(vbase + off);
      }
    }
  };
  if (!plain) load_round(lo, kr, vr, ok);

  pdl_wait();
  if (plain) load_round(lo, kr, vr, ok);

  // ---- q heads (warps 0-3), the step's k/v (warp 4 of the owning CTA) ----
  if (warp < C_GROUP) {
    const bf16* qsrc = a.q + ((size_t)b * a.nq + kvh * C_GROUP + warp) * CHD;
    if (plain) reinterpret_cast<uint2*>(q_s[warp])[lane] = reinterpret_cast<const uint2*>(qsrc)[lane];
    else c_norm_rope(qsrc, a.qw, a.cosc, a.sinc, pos, a.eps, q_s[warp], lane);
  }
  if (inject && warp == C_GROUP) {
    c_norm_rope(a.k_new + ((size_t)b * a.nkv + kvh) * CHD, a.kw, a.cosc, a.sinc, pos, a.eps, k_s, lane);
    reinterpret_cast<uint2*>(v_s)[lane] = reinterpret_cast<const uint2*>(a.v_new + ((size_t)b * a.nkv + kvh) * CHD)[lane];
    __syncwarp();
    const int page = pages[pos >> 4], slot = pos & 15;
    const int64_t dst = (int64_t)page * a.stride_page + ((int64_t)slot * a.nkv + kvh) * CHD;
    reinterpret_cast<uint2*>(a.kv + a.k_off + dst)[lane] = reinterpret_cast<uint2*>(k_s)[lane];
    reinterpret_cast<uint2*>(a.kv + a.v_off + dst)[lane] = reinterpret_cast<uint2*>(v_s)[lane];
  }
  __syncthreads();
  float qf[C_GROUP][8];
#pragma unroll
  for (int h = 0; h < C_GROUP; ++h) cunpack8(reinterpret_cast<const uint4*>(q_s[h])[l16], qf[h]);

  float m[C_GROUP], d[C_GROUP], o[C_GROUP][8];
#pragma unroll
  for (int h = 0; h < C_GROUP; ++h) {
    m[h] = -INFINITY;
    d[h] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[h][j] = 0.f;
  }

  for (int round = lo; round < hi; round += C_ROUND) {  // warp-uniform trip count
    uint4 krn[C_U], vrn[C_U];
    bool okn[C_U];
    const bool more = round + C_ROUND < hi;
    if (more) load_round(round + C_ROUND, krn, vrn, okn);  // next round in flight during this round's math
    float s[C_GROUP][C_U];
#pragma unroll
    for (int u = 0; u < C_U; ++u) {
      float kf[8];
      cunpack8(kr[u], kf);
#pragma unroll
      for (int h = 0; h < C_GROUP; ++h) {
        float p = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) p = fmaf(qf[h][j], kf[j], p);
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
        s[h][u] = ok[u] ? p * a.sm_scale_log2 : -INFINITY;
      }
    }
#pragma unroll
    for (int h = 0; h < C_GROUP; ++h) {
      float mn = m[h];
#pragma unroll
      for (int u = 0; u < C_U; ++u) mn = fmaxf(mn, s[h][u]);
      if (mn == -INFINITY) {
#pragma unroll
        for (int u = 0; u < C_U; ++u) s[h][u] = 0.f;
        continue;
      }
      const float sc = cex2(m[h] - mn);
      m[h] = mn;
      d[h] *= sc;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[h][j] *= sc;
#pragma unroll
      for (int u = 0; u < C_U; ++u) s[h][u] = cex2(s[h][u] - mn);
#pragma unroll
      for (int u = 0; u < C_U; ++u) d[h] += s[h][u];
    }
#pragma unroll
    for (int u = 0; u < C_U; ++u) {
      float vf[8];
      cunpack8(vr[u], vf);
#pragma unroll
      for (int h = 0; h < C_GROUP; ++h)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[h][j] = fmaf(s[h][u], vf[j], o[h][j]);
    }
    if (more) {
#pragma unroll
      for (int u = 0; u < C_U; ++u) {
        kr[u] = krn[u];
        vr[u] = vrn[u];
        ok[u] = okn[u];
      }
    }
  }

  // ---- merge: half-warps by shuffle, warps (+ injected token) through shared memory ----
#pragma unroll
  for (int h = 0; h < C_GROUP; ++h) {
    const float om = __shfl_xor_sync(0xffffffffu, m[h], 16), od = __shfl_xor_sync(0xffffffffu, d[h], 16);
    const float mn = fmaxf(m[h], om);
    const float wa = mn == -INFINITY ? 0.f : cex2(m[h] - mn), wb = mn == -INFINITY ? 0.f : cex2(om - mn);
    d[h] = d[h] * wa + od * wb;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float oo = __shfl_xor_sync(0xffffffffu, o[h][j], 16);
      o[h][j] = o[h][j] * wa + oo * wb;
    }
    m[h] = mn;
    if (half == 0) {
      if (l16 == 0) {
        st_m[warp][h] = m[h];
        st_d[warp][h] = d[h];
      }
      float4* dst = reinterpret_cast<float4*>(&st_o[warp][h][l16 * 8]);
      dst[0] = make_float4(o[h][0], o[h][1], o[h][2], o[h][3]);
      dst[1] = make_float4(o[h][4], o[h][5], o[h][6], o[h][7]);
    }
  }
  if (inject && warp < C_GROUP) {
    const int h = warp;
    const uint2 qr = reinterpret_cast<const uint2*>(q_s[h])[lane];
    const uint2 kr2 = reinterpret_cast<const uint2*>(k_s)[lane];
    float p = bf16_lo(qr.x) * bf16_lo(kr2.x);
    p = fmaf(bf16_hi(qr.x), bf16_hi(kr2.x), p);
    p = fmaf(bf16_lo(qr.y), bf16_lo(kr2.y), p);
    p = fmaf(bf16_hi(qr.y), bf16_hi(kr2.y), p);
    p = warp_sum(p);
    if (lane == 0) {
      st_m[C_WARPS][h] = p * a.sm_scale_log2;
      st_d[C_WARPS][h] = 1.f;
    }
    const uint2 vr2 = reinterpret_cast<const uint2*>(v_s)[lane];
    *reinterpret_cast<float4*>(&st_o[C_WARPS][h][lane * 4]) =
        make_float4(bf16_lo(vr2.x), bf16_hi(vr2.x), bf16_lo(vr2.y), bf16_hi(vr2.y));
  }
  __syncthreads();
  cluster_wait_open();
  const int nstates = inject ? C_WARPS + 1 : C_WARPS;
  const int t = threadIdx.x;
  if (t < CHD) {
#pragma unroll
    for (int h = 0; h < C_GROUP; ++h) {
      float mx = -INFINITY;
      for (int i = 0; i < nstates; ++i) mx = fmaxf(mx, st_m[i][h]);
      float dd = 0.f, oo = 0.f;
      if (mx != -INFINITY) {
        for (int i = 0; i < nstates; ++i) {
          const float w = cex2(st_m[i][h] - mx);
          dd = fmaf(st_d[i][h], w, dd);
          oo = fmaf(st_o[i][h][t], w, oo);
        }
      }
      // ---- DSMEM: this CTA's state into the leader's landing zone ----
      st_cluster_f32(map_to_rank(&c_o[rank][h][t], 0), oo);
      if (t == 0) {
        st_cluster_f32(map_to_rank(&c_m[rank][h], 0), mx);
        st_cluster_f32(map_to_rank(&c_d[rank][h], 0), dd);
      }
    }
  }
  cluster_sync_all();
  if (rank != 0 || t >= CHD) return;
#pragma unroll
  for (int h = 0; h < C_GROUP; ++h) {
    float mx = -INFINITY;
#pragma unroll
    for (int r = 0; r < C_CLUSTER; ++r) mx = fmaxf(mx, c_m[r][h]);
    float dd = 0.f, oo = 0.f;
#pragma unroll
    for (int r = 0; r < C_CLUSTER; ++r) {
      const float w = c_m[r][h] == -INFINITY ? 0.f : cex2(c_m[r][h] - mx);
      dd = fmaf(c_d[r][h], w, dd);
      oo = fmaf(c_o[r][h][t], w, oo);
    }
    a.out[((size_t)b * a.nq + kvh * C_GROUP + h) * CHD + t] = f2bf(dd > 0.f ? __fdividef(oo, dd) : 0.f);
  }
}

cudaError_t launch_decode_attention_cluster(const ClusterAttnArgs& a, int nkv, int bs, cudaStream_t stream) {
  const int pf_y = a.npf > 0 ? a.pf_y : 0;
  return launch(decode_attention_cluster_kernel, dim3(C_CLUSTER, nkv + pf_y, bs), dim3(C_THREADS), 0, stream, true, a);
}

}  // namespace pk
