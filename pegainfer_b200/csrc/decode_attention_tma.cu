// Fused decode attention, round-2 design: TMA page tiles + tensor-core (mma.sync) score / PV products.
//
// Why: the round-1 cluster kernel (decode_attention_cluster.cu) reads K/V with per-lane LDGs and reduces every
// q.k product with a 4-level shuffle tree -- ncu (profiles/r1_v4_attn_cluster_full_details.txt) shows ~215 issued
// instructions per token row and 2 warps per scheduler: latency-bound on its own instruction stream (6.8 cycles per
// issue), 6 % of DRAM bandwidth, 64 CTAs on 148 SMs.  Here:
//   * one thread-block cluster of up to 16 CTAs per (request, kv head) -> 128 CTAs for Qwen3-4B at bs 1; CTA r owns a
//     contiguous run of 16-token pages
//   * every page of the CTA is fetched by TMA (cp.async.bulk.tensor.4d over the page-first pool, box = one 64-column
//     half of one page of one kv head, 128-byte swizzle) into a ring of page slots, ALL requests issued at kernel
//     entry -- BEFORE griddepcontrol.wait: cached rows do not depend on the previous kernel, so the K/V stream
//     overlaps the tail of the qkv GEMV; nothing but the mbarrier wait sits between the previous kernel and the math
//   * S = Q K^T and O += P V run on the tensor cores (mma.sync.m16n8k16, bf16 x bf16 -> fp32): the GQA group's 4 query
//     heads are rows 0..3 of the 16-row A operand (K/V are read once per group), K tiles come out of shared memory
//     with ldmatrix, V with ldmatrix.trans; ~100 instructions per 16-token page per warp instead of ~3400.
//     P is split into bf16 hi + lo parts (two MMAs), so the PV product keeps ~16 mantissa bits of p: the reference's
//     decode kernel uses fp32 p (decode.cuh:62-145), and this stays within its rounding behaviour (no bf16 P).
//   * per-warp online-softmax states are merged through shared memory, CTA states through DISTRIBUTED shared memory
//     (st.shared::cluster) into four "head leader" CTAs (CTA h finishes query head h) -- no global-memory
//     synchronisation, final merge spread over 4 SMs.
// Also does QK-norm + RoPE of the 4 query heads and of the step's K, appends K/V to the paged cache and patches the
// new row into the shared-memory tile (the TMA request for that page was issued before the row existed).
// Plain mode (k_new == nullptr) is `paged_attention_decode_cuda`: q arrives normed + roped, K/V already appended,
// every request is issued after griddepcontrol.wait.
// Rounding points: q/k norm + RoPE as qk_norm_rope.cu; scores, softmax and PV in fp32; ONE bf16 rounding of O / d.
#include <cstdlib>
#include <cstring>

#include "decode_attention_cluster.cuh"
#include "tcgen05.cuh"

namespace pk {

namespace {

constexpr int T_HD = 128;
constexpr int T_WARPS = 8;
constexpr int T_THREADS = T_WARPS * 32;
constexpr int T_GROUP = 4;
constexpr int T_PAGE = 16;
constexpr int T_HALF = T_PAGE * 128;        // bytes of one 64-column half of a page tile (16 rows x 128 B)
constexpr int T_SLOT = 4 * T_HALF;          // K lo | K hi | V lo | V hi = 8 KB
constexpr int T_MAX_CLUSTER = 16;
constexpr int T_MAX_SLOTS = 24;

struct TmaAttnArgs {
  const bf16 *q, *k_new, *v_new;
  bf16* out;
  bf16* kv;
  int64_t k_off, v_off, stride_page;
  const int *page_indices, *page_indptr, *last_page_len, *positions, *request_indices;
  const bf16 *qw, *kw, *cosc, *sinc;
  float eps, sm_scale_log2;
  int nq, nkv, nslot;
};

__device__ __forceinline__ float t_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t t_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t t_cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t t_map_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void t_st_cluster(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void t_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// arrive (release, cluster scope) on an mbarrier that lives in another CTA of the cluster
__device__ __forceinline__ void t_mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool t_mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void t_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__device__ __forceinline__ void t_tma_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void t_wait_or_trap(uint64_t* bar, uint32_t parity) {  // a hung barrier becomes a trap
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void t_ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void t_ldsm4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void t_mma(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk c16 (0..15) of token row r (0..15) inside one K (or V) page tile: two 64-column halves,
// each [16 rows x 128 B] with the TMA 128-byte swizzle (chunk ^= row & 7)
__device__ __forceinline__ uint32_t t_sw(int r, int c16) {
  return (uint32_t)((c16 >> 3) * T_HALF + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
}

// warp-level QK RMSNorm + RoPE of one 128-wide head (qk_norm_rope.cu arithmetic), result to shared memory
__device__ __forceinline__ void t_norm_rope(const bf16* __restrict__ src, const bf16* __restrict__ w,
                                            const bf16* __restrict__ cosc, const bf16* __restrict__ sinc, int pos,
                                            float eps, bf16* dst, int lane) {
  const uint2 raw = reinterpret_cast<const uint2*>(src)[lane];
  const float v[4] = {bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y)};
  float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)T_HD + eps);
  const uint2 wr = reinterpret_cast<const uint2*>(w)[lane];
  const float wv[4] = {bf16_lo(wr.x), bf16_hi(wr.x), bf16_lo(wr.y), bf16_hi(wr.y)};
  const int ci = (lane & 15) * 4;
  const uint2 cr = reinterpret_cast<const uint2*>(cosc + (size_t)pos * T_HD + ci)[0];
  const uint2 sr = reinterpret_cast<const uint2*>(sinc + (size_t)pos * T_HD + ci)[0];
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

__global__ void __launch_bounds__(T_THREADS, 1)
decode_attention_tma_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                            const TmaAttnArgs a) {
  extern __shared__ uint8_t t_smem_raw[];
  const uint32_t raw_addr = smem_u32(t_smem_raw);
  uint8_t* smem = t_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // swizzled tiles need 1024-B alignment
  const int nslot = a.nslot;
  uint8_t* tiles = smem;                                                  // [nslot][8 KB]; later: per-warp states
  float* c_o = reinterpret_cast<float*>(smem + (size_t)nslot * T_SLOT);   // [16][128] landing zone (remote writes)
  float* c_m = c_o + T_MAX_CLUSTER * T_HD;                                // [16]
  float* c_d = c_m + T_MAX_CLUSTER;                                       // [16]
  float* st_m = c_d + T_MAX_CLUSTER;                                      // [8 warps][4 heads]
  float* st_d = st_m + T_WARPS * T_GROUP;
  bf16* q_s = reinterpret_cast<bf16*>(st_d + T_WARPS * T_GROUP);          // [4][128]
  bf16* k_s = q_s + T_GROUP * T_HD;
  bf16* v_s = k_s + T_HD;
  uint64_t* full = reinterpret_cast<uint64_t*>(v_s + T_HD);               // [nslot]
  uint64_t* merged = full + T_MAX_SLOTS;                                  // leaders: all CTAs' states have landed
  float* st_o = reinterpret_cast<float*>(tiles);                          // [8][4][128] (after the page loop)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int rank = (int)t_cluster_rank(), cs = (int)t_cluster_size();
  const bool plain = a.k_new == nullptr;
  const int kvh = blockIdx.y;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nslot; ++s) mbar_init(full + s, 1);
    // one arrival per (head this CTA leads, source CTA, source warp of the 4 that carry that head's 128 dims)
    const int heads_led = rank < T_GROUP ? (T_GROUP - rank + cs - 1) / cs : 0;
    mbar_init(merged, (uint32_t)max(1, heads_led * cs * 4));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    mbar_fence_init();
  }
  t_cluster_arrive();  // opening cluster barrier: completed (t_cluster_wait) before the first remote store
  __syncthreads();
  pdl_launch_dependents();  // the o_proj GEMV may start prefetching its weights
  if (plain) pdl_wait();    // ABI mode: metadata and the appended row come from the previous kernels

  const int b = (plain && a.request_indices) ? a.request_indices[blockIdx.z] : (int)blockIdx.z;
  const int npages = a.page_indptr[b + 1] - a.page_indptr[b];
  const int len = npages <= 0 ? 0 : (npages - 1) * T_PAGE + a.last_page_len[b];
  const int tot_pages = (len + T_PAGE - 1) / T_PAGE;
  const int ppc = (tot_pages + cs - 1) / cs;
  const int p_lo = min(tot_pages, rank * ppc), p_hi = min(tot_pages, p_lo + ppc);
  const int my_pages = p_hi - p_lo;
  const int* pages = a.page_indices + a.page_indptr[b];
  const int pos = plain ? -1 : a.positions[b];
  const int pos_lp = (pos >= 0 && (pos >> 4) >= p_lo && (pos >> 4) < p_hi) ? (pos >> 4) - p_lo : -1;

  auto issue_page = [&](int lp) {  // one lane: arm the slot's barrier and request the page's four half tiles
    const int slot = lp % nslot;
    const int page = __ldg(pages + p_lo + lp);
    const uint32_t dst = smem_u32(tiles + (size_t)slot * T_SLOT);
    mbar_expect_tx(full + slot, T_SLOT);
    t_tma_4d(dst, &map_k, 0, kvh, 0, page, full + slot);
    t_tma_4d(dst + T_HALF, &map_k, 64, kvh, 0, page, full + slot);
    t_tma_4d(dst + 2 * T_HALF, &map_v, 0, kvh, 0, page, full + slot);
    t_tma_4d(dst + 3 * T_HALF, &map_v, 64, kvh, 0, page, full + slot);
  };
  // first pass over the ring: each slot's owner warp (slot & 7) requests its page now
  if (lane == 0)
    for (int lp = warp; lp < my_pages && lp < nslot; lp += T_WARPS) issue_page(lp);
  // (slots >= 8 belong to warps slot & 7 as well: lp = warp + 8k covers them)

  if (!plain) pdl_wait();  // q / k_new / v_new come from the qkv GEMV

  // ---- q heads (warps 0-3), the step's k/v row (warp 4 of the CTA that owns its page) ----
  if (warp < T_GROUP) {
    const bf16* qsrc = a.q + ((size_t)b * a.nq + kvh * T_GROUP + warp) * T_HD;
    if (plain) reinterpret_cast<uint2*>(q_s + warp * T_HD)[lane] = reinterpret_cast<const uint2*>(qsrc)[lane];
    else t_norm_rope(qsrc, a.qw, a.cosc, a.sinc, pos, a.eps, q_s + warp * T_HD, lane);
  } else if (warp == T_GROUP && pos_lp >= 0) {
    t_norm_rope(a.k_new + ((size_t)b * a.nkv + kvh) * T_HD, a.kw, a.cosc, a.sinc, pos, a.eps, k_s, lane);
    reinterpret_cast<uint2*>(v_s)[lane] = reinterpret_cast<const uint2*>(a.v_new + ((size_t)b * a.nkv + kvh) * T_HD)[lane];
    __syncwarp();
    const int page = pages[pos >> 4], slot_row = pos & 15;
    const int64_t dst = (int64_t)page * a.stride_page + ((int64_t)slot_row * a.nkv + kvh) * T_HD;
    reinterpret_cast<uint2*>(a.kv + a.k_off + dst)[lane] = reinterpret_cast<uint2*>(k_s)[lane];
    reinterpret_cast<uint2*>(a.kv + a.v_off + dst)[lane] = reinterpret_cast<uint2*>(v_s)[lane];
  }
  __syncthreads();

  // ---- A operand: the group's 4 query heads are rows 0..3 of the 16-row tile, rows 4..15 are zero ----
  uint32_t qa[8][2];
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    qa[kk][0] = qa[kk][1] = 0u;
    if (g < T_GROUP) {
      qa[kk][0] = *reinterpret_cast<const uint32_t*>(q_s + g * T_HD + kk * 16 + 2 * tq);
      qa[kk][1] = *reinterpret_cast<const uint32_t*>(q_s + g * T_HD + kk * 16 + 8 + 2 * tq);
    }
  }
  float o[16][4];
#pragma unroll
  for (int d = 0; d < 16; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  // ldmatrix lane addressing (matrix i = lane / 8, row r = lane % 8)
  const int lm_i = lane >> 3, lm_r = lane & 7;
  const int k_tok = ((lm_i >> 1) << 3) + lm_r, k_c = lm_i & 1;   // K: matrices {tok 0-7 lo, tok 0-7 hi, tok 8-15 lo, tok 8-15 hi}
  const int v_tok = ((lm_i & 1) << 3) + lm_r, v_c = lm_i >> 1;   // V: {tok 0-7 | tok 8-15} x {dims 8d | 8(d+1)}

  // a slot is consumed by ONE warp (slot & 7), in page order: walk the ring passes, then this warp's slots
  for (int base = 0, pass = 0; base < my_pages; base += nslot, ++pass)
  for (int slot = warp; slot < nslot; slot += T_WARPS) {
    const int lp = base + slot;
    if (lp >= my_pages) break;
    t_wait_or_trap(full + slot, (uint32_t)(pass & 1));
    uint8_t* tile = tiles + (size_t)slot * T_SLOT;
    const int gp = p_lo + lp;
    if (lp == pos_lp) {  // patch the step's own row into the tile (the TMA saw the slot before it was written)
      const int r = pos & 15;
      if (lane < 16) *reinterpret_cast<uint4*>(tile + t_sw(r, lane)) = reinterpret_cast<const uint4*>(k_s)[lane];
      else *reinterpret_cast<uint4*>(tile + 2 * T_HALF + t_sw(r, lane - 16)) = reinterpret_cast<const uint4*>(v_s)[lane - 16];
    }
    if (gp == tot_pages - 1 && (len & 15) != 0) {  // rows past the context: zero V so that 0 * garbage stays 0
      for (int i = lane; i < (16 - (len & 15)) * 16; i += 32) {
        const int r = (len & 15) + (i >> 4);
        *reinterpret_cast<uint4*>(tile + 2 * T_HALF + t_sw(r, i & 15)) = make_uint4(0, 0, 0, 0);
      }
    }
    __syncwarp();
    const uint32_t kt = smem_u32(tile), vt = kt + 2 * T_HALF;
    // ---- S[16 x 16] = Q K^T ----
    float s[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      uint32_t b0, b1, b2, b3;
      t_ldsm4(kt + t_sw(k_tok, kk * 2 + k_c), b0, b1, b2, b3);
      t_mma(s[0], qa[kk][0], 0u, qa[kk][1], 0u, b0, b1);
      t_mma(s[1], qa[kk][0], 0u, qa[kk][1], 0u, b2, b3);
    }
    // ---- online softmax on row g (lanes g >= 4 carry all-zero rows: harmless) ----
    const int tb = gp * T_PAGE + 2 * tq;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      s[j][0] = (tb + j * 8 < len) ? s[j][0] * a.sm_scale_log2 : -INFINITY;
      s[j][1] = (tb + j * 8 + 1 < len) ? s[j][1] * a.sm_scale_log2 : -INFINITY;
      mx = fmaxf(mx, fmaxf(s[j][0], s[j][1]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);
    const float m_safe = m_new == -INFINITY ? 0.f : m_new;
    const float alpha = t_ex2(m_run - m_safe);
    m_run = m_new;
    float p[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      p[j][0] = t_ex2(s[j][0] - m_safe);
      p[j][1] = t_ex2(s[j][1] - m_safe);
    }
    l_run = l_run * alpha + ((p[0][0] + p[0][1]) + (p[1][0] + p[1][1]));
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      o[d][0] *= alpha;
      o[d][1] *= alpha;
    }
    // P = hi + lo (both bf16): two MMAs keep ~16 mantissa bits of p
    uint32_t ph[2], pl[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      ph[j] = pack_bf16(p[j][0], p[j][1]);
      pl[j] = pack_bf16(p[j][0] - bf16_lo(ph[j]), p[j][1] - bf16_hi(ph[j]));
    }
    // ---- O[16 x 128] += P V ----
#pragma unroll
    for (int dp = 0; dp < 8; ++dp) {
      uint32_t v0, v1, v2, v3;
      t_ldsm4_t(vt + t_sw(v_tok, dp * 2 + v_c), v0, v1, v2, v3);
      t_mma(o[2 * dp], ph[0], 0u, ph[1], 0u, v0, v1);
      t_mma(o[2 * dp], pl[0], 0u, pl[1], 0u, v0, v1);
      t_mma(o[2 * dp + 1], ph[0], 0u, ph[1], 0u, v2, v3);
      t_mma(o[2 * dp + 1], pl[0], 0u, pl[1], 0u, v2, v3);
    }
    // refill: this warp is the only reader of the slot -> request the page that reuses it
    if (lp + nslot < my_pages) {
      __syncwarp();
      if (lane == 0) {
        fence_proxy_async_smem();
        issue_page(lp + nslot);
      }
    }
  }
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);

  // ---- per-warp states -> shared memory (aliases the page ring: every warp must be done with its tiles) ----
  __syncthreads();
  if (g < T_GROUP) {
    if (tq == 0) {
      st_m[warp * T_GROUP + g] = m_run;
      st_d[warp * T_GROUP + g] = l_run;
    }
    float* dst = st_o + ((size_t)warp * T_GROUP + g) * T_HD + 2 * tq;
#pragma unroll
    for (int d = 0; d < 16; ++d) *reinterpret_cast<float2*>(dst + d * 8) = make_float2(o[d][0], o[d][1]);
  }
  __syncthreads();
  t_cluster_wait();  // every CTA of the cluster has started: its landing zone exists

  // ---- CTA merge (thread = one output dim of two heads), result pushed to the head's leader CTA over DSMEM ----
  {
    const int dim = threadIdx.x & (T_HD - 1);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int h = (threadIdx.x >> 7) * 2 + hh;
      float mx = -INFINITY;
#pragma unroll
      for (int w = 0; w < T_WARPS; ++w) mx = fmaxf(mx, st_m[w * T_GROUP + h]);
      float dd = 0.f, oo = 0.f;
      if (mx != -INFINITY) {
#pragma unroll
        for (int w = 0; w < T_WARPS; ++w) {
          const float mw = st_m[w * T_GROUP + h];
          if (mw != -INFINITY) {
            const float wt = t_ex2(mw - mx);
            dd = fmaf(st_d[w * T_GROUP + h], wt, dd);
            oo = fmaf(st_o[((size_t)w * T_GROUP + h) * T_HD + dim], wt, oo);
          }
        }
      }
      const uint32_t leader = (uint32_t)(h % cs);
      const int entry = (h / cs) * cs + rank;
      t_st_cluster(t_map_rank(smem_u32(c_o + entry * T_HD + dim), leader), oo);
      if (dim == 0) {
        t_st_cluster(t_map_rank(smem_u32(c_m + entry), leader), mx);
        t_st_cluster(t_map_rank(smem_u32(c_d + entry), leader), dd);
      }
      // this warp's 32 dims of head h are in flight to the leader: one release-arrive on its mbarrier
      __syncwarp();
      if (lane == 0) t_mbar_arrive_remote(t_map_rank(smem_u32(merged), leader));
    }
  }
  // Only the head leaders wait (their shared memory is the landing zone); every other CTA is done.
  if (rank >= T_GROUP || rank >= cs) return;
  for (uint32_t spins = 0; !t_mbar_try_wait_cluster(merged, 0); ++spins)
    if (spins > (1u << 26)) __trap();
  if (threadIdx.x >= T_HD) return;
  for (int h = rank; h < T_GROUP; h += cs) {  // this CTA leads heads rank, rank + cs, ...
    const int e0 = (h / cs) * cs;
    float mx = -INFINITY;
    for (int r = 0; r < cs; ++r) mx = fmaxf(mx, c_m[e0 + r]);
    float dd = 0.f, oo = 0.f;
    if (mx != -INFINITY) {
      for (int r = 0; r < cs; ++r) {
        const float mr = c_m[e0 + r];
        if (mr != -INFINITY) {
          const float wt = t_ex2(mr - mx);
          dd = fmaf(c_d[e0 + r], wt, dd);
          oo = fmaf(c_o[(e0 + r) * T_HD + threadIdx.x], wt, oo);
        }
      }
    }
    a.out[((size_t)blockIdx.z * a.nq + kvh * T_GROUP + h) * T_HD + threadIdx.x] = f2bf(dd > 0.f ? __fdividef(oo, dd) : 0.f);
  }
}

// 4-D view of one layer's K (or V) block of the page-first pool: {head dim 128, kv head, slot 16, page};
// box = {64, 1, 16, 1} = one 64-column half of one page of one kv head, 128-byte swizzle.
bool make_page_map(CUtensorMap* map, const bf16* base, int nkv, int64_t stride_page) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {(cuuint64_t)T_HD, (cuuint64_t)nkv, (cuuint64_t)T_PAGE, (cuuint64_t)1 << 20};  // page ids come from the page table
  cuuint64_t strides[3] = {(cuuint64_t)T_HD * 2, (cuuint64_t)nkv * T_HD * 2, (cuuint64_t)stride_page * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)T_PAGE, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t tma_attn_smem(int nslot) {
  return 1024 + (size_t)nslot * T_SLOT + sizeof(float) * (T_MAX_CLUSTER * T_HD + 2 * T_MAX_CLUSTER + 2 * T_WARPS * T_GROUP) +
         sizeof(bf16) * (T_GROUP + 2) * T_HD + sizeof(uint64_t) * T_MAX_SLOTS;
}

struct TmaAttnConfig {
  int nslot = 20;        // 20 x 8 KB: one ring pass up to 320 tokens per CTA (ctx 2560 at 8 CTAs): 6.9 vs 7.3 us at ctx 2304
  int max_cluster = 8;   // 16 (non-portable) measured +4.4 us per launch on B200: opt-in via PK_ATTN_CLUSTER=16
  bool ready = false, ok = false;
};

TmaAttnConfig& tma_attn_config() {
  static thread_local TmaAttnConfig c;
  if (!c.ready) {
    c.ready = true;
    const char* e = getenv("PK_ATTN_SLOTS");
    if (e) c.nslot = atoi(e);
    if (c.nslot < 2) c.nslot = 2;  // the per-warp states (16 KB) alias the ring
    if (c.nslot > T_MAX_SLOTS) c.nslot = T_MAX_SLOTS;
    const char* m = getenv("PK_ATTN_CLUSTER");
    if (m) c.max_cluster = atoi(m);
    if (c.max_cluster < 1) c.max_cluster = 1;
    if (c.max_cluster > T_MAX_CLUSTER) c.max_cluster = T_MAX_CLUSTER;
    const size_t smem = tma_attn_smem(c.nslot);
    c.ok = cudaFuncSetAttribute(decode_attention_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
    if (c.ok && c.max_cluster > 8)
      if (cudaFuncSetAttribute(decode_attention_tma_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
        c.max_cluster = 8;
    if (c.ok && c.max_cluster > 8) {  // can a 16-CTA cluster be co-scheduled at all with this footprint?
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(16, 1, 1);
      q.blockDim = dim3(T_THREADS);
      q.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 16;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      q.attrs = at;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, decode_attention_tma_kernel, &q) != cudaSuccess || n < 1) c.max_cluster = 8;
    }
    cudaGetLastError();
  }
  return c;
}

}  // namespace

// Returns cudaError as int; -2 when this kernel cannot take the call (caller falls back to the round-1 kernels).
int launch_decode_attention_tma(const ClusterAttnArgs& c, int nkv, int bs, cudaStream_t stream) {
  TmaAttnConfig& cfg = tma_attn_config();
  if (!cfg.ok) return -2;
  if ((c.stride_page * 2) % 16 != 0 || ((c.k_off * 2) % 16) != 0 || ((c.v_off * 2) % 16) != 0 ||
      (reinterpret_cast<uintptr_t>(c.kv) & 15) != 0 || (reinterpret_cast<uintptr_t>(c.q) & 7) != 0)
    return -2;
  CUtensorMap mk, mv;
  if (!make_page_map(&mk, c.kv + c.k_off, nkv, c.stride_page) || !make_page_map(&mv, c.kv + c.v_off, nkv, c.stride_page))
    return -2;
  TmaAttnArgs a{};
  a.q = c.q; a.k_new = c.k_new; a.v_new = c.v_new; a.out = c.out; a.kv = c.kv;
  a.k_off = c.k_off; a.v_off = c.v_off; a.stride_page = c.stride_page;
  a.page_indices = c.page_indices; a.page_indptr = c.page_indptr; a.last_page_len = c.last_page_len;
  a.positions = c.positions; a.request_indices = c.request_indices;
  a.qw = c.qw; a.kw = c.kw; a.cosc = c.cosc; a.sinc = c.sinc;
  a.eps = c.eps; a.sm_scale_log2 = c.sm_scale_log2;
  a.nq = c.nq; a.nkv = c.nkv; a.nslot = cfg.nslot;
  // cluster size: as many CTAs per (request, kv head) as keeps the grid within ~one wave of the SMs
  int cs = cfg.max_cluster;
  while (cs > 1 && (int64_t)cs * nkv * bs > (int64_t)sm_count() + sm_count() / 8) cs >>= 1;
  ThreadState& ts = tls();
  ts.launches++;
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(cs, nkv, bs);
  lc.blockDim = dim3(T_THREADS);
  lc.dynamicSmemBytes = tma_attn_smem(cfg.nslot);
  lc.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = ts.pdl ? 1 : 0;
  lc.attrs = at;
  lc.numAttrs = 2;
  return (int)cudaLaunchKernelEx(&lc, decode_attention_tma_kernel, mk, mv, a);
}

}  // namespace pk
