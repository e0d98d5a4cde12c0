// Persistent single-launch decode step (bs = 1, TP = 1): the whole token -- 36 x {qkv GEMV, attention,
// o GEMV, gate_up GEMV + SwiGLU, down GEMV}, lm_head GEMV and the greedy arg-max -- runs in ONE cooperative
// kernel, one CTA per SM.  Replaces the ~530-node CUDA graph of the reference's batch_decode_kernels
// (pegainfer-qwen3-4b/src/batch_decode.rs:82-295) and the 183-launch fused graph of decode_kernels_fused.
//
// Why: profiles/README.md (round 1, v2) -- with one kernel per GEMV every launch boundary costs ~3.5 us of
// idle HBM even with PDL (the next kernel cannot stage its activation vector until the previous grid has
// fully drained), ~0.5 ms of a 1.8 ms weight stream.  Here the weight stream never stops:
//   * a dedicated producer warp per CTA walks the CTA's row slices of ALL weight matrices of the token in
//     order and streams them through a 3 x 64 KB shared-memory ring with 1-D TMA bulk copies (mbarrier
//     complete_tx, L2 evict-first).  It never waits for data dependencies, only for ring slots, so while
//     the consumer warps sit in a grid barrier or run attention the ring keeps filling (192 KB per SM =
//     4.4 us of HBM time chip-wide) with the NEXT phase's weights.
//   * 8 consumer warps: GEMV rows (one row per warp per group, fp32 FMA, warp-shuffle reduce, fused
//     residual-add + RMSNorm prologue, SwiGLU epilogue), the split-KV attention items (two 4-warp teams),
//     and the arg-max.  Phases are separated by a grid barrier (atomic arrive + acquire poll).
// Rounding points are exactly those of decode_kernels_fused / the reference (see gemv.cu, decode_attention.cu).
// Activations written by other CTAs are read with ld.global.cg (L1 is not coherent across SMs).
#include <cooperative_groups.h>

#include "common.cuh"

namespace pk {

constexpr int PCW = 8;                 // consumer warps
constexpr int PTHREADS_C = PCW * 32;   // 256
constexpr int P_HD = 128;
constexpr int P_MAX_STAGES = 4;
constexpr int P_ATT_STATES = 5;        // 4 warp states + the injected new token
constexpr int kPPart = P_HD + 2;

struct PersistLayer {
  const bf16 *qkv, *o, *gate_up, *down, *in_ln, *post_ln, *q_norm, *k_norm;
};

struct PersistArgs {
  const PersistLayer* layers;
  int num_layers, H, qd, kd, I, V, nq, nkv;
  float eps, sm_scale_log2;
  const bf16 *embed, *lm_head, *final_norm, *cosc, *sinc, *zero;
  const uint32_t* token_ids;
  const int *positions, *page_indices, *page_indptr, *last_page_len;
  bf16* kv;
  int64_t layer_stride, kv_block_len, stride_page;
  int page_size;
  bf16 *Ha, *Hb, *q, *k, *v, *attn_out, *attn_proj, *mlp_act, *mlp_out, *logits;
  float* attn_partial;
  int* attn_counters;
  int attn_min_chunk, attn_max_chunks;
  unsigned* bar_counter;  // zeroed before every launch
  float* amax_val;
  int* amax_idx;
  unsigned* amax_ticket;
  int* sample_out;
  int kc, stages;
  unsigned long long* dbg;  // optional: per-phase globaltimer stamps of CTA 0 (profiling builds of the host)
};

__device__ __forceinline__ uint4 ld_cg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_cg8(const void* p) {
  uint2 r;
  asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bounded spins: a protocol bug must trap (launch failure), never hang the GPU
__device__ __forceinline__ void pmbar_wait(uint64_t* bar, uint32_t parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void bar_team(int team) {
  asm volatile("bar.sync %0, 128;" ::"r"(2 + team) : "memory");
}
__device__ __forceinline__ float pex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float pdot8(const uint4& w, const uint4& x, float acc) {
  float a = fmaf(bf16_lo(w.x), bf16_lo(x.x), acc);
  float b = bf16_hi(w.x) * bf16_hi(x.x);
  a = fmaf(bf16_lo(w.y), bf16_lo(x.y), a);
  b = fmaf(bf16_hi(w.y), bf16_hi(x.y), b);
  a = fmaf(bf16_lo(w.z), bf16_lo(x.z), a);
  b = fmaf(bf16_hi(w.z), bf16_hi(x.z), b);
  a = fmaf(bf16_lo(w.w), bf16_lo(x.w), a);
  b = fmaf(bf16_hi(w.w), bf16_hi(x.w), b);
  return a + b;
}
__device__ __forceinline__ void punpack8(const uint4& a, float* f) {
  f[0] = bf16_lo(a.x); f[1] = bf16_hi(a.x); f[2] = bf16_lo(a.y); f[3] = bf16_hi(a.y);
  f[4] = bf16_lo(a.z); f[5] = bf16_hi(a.z); f[6] = bf16_lo(a.w); f[7] = bf16_hi(a.w);
}

// ring cursor shared in form (not in memory) by producer and consumers: both walk the same sequence
struct Ring {
  int s;
  uint32_t ph;
  __device__ __forceinline__ void next(int stages) {
    if (++s == stages) {
      s = 0;
      ph ^= 1u;
    }
  }
};

struct Smem {
  uint8_t* ring;
  bf16* xs;          // activation vector of the running GEMV (aliased by the attention scratch)
  uint64_t* full;
  uint64_t* empty;
  float* red;        // 64 floats
  int slot_bytes, seg_bytes;
};

// ---------------------------------------------------------------- producer side of one GEMV phase
__device__ __forceinline__ void produce_gemv(const PersistArgs& a, const Smem& sm, Ring& r, const bf16* W, int M,
                                             int K, bool swiglu, int lane, uint64_t pol) {
  const int G = gridDim.x;
  const int rpg = swiglu ? PCW / 2 : PCW;
  const int r0 = (int)(((int64_t)blockIdx.x * M) / G), r1 = (int)(((int64_t)(blockIdx.x + 1) * M) / G);
  const int groups = (r1 - r0 + rpg - 1) / rpg;
  const int chunks = (K + a.kc - 1) / a.kc;
  for (int g = 0; g < groups; ++g) {
    int src = -1;
    if (lane < PCW) {
      const int out_row = r0 + g * rpg + (swiglu ? (lane & 3) : lane);
      if (out_row < r1) src = swiglu ? (lane < 4 ? out_row : M + out_row) : out_row;
    }
    const int nvalid = __popc(__ballot_sync(0xffffffffu, src >= 0));
    for (int c = 0; c < chunks; ++c) {
      const int k0 = c * a.kc;
      const uint32_t bytes = (uint32_t)(min(a.kc, K - k0) * 2);
      if (lane == 0) {
        pmbar_wait(sm.empty + r.s, r.ph ^ 1u);
        mbar_expect_tx(sm.full + r.s, bytes * (uint32_t)nvalid);
      }
      __syncwarp();
      if (src >= 0)
        bulk_g2s(sm.ring + (size_t)r.s * sm.slot_bytes + (size_t)lane * sm.seg_bytes, W + (size_t)src * K + k0,
                 bytes, sm.full + r.s, pol);
      r.next(a.stages);
    }
  }
}

// ---------------------------------------------------------------- activation staging (consumer warps)
__device__ __forceinline__ void stage_x_plain(const Smem& sm, const bf16* x, int K, int tid) {
  const int nv = K >> 3;
  for (int i = tid; i < nv; i += PTHREADS_C) reinterpret_cast<uint4*>(sm.xs)[i] = ld_cg16(reinterpret_cast<const uint4*>(x) + i);
  bar_consumers();
}
// x = bf16((h + r) * rsqrt(mean((h+r)^2) + eps) * w); CTA 0 stores bf16(h + r) to hidden_out
__device__ __forceinline__ void stage_x_norm(const PersistArgs& a, const Smem& sm, const bf16* h, const bf16* res,
                                             const bf16* w, bf16* hidden_out, int K, int tid, int warp, int lane) {
  constexpr int kMaxVec = 3;  // K <= 3 * 256 * 8 = 6144
  const int nv = K >> 3;
  float v[kMaxVec][8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = tid + j * PTHREADS_C;
    if (i < nv) {
      const uint4 hh = ld_cg16(reinterpret_cast<const uint4*>(h) + i), rr = ld_cg16(reinterpret_cast<const uint4*>(res) + i);
      v[j][0] = bf16_lo(hh.x) + bf16_lo(rr.x); v[j][1] = bf16_hi(hh.x) + bf16_hi(rr.x);
      v[j][2] = bf16_lo(hh.y) + bf16_lo(rr.y); v[j][3] = bf16_hi(hh.y) + bf16_hi(rr.y);
      v[j][4] = bf16_lo(hh.z) + bf16_lo(rr.z); v[j][5] = bf16_hi(hh.z) + bf16_hi(rr.z);
      v[j][6] = bf16_lo(hh.w) + bf16_lo(rr.w); v[j][7] = bf16_hi(hh.w) + bf16_hi(rr.w);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(v[j][e], v[j][e], ss);
    }
  }
  ss = warp_sum(ss);
  if (lane == 0) sm.red[warp] = ss;
  bar_consumers();
  float tot = 0.f;
#pragma unroll
  for (int w8 = 0; w8 < PCW; ++w8) tot += sm.red[w8];
  const float rinv = rsqrtf(tot / (float)K + a.eps);
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int i = tid + j * PTHREADS_C;
    if (i < nv) {
      const uint4 g = reinterpret_cast<const uint4*>(w)[i];
      uint4 o;
      o.x = pack_bf16(v[j][0] * rinv * bf16_lo(g.x), v[j][1] * rinv * bf16_hi(g.x));
      o.y = pack_bf16(v[j][2] * rinv * bf16_lo(g.y), v[j][3] * rinv * bf16_hi(g.y));
      o.z = pack_bf16(v[j][4] * rinv * bf16_lo(g.z), v[j][5] * rinv * bf16_hi(g.z));
      o.w = pack_bf16(v[j][6] * rinv * bf16_lo(g.w), v[j][7] * rinv * bf16_hi(g.w));
      reinterpret_cast<uint4*>(sm.xs)[i] = o;
      if (blockIdx.x == 0) {
        uint4 hs;
        hs.x = pack_bf16(v[j][0], v[j][1]);
        hs.y = pack_bf16(v[j][2], v[j][3]);
        hs.z = pack_bf16(v[j][4], v[j][5]);
        hs.w = pack_bf16(v[j][6], v[j][7]);
        reinterpret_cast<uint4*>(hidden_out)[i] = hs;
      }
    }
  }
  bar_consumers();
}

// ---------------------------------------------------------------- consumer side of one GEMV phase
// EPI 0: y[row] = bf16(acc) routed to up to 3 segments; EPI 1: SwiGLU; EPI 2: logits + running arg-max
template <int EPI>
__device__ __forceinline__ void consume_gemv(const PersistArgs& a, const Smem& sm, Ring& r, int M, int K, bf16* y0,
                                             bf16* y1, bf16* y2, int e0, int e1, int warp, int lane, float* best_v,
                                             int* best_i) {
  const int G = gridDim.x;
  constexpr int rpg = EPI == 1 ? PCW / 2 : PCW;
  const int r0 = (int)(((int64_t)blockIdx.x * M) / G), r1 = (int)(((int64_t)(blockIdx.x + 1) * M) / G);
  const int groups = (r1 - r0 + rpg - 1) / rpg;
  const int chunks = (K + a.kc - 1) / a.kc;
  for (int g = 0; g < groups; ++g) {
    const int out_row = r0 + g * rpg + (EPI == 1 ? (warp & 3) : warp);
    const bool has_row = out_row < r1;
    float acc4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < chunks; ++c) {
      pmbar_wait(sm.full + r.s, r.ph);
      if (has_row) {
        const int k0 = c * a.kc;
        const int nvec = min(a.kc, K - k0) >> 3;
        const uint4* wseg = reinterpret_cast<const uint4*>(sm.ring + (size_t)r.s * sm.slot_bytes + (size_t)warp * sm.seg_bytes);
        const uint4* xp = reinterpret_cast<const uint4*>(sm.xs + k0);
        int v0 = 0;
        for (; v0 + 128 <= nvec; v0 += 128) {
          uint4 wv[4], xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) wv[i] = wseg[v0 + lane + 32 * i];
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = xp[v0 + lane + 32 * i];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc4[i] = pdot8(wv[i], xv[i], acc4[i]);
        }
        for (int i = v0 + lane; i < nvec; i += 32) acc4[0] = pdot8(wseg[i], xp[i], acc4[0]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.empty + r.s);
      r.next(a.stages);
    }
    float acc = warp_sum((acc4[0] + acc4[1]) + (acc4[2] + acc4[3]));
    if (EPI == 0) {
      if (has_row && lane == 0) {
        if (out_row < e0) y0[out_row] = f2bf(acc);
        else if (out_row < e1) y1[out_row - e0] = f2bf(acc);
        else y2[out_row - e1] = f2bf(acc);
      }
    } else if (EPI == 1) {
      float* sw = sm.red + 16 + (g & 1) * 8;
      if (warp >= 4 && lane == 0) sw[warp - 4] = acc;
      bar_consumers();
      if (warp < 4 && has_row && lane == 0) {
        const float gt = round_bf16(acc), up = round_bf16(sw[warp]);
        y0[out_row] = f2bf(gt / (1.0f + expf(-gt)) * up);
      }
    } else {
      if (has_row) {
        const bf16 lg = f2bf(acc);
        if (lane == 0) y0[out_row] = lg;
        const float v = bf2f(lg);
        if (v > *best_v || (v == *best_v && out_row < *best_i)) {
          *best_v = v;
          *best_i = out_row;
        }
      }
    }
  }
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define PK_STAMP(slot)                                                                   \
  do {                                                                                   \
    if (a.dbg && tid == 0 && blockIdx.x == 0 && li < 4) a.dbg[li * 16 + (slot)] = gtimer(); \
  } while (0)

// ---------------------------------------------------------------- grid barrier (consumer warps only)
__device__ __forceinline__ void grid_barrier(const PersistArgs& a, unsigned& epoch, int tid) {
  __threadfence();
  bar_consumers();
  if (tid == 0) {
    ++epoch;
    atomicAdd(a.bar_counter, 1u);
    const unsigned target = epoch * gridDim.x;
    const long long t0 = clock64();
    while (ld_acquire_gpu(a.bar_counter) < target) {
      if (clock64() - t0 > 4000000000LL) __trap();
    }
  } else {
    ++epoch;
  }
  bar_consumers();
}

// ---------------------------------------------------------------- attention item (one 4-warp team)
// Same arithmetic as decode_attention_kernel's fused path: QK-norm + RoPE of q (and of the step's k in the
// chunk that owns the new position, which also appends k/v to the cache), online softmax over the chunk,
// team merge, then fp32 partial + ticket merge across chunks.
__device__ __forceinline__ void team_norm_rope(const bf16* src, const bf16* w, const bf16* cosc, const bf16* sinc,
                                               int pos, float eps, bf16* dst, int lane) {
  const uint2 raw = ld_cg8(reinterpret_cast<const uint2*>(src) + lane);
  const float v[4] = {bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y)};
  float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)P_HD + eps);
  const uint2 wr = reinterpret_cast<const uint2*>(w)[lane];
  const float wv[4] = {bf16_lo(wr.x), bf16_hi(wr.x), bf16_lo(wr.y), bf16_hi(wr.y)};
  const int ci = (lane & 15) * 4;
  const uint2 cr = reinterpret_cast<const uint2*>(cosc + (size_t)pos * P_HD + ci)[0];
  const uint2 sr = reinterpret_cast<const uint2*>(sinc + (size_t)pos * P_HD + ci)[0];
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

struct AttScratch {  // per team, aliased on the xs region
  float o[P_ATT_STATES][4][P_HD];
  float m[P_ATT_STATES][4], d[P_ATT_STATES][4];
  bf16 q_s[4][P_HD];
  bf16 k_s[P_HD], v_s[P_HD];
  float sw[64 * 4 + 8];
  int last;
};

constexpr int kU = 8;  // K/V rows per half-warp per 64-token round

// Per-step, per-lane addressing of round 0 of this team's attention item: identical for every layer (the
// page table is fixed within a step), so it is computed once and the K/V rows of layer l are requested
// BEFORE the qkv GEMV of layer l -- their HBM latency (several us while the weight stream saturates the
// memory system) hides behind that GEMV instead of sitting on the attention phase's critical path.
struct KvPrefetch {
  int64_t off[kU];   // element offset inside a layer's K (or V) plane; < 0: masked token
};

__device__ __forceinline__ void attention_item(const PersistArgs& a, const PersistLayer& L, int layer, int chunk_idx,
                                               int kvh, int chunk, int nchunks, int len, int pos,
                                               AttScratch* S, int team, int twarp, int lane, bool prefetched,
                                               const KvPrefetch& pf) {
  constexpr int GROUP = 4;
  const int half = lane >> 4, l16 = lane & 15;
  const int t128 = twarp * 32 + lane;
  const int lo = chunk_idx * chunk, hi = min(len, lo + chunk);
  const int* pages = a.page_indices + __ldg(a.page_indptr);
  const int64_t k_off = (int64_t)layer * a.layer_stride, v_off = k_off + a.kv_block_len;

  // q heads (and the new k/v when this chunk owns the position)
  team_norm_rope(a.q + (size_t)(kvh * GROUP + twarp) * P_HD, L.q_norm, a.cosc, a.sinc, pos, a.eps, S->q_s[twarp], lane);
  const bool inject = pos >= lo && pos < hi;
  if (inject && twarp == 0) {
    team_norm_rope(a.k + (size_t)kvh * P_HD, L.k_norm, a.cosc, a.sinc, pos, a.eps, S->k_s, lane);
    reinterpret_cast<uint2*>(S->v_s)[lane] = ld_cg8(reinterpret_cast<const uint2*>(a.v + (size_t)kvh * P_HD) + lane);
    __syncwarp();
    const int page = pages[pos / a.page_size], slot = pos % a.page_size;
    const int64_t dst = (int64_t)page * a.stride_page + ((int64_t)slot * a.nkv + kvh) * P_HD;
    reinterpret_cast<uint2*>(a.kv + k_off + dst)[lane] = reinterpret_cast<uint2*>(S->k_s)[lane];
    reinterpret_cast<uint2*>(a.kv + v_off + dst)[lane] = reinterpret_cast<uint2*>(S->v_s)[lane];
  }
  bar_team(team);
  float qf[GROUP][8];
#pragma unroll
  for (int h = 0; h < GROUP; ++h) punpack8(reinterpret_cast<const uint4*>(S->q_s[h])[l16], qf[h]);

  float m[GROUP], d[GROUP], o[GROUP][8];
#pragma unroll
  for (int h = 0; h < GROUP; ++h) {
    m[h] = -INFINITY;
    d[h] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[h][j] = 0.f;
  }
  const int new_pos = inject ? pos : -1;
  const bf16* kbase = a.kv + k_off + (int64_t)kvh * P_HD + l16 * 8;
  const bf16* vbase = a.kv + v_off + (int64_t)kvh * P_HD + l16 * 8;
  for (int round = lo; round < hi; round += 8 * kU) {
    const int base = round + twarp * 2 + half;
    uint4 kr[kU], vr[kU];
    bool ok[kU];
    if (prefetched && round == lo) {  // addresses precomputed once per step, lines already in L2
      const bf16* kl = a.kv + k_off;
      const bf16* vl = a.kv + v_off;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        ok[u] = pf.off[u] >= 0;
        kr[u] = make_uint4(0, 0, 0, 0);
        vr[u] = make_uint4(0, 0, 0, 0);
        if (ok[u]) {
          kr[u] = ld_cg16(kl + pf.off[u]);
          vr[u] = ld_cg16(vl + pf.off[u]);
        }
      }
    } else {
      int pg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pg[i] = (round + 16 * i < hi) ? __ldg(pages + (round >> 4) + i) : 0;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int t = base + u * 8;
        ok[u] = t < hi && t != new_pos;
        kr[u] = make_uint4(0, 0, 0, 0);
        vr[u] = make_uint4(0, 0, 0, 0);
        if (ok[u]) {
          const int64_t off = (int64_t)pg[u >> 1] * a.stride_page + (int64_t)(t & 15) * a.nkv * P_HD;
          kr[u] = ld_cg16(kbase + off);  // the cache is rewritten every step: never through L1
          vr[u] = ld_cg16(vbase + off);
        }
      }
    }
    float s[GROUP][kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float kf[8];
      punpack8(kr[u], kf);
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
      for (int u = 0; u < kU; ++u) mn = fmaxf(mn, s[h][u]);
      if (mn == -INFINITY) {
#pragma unroll
        for (int u = 0; u < kU; ++u) s[h][u] = 0.f;
        continue;
      }
      const float sc = pex2(m[h] - mn);
      m[h] = mn;
      d[h] *= sc;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[h][j] *= sc;
#pragma unroll
      for (int u = 0; u < kU; ++u) s[h][u] = pex2(s[h][u] - mn);
#pragma unroll
      for (int u = 0; u < kU; ++u) d[h] += s[h][u];
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float vf[8];
      punpack8(vr[u], vf);
#pragma unroll
      for (int h = 0; h < GROUP; ++h)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[h][j] = fmaf(s[h][u], vf[j], o[h][j]);
    }
  }
  // merge the two half-warps of this warp with shuffles, then the 4 warps (+ new token) through smem
#pragma unroll
  for (int h = 0; h < GROUP; ++h) {
    const float om = __shfl_xor_sync(0xffffffffu, m[h], 16), od = __shfl_xor_sync(0xffffffffu, d[h], 16);
    const float mn = fmaxf(m[h], om);
    const float wa = mn == -INFINITY ? 0.f : pex2(m[h] - mn), wb = mn == -INFINITY ? 0.f : pex2(om - mn);
    d[h] = d[h] * wa + od * wb;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float oo = __shfl_xor_sync(0xffffffffu, o[h][j], 16);
      o[h][j] = o[h][j] * wa + oo * wb;
    }
    m[h] = mn;
    if (half == 0) {
      if (l16 == 0) {
        S->m[twarp][h] = m[h];
        S->d[twarp][h] = d[h];
      }
      float4* dst = reinterpret_cast<float4*>(&S->o[twarp][h][l16 * 8]);
      dst[0] = make_float4(o[h][0], o[h][1], o[h][2], o[h][3]);
      dst[1] = make_float4(o[h][4], o[h][5], o[h][6], o[h][7]);
    }
  }
  if (inject) {
    const int h = twarp;
    const uint2 qr = reinterpret_cast<const uint2*>(S->q_s[h])[lane];
    const uint2 kr2 = reinterpret_cast<const uint2*>(S->k_s)[lane];
    float p = bf16_lo(qr.x) * bf16_lo(kr2.x);
    p = fmaf(bf16_hi(qr.x), bf16_hi(kr2.x), p);
    p = fmaf(bf16_lo(qr.y), bf16_lo(kr2.y), p);
    p = fmaf(bf16_hi(qr.y), bf16_hi(kr2.y), p);
    p = warp_sum(p);
    if (lane == 0) {
      S->m[4][h] = p * a.sm_scale_log2;
      S->d[4][h] = 1.f;
    }
    const uint2 vr2 = reinterpret_cast<const uint2*>(S->v_s)[lane];
    *reinterpret_cast<float4*>(&S->o[4][h][lane * 4]) =
        make_float4(bf16_lo(vr2.x), bf16_hi(vr2.x), bf16_lo(vr2.y), bf16_hi(vr2.y));
  }
  bar_team(team);
  const int nstates = inject ? 5 : 4;
  float M4[GROUP], D4[GROUP], O4[GROUP];
#pragma unroll
  for (int h = 0; h < GROUP; ++h) {
    float mx = -INFINITY;
    for (int i = 0; i < nstates; ++i) mx = fmaxf(mx, S->m[i][h]);
    float dd = 0.f, oo = 0.f;
    if (mx != -INFINITY) {
      for (int i = 0; i < nstates; ++i) {
        const float w = pex2(S->m[i][h] - mx);
        dd = fmaf(S->d[i][h], w, dd);
        oo = fmaf(S->o[i][h][t128], w, oo);
      }
    }
    M4[h] = mx; D4[h] = dd; O4[h] = oo;
  }
  if (nchunks == 1) {
#pragma unroll
    for (int h = 0; h < GROUP; ++h) a.attn_out[(size_t)(kvh * GROUP + h) * P_HD + t128] = f2bf(__fdividef(O4[h], D4[h]));
    bar_team(team);
    return;
  }
#pragma unroll
  for (int h = 0; h < GROUP; ++h) {
    float* p = a.attn_partial + ((size_t)chunk_idx * a.nq + kvh * GROUP + h) * kPPart;
    p[t128] = O4[h];
    if (t128 == 0) {
      p[P_HD] = M4[h];
      p[P_HD + 1] = D4[h];
    }
  }
  __threadfence();
  bar_team(team);
  if (t128 == 0) S->last = (atomicAdd(a.attn_counters + kvh, 1) == nchunks - 1);
  bar_team(team);
  if (S->last) {
    __threadfence();
    {
      const int h = twarp;
      const float* p0 = a.attn_partial + (size_t)(kvh * GROUP + h) * kPPart;
      const size_t cs = (size_t)a.nq * kPPart;
      const int c0 = lane, c1 = lane + 32;
      const float m0 = c0 < nchunks ? __ldcg(p0 + c0 * cs + P_HD) : -INFINITY;
      const float m1 = c1 < nchunks ? __ldcg(p0 + c1 * cs + P_HD) : -INFINITY;
      const float d0 = c0 < nchunks ? __ldcg(p0 + c0 * cs + P_HD + 1) : 0.f;
      const float d1 = c1 < nchunks ? __ldcg(p0 + c1 * cs + P_HD + 1) : 0.f;
      const float mx = warp_max(fmaxf(m0, m1));
      const float w0 = pex2(m0 - mx), w1 = pex2(m1 - mx);
      S->sw[c0 * GROUP + h] = w0;
      S->sw[c1 * GROUP + h] = w1;
      const float dd = warp_sum(d0 * w0 + d1 * w1);
      if (lane == 0) S->sw[256 + h] = dd;
    }
    bar_team(team);
#pragma unroll
    for (int h = 0; h < GROUP; ++h) {
      const float* p0 = a.attn_partial + (size_t)(kvh * GROUP + h) * kPPart + t128;
      const size_t cs = (size_t)a.nq * kPPart;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int c = 0;
      for (; c + 4 <= nchunks; c += 4) {
        const float v0 = __ldcg(p0 + (c + 0) * cs), v1 = __ldcg(p0 + (c + 1) * cs);
        const float v2 = __ldcg(p0 + (c + 2) * cs), v3 = __ldcg(p0 + (c + 3) * cs);
        a0 = fmaf(v0, S->sw[(c + 0) * GROUP + h], a0);
        a1 = fmaf(v1, S->sw[(c + 1) * GROUP + h], a1);
        a2 = fmaf(v2, S->sw[(c + 2) * GROUP + h], a2);
        a3 = fmaf(v3, S->sw[(c + 3) * GROUP + h], a3);
      }
      for (; c < nchunks; ++c) a0 = fmaf(__ldcg(p0 + c * cs), S->sw[c * GROUP + h], a0);
      a.attn_out[(size_t)(kvh * GROUP + h) * P_HD + t128] = f2bf(__fdividef((a0 + a1) + (a2 + a3), S->sw[256 + h]));
    }
    if (t128 == 0) a.attn_counters[kvh] = 0;
  }
  bar_team(team);
}

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(PTHREADS_C + 32, 1) decode_step_persistent_kernel(const PersistArgs a) {
  extern __shared__ __align__(128) uint8_t psmem[];
  Smem sm;
  sm.seg_bytes = a.kc * 2;
  sm.slot_bytes = PCW * sm.seg_bytes;
  sm.ring = psmem;
  uint8_t* after = psmem + (size_t)a.stages * sm.slot_bytes;
  sm.xs = reinterpret_cast<bf16*>(after);
  constexpr size_t kXsBytes = 2 * ((sizeof(AttScratch) + 127) & ~(size_t)127);
  sm.full = reinterpret_cast<uint64_t*>(after + kXsBytes);
  sm.empty = sm.full + P_MAX_STAGES;
  sm.red = reinterpret_cast<float*>(sm.empty + P_MAX_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(sm.full + s, 1);
      mbar_init(sm.empty + s, PCW);
    }
    mbar_fence_init();
  }
  __syncthreads();
  Ring ring{0, 0u};

  if (warp == PCW) {
    // ===== producer: every weight byte of the token, in consumption order =====
    const uint64_t pol = l2_evict_first_policy();
    for (int li = 0; li < a.num_layers; ++li) {
      const PersistLayer L = a.layers[li];
      produce_gemv(a, sm, ring, L.qkv, a.qd + 2 * a.kd, a.H, false, lane, pol);
      produce_gemv(a, sm, ring, L.o, a.H, a.qd, false, lane, pol);
      produce_gemv(a, sm, ring, L.gate_up, a.I, a.H, true, lane, pol);
      produce_gemv(a, sm, ring, L.down, a.H, a.I, false, lane, pol);
    }
    produce_gemv(a, sm, ring, a.lm_head, a.V, a.H, false, lane, pol);
    return;
  }

  // ===== consumers =====
  unsigned epoch = 0;
  bf16* Hcur = a.Ha;
  bf16* Hnext = a.Hb;
  const uint32_t tok = a.token_ids[0];
  const bf16* h_in = a.embed + (size_t)tok * a.H;  // layer 0 reads the embedding row directly
  const bf16* residual = a.zero;
  const int npages = a.page_indptr[1] - a.page_indptr[0];
  const int len = npages <= 0 ? 0 : (npages - 1) * a.page_size + a.last_page_len[0];
  int chunk = max(a.attn_min_chunk, (len + a.attn_max_chunks - 1) / a.attn_max_chunks);
  chunk = (chunk + 15) & ~15;
  const int nchunks = max(1, (len + chunk - 1) / chunk);
  const int team = warp >> 2, twarp = warp & 3;
  AttScratch* scratch = reinterpret_cast<AttScratch*>(reinterpret_cast<uint8_t*>(sm.xs) +
                                                      (size_t)team * ((sizeof(AttScratch) + 127) & ~(size_t)127));

  // this team's first attention item and its round-0 addressing (same for every layer)
  const int pos = a.positions[0];
  const int item0 = blockIdx.x * 2 + team;
  const bool has_item0 = item0 < nchunks * a.nkv;
  KvPrefetch pf;
  {
    const int half = lane >> 4, l16 = lane & 15;
    const int cidx = item0 / a.nkv, kvh0 = item0 % a.nkv;
    const int lo = cidx * chunk, hi = min(len, lo + chunk);
    const int* pages = a.page_indices + a.page_indptr[0];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int t = lo + twarp * 2 + half + u * 8;
      pf.off[u] = -1;
      if (has_item0 && t < hi && t != pos)
        pf.off[u] = (int64_t)__ldg(pages + (t >> 4)) * a.stride_page + (int64_t)(t & 15) * a.nkv * P_HD +
                    (int64_t)kvh0 * P_HD + l16 * 8;
    }
  }

  for (int li = 0; li < a.num_layers; ++li) {
    const PersistLayer L = a.layers[li];
    // pull this layer's K/V rows into L2 now (no registers held: 9 warps x 168 registers is the budget);
    // the attention phase after the qkv GEMV and the grid barrier then hits L2 instead of queueing
    // behind the saturated weight stream in DRAM
    {
      const bf16* kl = a.kv + (int64_t)li * a.layer_stride;
      const bf16* vl = kl + a.kv_block_len;
      if ((lane & 7) == 0) {  // one prefetch per 128-byte line (a 256-byte row = 2 lines)
#pragma unroll
        for (int u = 0; u < kU; ++u)
          if (pf.off[u] >= 0) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(kl + pf.off[u]));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vl + pf.off[u]));
          }
      }
    }
    // P1: q|k|v = W_qkv . RMSNorm(h + residual); Hnext = h + residual
    PK_STAMP(0);
    stage_x_norm(a, sm, h_in, residual, L.in_ln, Hnext, a.H, tid, warp, lane);
    PK_STAMP(1);
    consume_gemv<0>(a, sm, ring, a.qd + 2 * a.kd, a.H, a.q, a.k, a.v, a.qd, a.qd + a.kd, warp, lane, nullptr, nullptr);
    PK_STAMP(2);
    grid_barrier(a, epoch, tid);
    PK_STAMP(3);
    // P2: attention items (chunk, kv head), two teams per CTA
    for (int item = item0; item < nchunks * a.nkv; item += gridDim.x * 2)
      attention_item(a, L, li, item / a.nkv, item % a.nkv, chunk, nchunks, len, pos, scratch, team, twarp, lane,
                     item == item0, pf);
    PK_STAMP(4);
    grid_barrier(a, epoch, tid);
    PK_STAMP(5);
    // P3: attn_proj = W_o . attn_out
    stage_x_plain(sm, a.attn_out, a.qd, tid);
    PK_STAMP(6);
    consume_gemv<0>(a, sm, ring, a.H, a.qd, a.attn_proj, nullptr, nullptr, a.H, a.H, warp, lane, nullptr, nullptr);
    PK_STAMP(7);
    grid_barrier(a, epoch, tid);
    PK_STAMP(8);
    // P4: act = SwiGLU(W_gate_up . RMSNorm(Hnext + attn_proj)); Hcur = Hnext + attn_proj
    stage_x_norm(a, sm, Hnext, a.attn_proj, L.post_ln, Hcur, a.H, tid, warp, lane);
    PK_STAMP(9);
    consume_gemv<1>(a, sm, ring, a.I, a.H, a.mlp_act, nullptr, nullptr, 0, 0, warp, lane, nullptr, nullptr);
    PK_STAMP(10);
    grid_barrier(a, epoch, tid);
    PK_STAMP(11);
    // P5: mlp_out = W_down . act
    stage_x_plain(sm, a.mlp_act, a.I, tid);
    PK_STAMP(12);
    consume_gemv<0>(a, sm, ring, a.H, a.I, a.mlp_out, nullptr, nullptr, a.H, a.H, warp, lane, nullptr, nullptr);
    PK_STAMP(13);
    grid_barrier(a, epoch, tid);
    PK_STAMP(14);
    h_in = Hcur;
    residual = a.mlp_out;
  }
  // lm_head on RMSNorm(h + mlp_out) with the final norm weight, arg-max fused
  stage_x_norm(a, sm, h_in, residual, a.final_norm, Hnext, a.H, tid, warp, lane);
  float best_v = -INFINITY;
  int best_i = 0x7fffffff;
  consume_gemv<2>(a, sm, ring, a.V, a.H, a.logits, nullptr, nullptr, a.V, a.V, warp, lane, &best_v, &best_i);
  // CTA arg-max (all lanes of a warp hold the same candidate), then ticket reduce across CTAs
  if (lane == 0) {
    sm.red[warp] = best_v;
    reinterpret_cast<int*>(sm.red)[8 + warp] = best_i;
  }
  bar_consumers();
  if (tid == 0) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int w = 0; w < PCW; ++w) {
      const float v = sm.red[w];
      const int i = reinterpret_cast<int*>(sm.red)[8 + w];
      if (v > bv || (v == bv && i < bi)) {
        bv = v;
        bi = i;
      }
    }
    a.amax_val[blockIdx.x] = bv;
    a.amax_idx[blockIdx.x] = bi;
    __threadfence();
    const unsigned t = atomicAdd(a.amax_ticket, 1u);
    if (t == gridDim.x - 1) {
      __threadfence();
      float fv = -INFINITY;
      int fi = 0x7fffffff;
      for (unsigned c = 0; c < gridDim.x; ++c) {
        const float v = __ldcg(a.amax_val + c);
        const int i = __ldcg(a.amax_idx + c);
        if (v > fv || (v == fv && i < fi)) {
          fv = v;
          fi = i;
        }
      }
      a.sample_out[0] = fi == 0x7fffffff ? 0 : fi;
      *a.amax_ticket = 0;
    }
  }
}

}  // namespace pk

// ---------------------------------------------------------------- C ABI
extern "C" int pk_b200_decode_step_persistent(const pk_b200_decode_step_args* g, pk_stream stream) {
  using namespace pk;
  if (!g || g->head_dim != 128 || g->num_kv_heads <= 0 || g->num_q_heads != 4 * g->num_kv_heads) return -1;
  if (g->hidden_size % 8 || g->hidden_size > 6144 || g->intermediate_size % 8 || g->vocab_size <= 0) return -1;
  PersistArgs a{};
  a.layers = reinterpret_cast<const PersistLayer*>(g->layers_dev);
  a.num_layers = g->num_layers;
  a.H = g->hidden_size; a.qd = g->num_q_heads * 128; a.kd = g->num_kv_heads * 128; a.I = g->intermediate_size;
  a.V = g->vocab_size; a.nq = g->num_q_heads; a.nkv = g->num_kv_heads;
  a.eps = g->rms_eps; a.sm_scale_log2 = g->sm_scale * 1.44269504088896340736f;
  a.embed = (const bf16*)g->embed; a.lm_head = (const bf16*)g->lm_head; a.final_norm = (const bf16*)g->final_norm;
  a.cosc = (const bf16*)g->cos_cache; a.sinc = (const bf16*)g->sin_cache; a.zero = (const bf16*)g->zero_residual;
  a.token_ids = g->token_ids; a.positions = g->positions; a.page_indices = g->page_indices;
  a.page_indptr = g->page_indptr; a.last_page_len = g->last_page_len;
  a.kv = (bf16*)g->kv_data; a.layer_stride = g->layer_stride; a.kv_block_len = g->kv_block_len;
  a.stride_page = g->stride_page; a.page_size = g->page_size;
  a.Ha = (bf16*)g->hidden_a; a.Hb = (bf16*)g->hidden_b; a.q = (bf16*)g->q; a.k = (bf16*)g->k; a.v = (bf16*)g->v;
  a.attn_out = (bf16*)g->attn_out; a.attn_proj = (bf16*)g->attn_proj; a.mlp_act = (bf16*)g->mlp_act;
  a.mlp_out = (bf16*)g->mlp_out; a.logits = (bf16*)g->logits;
  a.attn_partial = g->attn_partial; a.attn_counters = g->attn_counters;
  a.attn_min_chunk = 64;
  a.attn_max_chunks = g->attn_max_chunks > 0 ? (g->attn_max_chunks > 64 ? 64 : g->attn_max_chunks) : 1;
  unsigned* ctl = reinterpret_cast<unsigned*>(g->sync_scratch);  // [0] barrier, [1] argmax ticket, then partials
  a.bar_counter = ctl;
  a.amax_ticket = ctl + 1;
  a.amax_val = reinterpret_cast<float*>(ctl + 16);
  a.amax_idx = reinterpret_cast<int*>(ctl + 16 + 256);
  a.sample_out = g->sample_out;
  a.dbg = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(g->sync_scratch) + 4096);  // [4 layers][16]
  const int sms = sm_count();
  if (sms > 256 || g->page_size != 16) return -1;
  a.kc = 4096;
  a.stages = 3;
  const size_t xs_bytes = 2 * ((sizeof(AttScratch) + 127) & ~(size_t)127);
  if ((size_t)a.I * 2 > xs_bytes || (size_t)a.qd * 2 > xs_bytes) return -1;
  auto smem_for = [&](int stages, int kc) {
    return (size_t)stages * PCW * kc * 2 + xs_bytes + 2 * P_MAX_STAGES * sizeof(uint64_t) + 64 * sizeof(float);
  };
  while (smem_for(a.stages, a.kc) > 226 * 1024 && a.kc > 1024) a.kc -= 512;
  const size_t smem = smem_for(a.stages, a.kc);
  static thread_local size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(decode_step_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return -2;
    configured = smem;
  }
  if (cudaMemsetAsync(ctl, 0, 4, stream) != cudaSuccess) return -3;  // barrier counter
  tls().launches++;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(PTHREADS_C + 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, decode_step_persistent_kernel, a);
}
