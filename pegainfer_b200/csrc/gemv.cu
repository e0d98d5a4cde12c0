// Decode GEMV: Y[M, N<=4] = W[M, K] x X[K, N] -- the op that is ~90 % of decode time
// (weights: 8.04 GB/token for Qwen3-4B).  Replaces `gemm_graphsafe_cuda` (csrc/linear.cu:65-78,
// a cuBLAS call in the reference) and adds the fused prologue/epilogue variants the B200 host
// mirror uses.  bf16 x bf16 -> fp32 accumulate -> one bf16 rounding, as CUBLAS_COMPUTE_32F.
//
// HBM-bound design (roofline: 2*M*K bytes / HBM bandwidth):
//  * grid = #SMs (x ctas_per_sm); CTA c owns the contiguous row range [c*M/G, (c+1)*M/G) so every
//    SM streams the same number of bytes (+-1 row).
//  * a dedicated producer warp streams the CTA's rows through a `stages`-deep shared-memory ring
//    with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), L2 evict-first (weights are
//    read exactly once per token); 8 consumer warps own one row each per group and reduce with
//    warp shuffles.  Bytes in flight per SM = stages * 16 KB, independent of register count.
//  * PDL: weights do not depend on the previous kernel, so the producer starts streaming BEFORE
//    griddepcontrol.wait; only the activation vector waits.  The kernel triggers its dependents
//    at entry so the next GEMV's weight prefetch overlaps this kernel's tail.
//  * optional prologue: x = RMSNorm(hidden + residual) * w (fused_add_rms_norm semantics, same
//    rounding points as csrc/flashinfer_norm.cu:71-105); optional epilogue: SwiGLU
//    (csrc/fused_proj.cu:44-63) on interleaved gate/up row groups.
#include <cstdlib>

#include "common.cuh"

namespace pk {

constexpr int kCW = 8;            // consumer warps (rows per group)
constexpr int kMaxStages = 12;
constexpr int kConsumerThreads = kCW * 32;

struct GemvArgs {
  const bf16* W;
  const bf16* X;  // [N, K] activations, or hidden_in when x_mode == 1
  bf16* Y[3];
  int seg_end[3];  // cumulative row ends of the up-to-3 output segments (seg_end[2] == M)
  int M, K;
  int x_mode;
  const bf16* residual;
  const bf16* norm_w;
  float eps;
  bf16* hidden_out;
  bf16* normed_out;
  int epi;      // 0 plain, 1 SwiGLU (W has 2*M rows: gate rows then up rows), 2 push partial rows to all TP peers,
                // 3 all-reduce of this CTA's rows inside the kernel (8-byte LL lines over NVLink)
  TpDev tp;     // x_mode 2 / epi 2 / epi 3: peer staging + flags
  const uint32_t* tp_step;  // epi 3: decode step counter (device word, identical on every rank)
  int tp_op;                // epi 3: collective index inside the step
  int stages;
  int kc;       // K elements per row segment per stage (multiple of 256); stage = 8 segments
};

// 8 bf16 x bf16 products onto `acc`, as two interleaved 4-long FMA chains (ILP: the consumer loop is
// latency-bound on dependent FMAs otherwise -- profiles/r1_gemv_v1.md)
__device__ __forceinline__ float dot8(const uint4& w, const uint4& x, float acc) {
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

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

#ifndef PK_GEMV_MIN_CTAS
#define PK_GEMV_MIN_CTAS 1  // occupancy hint of the streaming kernel (register cap = 64K / (288 * this))
#endif

template <int NTOK>
__global__ void __launch_bounds__(kConsumerThreads + 32, PK_GEMV_MIN_CTAS)
gemv_stream_kernel(const GemvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int stages = a.stages;
  const int kKC = a.kc;
  const int kSegBytes = kKC * 2, kStageBytes = kCW * kSegBytes;
  uint8_t* ring = smem;
  bf16* xs = reinterpret_cast<bf16*>(smem + (size_t)stages * kStageBytes);  // [NTOK][K]
  const size_t x_bytes = (((size_t)NTOK * a.K * 2) + 15) & ~(size_t)15;
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xs) + x_bytes);
  uint64_t* empty = full + kMaxStages;
  float* red = reinterpret_cast<float*>(empty + kMaxStages);  // 64 floats: reductions / SwiGLU swap
  float* ybuf = red + 64;                                      // epi 2: [NTOK][64] partial rows of this CTA

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int M = a.M, K = a.K;
  const bool glu = a.epi == 1 || a.epi == 4;  // interleaved gate / up row groups
  const int rows_per_group = glu ? kCW / 2 : kCW;
  const int r0 = (int)(((int64_t)blockIdx.x * M) / G);
  const int r1 = (int)(((int64_t)(blockIdx.x + 1) * M) / G);
  const int groups = (r1 - r0 + rows_per_group - 1) / rows_per_group;
  const int chunks = (K + kKC - 1) / kKC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kCW);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();  // let the next kernel start prefetching its weights

  if (warp == kCW) {
    // ===== producer warp: lanes 0..7 each stream one row segment per stage =====
    const uint64_t pol = l2_evict_first_policy();
    int s = 0;
    uint32_t ph = 0;
    for (int g = 0; g < groups; ++g) {
      // source row of consumer slot `lane` in this group (-1: none)
      int src = -1;
      if (lane < kCW) {
        const int out_row = r0 + g * rows_per_group + (glu ? (lane & 3) : lane);
        if (out_row < r1) src = glu ? (lane < 4 ? out_row : M + out_row) : out_row;
      }
      const unsigned valid = __ballot_sync(0xffffffffu, src >= 0);
      const int nvalid = __popc(valid);
      for (int c = 0; c < chunks; ++c) {
        const int k0 = c * kKC;
        const uint32_t bytes = (uint32_t)(min(kKC, K - k0) * 2);
        if (lane == 0) {
          mbar_wait(empty + s, ph ^ 1u);
          mbar_expect_tx(full + s, bytes * (uint32_t)nvalid);
        }
        __syncwarp();
        if (src >= 0)
          bulk_g2s(ring + (size_t)s * kStageBytes + (size_t)lane * kSegBytes,
                   a.W + (size_t)src * K + k0, bytes, full + s, pol);
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  pdl_wait();  // activations come from the previous kernel
  {
    const int tid = threadIdx.x;  // 0..255
    if (a.x_mode == 0) {
      const int nv = (NTOK * K) >> 3;
      for (int i = tid; i < nv; i += kConsumerThreads)
        reinterpret_cast<uint4*>(xs)[i] = reinterpret_cast<const uint4*>(a.X)[i];
    } else {
      // x = bf16((h + r) * rsqrt(mean((h+r)^2) + eps) * w); hidden_out = bf16(h + r).
      // One pass over global memory: each thread keeps its <= 5 vectors of the row in registers.
      // x_mode 2 (tensor parallel): r is the all-reduce of the peers' partial rows, summed here from the
      // local staging slots in rank order (fp32, one bf16 rounding == the collective's bf16 result) after
      // every rank's release flag for this sequence number has been observed.
      constexpr int kMaxVec = 5;  // K <= 5 * 256 * 8 = 10240
      const int nv = K >> 3;
      const uint8_t* stage_base = nullptr;
      if (a.x_mode == 2) {
        uint32_t* my_flags = a.tp.flags[a.tp.rank];
        const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(my_flags + 2 * kTpMaxCtas * kTpMaxWorld);
        const int slot = (int)(seq & 1u);
        if (tid < a.tp.world) {
          const uint32_t* f = my_flags + (size_t)(slot * kTpMaxCtas) * kTpMaxWorld + tid;
          for (uint32_t spins = 0; ld_acquire_sys(f) != seq; ++spins)
            if (spins > (1u << 27)) __trap();
        }
        consumer_bar();
        stage_base = a.tp.stage[a.tp.rank] + (size_t)slot * a.tp.world * a.tp.slot_bytes;
      }
      for (int n = 0; n < NTOK; ++n) {
        const uint4* h4 = reinterpret_cast<const uint4*>(a.X + (size_t)n * K);
        const uint4* r4 = reinterpret_cast<const uint4*>(a.residual + (size_t)n * K);
        float v[kMaxVec][8];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
          const int i = tid + j * kConsumerThreads;
          if (i < nv) {
            const uint4 h = h4[i];
            uint4 r;
            if (a.x_mode == 2) {
              float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              for (int rk = 0; rk < a.tp.world; ++rk) {
                const uint4 pv = __ldcg(reinterpret_cast<const uint4*>(stage_base + (size_t)rk * a.tp.slot_bytes +
                                                                       (size_t)n * K * 2) + i);
                acc8[0] += bf16_lo(pv.x); acc8[1] += bf16_hi(pv.x); acc8[2] += bf16_lo(pv.y); acc8[3] += bf16_hi(pv.y);
                acc8[4] += bf16_lo(pv.z); acc8[5] += bf16_hi(pv.z); acc8[6] += bf16_lo(pv.w); acc8[7] += bf16_hi(pv.w);
              }
              r.x = pack_bf16(acc8[0], acc8[1]); r.y = pack_bf16(acc8[2], acc8[3]);
              r.z = pack_bf16(acc8[4], acc8[5]); r.w = pack_bf16(acc8[6], acc8[7]);
            } else {
              r = r4[i];
            }
            v[j][0] = bf16_lo(h.x) + bf16_lo(r.x); v[j][1] = bf16_hi(h.x) + bf16_hi(r.x);
            v[j][2] = bf16_lo(h.y) + bf16_lo(r.y); v[j][3] = bf16_hi(h.y) + bf16_hi(r.y);
            v[j][4] = bf16_lo(h.z) + bf16_lo(r.z); v[j][5] = bf16_hi(h.z) + bf16_hi(r.z);
            v[j][6] = bf16_lo(h.w) + bf16_lo(r.w); v[j][7] = bf16_hi(h.w) + bf16_hi(r.w);
            if (a.x_mode == 3) {  // Qwen3.5: the residual sum is ROUNDED to bf16 before the norm (add_batch then norm)
#pragma unroll
              for (int e = 0; e < 8; ++e) v[j][e] = round_bf16(v[j][e]);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) ss = fmaf(v[j][e], v[j][e], ss);
          }
        }
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        consumer_bar();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kCW; ++w) tot += red[w];
        consumer_bar();
        const float rinv = rsqrtf(tot / (float)K + a.eps);
        const uint4* g4 = reinterpret_cast<const uint4*>(a.norm_w);
#pragma unroll
        for (int j = 0; j < kMaxVec; ++j) {
          const int i = tid + j * kConsumerThreads;
          if (i < nv) {
            const uint4 g = g4[i];
            const float wo = a.x_mode == 3 ? 1.0f : 0.0f;  // (1 + w) RMSNorm of Qwen3.5 (flashinfer_norm.cu:108-133)
            uint4 o;
            o.x = pack_bf16(v[j][0] * rinv * (wo + bf16_lo(g.x)), v[j][1] * rinv * (wo + bf16_hi(g.x)));
            o.y = pack_bf16(v[j][2] * rinv * (wo + bf16_lo(g.y)), v[j][3] * rinv * (wo + bf16_hi(g.y)));
            o.z = pack_bf16(v[j][4] * rinv * (wo + bf16_lo(g.z)), v[j][5] * rinv * (wo + bf16_hi(g.z)));
            o.w = pack_bf16(v[j][6] * rinv * (wo + bf16_lo(g.w)), v[j][7] * rinv * (wo + bf16_hi(g.w)));
            reinterpret_cast<uint4*>(xs + (size_t)n * K)[i] = o;
            if (blockIdx.x == 0) {
              uint4 hs;
              hs.x = pack_bf16(v[j][0], v[j][1]);
              hs.y = pack_bf16(v[j][2], v[j][3]);
              hs.z = pack_bf16(v[j][4], v[j][5]);
              hs.w = pack_bf16(v[j][6], v[j][7]);
              reinterpret_cast<uint4*>(a.hidden_out + (size_t)n * K)[i] = hs;
              if (a.normed_out) reinterpret_cast<uint4*>(a.normed_out + (size_t)n * K)[i] = o;
            }
          }
        }
      }
    }
    consumer_bar();
  }

  int s = 0;
  uint32_t ph = 0;
  for (int g = 0; g < groups; ++g) {
    const int out_row = r0 + g * rows_per_group + (glu ? (warp & 3) : warp);
    const bool has_row = out_row < r1;
    float acc4[NTOK][4];
#pragma unroll
    for (int n = 0; n < NTOK; ++n) acc4[n][0] = acc4[n][1] = acc4[n][2] = acc4[n][3] = 0.f;
    for (int c = 0; c < chunks; ++c) {
      mbar_wait(full + s, ph);
      if (has_row) {
        const int k0 = c * kKC;
        const int nvec = min(kKC, K - k0) >> 3;
        const uint4* wseg =
            reinterpret_cast<const uint4*>(ring + (size_t)s * kStageBytes + (size_t)warp * kSegBytes);
        int v0 = 0;
        for (; v0 + 128 <= nvec; v0 += 128) {
          // 2 KB of the segment: 4 vectors per lane, all shared-memory loads issued before the FMAs,
          // four independent accumulators per token
          uint4 wv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) wv[i] = wseg[v0 + lane + 32 * i];
#pragma unroll
          for (int n = 0; n < NTOK; ++n) {
            const uint4* xp = reinterpret_cast<const uint4*>(xs + (size_t)n * K + k0) + v0;
            uint4 xv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = xp[lane + 32 * i];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc4[n][i] = dot8(wv[i], xv[i], acc4[n][i]);
          }
        }
        for (int i = v0 + lane; i < nvec; i += 32) {
          const uint4 wv = wseg[i];
#pragma unroll
          for (int n = 0; n < NTOK; ++n) {
            const uint4 xv = reinterpret_cast<const uint4*>(xs + (size_t)n * K + k0)[i];
            acc4[n][0] = dot8(wv, xv, acc4[n][0]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
      if (++s == stages) {
        s = 0;
        ph ^= 1u;
      }
    }
    float acc[NTOK];
#pragma unroll
    for (int n = 0; n < NTOK; ++n) acc[n] = (acc4[n][0] + acc4[n][1]) + (acc4[n][2] + acc4[n][3]);
#pragma unroll
    for (int n = 0; n < NTOK; ++n) acc[n] = warp_sum(acc[n]);
    if (a.epi == 2 || a.epi == 3) {
      if (has_row && lane == 0) {
#pragma unroll
        for (int n = 0; n < NTOK; ++n) ybuf[n * 64 + (out_row - r0)] = acc[n];
      }
    } else if (a.epi == 0) {
      if (has_row && lane == 0) {
        const int seg = out_row < a.seg_end[0] ? 0 : (out_row < a.seg_end[1] ? 1 : 2);
        const int seg_lo = seg == 0 ? 0 : a.seg_end[seg - 1];
        const int seg_m = a.seg_end[seg] - seg_lo;
#pragma unroll
        for (int n = 0; n < NTOK; ++n)
          a.Y[seg][(size_t)n * seg_m + (out_row - seg_lo)] = f2bf(acc[n]);
      }
    } else {
      float* sw = red + 16 + (g & 1) * 16;  // [4 rows][NTOK] double-buffered by group parity
      if (warp >= 4 && lane == 0) {
#pragma unroll
        for (int n = 0; n < NTOK; ++n) sw[(warp - 4) * NTOK + n] = acc[n];
      }
      consumer_bar();
      if (warp < 4 && has_row && lane == 0) {
#pragma unroll
        for (int n = 0; n < NTOK; ++n) {
          const float gt = round_bf16(acc[n]);                   // gate_up_out is bf16 in the reference
          const float up = round_bf16(sw[warp * NTOK + n]);
          float sl = gt / (1.0f + expf(-gt));
          if (a.epi == 4) sl = round_bf16(sl);                   // Qwen3.5's unfused silu_mul rounds SiLU first (elementwise.cu:36-41)
          a.Y[0][(size_t)n * M + out_row] = f2bf(sl * up);
        }
      }
    }
  }
  if (a.epi == 3) {
    // ---- GEMV + all-reduce in one kernel, row tile by row tile.  Every rank runs the same grid, so CTA c owns rows
    // [r0, r1) on every rank: push my bf16 partials of these rows to the peers as 8-byte {2 x bf16, seq} lines (the
    // sequence number travels with the data: one posted NVLink store per line, no fence / flag / ticket), then poll
    // the peers' lines for the SAME rows in local memory and sum in rank order (fp32, one bf16 rounding: the
    // collective's bf16 result, bit-identical on every rank). ----
    consumer_bar();
    const int me = a.tp.rank, W = a.tp.world;
    const uint32_t seq = (*a.tp_step) * 256u + (uint32_t)a.tp_op + 1u;  // read after griddepcontrol.wait
    const int slot = a.tp_op & 1;
    const int nrows = r1 - r0;
    constexpr int NP = (NTOK + 1) / 2;  // token pairs per row
    const int items = nrows * NP;
    auto payload = [&](int it) -> uint32_t {
      const int r = it % nrows, pr = it / nrows;
      const float v0 = ybuf[(2 * pr) * 64 + r];
      const float v1 = (2 * pr + 1 < NTOK) ? ybuf[(2 * pr + 1) * 64 + r] : 0.f;
      return pack_bf16(v0, v1);
    };
    for (int idx = threadIdx.x; idx < items * (W - 1); idx += kConsumerThreads) {
      const int pi = idx / items, it = idx - pi * items;
      const int p = pi >= me ? pi + 1 : pi;
      const int r = it % nrows, pr = it / nrows;
      uint8_t* line = a.tp.stage[p] + (size_t)(slot * W + me) * a.tp.slot_bytes + a.tp.gll_off +
                      ((size_t)pr * M + r0 + r) * 8;
      st_volatile_v2(line, payload(it), seq);
    }
    const uint8_t* local = a.tp.stage[me] + (size_t)slot * W * a.tp.slot_bytes + a.tp.gll_off;
    for (int it = threadIdx.x; it < items; it += kConsumerThreads) {
      const int r = it % nrows, pr = it / nrows;
      float s0 = 0.f, s1 = 0.f;
      for (int q = 0; q < W; ++q) {
        uint32_t v;
        if (q == me) {
          v = payload(it);
        } else {
          const uint8_t* line = local + (size_t)q * a.tp.slot_bytes + ((size_t)pr * M + r0 + r) * 8;
          uint2 l;
          uint32_t spins = 0;
          do {
            l = ld_volatile_v2(line);
            if (++spins > (1u << 26)) __trap();  // a missing peer (~1 min) becomes a launch failure, not a hang
          } while (l.y != seq);
          v = l.x;
        }
        s0 += bf16_lo(v);
        s1 += bf16_hi(v);
      }
      a.Y[0][(size_t)(2 * pr) * M + r0 + r] = f2bf(s0);
      if (2 * pr + 1 < NTOK) a.Y[0][(size_t)(2 * pr + 1) * M + r0 + r] = f2bf(s1);
    }
  }
  if (a.epi == 2) {
    // ---- fused GEMV -> all-reduce (push half): store this CTA's partial rows straight into every rank's
    // staging slot [seq parity][my rank] over NVLink (consumer warp w serves peer w, 32 consecutive rows per
    // coalesced transaction), then the LAST CTA of the grid publishes the sequence flag to every rank with
    // release.sys.  The matching reduce + residual add + RMSNorm runs in the prologue (x_mode 2) of the next
    // GEMV on every rank: no standalone collective kernel.
    consumer_bar();
    const int me = a.tp.rank, W = a.tp.world;
    uint32_t* my_flags = a.tp.flags[me];
    uint32_t* ctl = my_flags + 2 * kTpMaxCtas * kTpMaxWorld;  // [0] = seq, [2] = CTA ticket of this op
    const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(ctl) + 1u;  // read after griddepcontrol.wait
    const int slot = (int)(seq & 1u);
    const int nrows = r1 - r0;
    for (int p = warp; p < W; p += kCW) {
      bf16* dst = reinterpret_cast<bf16*>(a.tp.stage[p] + (size_t)(slot * W + me) * a.tp.slot_bytes);
#pragma unroll
      for (int n = 0; n < NTOK; ++n)
        for (int r = lane; r < nrows; r += 32) dst[(size_t)n * M + r0 + r] = f2bf(ybuf[n * 64 + r]);
    }
    __threadfence_system();
    consumer_bar();
    if (threadIdx.x == 0) {
      __threadfence();
      const uint32_t t = atomicAdd(ctl + 2, 1u);
      if (t == gridDim.x - 1) {
        __threadfence_system();
        for (int p = 0; p < W; ++p)
          st_release_sys(a.tp.flags[p] + (size_t)(slot * kTpMaxCtas) * kTpMaxWorld + me, seq);
        ctl[2] = 0;
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(ctl) = seq;
      }
    }
  }
}

// Shapes the streaming kernel cannot take (K % 8 != 0, unaligned pointers, N > 4 without the
// tensor-core path): one warp per output element row, scalar loads.  Correctness only.
__global__ void gemv_generic_kernel(const bf16* __restrict__ W, const bf16* __restrict__ X,
                                    bf16* __restrict__ Y, int M, int N, int K) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t o = warp; o < (int64_t)M * N; o += nwarps) {
    const int m = (int)(o % M), n = (int)(o / M);
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(bf2f(W[(size_t)m * K + k]), bf2f(X[(size_t)n * K + k]), acc);
    acc = warp_sum(acc);
    if (lane == 0) Y[(size_t)n * M + m] = f2bf(acc);
  }
}

static int g_gemv_stages = 0, g_gemv_ctas_per_sm = 0, g_gemv_kc = 0;

static void gemv_tuning() {
  if (g_gemv_stages == 0) {
    const char* s = getenv("PK_GEMV_STAGES");
    g_gemv_stages = s ? atoi(s) : 6;  // 4 -> 6: -0.05 ms per Qwen3-4B decode step on B200 (profiles/README.md r2)
    if (g_gemv_stages < 2) g_gemv_stages = 2;
    if (g_gemv_stages > kMaxStages) g_gemv_stages = kMaxStages;
    const char* c = getenv("PK_GEMV_CTAS_PER_SM");
    g_gemv_ctas_per_sm = c ? atoi(c) : 2;
    if (g_gemv_ctas_per_sm < 1) g_gemv_ctas_per_sm = 1;
    const char* k = getenv("PK_GEMV_KC");
    g_gemv_kc = k ? atoi(k) : 1024;
    if (g_gemv_kc < 256 || g_gemv_kc % 256) g_gemv_kc = 1024;
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

void gemv_set_tuning(int stages, int ctas_per_sm, int kc) {
  gemv_tuning();
  if (stages >= 2 && stages <= kMaxStages) g_gemv_stages = stages;
  if (ctas_per_sm >= 1 && ctas_per_sm <= 4) g_gemv_ctas_per_sm = ctas_per_sm;
  if (kc >= 256 && kc % 256 == 0 && kc <= 16384) g_gemv_kc = kc;
}

bool gemv_stream_supported(const void* W, const void* X, int N, int K) {
  return N >= 1 && N <= 4 && K % 8 == 0 && al16(W) && al16(X) && (size_t)N * K * 2 <= 96 * 1024;
}

// CTA count of the streaming kernel: every CTA owns >= one row group, at most ctas_per_sm CTAs per SM.
int gemv_grid(int M, int epi) {
  gemv_tuning();
  const int rpg = (epi == 1 || epi == 4) ? kCW / 2 : kCW;
  int grid = sm_count() * g_gemv_ctas_per_sm;
  const int max_useful = (M + rpg - 1) / rpg;
  if (grid > max_useful) grid = max_useful;
  return grid < 1 ? 1 : grid;
}

template <int NTOK>
static cudaError_t launch_gemv_t(GemvArgs a, cudaStream_t stream) {
  gemv_tuning();
  int stages = g_gemv_stages;
  int kc = g_gemv_kc;
  while (kc > 256 && kc / 2 >= a.K) kc /= 2;  // no point in segments longer than a row
  const size_t x_bytes = (((size_t)NTOK * a.K * 2) + 15) & ~(size_t)15;
  const size_t tail = 2 * kMaxStages * sizeof(uint64_t) + (64 + 64 * 4) * sizeof(float);
  const size_t budget = (g_gemv_ctas_per_sm >= 2 ? 112 : 224) * 1024;
  while (stages > 2 && (size_t)stages * kCW * kc * 2 + x_bytes + tail > budget) --stages;
  while (kc > 256 && (size_t)stages * kCW * kc * 2 + x_bytes + tail > budget) kc -= 256;
  a.stages = stages;
  a.kc = kc;
  const size_t smem = (size_t)stages * kCW * kc * 2 + x_bytes + tail;
  auto kern = gemv_stream_kernel<NTOK>;
  static thread_local size_t configured[5] = {0, 0, 0, 0, 0};
  if (smem > configured[NTOK]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[NTOK] = smem;
  }
  const int grid = gemv_grid(a.M, a.epi);
  if ((a.epi == 2 || a.epi == 3) && (a.M + grid - 1) / grid > 64) return cudaErrorInvalidValue;
  return launch(kern, dim3(grid), dim3(kConsumerThreads + 32), smem, stream, true, a);
}

cudaError_t launch_gemv(const GemvArgs& a, int N, cudaStream_t stream) {
  switch (N) {
    case 1: return launch_gemv_t<1>(a, stream);
    case 2: return launch_gemv_t<2>(a, stream);
    case 3: return launch_gemv_t<3>(a, stream);
    case 4: return launch_gemv_t<4>(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

void launch_gemv_generic(const bf16* W, const bf16* X, bf16* Y, int M, int N, int K,
                         cudaStream_t stream) {
  int64_t warps = (int64_t)M * N;
  int64_t grid = (warps * 32 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  launch(gemv_generic_kernel, dim3((unsigned)grid), dim3(256), 0, stream, true, W, X, Y, M, N, K);
}

}  // namespace pk

namespace pk {
void launch_gemm(const bf16* W, const bf16* X, bf16* Y, int M, int N, int K, cudaStream_t stream);
}

extern "C" {

// ffi.rs:132-140.  N <= 4: HBM-streaming GEMV; larger N (graph buckets): tensor-core GEMM.
void gemm_graphsafe_cuda(const pk_bf16* W, const pk_bf16* X, pk_bf16* Y, int M, int N, int K,
                         pk_stream stream) {
  using namespace pk;
  if (M <= 0 || N <= 0 || K <= 0) return;
  if (N > 4) {
    launch_gemm((const bf16*)W, (const bf16*)X, (bf16*)Y, M, N, K, stream);
    return;
  }
  if (!gemv_stream_supported(W, X, N, K)) {
    launch_gemv_generic((const bf16*)W, (const bf16*)X, (bf16*)Y, M, N, K, stream);
    return;
  }
  GemvArgs a{};
  a.W = (const bf16*)W;
  a.X = (const bf16*)X;
  a.Y[0] = a.Y[1] = a.Y[2] = (bf16*)Y;
  a.seg_end[0] = a.seg_end[1] = a.seg_end[2] = M;
  a.M = M;
  a.K = K;
  launch_gemv(a, N, stream);
}

// ring depth (2..12 stages), CTAs per SM and K elements per row segment of the streaming GEMV; 0 keeps a value
int pk_b200_gemv_grid(int M, int epi) { return M > 0 ? pk::gemv_grid(M, epi) : 0; }

void pk_b200_set_gemv_tuning(int stages, int ctas_per_sm, int segment_elems) {
  pk::gemv_set_tuning(stages, ctas_per_sm, segment_elems);
}

int pk_b200_gemv_fused(const pk_b200_gemv_args* g, pk_stream stream) {
  using namespace pk;
  if (!g || g->M <= 0 || g->K <= 0) return -1;
  if (!gemv_stream_supported(g->W, g->X, g->N, g->K)) return -1;
  if (g->x_mode >= 1 && (g->hidden_out == nullptr || g->hidden_out == g->X || g->K > 10240)) return -1;
  if (g->x_mode < 0 || g->x_mode > 3 || g->epi < 0 || g->epi > 4) return -1;
  GemvArgs a{};
  a.W = (const bf16*)g->W;
  a.X = (const bf16*)g->X;
  int end = 0;
  for (int i = 0; i < 3; ++i) {
    a.Y[i] = (bf16*)(g->Y[i] ? g->Y[i] : g->Y[0]);
    end += g->seg_rows[i];
    a.seg_end[i] = end;
  }
  if (end == 0) a.seg_end[0] = a.seg_end[1] = g->M;
  a.seg_end[2] = g->M;
  a.M = g->M;
  a.K = g->K;
  a.x_mode = g->x_mode;
  a.residual = (const bf16*)g->residual;
  a.norm_w = (const bf16*)g->norm_w;
  a.eps = g->eps;
  a.hidden_out = (bf16*)g->hidden_out;
  a.normed_out = (bf16*)g->normed_out;
  a.epi = g->epi;
  if (g->x_mode == 2 || g->epi == 2) {
    const pk_tp_comm* comm = static_cast<const pk_tp_comm*>(g->tp_comm);
    if (!comm) return -1;
    if ((int64_t)g->N * (g->epi == 2 ? g->M : g->K) * 2 > comm->d.raw_bytes) return -2;
    a.tp = comm->d;
  }
  if (g->epi == 3) {
    const pk_tp_comm* comm = static_cast<const pk_tp_comm*>(g->tp_comm);
    if (!comm || !g->tp_step || g->tp_op < 0 || g->tp_op > 254) return -1;
    if ((int64_t)((g->N + 1) / 2) * g->M * 8 > comm->d.gll_bytes - 4096) return -2;  // the last 4 KB hold the top-1 lines
    a.tp = comm->d;
    a.tp_step = g->tp_step;
    a.tp_op = g->tp_op;
  }
  return (int)launch_gemv(a, g->N, stream);
}

}  // extern "C"
