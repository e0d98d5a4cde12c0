// Causal GQA prefill attention over the paged KV cache, second tcgen05 design: TWO 128-token query tiles per CTA in
// ping-pong, O accumulated in TMEM, P handed to the tensor core through TMEM (never through shared memory).
// Contract and rounding points are those of prefill_attention_tc.cu / prefill_attention.cu: S = Q K^T from bf16
// operands with fp32 accumulation, fp32 online softmax in the exp2 domain with the running maximum advancing per
// 128-token KV block, P rounded to bf16 for P V, denominator = row sum of the ROUNDED P (flashinfer prefill.cuh:956-985
// sums the converted fragment on the tensor cores), O / d rounded once to bf16.
//
// Why a second design (profiles/r1_v6_prefill_attn_tc_full_details.txt): the first kernel keeps ONE query tile per CTA
// and folds P V into register-resident O in the softmax warps -- ~790 instructions per softmax warp per KV block on
// 2 warps per scheduler, 68 % of cycles with no eligible warp, tensor pipe 15 % busy.  Here
//   * each CTA owns 256 query tokens of one head as two tiles; while tile A's softmax runs, the tensor core does tile
//     B's S or P V, so the MMA queue is never empty once the pipeline is primed;
//   * one softmax thread per query row (128 threads per tile): the row maximum needs no cross-thread exchange and S is
//     read from TMEM 1.5 times per block (second half kept in registers) instead of 2 times by twice the threads;
//   * O lives in TMEM and P V accumulates into it (enable_input_d); a separate correction warpgroup multiplies O by
//     alpha = exp2(m_old - m_new) between blocks, only for warps where some row's maximum moved (alpha == 1 is exact,
//     so skipping it changes nothing), off the softmax warps' critical path;
//   * P (bf16 pairs) is written with tcgen05.st over the first 64 columns of the tile's own S region and consumed as
//     the A operand from TMEM: no 32 KB shared-memory round trip, no generic->async proxy fence per block.
//
// Warps: 0-3 softmax tile 0 | 4-7 softmax tile 1 | 8-11 correction + epilogue | 12 MMA issuer | 13 TMA loader.
// TMEM (512 columns): S0 [0,128) | S1 [128,256) | O0 [256,384) | O1 [384,512); P_i aliases S_i[0,64).
// Shared memory: Q0, Q1 (32 KB each), K ring 2 x 32 KB, V ring 2 x 32 KB = 192 KB -> one CTA per SM.
// Per KV block j the in-order tensor queue is  P V(0,j) . S(0,j+1) . P V(1,j) . S(1,j+1): 4 x 512 cycles of MMA against
// two 128 x 128 softmax tiles (128 ex2 per thread, MUFU-bound at 2048 cycles per scheduler): the two are balanced by
// construction, the achievable rate is what the dependency chain S -> max -> exp -> P V of one tile leaves of it.
// Rows of the last page beyond kv_len are read as stored (their P is 0): the pool must hold finite values there.
#include "tcgen05.cuh"

namespace pk {
namespace fa2 {

constexpr int HD = 128;                // head dim
constexpr int TQ = 128;                // query tokens per tile (UMMA M)
constexpr int TKV = 128;               // kv tokens per block
constexpr int PAGE = 16;               // tokens per page
constexpr int HALF_BYTES = 128 * 128;  // one 64-column half of a [128 x 128] bf16 tile
constexpr int TILE_B = 2 * HALF_BYTES;
constexpr int W_CORR = 8, W_MMA = 12, W_LOAD = 14;  // warps 12 / 13 issue the MMAs of tile 0 / tile 1
constexpr int NTHREADS = 15 * 32;
constexpr uint32_t TMEM_COLS = 512;

struct Args {
  const bf16* q;
  bf16* out;
  const int* page_indices;
  const int* page_indptr;
  const int* last_page_len;
  const int* q_indptr;
  int seq_len, batch_size, nq, nkv;
  float sm_scale_log2;
  int dbg;  // PK_FA2_DBG ablation bits (timing experiments only; results are wrong when set): 1 no P V MMAs, 2 no S MMAs, 4 no exponentials, 8 no O rescale, 32 trace, 64 skip the first TMEM read of the max phase
};

// PK_FA2_DBG bit 32: the heaviest CTA of head 0 records %clock64 at the hand-over points of tile 0 (tools/fa2_trace.py)
__device__ unsigned long long g_fa2_trace[3 * 128 * 8];
#define FA2_TRACE(role, j, slot)                                                                                  \
  do {                                                                                                            \
    if (trace_on && lane == 0 && (j) < 128) g_fa2_trace[((role) * 128 + (j)) * 8 + (slot)] = clock64();           \
  } while (0)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// {x0, x1} * s + c on the packed fp32 pipe (FFMA2, scalars broadcast)
__device__ __forceinline__ void fma2(float& x0, float& x1, float s, float c) {
  unsigned long long v, sv, cv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(sv) : "f"(s));
  asm("mov.b64 %0, {%1, %1};" : "=l"(cv) : "f"(c));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(v), "l"(sv), "l"(cv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(v));
}
// acc0 += lo bf16 of pk, acc1 += hi bf16 of pk, exactly (fp32 accumulate of the ROUNDED values, FHADD.BF16: no unpack)
__device__ __forceinline__ void add_bf16_pair(float& acc0, float& acc1, uint32_t pk) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.bf16 %0, lo, %0;\n\tadd.rn.f32.bf16 %1, hi, %1;\n\t}"
      : "+f"(acc0), "+f"(acc1)
      : "r"(pk));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
// A hung barrier becomes a trap (reported as a launch failure) instead of a wedged GPU.
__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ uint32_t sw_off(int r, int c16) {
  return (uint32_t)((c16 >> 3) * HALF_BYTES + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
}
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (P) is read from tensor memory, lane = row, one 32-bit column = two
// consecutive K elements
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
// one wait for two 32-column loads in flight; in/out operands pin every use of the registers behind the wait
__device__ __forceinline__ void tmem_ld_wait_2x32(uint32_t* a, uint32_t* b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]),
                 "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(a[16]), "+r"(a[17]), "+r"(a[18]),
                 "+r"(a[19]), "+r"(a[20]), "+r"(a[21]), "+r"(a[22]), "+r"(a[23]), "+r"(a[24]), "+r"(a[25]), "+r"(a[26]), "+r"(a[27]),
                 "+r"(a[28]), "+r"(a[29]), "+r"(a[30]), "+r"(a[31]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]),
                 "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]),
                 "+r"(b[15]), "+r"(b[16]), "+r"(b[17]), "+r"(b[18]), "+r"(b[19]), "+r"(b[20]), "+r"(b[21]), "+r"(b[22]), "+r"(b[23]),
                 "+r"(b[24]), "+r"(b[25]), "+r"(b[26]), "+r"(b[27]), "+r"(b[28]), "+r"(b[29]), "+r"(b[30]), "+r"(b[31])
               :
               : "memory");
}
// one lane of the (converged) warp; the compiler knows a single thread is active under this predicate, so the
// uniform-register operands of tcgen05.mma / commit need no per-thread waterfall (unlike `lane == 0`)
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(NTHREADS, 1)
prefill_attention_tc2_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, const Args a) {
  extern __shared__ uint8_t fsm_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fsm_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem;                 // tile i at Qs + i * TILE_B
  uint8_t* KVs = smem + 2 * TILE_B;   // K stage s at KVs + s * 2 * TILE_B, V right after
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_B);
  uint64_t* k_full = bars;            // [2] TMA transaction barriers
  uint64_t* k_empty = bars + 2;       // [2] tcgen05.commit after the last S of the block
  uint64_t* v_full = bars + 4;        // [2]
  uint64_t* v_empty = bars + 6;       // [2] tcgen05.commit after the last P V of the block
  uint64_t* s_full = bars + 8;        // [tile] S(j) complete                        MMA -> softmax
  uint64_t* po_ready = bars + 10;     // [tile] P(j) stored AND O scaled by alpha(j) (4 + 4 warp arrivals)  softmax, correction -> MMA
  uint64_t* a_ready = bars + 12;      // [tile] alpha(j) published (128 arrivals)     softmax -> correction
  uint64_t* pv_done = bars + 16;      // [tile] P V(j) complete                       MMA -> correction
  uint64_t* d_ready = bars + 18;      // [tile] final denominators published (128)   softmax -> correction
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* alpha_s = reinterpret_cast<float*>(bars + 22);  // [tile][128]
  float* denom_s = alpha_s + 2 * TQ;                     // [tile][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x;  // heads vary fastest: every head's heaviest tile pair is scheduled before any lighter one
  const int kvh = head / (a.nq / a.nkv);

  // ---- locate (request, 256-token tile pair): late (heavy) pairs first ----
  int b = 0, q_start = 0, qo_len = 0, pair = -1;
  {
    int total = 0;
    for (int i = 0; i < a.batch_size; ++i) total += (a.q_indptr[i + 1] - a.q_indptr[i] + 2 * TQ - 1) / (2 * TQ);
    int idx = total - 1 - (int)blockIdx.y;
    if (idx < 0) return;
    for (int i = 0; i < a.batch_size; ++i) {
      const int len = a.q_indptr[i + 1] - a.q_indptr[i];
      const int nt = (len + 2 * TQ - 1) / (2 * TQ);
      if (idx < nt) {
        b = i; q_start = a.q_indptr[i]; qo_len = len; pair = idx;
        break;
      }
      idx -= nt;
    }
    if (pair < 0) return;
  }
  const int np = a.page_indptr[b + 1] - a.page_indptr[b];
  const int kv_len = np <= 0 ? 0 : (np - 1) * PAGE + a.last_page_len[b];
  const int* pages = a.page_indices + a.page_indptr[b];
  const int t0 = pair * 2 * TQ;
  const int causal_off = kv_len - qo_len;  // query token t attends kv <= t + causal_off
  // blocks each tile needs (tile 1 may not exist)
  const int nblk0 = (min(kv_len, causal_off + min(qo_len, t0 + TQ)) + TKV - 1) / TKV;
  const int nblk1 = t0 + TQ < qo_len ? (min(kv_len, causal_off + min(qo_len, t0 + 2 * TQ)) + TKV - 1) / TKV : 0;
  auto nblk = [&](int i) { return i ? nblk1 : nblk0; };
  const int nb = max(nblk0, nblk1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full + s, 1);
      mbar_init(k_empty + s, 2);  // one arrival per tile (a tile that does not read the block is arrived for)
      mbar_init(v_full + s, 1);
      mbar_init(v_empty + s, 2);
      mbar_init(s_full + s, 1);
      mbar_init(po_ready + s, 8);  // one elected arrival per warp: 4 softmax + 4 correction
      mbar_init(a_ready + s, 4);
      mbar_init(pv_done + s, 1);
      mbar_init(d_ready + s, 4);
    }
    mbar_fence_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
  pdl_launch_dependents();
  pdl_wait();  // q (and the appended K/V rows) come from the previous kernels

  tc_fence_before();
  __syncthreads();  // barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool trace_on = (a.dbg & 32) && blockIdx.x == 0 && blockIdx.y == 0;

  if (warp != W_LOAD) {
    // ---- both Q tiles -> swizzled smem (every warp but the loader, whose K / V loads are already in flight), rows past
    // the request are zero ----
    for (int idx = threadIdx.x; idx < 2 * TQ * 16; idx += NTHREADS - 32) {
      const int r = idx >> 4, c = idx & 15;  // r in [0, 256)
      const bool valid = t0 + r < qo_len;
      const bf16* src = a.q + ((size_t)(q_start + (valid ? t0 + r : 0)) * a.nq + head) * HD + c * 8;
      cp_async16(smem_u32(Qs) + (uint32_t)((r >> 7) * TILE_B) + sw_off(r & 127, c), src, valid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS - 32) : "memory");
  }

  if (warp == W_LOAD) {
    // =========================== TMA loader ===========================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
      for (int j = 0; j < nb; ++j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(((j >> 1) & 1) ^ 1);
        const int p0 = j * (TKV / PAGE);
        const int npg = min(TKV / PAGE, np - p0);
        const uint32_t kdst = smem_u32(KVs + (size_t)s * 2 * TILE_B), vdst = kdst + TILE_B;
        wait_or_trap(k_empty + s, ph);
        mbar_expect_tx(k_full + s, (uint32_t)npg * 2 * (64 * PAGE * 2));
        for (int p = 0; p < npg; ++p) {
          const int page = __ldg(pages + p0 + p);
          const uint32_t o = (uint32_t)(p * PAGE * 128);
          tma_load_4d(kdst + o, &map_k, 0, kvh, 0, page, k_full + s);
          tma_load_4d(kdst + HALF_BYTES + o, &map_k, 64, kvh, 0, page, k_full + s);
        }
        wait_or_trap(v_empty + s, ph);
        mbar_expect_tx(v_full + s, (uint32_t)npg * 2 * (64 * PAGE * 2));
        for (int p = 0; p < npg; ++p) {
          const int page = __ldg(pages + p0 + p);
          const uint32_t o = (uint32_t)(p * PAGE * 128);
          tma_load_4d(vdst + o, &map_v, 0, kvh, 0, page, v_full + s);
          tma_load_4d(vdst + HALF_BYTES + o, &map_v, 64, kvh, 0, page, v_full + s);
        }
      }
    }
  } else if (warp >= W_MMA) {
    // =========================== MMA issuers: one single-thread pipeline per tile ===========================
    // The trace of the one-issuer version (tools/fa2_trace.py) showed the issuing thread busy ~100 % of the time: every
    // tcgen05.mma stalls until the tensor queue has room, so issue time ~ execution time, and the waits / commits
    // between the batches (~320 cycles each) left the pipe dry.  Two issuers (disjoint TMEM regions) interleave in the
    // hardware queue, the K / V arrivals are checked before the P hand-over, and the descriptors are precomputed.
    {
      const int i = warp - W_MMA;
      const int n_i = nblk(i), n_o = nblk(1 - i);
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TKV >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      constexpr uint32_t idesc_pv = idesc_s | (1u << 16);  // B (= V) MN-major
      const uint64_t q_desc = make_sw128_desc(smem_u32(Qs) + (uint32_t)(i * TILE_B));
      const uint32_t t_sp = tmem_base + (uint32_t)(i * TKV);           // S, and P over its first 64 columns
      const uint32_t t_o = tmem_base + (uint32_t)(2 * TKV + i * HD);  // O
      auto issue_s = [&](int j) {
        const int s = j & 1;
        const uint64_t k_desc = make_sw128_desc(smem_u32(KVs + (size_t)s * 2 * TILE_B));
        if (elect_one()) {
          if (!(a.dbg & 2)) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
              const uint64_t o = (uint64_t)((k >> 2) * (HALF_BYTES >> 4) + (k & 3) * 2);  // descriptor address field: 16-byte units
              umma_bf16(t_sp, q_desc + o, k_desc + o, idesc_s, k > 0 ? 1u : 0u);
            }
          }
          umma_commit(s_full + i);
          umma_commit(k_empty + s);
          if (j >= n_o) umma_commit(k_empty + s);  // the other tile does not read K(j): arrive for it
        }
        __syncwarp();
        if (i == 0) FA2_TRACE(1, j, 5);
      };
      if (n_i > 0) {
        // Stagger the tiles by half a block: tile 1 starts when tile 0 has finished its first softmax, so that one
        // tile's exponentials run under the other's MMAs instead of both contending for the MUFU and then for the
        // tensor pipe (nothing re-synchronises them afterwards; measured with tools/fa2_trace.py).
        if (i == 1 && !(a.dbg & 128)) wait_or_trap(po_ready + 0, 0);
        wait_or_trap(k_full, 0);
        tc_fence_after();
        issue_s(0);
      }
      for (int j = 0; j < n_i; ++j) {
        const int s = j & 1;
        // operands first (long since there in steady state), then the hand-over that is on the critical path
        wait_or_trap(v_full + s, (uint32_t)((j >> 1) & 1));
        if (j + 1 < n_i) wait_or_trap(k_full + (s ^ 1), (uint32_t)(((j + 1) >> 1) & 1));
        if (i == 0) FA2_TRACE(1, j, 0);
        wait_or_trap(po_ready + i, (uint32_t)(j & 1));  // P(j) stored and O scaled by alpha(j)
        tc_fence_after();
        if (i == 0) FA2_TRACE(1, j, 1);
        const uint64_t v_desc = make_sw128_mn_desc(smem_u32(KVs + (size_t)s * 2 * TILE_B) + TILE_B, HALF_BYTES, 1024);
        const int ksteps = (a.dbg & 1) ? 0 : min(TKV / 16, np - j * (TKV / PAGE));  // one k-step = one 16-token page
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TKV / 16; ++k)
            if (k < ksteps) umma_bf16_ts(t_o, t_sp + (uint32_t)(k * 8), v_desc + (uint64_t)(k * (2048 >> 4)), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(pv_done + i);
          umma_commit(v_empty + s);
          if (j >= n_o) umma_commit(v_empty + s);
        }
        __syncwarp();
        if (i == 0) FA2_TRACE(1, j, 3);
        if (j + 1 < n_i) issue_s(j + 1);  // in order behind P V(j): it overwrites the S / P columns P V(j) reads
      }
    }
  } else if (warp >= W_CORR) {
    // =========================== correction + epilogue ===========================
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
    for (int j = 0; j < nb; ++j)
      for (int i = 0; i < 2; ++i) {
        if (j >= nblk(i)) continue;
        if (j > 0) {
          wait_or_trap(a_ready + i, (uint32_t)((j - 1) & 1));
          if (i == 0 && q4 == 0) FA2_TRACE(2, j, 0);
          const float alpha = alpha_s[i * TQ + r];
          const bool need = !__all_sync(0xffffffffu, alpha == 1.0f);
          wait_or_trap(pv_done + i, (uint32_t)((j - 1) & 1));  // O holds blocks < j
          tc_fence_after();
          if (i == 0 && q4 == 0) FA2_TRACE(2, j, 1);
          if (need && !(a.dbg & 8)) {
            const uint32_t t_o = lane_base + (uint32_t)(2 * TKV + i * HD);
#pragma unroll
            for (int c = 0; c < HD; c += 64) {
              uint32_t t[64];
              tmem_ld32_nowait(t_o + c, t);
              tmem_ld32_nowait(t_o + c + 32, t + 32);
              tmem_ld_wait64(t);
#pragma unroll
              for (int x = 0; x < 64; ++x) t[x] = __float_as_uint(__uint_as_float(t[x]) * alpha);
              tmem_st32(t_o + c, t);
              tmem_st32(t_o + c + 32, t + 32);
            }
            tmem_st_wait();
          }
          tc_fence_before();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(po_ready + i);  // block 0: nothing to scale, the MMA warp still expects both halves of the hand-over
        if (i == 0 && q4 == 0) FA2_TRACE(2, j, 2);
      }
    // epilogue: O / d -> bf16 -> global, one row per thread
    for (int i = 0; i < 2; ++i) {
      if (nblk(i) == 0) continue;
      wait_or_trap(d_ready + i, 0);
      wait_or_trap(pv_done + i, (uint32_t)((nblk(i) - 1) & 1));
      tc_fence_after();
      const int tok = t0 + i * TQ + r;
      const float inv = __fdividef(1.f, denom_s[i * TQ + r]);
      const uint32_t t_o = lane_base + (uint32_t)(2 * TKV + i * HD);
      uint4* dst = reinterpret_cast<uint4*>(a.out + ((size_t)(q_start + (tok < qo_len ? tok : 0)) * a.nq + head) * HD);
#pragma unroll
      for (int c = 0; c < HD; c += 64) {
        uint32_t t[64];
        tmem_ld32_nowait(t_o + c, t);
        tmem_ld32_nowait(t_o + c + 32, t + 32);
        tmem_ld_wait64(t);
        if (tok < qo_len) {
#pragma unroll
          for (int x = 0; x < 64; x += 8)
            dst[(c + x) >> 3] = make_uint4(pack_bf16(__uint_as_float(t[x]) * inv, __uint_as_float(t[x + 1]) * inv),
                                           pack_bf16(__uint_as_float(t[x + 2]) * inv, __uint_as_float(t[x + 3]) * inv),
                                           pack_bf16(__uint_as_float(t[x + 4]) * inv, __uint_as_float(t[x + 5]) * inv),
                                           pack_bf16(__uint_as_float(t[x + 6]) * inv, __uint_as_float(t[x + 7]) * inv));
        }
      }
    }
  } else {
    // =========================== softmax: one thread per query row ===========================
    const int i = warp >> 2;  // tile
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const int tok = t0 + i * TQ + r;
    const int lim = tok < qo_len ? min(kv_len - 1, tok + causal_off) : -1;  // last kv index this row may see
    const uint32_t t_s = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(i * TKV);
    const float scale = a.sm_scale_log2;
    float m = -INFINITY, d = 0.f;
    const int n_i = nblk(i);
    for (int j = 0; j < n_i; ++j) {
      wait_or_trap(s_full + i, (uint32_t)(j & 1));
      tc_fence_after();
      if (warp == 0) FA2_TRACE(0, j, 0);
      const bool masked = !__all_sync(0xffffffffu, j * TKV + TKV - 1 <= lim);  // warp-uniform
      // ---- row maximum in raw score units (the scale is positive): columns 64..127 first (dropped), then 0..63 (kept) ----
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      uint32_t v[64];
      if (!(a.dbg & 64)) {
        tmem_ld32_nowait(t_s + 64, v);
        tmem_ld32_nowait(t_s + 96, v + 32);
        tmem_ld_wait64(v);
      }
      if (masked) {
        const int c0 = j * TKV + 64;
#pragma unroll
        for (int x = 0; x < 64; ++x) mx4[x & 3] = fmaxf(mx4[x & 3], c0 + x <= lim ? __uint_as_float(v[x]) : -INFINITY);
      } else {
#pragma unroll
        for (int x = 0; x < 64; ++x) mx4[x & 3] = fmaxf(mx4[x & 3], __uint_as_float(v[x]));
      }
      tmem_ld32_nowait(t_s, v);
      tmem_ld32_nowait(t_s + 32, v + 32);
      tmem_ld_wait64(v);
      if (masked) {
        const int c0 = j * TKV;
#pragma unroll
        for (int x = 0; x < 64; ++x) {
          if (c0 + x > lim) v[x] = 0xff800000u;  // -inf -> P = 0
          mx4[x & 3] = fmaxf(mx4[x & 3], __uint_as_float(v[x]));
        }
      } else {
#pragma unroll
        for (int x = 0; x < 64; ++x) mx4[x & 3] = fmaxf(mx4[x & 3], __uint_as_float(v[x]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * scale;
      const float m_new = fmaxf(m, mx);
      const float ref = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = ex2(m - ref);  // m = -inf -> 0
      m = m_new;
      if (warp == 0) FA2_TRACE(0, j, 1);
      if (j > 0) {
        alpha_s[i * TQ + r] = alpha;
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready + i);  // release: the correction warps read alpha after their wait
      }
      if (a.dbg & 4) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(po_ready + i);
        continue;
      }
      // ---- P = bf16(exp2(s * scale - ref)); the denominator sums the ROUNDED values (FHADD.BF16 straight off the pair) ----
      const float nref = -ref;
      float ds4[4] = {0.f, 0.f, 0.f, 0.f};
      auto p_pair = [&](uint32_t u0, uint32_t u1, int x) -> uint32_t {
        float s0 = __uint_as_float(u0), s1 = __uint_as_float(u1);
        fma2(s0, s1, scale, nref);
        const uint32_t pk = pack_bf16(ex2(s0), ex2(s1));
        add_bf16_pair(ds4[(x & 1) * 2], ds4[(x & 1) * 2 + 1], pk);
        return pk;
      };
      uint32_t w[32];
      tmem_ld32_nowait(t_s + 64, w);  // columns 64..95 again (still intact): in flight under the first half's exponentials
#pragma unroll
      for (int x = 0; x < 32; ++x) v[x] = p_pair(v[2 * x], v[2 * x + 1], x);
      tmem_st32(t_s, v);  // P columns 0..31 = kv 0..63 of the block (over S columns already consumed)
      tmem_ld32_nowait(t_s + 96, v + 32);
      tmem_ld_wait_2x32(w, v + 32);
      if (masked) {
        const int c0 = j * TKV + 64;
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          if (c0 + x > lim) w[x] = 0xff800000u;
          if (c0 + 32 + x > lim) v[32 + x] = 0xff800000u;
        }
      }
#pragma unroll
      for (int x = 0; x < 16; ++x) w[x] = p_pair(w[2 * x], w[2 * x + 1], x);
#pragma unroll
      for (int x = 0; x < 16; ++x) w[16 + x] = p_pair(v[32 + 2 * x], v[32 + 2 * x + 1], x);
      tmem_st32(t_s + 32, w);  // P columns 32..63 = kv 64..127
      if (warp == 0) FA2_TRACE(0, j, 2);
      tmem_st_wait();
      d = fmaf(d, alpha, (ds4[0] + ds4[1]) + (ds4[2] + ds4[3]));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(po_ready + i);
      if (warp == 0) FA2_TRACE(0, j, 3);
    }
    if (n_i > 0) {
      denom_s[i * TQ + r] = d;
      __syncwarp();
      if (lane == 0) mbar_arrive(d_ready + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// 4-D view of one layer's K (or V) block of the page-first pool: {head dim 128, kv head, slot 16, page};
// box = {64, 1, 16, 1} = one 64-column half of one page of one kv head, 128-byte swizzle.
static bool make_kv_map(CUtensorMap* map, const bf16* base, int nkv, int64_t stride_page) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {(cuuint64_t)HD, (cuuint64_t)nkv, (cuuint64_t)PAGE, (cuuint64_t)1 << 20};
  cuuint64_t strides[3] = {(cuuint64_t)HD * 2, (cuuint64_t)nkv * HD * 2, (cuuint64_t)stride_page * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)PAGE, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace fa2

// copies the PK_FA2_DBG=32 trace ([role 0 softmax / 1 MMA / 2 correction][block][8] clock64 values) to the host
extern "C" int pk_b200_fa2_trace_copy(void* host_out, int bytes) {
  if (bytes > (int)sizeof(fa2::g_fa2_trace)) bytes = (int)sizeof(fa2::g_fa2_trace);
  return (int)cudaMemcpyFromSymbol(host_out, fa2::g_fa2_trace, bytes);
}

// Launch for the paged batch-prefill entry (prefill_attention.cu dispatches here).  Returns cudaError as int, -2 when
// the pool cannot be described by a TMA tensor map (the caller falls back).
int launch_prefill_tc2(const bf16* q, bf16* out, const bf16* k_base, const bf16* v_base, const int* page_indices,
                       const int* page_indptr, const int* last_page_len, const int* q_indptr, int seq_len, int batch_size, int nq,
                       int nkv, int page_size, int64_t stride_page, float sm_scale_log2, cudaStream_t stream) {
  using namespace fa2;
  if (page_size != PAGE || (stride_page * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(k_base) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(v_base) & 15) != 0 || (reinterpret_cast<uintptr_t>(q) & 15) != 0)
    return -2;
  CUtensorMap mk, mv;
  if (!make_kv_map(&mk, k_base, nkv, stride_page) || !make_kv_map(&mv, v_base, nkv, stride_page)) return -2;
  Args a{};
  a.q = q; a.out = out;
  a.page_indices = page_indices; a.page_indptr = page_indptr; a.last_page_len = last_page_len; a.q_indptr = q_indptr;
  a.seq_len = seq_len; a.batch_size = batch_size; a.nq = nq; a.nkv = nkv;
  a.sm_scale_log2 = sm_scale_log2;
  {
    const char* e = getenv("PK_FA2_DBG");
    a.dbg = e ? atoi(e) : 0;
  }
  constexpr size_t smem = 6 * TILE_B + 1024 + 256 + 2048;
  static thread_local bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(prefill_attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cfg = true;
  }
  const int pairs = (seq_len + 2 * TQ - 1) / (2 * TQ) + batch_size;  // upper bound; surplus CTAs exit at once
  return (int)launch(prefill_attention_tc2_kernel, dim3(nq, pairs), dim3(NTHREADS), smem, stream, true, mk, mv, a);
}

}  // namespace pk
