// Causal GQA prefill attention over the paged KV cache on the 5th-generation tensor cores (tcgen05 + TMEM).
// Same contract and rounding points as prefill_attention.cu (the mma.sync version): S = Q K^T from bf16 operands with
// fp32 accumulation, fp32 online softmax in the exp2 domain, P rounded to bf16 for the P V product, denominator =
// row sum of the ROUNDED P, O accumulated in fp32, O/d rounded once to bf16 (flashinfer prefill.cuh:956-985).
// The running maximum advances per 128-token KV block (the mma.sync kernel: per 64) -- same mathematics, P differs
// by its bf16 rounding relative to a different reference, covered by the test tolerance.
//
// One CTA = 128 query tokens of ONE q head (TMEM lane = token).  Per 128-token KV block j:
//   loader warp (9)    : one lane issues 4-D TMA tile loads straight out of the paged pool -- per 16-token page two
//                        {64 x 16} boxes for K and two for V, 128-byte swizzle, mbarrier complete_tx; separate
//                        2-stage rings for K (freed by S) and V (freed by P V)
//   MMA warp (8)       : S[b] = Q K^T (M128 x N128 x K16 x 8, both operands K-major) into TMEM buffer b = j & 1,
//                        issued one block AHEAD of the softmax; then P V (A = P K-major, B = V MN-major: the V tile
//                        is [kv token][head dim], i.e. N-contiguous) into a third TMEM region; only the k-steps whose
//                        16-token page was loaded are issued
//   softmax warps (0-7): two threads per query row (warps w and w+4 share TMEM lane quadrant w): each owns 64 of the
//                        128 score columns and 64 of the 128 output columns.  Each reads the whole S row from TMEM
//                        for the row maximum (the partner's half is reduced and dropped), so the loop has no
//                        cross-warp exchange; the denominators are combined once at the end.  bf16 P -> swizzled
//                        smem; the previous
//                        block's O = O * alpha + (P V) fold (O in 64 registers per thread) is deferred until just
//                        before P is rewritten, so P V runs under the next block's exponentials.  Blocks fully
//                        below the diagonal skip the per-element mask.
// TMEM: S0 [0,128) | S1 [128,256) | PV [256,384) of a 512-column allocation.  Shared memory: Q 32 KB, P 32 KB,
// 2 x (K 32 KB + V 32 KB) = 192 KB -> one CTA per SM.
// Rows of the last page beyond kv_len are read as stored (their P is 0): the pool must hold finite values there
// (it is zero-initialised by the host, as the reference's KvPool).
// Tensor-bound: 4 * 128 * kv flop per (query token, head); causal blocks past the diagonal are skipped.
#include "tcgen05.cuh"

namespace pk {

constexpr int THD = 128;                // head dim
constexpr int TQ = 128;                 // query tokens per CTA (UMMA M)
constexpr int TKV = 128;                // kv tokens per block (UMMA N of S, K extent of P V)
constexpr int TPAGE = 16;               // tokens per page (the only page size of the reference)
constexpr int HALF_BYTES = 128 * 128;   // one 64-column half of a [128 x 128] bf16 tile: 128 rows x 128 B
constexpr int TILE_B = 2 * HALF_BYTES;  // 32 KB
constexpr int T_SM_WARPS = 8;           // softmax warps
constexpr int T_SM_THREADS = T_SM_WARPS * 32;
constexpr int T_MMA_WARP = 8, T_LOAD_WARP = 9;
constexpr int T_THREADS = 10 * 32;
constexpr uint32_t T_TMEM_COLS = 512;

struct PrefillTcArgs {
  const bf16* q;
  bf16* out;
  const int* page_indices;
  const int* page_indptr;
  const int* last_page_len;
  const int* q_indptr;
  int seq_len, batch_size, nq, nkv;
  float sm_scale_log2;
};

__device__ __forceinline__ float tex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tcp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
// A hung barrier becomes a trap (reported as a launch failure) instead of a wedged GPU.
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 27)) __trap();
}
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// byte offset of 16-byte chunk `c16` (0..15) of row `r` in a [128 rows x 128 cols] bf16 tile stored as two
// 64-column halves, each half = 128 rows x 128 B with the 128-byte swizzle (chunk ^= row & 7)
__device__ __forceinline__ uint32_t sw_off(int r, int c16) {
  return (uint32_t)((c16 >> 3) * HALF_BYTES + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
}
__global__ void __launch_bounds__(T_THREADS, 1)
prefill_attention_tc_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                            const PrefillTcArgs a) {
  extern __shared__ uint8_t tsm_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tsm_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem;
  uint8_t* Ps = smem + TILE_B;
  uint8_t* KVs = smem + 2 * TILE_B;  // stage s: K at KVs + s*2*TILE_B, V right after
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_B);
  // K and V have separate 2-stage rings: K(j) is free as soon as S(j) retires (one block ahead of the softmax), V(j)
  // only after P V(j) -- with a joint stage the next K load could not start before the previous P V finished and
  // every S waited for a TMA round trip.
  uint64_t* k_full = bars;        // [2], TMA transaction barrier
  uint64_t* k_empty = bars + 2;   // [2], count 1 (tcgen05.commit after S)
  uint64_t* v_full = bars + 4;    // [2], TMA transaction barrier
  uint64_t* v_empty = bars + 6;   // [2], count 1 (tcgen05.commit after P V)
  uint64_t* s_full = bars + 8;    // [2], count 1
  uint64_t* p_full = bars + 10;   // count 256
  uint64_t* pv_full = bars + 11;  // count 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* xch = reinterpret_cast<float*>(bars + 14);  // [2 parities][2 halves][128 rows] row maxima / denominators

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x;  // heads vary fastest: all heads' heaviest tiles are scheduled first
  const int kvh = head / (a.nq / a.nkv);

  // ---- locate (request, 128-token tile): late (heavy) tiles first ----
  int b = 0, q_start = 0, qo_len = a.seq_len, tile_local = -1;
  {
    int total_tiles = 0;
    for (int i = 0; i < a.batch_size; ++i) total_tiles += (a.q_indptr[i + 1] - a.q_indptr[i] + TQ - 1) / TQ;
    int idx = total_tiles - 1 - (int)blockIdx.y;
    if (idx < 0) return;
    for (int i = 0; i < a.batch_size; ++i) {
      const int len = a.q_indptr[i + 1] - a.q_indptr[i];
      const int nt = (len + TQ - 1) / TQ;
      if (idx < nt) {
        b = i; q_start = a.q_indptr[i]; qo_len = len; tile_local = idx;
        break;
      }
      idx -= nt;
    }
    if (tile_local < 0) return;
  }
  const int np = a.page_indptr[b + 1] - a.page_indptr[b];
  const int kv_len = np <= 0 ? 0 : (np - 1) * TPAGE + a.last_page_len[b];
  const int* pages = a.page_indices + a.page_indptr[b];
  const int t0 = tile_local * TQ;
  const int causal_off = kv_len - qo_len;  // query token t attends kv <= t + causal_off
  const int kv_end = min(kv_len, causal_off + min(qo_len, t0 + TQ));
  const int n_blocks = (kv_end + TKV - 1) / TKV;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full + s, 1);
      mbar_init(k_empty + s, 1);
      mbar_init(v_full + s, 1);
      mbar_init(v_empty + s, 1);
      mbar_init(s_full + s, 1);
    }
    mbar_init(p_full, T_SM_THREADS);
    mbar_init(pv_full, 1);
    mbar_fence_init();
  }
  if (warp == T_MMA_WARP) tmem_alloc(tmem_slot, T_TMEM_COLS);
  pdl_launch_dependents();
  pdl_wait();  // q (and the appended K/V rows) come from the previous kernels

  // ---- Q tile -> swizzled smem (all threads), rows past the request are zero ----
  for (int idx = threadIdx.x; idx < TQ * 16; idx += T_THREADS) {
    const int r = idx >> 4, c = idx & 15;
    const bool valid = t0 + r < qo_len;
    const bf16* src = a.q + ((size_t)(q_start + (valid ? t0 + r : 0)) * a.nq + head) * THD + c * 8;
    tcp_async16(smem_u32(Qs) + sw_off(r, c), src, valid);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == T_LOAD_WARP) {
    // =========================== TMA loader ===========================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
      for (int j = 0; j < n_blocks; ++j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(((j >> 1) & 1) ^ 1);
        const int p0 = j * (TKV / TPAGE);
        const int npg = min(TKV / TPAGE, np - p0);  // pages of this block that exist
        const uint32_t kdst = smem_u32(KVs + (size_t)s * 2 * TILE_B), vdst = kdst + TILE_B;
        mbar_wait_or_trap(k_empty + s, ph);
        mbar_expect_tx(k_full + s, (uint32_t)npg * 2 * (64 * TPAGE * 2));
        for (int p = 0; p < npg; ++p) {
          const int page = __ldg(pages + p0 + p);
          const uint32_t o = (uint32_t)(p * TPAGE * 128);  // 16 rows x 128 B inside each half
          tma_load_4d(kdst + o, &map_k, 0, kvh, 0, page, k_full + s);
          tma_load_4d(kdst + HALF_BYTES + o, &map_k, 64, kvh, 0, page, k_full + s);
        }
        mbar_wait_or_trap(v_empty + s, ph);
        mbar_expect_tx(v_full + s, (uint32_t)npg * 2 * (64 * TPAGE * 2));
        for (int p = 0; p < npg; ++p) {
          const int page = __ldg(pages + p0 + p);
          const uint32_t o = (uint32_t)(p * TPAGE * 128);
          tma_load_4d(vdst + o, &map_v, 0, kvh, 0, page, v_full + s);
          tma_load_4d(vdst + HALF_BYTES + o, &map_v, 64, kvh, 0, page, v_full + s);
        }
      }
    }
  } else if (warp == T_MMA_WARP) {
    // =========================== MMA issuer ===========================
    // D = f32, A = B = bf16, M = 128, N = 128; P V additionally reads B (= V) MN-major
    constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TKV >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
    constexpr uint32_t idesc_pv = idesc_s | (1u << 16);
    const uint32_t q_addr = smem_u32(Qs), p_addr = smem_u32(Ps);
    auto issue_s = [&](int j) {
      const int s = j & 1;
      mbar_wait_or_trap(k_full + s, (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      if (lane == 0) {
        const uint32_t k_addr = smem_u32(KVs + (size_t)s * 2 * TILE_B);
        const uint32_t d = tmem_base + (uint32_t)(s * TKV);
#pragma unroll
        for (int k = 0; k < THD / 16; ++k) {
          const uint32_t o = (uint32_t)((k >> 2) * HALF_BYTES + (k & 3) * 32);
          umma_bf16(d, make_sw128_desc(q_addr + o), make_sw128_desc(k_addr + o), idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(s_full + s);
        umma_commit(k_empty + s);  // the K stage can be refilled as soon as S(j) has retired
      }
      __syncwarp();
    };
    if (n_blocks > 0) issue_s(0);
    for (int j = 0; j < n_blocks; ++j) {
      if (j + 1 < n_blocks) issue_s(j + 1);  // runs under the softmax of block j
      mbar_wait_or_trap(p_full, (uint32_t)(j & 1));
      mbar_wait_or_trap(v_full + (j & 1), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      if (lane == 0) {
        const int s = j & 1;
        const uint32_t v_addr = smem_u32(KVs + (size_t)s * 2 * TILE_B) + TILE_B;
        const uint32_t d = tmem_base + 2 * TKV;
        const int ksteps = min(TKV / 16, np - j * (TKV / TPAGE));  // one k-step = one 16-token page
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t ao = (uint32_t)((k >> 2) * HALF_BYTES + (k & 3) * 32);  // P: K-major, K = kv token
          const uint32_t bo = (uint32_t)(k * 2048);                              // V: 16 tokens = two 8-row groups
          umma_bf16(d, make_sw128_desc(p_addr + ao), make_sw128_mn_desc(v_addr + bo, HALF_BYTES, 1024), idesc_pv,
                    k > 0 ? 1u : 0u);
        }
        umma_commit(pv_full);
        umma_commit(v_empty + s);
      }
      __syncwarp();
    }
  } else {
    // =========================== softmax warps: two threads per query row ===========================
    const int r = (warp & 3) * 32 + lane;  // query row = TMEM lane
    const int hh = warp >> 2;              // which 64 columns (of S, and of O) this thread owns
    const int tok = t0 + r;
    const bool row_ok = tok < qo_len;
    const int lim = row_ok ? min(kv_len - 1, tok + causal_off) : -1;  // last kv index this row may see
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(hh * 64);
    const uint32_t t_lane_other = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((hh ^ 1) * 64);
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m = -INFINITY, d = 0.f, alpha_prev = 0.f;
    const uint32_t t_pv = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(2 * TKV + hh * 64);
    // O = O * alpha + P V of block jj (this thread's 64 output columns), 32 columns at a time
    auto fold_pv = [&](int jj, float alpha) {
      mbar_wait_or_trap(pv_full, (uint32_t)(jj & 1));
      tc_fence_after();
      {
        uint32_t t[64];
        tmem_ld32_nowait(t_pv, t);
        tmem_ld32_nowait(t_pv + 32, t + 32);
        tmem_ld_wait64(t);
#pragma unroll
        for (int i = 0; i < 64; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(t[i]));
      }
      tc_fence_before();  // the PV region is rewritten only after our next p_full arrival
    };
    for (int j = 0; j < n_blocks; ++j) {
      const int sb = j & 1;
      mbar_wait_or_trap(s_full + sb, (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      // Row maximum over all 128 columns, in raw score units (the scale is positive, so max and scale commute
      // exactly): the partner thread's 64 columns are reduced and dropped, ours are kept.  Both threads of a row
      // compute the same maximum from the same data -- no exchange, no barrier inside the loop.
      const bool masked = !__all_sync(0xffffffffu, j * TKV + TKV - 1 <= lim);  // warp-uniform
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      {
        uint32_t w[64];
        tmem_ld32_nowait(t_lane_other + (uint32_t)(sb * TKV), w);
        tmem_ld32_nowait(t_lane_other + (uint32_t)(sb * TKV + 32), w + 32);
        tmem_ld_wait64(w);
        if (masked) {
          const int c_other = j * TKV + (hh ^ 1) * 64;
#pragma unroll
          for (int i = 0; i < 64; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], c_other + i <= lim ? __uint_as_float(w[i]) : -INFINITY);
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(w[i]));
        }
      }
      uint32_t v[64];
      tmem_ld32_nowait(t_lane + (uint32_t)(sb * TKV), v);
      tmem_ld32_nowait(t_lane + (uint32_t)(sb * TKV + 32), v + 32);
      tmem_ld_wait64(v);
      if (masked) {
        const int col0 = j * TKV + hh * 64;  // kv index of v[0]
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          if (col0 + i > lim) v[i] = 0xff800000u;  // -inf -> P = 0
          mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(v[i]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * a.sm_scale_log2;
      const float m_new = fmaxf(m, mx);
      const float ref = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = tex2(m - ref);  // m = -inf -> 0
      m = m_new;
      // P = bf16(exp2(s - ref)), packed in place (v[0..31]); masked scores are -inf -> P = 0.  The rounded values
      // (the denominator sums ROUNDED P) are recovered from the packed pair with two integer ops, so the only
      // special-function work per element is the exp2 itself.
      float ds4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const uint32_t pk = pack_bf16(tex2(fmaf(__uint_as_float(v[2 * i]), a.sm_scale_log2, -ref)),
                                      tex2(fmaf(__uint_as_float(v[2 * i + 1]), a.sm_scale_log2, -ref)));
        ds4[i & 3] += __uint_as_float(pk << 16) + __uint_as_float(pk & 0xffff0000u);
        v[i] = pk;
      }
      tc_fence_before();  // our tcgen05.ld of S[sb] are done before the MMA warp may overwrite it (block j+2)
      // Block j-1's P V ran under everything above; fold it into O now -- its completion also frees the P buffer.
      if (j > 0) fold_pv(j - 1, alpha_prev);
#pragma unroll
      for (int q8 = 0; q8 < 8; ++q8)
        *reinterpret_cast<uint4*>(Ps + sw_off(r, hh * 8 + q8)) = make_uint4(v[q8 * 4], v[q8 * 4 + 1], v[q8 * 4 + 2], v[q8 * 4 + 3]);
      d = fmaf(d, alpha, (ds4[0] + ds4[1]) + (ds4[2] + ds4[3]));
      fence_proxy_async_smem();  // P visible to the tensor core's operand reads
      mbar_arrive(p_full);
      alpha_prev = alpha;
    }
    if (n_blocks > 0) fold_pv(n_blocks - 1, alpha_prev);
    // denominators of the two column halves
    float* xb = xch;
    xb[hh * 128 + r] = d;
    softmax_bar();
    d += xb[(hh ^ 1) * 128 + r];
    if (row_ok) {
      const float inv = __fdividef(1.f, d);
      uint4* dst = reinterpret_cast<uint4*>(a.out + ((size_t)(q_start + tok) * a.nq + head) * THD + hh * 64);
#pragma unroll
      for (int c = 0; c < 64; c += 8)
        dst[c >> 3] = make_uint4(pack_bf16(o[c] * inv, o[c + 1] * inv), pack_bf16(o[c + 2] * inv, o[c + 3] * inv),
                                 pack_bf16(o[c + 4] * inv, o[c + 5] * inv), pack_bf16(o[c + 6] * inv, o[c + 7] * inv));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == T_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, T_TMEM_COLS);
  }
}

// 4-D view of one layer's K (or V) block of the page-first pool: {head dim 128, kv head, slot 16, page};
// box = {64, 1, 16, 1} = one 64-column half of one page of one kv head, 128-byte swizzle.
static bool make_kv_map(CUtensorMap* map, const bf16* base, int nkv, int64_t stride_page) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {(cuuint64_t)THD, (cuuint64_t)nkv, (cuuint64_t)TPAGE, (cuuint64_t)1 << 20};  // page ids come from the page table; the extent only bounds them
  cuuint64_t strides[3] = {(cuuint64_t)THD * 2, (cuuint64_t)nkv * THD * 2, (cuuint64_t)stride_page * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)TPAGE, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Launch for the paged batch-prefill entry (prefill_attention.cu dispatches here).  Returns cudaError as int,
// -2 when the pool cannot be described by a TMA tensor map (caller falls back to the mma.sync kernel).
int launch_prefill_tc(const bf16* q, bf16* out, const bf16* k_base, const bf16* v_base, const int* page_indices,
                      const int* page_indptr, const int* last_page_len, const int* q_indptr, int seq_len, int batch_size,
                      int nq, int nkv, int page_size, int64_t stride_page, float sm_scale_log2, cudaStream_t stream) {
  if (page_size != TPAGE || (stride_page * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(k_base) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(v_base) & 15) != 0 || (reinterpret_cast<uintptr_t>(q) & 15) != 0)
    return -2;
  CUtensorMap mk, mv;
  if (!make_kv_map(&mk, k_base, nkv, stride_page) || !make_kv_map(&mv, v_base, nkv, stride_page)) return -2;
  PrefillTcArgs a{};
  a.q = q; a.out = out;
  a.page_indices = page_indices; a.page_indptr = page_indptr; a.last_page_len = last_page_len; a.q_indptr = q_indptr;
  a.seq_len = seq_len; a.batch_size = batch_size; a.nq = nq; a.nkv = nkv;
  a.sm_scale_log2 = sm_scale_log2;
  constexpr size_t smem = 6 * TILE_B + 1024 + 256 + 2048;
  static thread_local bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(prefill_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cfg = true;
  }
  const int tiles = (seq_len + TQ - 1) / TQ + batch_size;  // upper bound; surplus CTAs exit at once
  return (int)launch(prefill_attention_tc_kernel, dim3(nq, tiles), dim3(T_THREADS), smem, stream, true, mk, mv, a);
}

}  // namespace pk
