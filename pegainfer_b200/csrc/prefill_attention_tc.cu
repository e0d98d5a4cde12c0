// Causal GQA prefill attention over the paged KV cache on the 5th-generation tensor cores (tcgen05 + TMEM).
// Same contract and rounding points as prefill_attention.cu (the mma.sync version): S = Q K^T from bf16 operands with
// fp32 accumulation, fp32 online softmax in the exp2 domain, P rounded to bf16 for the P V product, denominator =
// row sum of the ROUNDED P, O accumulated in fp32, O/d rounded once to bf16 (flashinfer prefill.cuh:956-985).
// The running maximum advances per 128-token KV block (the mma.sync kernel: per 64) -- same mathematics, P differs
// by its bf16 rounding relative to a different reference, covered by the test tolerance.
//
// One CTA = 128 query tokens of ONE q head (TMEM lane = token).  Per 128-token KV block j:
//   loader warps (5-7): gather the block's K and V rows from their 16-token pages with 16-byte cp.async copies
//                       straight into the 128-byte-swizzled shared-memory image UMMA expects, 2-stage ring
//   MMA warp (4)      : S[b] = Q K^T   (M128 x N128 x K16 x 8, both operands K-major) into TMEM buffer b = j & 1,
//                       issued one block AHEAD of the softmax; then O_j = P V (A = P K-major, B = V MN-major: the
//                       V tile is [kv token][head dim], i.e. N-contiguous) into a third TMEM region
//   softmax warps (0-3): thread = query row: tcgen05.ld S (two passes: max, then exp2), bf16 P -> swizzled smem,
//                       then O = O * alpha + (P V) with O held in 128 registers per thread
// TMEM: S0 [0,128) | S1 [128,256) | PV [256,384) of a 512-column allocation.  Shared memory: Q 32 KB, P 32 KB,
// 2 x (K 32 KB + V 32 KB) = 192 KB -> one CTA per SM.
// Tensor-bound: 4 * 128 * kv flop per (query token, head); causal blocks past the diagonal are skipped.
#include "tcgen05.cuh"

namespace pk {

constexpr int THD = 128;            // head dim
constexpr int TQ = 128;             // query tokens per CTA (UMMA M)
constexpr int TKV = 128;            // kv tokens per block (UMMA N of S, K extent of P V)
constexpr int HALF_BYTES = 128 * 128;  // one 64-column half of a [128 x 128] bf16 tile: 128 rows x 128 B
constexpr int TILE_B = 2 * HALF_BYTES;  // 32 KB
constexpr int T_SOFTMAX_WARPS = 4, T_LOADER_WARPS = 3;
constexpr int T_THREADS = (T_SOFTMAX_WARPS + 1 + T_LOADER_WARPS) * 32;  // 256
constexpr int T_LOADERS = T_LOADER_WARPS * 32;
constexpr uint32_t T_TMEM_COLS = 512;

struct PrefillTcArgs {
  const bf16* q;
  bf16* out;
  const bf16* k_base;  // pool + k_off
  const bf16* v_base;
  const int* page_indices;
  const int* page_indptr;
  const int* last_page_len;
  const int* q_indptr;
  int seq_len, batch_size, nq, nkv, page_size;
  int64_t stride_page;
  float sm_scale_log2;
  int v_desc_mode;  // 0: LBO = half stride, SBO = 8-row group (canonical); 1: swapped (bring-up switch)
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
// A hung barrier becomes a trap (reported as a launch failure) instead of a wedged GPU.
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 27)) __trap();
}
// byte offset of 16-byte chunk `c16` (0..15) of row `r` in a [128 rows x 128 cols] bf16 tile stored as two
// 64-column halves, each half = 128 rows x 128 B with the 128-byte swizzle (chunk ^= row & 7)
__device__ __forceinline__ uint32_t sw_off(int r, int c16) {
  return (uint32_t)((c16 >> 3) * HALF_BYTES + r * 128 + (((c16 & 7) ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(T_THREADS, 1)
prefill_attention_tc_kernel(const PrefillTcArgs a) {
  extern __shared__ uint8_t tsm_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tsm_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem;
  uint8_t* Ps = smem + TILE_B;
  uint8_t* KVs = smem + 2 * TILE_B;  // stage s: K at KVs + s*2*TILE_B, V right after
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_B);
  uint64_t* kv_full = bars;       // [2], count T_LOADERS
  uint64_t* kv_empty = bars + 2;  // [2], count 1 (tcgen05.commit)
  uint64_t* s_full = bars + 4;    // [2], count 1
  uint64_t* p_full = bars + 6;    // count 128
  uint64_t* pv_full = bars + 7;   // count 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int kvh = head / (a.nq / a.nkv);

  // ---- locate (request, 128-token tile): late (heavy) tiles first ----
  int b = 0, q_start = 0, qo_len = a.seq_len, tile_local = -1;
  {
    int total_tiles = 0;
    for (int i = 0; i < a.batch_size; ++i) total_tiles += (a.q_indptr[i + 1] - a.q_indptr[i] + TQ - 1) / TQ;
    int idx = total_tiles - 1 - (int)blockIdx.x;
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
  const int kv_len = np <= 0 ? 0 : (np - 1) * a.page_size + a.last_page_len[b];
  const int* pages = a.page_indices + a.page_indptr[b];
  const int t0 = tile_local * TQ;
  const int causal_off = kv_len - qo_len;  // query token t attends kv <= t + causal_off
  const int kv_end = min(kv_len, causal_off + min(qo_len, t0 + TQ));
  const int n_blocks = (kv_end + TKV - 1) / TKV;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full + s, T_LOADERS);
      mbar_init(kv_empty + s, 1);
      mbar_init(s_full + s, 1);
    }
    mbar_init(p_full, T_SOFTMAX_WARPS * 32);
    mbar_init(pv_full, 1);
    mbar_fence_init();
  }
  if (warp == T_SOFTMAX_WARPS) tmem_alloc(tmem_slot, T_TMEM_COLS);
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

  if (warp >= T_SOFTMAX_WARPS + 1) {
    // =========================== loader warps ===========================
    const int ltid = threadIdx.x - (T_SOFTMAX_WARPS + 1) * 32;
    for (int j = 0; j < n_blocks; ++j) {
      const int s = j & 1;
      mbar_wait_or_trap(kv_empty + s, (uint32_t)(((j >> 1) & 1) ^ 1));
      const uint32_t kdst = smem_u32(KVs + (size_t)s * 2 * TILE_B), vdst = kdst + TILE_B;
      for (int idx = ltid; idx < TKV * 32; idx += T_LOADERS) {
        const int r = idx >> 5, c = idx & 31;  // c < 16: K chunk, else V chunk
        const int kv = j * TKV + r;
        const bool valid = kv < kv_len;
        int64_t off = 0;
        if (valid) {
          const int page = __ldg(pages + kv / a.page_size), slot = kv % a.page_size;
          off = (int64_t)page * a.stride_page + ((int64_t)slot * a.nkv + kvh) * THD + (c & 15) * 8;
        }
        if (c < 16) tcp_async16(kdst + sw_off(r, c), a.k_base + off, valid);
        else tcp_async16(vdst + sw_off(r, c - 16), a.v_base + off, valid);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      // Publish block j before touching the next stage: the MMA warp issues S(j+1) ahead of P V(j), so waiting for
      // kv_empty (= P V(j-1) retired) with block j still unannounced would deadlock.  Block j+1 streams in while
      // block j is being computed (two stages).
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();
      mbar_arrive(kv_full + s);
    }
  } else if (warp == T_SOFTMAX_WARPS) {
    // =========================== MMA issuer ===========================
    // D = f32, A = B = bf16, M = 128, N = 128; P V additionally reads B (= V) MN-major
    constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TKV >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
    constexpr uint32_t idesc_pv = idesc_s | (1u << 16);
    const uint32_t q_addr = smem_u32(Qs), p_addr = smem_u32(Ps);
    auto issue_s = [&](int j) {
      const int s = j & 1;
      mbar_wait_or_trap(kv_full + s, (uint32_t)((j >> 1) & 1));
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
      }
      __syncwarp();
    };
    if (n_blocks > 0) issue_s(0);
    for (int j = 0; j < n_blocks; ++j) {
      if (j + 1 < n_blocks) issue_s(j + 1);  // runs under the softmax of block j
      mbar_wait_or_trap(p_full, (uint32_t)(j & 1));
      tc_fence_after();
      if (lane == 0) {
        const int s = j & 1;
        const uint32_t v_addr = smem_u32(KVs + (size_t)s * 2 * TILE_B) + TILE_B;
        const uint32_t d = tmem_base + 2 * TKV;
#pragma unroll
        for (int k = 0; k < TKV / 16; ++k) {
          const uint32_t ao = (uint32_t)((k >> 2) * HALF_BYTES + (k & 3) * 32);  // P: K-major, K = kv token
          const uint32_t bo = (uint32_t)(k * 2048);                              // V: 16 tokens = two 8-row groups
          const uint64_t bdesc = a.v_desc_mode == 0 ? make_sw128_mn_desc(v_addr + bo, HALF_BYTES, 1024)
                                                    : make_sw128_mn_desc(v_addr + bo, 1024, HALF_BYTES);
          umma_bf16(d, make_sw128_desc(p_addr + ao), bdesc, idesc_pv, k > 0 ? 1u : 0u);
        }
        umma_commit(pv_full);
        umma_commit(kv_empty + s);  // K_j (read by S_j) and V_j are free once everything issued so far retires
      }
      __syncwarp();
    }
  } else {
    // =========================== softmax warps: thread = query row ===========================
    const int r = threadIdx.x;  // 0..127 = TMEM lane
    const int tok = t0 + r;
    const bool row_ok = tok < qo_len;
    const int lim = row_ok ? min(kv_len - 1, tok + causal_off) : -1;  // last kv index this row may see
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    float o[THD];
#pragma unroll
    for (int i = 0; i < THD; ++i) o[i] = 0.f;
    float m = -INFINITY, d = 0.f;
    for (int j = 0; j < n_blocks; ++j) {
      const int sb = j & 1;
      mbar_wait_or_trap(s_full + sb, (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint32_t t_s = t_lane + (uint32_t)(sb * TKV);
      const int col0 = j * TKV;
      // pass 1: row maximum of the masked, scaled scores
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < TKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_s + (uint32_t)c, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + c + i <= lim) mx = fmaxf(mx, __uint_as_float(v[i]) * a.sm_scale_log2);
      }
      const float m_new = fmaxf(m, mx);
      const float ref = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = tex2(m - ref);  // m = -inf -> 0
      m = m_new;
      d *= alpha;
      // pass 2: P = bf16(exp2(s - ref)) -> swizzled smem (A operand of P V), d += rounded P
#pragma unroll 1
      for (int c = 0; c < TKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_s + (uint32_t)c, v);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = col0 + c + i <= lim ? round_bf16(tex2(__uint_as_float(v[i]) * a.sm_scale_log2 - ref)) : 0.f;
          const float p1 = col0 + c + i + 1 <= lim ? round_bf16(tex2(__uint_as_float(v[i + 1]) * a.sm_scale_log2 - ref)) : 0.f;
          d += p0 + p1;
          pk[i >> 1] = pack_bf16(p0, p1);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          *reinterpret_cast<uint4*>(Ps + sw_off(r, (c >> 3) + q4)) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
      }
      tc_fence_before();        // our tcgen05.ld of S[sb] are done before the MMA warp may overwrite it (block j+2)
      fence_proxy_async_smem();  // P visible to the tensor core's operand reads
      mbar_arrive(p_full);
      // O = O * alpha + P V
      mbar_wait_or_trap(pv_full, (uint32_t)(j & 1));
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < THD; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_lane + (uint32_t)(2 * TKV + c), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] = fmaf(o[c + i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();  // PV region may be overwritten by block j+1 only after p_full(j+1), which follows this
    }
    if (row_ok) {
      const float inv = __fdividef(1.f, d);
      uint4* dst = reinterpret_cast<uint4*>(a.out + ((size_t)(q_start + tok) * a.nq + head) * THD);
#pragma unroll
      for (int c = 0; c < THD; c += 8)
        dst[c >> 3] = make_uint4(pack_bf16(o[c] * inv, o[c + 1] * inv), pack_bf16(o[c + 2] * inv, o[c + 3] * inv),
                                 pack_bf16(o[c + 4] * inv, o[c + 5] * inv), pack_bf16(o[c + 6] * inv, o[c + 7] * inv));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == T_SOFTMAX_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, T_TMEM_COLS);
  }
}

// Launch for the paged batch-prefill entry (prefill_attention.cu dispatches here).  Returns cudaError as int.
int launch_prefill_tc(const bf16* q, bf16* out, const bf16* k_base, const bf16* v_base, const int* page_indices,
                      const int* page_indptr, const int* last_page_len, const int* q_indptr, int seq_len, int batch_size,
                      int nq, int nkv, int page_size, int64_t stride_page, float sm_scale_log2, int v_desc_mode,
                      cudaStream_t stream) {
  PrefillTcArgs a{};
  a.q = q; a.out = out; a.k_base = k_base; a.v_base = v_base;
  a.page_indices = page_indices; a.page_indptr = page_indptr; a.last_page_len = last_page_len; a.q_indptr = q_indptr;
  a.seq_len = seq_len; a.batch_size = batch_size; a.nq = nq; a.nkv = nkv; a.page_size = page_size;
  a.stride_page = stride_page; a.sm_scale_log2 = sm_scale_log2; a.v_desc_mode = v_desc_mode;
  constexpr size_t smem = 6 * TILE_B + 1024 + 256;
  static thread_local bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(prefill_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cfg = true;
  }
  const int tiles = (seq_len + TQ - 1) / TQ + batch_size;  // upper bound; surplus CTAs exit at once
  return (int)launch(prefill_attention_tc_kernel, dim3(tiles, nq), dim3(T_THREADS), smem, stream, true, a);
}

}  // namespace pk
