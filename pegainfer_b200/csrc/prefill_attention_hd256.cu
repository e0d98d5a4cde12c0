// Causal GQA prefill attention over the paged KV cache for head_dim 256 (Qwen3.5 full-attention layers):
// `batch_prefill_paged_cuda_hd256` (ffi.rs:1309-1334; the reference instantiates FlashInfer FA2
// BatchPrefillWithPagedKVCache for HD 256, csrc/paged_attention.cu:749-838 and prefill_attention_hd256.cu:24-386).
//
// Same structure as the mma.sync kernel of prefill_attention.cu, re-derived for 512-byte rows:
//   * warp = 16 query tokens of one q head, CTA = (8 / GROUP) token blocks x the GROUP q heads of one kv head, so a K/V
//     tile staged in shared memory is reused by the whole GQA group
//   * K/V tiles of 32 kv tokens x 256 dims gathered from 16-token pages with cp.async, double buffered, rows XOR-swizzled
//     in 16-byte chunks (conflict-free ldmatrix / ldmatrix.trans)
//   * the 16 x 256 Q tile of a warp lives in shared memory and its A fragments are re-read per K step (64 registers of
//     fragments + 128 of O accumulators would not fit next to the score tile)
// Rounding points follow flashinfer/attention/prefill.cuh like the HD-128 kernels: S from bf16 operands with fp32
// accumulation, fp32 softmax with exp2 and a running max, P rounded to bf16 for the PV product, denominator = row sum
// of the ROUNDED P, O accumulated in fp32, O / d rounded once.
// The caller's tile plan (request / qo-tile / kv-tile indices) is ignored consistently: tiles come from q_indptr.
#include "common.cuh"

namespace pk {

namespace {

constexpr int H2D = 256;                 // head dim
constexpr int H2_ROW = H2D * 2;          // bytes per row
constexpr int H2_KV_TILE = 32;           // kv tokens per shared-memory tile
constexpr int H2_PWARPS = 8;
constexpr int H2_PTHREADS = H2_PWARPS * 32;
constexpr int H2_TILE_BYTES = H2_KV_TILE * H2_ROW;   // 16 KB
constexpr int H2_QTILE_BYTES = 16 * H2_ROW;          // 8 KB per warp

struct Hd256PrefillArgs {
  const bf16* q;
  bf16* out;
  const bf16* k_base;
  const bf16* v_base;
  const int *page_indices, *page_indptr, *last_page_len, *q_indptr;
  int seq_len, batch_size, nq, nkv;
  int64_t stride_page;
  float sm_scale_log2;
};

__device__ __forceinline__ void h2p_cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void h2p_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void h2p_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void h2p_ldsm4(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void h2p_ldsm4_t(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void h2p_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float h2p_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of 16-byte chunk c (0..31) of row r inside a tile of 512-byte rows (chunk index XOR-swizzled in its low 3 bits)
__device__ __forceinline__ int h2p_sw(int r, int c) { return r * H2_ROW + (((c & ~7) | ((c & 7) ^ (r & 7))) << 4); }

template <int GROUP>
__global__ void __launch_bounds__(H2_PTHREADS, 1)
prefill_attention_hd256_kernel(const Hd256PrefillArgs a) {
  constexpr int TOK_BLOCKS = H2_PWARPS / GROUP;
  constexpr int TOK_PER_CTA = 16 * TOK_BLOCKS;
  extern __shared__ __align__(128) uint8_t h2p_smem[];
  uint8_t* Ks = h2p_smem;                              // [2][32 rows][512 B]
  uint8_t* Vs = h2p_smem + 2 * H2_TILE_BYTES;
  uint8_t* Qs = h2p_smem + 4 * H2_TILE_BYTES;          // [8 warps][16 rows][512 B]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int kvh = blockIdx.y;

  // ---- locate (request, token tile): heavy (late) tiles first ----
  int b = 0, q_start = 0, qo_len = 0, tile_local = -1;
  {
    int total_tiles = 0;
    for (int i = 0; i < a.batch_size; ++i) total_tiles += (a.q_indptr[i + 1] - a.q_indptr[i] + TOK_PER_CTA - 1) / TOK_PER_CTA;
    int idx = total_tiles - 1 - (int)blockIdx.x;
    if (idx < 0) return;
    for (int i = 0; i < a.batch_size; ++i) {
      const int len = a.q_indptr[i + 1] - a.q_indptr[i];
      const int nt = (len + TOK_PER_CTA - 1) / TOK_PER_CTA;
      if (idx < nt) {
        b = i; q_start = a.q_indptr[i]; qo_len = len; tile_local = idx;
        break;
      }
      idx -= nt;
    }
    if (tile_local < 0) return;
  }
  const int np = a.page_indptr[b + 1] - a.page_indptr[b];
  const int kv_len = np <= 0 ? 0 : (np - 1) * 16 + a.last_page_len[b];
  const int* pages = a.page_indices + a.page_indptr[b];
  const int t0 = tile_local * TOK_PER_CTA;
  const int causal_off = kv_len - qo_len;  // query token t attends kv <= t + causal_off
  const int kv_end = min(kv_len, causal_off + min(qo_len, t0 + TOK_PER_CTA));
  const int n_tiles = (kv_end + H2_KV_TILE - 1) / H2_KV_TILE;

  pdl_wait();

  // ---- this warp's 16 query rows of one head -> shared memory (swizzled), A fragments are re-read per K step ----
  const int tb = warp / GROUP, hq = warp % GROUP;
  const int head = kvh * GROUP + hq;
  const int tok_lo = t0 + tb * 16 + g, tok_hi = tok_lo + 8;
  const bool ok_lo = tok_lo < qo_len, ok_hi = tok_hi < qo_len;
  uint8_t* Qw = Qs + warp * H2_QTILE_BYTES;
  for (int i = lane; i < 16 * 32; i += 32) {
    const int r = i >> 5, c = i & 31;
    const int tok = t0 + tb * 16 + r;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (tok < qo_len) v = *reinterpret_cast<const uint4*>(a.q + ((size_t)(q_start + tok) * a.nq + head) * H2D + c * 8);
    *reinterpret_cast<uint4*>(Qw + h2p_sw(r, c)) = v;
  }
  __syncwarp();

  float o[32][4];
#pragma unroll
  for (int i = 0; i < 32; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, d_lo = 0.f, d_hi = 0.f;

  auto load_tile = [&](int j, int stage) {
    // 32 rows x 32 chunks for K and for V; 256 threads x 4 chunks each
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * H2_PTHREADS;  // 0..1023
      const int r = idx >> 5, c = idx & 31;
      const int kv = j * H2_KV_TILE + r;
      const bool valid = kv < kv_len;
      int64_t off = 0;
      if (valid) {
        const int page = __ldg(pages + (kv >> 4)), slot = kv & 15;
        off = (int64_t)page * a.stride_page + ((int64_t)slot * a.nkv + kvh) * H2D + c * 8;
      }
      h2p_cp_async16(Ks + stage * H2_TILE_BYTES + h2p_sw(r, c), a.k_base + off, valid);
      h2p_cp_async16(Vs + stage * H2_TILE_BYTES + h2p_sw(r, c), a.v_base + off, valid);
    }
    h2p_commit();
  };

  const int mat = lane >> 3, rr = lane & 7;
  if (n_tiles > 0) load_tile(0, 0);
  for (int j = 0; j < n_tiles; ++j) {
    const int stage = j & 1;
    if (j + 1 < n_tiles) {
      load_tile(j + 1, stage ^ 1);
      h2p_wait<1>();
    } else {
      h2p_wait<0>();
    }
    __syncthreads();
    const uint8_t* Kt = Ks + stage * H2_TILE_BYTES;
    const uint8_t* Vt = Vs + stage * H2_TILE_BYTES;

    // ---- S = Q K^T : 16 x 32 per warp, 16 K steps ----
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll 4
    for (int ks = 0; ks < 16; ++ks) {
      uint32_t qa[4];
      h2p_ldsm4(qa, Qw + h2p_sw((mat & 1) * 8 + rr, ks * 2 + (mat >> 1)));
#pragma unroll
      for (int npair = 0; npair < 2; ++npair) {
        const int row = npair * 16 + (mat >> 1) * 8 + rr;
        uint32_t kb[4];
        h2p_ldsm4(kb, Kt + h2p_sw(row, ks * 2 + (mat & 1)));
        h2p_mma(s[npair * 2], qa, kb[0], kb[1]);
        h2p_mma(s[npair * 2 + 1], qa, kb[2], kb[3]);
      }
    }
    // ---- mask + online softmax (scaled log2 domain) ----
    const int lim_lo = min(kv_len - 1, tok_lo + causal_off);
    const int lim_hi = min(kv_len - 1, tok_hi + causal_off);
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int kv0 = j * H2_KV_TILE + nt * 8 + t4 * 2;
      s[nt][0] = (ok_lo && kv0 <= lim_lo) ? s[nt][0] * a.sm_scale_log2 : -INFINITY;
      s[nt][1] = (ok_lo && kv0 + 1 <= lim_lo) ? s[nt][1] * a.sm_scale_log2 : -INFINITY;
      s[nt][2] = (ok_hi && kv0 <= lim_hi) ? s[nt][2] * a.sm_scale_log2 : -INFINITY;
      s[nt][3] = (ok_hi && kv0 + 1 <= lim_hi) ? s[nt][3] * a.sm_scale_log2 : -INFINITY;
      mx_lo = fmaxf(mx_lo, fmaxf(s[nt][0], s[nt][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s[nt][2], s[nt][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    const float ref_lo = mn_lo == -INFINITY ? 0.f : mn_lo, ref_hi = mn_hi == -INFINITY ? 0.f : mn_hi;
    const float sc_lo = h2p_ex2(m_lo - ref_lo), sc_hi = h2p_ex2(m_hi - ref_hi);  // m = -inf -> 0
    m_lo = mn_lo;
    m_hi = mn_hi;
    d_lo *= sc_lo;
    d_hi *= sc_hi;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      o[i][0] *= sc_lo; o[i][1] *= sc_lo; o[i][2] *= sc_hi; o[i][3] *= sc_hi;
    }
    uint32_t pa[2][4];  // P as bf16 A fragments, 2 K steps of 16 kv tokens
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float p0 = round_bf16(h2p_ex2(s[nt][0] - ref_lo)), p1 = round_bf16(h2p_ex2(s[nt][1] - ref_lo));
      const float p2 = round_bf16(h2p_ex2(s[nt][2] - ref_hi)), p3 = round_bf16(h2p_ex2(s[nt][3] - ref_hi));
      d_lo += p0 + p1;
      d_hi += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    // ---- O += P V : 16 x 256 per warp ----
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int dp = 0; dp < 16; ++dp) {
        const int row = ks * 16 + (mat & 1) * 8 + rr;
        uint32_t vb[4];
        h2p_ldsm4_t(vb, Vt + h2p_sw(row, dp * 2 + (mat >> 1)));
        h2p_mma(o[dp * 2], pa[ks], vb[0], vb[1]);
        h2p_mma(o[dp * 2 + 1], pa[ks], vb[2], vb[3]);
      }
    }
    __syncthreads();  // everyone done with this stage before it is refilled
  }

  // ---- finalize ----
  d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 1);
  d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 2);
  d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 1);
  d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 2);
  const float inv_lo = (ok_lo && d_lo > 0.f) ? __fdividef(1.f, d_lo) : 0.f;
  const float inv_hi = (ok_hi && d_hi > 0.f) ? __fdividef(1.f, d_hi) : 0.f;
  bf16* out_lo = a.out + ((size_t)(q_start + tok_lo) * a.nq + head) * H2D;
  bf16* out_hi = a.out + ((size_t)(q_start + tok_hi) * a.nq + head) * H2D;
#pragma unroll
  for (int nt = 0; nt < 32; ++nt) {
    const int c = nt * 8 + t4 * 2;
    if (ok_lo) *reinterpret_cast<uint32_t*>(out_lo + c) = pack_bf16(o[nt][0] * inv_lo, o[nt][1] * inv_lo);
    if (ok_hi) *reinterpret_cast<uint32_t*>(out_hi + c) = pack_bf16(o[nt][2] * inv_hi, o[nt][3] * inv_hi);
  }
}

}  // namespace
}  // namespace pk

using namespace pk;

extern "C" int batch_prefill_paged_cuda_hd256(const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data, int64_t k_offset_elems,
                                              int64_t v_offset_elems, const int* page_indices, const int* page_indptr,
                                              const int* last_page_len_d, const int* q_indptr, const int* request_indices,
                                              const int* qo_tile_indices, const int* kv_tile_indices, const int* kv_chunk_size_ptr,
                                              const uint32_t* total_num_rows, int num_qo_heads, int num_kv_heads, int head_dim,
                                              int page_size, int seq_len, int batch_size, int padded_batch_size, int64_t stride_page,
                                              float sm_scale, pk_stream stream) {
  (void)request_indices; (void)qo_tile_indices; (void)kv_tile_indices; (void)kv_chunk_size_ptr; (void)total_num_rows;
  (void)padded_batch_size;
  if (head_dim != H2D || page_size != 16 || !q_indptr || num_kv_heads <= 0 || num_qo_heads % num_kv_heads != 0) return -1;
  if (seq_len <= 0 || batch_size <= 0) return 0;
  Hd256PrefillArgs a{};
  a.q = (const bf16*)q; a.out = (bf16*)output;
  a.k_base = (const bf16*)kv_data + k_offset_elems; a.v_base = (const bf16*)kv_data + v_offset_elems;
  a.page_indices = page_indices; a.page_indptr = page_indptr; a.last_page_len = last_page_len_d; a.q_indptr = q_indptr;
  a.seq_len = seq_len; a.batch_size = batch_size; a.nq = num_qo_heads; a.nkv = num_kv_heads;
  a.stride_page = stride_page;
  a.sm_scale_log2 = sm_scale * 1.44269504088896340736f;
  const int group = num_qo_heads / num_kv_heads;
  constexpr size_t smem = 4 * H2_TILE_BYTES + H2_PWARPS * H2_QTILE_BYTES;
#define PK_LAUNCH_HD256(G)                                                                            \
  {                                                                                                   \
    auto kern = prefill_attention_hd256_kernel<G>;                                                    \
    static thread_local bool cfg = false;                                                             \
    if (!cfg) {                                                                                       \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
      cfg = true;                                                                                     \
    }                                                                                                 \
    const int tok_per_cta = 16 * (H2_PWARPS / G);                                                     \
    const int tiles = (seq_len + tok_per_cta - 1) / tok_per_cta + batch_size;                         \
    return (int)launch(kern, dim3(tiles, num_kv_heads), dim3(H2_PTHREADS), smem, stream, true, a);    \
  }
  switch (group) {
    case 1: PK_LAUNCH_HD256(1)
    case 2: PK_LAUNCH_HD256(2)
    case 4: PK_LAUNCH_HD256(4)
    case 8: PK_LAUNCH_HD256(8)
    default: return -1;
  }
#undef PK_LAUNCH_HD256
}
