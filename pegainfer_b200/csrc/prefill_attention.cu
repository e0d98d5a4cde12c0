// Causal GQA prefill attention over the paged KV cache (flash-attention, online softmax).
// Replaces csrc/paged_attention.cu:343-608 (FlashInfer FA2 BatchPrefillWithPagedKVCache /
// SinglePrefillWithKVCache instantiations) of the reference.
//
// Rounding points follow flashinfer/attention/prefill.cuh: S = QK^T from bf16 operands with fp32
// accumulation; softmax in fp32 with exp2 and a running max; P is rounded to bf16 for the PV
// product and the denominator is the row sum of the ROUNDED P (prefill.cuh:956-985); O
// accumulates in fp32 and O/d is rounded once to bf16.
//
// This file: warp-level mma.sync.m16n8k16 bf16 tensor-core tiles (one warp = 16 query
// tokens of one q head; a CTA covers 8/GROUP token blocks x GROUP heads of one kv head so every
// K/V tile staged in shared memory is reused by the whole GQA group), cp.async double-buffered
// 64-token K/V tiles gathered from 16-token pages, XOR-swizzled rows (conflict-free ldmatrix).
// The paged batch-prefill entry runs the tcgen05/TMEM kernel of prefill_attention_tc.cu by default; this kernel
// serves single_prefill_cuda (contiguous K/V), page sizes other than 16 and PK_PREFILL_ATTN=legacy (A/B).
//
// The caller's tile plan (request_indices / qo_tile_indices / kv_tile_indices) is ignored
// consistently: tiles are derived from q_indptr on the device; results do not depend on tiling.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace pk {

constexpr int PHD = 128;       // head dim
constexpr int KV_TILE = 64;    // kv tokens per smem tile
constexpr int PWARPS = 8;
constexpr int PTHREADS = PWARPS * 32;
constexpr int TILE_BYTES = KV_TILE * PHD * 2;  // 16 KB

struct PrefillArgs {
  const bf16* q;
  bf16* out;
  const bf16* k_base;  // paged: pool + k_off ; contiguous: k_cache
  const bf16* v_base;
  const int* page_indices;
  const int* page_indptr;
  const int* last_page_len;
  const int* q_indptr;  // nullptr => single request [0, seq_len)
  int seq_len;          // total q tokens
  int batch_size;
  int nq, nkv, page_size;
  int64_t stride_page;
  float sm_scale_log2;
  // contiguous (single_prefill) mode
  int contiguous;
  int kv_len_single;
  int max_seq_len;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src),
               "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int GROUP>
__global__ void __launch_bounds__(PTHREADS)
prefill_attention_kernel(const PrefillArgs a) {
  constexpr int TOK_BLOCKS = PWARPS / GROUP;     // 16-token blocks per CTA
  constexpr int TOK_PER_CTA = 16 * TOK_BLOCKS;
  extern __shared__ __align__(128) uint8_t psm[];
  uint8_t* Ks = psm;                     // [2][KV_TILE][256 B] swizzled
  uint8_t* Vs = psm + 2 * TILE_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int kvh = blockIdx.y;

  // ---- locate (request, token tile): heavy (late) tiles first ----
  int b = 0, q_start = 0, qo_len = a.seq_len, tile_local = -1;
  {
    int total_tiles = 0;
    if (a.q_indptr) {
      for (int i = 0; i < a.batch_size; ++i)
        total_tiles += (a.q_indptr[i + 1] - a.q_indptr[i] + TOK_PER_CTA - 1) / TOK_PER_CTA;
    } else {
      total_tiles = (a.seq_len + TOK_PER_CTA - 1) / TOK_PER_CTA;
    }
    int idx = total_tiles - 1 - (int)blockIdx.x;
    if (idx < 0) return;
    if (a.q_indptr) {
      for (int i = 0; i < a.batch_size; ++i) {
        const int len = a.q_indptr[i + 1] - a.q_indptr[i];
        const int nt = (len + TOK_PER_CTA - 1) / TOK_PER_CTA;
        if (idx < nt) {
          b = i; q_start = a.q_indptr[i]; qo_len = len; tile_local = idx;
          break;
        }
        idx -= nt;
      }
    } else {
      tile_local = idx;
    }
    if (tile_local < 0) return;
  }
  int kv_len;
  const int* pages = nullptr;
  if (a.contiguous) {
    kv_len = a.kv_len_single;
  } else {
    const int np = a.page_indptr[b + 1] - a.page_indptr[b];
    kv_len = np <= 0 ? 0 : (np - 1) * a.page_size + a.last_page_len[b];
    pages = a.page_indices + a.page_indptr[b];
  }
  const int t0 = tile_local * TOK_PER_CTA;
  const int causal_off = kv_len - qo_len;  // query token t attends kv <= t + causal_off
  const int kv_end = min(kv_len, causal_off + min(qo_len, t0 + TOK_PER_CTA));
  const int n_tiles = (kv_end + KV_TILE - 1) / KV_TILE;

  pdl_wait();

  // ---- this warp's 16 query rows of one head -> A fragments ----
  const int tb = warp / GROUP, hq = warp % GROUP;
  const int head = kvh * GROUP + hq;
  const int tok_lo = t0 + tb * 16 + g, tok_hi = tok_lo + 8;  // rows g and g+8 (request-local)
  const bool ok_lo = tok_lo < qo_len, ok_hi = tok_hi < qo_len;
  uint32_t qa[8][4];
  {
    const bf16* q_lo = a.q + ((size_t)(q_start + tok_lo) * a.nq + head) * PHD;
    const bf16* q_hi = a.q + ((size_t)(q_start + tok_hi) * a.nq + head) * PHD;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const int c = ks * 16 + t4 * 2;
      qa[ks][0] = ok_lo ? *reinterpret_cast<const uint32_t*>(q_lo + c) : 0u;
      qa[ks][1] = ok_hi ? *reinterpret_cast<const uint32_t*>(q_hi + c) : 0u;
      qa[ks][2] = ok_lo ? *reinterpret_cast<const uint32_t*>(q_lo + c + 8) : 0u;
      qa[ks][3] = ok_hi ? *reinterpret_cast<const uint32_t*>(q_hi + c + 8) : 0u;
    }
  }

  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, d_lo = 0.f, d_hi = 0.f;

  auto load_tile = [&](int j, int stage) {
    // 64 rows x 16 chunks for K and for V; 256 threads x 4 chunks each
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * PTHREADS;  // 0..1023
      const int r = idx >> 4, c = idx & 15;
      const int kv = j * KV_TILE + r;
      const bool valid = kv < kv_len;
      int64_t off = 0;
      if (valid) {
        if (a.contiguous) {
          off = ((int64_t)kvh * a.max_seq_len + kv) * PHD + c * 8;
        } else {
          const int page = __ldg(pages + kv / a.page_size), slot = kv % a.page_size;
          off = (int64_t)page * a.stride_page + ((int64_t)slot * a.nkv + kvh) * PHD + c * 8;
        }
      }
      const int sw = (c ^ (r & 7)) * 16;
      cp_async16(Ks + stage * TILE_BYTES + r * 256 + sw, a.k_base + off, valid);
      cp_async16(Vs + stage * TILE_BYTES + r * 256 + sw, a.v_base + off, valid);
    }
    cp_async_commit();
  };

  if (n_tiles > 0) load_tile(0, 0);
  for (int j = 0; j < n_tiles; ++j) {
    const int stage = j & 1;
    if (j + 1 < n_tiles) {
      load_tile(j + 1, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint8_t* Kt = Ks + stage * TILE_BYTES;
    const uint8_t* Vt = Vs + stage * TILE_BYTES;

    // ---- S = Q K^T : 16 x 64 per warp ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    {
      const int mat = lane >> 3, rr = lane & 7;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          const int row = np * 16 + (mat >> 1) * 8 + rr;
          const int chunk = ks * 2 + (mat & 1);
          uint32_t kb[4];
          ldmatrix_x4(kb, Kt + row * 256 + ((chunk ^ (row & 7)) * 16));
          mma_bf16(s[np * 2], qa[ks], kb[0], kb[1]);
          mma_bf16(s[np * 2 + 1], qa[ks], kb[2], kb[3]);
        }
      }
    }
    // ---- mask + online softmax (scaled log2 domain) ----
    const int lim_lo = min(kv_len - 1, tok_lo + causal_off);
    const int lim_hi = min(kv_len - 1, tok_hi + causal_off);
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int kv0 = j * KV_TILE + nt * 8 + t4 * 2;
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
    const float sc_lo = ex2f(m_lo - ref_lo), sc_hi = ex2f(m_hi - ref_hi);  // m=-inf -> 0
    m_lo = mn_lo;
    m_hi = mn_hi;
    d_lo *= sc_lo;
    d_hi *= sc_hi;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i][0] *= sc_lo; o[i][1] *= sc_lo; o[i][2] *= sc_hi; o[i][3] *= sc_hi;
    }
    uint32_t pa[4][4];  // P as bf16 A fragments, 4 k-steps of 16 kv
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = round_bf16(ex2f(s[nt][0] - ref_lo)), p1 = round_bf16(ex2f(s[nt][1] - ref_lo));
      const float p2 = round_bf16(ex2f(s[nt][2] - ref_hi)), p3 = round_bf16(ex2f(s[nt][3] - ref_hi));
      d_lo += p0 + p1;
      d_hi += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    // ---- O += P V ----
    {
      const int mat = lane >> 3, rr = lane & 7;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int dp = 0; dp < 8; ++dp) {
          const int row = ks * 16 + (mat & 1) * 8 + rr;
          const int chunk = dp * 2 + (mat >> 1);
          uint32_t vb[4];
          ldmatrix_x4_trans(vb, Vt + row * 256 + ((chunk ^ (row & 7)) * 16));
          mma_bf16(o[dp * 2], pa[ks], vb[0], vb[1]);
          mma_bf16(o[dp * 2 + 1], pa[ks], vb[2], vb[3]);
        }
      }
    }
    __syncthreads();  // everyone done with this stage before it is refilled
  }

  // ---- finalize ----
  d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 1);
  d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 2);
  d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 1);
  d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 2);
  const float inv_lo = ok_lo ? __fdividef(1.f, d_lo) : 0.f;
  const float inv_hi = ok_hi ? __fdividef(1.f, d_hi) : 0.f;
  bf16* out_lo = a.out + ((size_t)(q_start + tok_lo) * a.nq + head) * PHD;
  bf16* out_hi = a.out + ((size_t)(q_start + tok_hi) * a.nq + head) * PHD;
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    const int c = nt * 8 + t4 * 2;
    if (ok_lo) *reinterpret_cast<uint32_t*>(out_lo + c) = pack_bf16(o[nt][0] * inv_lo, o[nt][1] * inv_lo);
    if (ok_hi) *reinterpret_cast<uint32_t*>(out_hi + c) = pack_bf16(o[nt][2] * inv_hi, o[nt][3] * inv_hi);
  }
}

static int launch_prefill(const PrefillArgs& a, cudaStream_t stream) {
  if (a.nkv <= 0 || a.nq % a.nkv != 0) return -1;
  const int group = a.nq / a.nkv;
  const size_t smem = 4 * TILE_BYTES;
#define PK_LAUNCH_PREFILL(G)                                                                       \
  {                                                                                                \
    auto kern = prefill_attention_kernel<G>;                                                       \
    static thread_local bool cfg = false;                                                          \
    if (!cfg) {                                                                                    \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
      cfg = true;                                                                                  \
    }                                                                                              \
    const int tok_per_cta = 16 * (PWARPS / G);                                                     \
    const int tiles = (a.seq_len + tok_per_cta - 1) / tok_per_cta + (a.q_indptr ? a.batch_size : 0); \
    return (int)launch(kern, dim3(tiles, a.nkv), dim3(PTHREADS), smem, stream, true, a);           \
  }
  switch (group) {
    case 1: PK_LAUNCH_PREFILL(1)
    case 2: PK_LAUNCH_PREFILL(2)
    case 4: PK_LAUNCH_PREFILL(4)
    case 8: PK_LAUNCH_PREFILL(8)
    default: return -1;
  }
#undef PK_LAUNCH_PREFILL
}

// flashinfer/utils.cuh:384-403 FA2DetermineCtaTileQ (compute capability >= 8 branch)
static int fa2_cta_tile_q(int64_t packed_qo_len, int head_dim) {
  if (packed_qo_len > 64 && head_dim < 256) return 128;
  return packed_qo_len > 16 ? 64 : 16;
}
static int resolve_cta_tile_q(int64_t packed_qo_len, int head_dim, int override_) {
  if (override_ == 0) return fa2_cta_tile_q(packed_qo_len, head_dim);
  if (override_ == 16 || override_ == 64 || override_ == 128) return override_;
  return 0;
}

// prefill_attention_tc.cu
int launch_prefill_tc(const bf16* q, bf16* out, const bf16* k_base, const bf16* v_base, const int* page_indices,
                      const int* page_indptr, const int* last_page_len, const int* q_indptr, int seq_len, int batch_size,
                      int nq, int nkv, int page_size, int64_t stride_page, float sm_scale_log2, cudaStream_t stream);

// prefill_attention_tc2.cu
int launch_prefill_tc2(const bf16* q, bf16* out, const bf16* k_base, const bf16* v_base, const int* page_indices,
                       const int* page_indptr, const int* last_page_len, const int* q_indptr, int seq_len, int batch_size, int nq,
                       int nkv, int page_size, int64_t stride_page, float sm_scale_log2, cudaStream_t stream);

// PK_PREFILL_ATTN = tc2 (default: two query tiles per CTA, O and P in TMEM) | tc (one tile per CTA, O in registers) |
// legacy (the mma.sync kernel) for the paged batch-prefill entry.  Read per call: the A/B tools flip it in-process.
static int prefill_attn_impl() {
  const char* e = getenv("PK_PREFILL_ATTN");
  if (e && strcmp(e, "legacy") == 0) return 0;
  if (e && strcmp(e, "tc") == 0) return 1;
  return 2;
}
static int launch_prefill_tensor(int impl, const bf16* q, bf16* out, const bf16* k_base, const bf16* v_base, const int* page_indices,
                                 const int* page_indptr, const int* last_page_len, const int* q_indptr, int seq_len, int batch_size,
                                 int nq, int nkv, int page_size, int64_t stride_page, float sm_scale_log2, cudaStream_t stream) {
  return (impl == 2 ? launch_prefill_tc2 : launch_prefill_tc)(q, out, k_base, v_base, page_indices, page_indptr, last_page_len, q_indptr,
                                                              seq_len, batch_size, nq, nkv, page_size, stride_page, sm_scale_log2, stream);
}
}  // namespace pk

using namespace pk;

extern "C" {

int batch_prefill_paged_num_tiles(int seq_len, int num_qo_heads, int num_kv_heads, int head_dim) {
  const int64_t packed = (int64_t)seq_len * (num_qo_heads / num_kv_heads);
  const int t = fa2_cta_tile_q(packed, head_dim);
  return (int)((packed + t - 1) / t);
}
int batch_prefill_paged_num_tiles_with_cta_tile_q(int seq_len, int num_qo_heads, int num_kv_heads,
                                                  int head_dim, int cta_tile_q_override) {
  const int64_t packed = (int64_t)seq_len * (num_qo_heads / num_kv_heads);
  const int t = resolve_cta_tile_q(packed, head_dim, cta_tile_q_override);
  if (t == 0) return -1;
  return (int)((packed + t - 1) / t);
}
int batch_prefill_cta_tile_q(int total_seq_len, int num_qo_heads, int num_kv_heads, int head_dim) {
  return fa2_cta_tile_q((int64_t)total_seq_len * (num_qo_heads / num_kv_heads), head_dim);
}
int batch_prefill_cta_tile_q_with_override(int total_seq_len, int num_qo_heads, int num_kv_heads,
                                           int head_dim, int cta_tile_q_override) {
  return resolve_cta_tile_q((int64_t)total_seq_len * (num_qo_heads / num_kv_heads), head_dim,
                            cta_tile_q_override);
}

int batch_prefill_paged_cuda_with_cta_tile_q(
    const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data, int64_t k_offset_elems,
    int64_t v_offset_elems, const int* page_indices, const int* page_indptr,
    const int* last_page_len_d, const int* q_indptr, const int* request_indices,
    const int* qo_tile_indices, const int* kv_tile_indices, const int* kv_chunk_size_ptr,
    const uint32_t* total_num_rows, int num_qo_heads, int num_kv_heads, int head_dim,
    int page_size, int seq_len, int batch_size, int padded_batch_size, int64_t stride_page,
    float sm_scale, int cta_tile_q_override, pk_stream stream) {
  (void)request_indices; (void)qo_tile_indices; (void)kv_tile_indices; (void)kv_chunk_size_ptr;
  (void)total_num_rows; (void)padded_batch_size;
  if (head_dim != PHD) return -1;
  if (resolve_cta_tile_q((int64_t)seq_len * (num_qo_heads / (num_kv_heads > 0 ? num_kv_heads : 1)),
                         head_dim, cta_tile_q_override) == 0)
    return -1;  // invalid tile override, as the reference
  if (seq_len <= 0 || batch_size <= 0) return 0;
  if (prefill_attn_impl() >= 1 && q_indptr && page_size == 16 && num_kv_heads > 0 && num_qo_heads % num_kv_heads == 0) {
    const int rc = launch_prefill_tensor(prefill_attn_impl(), (const bf16*)q, (bf16*)output, (const bf16*)kv_data + k_offset_elems,
                                     (const bf16*)kv_data + v_offset_elems, page_indices, page_indptr, last_page_len_d,
                                     q_indptr, seq_len, batch_size, num_qo_heads, num_kv_heads, page_size, stride_page,
                                     sm_scale * 1.44269504088896340736f, stream);
    if (rc != -2) return rc;  // -2: pool not expressible as a TMA tensor map -> mma.sync kernel below
  }
  PrefillArgs a{};
  a.q = (const bf16*)q; a.out = (bf16*)output;
  a.k_base = (const bf16*)kv_data + k_offset_elems;
  a.v_base = (const bf16*)kv_data + v_offset_elems;
  a.page_indices = page_indices; a.page_indptr = page_indptr; a.last_page_len = last_page_len_d;
  a.q_indptr = q_indptr; a.seq_len = seq_len; a.batch_size = batch_size;
  a.nq = num_qo_heads; a.nkv = num_kv_heads; a.page_size = page_size; a.stride_page = stride_page;
  a.sm_scale_log2 = sm_scale * 1.44269504088896340736f;
  return launch_prefill(a, stream);
}

int batch_prefill_paged_cuda(const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data,
                             int64_t k_offset_elems, int64_t v_offset_elems,
                             const int* page_indices, const int* page_indptr,
                             const int* last_page_len_d, const int* q_indptr,
                             const int* request_indices, const int* qo_tile_indices,
                             const int* kv_tile_indices, const int* kv_chunk_size_ptr,
                             const uint32_t* total_num_rows, int num_qo_heads, int num_kv_heads,
                             int head_dim, int page_size, int seq_len, int batch_size,
                             int padded_batch_size, int64_t stride_page, float sm_scale,
                             pk_stream stream) {
  return batch_prefill_paged_cuda_with_cta_tile_q(
      q, output, kv_data, k_offset_elems, v_offset_elems, page_indices, page_indptr,
      last_page_len_d, q_indptr, request_indices, qo_tile_indices, kv_tile_indices,
      kv_chunk_size_ptr, total_num_rows, num_qo_heads, num_kv_heads, head_dim, page_size, seq_len,
      batch_size, padded_batch_size, stride_page, sm_scale, 0, stream);
}

int single_prefill_cuda(const pk_bf16* q, pk_bf16* output, const pk_bf16* k_cache,
                        const pk_bf16* v_cache, int num_qo_heads, int num_kv_heads, int head_dim,
                        int seq_len, int kv_len, int max_seq_len, float sm_scale,
                        pk_stream stream) {
  if (head_dim != PHD) return -1;
  if (seq_len <= 0) return 0;
  PrefillArgs a{};
  a.q = (const bf16*)q; a.out = (bf16*)output;
  a.k_base = (const bf16*)k_cache; a.v_base = (const bf16*)v_cache;
  a.seq_len = seq_len; a.batch_size = 1;
  a.nq = num_qo_heads; a.nkv = num_kv_heads;
  a.sm_scale_log2 = sm_scale * 1.44269504088896340736f;
  a.contiguous = 1; a.kv_len_single = kv_len; a.max_seq_len = max_seq_len;
  return launch_prefill(a, stream);
}

int pk_b200_prefill_attention_tc(const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data, int64_t k_offset_elems,
                                 int64_t v_offset_elems, const int* page_indices, const int* page_indptr,
                                 const int* last_page_len_d, const int* q_indptr, int num_qo_heads, int num_kv_heads,
                                 int head_dim, int page_size, int seq_len, int batch_size, int64_t stride_page,
                                 float sm_scale, pk_stream stream) {
  if (head_dim != PHD || !q_indptr || num_kv_heads <= 0 || num_qo_heads % num_kv_heads != 0 || page_size != 16) return -1;
  if (seq_len <= 0 || batch_size <= 0) return 0;
  const int impl = prefill_attn_impl() == 1 ? 1 : 2;
  const int rc = launch_prefill_tensor(impl, (const bf16*)q, (bf16*)output, (const bf16*)kv_data + k_offset_elems,
                                       (const bf16*)kv_data + v_offset_elems, page_indices, page_indptr, last_page_len_d,
                                       q_indptr, seq_len, batch_size, num_qo_heads, num_kv_heads, page_size, stride_page,
                                       sm_scale * 1.44269504088896340736f, stream);
  return rc == -2 ? -1 : rc;
}

}  // extern "C"
