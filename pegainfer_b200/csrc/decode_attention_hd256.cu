// Paged GQA decode attention for head_dim 256 (Qwen3.5 full-attention layers): `paged_attention_decode_cuda_hd256`
// (ffi.rs:1286-1307; the reference instantiates FlashInfer BatchDecodeWithPagedKVCache for HD 256,
// csrc/paged_attention.cu:682-760).  Same design as decode_attention_cluster.cu, re-derived for 512-byte rows:
//   * one 8-CTA thread-block cluster per (request, kv head), CTA r owns an eighth of the context
//   * one WARP per token row (32 lanes x 16 B = one 256-element K row), the GQA group's 4 q heads in registers so
//     K and V are read once per group; 4 rows per warp in flight, the next round requested before this one is reduced
//   * CTA states merged into the leader CTA over distributed shared memory, no global synchronisation
// q arrives normed + roped (qk_norm_partial_rope_batched_decode_hd256_cuda), the step's K/V row is already in the
// pool: nothing is requested before griddepcontrol.wait.  fp32 throughout, one bf16 rounding of O / d.
#include "common.cuh"

namespace pk {

constexpr int H2 = 256;
constexpr int H2_WARPS = 8, H2_THREADS = H2_WARPS * 32;
constexpr int H2_U = 4;                      // token rows per warp per round
constexpr int H2_ROUND = H2_WARPS * H2_U;    // 32 tokens per CTA round
constexpr int H2_CLUSTER = 8, H2_GROUP = 4;

struct Hd256Args {
  const bf16* q;
  bf16* out;
  const bf16* kv;
  int64_t k_off, v_off, stride_page;
  const int *page_indices, *page_indptr, *last_page_len, *request_indices;
  float sm_scale_log2;
  int nq, nkv;
};

__device__ __forceinline__ float h2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void h2_unpack8(const uint4& a, float* f) {
  f[0] = bf16_lo(a.x); f[1] = bf16_hi(a.x); f[2] = bf16_lo(a.y); f[3] = bf16_hi(a.y);
  f[4] = bf16_lo(a.z); f[5] = bf16_hi(a.z); f[6] = bf16_lo(a.w); f[7] = bf16_hi(a.w);
}
__device__ __forceinline__ uint32_t h2_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t h2_map_to_rank(const void* smem_ptr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(rank));
  return r;
}
__device__ __forceinline__ void h2_st_cluster(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// split cluster barrier: arrive at kernel entry, wait right before the first remote shared-memory store, so every
// CTA of the cluster is known to have started (its shared memory exists) before anyone writes into it
__device__ __forceinline__ void h2_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void h2_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void h2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(H2_CLUSTER, 1, 1) __launch_bounds__(H2_THREADS, 1)
decode_attention_hd256_kernel(const Hd256Args a) {
  // dynamic shared memory (66 KB): per-warp states of this CTA, then the leader-side landing zone for the cluster
  extern __shared__ __align__(16) uint8_t h2_smem[];
  typedef float (*OArr)[H2_GROUP][H2];
  typedef float (*SArr)[H2_GROUP];
  OArr st_o = reinterpret_cast<OArr>(h2_smem);                                            // [H2_WARPS]
  OArr c_o = reinterpret_cast<OArr>(h2_smem + sizeof(float) * H2_WARPS * H2_GROUP * H2);  // [H2_CLUSTER]
  SArr st_m = reinterpret_cast<SArr>(h2_smem + sizeof(float) * (H2_WARPS + H2_CLUSTER) * H2_GROUP * H2);
  SArr st_d = st_m + H2_WARPS;
  SArr c_m = st_d + H2_WARPS;
  SArr c_d = c_m + H2_CLUSTER;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)h2_cluster_rank();
  const int kvh = blockIdx.y;
  const int b = a.request_indices ? a.request_indices[blockIdx.z] : (int)blockIdx.z;
  const int npages = a.page_indptr[b + 1] - a.page_indptr[b];
  const int len = npages <= 0 ? 0 : (npages - 1) * 16 + a.last_page_len[b];
  int chunk = (len + H2_CLUSTER - 1) / H2_CLUSTER;
  chunk = (chunk + H2_ROUND - 1) / H2_ROUND * H2_ROUND;
  const int lo = min(len, rank * chunk), hi = min(len, lo + chunk);
  const int* pages = a.page_indices + a.page_indptr[b];
  const bf16* kbase = a.kv + a.k_off + (int64_t)kvh * H2 + lane * 8;
  const bf16* vbase = a.kv + a.v_off + (int64_t)kvh * H2 + lane * 8;

  h2_cluster_arrive();  // opening barrier (completed by h2_cluster_wait before the DSMEM stores)
  pdl_launch_dependents();
  pdl_wait();

  // the group's 4 query heads: lane holds its 8 dims of each
  float qf[H2_GROUP][8];
#pragma unroll
  for (int h = 0; h < H2_GROUP; ++h) {
    const uint4 qv = reinterpret_cast<const uint4*>(a.q + ((size_t)b * a.nq + kvh * H2_GROUP + h) * H2)[lane];
    h2_unpack8(qv, qf[h]);
  }
  uint4 kr[H2_U], vr[H2_U];
  bool ok[H2_U];
  auto load_round = [&](int round, uint4* kk, uint4* vv, bool* okk) {
#pragma unroll
    for (int u = 0; u < H2_U; ++u) {
      const int t = round + warp + H2_WARPS * u;  // warp-uniform
      okk[u] = t < hi;
      kk[u] = make_uint4(0, 0, 0, 0);
      vv[u] = make_uint4(0, 0, 0, 0);
      if (okk[u]) {
        const int page = __ldg(pages + (t >> 4));
        const int64_t off = (int64_t)page * a.stride_page + (int64_t)(t & 15) * a.nkv * H2;
        kk[u] = ldg_stream(kbase + off);
        vv[u] = ldg_stream(vbase + off);
      }
    }
  };
  load_round(lo, kr, vr, ok);

  float m[H2_GROUP], d[H2_GROUP], o[H2_GROUP][8];
#pragma unroll
  for (int h = 0; h < H2_GROUP; ++h) {
    m[h] = -INFINITY;
    d[h] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[h][j] = 0.f;
  }
  for (int round = lo; round < hi; round += H2_ROUND) {  // CTA-uniform trip count
    uint4 krn[H2_U], vrn[H2_U];
    bool okn[H2_U];
    const bool more = round + H2_ROUND < hi;
    if (more) load_round(round + H2_ROUND, krn, vrn, okn);
    float s[H2_GROUP][H2_U];
#pragma unroll
    for (int u = 0; u < H2_U; ++u) {
      float kf[8];
      h2_unpack8(kr[u], kf);
#pragma unroll
      for (int h = 0; h < H2_GROUP; ++h) {
        float p = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) p = fmaf(qf[h][j], kf[j], p);
        p = warp_sum(p);
        s[h][u] = ok[u] ? p * a.sm_scale_log2 : -INFINITY;
      }
    }
#pragma unroll
    for (int h = 0; h < H2_GROUP; ++h) {
      float mn = m[h];
#pragma unroll
      for (int u = 0; u < H2_U; ++u) mn = fmaxf(mn, s[h][u]);
      if (mn == -INFINITY) {  // nothing valid yet: keep the state empty, contribute zero weights
#pragma unroll
        for (int u = 0; u < H2_U; ++u) s[h][u] = 0.f;
        continue;
      }
      const float sc = h2_ex2(m[h] - mn);
      m[h] = mn;
      d[h] *= sc;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[h][j] *= sc;
#pragma unroll
      for (int u = 0; u < H2_U; ++u) {
        s[h][u] = h2_ex2(s[h][u] - mn);
        d[h] += s[h][u];
      }
    }
#pragma unroll
    for (int u = 0; u < H2_U; ++u) {
      float vf[8];
      h2_unpack8(vr[u], vf);
#pragma unroll
      for (int h = 0; h < H2_GROUP; ++h)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[h][j] = fmaf(s[h][u], vf[j], o[h][j]);
    }
    if (more) {
#pragma unroll
      for (int u = 0; u < H2_U; ++u) {
        kr[u] = krn[u];
        vr[u] = vrn[u];
        ok[u] = okn[u];
      }
    }
  }

  // ---- merge the 8 warps of this CTA through shared memory ----
#pragma unroll
  for (int h = 0; h < H2_GROUP; ++h) {
    if (lane == 0) {
      st_m[warp][h] = m[h];
      st_d[warp][h] = d[h];
    }
    float4* dst = reinterpret_cast<float4*>(&st_o[warp][h][lane * 8]);
    dst[0] = make_float4(o[h][0], o[h][1], o[h][2], o[h][3]);
    dst[1] = make_float4(o[h][4], o[h][5], o[h][6], o[h][7]);
  }
  __syncthreads();
  h2_cluster_wait();
  const int t = threadIdx.x;  // 256 threads = 256 output dims
#pragma unroll
  for (int h = 0; h < H2_GROUP; ++h) {
    float mx = -INFINITY;
#pragma unroll
    for (int w = 0; w < H2_WARPS; ++w) mx = fmaxf(mx, st_m[w][h]);
    float dd = 0.f, oo = 0.f;
#pragma unroll
    for (int w = 0; w < H2_WARPS; ++w) {
      const float wt = st_m[w][h] == -INFINITY ? 0.f : h2_ex2(st_m[w][h] - mx);
      dd = fmaf(st_d[w][h], wt, dd);
      oo = fmaf(st_o[w][h][t], wt, oo);
    }
    // (max, denominator, numerator) of this CTA -> the leader's landing zone over DSMEM
    h2_st_cluster(h2_map_to_rank(&c_o[rank][h][t], 0), oo);
    if (t == 0) {
      h2_st_cluster(h2_map_to_rank(&c_m[rank][h], 0), mx);
      h2_st_cluster(h2_map_to_rank(&c_d[rank][h], 0), dd);
    }
  }
  h2_cluster_sync();
  if (rank != 0) return;
#pragma unroll
  for (int h = 0; h < H2_GROUP; ++h) {
    float mx = -INFINITY;
#pragma unroll
    for (int r = 0; r < H2_CLUSTER; ++r) mx = fmaxf(mx, c_m[r][h]);
    float dd = 0.f, oo = 0.f;
#pragma unroll
    for (int r = 0; r < H2_CLUSTER; ++r) {
      const float wt = c_m[r][h] == -INFINITY ? 0.f : h2_ex2(c_m[r][h] - mx);
      dd = fmaf(c_d[r][h], wt, dd);
      oo = fmaf(c_o[r][h][t], wt, oo);
    }
    a.out[((size_t)b * a.nq + kvh * H2_GROUP + h) * H2 + t] = f2bf(dd > 0.f ? __fdividef(oo, dd) : 0.f);  // len == 0 -> zeros
  }
}

}  // namespace pk

using namespace pk;

extern "C" int paged_attention_decode_cuda_hd256(const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data, int64_t k_offset_elems,
                                                 int64_t v_offset_elems, const int* page_indices, const int* page_indptr,
                                                 const int* last_page_len_d, const int* request_indices, const int* kv_tile_indices,
                                                 const int* kv_chunk_size_ptr, int num_qo_heads, int num_kv_heads, int head_dim,
                                                 int page_size, int batch_size, int64_t stride_page, float sm_scale, pk_stream stream) {
  (void)kv_tile_indices;
  (void)kv_chunk_size_ptr;
  // the reference's only instantiation: HD 256, page 16, GQA group 4 (Qwen3.5-4B: 16 q heads / 4 kv heads)
  if (head_dim != H2 || page_size != 16 || num_kv_heads <= 0 || num_qo_heads != H2_GROUP * num_kv_heads) return -1;
  if (batch_size <= 0) return 0;
  Hd256Args a{};
  a.q = (const bf16*)q; a.out = (bf16*)output; a.kv = (const bf16*)kv_data;
  a.k_off = k_offset_elems; a.v_off = v_offset_elems; a.stride_page = stride_page;
  a.page_indices = page_indices; a.page_indptr = page_indptr; a.last_page_len = last_page_len_d;
  a.request_indices = request_indices;
  a.sm_scale_log2 = sm_scale * 1.44269504088896340736f;
  a.nq = num_qo_heads; a.nkv = num_kv_heads;
  constexpr size_t smem = sizeof(float) * ((H2_WARPS + H2_CLUSTER) * H2_GROUP * H2 + 2 * (H2_WARPS + H2_CLUSTER) * H2_GROUP);
  static thread_local bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(decode_attention_hd256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cfg = true;
  }
  return (int)launch(decode_attention_hd256_kernel, dim3(H2_CLUSTER, num_kv_heads, batch_size), dim3(H2_THREADS), smem, stream, true, a);
}
