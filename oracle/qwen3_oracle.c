/*
 * qwen3_oracle.c -- CPU restatement of pegainfer's Qwen3 forward-pass hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in pegainfer_b200/ (the product) may link,
 * import or execute this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the
 * checker / reported baseline.
 *
 * Every function restates one reference op with the SAME rounding points as
 * the reference CUDA code (cited per function, paths relative to the
 * reference repo root).  All tensors are bf16 bit patterns (uint16_t) unless
 * noted; all accumulation is fp32 (FMA-contracted, like the GPU).  Reduction
 * ORDER is the oracle's own (the reference's cuBLAS / FlashInfer orders are
 * unspecified), so parity against CUDA is stated with a tolerance of a few
 * bf16 ulps per op, not bit equality -- see DESIGN.md "Parity".
 *
 * Pinning: tests/test_oracle_golden.py checks this file against every
 * weight-free known-answer test the reference holds for the path
 * (pegainfer-server/src/ops/tests.rs, pegainfer-kernels/src/ops/embedding.rs,
 * pegainfer-core/src/kv_pool.rs) and against HF transformers Qwen3 fixtures
 * (tests/golden/, generator committed); tests/test_ref_kernels_gpu.py pins it
 * against the reference's own CUDA kernels (oracle/_ref) on the GPU box.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---- OpenMP control (bench.py's CPU arm states and pins its thread count; torchrun exports
 * OMP_NUM_THREADS=1 to every rank, which must not silently turn the "all host cores" baseline into a
 * single-threaded one) ----------------------------------------------------- */
#ifdef _OPENMP
#include <omp.h>
ORC_API void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
ORC_API int orc_get_max_threads(void) { return omp_get_max_threads(); }
#else
ORC_API void orc_set_num_threads(int n) { (void)n; }
ORC_API int orc_get_max_threads(void) { return 1; }
#endif

/* ---- bf16 helpers: __float2bfloat16 is round-to-nearest-even ------------ */
static inline float bf2f(uint16_t h) {
  uint32_t u = ((uint32_t)h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fff; /* NaN, as cuda_bf16.h */
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

ORC_API void orc_f32_to_bf16(const float* in, uint16_t* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = f2bf(in[i]);
}
ORC_API void orc_bf16_to_f32(const uint16_t* in, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = bf2f(in[i]);
}

/* Row-parallel copy with the SAME static row partition orc_gemm uses: the destination's pages are first touched by
 * the thread that will stream those rows, so on a multi-socket host every thread reads local memory (bench.py's CPU
 * arm re-homes the weight matrices with this; timing hygiene only, no arithmetic). */
ORC_API void orc_spread_rows(uint16_t* dst, const uint16_t* src, int64_t rows, int64_t cols) {
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < rows; ++m) memcpy(dst + m * cols, src + m * cols, (size_t)cols * 2);
}

/* fp32 dot product, 16 independent partial sums (vectorisable), fixed tree. */
static inline float dot_f32(const float* a, const float* b, int n) {
  float acc[16];
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  int k = 0;
  for (; k + 16 <= n; k += 16)
    for (int j = 0; j < 16; ++j) acc[j] = fmaf(a[k + j], b[k + j], acc[j]);
  for (int j = 0; k < n; ++k, ++j) acc[j] = fmaf(a[k], b[k], acc[j]);
  for (int s = 8; s > 0; s >>= 1)
    for (int j = 0; j < s; ++j) acc[j] += acc[j + s];
  return acc[0];
}

/* ---- a1: embedding  (csrc/elementwise.cu:49-112) ------------------------- */
ORC_API void orc_embedding_batched(const uint16_t* embed, const uint32_t* ids, uint16_t* out,
                                   int hidden, int seq_len) {
  for (int t = 0; t < seq_len; ++t)
    memcpy(out + (size_t)t * hidden, embed + (size_t)ids[t] * hidden, (size_t)hidden * 2);
}
ORC_API void orc_embedding_batched_vocab_shard(const uint16_t* embed, const uint32_t* ids,
                                               uint16_t* out, int hidden, int seq_len,
                                               uint32_t vocab_start, uint32_t part_vocab) {
  for (int t = 0; t < seq_len; ++t) {
    uint32_t id = ids[t];
    if (id >= vocab_start && id < vocab_start + part_vocab)
      memcpy(out + (size_t)t * hidden, embed + (size_t)(id - vocab_start) * hidden,
             (size_t)hidden * 2);
    else
      memset(out + (size_t)t * hidden, 0, (size_t)hidden * 2);
  }
}

/* ---- a2: RMSNorm (csrc/flashinfer_norm.cu:49-65 -> flashinfer/norm.cuh:36-111)
 * fp32 sum of squares, rsqrt(sum/d + eps), x*rms*w in fp32, ONE bf16 rounding. */
ORC_API void orc_rms_norm_batched(const uint16_t* x, const uint16_t* w, uint16_t* out, int hidden,
                                  int seq_len, float eps) {
#pragma omp parallel for schedule(static)
  for (int t = 0; t < seq_len; ++t) {
    const uint16_t* xr = x + (size_t)t * hidden;
    float ss = 0.f;
    for (int i = 0; i < hidden; ++i) {
      float v = bf2f(xr[i]);
      ss = fmaf(v, v, ss);
    }
    float r = 1.0f / sqrtf(ss / (float)hidden + eps);
    for (int i = 0; i < hidden; ++i)
      out[(size_t)t * hidden + i] = f2bf(bf2f(xr[i]) * r * (0.f + bf2f(w[i])));
  }
}

/* ---- a9: fused add + RMSNorm (flashinfer_norm.cu:71-105 -> norm.cuh:386-477)
 * x = f32(hidden)+f32(residual); hidden = bf16(x); sum of squares on UNROUNDED
 * x; out = bf16(x*rms*w). */
ORC_API void orc_fused_add_rms_norm_batched(uint16_t* hidden, const uint16_t* residual,
                                            const uint16_t* w, uint16_t* out, int hdim, int bs,
                                            float eps) {
#pragma omp parallel for schedule(static)
  for (int t = 0; t < bs; ++t) {
    float* xs = (float*)malloc(sizeof(float) * (size_t)hdim);
    float ss = 0.f;
    for (int i = 0; i < hdim; ++i) {
      /* FlashInfer is called with (input=out(=residual copy), residual=hidden):
       * x = float(input) ; x += float(residual)  -> residual + hidden */
      float xv = bf2f(residual[(size_t)t * hdim + i]);
      xv += bf2f(hidden[(size_t)t * hdim + i]);
      ss = fmaf(xv, xv, ss);
      hidden[(size_t)t * hdim + i] = f2bf(xv);
      xs[i] = xv;
    }
    float r = 1.0f / sqrtf(ss / (float)hdim + eps);
    for (int i = 0; i < hdim; ++i) out[(size_t)t * hdim + i] = f2bf(xs[i] * r * (0.f + bf2f(w[i])));
    free(xs);
  }
}

/* ---- a11: residual add (csrc/elementwise.cu:8-20) ------------------------ */
ORC_API void orc_add(const uint16_t* a, const uint16_t* b, uint16_t* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = f2bf(bf2f(a[i]) + bf2f(b[i]));
}

/* ---- a10: SiLU*mul (csrc/fused_proj.cu:44-63): g/(1+expf(-g))*u, one rounding */
ORC_API void orc_silu_mul_fused(const uint16_t* gate_up, uint16_t* out, int inter, int bs) {
  for (int t = 0; t < bs; ++t)
    for (int i = 0; i < inter; ++i) {
      float g = bf2f(gate_up[(size_t)t * 2 * inter + i]);
      float u = bf2f(gate_up[(size_t)t * 2 * inter + inter + i]);
      float s = g / (1.0f + expf(-g));
      out[(size_t)t * inter + i] = f2bf(s * u);
    }
}
/* unfused variant rounds SiLU to bf16 first (csrc/elementwise.cu:27-42; Qwen3.5 only) */
ORC_API void orc_silu_mul(const uint16_t* gate, const uint16_t* up, uint16_t* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    float g = bf2f(gate[i]);
    float s = g / (1.0f + expf(-g));
    out[i] = f2bf(bf2f(f2bf(s)) * bf2f(up[i]));
  }
}

/* ---- a3: GEMM / GEMV (csrc/linear.cu:48-78; cublasGemmEx bf16 x bf16 ->
 * COMPUTE_32F, bf16 out).  Y[n*M+m] = sum_k W[m*K+k] * X[n*K+k]. */
ORC_API void orc_gemm(const uint16_t* W, const uint16_t* X, uint16_t* Y, int M, int N, int K) {
  float* xf = (float*)malloc(sizeof(float) * (size_t)N * K);
  for (size_t i = 0; i < (size_t)N * K; ++i) xf[i] = bf2f(X[i]);
#pragma omp parallel
  {
    float* wr = (float*)malloc(sizeof(float) * (size_t)K);
#pragma omp for schedule(static)
    for (int m = 0; m < M; ++m) {
      const uint16_t* wrow = W + (size_t)m * K;
      for (int k = 0; k < K; ++k) wr[k] = bf2f(wrow[k]);
      for (int n = 0; n < N; ++n) Y[(size_t)n * M + m] = f2bf(dot_f32(wr, xf + (size_t)n * K, K));
    }
    free(wr);
  }
  free(xf);
}

/* ---- RoPE tables (pegainfer-core/src/weight_loader.rs:210-244) ----------- */
ORC_API void orc_precompute_rope(uint16_t* cos_out, uint16_t* sin_out, int head_dim, int max_pos,
                                 float theta) {
  int half = head_dim / 2;
  for (int pos = 0; pos < max_pos; ++pos)
    for (int i = 0; i < half; ++i) {
      float inv = 1.0f / powf(theta, (float)i * 2.0f / (float)head_dim);
      float fr = (float)pos * inv;
      uint16_t c = f2bf(cosf(fr)), s = f2bf(sinf(fr));
      cos_out[(size_t)pos * head_dim + i] = c;
      cos_out[(size_t)pos * head_dim + i + half] = c;
      sin_out[(size_t)pos * head_dim + i] = s;
      sin_out[(size_t)pos * head_dim + i + half] = s;
    }
}

/* ---- a4: per-head QK RMSNorm + NeoX RoPE (csrc/prefill_attention.cu:12-88)
 * normed=bf16(x*inv_rms); bf16(normed*w); rope in fp32 from bf16 cos/sin; bf16.
 * positions != NULL: pos = positions[t] (decode / multi-request), else start_pos+t. */
ORC_API void orc_qk_norm_rope(uint16_t* q, uint16_t* k, const uint16_t* qw, const uint16_t* kw,
                              const uint16_t* cosc, const uint16_t* sinc, const int* positions,
                              int start_pos, int nq, int nkv, int hd, int seq_len, float eps) {
  int half = hd / 2;
#pragma omp parallel for schedule(static)
  for (int t = 0; t < seq_len; ++t) {
    float nrm[512];
    int pos = positions ? positions[t] : start_pos + t;
    for (int h = 0; h < nq + nkv; ++h) {
      int isq = h < nq;
      uint16_t* d = isq ? q + (size_t)t * nq * hd + (size_t)h * hd
                        : k + (size_t)t * nkv * hd + (size_t)(h - nq) * hd;
      const uint16_t* w = isq ? qw : kw;
      float ss = 0.f;
      for (int i = 0; i < hd; ++i) {
        float v = bf2f(d[i]);
        ss += v * v;
      }
      float r = 1.0f / sqrtf(ss / (float)hd + eps);
      for (int i = 0; i < hd; ++i) {
        float n1 = bf2f(f2bf(bf2f(d[i]) * r));
        nrm[i] = bf2f(f2bf(n1 * bf2f(w[i])));
      }
      for (int i = 0; i < half; ++i) {
        float lo = nrm[i], hi = nrm[i + half];
        float c = bf2f(cosc[(size_t)pos * hd + i]), s = bf2f(sinc[(size_t)pos * hd + i]);
        d[i] = f2bf(lo * c - hi * s);
        d[i + half] = f2bf(lo * s + hi * c);
      }
    }
  }
}

/* ---- paged KV addressing (pegainfer-core/src/kv_pool.rs:14-75,
 * csrc/paged_attention.cu:37-66): page-first pool, NHD inside a block. */
static inline size_t kv_elem_off(const int* page_indices, const int* page_indptr, int b,
                                 int64_t tok, int head, int nkv, int hd, int page_size,
                                 int64_t stride_page) {
  int64_t pg = page_indices[page_indptr[b] + tok / page_size];
  int64_t slot = tok % page_size;
  return (size_t)(pg * stride_page + slot * nkv * hd + (int64_t)head * hd);
}
static inline int kv_len_of(const int* page_indptr, const int* last_page_len, int b, int page_size) {
  int np = page_indptr[b + 1] - page_indptr[b];
  return np <= 0 ? 0 : (np - 1) * page_size + last_page_len[b];
}

/* ---- a5: scatter/append (csrc/paged_attention.cu:274-311 -> page.cuh:259-284) */
ORC_API int orc_paged_kv_scatter(uint16_t* kv, int64_t k_off, int64_t v_off,
                                 const int* page_indices, const int* page_indptr,
                                 const int* last_page_len, const uint16_t* src_k,
                                 const uint16_t* src_v, const int* batch_indices,
                                 const int* positions, int nnz, int nkv, int hd, int page_size,
                                 int64_t stride_page, int64_t src_stride_n, int64_t src_stride_h) {
  (void)last_page_len;
  for (int i = 0; i < nnz; ++i)
    for (int h = 0; h < nkv; ++h) {
      size_t o = kv_elem_off(page_indices, page_indptr, batch_indices[i], positions[i], h, nkv, hd,
                             page_size, stride_page);
      memcpy(kv + k_off + o, src_k + (size_t)i * src_stride_n + (size_t)h * src_stride_h,
             (size_t)hd * 2);
      memcpy(kv + v_off + o, src_v + (size_t)i * src_stride_n + (size_t)h * src_stride_h,
             (size_t)hd * 2);
    }
  return 0;
}

/* one (request, q head, kv range) softmax-attention state in base-2 units
 * (flashinfer/attention/decode.cuh:62-145, state.cuh:30-80). */
static void attn_range(const float* qf, const uint16_t* kv, int64_t k_off, int64_t v_off,
                       const int* page_indices, const int* page_indptr, int b, int kvh, int nkv,
                       int hd, int page_size, int64_t stride_page, int lo, int hi, float scale_log2,
                       float* o, float* m_out, float* d_out) {
  int n = hi - lo;
  float* s = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  float kf[512];
  float m = -INFINITY;
  for (int j = 0; j < n; ++j) {
    size_t off = kv_elem_off(page_indices, page_indptr, b, lo + j, kvh, nkv, hd, page_size,
                             stride_page);
    for (int i = 0; i < hd; ++i) kf[i] = bf2f(kv[k_off + off + i]);
    s[j] = dot_f32(qf, kf, hd) * scale_log2;
    if (s[j] > m) m = s[j];
  }
  float d = 0.f;
  for (int i = 0; i < hd; ++i) o[i] = 0.f;
  for (int j = 0; j < n; ++j) {
    float p = exp2f(s[j] - m);
    d += p;
    size_t off = kv_elem_off(page_indices, page_indptr, b, lo + j, kvh, nkv, hd, page_size,
                             stride_page);
    for (int i = 0; i < hd; ++i) o[i] = fmaf(p, bf2f(kv[v_off + off + i]), o[i]);
  }
  free(s);
  *m_out = m;
  *d_out = d;
}

/* ---- a6: decode attention, non-partition (csrc/paged_attention.cu:77-145) */
ORC_API int orc_paged_attention_decode(const uint16_t* q, uint16_t* out, const uint16_t* kv,
                                       int64_t k_off, int64_t v_off, const int* page_indices,
                                       const int* page_indptr, const int* last_page_len,
                                       const int* request_indices, const int* kv_tile_indices,
                                       const int* kv_chunk_size, int nq, int nkv, int hd,
                                       int page_size, int bs, int64_t stride_page, float sm_scale) {
  (void)kv_tile_indices;
  (void)kv_chunk_size;
  int group = nq / nkv;
  float sl2 = sm_scale * 1.44269504088896340736f;
#pragma omp parallel for collapse(2) schedule(dynamic)
  for (int bx = 0; bx < bs; ++bx)
    for (int h = 0; h < nq; ++h) {
      int b = request_indices[bx];
      int len = kv_len_of(page_indptr, last_page_len, b, page_size);
      float qf[512], o[512], m, d;
      for (int i = 0; i < hd; ++i) qf[i] = bf2f(q[((size_t)b * nq + h) * hd + i]);
      attn_range(qf, kv, k_off, v_off, page_indices, page_indptr, b, h / group, nkv, hd, page_size,
                 stride_page, 0, len, sl2, o, &m, &d);
      for (int i = 0; i < hd; ++i) out[((size_t)bx * nq + h) * hd + i] = f2bf(o[i] / d);
    }
  return 0;
}

/* ---- a7: decode attention, split-KV (csrc/paged_attention.cu:158-230 +
 * cascade.cuh VariableLengthMergeStates): per-chunk NORMALISED partial rounded
 * to bf16 in tmp_v, base-2 LSE in tmp_s, fp32 merge, bf16 out. */
ORC_API int orc_paged_attention_decode_split_kv(
    const uint16_t* q, uint16_t* out, const uint16_t* kv, int64_t k_off, int64_t v_off,
    const int* page_indices, const int* page_indptr, const int* last_page_len,
    const int* request_indices, const int* kv_tile_indices, const int* kv_chunk_size,
    const int* o_indptr, const uint8_t* block_valid_mask, uint16_t* tmp_v, float* tmp_s, int nq,
    int nkv, int hd, int page_size, int bs, int padded_slots, int64_t stride_page, float sm_scale) {
  int group = nq / nkv;
  int chunk = kv_chunk_size[0];
  float sl2 = sm_scale * 1.44269504088896340736f;
#pragma omp parallel for collapse(2) schedule(dynamic)
  for (int slot = 0; slot < padded_slots; ++slot)
    for (int h = 0; h < nq; ++h) {
      if (block_valid_mask && !block_valid_mask[slot]) continue;
      int b = request_indices[slot];
      int tile = kv_tile_indices[slot];
      int len = kv_len_of(page_indptr, last_page_len, b, page_size);
      int lo = tile * chunk, hi = (tile + 1) * chunk;
      if (hi > len) hi = len;
      float qf[512], o[512], m, d;
      for (int i = 0; i < hd; ++i) qf[i] = bf2f(q[((size_t)b * nq + h) * hd + i]);
      attn_range(qf, kv, k_off, v_off, page_indices, page_indptr, b, h / group, nkv, hd, page_size,
                 stride_page, lo, hi, sl2, o, &m, &d);
      for (int i = 0; i < hd; ++i) tmp_v[((size_t)slot * nq + h) * hd + i] = f2bf(o[i] / d);
      tmp_s[(size_t)slot * nq + h] = m + log2f(d);
    }
  for (int b = 0; b < bs; ++b)
    for (int h = 0; h < nq; ++h) {
      /* state_t::init + merge (state.cuh:37-66) with other_d = 1 */
      float m = -INFINITY, d = 1.f, o[512];
      for (int i = 0; i < hd; ++i) o[i] = 0.f;
      for (int s = o_indptr[b]; s < o_indptr[b + 1]; ++s) {
        float om = tmp_s[(size_t)s * nq + h];
        float mn = om > m ? om : m;
        float a = exp2f(m - mn), c = exp2f(om - mn);
        d = d * a + c;
        for (int i = 0; i < hd; ++i)
          o[i] = o[i] * a + bf2f(tmp_v[((size_t)s * nq + h) * hd + i]) * c;
        m = mn;
      }
      for (int i = 0; i < hd; ++i) out[((size_t)b * nq + h) * hd + i] = f2bf(o[i] / d);
    }
  return 0;
}

/* ---- a8: causal GQA prefill attention over paged KV
 * (csrc/paged_attention.cu:399-500 -> flashinfer/attention/prefill.cuh FA2):
 * S fp32 from bf16 products; p = bf16(exp2(s*c - m*c)); denominator = sum of
 * the bf16-rounded p (prefill.cuh:956-985 rowsum on the f16 fragment);
 * O = sum p*v in fp32; one bf16 rounding of O/d.  The oracle uses the final
 * row max for m (FA2 uses the running max; same relative rounding). */
ORC_API int orc_batch_prefill_paged(const uint16_t* q, uint16_t* out, const uint16_t* kv,
                                    int64_t k_off, int64_t v_off, const int* page_indices,
                                    const int* page_indptr, const int* last_page_len,
                                    const int* q_indptr, int nq, int nkv, int hd, int page_size,
                                    int batch_size, int64_t stride_page, float sm_scale) {
  int group = nq / nkv;
  float sl2 = sm_scale * 1.44269504088896340736f;
  for (int b = 0; b < batch_size; ++b) {
    int qo_len = q_indptr[b + 1] - q_indptr[b];
    int kv_len = kv_len_of(page_indptr, last_page_len, b, page_size);
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int t = 0; t < qo_len; ++t)
      for (int h = 0; h < nq; ++h) {
        int n = kv_len - qo_len + t + 1; /* causal: kv_idx <= t + kv_len - qo_len */
        size_t row = (size_t)(q_indptr[b] + t) * nq + h;
        float qf[512], kf[512], o[512];
        for (int i = 0; i < hd; ++i) qf[i] = bf2f(q[row * hd + i]);
        float* s = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        float m = -INFINITY;
        for (int j = 0; j < n; ++j) {
          size_t off = kv_elem_off(page_indices, page_indptr, b, j, h / group, nkv, hd, page_size,
                                   stride_page);
          for (int i = 0; i < hd; ++i) kf[i] = bf2f(kv[k_off + off + i]);
          s[j] = dot_f32(qf, kf, hd);
          if (s[j] > m) m = s[j];
        }
        float d = 0.f;
        for (int i = 0; i < hd; ++i) o[i] = 0.f;
        for (int j = 0; j < n; ++j) {
          float p = bf2f(f2bf(exp2f(s[j] * sl2 - m * sl2)));
          d += p;
          size_t off = kv_elem_off(page_indices, page_indptr, b, j, h / group, nkv, hd, page_size,
                                   stride_page);
          for (int i = 0; i < hd; ++i) o[i] = fmaf(p, bf2f(kv[v_off + off + i]), o[i]);
        }
        free(s);
        for (int i = 0; i < hd; ++i) out[row * hd + i] = f2bf(o[i] / d);
      }
  }
  return 0;
}

/* ---- a12: arg-max over bf16 logits; lowest index wins ties
 * (csrc/argmax.cu:5-49; flashinfer_top1 tie order is not index-defined). */
ORC_API void orc_argmax(const uint16_t* x, int* out, int n) {
  float best = -INFINITY;
  int bi = 0;
  for (int i = 0; i < n; ++i) {
    float v = bf2f(x[i]);
    if (v > best) {
      best = v;
      bi = i;
    }
  }
  out[0] = bi;
}

/* ---- TP all-reduce model (pegainfer-qwen3-4b/src/weights.rs:396-405): SUM
 * over ranks of bf16 partials.  NCCL's order is unspecified ("parity
 * unpinned"); the oracle sums in fp32 in rank order and rounds once. */
ORC_API void orc_all_reduce_sum(const uint16_t* const* parts, int world, uint16_t* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    float a = 0.f;
    for (int r = 0; r < world; ++r) a += bf2f(parts[r][i]);
    out[i] = f2bf(a);
  }
}
