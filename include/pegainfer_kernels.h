/*
 * pegainfer_kernels.h -- C ABI of libpegainfer_kernels_b200.so
 *
 * Drop-in boundary for the Qwen3 forward-pass hot path of xiaguan/pegainfer:
 * every entry point below has the NAME, ARGUMENT ORDER and ERROR CONVENTION of
 * the `unsafe extern "C"` declaration in the reference's
 * pegainfer-kernels/src/ffi.rs that it replaces (file:line cited per symbol), so
 * the reference's Rust `ops::*` wrappers bind to this library unchanged
 * (INTEGRATION.md shows the build.rs change).  Section "B200 extensions" adds the
 * fused / tensor-parallel entry points the reference does not have.
 *
 * Conventions (reference: SURVEY.md 8b):
 *  - pk_bf16 = uint16_t bf16 bit pattern (`pub type Half = u16`, ffi.rs:4).
 *  - All pointers are DEVICE pointers unless a comment says host.
 *  - Activations are `HiddenStates[dim, tokens]`: token t occupies elements
 *    [t*dim, (t+1)*dim) (pegainfer-kernels/src/tensor.rs:210-217); weights are
 *    row-major [out, in] (tensor.rs:136-141).
 *  - Buffers are owned by the caller.  The library owns only per-thread scratch
 *    created by cublas_init() and never allocates inside a launch, so every
 *    launch entry point is CUDA-graph-capture safe.
 *  - Error styles kept from the reference: `void` (errors surface at the next
 *    sync), `pk_curesult` (= cudaGetLastError cast), `int` (cudaError, 0 = ok,
 *    -1 = invalid argument).
 *  - There is NO CPU fallback: without a CUDA device every launch fails.
 */
#ifndef PEGAINFER_KERNELS_H_
#define PEGAINFER_KERNELS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t pk_bf16;
typedef struct CUstream_st* pk_stream; /* CUstream == cudaStream_t */
typedef int pk_curesult;               /* CUresult */

/* ---- context / handles: ffi.rs:159-161, csrc/linear.cu:14-42 ----------------
 * Call cuda_set_device + cublas_init once per rank thread.  cublas_init keeps its
 * historical name; here it creates the per-thread split-KV / top-1 scratch (the
 * reference allocates its 32 MB cuBLAS workspace at the same point). */
int  cuda_set_device(int device_ordinal);
void cublas_init(void);
void cublas_destroy(void);

/* ---- embedding: ffi.rs:78-96,143-149; csrc/elementwise.cu:49-112 ------------ */
pk_curesult embedding_batched_cuda(const pk_bf16* embed, const uint32_t* token_ids, pk_bf16* out,
                                   int hidden_size, int seq_len, pk_stream stream);
pk_curesult embedding_decode_cuda(const pk_bf16* embed, const uint32_t* token_id, pk_bf16* out,
                                  int hidden_size, pk_stream stream);
pk_curesult embedding_batched_vocab_shard_cuda(const pk_bf16* embed, const uint32_t* token_ids,
                                               pk_bf16* out, int hidden_size, int seq_len,
                                               uint32_t vocab_start, uint32_t part_vocab_size,
                                               pk_stream stream);

/* ---- norms: ffi.rs:22-68; csrc/flashinfer_norm.cu:49-105 --------------------
 * fused_add_rms_norm: hidden += residual (bf16-rounded store); out = RMSNorm of the
 * UNROUNDED fp32 sum.  No staging memcpy (the reference does one). */
void rms_norm_cuda(const pk_bf16* x, const pk_bf16* weight, pk_bf16* out, int n, float eps,
                   pk_stream stream);
void rms_norm_batched_cuda(const pk_bf16* x, const pk_bf16* weight, pk_bf16* out, int hidden_dim,
                           int seq_len, float eps, pk_stream stream);
void fused_add_rms_norm_cuda(pk_bf16* hidden, const pk_bf16* residual, const pk_bf16* weight,
                             pk_bf16* out, int n, float eps, pk_stream stream);
void fused_add_rms_norm_batched_cuda(pk_bf16* hidden, const pk_bf16* residual,
                                     const pk_bf16* weight, pk_bf16* out, int hidden_dim,
                                     int batch_size, float eps, pk_stream stream);

/* ---- elementwise: ffi.rs:41-47,70-76,151-157 -------------------------------- */
pk_curesult add_cuda(const pk_bf16* a, const pk_bf16* b, pk_bf16* out, int n, pk_stream stream);
pk_curesult silu_mul_triton_aot_cuda(const pk_bf16* gate, const pk_bf16* up, pk_bf16* out, int n,
                                     pk_stream stream); /* historical name; rounds SiLU to bf16 */
void silu_mul_fused_cuda(const pk_bf16* gate_up, pk_bf16* out, int intermediate_size, int bs,
                         pk_stream stream);

/* ---- GEMM / GEMV: ffi.rs:122-140; csrc/linear.cu:48-78 ----------------------
 * Y[M,N] (col-major, ld=M) = W[M,K] (row-major) * X[K,N] (col-major, ld=K);
 * bf16 in, fp32 accumulate, one bf16 rounding.
 * Both names dispatch alike and are capture safe: N <= 4 runs the HBM-streaming GEMV,
 * larger N the tcgen05 tensor-core GEMM. */
void gemm_cuda(const pk_bf16* W, const pk_bf16* X, pk_bf16* Y, int M, int N, int K,
               pk_stream stream);
void gemm_graphsafe_cuda(const pk_bf16* W, const pk_bf16* X, pk_bf16* Y, int M, int N, int K,
                         pk_stream stream);

/* ---- QK-norm + RoPE: ffi.rs:164-178,1143-1157; csrc/prefill_attention.cu:12-159
 * In place on q [nq*hd, T] and k [nkv*hd, T]; head_dim must be 128 (as the reference). */
void prefill_qk_norm_rope_only_cuda(pk_bf16* q_batch, pk_bf16* k_batch,
                                    const pk_bf16* q_norm_weight, const pk_bf16* k_norm_weight,
                                    const pk_bf16* cos_cache, const pk_bf16* sin_cache,
                                    int num_q_heads, int num_kv_heads, int head_dim, int seq_len,
                                    int start_pos, float rms_eps, pk_stream stream);
void qk_norm_rope_batched_decode_cuda(pk_bf16* q, pk_bf16* k, const pk_bf16* q_norm_weight,
                                      const pk_bf16* k_norm_weight, const pk_bf16* cos_cache,
                                      const pk_bf16* sin_cache, const int* positions,
                                      int num_q_heads, int num_kv_heads, int head_dim,
                                      int batch_size, float rms_eps, pk_stream stream);

/* ---- paged KV + attention: ffi.rs:1160-1283,1337-1385; csrc/paged_attention.cu */
int paged_kv_scatter_cuda(const pk_bf16* kv_data, int64_t k_offset_elems, int64_t v_offset_elems,
                          const int* page_indices, const int* page_indptr,
                          const int* last_page_len_d, const pk_bf16* src_k, const pk_bf16* src_v,
                          const int* batch_indices, const int* positions, int nnz,
                          int num_kv_heads, int head_dim, int page_size, int64_t stride_page,
                          int64_t src_stride_n, int64_t src_stride_h, pk_stream stream);

/* host-only planning helpers (csrc/paged_attention.cu:343-397) */
int batch_prefill_paged_num_tiles(int seq_len, int num_qo_heads, int num_kv_heads, int head_dim);
int batch_prefill_paged_num_tiles_with_cta_tile_q(int seq_len, int num_qo_heads, int num_kv_heads,
                                                  int head_dim, int cta_tile_q_override);
int batch_prefill_cta_tile_q(int total_seq_len, int num_qo_heads, int num_kv_heads, int head_dim);
int batch_prefill_cta_tile_q_with_override(int total_seq_len, int num_qo_heads, int num_kv_heads,
                                           int head_dim, int cta_tile_q_override);

int batch_prefill_paged_cuda(const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data,
                             int64_t k_offset_elems, int64_t v_offset_elems,
                             const int* page_indices, const int* page_indptr,
                             const int* last_page_len_d, const int* q_indptr,
                             const int* request_indices, const int* qo_tile_indices,
                             const int* kv_tile_indices, const int* kv_chunk_size_ptr,
                             const uint32_t* total_num_rows, int num_qo_heads, int num_kv_heads,
                             int head_dim, int page_size, int seq_len, int batch_size,
                             int padded_batch_size, int64_t stride_page, float sm_scale,
                             pk_stream stream);
int batch_prefill_paged_cuda_with_cta_tile_q(
    const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data, int64_t k_offset_elems,
    int64_t v_offset_elems, const int* page_indices, const int* page_indptr,
    const int* last_page_len_d, const int* q_indptr, const int* request_indices,
    const int* qo_tile_indices, const int* kv_tile_indices, const int* kv_chunk_size_ptr,
    const uint32_t* total_num_rows, int num_qo_heads, int num_kv_heads, int head_dim,
    int page_size, int seq_len, int batch_size, int padded_batch_size, int64_t stride_page,
    float sm_scale, int cta_tile_q_override, pk_stream stream);
int single_prefill_cuda(const pk_bf16* q, pk_bf16* output, const pk_bf16* k_cache,
                        const pk_bf16* v_cache, int num_qo_heads, int num_kv_heads, int head_dim,
                        int seq_len, int kv_len, int max_seq_len, float sm_scale,
                        pk_stream stream);

int paged_attention_decode_cuda(const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data,
                                int64_t k_offset_elems, int64_t v_offset_elems,
                                const int* page_indices, const int* page_indptr,
                                const int* last_page_len_d, const int* request_indices,
                                const int* kv_tile_indices, const int* kv_chunk_size_ptr,
                                int num_qo_heads, int num_kv_heads, int head_dim, int page_size,
                                int batch_size, int64_t stride_page, float sm_scale,
                                pk_stream stream);
int paged_attention_decode_split_kv_cuda(
    const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data, int64_t k_offset_elems,
    int64_t v_offset_elems, const int* page_indices, const int* page_indptr,
    const int* last_page_len_d, const int* request_indices, const int* kv_tile_indices,
    const int* kv_chunk_size_ptr, const int* o_indptr, const uint8_t* block_valid_mask,
    pk_bf16* tmp_v, float* tmp_s, int num_qo_heads, int num_kv_heads, int head_dim, int page_size,
    int batch_size, int padded_batch_size, int64_t stride_page, float sm_scale, pk_stream stream);

/* ---- sampling: ffi.rs:98-108; csrc/argmax.cu, csrc/flashinfer_top1.cu --------
 * Both pick the LOWEST index among equal maxima (the reference's radix top-1 leaves
 * tie order undefined).  row_states_scratch: >= 4 KiB (the reference passes 1 MiB).
 * flashinfer_top1_cuda also leaves the winner's bf16 VALUE in top1_value_scratch[0] on every path (the reference uses it
 * as scratch only): the vocab-sharded tensor-parallel greedy token compares it across ranks (pk_tp_top1_exchange). */
void argmax_cuda(const pk_bf16* x, int* out, int n, pk_stream stream);
void flashinfer_top1_cuda(const pk_bf16* logits, pk_bf16* top1_value_scratch,
                          uint8_t* row_states_scratch, int* output, int vocab_size,
                          pk_stream stream);

/* Non-greedy sampling: ffi.rs:110-120; csrc/flashinfer_sampling.cu:13-110.  probs = softmax(logits *
 * inv_temperature) (fp32, written to probs_scratch[vocab]); joint top-k / top-p filtering; multinomial draw
 * from the renormalised eligible set.  Distribution as the reference; the random stream is splitmix64(seed),
 * not FlashInfer's Philox.  valid_scratch[0] is set to 1. */
void gpu_sample_flashinfer_cuda(const pk_bf16* logits, float* probs_scratch, uint8_t* valid_scratch,
                                int* output, int vocab_size, float inv_temperature, int top_k, float top_p,
                                uint64_t seed, pk_stream stream);

/* =============================== B200 extensions ===============================
 * Not in ffi.rs.  The fused decode layer used by the host mirror
 * (pegainfer_b200/csrc/host) and the TP hook north_star asks pegainfer-comm to gain. */

/* Library / build identification ("pegainfer-kernels-b200 sm_100a ..."). */
const char* pk_b200_version(void);
/* Number of kernels launched by this library on the calling thread since the last
 * reset (bench.py's gpu_launches; counts launches recorded into a capture too). */
int64_t pk_b200_launch_count(int reset);
/* Enable (1) / disable (0) programmatic dependent launch on this thread's launches. */
void pk_b200_set_pdl(int enable);

/* Tensor-core GEMM with up to three output segments (the fused q|k|v projection of prefill):
 * rows [0, seg_rows[0]) of W -> Y[0] (a HiddenStates [seg_rows[0], N]), the next seg_rows[1] -> Y[1], the
 * rest -> Y[2].  seg_rows must sum to M.  Same arithmetic as gemm_cuda. */
int pk_b200_gemm_segments(const pk_bf16* W, const pk_bf16* X, pk_bf16* const* Y, const int* seg_rows, int M, int N,
                          int K, pk_stream stream);

/* gate_up projection + SwiGLU in one tensor-core launch (CTA-pair kernel): W = [gate (M rows); up (M rows)] x [K],
 * Y[tok][M] = bf16(silu(bf16(gate.x)) * bf16(up.x)) -- gemm_cuda + silu_mul_fused_cuda with the same rounding points.
 * Returns 0, or -2 when the shape is not for this kernel (run the two-kernel sequence instead). */
int pk_b200_gemm_swiglu(const pk_bf16* W, const pk_bf16* X, pk_bf16* Y, int M, int N, int K, pk_stream stream);
/* Which prefill GEMM kernel a [M features] x [N tokens] x [K] problem gets on a GPU with `sms` SMs (the choice gemm_cuda /
 * pk_b200_gemm_segments / pk_b200_gemm_swiglu make internally): 256 / 320 / 128 = tile width of the CTA-pair tcgen05 kernel
 * (gemm2.cu), -2 = the single-CTA kernel (gemm.cu, split-K when skinny).  Pure host arithmetic, no CUDA call: the wave-count
 * cost model calibrated on B200 measurements (profiles/README.md round 2) is pinned by a CPU test. */
int pk_b200_gemm_plan(int M, int N, int K, int swiglu, int sms);

/* GEMV with fused prologue/epilogue for decode (N == 1..4 tokens), one launch:
 *   x_mode 0: x = X as is ([N, K]).
 *   x_mode 1: X is `hidden` [N, K]; x = RMSNorm(hidden + residual) * norm_w computed in the
 *             prologue exactly as fused_add_rms_norm_batched_cuda would (unrounded fp32 sum,
 *             one rounding); CTA 0 also stores bf16(hidden + residual) to hidden_out (must NOT
 *             alias X: other CTAs still read it) and x to normed_out (optional).
 *   epi 0:    Y = bf16(W x); rows are routed to up to three outputs: rows [0, seg_rows[0]) ->
 *             Y[0], the next seg_rows[1] -> Y[1], the rest -> Y[2] (the fused q|k|v projection).
 *   epi 1:    SwiGLU: W = [M gate rows | M up rows];
 *             Y[0][m] = bf16(silu(bf16(gate_m.x)) * bf16(up_m.x))  (csrc/fused_proj.cu:44-63).
 *   x_mode 3: Qwen3.5 variant of x_mode 1: hidden_out = bf16(hidden + residual) and the norm runs on that ROUNDED sum with
 *             (1 + w) weights (add_batch + rms_norm_batch_offset of pegainfer-qwen35-4b/src/batch_decode.rs:241-253).
 *   epi 4:    like epi 1 with SiLU rounded to bf16 before the multiply (silu_mul_triton_aot_cuda, elementwise.cu:36-41).
 *   Tensor parallel, GEMV fused with its all-reduce over NVLink peer memory (no collective launch):
 *   epi 2:    the bf16 partial rows are pushed into every rank's staging slot and the grid's last CTA
 *             publishes the sequence flag (release.sys); Y is not written.
 *   x_mode 2: like x_mode 1, but `residual` is the all-reduce result: after all ranks' flags are seen the
 *             prologue sums the `world` partials from local staging in rank order (fp32, one bf16 rounding),
 *             adds the residual stream X and applies RMSNorm.  Pairs with the previous epi-2 launch.
 *   epi 3:    GEMV and its all-reduce in ONE kernel, row tile by row tile (the default for TP decode): every CTA
 *             rounds its partial rows to bf16, stores them as 8-byte {2 x bf16, seq} lines straight into every
 *             peer's staging area over NVLink (flag travels with the data: no fence, no flag store, no ticket),
 *             then polls ITS OWN rows' lines from the peers in local memory, sums in rank order (fp32, one bf16
 *             rounding -- bit-identical on every rank, and to the standalone collective) and writes the REDUCED
 *             rows to Y[0].  All ranks launch the same grid, so CTA c owns the same rows everywhere.  seq =
 *             (*tp_step) * 256 + tp_op + 1: `tp_step` is a device word the host advances once per decode step on
 *             every rank (part of the step's metadata block), `tp_op` the op's index inside the step (baked into
 *             the CUDA graph); the staging slot alternates with tp_op's parity. */
typedef struct {
  const pk_bf16* W;
  const pk_bf16* X;
  pk_bf16* Y[3];
  int seg_rows[3];
  int M, N, K;
  int x_mode;
  const pk_bf16* residual;
  const pk_bf16* norm_w;
  float eps;
  pk_bf16* hidden_out;
  pk_bf16* normed_out;
  int epi;
  void* tp_comm; /* pk_tp_comm*, required for x_mode 2 / epi 2 / epi 3 (tensor parallel) */
  const uint32_t* tp_step; /* epi 3: device word, the decode step counter (same value on every rank) */
  int tp_op;               /* epi 3: index of this collective inside the step (0..254) */
} pk_b200_gemv_args;
int pk_b200_gemv_fused(const pk_b200_gemv_args* args, pk_stream stream);
/* Ring depth (2..12 stages of 8 row segments), CTAs per SM and K elements per row segment (multiple of
 * 256) of the streaming GEMV; 0 keeps a value.  Defaults: PK_GEMV_STAGES / PK_GEMV_CTAS_PER_SM /
 * PK_GEMV_KC or the built-in tuning. */
void pk_b200_set_gemv_tuning(int stages, int ctas_per_sm, int segment_elems);

/* QK-norm + RoPE + KV append + split-KV GQA decode attention + merge in ONE launch.
 * q/k/v are the raw projections of the step ([dim, bs]); k is normed/roped and both k, v
 * are appended at `positions` before use.  partial_* : fp32 scratch
 * [bs * max_chunks * nq * (hd + 2)], counters: int[bs * nkv] zero-initialised (self-resetting). */
int pk_b200_decode_attention_fused(
    const pk_bf16* q, const pk_bf16* k, const pk_bf16* v, pk_bf16* output, pk_bf16* kv_data,
    int64_t k_offset_elems, int64_t v_offset_elems, const int* page_indices,
    const int* page_indptr, const int* last_page_len_d, const int* positions,
    const pk_bf16* q_norm_weight, const pk_bf16* k_norm_weight, const pk_bf16* cos_cache,
    const pk_bf16* sin_cache, float rms_eps, float* partial_scratch, int* counters,
    int chunk_tokens, int max_chunks, int num_qo_heads, int num_kv_heads, int head_dim,
    int page_size, int batch_size, int64_t stride_page, float sm_scale, pk_stream stream);

/* Causal GQA prefill attention over the paged cache on tcgen05/TMEM, callable directly: the kernels behind
 * batch_prefill_paged_cuda* (replaces csrc/prefill_attention.cu:24-386 / FlashInfer BatchPrefillWithPagedKVCache).
 * PK_PREFILL_ATTN (read per call) selects tc2 (default; prefill_attention_tc2.cu: two 128-token query tiles per CTA in
 * ping-pong, O accumulated in TMEM, P handed to the tensor core through TMEM, correction warpgroup), tc
 * (prefill_attention_tc.cu: one tile per CTA, O in registers) or, for the ABI entry only, legacy (mma.sync).
 * Same inputs as the ABI entry minus the FlashInfer tile plan (tiles are derived from q_indptr on the device).  K/V
 * pages are fetched with 4-D TMA tile loads, so the pool base + offsets must be 16-byte aligned.  Returns
 * 0 / cudaError / -1 for unsupported shapes (head_dim != 128, page_size != 16). */
int pk_b200_prefill_attention_tc(const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data, int64_t k_offset_elems,
                                 int64_t v_offset_elems, const int* page_indices, const int* page_indptr,
                                 const int* last_page_len_d, const int* q_indptr, int num_qo_heads, int num_kv_heads,
                                 int head_dim, int page_size, int seq_len, int batch_size, int64_t stride_page,
                                 float sm_scale, pk_stream stream);

/* Measurement aid of prefill_attention_tc2.cu: with PK_FA2_DBG=32 the heaviest CTA of head 0 stamps %clock64 at the
 * hand-over points of its tile 0; this copies the stamps ([role: softmax, MMA issuer, correction][block < 128][8] u64) to
 * the host (tools/fa2_trace.py, profiles/r2_prefill_attention.md).  The other PK_FA2_DBG bits (1 no P V MMAs, 2 no S
 * MMAs, 4 no exponentials, 8 no O rescale, 64 skip one TMEM read, 128 no tile stagger) are timing ablations and give
 * wrong results by construction; unset = the product path.  Returns cudaError. */
int pk_b200_fa2_trace_copy(void* host_out, int bytes);

/* pk_b200_decode_attention_fused plus an L2 prefetch of weights that FOLLOWING launches will stream.
 * bs-1 decode attention is latency-bound (a few MB of K/V against ~10 us of dependent steps) and leaves HBM
 * idle; extra clusters of the same launch use that window to pull weight rows into the 126 MB L2
 * (cp.async.bulk.prefetch.L2), after the previous kernel has drained.  A span describes one row-major weight
 * block and how the GEMV that will read it cuts it: `slices` = that GEMV's grid (pk_b200_gemv_grid), and the
 * first `prefetch_rows` rows of every slice are requested so all of its CTAs gain equally.  Results are
 * identical with or without spans.  At most 4 spans; row_bytes % 16 == 0, base 16-byte aligned (else -1). */
typedef struct pk_b200_prefetch_span {
  const void* base;      /* first row of the weight block */
  int32_t rows;          /* rows of the block (= the GEMV's M for that block) */
  int32_t row_bytes;     /* K * 2 */
  int32_t slices;        /* grid of the GEMV that will read it */
  int32_t prefetch_rows; /* leading rows of each slice to request */
} pk_b200_prefetch_span;
int pk_b200_decode_attention_fused_prefetch(
    const pk_bf16* q, const pk_bf16* k, const pk_bf16* v, pk_bf16* output, pk_bf16* kv_data,
    int64_t k_offset_elems, int64_t v_offset_elems, const int* page_indices,
    const int* page_indptr, const int* last_page_len_d, const int* positions,
    const pk_bf16* q_norm_weight, const pk_bf16* k_norm_weight, const pk_bf16* cos_cache,
    const pk_bf16* sin_cache, float rms_eps, float* partial_scratch, int* counters,
    int chunk_tokens, int max_chunks, int num_qo_heads, int num_kv_heads, int head_dim,
    int page_size, int batch_size, int64_t stride_page, float sm_scale,
    const pk_b200_prefetch_span* spans, int num_spans, pk_stream stream);
/* Grid (row slices) pk_b200_gemv_fused / gemm_cuda(N <= 4) uses for M output rows with epilogue `epi`. */
int pk_b200_gemv_grid(int M, int epi);

/* ---- Qwen3.5 hybrid-layer ops (next scope row; ffi.rs:181-226,981-1039) --------
 * Memory-bound pieces of the gated-delta-net / gated HD-256 attention layers behind the reference's names
 * (csrc/{flashinfer_norm,norm,conv1d,gated_delta_rule,prefill_attention_hd256}.cu).  HD-256 attention and the
 * chunk-wise prefill of the delta rule are not provided yet.  Arithmetic: oracle/qwen35_oracle.py. */
void rms_norm_batched_offset_cuda(const pk_bf16* x, const pk_bf16* weight, pk_bf16* out, int hidden_dim,
                                  int seq_len, float eps, pk_stream stream); /* out = x*rsqrt(mean+eps)*(1+w) */
void rms_norm_offset_cuda(const pk_bf16* x, const pk_bf16* weight, pk_bf16* out, int n, float eps,
                          pk_stream stream);
void rms_norm_gated_cuda(const pk_bf16* x, const float* weight, const pk_bf16* gate, pk_bf16* out,
                         int num_heads, int head_dim, float eps, pk_stream stream);
void gated_delta_rule_decode_cuda(const pk_bf16* qkv, const pk_bf16* b_proj, const pk_bf16* a_proj,
                                  const pk_bf16* dt_bias, const float* A_log, float* state,
                                  pk_bf16* output, int num_key_heads, int num_value_heads, int key_dim,
                                  int val_dim, pk_stream stream); /* key_dim == val_dim == 128 */
/* B200 extension: the same recurrence over a whole sequence (qkv_seq [T, 2*nk*128 + nv*128], b_seq / a_seq [T, nv],
 * out_seq [T, nv*128]) with the state held in registers -- correctness-first prefill until the chunk-wise kernel exists. */
int pk_b200_gated_delta_rule_prefill_recurrent(const pk_bf16* qkv_seq, const pk_bf16* b_seq, const pk_bf16* a_seq,
                                               const pk_bf16* dt_bias, const float* A_log, float* state,
                                               pk_bf16* out_seq, int num_key_heads, int num_value_heads,
                                               int key_dim, int val_dim, int seq_len, pk_stream stream);
void conv1d_prefill_cuda(const pk_bf16* x_seq, const pk_bf16* conv_weight, pk_bf16* conv_state,
                         pk_bf16* out_seq, int num_channels, int seq_len, int kernel_size,
                         pk_stream stream);
void prefill_attention_hd256_prep_cuda(const pk_bf16* q_full_batch, const pk_bf16* k_batch,
                                       const pk_bf16* v_batch, const pk_bf16* q_norm_weight,
                                       const pk_bf16* k_norm_weight, const pk_bf16* cos_cache,
                                       const pk_bf16* sin_cache, pk_bf16* q_batch_out, pk_bf16* k_cache,
                                       pk_bf16* v_cache, int num_q_heads, int num_kv_heads, int seq_len,
                                       const int* start_pos_ptr, int rotary_dim, float rms_eps,
                                       int max_seq_len, pk_stream stream);
void attention_gate_batch_hd256_cuda(const pk_bf16* q_full_batch, pk_bf16* attn_out, int num_q_heads,
                                     int seq_len, pk_stream stream);
void qk_norm_partial_rope_batched_decode_hd256_cuda(const pk_bf16* q_full_batch, pk_bf16* k_batch,
                                                    const pk_bf16* q_norm_weight,
                                                    const pk_bf16* k_norm_weight, const pk_bf16* cos_cache,
                                                    const pk_bf16* sin_cache, const int* positions,
                                                    pk_bf16* q_batch_out, int num_q_heads, int num_kv_heads,
                                                    int batch_size, int rotary_dim, float rms_eps,
                                                    pk_stream stream);

/* HD-256 causal GQA prefill attention over the paged cache (ffi.rs:1309-1334; FlashInfer FA2 HD 256 in the reference):
 * same arguments and planning contract as batch_prefill_paged_cuda (the tile plan is ignored consistently). */
int batch_prefill_paged_cuda_hd256(const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data, int64_t k_offset_elems,
                                   int64_t v_offset_elems, const int* page_indices, const int* page_indptr,
                                   const int* last_page_len_d, const int* q_indptr, const int* request_indices,
                                   const int* qo_tile_indices, const int* kv_tile_indices, const int* kv_chunk_size_ptr,
                                   const uint32_t* total_num_rows, int num_qo_heads, int num_kv_heads, int head_dim,
                                   int page_size, int seq_len, int batch_size, int padded_batch_size, int64_t stride_page,
                                   float sm_scale, pk_stream stream);

/* HD-256 paged decode attention (ffi.rs:1286-1307): q normed + roped, K/V already appended; GQA group 4, page 16
 * (the reference's only instantiation), else -1. */
int paged_attention_decode_cuda_hd256(const pk_bf16* q, pk_bf16* output, const pk_bf16* kv_data,
                                      int64_t k_offset_elems, int64_t v_offset_elems, const int* page_indices,
                                      const int* page_indptr, const int* last_page_len_d,
                                      const int* request_indices, const int* kv_tile_indices,
                                      const int* kv_chunk_size_ptr, int num_qo_heads, int num_kv_heads,
                                      int head_dim, int page_size, int batch_size, int64_t stride_page,
                                      float sm_scale, pk_stream stream);

/* ---- TP all-reduce hook (pegainfer-qwen3-4b/src/weights.rs:396-405) ----------
 * One-shot all-reduce over NVLink peer memory: every rank owns a symmetric staging
 * buffer; peers' buffers are mapped (cudaIpc / peer access).  In-place SUM over
 * [H, T] bf16 on the rank's stream, graph-capturable.
 *   pk_tp_comm: opaque; created from the world's staging + flag pointers as mapped in
 *   THIS process (index = rank). */
typedef struct pk_tp_comm pk_tp_comm;
/* Bytes of the per-rank flag array (zero-initialised device memory). */
int64_t pk_tp_flag_bytes(void);
/* staging_ptrs[p] / flag_ptrs[p]: rank p's staging / flag buffers as mapped in THIS process
 * (own buffers for p == rank, cudaIpc-opened for peers).  staging_bytes: size of each staging
 * buffer; it is cut into 2 slots x world source regions. */
pk_tp_comm* pk_tp_comm_create(int rank, int world, void* const* staging_ptrs,
                              void* const* flag_ptrs, int64_t staging_bytes);
void pk_tp_comm_destroy(pk_tp_comm* comm);
/* Largest row count one call can reduce for rows of hidden_dim bf16 (callers chunk above it). */
int64_t pk_tp_max_rows(pk_tp_comm* comm, int hidden_dim);
/* Vocab-sharded lm_head: combine every rank's (max logit, index inside its shard) into the global greedy token on
 * every rank (highest value, lowest global index on ties).  `index_inout[b]`: local index in, global winner out.
 * `step_counter` != NULL (CUDA-graph path): sequence = (*step_counter) * 256 + seq_or_op + 1; NULL: `seq_or_op` is
 * the absolute sequence number chosen by the host (must differ from call to call and be identical on all ranks). */
int pk_tp_top1_exchange(pk_tp_comm* comm, const pk_bf16* local_max, int* index_inout, int batch_size, int vocab_offset,
                        const uint32_t* step_counter, uint32_t seq_or_op, pk_stream stream);
/* all_reduce_hidden(&mut HiddenStates): in-place SUM (fp32 in rank order, one bf16 rounding). */
int pk_tp_all_reduce(pk_tp_comm* comm, pk_bf16* hidden, int64_t n, pk_stream stream);
int pk_tp_all_reduce_rows(pk_tp_comm* comm, pk_bf16* hidden, int hidden_dim, int rows,
                          pk_stream stream);
/* all-reduce(partial) fused with `hidden += sum; out = RMSNorm(hidden) * weight`. */
int pk_tp_all_reduce_add_rms_norm(pk_tp_comm* comm, pk_bf16* hidden, const pk_bf16* partial,
                                  const pk_bf16* weight, pk_bf16* out, int hidden_dim,
                                  int batch_size, float eps, pk_stream stream);
/* cudaIpc plumbing for one-process-per-GPU peers (64-byte handles travel over torch.distributed) */
int pk_tp_ipc_export(void* dev_ptr, void* handle_out_64);
int pk_tp_ipc_open(const void* handle_64, void** dev_ptr_out);
int pk_tp_ipc_close(void* dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* PEGAINFER_KERNELS_H_ */
