// Qwen3.5 hybrid-layer ops behind the reference's ABI names (the next scope row).
// Covered here: the memory-bound pieces -- (1+w) RMSNorm, gated per-head RMSNorm, causal depthwise conv1d + SiLU, the
// recurrent gated-delta-rule decode step, the HD-256 QK-norm + partial-RoPE preparation (prefill and batched decode)
// and the sigmoid output gate.  NOT yet here: HD-256 attention and the chunk-wise prefill of the delta rule
// (oracle/qwen35_chunkwise.py is its specification).  Arithmetic and rounding points follow the reference kernels
// cited per function; oracle/qwen35_oracle.py restates them and is pinned to HF.
// STATUS: compiled for sm_100a and exported; the GPU parity tests (tests/test_qwen35_ops_gpu.py) are opt-in
// (PK_TEST_QWEN35=1) until they have run on hardware.
#include <algorithm>

#include "common.cuh"

namespace pk {

// ---------------------------------------------------------------- (1+w) RMSNorm: flashinfer_norm.cu:108-133 (GemmaRMSNorm)
// out = bf16(x * rsqrt(mean(x^2) + eps) * (1 + w)), fp32 throughout, one rounding.  One CTA per row; the row is read
// once and parked in shared memory as fp32.
__global__ void rms_norm_offset_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ out,
                                       int dim, float eps) {
  extern __shared__ float q35_row[];  // dim floats + 33
  float* red = q35_row + dim;
  const size_t base = (size_t)blockIdx.x * dim;
  pdl_wait();
  float ss = 0.f;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    const float v = bf2f(x[base + i]);
    q35_row[i] = v;
    ss = fmaf(v, v, ss);
  }
  const float r = rsqrtf(block_sum(ss, red) / (float)dim + eps);
  for (int i = threadIdx.x; i < dim; i += blockDim.x) out[base + i] = f2bf(q35_row[i] * r * (1.0f + bf2f(w[i])));
}

// ---------------------------------------------------------------- gated per-head RMSNorm: norm.cu:17-61
// out = bf16(x * rsqrt(mean_head(x^2) + eps) * w_f32 * silu(gate)); one warp per head (head_dim <= 1024).
__global__ void rms_norm_gated_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const bf16* __restrict__ gate,
                                      bf16* __restrict__ out, int num_heads, int head_dim, float eps) {
  const int head = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (head >= num_heads) return;
  pdl_wait();
  const size_t base = (size_t)head * head_dim;
  float ss = 0.f;
  for (int i = lane; i < head_dim; i += 32) {
    const float v = bf2f(x[base + i]);
    ss = fmaf(v, v, ss);
  }
  const float r = rsqrtf(warp_sum(ss) / (float)head_dim + eps);
  for (int i = lane; i < head_dim; i += 32) {
    const float g = bf2f(gate[base + i]);
    out[base + i] = f2bf(bf2f(x[base + i]) * r * w[i] * (g / (1.0f + expf(-g))));
  }
}

// ---------------------------------------------------------------- causal depthwise conv1d + SiLU: conv1d.cu:19-84
// y[t, c] = bf16(silu(bf16(sum_k w[c, k] * x[t - (K-1) + k, c]))), history from conv_state [C, K-1] (oldest first).
// The state update is a second launch: in one launch (as the reference does it) the last position's threads would
// rewrite the state while other CTAs still read it for the first positions.
__global__ void conv1d_silu_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ state,
                                   bf16* __restrict__ out, int C, int T, int K) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)C * T) return;
  pdl_wait();
  const int c = (int)(idx % C), t = (int)(idx / C);
  const int sw = K - 1;
  float sum = 0.f;
  for (int k = 0; k < K; ++k) {
    const int src = t - sw + k;
    float v = 0.f;
    if (src >= 0) v = bf2f(x[(size_t)src * C + c]);
    else if (sw + src >= 0) v = bf2f(state[(size_t)c * sw + sw + src]);
    sum += v * bf2f(w[(size_t)c * K + k]);
  }
  const float s = round_bf16(sum);
  out[(size_t)t * C + c] = f2bf(s / (1.0f + expf(-s)));
}
__global__ void conv1d_state_kernel(const bf16* __restrict__ x, bf16* __restrict__ state, int C, int T, int K) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int sw = K - 1;
  bf16 nw[8];
  for (int i = 0; i < sw && i < 8; ++i) {
    const int src = T - sw + i;  // position of the i-th newest-window entry
    nw[i] = src >= 0 ? x[(size_t)src * C + c] : (sw + src >= 0 ? state[(size_t)c * sw + sw + src] : f2bf(0.f));
  }
  for (int i = 0; i < sw && i < 8; ++i) state[(size_t)c * sw + i] = nw[i];
}

// ---------------------------------------------------------------- gated delta rule, one decode step: gated_delta_rule.cu:27-166
// One CTA per value head, 512 threads = 128 value columns x 4 key slices; every state element is read ONCE into
// registers (32 per thread), decayed, corrected and written back once (the reference re-reads it for the second pass).
// HBM-bound: 2 * 4 * dk * dv bytes per head per token.
constexpr int GK = 128, GV = 128, GS = 4, GJ = GK / GS;
__global__ void __launch_bounds__(GV * GS)
gated_delta_rule_decode_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ b_proj, const bf16* __restrict__ a_proj,
                               const bf16* __restrict__ dt_bias, const float* __restrict__ a_log, float* __restrict__ state,
                               bf16* __restrict__ output, int nk, int nv) {
  __shared__ float sq[GK], sk[GK], part[GS][GV], red[40];
  __shared__ float s_decay, s_beta;
  const int vh = blockIdx.x, col = threadIdx.x & (GV - 1), sl = threadIdx.x >> 7;
  const int kh = vh * nk / nv;
  pdl_wait();
  // L2-normalise q and k of the key head (eps 1e-12), q additionally scaled by rsqrt(dk)
  float qv = 0.f, kv = 0.f;
  if (sl == 0) {
    qv = bf2f(qkv[(size_t)kh * GK + col]);
    kv = bf2f(qkv[(size_t)nk * GK + (size_t)kh * GK + col]);
  }
  const float qn = block_sum(qv * qv, red);
  const float kn = block_sum(kv * kv, red);
  if (sl == 0) {
    sq[col] = qv * rsqrtf(qn + 1e-12f) * rsqrtf((float)GK);
    sk[col] = kv * rsqrtf(kn + 1e-12f);
  }
  if (threadIdx.x == 0) {
    const float x = bf2f(a_proj[vh]) + bf2f(dt_bias[vh]);
    const float sp = x > 20.0f ? x : logf(1.0f + expf(x));
    s_decay = expf(-expf(a_log[vh]) * sp);
    s_beta = 1.0f / (1.0f + expf(-bf2f(b_proj[vh])));
  }
  __syncthreads();
  const float decay = s_decay, beta = s_beta;
  const float vv = bf2f(qkv[(size_t)2 * nk * GK + (size_t)vh * GV + col]);
  float* st = state + ((size_t)vh * GK + (size_t)sl * GJ) * GV + col;
  float s[GJ];
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < GJ; ++j) {
    s[j] = st[(size_t)j * GV] * decay;
    acc = fmaf(s[j], sk[sl * GJ + j], acc);
  }
  part[sl][col] = acc;
  __syncthreads();
  const float delta = (vv - (part[0][col] + part[1][col] + part[2][col] + part[3][col])) * beta;
  __syncthreads();
  acc = 0.f;
#pragma unroll
  for (int j = 0; j < GJ; ++j) {
    s[j] = fmaf(delta, sk[sl * GJ + j], s[j]);
    st[(size_t)j * GV] = s[j];
    acc = fmaf(s[j], sq[sl * GJ + j], acc);
  }
  part[sl][col] = acc;
  __syncthreads();
  if (sl == 0) output[(size_t)vh * GV + col] = f2bf(part[0][col] + part[1][col] + part[2][col] + part[3][col]);
}

// Sequence form of the same recurrence (our extension, the correctness-first prefill until the chunk-wise tensor-core
// kernel of oracle/qwen35_chunkwise.py exists): one CTA per value head walks the T tokens with the 128 x 128 fp32 state
// held in REGISTERS the whole time (32 elements per thread), so the state costs one HBM read and one write per call and
// each token only streams its q/k/v rows.  Latency-bound (~4 block barriers per token), not a roofline kernel.
__global__ void __launch_bounds__(GV * GS)
gated_delta_rule_seq_kernel(const bf16* __restrict__ qkv_seq, const bf16* __restrict__ b_seq, const bf16* __restrict__ a_seq,
                            const bf16* __restrict__ dt_bias, const float* __restrict__ a_log, float* __restrict__ state,
                            bf16* __restrict__ out_seq, int nk, int nv, int T) {
  __shared__ float sq[GK], sk[GK], part[GS][GV], red[40];
  __shared__ float s_decay, s_beta;
  const int vh = blockIdx.x, col = threadIdx.x & (GV - 1), sl = threadIdx.x >> 7;
  const int kh = vh * nk / nv;
  const int qkv_dim = 2 * nk * GK + nv * GV;
  pdl_wait();
  float* st = state + ((size_t)vh * GK + (size_t)sl * GJ) * GV + col;
  float s[GJ];
#pragma unroll
  for (int j = 0; j < GJ; ++j) s[j] = st[(size_t)j * GV];
  const float bias = bf2f(dt_bias[vh]), neg_exp_a = -expf(a_log[vh]);
  for (int t = 0; t < T; ++t) {
    const bf16* row = qkv_seq + (size_t)t * qkv_dim;
    float qv = 0.f, kv = 0.f;
    if (sl == 0) {
      qv = bf2f(row[(size_t)kh * GK + col]);
      kv = bf2f(row[(size_t)nk * GK + (size_t)kh * GK + col]);
    }
    const float qn = block_sum(qv * qv, red);
    const float kn = block_sum(kv * kv, red);
    if (sl == 0) {
      sq[col] = qv * rsqrtf(qn + 1e-12f) * rsqrtf((float)GK);
      sk[col] = kv * rsqrtf(kn + 1e-12f);
    }
    if (threadIdx.x == 0) {
      const float x = bf2f(a_seq[(size_t)t * nv + vh]) + bias;
      const float sp = x > 20.0f ? x : logf(1.0f + expf(x));
      s_decay = expf(neg_exp_a * sp);
      s_beta = 1.0f / (1.0f + expf(-bf2f(b_seq[(size_t)t * nv + vh])));
    }
    __syncthreads();
    const float decay = s_decay, beta = s_beta;
    const float vv = bf2f(row[(size_t)2 * nk * GK + (size_t)vh * GV + col]);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < GJ; ++j) {
      s[j] *= decay;
      acc = fmaf(s[j], sk[sl * GJ + j], acc);
    }
    part[sl][col] = acc;
    __syncthreads();
    const float delta = (vv - (part[0][col] + part[1][col] + part[2][col] + part[3][col])) * beta;
    __syncthreads();
    acc = 0.f;
#pragma unroll
    for (int j = 0; j < GJ; ++j) {
      s[j] = fmaf(delta, sk[sl * GJ + j], s[j]);
      acc = fmaf(s[j], sq[sl * GJ + j], acc);
    }
    part[sl][col] = acc;
    __syncthreads();
    if (sl == 0) out_seq[(size_t)t * nv * GV + (size_t)vh * GV + col] = f2bf(part[0][col] + part[1][col] + part[2][col] + part[3][col]);
    // the next token's block_sum starts with a barrier, which also protects sq / sk / part
  }
#pragma unroll
  for (int j = 0; j < GJ; ++j) st[(size_t)j * GV] = s[j];
}

// ---------------------------------------------------------------- delta rule over a sequence, version 2 (round 2)
// The per-token critical path of gated_delta_rule_seq_kernel is six block barriers (two norm reductions among them) on a
// 512-thread CTA, and only 32 CTAs exist (one per value head): 61 ms TTFT at 1024 tokens, ~60 % of it here.  Everything
// that does not depend on the state -- the L2 norms of q and k, the q scale, decay = exp(-exp(A_log) softplus(a + dt_bias))
// and beta = sigmoid(b) -- is hoisted into a token-parallel pre-pass (fp32, same formulas); the sequential kernel then
// needs TWO barriers per token (the two 4-way partial sums), runs 4 CTAs per value head (32 state columns each, 128
// threads: cheaper barriers, 128 CTAs) and stages the next token's rows while the current one is reduced.
__global__ void __launch_bounds__(GK) gdr_prepare_kernel(const bf16* __restrict__ qkv_seq, const bf16* __restrict__ b_seq,
                                                         const bf16* __restrict__ a_seq, const bf16* __restrict__ dt_bias,
                                                         const float* __restrict__ a_log, float* __restrict__ qn, float* __restrict__ kn,
                                                         float* __restrict__ decay, float* __restrict__ beta, int nk, int nv, int T) {
  __shared__ float red[40];
  pdl_wait();
  const int t = blockIdx.x, kh = blockIdx.y, d = threadIdx.x;
  const int qkv_dim = 2 * nk * GK + nv * GV;
  if (kh == nk) {  // one extra block row: the per-(token, value head) scalars
    for (int vh = d; vh < nv; vh += GK) {
      const float x = bf2f(a_seq[(size_t)t * nv + vh]) + bf2f(dt_bias[vh]);
      const float sp = x > 20.0f ? x : logf(1.0f + expf(x));
      decay[(size_t)t * nv + vh] = expf(-expf(a_log[vh]) * sp);
      beta[(size_t)t * nv + vh] = 1.0f / (1.0f + expf(-bf2f(b_seq[(size_t)t * nv + vh])));
    }
    return;
  }
  const bf16* row = qkv_seq + (size_t)t * qkv_dim;
  const float qv = bf2f(row[(size_t)kh * GK + d]), kv = bf2f(row[(size_t)nk * GK + (size_t)kh * GK + d]);
  const float qs = block_sum(qv * qv, red);
  const float ks = block_sum(kv * kv, red);
  qn[((size_t)t * nk + kh) * GK + d] = qv * rsqrtf(qs + 1e-12f) * rsqrtf((float)GK);
  kn[((size_t)t * nk + kh) * GK + d] = kv * rsqrtf(ks + 1e-12f);
}

constexpr int G2C = 32;            // state columns per CTA
constexpr int G2T = G2C * GS;      // 128 threads: col = tid & 31, slice = tid >> 5 (32 key dims each)
__global__ void __launch_bounds__(G2T)
gated_delta_rule_seq2_kernel(const bf16* __restrict__ qkv_seq, const float* __restrict__ qn, const float* __restrict__ kn,
                             const float* __restrict__ decay_s, const float* __restrict__ beta_s, float* __restrict__ state,
                             bf16* __restrict__ out_seq, int nk, int nv, int T) {
  __shared__ float sq[2][GK], sk[2][GK], pa[GS][G2C], pb[GS][G2C];
  const int vh = blockIdx.x, cb = blockIdx.y, col = threadIdx.x & (G2C - 1), sl = threadIdx.x >> 5;
  const int kh = vh * nk / nv;
  const int qkv_dim = 2 * nk * GK + nv * GV;
  const int vcol = cb * G2C + col;
  pdl_wait();
  float* st = state + ((size_t)vh * GK + (size_t)sl * GJ) * GV + vcol;
  float s[GJ];
#pragma unroll
  for (int j = 0; j < GJ; ++j) s[j] = st[(size_t)j * GV];
  // stage token 0's normalised q / k rows (128 floats each: one element per thread)
  sq[0][threadIdx.x] = qn[((size_t)0 * nk + kh) * GK + threadIdx.x];
  sk[0][threadIdx.x] = kn[((size_t)0 * nk + kh) * GK + threadIdx.x];
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const int cur = t & 1;
    // requests for the next token go out before this token's dependent chain
    float nq = 0.f, nkk = 0.f;
    if (t + 1 < T) {
      nq = qn[((size_t)(t + 1) * nk + kh) * GK + threadIdx.x];
      nkk = kn[((size_t)(t + 1) * nk + kh) * GK + threadIdx.x];
    }
    const float decay = decay_s[(size_t)t * nv + vh], beta = beta_s[(size_t)t * nv + vh];
    const float vv = bf2f(qkv_seq[(size_t)t * qkv_dim + (size_t)2 * nk * GK + (size_t)vh * GV + vcol]);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < GJ; ++j) {
      s[j] *= decay;
      acc = fmaf(s[j], sk[cur][sl * GJ + j], acc);
    }
    pa[sl][col] = acc;
    __syncthreads();  // barrier 1: the four slices' k.S partial sums
    const float delta = (vv - (pa[0][col] + pa[1][col] + pa[2][col] + pa[3][col])) * beta;
    acc = 0.f;
#pragma unroll
    for (int j = 0; j < GJ; ++j) {
      s[j] = fmaf(delta, sk[cur][sl * GJ + j], s[j]);
      acc = fmaf(s[j], sq[cur][sl * GJ + j], acc);
    }
    pb[sl][col] = acc;
    sq[cur ^ 1][threadIdx.x] = nq;  // the other buffer: nobody reads it during this token
    sk[cur ^ 1][threadIdx.x] = nkk;
    __syncthreads();  // barrier 2: q.S partial sums + the next token's rows
    if (sl == 0) out_seq[(size_t)t * nv * GV + (size_t)vh * GV + vcol] = f2bf(pb[0][col] + pb[1][col] + pb[2][col] + pb[3][col]);
    // pa is rewritten after barrier 2, pb after the next barrier 1: both reads above are behind a barrier by then
  }
#pragma unroll
  for (int j = 0; j < GJ; ++j) st[(size_t)j * GV] = s[j];
}

// ---------------------------------------------------------------- HD-256 QK prep: prefill_attention_hd256.cu:7-113,176-262
// One warp per (head, token): lane owns 8 consecutive dims (one 16-byte vector), so the per-head RMS is a warp
// reduction and the RoPE partner (dim +- rotary_dim/2) sits rotary_dim/16 lanes away -- no shared memory, no barrier.
// normed = bf16(x * (1/sqrt(mean+eps)) * (1+w)); rotary dims: bf16(lo*c - hi*s), bf16(lo*s + hi*c) with bf16 cos/sin
// read at [pos * rotary_dim + d], d < rotary_dim/2.
constexpr int HD2 = 256;
__device__ __forceinline__ void hd256_norm_rope(const bf16* __restrict__ src, const bf16* __restrict__ w,
                                                const bf16* __restrict__ cosc, const bf16* __restrict__ sinc, int pos,
                                                int rotary_dim, float eps, bf16* __restrict__ dst, int lane) {
  const uint4 raw = reinterpret_cast<const uint4*>(src)[lane];
  const uint4 wr = reinterpret_cast<const uint4*>(w)[lane];
  float v[8] = {bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y), bf16_lo(raw.z), bf16_hi(raw.z), bf16_lo(raw.w), bf16_hi(raw.w)};
  const float wv[8] = {bf16_lo(wr.x), bf16_hi(wr.x), bf16_lo(wr.y), bf16_hi(wr.y), bf16_lo(wr.z), bf16_hi(wr.z), bf16_lo(wr.w), bf16_hi(wr.w)};
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
  const float inv = 1.0f / sqrtf(warp_sum(ss) / (float)HD2 + eps);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = round_bf16(v[j] * inv * (1.0f + wv[j]));
  const int half_lanes = rotary_dim >> 4;  // lanes per rotary half (rotary_dim / 2 / 8)
  const bool rot = lane < 2 * half_lanes, is_lo = lane < half_lanes;
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float other = __shfl_sync(0xffffffffu, v[j], rot ? (is_lo ? lane + half_lanes : lane - half_lanes) : lane);
    o[j] = v[j];
    if (rot) {
      const int d = (is_lo ? lane : lane - half_lanes) * 8 + j;  // index inside the first rotary half
      const float c = bf2f(cosc[(size_t)pos * rotary_dim + d]), s = bf2f(sinc[(size_t)pos * rotary_dim + d]);
      o[j] = is_lo ? v[j] * c - other * s : other * s + v[j] * c;
    }
  }
  uint4 res;
  res.x = pack_bf16(o[0], o[1]); res.y = pack_bf16(o[2], o[3]); res.z = pack_bf16(o[4], o[5]); res.w = pack_bf16(o[6], o[7]);
  reinterpret_cast<uint4*>(dst)[lane] = res;
}

// prefill: q_full [T][nq][q(256) | gate(256)] -> q_out [T][nq][256]; k -> k_cache [nkv][max_seq][256] at start + t;
// v copied alongside.
__global__ void hd256_prefill_prep_kernel(const bf16* __restrict__ q_full, const bf16* __restrict__ k, const bf16* __restrict__ v,
                                          const bf16* __restrict__ qw, const bf16* __restrict__ kw, const bf16* __restrict__ cosc,
                                          const bf16* __restrict__ sinc, bf16* __restrict__ q_out, bf16* __restrict__ k_cache,
                                          bf16* __restrict__ v_cache, int nq, int nkv, int T, const int* __restrict__ start_pos_ptr,
                                          int rotary_dim, float eps, int max_seq) {
  const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // (token, head) with head < nq + 2 nkv
  const int lane = threadIdx.x & 31, per_tok = nq + 2 * nkv;
  if (item >= (int64_t)T * per_tok) return;
  pdl_wait();
  const int t = (int)(item / per_tok), h = (int)(item % per_tok);
  const int pos = *start_pos_ptr + t;
  if (h < nq) {
    hd256_norm_rope(q_full + ((size_t)t * nq + h) * 2 * HD2, qw, cosc, sinc, pos, rotary_dim, eps, q_out + ((size_t)t * nq + h) * HD2, lane);
  } else if (h < nq + nkv) {
    const int kh = h - nq;
    hd256_norm_rope(k + ((size_t)t * nkv + kh) * HD2, kw, cosc, sinc, pos, rotary_dim, eps, k_cache + ((size_t)kh * max_seq + pos) * HD2, lane);
  } else {
    const int kh = h - nq - nkv;
    reinterpret_cast<uint4*>(v_cache + ((size_t)kh * max_seq + pos) * HD2)[lane] = reinterpret_cast<const uint4*>(v + ((size_t)t * nkv + kh) * HD2)[lane];
  }
}

// batched decode: per-request positions, K normalised + roped in place
__global__ void hd256_decode_prep_kernel(const bf16* __restrict__ q_full, bf16* __restrict__ k, const bf16* __restrict__ qw,
                                         const bf16* __restrict__ kw, const bf16* __restrict__ cosc, const bf16* __restrict__ sinc,
                                         const int* __restrict__ positions, bf16* __restrict__ q_out, int nq, int nkv, int bs,
                                         int rotary_dim, float eps) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31, per_tok = nq + nkv;
  if (item >= bs * per_tok) return;
  pdl_wait();
  const int t = item / per_tok, h = item % per_tok, pos = positions[t];
  if (h < nq) hd256_norm_rope(q_full + ((size_t)t * nq + h) * 2 * HD2, qw, cosc, sinc, pos, rotary_dim, eps, q_out + ((size_t)t * nq + h) * HD2, lane);
  else hd256_norm_rope(k + ((size_t)t * nkv + (h - nq)) * HD2, kw, cosc, sinc, pos, rotary_dim, eps, k + ((size_t)t * nkv + (h - nq)) * HD2, lane);
}

// attn_out[t][h][d] *= sigmoid(gate[t][h][d]), gate = second half of the head's q_full rows: prefill_attention_hd256.cu:134-157
__global__ void hd256_gate_kernel(const bf16* __restrict__ q_full, bf16* __restrict__ attn_out, int nq, int T) {
  const int64_t i8 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one 8-element vector
  if (i8 >= (int64_t)T * nq * (HD2 / 8)) return;
  pdl_wait();
  const int64_t row = i8 / (HD2 / 8);  // (t, h)
  const int vcol = (int)(i8 % (HD2 / 8));
  const uint4 g = reinterpret_cast<const uint4*>(q_full + row * 2 * HD2 + HD2)[vcol];
  uint4 o = reinterpret_cast<uint4*>(attn_out + row * HD2)[vcol];
  auto f = [](uint32_t ov, uint32_t gv) {
    const float a = bf16_lo(ov) * (1.0f / (1.0f + expf(-bf16_lo(gv)))), b = bf16_hi(ov) * (1.0f / (1.0f + expf(-bf16_hi(gv))));
    return pack_bf16(a, b);
  };
  o.x = f(o.x, g.x); o.y = f(o.y, g.y); o.z = f(o.z, g.z); o.w = f(o.w, g.w);
  reinterpret_cast<uint4*>(attn_out + row * HD2)[vcol] = o;
}

}  // namespace pk

using namespace pk;

extern "C" {

void rms_norm_batched_offset_cuda(const pk_bf16* x, const pk_bf16* weight, pk_bf16* out, int hidden_dim, int seq_len, float eps,
                                  pk_stream stream) {
  if (hidden_dim <= 0 || seq_len <= 0) return;
  const size_t smem = sizeof(float) * ((size_t)hidden_dim + 40);
  if (smem > 48 * 1024) cudaFuncSetAttribute(rms_norm_offset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch(rms_norm_offset_kernel, dim3(seq_len), dim3(256), smem, stream, true, (const bf16*)x, (const bf16*)weight, (bf16*)out,
         hidden_dim, eps);
}
void rms_norm_offset_cuda(const pk_bf16* x, const pk_bf16* weight, pk_bf16* out, int n, float eps, pk_stream stream) {
  rms_norm_batched_offset_cuda(x, weight, out, n, 1, eps, stream);
}
void rms_norm_gated_cuda(const pk_bf16* x, const float* weight, const pk_bf16* gate, pk_bf16* out, int num_heads, int head_dim,
                         float eps, pk_stream stream) {
  if (num_heads <= 0 || head_dim <= 0) return;
  launch(rms_norm_gated_kernel, dim3((num_heads + 7) / 8), dim3(256), 0, stream, true, (const bf16*)x, weight, (const bf16*)gate,
         (bf16*)out, num_heads, head_dim, eps);
}
void conv1d_prefill_cuda(const pk_bf16* x_seq, const pk_bf16* conv_weight, pk_bf16* conv_state, pk_bf16* out_seq, int num_channels,
                         int seq_len, int kernel_size, pk_stream stream) {
  if (num_channels <= 0 || seq_len <= 0) return;
  if (kernel_size < 1 || kernel_size > 9) unsupported("conv1d_prefill_cuda", "kernel_size must be 1..9");
  const int64_t total = (int64_t)num_channels * seq_len;
  launch(conv1d_silu_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, true, (const bf16*)x_seq,
         (const bf16*)conv_weight, (const bf16*)conv_state, (bf16*)out_seq, num_channels, seq_len, kernel_size);
  launch(conv1d_state_kernel, dim3((num_channels + 255) / 256), dim3(256), 0, stream, false, (const bf16*)x_seq, (bf16*)conv_state,
         num_channels, seq_len, kernel_size);
}
void gated_delta_rule_decode_cuda(const pk_bf16* qkv, const pk_bf16* b_proj, const pk_bf16* a_proj, const pk_bf16* dt_bias,
                                  const float* A_log, float* state, pk_bf16* output, int num_key_heads, int num_value_heads,
                                  int key_dim, int val_dim, pk_stream stream) {
  if (num_value_heads <= 0 || num_key_heads <= 0) return;
  if (key_dim != GK || val_dim != GV)  // the reference's fixed 128 x 128 heads (gated_delta_rule.cu asserts the same)
    unsupported("gated_delta_rule_decode_cuda", "key_dim and val_dim must be 128");
  launch(gated_delta_rule_decode_kernel, dim3(num_value_heads), dim3(GV * GS), 0, stream, true, (const bf16*)qkv, (const bf16*)b_proj,
         (const bf16*)a_proj, (const bf16*)dt_bias, A_log, state, (bf16*)output, num_key_heads, num_value_heads);
}
int pk_b200_gated_delta_rule_prefill_recurrent(const pk_bf16* qkv_seq, const pk_bf16* b_seq, const pk_bf16* a_seq, const pk_bf16* dt_bias,
                                               const float* A_log, float* state, pk_bf16* out_seq, int num_key_heads,
                                               int num_value_heads, int key_dim, int val_dim, int seq_len, pk_stream stream) {
  if (key_dim != GK || val_dim != GV || num_value_heads <= 0 || num_key_heads <= 0) return -1;
  if (seq_len <= 0) return 0;
  // version 2 (token-parallel pre-pass + 2-barrier sequential kernel on 4 CTAs per head) when the thread's workspace
  // is there (cublas_init); the sequence is cut into pieces that fit it, the state carries over in `state`.
  ThreadState& ts = tls();
  static const bool use_v1 = [] { const char* e = getenv("PK_GDR_SEQ"); return e && atoi(e) == 1; }();
  const size_t per_tok = ((size_t)2 * num_key_heads * GK + 2 * num_value_heads) * sizeof(float);
  if (!use_v1 && ts.gemm_part && ts.gemm_part_bytes >= per_tok * 64) {
    const int max_t = (int)std::min<size_t>(ts.gemm_part_bytes / per_tok, 1 << 20);
    for (int t0 = 0; t0 < seq_len; t0 += max_t) {
      const int n = std::min(max_t, seq_len - t0);
      float* qn = ts.gemm_part;
      float* kn = qn + (size_t)n * num_key_heads * GK;
      float* dec = kn + (size_t)n * num_key_heads * GK;
      float* bet = dec + (size_t)n * num_value_heads;
      const int qkv_dim = 2 * num_key_heads * GK + num_value_heads * GV;
      const bf16* qs = (const bf16*)qkv_seq + (size_t)t0 * qkv_dim;
      const bf16* bs = (const bf16*)b_seq + (size_t)t0 * num_value_heads;
      const bf16* as = (const bf16*)a_seq + (size_t)t0 * num_value_heads;
      cudaError_t e = launch(gdr_prepare_kernel, dim3(n, num_key_heads + 1), dim3(GK), 0, stream, true, qs, bs, as, (const bf16*)dt_bias, A_log,
                             qn, kn, dec, bet, num_key_heads, num_value_heads, n);
      if (e != cudaSuccess) return (int)e;
      e = launch(gated_delta_rule_seq2_kernel, dim3(num_value_heads, GV / G2C), dim3(G2T), 0, stream, true, qs, (const float*)qn,
                 (const float*)kn, (const float*)dec, (const float*)bet, state, (bf16*)out_seq + (size_t)t0 * num_value_heads * GV,
                 num_key_heads, num_value_heads, n);
      if (e != cudaSuccess) return (int)e;
    }
    return 0;
  }
  return (int)launch(gated_delta_rule_seq_kernel, dim3(num_value_heads), dim3(GV * GS), 0, stream, true, (const bf16*)qkv_seq,
                     (const bf16*)b_seq, (const bf16*)a_seq, (const bf16*)dt_bias, A_log, state, (bf16*)out_seq, num_key_heads,
                     num_value_heads, seq_len);
}
void prefill_attention_hd256_prep_cuda(const pk_bf16* q_full_batch, const pk_bf16* k_batch, const pk_bf16* v_batch,
                                       const pk_bf16* q_norm_weight, const pk_bf16* k_norm_weight, const pk_bf16* cos_cache,
                                       const pk_bf16* sin_cache, pk_bf16* q_batch_out, pk_bf16* k_cache, pk_bf16* v_cache,
                                       int num_q_heads, int num_kv_heads, int seq_len, const int* start_pos_ptr, int rotary_dim,
                                       float rms_eps, int max_seq_len, pk_stream stream) {
  if (seq_len <= 0) return;
  if (rotary_dim % 16 != 0 || rotary_dim > HD2) unsupported("prefill_attention_hd256_prep_cuda", "rotary_dim must be a multiple of 16, <= 256");
  const int64_t items = (int64_t)seq_len * (num_q_heads + 2 * num_kv_heads);
  launch(hd256_prefill_prep_kernel, dim3((unsigned)((items + 7) / 8)), dim3(256), 0, stream, true, (const bf16*)q_full_batch,
         (const bf16*)k_batch, (const bf16*)v_batch, (const bf16*)q_norm_weight, (const bf16*)k_norm_weight, (const bf16*)cos_cache,
         (const bf16*)sin_cache, (bf16*)q_batch_out, (bf16*)k_cache, (bf16*)v_cache, num_q_heads, num_kv_heads, seq_len, start_pos_ptr,
         rotary_dim, rms_eps, max_seq_len);
}
void qk_norm_partial_rope_batched_decode_hd256_cuda(const pk_bf16* q_full_batch, pk_bf16* k_batch, const pk_bf16* q_norm_weight,
                                                    const pk_bf16* k_norm_weight, const pk_bf16* cos_cache, const pk_bf16* sin_cache,
                                                    const int* positions, pk_bf16* q_batch_out, int num_q_heads, int num_kv_heads,
                                                    int batch_size, int rotary_dim, float rms_eps, pk_stream stream) {
  if (batch_size <= 0) return;
  if (rotary_dim % 16 != 0 || rotary_dim > HD2)
    unsupported("qk_norm_partial_rope_batched_decode_hd256_cuda", "rotary_dim must be a multiple of 16, <= 256");
  const int items = batch_size * (num_q_heads + num_kv_heads);
  launch(hd256_decode_prep_kernel, dim3((items + 7) / 8), dim3(256), 0, stream, true, (const bf16*)q_full_batch, (bf16*)k_batch,
         (const bf16*)q_norm_weight, (const bf16*)k_norm_weight, (const bf16*)cos_cache, (const bf16*)sin_cache, positions,
         (bf16*)q_batch_out, num_q_heads, num_kv_heads, batch_size, rotary_dim, rms_eps);
}
void attention_gate_batch_hd256_cuda(const pk_bf16* q_full_batch, pk_bf16* attn_out, int num_q_heads, int seq_len, pk_stream stream) {
  if (num_q_heads <= 0 || seq_len <= 0) return;
  const int64_t vecs = (int64_t)seq_len * num_q_heads * (HD2 / 8);
  launch(hd256_gate_kernel, dim3((unsigned)((vecs + 255) / 256)), dim3(256), 0, stream, true, (const bf16*)q_full_batch, (bf16*)attn_out,
         num_q_heads, seq_len);
}

}  // extern "C"
