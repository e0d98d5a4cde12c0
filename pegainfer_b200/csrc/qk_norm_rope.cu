// Per-head QK RMSNorm + NeoX RoPE, in place on Q and K.  Replaces
// csrc/prefill_attention.cu:12-159 of the reference, same rounding points:
//   normed = bf16(x * inv_rms);  t = bf16(f32(normed) * w);
//   lo' = bf16(lo*c - hi*s), hi' = bf16(lo*s + hi*c)  with bf16 cos/sin tables.
// One WARP per (head, token): lane l owns elements 4l..4l+3 (one 8-byte access); the RoPE
// partner (i, i+64) lives in lane l^16, exchanged with a shuffle -- no shared memory, no
// block barrier (the reference uses a 128-thread block and two __syncthreads per head).
#include "common.cuh"

namespace pk {

__global__ void qk_norm_rope_kernel(bf16* __restrict__ q, bf16* __restrict__ k,
                                    const bf16* __restrict__ qw, const bf16* __restrict__ kw,
                                    const bf16* __restrict__ cosc, const bf16* __restrict__ sinc,
                                    const int* __restrict__ positions, int start_pos, int nq,
                                    int nkv, int tokens, float eps) {
  constexpr int HD = 128;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int heads = nq + nkv;
  if (warp >= heads * tokens) return;
  const int token = warp / heads, head = warp - token * heads;
  const bool is_q = head < nq;
  bf16* row = is_q ? q + ((size_t)token * nq + head) * HD
                   : k + ((size_t)token * nkv + (head - nq)) * HD;
  const bf16* w = is_q ? qw : kw;
  pdl_wait();
  const int pos = positions ? __ldg(positions + token) : start_pos + token;

  const uint2 raw = reinterpret_cast<const uint2*>(row)[lane];
  float v[4] = {bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y)};
  float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss / (float)HD + eps);
  const uint2 wr = reinterpret_cast<const uint2*>(w)[lane];
  const float wv[4] = {bf16_lo(wr.x), bf16_hi(wr.x), bf16_lo(wr.y), bf16_hi(wr.y)};
  float t[4], o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) t[j] = round_bf16(round_bf16(v[j] * inv) * wv[j]);
  // table layout [pos*128 + i] duplicated for i and i+64 (weight_loader.rs:210-244)
  const int ci = (lane & 15) * 4;
  const uint2 cr = reinterpret_cast<const uint2*>(cosc + (size_t)pos * HD + ci)[0];
  const uint2 sr = reinterpret_cast<const uint2*>(sinc + (size_t)pos * HD + ci)[0];
  const float c[4] = {bf16_lo(cr.x), bf16_hi(cr.x), bf16_lo(cr.y), bf16_hi(cr.y)};
  const float s[4] = {bf16_lo(sr.x), bf16_hi(sr.x), bf16_lo(sr.y), bf16_hi(sr.y)};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float other = __shfl_xor_sync(0xffffffffu, t[j], 16);
    // lanes 0..15 hold "lo" (d < 64): lo*c - hi*s ; lanes 16..31 hold "hi": lo*s + hi*c
    o[j] = lane < 16 ? t[j] * c[j] - other * s[j] : other * s[j] + t[j] * c[j];
  }
  uint2 res;
  res.x = pack_bf16(o[0], o[1]);
  res.y = pack_bf16(o[2], o[3]);
  reinterpret_cast<uint2*>(row)[lane] = res;
}

static void launch_qk(pk_bf16* q, pk_bf16* k, const pk_bf16* qw, const pk_bf16* kw,
                      const pk_bf16* cosc, const pk_bf16* sinc, const int* positions,
                      int start_pos, int nq, int nkv, int head_dim, int tokens, float eps,
                      pk_stream stream) {
  if (head_dim != 128 || tokens <= 0) return;  // HEAD_DIM is hard-coded 128 in the reference too
  const int warps = (nq + nkv) * tokens;
  const int block = 128;
  launch(qk_norm_rope_kernel, dim3((warps * 32 + block - 1) / block), dim3(block), 0, stream, true,
         (bf16*)q, (bf16*)k, (const bf16*)qw, (const bf16*)kw, (const bf16*)cosc,
         (const bf16*)sinc, positions, start_pos, nq, nkv, tokens, eps);
}

}  // namespace pk

extern "C" {

void prefill_qk_norm_rope_only_cuda(pk_bf16* q_batch, pk_bf16* k_batch,
                                    const pk_bf16* q_norm_weight, const pk_bf16* k_norm_weight,
                                    const pk_bf16* cos_cache, const pk_bf16* sin_cache,
                                    int num_q_heads, int num_kv_heads, int head_dim, int seq_len,
                                    int start_pos, float rms_eps, pk_stream stream) {
  pk::launch_qk(q_batch, k_batch, q_norm_weight, k_norm_weight, cos_cache, sin_cache, nullptr,
                start_pos, num_q_heads, num_kv_heads, head_dim, seq_len, rms_eps, stream);
}

void qk_norm_rope_batched_decode_cuda(pk_bf16* q, pk_bf16* k, const pk_bf16* q_norm_weight,
                                      const pk_bf16* k_norm_weight, const pk_bf16* cos_cache,
                                      const pk_bf16* sin_cache, const int* positions,
                                      int num_q_heads, int num_kv_heads, int head_dim,
                                      int batch_size, float rms_eps, pk_stream stream) {
  pk::launch_qk(q, k, q_norm_weight, k_norm_weight, cos_cache, sin_cache, positions, 0,
                num_q_heads, num_kv_heads, head_dim, batch_size, rms_eps, stream);
}

}  // extern "C"
