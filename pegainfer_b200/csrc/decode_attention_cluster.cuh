// Argument block of the cluster decode-attention kernel (decode_attention_cluster.cu), shared with the entry point
// in decode_attention.cu.
#pragma once
#include "common.cuh"

namespace pk {

// L2 prefetch span: the weight rows a FOLLOWING GEMV launch will stream.  The rows are cut into `slices` row ranges
// exactly as that GEMV's grid cuts them ([s*rows/slices, (s+1)*rows/slices)); the first `pf_rows` rows of every
// slice are requested, so every GEMV CTA finds the same share of its slice in L2.
struct PfSpan {
  const uint8_t* base;
  int rows, row_bytes, slices, pf_rows;
};
constexpr int kMaxPfSpans = 4;

struct ClusterAttnArgs {
  const bf16 *q, *k_new, *v_new;
  bf16* out;
  bf16* kv;
  int64_t k_off, v_off, stride_page;
  const int *page_indices, *page_indptr, *last_page_len, *positions;
  const int* request_indices;  // plain mode (k_new == nullptr): batch slot -> request, may be null (identity)
  const bf16 *qw, *kw, *cosc, *sinc;
  float eps, sm_scale_log2;
  int nq, nkv;
  int npf, pf_y;  // prefetch spans; extra cluster rows (blockIdx.y >= nkv) that only issue L2 prefetches
  PfSpan pf[kMaxPfSpans];
};
cudaError_t launch_decode_attention_cluster(const ClusterAttnArgs& a, int nkv, int bs, cudaStream_t stream);
// round-2 kernel (decode_attention_tma.cu): TMA page tiles + mma.sync, up to 16-CTA clusters.  Same argument block
// (prefetch spans ignored).  Returns a cudaError, or -2 when it cannot take the call.
int launch_decode_attention_tma(const ClusterAttnArgs& a, int nkv, int bs, cudaStream_t stream);

}  // namespace pk
