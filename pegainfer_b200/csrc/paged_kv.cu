// Append K/V token rows into the page-first KV pool.  Replaces csrc/paged_attention.cu:274-311
// (FlashInfer page.cuh:259-284 AppendPagedKVCache) of the reference.
//
// Pool layout (pegainfer-core/src/kv_pool.rs:66-75): [page][layer][K | V][slot 0..15][kv head][128].
// Token i goes to entry e = page_indptr[batch_indices[i]] * page_size + positions[i]:
// page = page_indices[e / page_size], slot = e % page_size.  Pure 16-byte copies.
#include "common.cuh"

namespace pk {

__global__ void paged_kv_scatter_kernel(bf16* __restrict__ kv, int64_t k_off, int64_t v_off,
                                        const int* __restrict__ page_indices,
                                        const int* __restrict__ page_indptr,
                                        const bf16* __restrict__ src_k,
                                        const bf16* __restrict__ src_v,
                                        const int* __restrict__ batch_indices,
                                        const int* __restrict__ positions, int nnz, int nkv, int hd,
                                        int page_size, int64_t stride_page, int64_t src_stride_n,
                                        int64_t src_stride_h) {
  pdl_wait();
  const int vec_per_head = hd >> 3;
  const int64_t total = (int64_t)nnz * nkv * vec_per_head;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % vec_per_head);
    const int h = (int)((idx / vec_per_head) % nkv);
    const int i = (int)(idx / ((int64_t)vec_per_head * nkv));
    const int b = __ldg(batch_indices + i);
    const int entry = __ldg(page_indptr + b) * page_size + __ldg(positions + i);
    const int page = __ldg(page_indices + entry / page_size);
    const int slot = entry % page_size;
    const int64_t dst = (int64_t)page * stride_page + ((int64_t)slot * nkv + h) * hd + c * 8;
    const int64_t src = (int64_t)i * src_stride_n + (int64_t)h * src_stride_h + c * 8;
    *reinterpret_cast<uint4*>(kv + k_off + dst) = *reinterpret_cast<const uint4*>(src_k + src);
    *reinterpret_cast<uint4*>(kv + v_off + dst) = *reinterpret_cast<const uint4*>(src_v + src);
  }
}

}  // namespace pk

extern "C" int paged_kv_scatter_cuda(const pk_bf16* kv_data, int64_t k_offset_elems,
                                     int64_t v_offset_elems, const int* page_indices,
                                     const int* page_indptr, const int* last_page_len_d,
                                     const pk_bf16* src_k, const pk_bf16* src_v,
                                     const int* batch_indices, const int* positions, int nnz,
                                     int num_kv_heads, int head_dim, int page_size,
                                     int64_t stride_page, int64_t src_stride_n,
                                     int64_t src_stride_h, pk_stream stream) {
  (void)last_page_len_d;
  if (nnz <= 0) return 0;
  if (head_dim % 8 != 0 || src_stride_n % 8 != 0 || src_stride_h % 8 != 0 || stride_page % 8 != 0)
    return (int)cudaErrorInvalidValue;
  const int64_t total = (int64_t)nnz * num_kv_heads * (head_dim / 8);
  int64_t grid = (total + 255) / 256;
  const int64_t cap = (int64_t)pk::sm_count() * 8;
  if (grid > cap) grid = cap;
  pk::launch(pk::paged_kv_scatter_kernel, dim3((unsigned)grid), dim3(256), 0, stream, true,
             (pk::bf16*)kv_data, k_offset_elems, v_offset_elems, page_indices, page_indptr,
             (const pk::bf16*)src_k, (const pk::bf16*)src_v, batch_indices, positions, nnz,
             num_kv_heads, head_dim, page_size, stride_page, src_stride_n, src_stride_h);
  return (int)cudaGetLastError();
}
