// C++ mirror of the reference's Rust host side for the Qwen3 hot path (no rustc in this image;
// the reference is compiled code, so the host side above the C ABI is C++).
//
//   tensor types        pegainfer-kernels/src/tensor.rs:12-232   (DeviceContext, DeviceVec,
//                                                                  DeviceMatrix, HiddenStates)
//   paging              pegainfer-core/src/{kv_pool.rs,page_pool.rs}
//   CUDA graph state    pegainfer-core/src/cuda_graph.rs:12-58
//   weights / TP shard  pegainfer-qwen3-4b/src/weights.rs:83-358, config.rs:114-158,
//                       pegainfer-core/src/weight_loader.rs:130-244
//   prefill DAG         pegainfer-qwen3-4b/src/prefill.rs:17-285 + ops/attention.rs:17-303
//   decode DAG          pegainfer-qwen3-4b/src/batch_decode.rs:17-295, batch_decode_buffers.rs
//
// The kernels are reached ONLY through the pegainfer-kernels C ABI (include/pegainfer_kernels.h),
// resolved with dlopen so the same host code drives either libpegainfer_kernels_b200.so or the
// reference's own kernels (oracle/_ref/libkernels_ref.so) for A/B runs.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "pegainfer_kernels.h"

namespace pq {

// ---------------------------------------------------------------- kernel library (ffi.rs)
struct KernelLib {
  void* handle = nullptr;
  std::string path;
#define PQ_FN(name) decltype(&::name) name = nullptr;
  PQ_FN(cuda_set_device) PQ_FN(cublas_init) PQ_FN(cublas_destroy)
  PQ_FN(embedding_batched_cuda) PQ_FN(rms_norm_cuda) PQ_FN(rms_norm_batched_cuda)
  PQ_FN(fused_add_rms_norm_batched_cuda) PQ_FN(add_cuda) PQ_FN(silu_mul_fused_cuda)
  PQ_FN(gemm_cuda) PQ_FN(gemm_graphsafe_cuda) PQ_FN(prefill_qk_norm_rope_only_cuda)
  PQ_FN(qk_norm_rope_batched_decode_cuda) PQ_FN(paged_kv_scatter_cuda)
  PQ_FN(batch_prefill_cta_tile_q_with_override) PQ_FN(batch_prefill_paged_cuda_with_cta_tile_q)
  PQ_FN(paged_attention_decode_cuda) PQ_FN(paged_attention_decode_split_kv_cuda)
  PQ_FN(flashinfer_top1_cuda)
  // B200 extensions (null when driving the reference's kernels)
  PQ_FN(pk_b200_launch_count) PQ_FN(pk_b200_set_pdl) PQ_FN(pk_b200_gemv_fused) PQ_FN(pk_b200_gemm_segments)
  PQ_FN(pk_b200_decode_attention_fused) PQ_FN(pk_tp_all_reduce_rows)
  PQ_FN(pk_tp_all_reduce_add_rms_norm) PQ_FN(pk_tp_max_rows)
  PQ_FN(pk_b200_decode_attention_fused_prefetch) PQ_FN(pk_b200_gemv_grid) PQ_FN(pk_tp_top1_exchange) PQ_FN(pk_b200_gemm_swiglu)
#undef PQ_FN
  bool has_extensions() const { return pk_b200_gemv_fused != nullptr; }
  std::string load(const std::string& p);  // returns error text, empty on success
  ~KernelLib();
};

// ---------------------------------------------------------------- tensor.rs
struct DeviceContext {
  int device = 0;
  cudaStream_t stream = nullptr;
};

struct DeviceBuf {  // owning CudaSlice<T>
  void* ptr = nullptr;
  size_t bytes = 0;
  DeviceBuf() = default;
  DeviceBuf(const DeviceBuf&) = delete;
  DeviceBuf& operator=(const DeviceBuf&) = delete;
  DeviceBuf(DeviceBuf&& o) noexcept { *this = std::move(o); }
  DeviceBuf& operator=(DeviceBuf&& o) noexcept;
  ~DeviceBuf();
  bool alloc_zeros(size_t nbytes);
  bool alloc_uninit(size_t nbytes);
  pk_bf16* bf() const { return static_cast<pk_bf16*>(ptr); }
  int* i32() const { return static_cast<int*>(ptr); }
};

struct DeviceVec {  // tensor.rs DeviceVec: 1-D bf16
  DeviceBuf data;
  size_t len = 0;
};
struct DeviceMatrix {  // row-major [rows, cols] bf16 (tensor.rs:136-141)
  DeviceBuf data;
  size_t rows = 0, cols = 0;
};
struct HiddenStates {  // [hidden_dim, seq_len]: token t at offset t*hidden_dim (tensor.rs:210-217)
  DeviceBuf data;
  size_t hidden_dim = 0, seq_len = 0;
  bool zeros(size_t dim, size_t tokens) {
    hidden_dim = dim;
    seq_len = tokens;
    return data.alloc_zeros(dim * tokens * 2);
  }
};

// ---------------------------------------------------------------- config.rs
struct Config {
  int hidden_size, intermediate_size, num_hidden_layers, num_attention_heads, num_key_value_heads,
      head_dim, vocab_size;
  float rms_norm_eps, rope_theta;
  bool tie_word_embeddings;
};
struct TensorParallelConfig {
  int rank = 0, world_size = 1;
  bool is_sharded() const { return world_size > 1; }
  void shard_range(int total, int* off, int* len) const {
    *len = total / world_size;
    *off = rank * *len;
  }
  std::string validate_for(const Config& c) const;
};

// ---------------------------------------------------------------- kv_pool.rs / page_pool.rs
struct KvLayout {
  int page_size, num_layers, num_kv_heads, head_dim;
  int64_t kv_block_len, layer_stride, page_stride;
  static KvLayout make(int num_layers, int num_kv_heads, int head_dim, int page_size);
};
struct PagePool {  // fixed-page allocator; ascending ids first (page_pool.rs:34-36)
  std::vector<int> free_list;
  int capacity = 0;
  void init(int n);
  bool acquire(int n, std::vector<int>* out);
  void release(const std::vector<int>& pages);
};
struct KvState {  // per-request pages + seq_len (kv_pool.rs:130-230)
  std::vector<int> pages;
  int seq_len = 0;
  bool live = false;
  int last_page_len(int page_size) const {
    if (seq_len == 0) return 0;
    int r = seq_len % page_size;
    return r == 0 ? page_size : r;
  }
};

// ---------------------------------------------------------------- cuda_graph.rs
struct CudaGraphState {
  cudaGraphExec_t exec = nullptr;
  bool captured() const { return exec != nullptr; }
  ~CudaGraphState();
};

// ---------------------------------------------------------------- weights.rs
struct Attention {
  DeviceMatrix qkv_proj, o_proj;
  DeviceVec q_norm, k_norm;
  int q_dim = 0, kv_dim = 0;
};
struct MLP {
  DeviceMatrix gate_up_proj, down_proj;
};
struct TransformerBlock {
  DeviceVec input_layernorm, post_attention_layernorm;
  Attention attention;
  MLP mlp;
};

struct RuntimeConfig {
  int device_ordinal = 0;
  int tp_rank = 0, tp_world = 1;
  int enable_cuda_graph = 1;
  int mode = 1;        // 0: reference op sequence through the ffi.rs ABI only; 1: fused B200 path
  int num_pages = 0;   // 0: 85 % of free memory (weights.rs:309-334), capped
  int max_batch = 4;
  int enable_pdl = 1;
};

}  // namespace pq
