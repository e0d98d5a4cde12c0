// See qwen3_host.hpp.  One Qwen3Model per process/GPU (one process per GPU under torchrun);
// all work on one stream, as the reference (tensor.rs:37-47).
#include "qwen3_host.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>

namespace pq {

// ===================================================================== KernelLib
std::string KernelLib::load(const std::string& p) {
  path = p;
  handle = dlopen(p.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!handle) return std::string("dlopen failed: ") + dlerror();
  std::string missing;
#define PQ_REQ(name)                                                   \
  name = reinterpret_cast<decltype(name)>(dlsym(handle, #name));       \
  if (!name) missing += std::string(" ") + #name;
#define PQ_OPT(name) name = reinterpret_cast<decltype(name)>(dlsym(handle, #name));
  PQ_REQ(cuda_set_device) PQ_REQ(cublas_init) PQ_REQ(cublas_destroy)
  PQ_REQ(embedding_batched_cuda) PQ_REQ(rms_norm_cuda) PQ_REQ(rms_norm_batched_cuda)
  PQ_REQ(fused_add_rms_norm_batched_cuda) PQ_REQ(add_cuda) PQ_REQ(silu_mul_fused_cuda)
  PQ_REQ(gemm_cuda) PQ_REQ(gemm_graphsafe_cuda) PQ_REQ(prefill_qk_norm_rope_only_cuda)
  PQ_REQ(qk_norm_rope_batched_decode_cuda) PQ_REQ(paged_kv_scatter_cuda)
  PQ_REQ(batch_prefill_cta_tile_q_with_override) PQ_REQ(batch_prefill_paged_cuda_with_cta_tile_q)
  PQ_REQ(paged_attention_decode_cuda) PQ_REQ(paged_attention_decode_split_kv_cuda)
  PQ_REQ(flashinfer_top1_cuda)
  PQ_OPT(pk_b200_launch_count) PQ_OPT(pk_b200_set_pdl) PQ_OPT(pk_b200_gemv_fused) PQ_OPT(pk_b200_gemm_segments)
  PQ_OPT(pk_b200_decode_attention_fused) PQ_OPT(pk_tp_all_reduce_rows)
  PQ_OPT(pk_tp_all_reduce_add_rms_norm) PQ_OPT(pk_tp_max_rows)
  PQ_OPT(pk_b200_decode_attention_fused_prefetch) PQ_OPT(pk_b200_gemv_grid) PQ_OPT(pk_tp_top1_exchange) PQ_OPT(pk_b200_gemm_swiglu)
#undef PQ_REQ
#undef PQ_OPT
  if (!missing.empty()) return "kernel library " + p + " lacks:" + missing;
  return "";
}
KernelLib::~KernelLib() {
  if (handle) dlclose(handle);
}

// ===================================================================== tensor.rs
DeviceBuf& DeviceBuf::operator=(DeviceBuf&& o) noexcept {
  if (this != &o) {
    if (ptr) cudaFree(ptr);
    ptr = o.ptr;
    bytes = o.bytes;
    o.ptr = nullptr;
    o.bytes = 0;
  }
  return *this;
}
DeviceBuf::~DeviceBuf() {
  if (ptr) cudaFree(ptr);
}
bool DeviceBuf::alloc_uninit(size_t nbytes) {
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  bytes = 0;
  if (nbytes == 0) nbytes = 16;
  if (cudaMalloc(&ptr, nbytes) != cudaSuccess) {
    ptr = nullptr;
    return false;
  }
  bytes = nbytes;
  return true;
}
bool DeviceBuf::alloc_zeros(size_t nbytes) {
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  bytes = 0;
  if (nbytes == 0) nbytes = 16;
  if (cudaMalloc(&ptr, nbytes) != cudaSuccess) {
    ptr = nullptr;
    return false;
  }
  bytes = nbytes;
  return cudaMemset(ptr, 0, nbytes) == cudaSuccess;
}

std::string TensorParallelConfig::validate_for(const Config& c) const {
  char b[160];
  if (world_size < 1) return "tensor_parallel.world_size must be >= 1";
  if (rank < 0 || rank >= world_size) {
    snprintf(b, sizeof b, "tensor_parallel.rank %d must be < world_size %d", rank, world_size);
    return b;
  }
  const std::pair<const char*, int> chk[] = {{"num_attention_heads", c.num_attention_heads},
                                             {"num_key_value_heads", c.num_key_value_heads},
                                             {"intermediate_size", c.intermediate_size}};
  for (auto& kv : chk)
    if (kv.second % world_size) {
      snprintf(b, sizeof b, "%s=%d not divisible by tp world_size=%d", kv.first, kv.second, world_size);
      return b;
    }
  return "";
}

KvLayout KvLayout::make(int num_layers, int num_kv_heads, int head_dim, int page_size) {
  KvLayout l;
  l.page_size = page_size;
  l.num_layers = num_layers;
  l.num_kv_heads = num_kv_heads;
  l.head_dim = head_dim;
  l.kv_block_len = (int64_t)page_size * num_kv_heads * head_dim;
  l.layer_stride = 2 * l.kv_block_len;
  l.page_stride = (int64_t)num_layers * l.layer_stride;
  return l;
}
void PagePool::init(int n) {
  capacity = n;
  free_list.clear();
  for (int i = n - 1; i >= 0; --i) free_list.push_back(i);  // pop_back hands out ascending ids
}
bool PagePool::acquire(int n, std::vector<int>* out) {
  if ((int)free_list.size() < n) return false;
  for (int i = 0; i < n; ++i) {
    out->push_back(free_list.back());
    free_list.pop_back();
  }
  return true;
}
void PagePool::release(const std::vector<int>& pages) {
  for (int p : pages) free_list.push_back(p);
}
CudaGraphState::~CudaGraphState() {
  if (exec) cudaGraphExecDestroy(exec);
}

// ===================================================================== model
static const int kBuckets[] = {1, 2, 4, 8, 16, 32, 64};  // batch_decode_buffers.rs:12
static const int kSplitChunkTokens = 256, kSplitMaxChunks = 64, kSplitMaxBs = 2, kSplitMinSeq = 1024;
static const int kPrefillCtaTileQ = 64;  // config.rs:5
static const int kRopePositions = 4096;  // weights.rs:300
static const int kPageSize = 16;         // weights.rs:309

static int bucket_for(int bs);

struct Qwen3Model {
  KernelLib k;
  DeviceContext ctx;
  Config config{};
  RuntimeConfig rt{};
  TensorParallelConfig tp{};
  pk_tp_comm* tp_comm = nullptr;
  std::string err;

  DeviceMatrix embed_tokens, lm_head;
  bool has_lm_head = false;
  std::vector<TransformerBlock> layers;
  DeviceVec norm, cos_cache, sin_cache;
  bool finalized = false;

  // KvPool
  KvLayout layout{};
  DeviceBuf kv_buffer;
  PagePool pool;
  int padding_page = 0;
  std::vector<KvState> kv_states;

  // ---- BatchDecodeBuffers (batch_decode_buffers.rs:51-171) ----
  int max_bs = 0;
  HiddenStates normed, q, kbuf, v, attn_out, attn_proj, gate_up_out, mlp_act, mlp_out, hidden, hidden_b,
      logits;
  DeviceBuf zero_residual;
  DeviceBuf meta_d;  // packed per-step metadata, one H2D copy
  void* meta_h = nullptr;
  size_t meta_bytes = 0;
  int max_total_pages = 0;
  // offsets (in ints) into the packed metadata
  struct MetaOff {
    int token_ids, positions, page_indptr, last_page_len, request_indices, kv_tile_indices,
        kv_chunk_size, split_request, split_tile, split_chunk, split_o_indptr, page_indices,
        split_mask_bytes, step_seq, total_ints;
  } mo{};
  DeviceBuf split_tmp_v, split_tmp_s;
  DeviceBuf attn_partial, attn_counters;
  int attn_max_chunks = 1;
  DeviceBuf sample_out, top1_val, top1_states;
  int* sample_h = nullptr;
  std::map<int, std::unique_ptr<CudaGraphState>> graphs;
  int max_seq_len_step = 0, split_padded_slots = 0;
  uint32_t step_counter = 0;  // decode steps issued so far: the sequence base of the GEMV-fused all-reduces (every rank
                              // issues the same steps in the same order, so the counters agree across ranks)

  int local_heads() const { return config.num_attention_heads / tp.world_size; }
  int local_kv_heads() const { return config.num_key_value_heads / tp.world_size; }
  int local_inter() const { return config.intermediate_size / tp.world_size; }
  int local_q_dim() const { return local_heads() * config.head_dim; }
  int local_kv_dim() const { return local_kv_heads() * config.head_dim; }
  // Vocab-sharded lm_head: on the fused TP path every rank streams only rows [rank * V/N, (rank+1) * V/N)
  // of the output projection and the greedy token is agreed on with a (max, index) exchange; the reference keeps
  // lm_head replicated (weights.rs:104-119), which is what mode 0 (its op sequence) still does.
  bool vocab_sharded() const {
    return tp.is_sharded() && rt.mode >= 1 && tp_comm && k.pk_tp_top1_exchange && config.vocab_size % tp.world_size == 0 &&
           (config.vocab_size / tp.world_size) % 8 == 0;
  }
  int local_vocab() const { return vocab_sharded() ? config.vocab_size / tp.world_size : config.vocab_size; }
  int vocab_offset() const { return vocab_sharded() ? tp.rank * (config.vocab_size / tp.world_size) : 0; }
  const DeviceMatrix& output_projection_full() const { return has_lm_head ? lm_head : embed_tokens; }
  // this rank's rows of the output projection: the whole matrix, or its vocabulary shard
  const pk_bf16* output_rows() const {
    const DeviceMatrix& m = output_projection_full();
    return (vocab_sharded() && !lm_head_is_shard) ? m.data.bf() + (size_t)vocab_offset() * m.cols : m.data.bf();
  }
  bool lm_head_is_shard = false;  // untied lm_head uploaded as the local shard only
  uint32_t host_sample_seq = 0;

  bool fail(const std::string& m) {
    err = m;
    return false;
  }
  bool cu(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
  }

  // ---- ops (pegainfer-kernels/src/ops/linear.rs) ----
  void gemm_rows_into(const DeviceMatrix& w, size_t row_off, size_t nrows, const pk_bf16* x, int T,
                      pk_bf16* out) {
    const pk_bf16* wsub = w.data.bf() + row_off * w.cols;
    if (T == 1)
      k.gemm_graphsafe_cuda(wsub, x, out, (int)nrows, 1, (int)w.cols, ctx.stream);
    else
      k.gemm_cuda(wsub, x, out, (int)nrows, T, (int)w.cols, ctx.stream);
  }
  // decode: every GEMM of the graph goes through the graph-safe entry (ops/linear.rs:27-36 for
  // seq_len == 1; wider buckets take the same entry here because our gemm paths are capture-safe)
  void gemm_decode(const DeviceMatrix& w, size_t row_off, size_t nrows, const pk_bf16* x, int bs,
                   pk_bf16* out) {
    k.gemm_graphsafe_cuda(w.data.bf() + row_off * w.cols, x, out, (int)nrows, bs, (int)w.cols,
                          ctx.stream);
  }

  bool all_reduce_hidden(pk_bf16* h, int dim, int T) {  // weights.rs:396-405
    if (!tp.is_sharded()) return true;
    if (!tp_comm || !k.pk_tp_all_reduce_rows) return fail("tensor parallel without a TP communicator");
    const int64_t max_rows = k.pk_tp_max_rows(tp_comm, dim);
    for (int t0 = 0; t0 < T; t0 += (int)max_rows) {
      const int n = std::min<int64_t>(max_rows, T - t0);
      if (k.pk_tp_all_reduce_rows(tp_comm, h + (size_t)t0 * dim, dim, n, ctx.stream) != 0)
        return fail("pk_tp_all_reduce_rows failed");
    }
    return true;
  }

  bool load_tensor(const std::string& name, const void* data, int rows, int cols);
  bool finalize();
  bool create_decode_buffers();
  bool prefill(int n_req, const uint32_t* tokens, const int* lens, const int* kv_ids, void** logits_out);
  bool decode(int bs, const uint32_t* tokens, const int* kv_ids, void** logits_out, int* sampled);
  bool build_step_meta(int bs, int padded, const uint32_t* tokens, const int* kv_ids, int* mh, bool* split_out);
  bool run_step_kernels(int padded, bool split);
  bool decode_burst(int kv_id, uint32_t first_token, int K, uint32_t* tokens_out, float* ms_total);
  bool bench_gemv_pass(int iters, float* ms_per_pass, int* launches_per_pass);
  int64_t launches_per_step = 0;
  bool decode_kernels_compat(int bs, bool split);
  bool decode_kernels_fused(int bs);
  bool decode_kernels_wide(int bs);
  bool unified_step(int n_prefill, const uint32_t* prompt_tokens, const int* lens, const int* prefill_kv_ids, int n_decode,
                    const uint32_t* decode_tokens, const int* decode_kv_ids, void** prefill_logits, void** decode_logits);
  HiddenStates pf_hid, pf_hid_out, pf_nrm, pf_q, pf_k, pf_v, pf_o, pf_gu, pf_act, pf_att;
  DeviceBuf pf_plan;
  std::vector<int> plan_pack;
  int prefill_capacity = 0;
  bool ensure_capacity(KvState& s, int tokens);
  bool sample_greedy(const pk_bf16* logits, int* out);
  ~Qwen3Model();
};

Qwen3Model::~Qwen3Model() {
  if (ctx.stream) cudaStreamSynchronize(ctx.stream);
  graphs.clear();
  if (meta_h) cudaFreeHost(meta_h);
  if (sample_h) cudaFreeHost(sample_h);
  if (k.cublas_destroy) k.cublas_destroy();
  if (ctx.stream) cudaStreamDestroy(ctx.stream);
}

// ---- weight upload with TP sharding (weights.rs:121-291, weight_loader.rs:130-206) ----
static bool upload_rows(DeviceMatrix& dst, size_t dst_row, const void* src, int src_cols, int row_off,
                        int rows) {
  const char* s = static_cast<const char*>(src) + (size_t)row_off * src_cols * 2;
  return cudaMemcpy(dst.data.bf() + dst_row * dst.cols, s, (size_t)rows * src_cols * 2,
                    cudaMemcpyDefault) == cudaSuccess;
}
static bool upload_cols(DeviceMatrix& dst, const void* src, int src_rows, int src_cols, int col_off,
                        int cols) {
  const char* s = static_cast<const char*>(src) + (size_t)col_off * 2;
  return cudaMemcpy2D(dst.data.ptr, (size_t)cols * 2, s, (size_t)src_cols * 2, (size_t)cols * 2, src_rows,
                      cudaMemcpyDefault) == cudaSuccess;
}
static bool alloc_matrix(DeviceMatrix& m, size_t rows, size_t cols) {
  if (m.data.ptr && m.rows == rows && m.cols == cols) return true;
  m.rows = rows;
  m.cols = cols;
  return m.data.alloc_zeros(rows * cols * 2);
}
static bool upload_vec(DeviceVec& v, const void* src, size_t n) {
  v.len = n;
  if (!v.data.alloc_zeros(n * 2)) return false;
  return cudaMemcpy(v.data.ptr, src, n * 2, cudaMemcpyDefault) == cudaSuccess;
}

bool Qwen3Model::load_tensor(const std::string& name, const void* data, int rows, int cols) {
  const Config& c = config;
  const int H = c.hidden_size;
  if ((int)layers.size() != c.num_hidden_layers) layers.resize(c.num_hidden_layers);
  int q_off, q_rows, kv_off, kv_rows, i_off, i_rows;
  tp.shard_range(c.num_attention_heads * c.head_dim, &q_off, &q_rows);
  tp.shard_range(c.num_key_value_heads * c.head_dim, &kv_off, &kv_rows);
  tp.shard_range(c.intermediate_size, &i_off, &i_rows);
  bool ok = true;
  // full-shape check for every tensor (weight_loader.rs:130-206 asserts shapes; a mismatched checkpoint must fail
  // with the tensor name instead of reading out of bounds)
  auto want = [&](int er, int ec) -> bool {
    if (rows == er && cols == ec) return true;
    char b[200];
    snprintf(b, sizeof b, "%s: shape [%d, %d] does not match the expected [%d, %d]", name.c_str(), rows, cols, er, ec);
    fail(b);
    return false;
  };
  const int Q = c.num_attention_heads * c.head_dim, KV = c.num_key_value_heads * c.head_dim;
  if (name == "model.embed_tokens.weight") {
    if (!want(c.vocab_size, H)) return false;
    ok = alloc_matrix(embed_tokens, rows, cols) && upload_rows(embed_tokens, 0, data, cols, 0, rows);
  } else if (name == "lm_head.weight") {
    if (c.tie_word_embeddings) return true;  // tied: weights.rs:104-107
    if (!want(c.vocab_size, H)) return false;
    has_lm_head = true;
    if (vocab_sharded()) {  // keep only this rank's vocabulary rows
      lm_head_is_shard = true;
      ok = alloc_matrix(lm_head, local_vocab(), cols) && upload_rows(lm_head, 0, data, cols, vocab_offset(), local_vocab());
    } else {
      ok = alloc_matrix(lm_head, rows, cols) && upload_rows(lm_head, 0, data, cols, 0, rows);
    }
  } else if (name == "model.norm.weight") {
    if (!want(1, H)) return false;
    ok = upload_vec(norm, data, H);
  } else if (name.rfind("model.layers.", 0) == 0) {
    const size_t dot = name.find('.', 13);
    if (dot == std::string::npos) return fail("unknown tensor " + name);
    const int li = atoi(name.substr(13, dot - 13).c_str());
    if (li < 0 || li >= c.num_hidden_layers) return fail("layer index out of range: " + name);
    TransformerBlock& L = layers[li];
    const std::string sub = name.substr(dot + 1);
    Attention& A = L.attention;
    A.q_dim = q_rows;
    A.kv_dim = kv_rows;
    if (sub == "input_layernorm.weight") ok = want(1, H) && upload_vec(L.input_layernorm, data, H);
    else if (sub == "post_attention_layernorm.weight") ok = want(1, H) && upload_vec(L.post_attention_layernorm, data, H);
    else if (sub == "self_attn.q_norm.weight") ok = want(1, c.head_dim) && upload_vec(A.q_norm, data, c.head_dim);
    else if (sub == "self_attn.k_norm.weight") ok = want(1, c.head_dim) && upload_vec(A.k_norm, data, c.head_dim);
    else if (sub == "self_attn.q_proj.weight" || sub == "self_attn.k_proj.weight" ||
             sub == "self_attn.v_proj.weight") {
      // DeviceMatrix::vstack([q, k, v]) (weights.rs:182): written straight into the fused matrix
      if (!want(sub[10] == 'q' ? Q : KV, H)) return false;
      ok = alloc_matrix(A.qkv_proj, (size_t)q_rows + 2 * kv_rows, H);
      if (sub[10] == 'q') ok = ok && upload_rows(A.qkv_proj, 0, data, cols, q_off, q_rows);
      else if (sub[10] == 'k') ok = ok && upload_rows(A.qkv_proj, q_rows, data, cols, kv_off, kv_rows);
      else ok = ok && upload_rows(A.qkv_proj, (size_t)q_rows + kv_rows, data, cols, kv_off, kv_rows);
    } else if (sub == "self_attn.o_proj.weight") {  // column shard (weight_loader.rs:168-206)
      if (!want(H, Q)) return false;
      ok = alloc_matrix(A.o_proj, H, q_rows) && upload_cols(A.o_proj, data, rows, cols, q_off, q_rows);
    } else if (sub == "mlp.gate_proj.weight" || sub == "mlp.up_proj.weight") {
      if (!want(c.intermediate_size, H)) return false;
      ok = alloc_matrix(L.mlp.gate_up_proj, (size_t)2 * i_rows, H);
      ok = ok && upload_rows(L.mlp.gate_up_proj, sub[4] == 'g' ? 0 : i_rows, data, cols, i_off, i_rows);
    } else if (sub == "mlp.down_proj.weight") {
      if (!want(H, c.intermediate_size)) return false;
      ok = alloc_matrix(L.mlp.down_proj, H, i_rows) && upload_cols(L.mlp.down_proj, data, rows, cols, i_off, i_rows);
    } else {
      return fail("unknown tensor " + name);
    }
    if (!ok && !err.empty() && err.find("does not match") != std::string::npos) return false;
  } else {
    return fail("unknown tensor " + name);
  }
  if (!ok) return fail("upload failed for " + name + ": " + cudaGetErrorString(cudaGetLastError()));
  return true;
}

static uint16_t f2bf_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fff;
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

bool Qwen3Model::finalize() {
  const Config& c = config;
  // precompute_rope (weight_loader.rs:210-244): f32 powf/cos/sin, bf16, half-split duplicated
  {
    const int hd = c.head_dim, half = hd / 2;
    std::vector<uint16_t> cs((size_t)kRopePositions * hd), sn((size_t)kRopePositions * hd);
    std::vector<float> inv(half);
    for (int i = 0; i < half; ++i) inv[i] = 1.0f / powf(c.rope_theta, (float)i * 2.0f / (float)hd);
    for (int pos = 0; pos < kRopePositions; ++pos)
      for (int i = 0; i < half; ++i) {
        const float fr = (float)pos * inv[i];
        const uint16_t cv = f2bf_host(cosf(fr)), sv = f2bf_host(sinf(fr));
        cs[(size_t)pos * hd + i] = cs[(size_t)pos * hd + i + half] = cv;
        sn[(size_t)pos * hd + i] = sn[(size_t)pos * hd + i + half] = sv;
      }
    if (!upload_vec(cos_cache, cs.data(), cs.size()) || !upload_vec(sin_cache, sn.data(), sn.size()))
      return fail("rope upload failed");
  }
  for (int i = 0; i < c.num_hidden_layers; ++i) {
    const TransformerBlock& L = layers[i];
    if (!L.attention.qkv_proj.data.ptr || !L.attention.o_proj.data.ptr || !L.mlp.gate_up_proj.data.ptr ||
        !L.mlp.down_proj.data.ptr || !L.input_layernorm.data.ptr || !L.post_attention_layernorm.data.ptr ||
        !L.attention.q_norm.data.ptr || !L.attention.k_norm.data.ptr)
      return fail("layer " + std::to_string(i) + " is missing tensors");
  }
  if (!embed_tokens.data.ptr || !norm.data.ptr) return fail("embed_tokens / norm missing");
  if (!c.tie_word_embeddings && !has_lm_head) return fail("lm_head.weight missing for an untied model");

  // KV pool: 85 % of free memory in 16-token pages (weights.rs:309-334) unless told otherwise
  layout = KvLayout::make(c.num_hidden_layers, local_kv_heads(), c.head_dim, kPageSize);
  int num_pages = rt.num_pages;
  if (num_pages <= 0) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    num_pages = (int)std::max<size_t>(64, (size_t)((double)free_b * 0.85) / ((size_t)layout.page_stride * 2));
  }
  if (!kv_buffer.alloc_zeros((size_t)num_pages * layout.page_stride * 2)) return fail("KvPool alloc failed");
  pool.init(num_pages);
  std::vector<int> pad;
  pool.acquire(1, &pad);  // page 0 reserved as the graph-padding page (kv_pool.rs:99-102)
  padding_page = pad[0];
  if (!create_decode_buffers()) return false;
  finalized = true;
  return cu(cudaStreamSynchronize(ctx.stream), "finalize sync");
}

bool Qwen3Model::create_decode_buffers() {
  const Config& c = config;
  max_bs = std::max(1, std::min(rt.max_batch, 64));
  // graph buckets pad the batch up to {1,2,4,...,64} (batch_decode_buffers.rs:12): the buffers must hold the bucket
  if (rt.enable_cuda_graph && bucket_for(max_bs) > 0) max_bs = bucket_for(max_bs);
  const int bs = max_bs, H = c.hidden_size, qd = local_q_dim(), kd = local_kv_dim();
  bool ok = normed.zeros(H, bs) && q.zeros(qd, bs) && kbuf.zeros(kd, bs) && v.zeros(kd, bs) &&
            attn_out.zeros(qd, bs) && attn_proj.zeros(H, bs) && gate_up_out.zeros(2 * local_inter(), bs) &&
            mlp_act.zeros(local_inter(), bs) && mlp_out.zeros(H, bs) && hidden.zeros(H, bs) &&
            hidden_b.zeros(H, bs) && logits.zeros(local_vocab(), bs) && zero_residual.alloc_zeros((size_t)H * bs * 2);
  // a request never holds more than 4096 / 16 pages (RoPE table bound), so this covers any legal batch
  max_total_pages = max_bs * (kRopePositions / kPageSize);
  const int slots = bs * kSplitMaxChunks;
  int o = 0;
  mo.token_ids = o; o += bs;
  mo.positions = o; o += bs;
  mo.page_indptr = o; o += bs + 1;
  mo.last_page_len = o; o += bs;
  mo.request_indices = o; o += bs;
  mo.kv_tile_indices = o; o += bs;
  mo.kv_chunk_size = o; o += bs;
  mo.split_request = o; o += slots;
  mo.split_tile = o; o += slots;
  mo.split_chunk = o; o += 1;
  mo.step_seq = o; o += 1;
  mo.split_o_indptr = o; o += bs + 1;
  mo.split_mask_bytes = o; o += (slots + 3) / 4;
  mo.page_indices = o; o += max_total_pages + bs;
  mo.total_ints = o;
  meta_bytes = (size_t)o * 4;
  ok = ok && meta_d.alloc_zeros(meta_bytes) && cudaMallocHost(&meta_h, meta_bytes) == cudaSuccess;
  ok = ok && split_tmp_v.alloc_zeros((size_t)slots * qd * 2) &&
       split_tmp_s.alloc_zeros((size_t)slots * local_heads() * 4);
  // fused attention scratch: ~2 CTAs per SM across (chunks, kv heads, requests)
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx.device);
  attn_max_chunks = std::max(1, std::min(64, (2 * sms + local_kv_heads() - 1) / local_kv_heads()));
  ok = ok && attn_partial.alloc_zeros((size_t)bs * attn_max_chunks * local_heads() * 130 * 4) &&
       attn_counters.alloc_zeros((size_t)bs * local_kv_heads() * 4 + 64);
  ok = ok && sample_out.alloc_zeros(64 * 4) && top1_val.alloc_zeros(256) && top1_states.alloc_zeros(1 << 20) &&
       cudaMallocHost((void**)&sample_h, 64 * 4) == cudaSuccess;
  if (!ok) return fail(std::string("decode buffer allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
  return true;
}

bool Qwen3Model::ensure_capacity(KvState& s, int tokens) {
  const int need = (tokens + kPageSize - 1) / kPageSize;
  const int held = (int)s.pages.size();
  if (need <= held) return true;
  if (!pool.acquire(need - held, &s.pages)) {
    char b[128];
    snprintf(b, sizeof b, "KvState: out of pages (need %d more, %d available)", need - held,
             (int)pool.free_list.size());
    return fail(b);
  }
  return true;
}

// ===================================================================== prefill (prefill.rs)
bool Qwen3Model::prefill(int n_req, const uint32_t* tokens, const int* lens, const int* kv_ids,
                         void** logits_out) {
  return unified_step(n_req, tokens, lens, kv_ids, 0, nullptr, nullptr, logits_out, nullptr);
}

// Unified step (unified_forward.rs:78-567): prefill prompts and decode tokens in ONE forward pass.  Token order as the
// reference: all prompt tokens, then one token per decode request.  Projections / MLP / norms run over all tokens
// together; attention splits -- prompt rows take QK-norm + RoPE, KV scatter and the paged prefill kernel, decode rows
// the fused decode-attention launch (or, on a kernel library without the extensions, the reference's three launches).
// With n_decode == 0 this IS batch_prefill (prefill.rs:220-285).  Returns last-token logits per prompt and per decode row.
bool Qwen3Model::unified_step(int n_req, const uint32_t* tokens, const int* lens, const int* kv_ids, int n_dec,
                              const uint32_t* dec_tokens, const int* dec_kv_ids, void** logits_out, void** dec_logits_out) {
  const Config& c = config;
  const int H = c.hidden_size, qd = local_q_dim(), kd = local_kv_dim(), I = local_inter();
  const int nh = local_heads(), nkv = local_kv_heads(), hd = c.head_dim;
  // ---- plan, then commit (the reference validates the whole batch before any KvState changes): every check that
  // can reject the call runs first; seq_len / pages are only touched once nothing below can fail for a
  // data-dependent reason, and a CUDA failure later rolls them back (PrefillTxn). ----
  int T = 0;
  std::vector<int> starts(n_req);
  if (n_req <= 0) return fail("empty prefill batch");
  if (n_req + n_dec > max_bs) return fail("more prompts + decode requests than max_batch");
  int pages_needed = 0;
  for (int i = 0; i < n_dec; ++i) {
    if (dec_kv_ids[i] < 0 || dec_kv_ids[i] >= (int)kv_states.size() || !kv_states[dec_kv_ids[i]].live) return fail("bad kv id");
    for (int j = 0; j < i; ++j)
      if (dec_kv_ids[j] == dec_kv_ids[i]) return fail("duplicate kv id in one unified step");
    for (int j = 0; j < n_req; ++j)
      if (kv_ids[j] == dec_kv_ids[i]) return fail("a request cannot be prefilled and decoded in the same step");
    const KvState& s = kv_states[dec_kv_ids[i]];
    if (s.seq_len + 1 > kRopePositions) return fail("position beyond the 4096-entry RoPE table");
    pages_needed += std::max(0, (s.seq_len + 1 + kPageSize - 1) / kPageSize - (int)s.pages.size());
  }
  for (int i = 0; i < n_req; ++i) {
    if (kv_ids[i] < 0 || kv_ids[i] >= (int)kv_states.size() || !kv_states[kv_ids[i]].live)
      return fail("bad kv id");
    for (int j = 0; j < i; ++j)
      if (kv_ids[j] == kv_ids[i]) return fail("duplicate kv id in one prefill batch");
    if (lens[i] <= 0) return fail("empty prompt");
    const KvState& s = kv_states[kv_ids[i]];
    starts[i] = s.seq_len;
    if (starts[i] + lens[i] > kRopePositions) return fail("position beyond the 4096-entry RoPE table");
    pages_needed += std::max(0, (starts[i] + lens[i] + kPageSize - 1) / kPageSize - (int)s.pages.size());
    T += lens[i];
  }
  if (pages_needed > (int)pool.free_list.size()) {
    char b[128];
    snprintf(b, sizeof b, "KvState: out of pages (need %d more, %d available)", pages_needed,
             (int)pool.free_list.size());
    return fail(b);
  }
  struct PrefillTxn {  // undo record: restores seq_len and returns the pages acquired by this call
    Qwen3Model* m;
    std::vector<std::pair<int, std::pair<int, size_t>>> old;  // kv id -> (seq_len, page count)
    bool committed = false;
    ~PrefillTxn() {
      if (committed) return;
      for (auto& o : old) {
        KvState& s = m->kv_states[o.first];
        std::vector<int> extra(s.pages.begin() + o.second.second, s.pages.end());
        m->pool.release(extra);
        s.pages.resize(o.second.second);
        s.seq_len = o.second.first;
      }
    }
  } txn{this, {}, false};
  for (int i = 0; i < n_req; ++i) {
    KvState& s = kv_states[kv_ids[i]];
    txn.old.push_back({kv_ids[i], {s.seq_len, s.pages.size()}});
    if (!ensure_capacity(s, starts[i] + lens[i])) return false;  // cannot fail: counted above
    s.seq_len += lens[i];
  }
  const int Tp = T;  // prompt tokens; decode rows follow
  std::vector<int> dec_pos(n_dec);
  for (int i = 0; i < n_dec; ++i) {
    KvState& s = kv_states[dec_kv_ids[i]];
    txn.old.push_back({dec_kv_ids[i], {s.seq_len, s.pages.size()}});
    dec_pos[i] = s.seq_len;
    if (!ensure_capacity(s, s.seq_len + 1)) return false;
    s.seq_len += 1;
  }
  T += n_dec;
  // ---- PrefillPagedPlan::new_batch_with_cta_tile_q (ops/attention.rs:208-302) ----
  std::vector<int> page_indices, page_indptr{0}, last_page_len, kv_chunk, batch_indices, positions, q_indptr{0},
      req_idx, qo_tile, kv_tile;
  const int group = nh / nkv;
  const int cta_tile_q = k.batch_prefill_cta_tile_q_with_override(Tp, nh, nkv, hd, kPrefillCtaTileQ);
  if (cta_tile_q <= 0) return fail("invalid prefill CTA tile override");
  for (int i = 0; i < n_req; ++i) {
    const KvState& s = kv_states[kv_ids[i]];
    page_indices.insert(page_indices.end(), s.pages.begin(), s.pages.end());
    page_indptr.push_back((int)page_indices.size());
    last_page_len.push_back(s.last_page_len(kPageSize));
    kv_chunk.push_back(starts[i] + lens[i]);
    for (int t = 0; t < lens[i]; ++t) {
      batch_indices.push_back(i);
      positions.push_back(starts[i] + t);
    }
    q_indptr.push_back(q_indptr.back() + lens[i]);
    const int nt = (lens[i] * group + cta_tile_q - 1) / cta_tile_q;
    for (int t = 0; t < nt; ++t) {
      req_idx.push_back(i);
      qo_tile.push_back(t);
      kv_tile.push_back(0);
    }
  }
  const int num_tiles = (int)req_idx.size();
  // one packed upload instead of the reference's 11 clone_htod calls
  std::vector<int> pack;
  auto put = [&](const std::vector<int>& v) {
    const int off = (int)pack.size();
    pack.insert(pack.end(), v.begin(), v.end());
    while (pack.size() % 4) pack.push_back(0);
    return off;
  };
  const int o_pi = put(page_indices), o_ip = put(page_indptr), o_lpl = put(last_page_len), o_kc = put(kv_chunk),
            o_bi = put(batch_indices), o_pos = put(positions), o_qi = put(q_indptr), o_ri = put(req_idx),
            o_qt = put(qo_tile), o_kt = put(kv_tile), o_tnr = put(std::vector<int>{Tp});
  // decode rows of the step: CSR page table, last_page_len, positions (batch_decode.rs:26-59 semantics)
  std::vector<int> d_pi, d_ip{0}, d_lpl, d_req, d_tile, d_chunk;
  for (int i = 0; i < n_dec; ++i) {
    const KvState& s = kv_states[dec_kv_ids[i]];
    d_pi.insert(d_pi.end(), s.pages.begin(), s.pages.end());
    d_ip.push_back((int)d_pi.size());
    d_lpl.push_back(s.last_page_len(kPageSize));
    d_req.push_back(i);
    d_tile.push_back(0);
    d_chunk.push_back(s.seq_len);
  }
  const int o_dpi = put(d_pi), o_dip = put(d_ip), o_dlpl = put(d_lpl), o_dpos = put(dec_pos), o_dreq = put(d_req),
            o_dtile = put(d_tile), o_dchunk = put(d_chunk);
  const int o_tok = (int)pack.size();
  pack.insert(pack.end(), reinterpret_cast<const int*>(tokens), reinterpret_cast<const int*>(tokens) + Tp);
  if (n_dec > 0) pack.insert(pack.end(), reinterpret_cast<const int*>(dec_tokens), reinterpret_cast<const int*>(dec_tokens) + n_dec);
  plan_pack.swap(pack);  // keep the host block alive until the async copy has been consumed

  // ---- PrefillBuffers (prefill.rs:17-51).  The reference re-allocates nine buffers per prefill call
  // (its doc comment cites the cuMemAlloc cost in TTFT); here they are a grow-only arena owned by the
  // model: every element is overwritten before it is read, so no zero fill either. ----
  if (T > prefill_capacity) {
    cudaStreamSynchronize(ctx.stream);
    const int cap = ((T + 255) / 256) * 256;
    auto grow = [&](HiddenStates& hs, size_t dim) {
      hs.hidden_dim = dim;
      hs.seq_len = cap;
      return hs.data.alloc_uninit(dim * cap * 2);
    };
    if (!(grow(pf_hid, H) && grow(pf_hid_out, H) && grow(pf_nrm, H) && grow(pf_q, qd) && grow(pf_k, kd) &&
          grow(pf_v, kd) && grow(pf_o, H) && grow(pf_gu, 2 * I) && grow(pf_act, I) && grow(pf_att, qd)))
      return fail("prefill buffer allocation failed");
    prefill_capacity = cap;
  }
  if (plan_pack.size() * 4 > pf_plan.bytes) {
    cudaStreamSynchronize(ctx.stream);
    if (!pf_plan.alloc_uninit(plan_pack.size() * 4 + 65536)) return fail("prefill plan buffer allocation failed");
  }
  HiddenStates &hid = pf_hid, &hid_out = pf_hid_out, &nrm = pf_nrm, &qb = pf_q, &kb = pf_k, &vb = pf_v, &ob = pf_o,
               &gu = pf_gu, &act = pf_act, &att = pf_att;
  cudaStream_t st = ctx.stream;
  if (!cu(cudaMemcpyAsync(pf_plan.ptr, plan_pack.data(), plan_pack.size() * 4, cudaMemcpyHostToDevice, st), "plan H2D"))
    return false;
  const int* P = pf_plan.i32();
  const float eps = c.rms_norm_eps;
  const float sm_scale = 1.0f / sqrtf((float)hd);
  k.embedding_batched_cuda(embed_tokens.data.bf(), reinterpret_cast<const uint32_t*>(P + o_tok), hid.data.bf(), H, T, st);
  pk_bf16* hcur = hid.data.bf();
  pk_bf16* hnext = hid_out.data.bf();
  for (int li = 0; li < c.num_hidden_layers; ++li) {  // forward_layer_batch_paged (prefill.rs:73-188)
    TransformerBlock& L = layers[li];
    k.rms_norm_batched_cuda(hcur, L.input_layernorm.data.bf(), nrm.data.bf(), H, T, eps, st);
    if (rt.mode >= 1 && k.pk_b200_gemm_segments && T > 1) {
      // one tensor-core launch over the stacked [q;k;v] weight, three outputs (same arithmetic)
      pk_bf16* outs[3] = {qb.data.bf(), kb.data.bf(), vb.data.bf()};
      const int segs[3] = {qd, kd, kd};
      if (k.pk_b200_gemm_segments(L.attention.qkv_proj.data.bf(), nrm.data.bf(), outs, segs, qd + 2 * kd, T, H, st) != 0)
        return fail("pk_b200_gemm_segments failed");
    } else {
      gemm_rows_into(L.attention.qkv_proj, 0, qd, nrm.data.bf(), T, qb.data.bf());
      gemm_rows_into(L.attention.qkv_proj, qd, kd, nrm.data.bf(), T, kb.data.bf());
      gemm_rows_into(L.attention.qkv_proj, qd + kd, kd, nrm.data.bf(), T, vb.data.bf());
    }
    // prompt rows [0, Tp): prefill_attention_paged_into (ops/attention.rs:310-458)
    if (n_req == 1)
      k.prefill_qk_norm_rope_only_cuda(qb.data.bf(), kb.data.bf(), L.attention.q_norm.data.bf(),
                                       L.attention.k_norm.data.bf(), cos_cache.data.bf(), sin_cache.data.bf(), nh,
                                       nkv, hd, Tp, starts[0], eps, st);
    else
      k.qk_norm_rope_batched_decode_cuda(qb.data.bf(), kb.data.bf(), L.attention.q_norm.data.bf(),
                                         L.attention.k_norm.data.bf(), cos_cache.data.bf(), sin_cache.data.bf(),
                                         P + o_pos, nh, nkv, hd, Tp, eps, st);
    const int64_t k_off = (int64_t)li * layout.layer_stride, v_off = k_off + layout.kv_block_len;
    if (k.paged_kv_scatter_cuda(kv_buffer.bf(), k_off, v_off, P + o_pi, P + o_ip, P + o_lpl, kb.data.bf(),
                                vb.data.bf(), P + o_bi, P + o_pos, Tp, nkv, hd, kPageSize, layout.page_stride, kd, hd,
                                st) != 0)
      return fail("paged_kv_scatter_cuda failed");
    if (k.batch_prefill_paged_cuda_with_cta_tile_q(
            qb.data.bf(), att.data.bf(), kv_buffer.bf(), k_off, v_off, P + o_pi, P + o_ip, P + o_lpl, P + o_qi,
            P + o_ri, P + o_qt, P + o_kt, P + o_kc, reinterpret_cast<const uint32_t*>(P + o_tnr), nh, nkv, hd,
            kPageSize, Tp, n_req, num_tiles, layout.page_stride, sm_scale, cta_tile_q, st) != 0)
      return fail("batch_prefill_paged_cuda failed");
    if (n_dec > 0) {  // decode rows [Tp, T): one query token per request over its paged context (unified_forward.rs:318-519)
      pk_bf16 *qd_ = qb.data.bf() + (size_t)Tp * qd, *kd_ = kb.data.bf() + (size_t)Tp * kd, *vd_ = vb.data.bf() + (size_t)Tp * kd;
      pk_bf16* od_ = att.data.bf() + (size_t)Tp * qd;
      if (rt.mode >= 1 && k.pk_b200_decode_attention_fused) {
        if (k.pk_b200_decode_attention_fused(qd_, kd_, vd_, od_, kv_buffer.bf(), k_off, v_off, P + o_dpi, P + o_dip, P + o_dlpl, P + o_dpos,
                                             L.attention.q_norm.data.bf(), L.attention.k_norm.data.bf(), cos_cache.data.bf(),
                                             sin_cache.data.bf(), eps, static_cast<float*>(attn_partial.ptr), attn_counters.i32(), 64,
                                             attn_max_chunks, nh, nkv, hd, kPageSize, n_dec, layout.page_stride, sm_scale, st) != 0)
          return fail("pk_b200_decode_attention_fused failed");
      } else {
        k.qk_norm_rope_batched_decode_cuda(qd_, kd_, L.attention.q_norm.data.bf(), L.attention.k_norm.data.bf(), cos_cache.data.bf(),
                                           sin_cache.data.bf(), P + o_dpos, nh, nkv, hd, n_dec, eps, st);
        if (k.paged_kv_scatter_cuda(kv_buffer.bf(), k_off, v_off, P + o_dpi, P + o_dip, P + o_dlpl, kd_, vd_, P + o_dreq, P + o_dpos, n_dec,
                                    nkv, hd, kPageSize, layout.page_stride, kd, hd, st) != 0)
          return fail("paged_kv_scatter_cuda (unified decode rows) failed");
        if (k.paged_attention_decode_cuda(qd_, od_, kv_buffer.bf(), k_off, v_off, P + o_dpi, P + o_dip, P + o_dlpl, P + o_dreq, P + o_dtile,
                                          P + o_dchunk, nh, nkv, hd, kPageSize, n_dec, layout.page_stride, sm_scale, st) != 0)
          return fail("paged_attention_decode_cuda (unified decode rows) failed");
      }
    }
    k.gemm_cuda(L.attention.o_proj.data.bf(), att.data.bf(), ob.data.bf(), H, T, qd, st);
    if (T == 1) {}  // (gemm_into picks the graph-safe entry for T == 1; gemm_cuda handles it too)
    if (!all_reduce_hidden(ob.data.bf(), H, T)) return false;
    k.fused_add_rms_norm_batched_cuda(hcur, ob.data.bf(), L.post_attention_layernorm.data.bf(), nrm.data.bf(), H, T,
                                      eps, st);
    // gate_up GEMM with the SwiGLU activation in its epilogue (one launch, same rounding points) when the kernel
    // library offers it and takes the shape; else the reference's two launches
    if (!(rt.mode >= 1 && k.pk_b200_gemm_swiglu &&
          k.pk_b200_gemm_swiglu(L.mlp.gate_up_proj.data.bf(), nrm.data.bf(), act.data.bf(), I, T, H, st) == 0)) {
      k.gemm_cuda(L.mlp.gate_up_proj.data.bf(), nrm.data.bf(), gu.data.bf(), 2 * I, T, H, st);
      k.silu_mul_fused_cuda(gu.data.bf(), act.data.bf(), I, T, st);
    }
    k.gemm_cuda(L.mlp.down_proj.data.bf(), act.data.bf(), ob.data.bf(), H, T, I, st);
    if (!all_reduce_hidden(ob.data.bf(), H, T)) return false;
    k.add_cuda(hcur, ob.data.bf(), hnext, H * T, st);  // prefill.rs:183 (rounds the residual sum)
    std::swap(hcur, hnext);
  }
  // last-token logits per request (prefill.rs:267-282): extract_vec -> rms_norm -> linear; then one row per decode token
  int off = 0;
  for (int i = 0; i < n_req + n_dec; ++i) {
    const int last = i < n_req ? off + lens[i] - 1 : Tp + (i - n_req);
    pk_bf16* lg = logits.data.bf() + (size_t)i * local_vocab();
    k.rms_norm_cuda(hcur + (size_t)last * H, norm.data.bf(), normed.data.bf() + (size_t)i * H, H, eps, st);
    k.gemm_graphsafe_cuda(output_rows(), normed.data.bf() + (size_t)i * H, lg, local_vocab(), 1, H, st);
    if (i < n_req) {
      logits_out[i] = lg;
      off += lens[i];
    } else if (dec_logits_out) {
      dec_logits_out[i - n_req] = lg;
    }
  }
  if (!cu(cudaStreamSynchronize(st), "prefill sync")) return false;
  txn.committed = true;
  return true;
}

// ===================================================================== decode (batch_decode.rs)
static int bucket_for(int bs) {
  for (int b : kBuckets)
    if (b >= bs) return b;
  return -1;
}

bool Qwen3Model::decode_kernels_compat(int bs, bool split) {
  const Config& c = config;
  const int H = c.hidden_size, qd = local_q_dim(), kd = local_kv_dim(), I = local_inter();
  const int nh = local_heads(), nkv = local_kv_heads(), hd = c.head_dim;
  const float eps = c.rms_norm_eps, sm_scale = 1.0f / sqrtf((float)hd);
  cudaStream_t st = ctx.stream;
  const int* M = meta_d.i32();
  k.embedding_batched_cuda(embed_tokens.data.bf(), reinterpret_cast<const uint32_t*>(M + mo.token_ids),
                           hidden.data.bf(), H, bs, st);
  k.rms_norm_batched_cuda(hidden.data.bf(), layers[0].input_layernorm.data.bf(), normed.data.bf(), H, bs, eps, st);
  for (int li = 0; li < c.num_hidden_layers; ++li) {  // batch_decode_layer (batch_decode.rs:148-295)
    TransformerBlock& L = layers[li];
    gemm_decode(L.attention.qkv_proj, 0, qd, normed.data.bf(), bs, q.data.bf());
    gemm_decode(L.attention.qkv_proj, qd, kd, normed.data.bf(), bs, kbuf.data.bf());
    gemm_decode(L.attention.qkv_proj, qd + kd, kd, normed.data.bf(), bs, v.data.bf());
    k.qk_norm_rope_batched_decode_cuda(q.data.bf(), kbuf.data.bf(), L.attention.q_norm.data.bf(),
                                       L.attention.k_norm.data.bf(), cos_cache.data.bf(), sin_cache.data.bf(),
                                       M + mo.positions, nh, nkv, hd, bs, eps, st);
    const int64_t k_off = (int64_t)li * layout.layer_stride, v_off = k_off + layout.kv_block_len;
    if (k.paged_kv_scatter_cuda(kv_buffer.bf(), k_off, v_off, M + mo.page_indices, M + mo.page_indptr,
                                M + mo.last_page_len, kbuf.data.bf(), v.data.bf(), M + mo.request_indices,
                                M + mo.positions, bs, nkv, hd, kPageSize, layout.page_stride, kd, hd, st) != 0)
      return fail("paged_kv_scatter_cuda (batch decode) failed");
    int rc;
    if (split)
      rc = k.paged_attention_decode_split_kv_cuda(
          q.data.bf(), attn_out.data.bf(), kv_buffer.bf(), k_off, v_off, M + mo.page_indices, M + mo.page_indptr,
          M + mo.last_page_len, M + mo.split_request, M + mo.split_tile, M + mo.split_chunk, M + mo.split_o_indptr,
          reinterpret_cast<const uint8_t*>(M + mo.split_mask_bytes), split_tmp_v.bf(),
          static_cast<float*>(split_tmp_s.ptr), nh, nkv, hd, kPageSize, bs, split_padded_slots, layout.page_stride,
          sm_scale, st);
    else
      rc = k.paged_attention_decode_cuda(q.data.bf(), attn_out.data.bf(), kv_buffer.bf(), k_off, v_off,
                                         M + mo.page_indices, M + mo.page_indptr, M + mo.last_page_len,
                                         M + mo.request_indices, M + mo.kv_tile_indices, M + mo.kv_chunk_size, nh, nkv,
                                         hd, kPageSize, bs, layout.page_stride, sm_scale, st);
    if (rc != 0) return fail("paged_attention_decode failed");
    gemm_decode(L.attention.o_proj, 0, H, attn_out.data.bf(), bs, attn_proj.data.bf());
    if (!all_reduce_hidden(attn_proj.data.bf(), H, bs)) return false;
    k.fused_add_rms_norm_batched_cuda(hidden.data.bf(), attn_proj.data.bf(), L.post_attention_layernorm.data.bf(),
                                      normed.data.bf(), H, bs, eps, st);
    gemm_decode(L.mlp.gate_up_proj, 0, 2 * I, normed.data.bf(), bs, gate_up_out.data.bf());
    k.silu_mul_fused_cuda(gate_up_out.data.bf(), mlp_act.data.bf(), I, bs, st);
    gemm_decode(L.mlp.down_proj, 0, H, mlp_act.data.bf(), bs, mlp_out.data.bf());
    if (!all_reduce_hidden(mlp_out.data.bf(), H, bs)) return false;
    const DeviceVec& nw = li + 1 < c.num_hidden_layers ? layers[li + 1].input_layernorm : norm;
    k.fused_add_rms_norm_batched_cuda(hidden.data.bf(), mlp_out.data.bf(), nw.data.bf(), normed.data.bf(), H, bs, eps, st);
  }
  k.gemm_graphsafe_cuda(output_rows(), normed.data.bf(), logits.data.bf(), local_vocab(), bs, H, ctx.stream);
  return true;
}

// B200 path: 5 launches per layer (TP: 7), residual-add + RMSNorm folded into the consuming GEMV's
// prologue, SwiGLU into the gate_up epilogue, QK-norm/RoPE/append/split-KV/merge into one attention
// launch, PDL between all of them.  Same rounding points as decode_kernels_compat.
bool Qwen3Model::decode_kernels_fused(int bs) {
  const Config& c = config;
  const int H = c.hidden_size, qd = local_q_dim(), kd = local_kv_dim(), I = local_inter();
  const int nh = local_heads(), nkv = local_kv_heads(), hd = c.head_dim;
  const float eps = c.rms_norm_eps, sm_scale = 1.0f / sqrtf((float)hd);
  cudaStream_t st = ctx.stream;
  const int* M = meta_d.i32();
  const bool tp_on = tp.is_sharded();
  pk_bf16* Ha = hidden.data.bf();
  pk_bf16* Hb = hidden_b.data.bf();
  k.embedding_batched_cuda(embed_tokens.data.bf(), reinterpret_cast<const uint32_t*>(M + mo.token_ids), Ha, H, bs, st);
  int tp_op = 0;
  auto gemv = [&](const pk_bf16* W, const pk_bf16* X, int Mrows, int K, pk_bf16* y0, pk_bf16* y1, pk_bf16* y2,
                  int s0, int s1, int s2, int x_mode, const pk_bf16* residual, const pk_bf16* nw, pk_bf16* hout,
                  int epi) -> bool {
    pk_b200_gemv_args g{};
    g.W = W; g.X = X;
    g.Y[0] = y0; g.Y[1] = y1; g.Y[2] = y2;
    g.seg_rows[0] = s0; g.seg_rows[1] = s1; g.seg_rows[2] = s2;
    g.M = Mrows; g.N = bs; g.K = K;
    g.x_mode = x_mode; g.residual = residual; g.norm_w = nw; g.eps = eps;
    g.hidden_out = hout; g.normed_out = nullptr; g.epi = epi;
    g.tp_comm = tp_comm;
    if (epi == 3) {  // all-reduce inside the GEMV: sequence = (step counter, op index)
      g.tp_step = reinterpret_cast<const uint32_t*>(M + mo.step_seq);
      g.tp_op = tp_op++;
    }
    if (k.pk_b200_gemv_fused(&g, st) != 0) return fail("pk_b200_gemv_fused rejected its arguments");
    return true;
  };
  // Fused attention launch; when the kernel library offers it, the launch also prefetches into L2 the weights the
  // next two GEMVs will stream (all of o_proj, the leading rows of every gate_up slice): bs-1 attention is
  // latency-bound and would otherwise leave HBM idle for ~10 us per layer.  PK_PF_O / PK_PF_GU = rows per GEMV
  // slice to request (0 disables), defaults tuned on B200 (profiles/README.md).
  static const int pf_o_rows = [] { const char* e = getenv("PK_PF_O"); return e ? atoi(e) : 0; }();
  static const int pf_gu_rows = [] { const char* e = getenv("PK_PF_GU"); return e ? atoi(e) : 0; }();
  const bool can_prefetch = k.pk_b200_decode_attention_fused_prefetch && k.pk_b200_gemv_grid && (pf_o_rows > 0 || pf_gu_rows > 0);
  const int o_slices = can_prefetch ? k.pk_b200_gemv_grid(H, 0) : 1;
  const int gu_slices = can_prefetch ? k.pk_b200_gemv_grid(I, 1) : 1;
  auto attention = [&](TransformerBlock& L, int64_t k_off, int64_t v_off) -> bool {
    if (!can_prefetch) {
      if (k.pk_b200_decode_attention_fused(
              q.data.bf(), kbuf.data.bf(), v.data.bf(), attn_out.data.bf(), kv_buffer.bf(), k_off, v_off,
              M + mo.page_indices, M + mo.page_indptr, M + mo.last_page_len, M + mo.positions,
              L.attention.q_norm.data.bf(), L.attention.k_norm.data.bf(), cos_cache.data.bf(), sin_cache.data.bf(), eps,
              static_cast<float*>(attn_partial.ptr), attn_counters.i32(), 64, attn_max_chunks, nh, nkv, hd, kPageSize,
              bs, layout.page_stride, sm_scale, st) != 0)
        return fail("pk_b200_decode_attention_fused failed");
      return true;
    }
    pk_b200_prefetch_span sp[3];
    int ns = 0;
    if (pf_o_rows > 0) sp[ns++] = pk_b200_prefetch_span{L.attention.o_proj.data.bf(), H, qd * 2, o_slices, pf_o_rows};
    if (pf_gu_rows > 0) {
      const pk_bf16* gu = L.mlp.gate_up_proj.data.bf();
      sp[ns++] = pk_b200_prefetch_span{gu, I, H * 2, gu_slices, pf_gu_rows};
      sp[ns++] = pk_b200_prefetch_span{gu + (size_t)I * H, I, H * 2, gu_slices, pf_gu_rows};
    }
    if (k.pk_b200_decode_attention_fused_prefetch(
            q.data.bf(), kbuf.data.bf(), v.data.bf(), attn_out.data.bf(), kv_buffer.bf(), k_off, v_off,
            M + mo.page_indices, M + mo.page_indptr, M + mo.last_page_len, M + mo.positions,
            L.attention.q_norm.data.bf(), L.attention.k_norm.data.bf(), cos_cache.data.bf(), sin_cache.data.bf(), eps,
            static_cast<float*>(attn_partial.ptr), attn_counters.i32(), 64, attn_max_chunks, nh, nkv, hd, kPageSize, bs,
            layout.page_stride, sm_scale, sp, ns, st) != 0)
      return fail("pk_b200_decode_attention_fused_prefetch failed");
    return true;
  };
  // Tensor parallel, three variants (all custom kernels over NVLink peer memory, no NCCL on the data path; PK_TP_MODE):
  //  * kernel: one-shot all-reduce kernel fused with add + RMSNorm between the GEMVs (7 launches/layer).
  //  * fused: the row-parallel GEMVs (o_proj, down_proj) push their partial rows to every rank
  //    (epi 2) and the following GEMV's prologue reduces them (x_mode 2): 5 launches per layer, the layer's two
  //    all-reduces live inside the GEMVs.
  //  Measured (Qwen3-8B, bs 1, profiles/README.md): 2 x B200 398 vs 367 tok/s, 8 x B200 411 vs 350 tok/s -- every one
  //  of the push epilogue's ~300 CTAs pays a system-scope fence after its remote stores and the grid ticket
  //  serialises behind them, which costs more than a single-CTA collective launch.
  //  * ll (default): the row-parallel GEMVs reduce their own rows INSIDE the kernel (epi 3: 8-byte {data, seq} lines
  //    pushed to the peers, each CTA polls and sums its own rows) and write the reduced vector; the following GEMV
  //    takes the plain residual prologue of the single-GPU path.  5 launches per layer, no fence, no flag, no ticket.
  static const int tp_mode_env = [] {  // 0 kernel, 1 fused (flag protocol), 2 ll
    const char* e = getenv("PK_TP_MODE");
    if (e && strcmp(e, "kernel") == 0) return 0;
    if (e && strcmp(e, "fused") == 0) return 1;
    const char* f = getenv("PK_TP_FUSED");  // round-1 switch
    if (f && !e) return atoi(f) > 0 ? 1 : 0;
    return 2;
  }();
  const bool tp_ll = tp_on && tp_mode_env == 2;
  const bool tp_fuse = tp_on && tp_mode_env >= 1;
  const int red_mode = (tp_on && !tp_ll) ? 2 : 1;
  const int push_epi = tp_on ? (tp_ll ? 3 : 2) : 0;
  const pk_bf16* prev_residual = zero_residual.bf();  // layer 0: hidden + 0
  if (tp_on && !tp_fuse) {
    k.rms_norm_batched_cuda(Ha, layers[0].input_layernorm.data.bf(), normed.data.bf(), H, bs, eps, st);
    for (int li = 0; li < c.num_hidden_layers; ++li) {
      TransformerBlock& L = layers[li];
      const int64_t k_off = (int64_t)li * layout.layer_stride, v_off = k_off + layout.kv_block_len;
      if (!gemv(L.attention.qkv_proj.data.bf(), normed.data.bf(), qd + 2 * kd, H, q.data.bf(), kbuf.data.bf(),
                v.data.bf(), qd, kd, kd, 0, nullptr, nullptr, nullptr, 0))
        return false;
      if (!attention(L, k_off, v_off)) return false;
      if (!gemv(L.attention.o_proj.data.bf(), attn_out.data.bf(), H, qd, attn_proj.data.bf(), nullptr, nullptr, H, 0, 0,
                0, nullptr, nullptr, nullptr, 0))
        return false;
      if (k.pk_tp_all_reduce_add_rms_norm(tp_comm, Ha, attn_proj.data.bf(), L.post_attention_layernorm.data.bf(),
                                          normed.data.bf(), H, bs, eps, st) != 0)
        return fail("pk_tp_all_reduce_add_rms_norm failed");
      if (!gemv(L.mlp.gate_up_proj.data.bf(), normed.data.bf(), I, H, mlp_act.data.bf(), nullptr, nullptr, I, 0, 0, 0,
                nullptr, nullptr, nullptr, 1))
        return false;
      if (!gemv(L.mlp.down_proj.data.bf(), mlp_act.data.bf(), H, I, mlp_out.data.bf(), nullptr, nullptr, H, 0, 0, 0,
                nullptr, nullptr, nullptr, 0))
        return false;
      const DeviceVec& nw = li + 1 < c.num_hidden_layers ? layers[li + 1].input_layernorm : norm;
      if (k.pk_tp_all_reduce_add_rms_norm(tp_comm, Ha, mlp_out.data.bf(), nw.data.bf(), normed.data.bf(), H, bs, eps,
                                          st) != 0)
        return fail("pk_tp_all_reduce_add_rms_norm failed");
    }
    if (!gemv(output_rows(), normed.data.bf(), local_vocab(), H, logits.data.bf(), nullptr, nullptr,
              local_vocab(), 0, 0, 0, nullptr, nullptr, nullptr, 0))
      return false;
  } else {
    for (int li = 0; li < c.num_hidden_layers; ++li) {
      TransformerBlock& L = layers[li];
      const int64_t k_off = (int64_t)li * layout.layer_stride, v_off = k_off + layout.kv_block_len;
      // q|k|v = W_qkv . RMSNorm(Ha + prev_residual); Hb = Ha + prev_residual
      if (!gemv(L.attention.qkv_proj.data.bf(), Ha, qd + 2 * kd, H, q.data.bf(), kbuf.data.bf(), v.data.bf(), qd, kd,
                kd, li == 0 ? 1 : red_mode, prev_residual, L.input_layernorm.data.bf(), Hb, 0))
        return false;
      if (!attention(L, k_off, v_off)) return false;
      if (!gemv(L.attention.o_proj.data.bf(), attn_out.data.bf(), H, qd, attn_proj.data.bf(), nullptr, nullptr, H, 0, 0,
                0, nullptr, nullptr, nullptr, push_epi))
        return false;
      // act = SwiGLU(W_gate_up . RMSNorm(Hb + attn_proj)); Ha = Hb + attn_proj
      if (!gemv(L.mlp.gate_up_proj.data.bf(), Hb, I, H, mlp_act.data.bf(), nullptr, nullptr, I, 0, 0, red_mode,
                attn_proj.data.bf(), L.post_attention_layernorm.data.bf(), Ha, 1))
        return false;
      if (!gemv(L.mlp.down_proj.data.bf(), mlp_act.data.bf(), H, I, mlp_out.data.bf(), nullptr, nullptr, H, 0, 0, 0,
                nullptr, nullptr, nullptr, push_epi))
        return false;
      prev_residual = mlp_out.data.bf();
    }
    if (!gemv(output_rows(), Ha, local_vocab(), H, logits.data.bf(), nullptr, nullptr, local_vocab(), 0,
              0, red_mode, prev_residual, norm.data.bf(), Hb, 0))
      return false;
  }
  // greedy token for every request inside the same graph
  for (int b = 0; b < bs; ++b)
    k.flashinfer_top1_cuda(logits.data.bf() + (size_t)b * local_vocab(), static_cast<pk_bf16*>(top1_val.ptr) + b,
                           static_cast<uint8_t*>(top1_states.ptr) + (size_t)b * 8192, sample_out.i32() + b,
                           local_vocab(), st);
  if (vocab_sharded() &&
      k.pk_tp_top1_exchange(tp_comm, static_cast<pk_bf16*>(top1_val.ptr), sample_out.i32(), bs, vocab_offset(),
                            reinterpret_cast<const uint32_t*>(M + mo.step_seq), 250u, st) != 0)
    return fail("pk_tp_top1_exchange failed");
  return true;
}

// B200 path for the reference's decode buckets above 4 requests (batch_decode_buffers.rs:12: 8, 16, 32, 64): the
// projections run on the tensor cores (skinny split-K GEMM: every SM streams a K slice of the weights), norms stay
// separate launches over [H, bs], QK-norm + RoPE + KV append + attention + merge is the one fused attention launch over
// all requests (cluster size shrinks with the batch), and every row's greedy token is taken inside the graph.
// Same rounding points as decode_kernels_compat (it IS that op sequence with three fusions).
bool Qwen3Model::decode_kernels_wide(int bs) {
  const Config& c = config;
  const int H = c.hidden_size, qd = local_q_dim(), kd = local_kv_dim(), I = local_inter();
  const int nh = local_heads(), nkv = local_kv_heads(), hd = c.head_dim;
  const float eps = c.rms_norm_eps, sm_scale = 1.0f / sqrtf((float)hd);
  cudaStream_t st = ctx.stream;
  const int* M = meta_d.i32();
  k.embedding_batched_cuda(embed_tokens.data.bf(), reinterpret_cast<const uint32_t*>(M + mo.token_ids), hidden.data.bf(), H, bs, st);
  k.rms_norm_batched_cuda(hidden.data.bf(), layers[0].input_layernorm.data.bf(), normed.data.bf(), H, bs, eps, st);
  for (int li = 0; li < c.num_hidden_layers; ++li) {
    TransformerBlock& L = layers[li];
    pk_bf16* outs[3] = {q.data.bf(), kbuf.data.bf(), v.data.bf()};
    const int segs[3] = {qd, kd, kd};
    if (k.pk_b200_gemm_segments(L.attention.qkv_proj.data.bf(), normed.data.bf(), outs, segs, qd + 2 * kd, bs, H, st) != 0)
      return fail("pk_b200_gemm_segments failed");
    const int64_t k_off = (int64_t)li * layout.layer_stride, v_off = k_off + layout.kv_block_len;
    if (k.pk_b200_decode_attention_fused(q.data.bf(), kbuf.data.bf(), v.data.bf(), attn_out.data.bf(), kv_buffer.bf(), k_off, v_off,
                                         M + mo.page_indices, M + mo.page_indptr, M + mo.last_page_len, M + mo.positions,
                                         L.attention.q_norm.data.bf(), L.attention.k_norm.data.bf(), cos_cache.data.bf(),
                                         sin_cache.data.bf(), eps, static_cast<float*>(attn_partial.ptr), attn_counters.i32(), 64,
                                         attn_max_chunks, nh, nkv, hd, kPageSize, bs, layout.page_stride, sm_scale, st) != 0)
      return fail("pk_b200_decode_attention_fused failed");
    gemm_decode(L.attention.o_proj, 0, H, attn_out.data.bf(), bs, attn_proj.data.bf());
    if (!all_reduce_hidden(attn_proj.data.bf(), H, bs)) return false;
    k.fused_add_rms_norm_batched_cuda(hidden.data.bf(), attn_proj.data.bf(), L.post_attention_layernorm.data.bf(), normed.data.bf(), H,
                                      bs, eps, st);
    gemm_decode(L.mlp.gate_up_proj, 0, 2 * I, normed.data.bf(), bs, gate_up_out.data.bf());
    k.silu_mul_fused_cuda(gate_up_out.data.bf(), mlp_act.data.bf(), I, bs, st);
    gemm_decode(L.mlp.down_proj, 0, H, mlp_act.data.bf(), bs, mlp_out.data.bf());
    if (!all_reduce_hidden(mlp_out.data.bf(), H, bs)) return false;
    const DeviceVec& nw = li + 1 < c.num_hidden_layers ? layers[li + 1].input_layernorm : norm;
    k.fused_add_rms_norm_batched_cuda(hidden.data.bf(), mlp_out.data.bf(), nw.data.bf(), normed.data.bf(), H, bs, eps, st);
  }
  k.gemm_graphsafe_cuda(output_rows(), normed.data.bf(), logits.data.bf(), local_vocab(), bs, H, st);
  for (int b = 0; b < bs; ++b)
    k.flashinfer_top1_cuda(logits.data.bf() + (size_t)b * local_vocab(), static_cast<pk_bf16*>(top1_val.ptr) + b,
                           static_cast<uint8_t*>(top1_states.ptr) + (size_t)b * 8192, sample_out.i32() + b, local_vocab(), st);
  if (vocab_sharded() &&
      k.pk_tp_top1_exchange(tp_comm, static_cast<pk_bf16*>(top1_val.ptr), sample_out.i32(), bs, vocab_offset(),
                            reinterpret_cast<const uint32_t*>(M + mo.step_seq), 250u, st) != 0)
    return fail("pk_tp_top1_exchange failed");
  return true;
}

// Advance the KV states by one token and write the step's packed metadata block into `mh`
// (batch_decode.rs:26-59 + sync_paged_meta / sync_split_kv_meta of batch_decode_buffers.rs:177-279).
bool Qwen3Model::build_step_meta(int bs, int padded, const uint32_t* tokens, const int* kv_ids, int* mh,
                                 bool* split_out) {
  const bool fused = rt.mode >= 1;
  std::vector<int> positions(bs);
  // validate the whole step first, then commit (no request is advanced unless every request can be)
  int pages_needed = 0, table_pages = 0;
  for (int b = 0; b < bs; ++b) {
    if (kv_ids[b] < 0 || kv_ids[b] >= (int)kv_states.size() || !kv_states[kv_ids[b]].live) return fail("bad kv id");
    for (int j = 0; j < b; ++j)
      if (kv_ids[j] == kv_ids[b]) return fail("duplicate kv id in one decode batch");
    const KvState& s = kv_states[kv_ids[b]];
    if (s.seq_len + 1 > kRopePositions) return fail("position beyond the 4096-entry RoPE table");
    const int need = (s.seq_len + 1 + kPageSize - 1) / kPageSize;
    pages_needed += std::max(0, need - (int)s.pages.size());
    table_pages += std::max(need, (int)s.pages.size());
  }
  if (pages_needed > (int)pool.free_list.size()) {
    char eb[128];
    snprintf(eb, sizeof eb, "KvState: out of pages (need %d more, %d available)", pages_needed,
             (int)pool.free_list.size());
    return fail(eb);
  }
  if (table_pages + (padded - bs) > max_total_pages + max_bs) return fail("page table overflow");
  for (int b = 0; b < bs; ++b) {
    KvState& s = kv_states[kv_ids[b]];
    positions[b] = s.seq_len;
    if (!ensure_capacity(s, s.seq_len + 1)) return false;  // cannot fail: counted above
    s.seq_len += 1;
  }
  memset(mh, 0, meta_bytes);
  mh[mo.step_seq] = (int)(++step_counter & 0xffffffu);
  int np = 0;
  max_seq_len_step = 0;
  mh[mo.page_indptr] = 0;
  for (int b = 0; b < padded; ++b) {
    if (b < bs) {
      const KvState& s = kv_states[kv_ids[b]];
      memcpy(mh + mo.page_indices + np, s.pages.data(), s.pages.size() * 4);
      np += (int)s.pages.size();
      mh[mo.last_page_len + b] = s.last_page_len(kPageSize);
      mh[mo.kv_chunk_size + b] = s.seq_len;
      mh[mo.token_ids + b] = tokens ? (int)tokens[b] : 0;
      mh[mo.positions + b] = positions[b];
      max_seq_len_step = std::max(max_seq_len_step, s.seq_len);
    } else {  // padding slot: the padding page, seq_len 1
      mh[mo.page_indices + np++] = padding_page;
      mh[mo.last_page_len + b] = 1;
      mh[mo.kv_chunk_size + b] = 1;
    }
    mh[mo.page_indptr + b + 1] = np;
    mh[mo.request_indices + b] = b;
  }
  const bool split = !fused && padded <= kSplitMaxBs && max_seq_len_step >= kSplitMinSeq;
  {
    const int chunk = std::max(kSplitChunkTokens, (max_seq_len_step + kSplitMaxChunks - 1) / kSplitMaxChunks);
    split_padded_slots = padded * kSplitMaxChunks;
    uint8_t* mask = reinterpret_cast<uint8_t*>(mh + mo.split_mask_bytes);
    int ns = 0;
    mh[mo.split_o_indptr] = 0;
    for (int b = 0; b < bs; ++b) {
      const int len = kv_states[kv_ids[b]].seq_len;
      const int chunks = std::max(1, (len + chunk - 1) / chunk);
      for (int cidx = 0; cidx < chunks && ns < split_padded_slots; ++cidx) {
        mh[mo.split_request + ns] = b;
        mh[mo.split_tile + ns] = cidx;
        mask[ns++] = 1;
      }
      mh[mo.split_o_indptr + b + 1] = ns;
    }
    for (int b = bs; b < padded; ++b) mh[mo.split_o_indptr + b + 1] = ns;
    mh[mo.split_chunk] = chunk;
  }
  *split_out = split;
  return true;
}

// CudaGraphState::run_or_capture (cuda_graph.rs:28-58) around the decode body
bool Qwen3Model::run_step_kernels(int padded, bool split) {
  const bool fused = rt.mode >= 1;
  cudaStream_t st = ctx.stream;
  auto body = [&]() {
    if (fused && padded > 4) return decode_kernels_wide(padded);
    return fused ? decode_kernels_fused(padded) : decode_kernels_compat(padded, split);
  };
  if (!rt.enable_cuda_graph) return body();
  const int key = padded * 4 + (fused ? 2 : (split ? 1 : 0));
  auto& g = graphs[key];
  if (!g) g.reset(new CudaGraphState());
  if (!g->captured()) {
    const int64_t before = k.pk_b200_launch_count ? k.pk_b200_launch_count(0) : 0;
    if (!cu(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal), "begin capture")) return false;
    const bool ok = body();
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (!ok) {
      if (graph) cudaGraphDestroy(graph);
      return false;
    }
    if (!cu(e, "end capture")) return false;
    e = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (!cu(e, "graph instantiate")) return false;
    if (k.pk_b200_launch_count) launches_per_step = k.pk_b200_launch_count(0) - before;
  }
  return cu(cudaGraphLaunch(g->exec, st), "graph launch");
}

bool Qwen3Model::decode(int bs, const uint32_t* tokens, const int* kv_ids, void** logits_out, int* sampled) {
  if (!finalized) return fail("model not finalized");
  if (bs <= 0 || bs > max_bs) return fail("batch size out of range");
  const bool fused = rt.mode >= 1;
  // fused path: 1..4 requests run the GEMV graph of their exact size; larger batches pad to the reference's buckets
  const int padded = (rt.enable_cuda_graph && (!fused || bs > 4)) ? bucket_for(bs) : bs;
  if (padded < 0 || padded > max_bs) return fail("batch exceeds max_batch bucket");
  bool split = false;
  if (!build_step_meta(bs, padded, tokens, kv_ids, static_cast<int*>(meta_h), &split)) return false;
  cudaStream_t st = ctx.stream;
  if (!cu(cudaMemcpyAsync(meta_d.ptr, meta_h, meta_bytes, cudaMemcpyHostToDevice, st), "meta H2D")) return false;
  if (!run_step_kernels(padded, split)) return false;
  if (logits_out) *logits_out = logits.data.ptr;
  if (sampled) {
    if (fused) {
      if (!cu(cudaMemcpyAsync(sample_h, sample_out.ptr, (size_t)bs * 4, cudaMemcpyDeviceToHost, st), "sample D2H"))
        return false;
      if (!cu(cudaStreamSynchronize(st), "decode sync")) return false;
      for (int b = 0; b < bs; ++b) sampled[b] = sample_h[b];
    } else {
      for (int b = 0; b < bs; ++b)  // per request: top-1 kernel, sync, 4-byte D2H (sampling.rs:161-172)
        if (!sample_greedy(logits.data.bf() + (size_t)b * local_vocab(), sampled + b)) return false;
    }
  }
  return true;
}

// K greedy decode steps of ONE request with everything resident in HBM: the per-step metadata blocks
// are staged on the device up front, each step's token id is the previous step's on-device arg-max,
// and there is no host synchronisation inside the timed region (CUDA events on the launch stream).
bool Qwen3Model::decode_burst(int kv_id, uint32_t first_token, int K, uint32_t* tokens_out, float* ms_total) {
  if (!finalized || rt.mode < 1) return fail("decode_burst needs the fused path");
  if (K <= 0) return fail("K must be positive");
  {  // all K steps must fit before the first one advances the request
    if (kv_id < 0 || kv_id >= (int)kv_states.size() || !kv_states[kv_id].live) return fail("bad kv id");
    const KvState& s = kv_states[kv_id];
    if (s.seq_len + K > kRopePositions) return fail("position beyond the 4096-entry RoPE table");
    const int need = (s.seq_len + K + kPageSize - 1) / kPageSize - (int)s.pages.size();
    if (need > (int)pool.free_list.size()) return fail("KvState: out of pages for the burst");
  }
  DeviceBuf staged;
  if (!staged.alloc_zeros((size_t)K * meta_bytes)) return fail("burst staging alloc failed");
  std::vector<int> host((size_t)K * mo.total_ints);
  bool split = false;
  for (int i = 0; i < K; ++i) {
    const uint32_t t = i == 0 ? first_token : 0u;
    if (!build_step_meta(1, 1, &t, &kv_id, host.data() + (size_t)i * mo.total_ints, &split)) return false;
  }
  cudaStream_t st = ctx.stream;
  DeviceBuf toks;
  if (!toks.alloc_zeros((size_t)K * 4)) return fail("burst alloc failed");
  if (!cu(cudaMemcpyAsync(staged.ptr, host.data(), (size_t)K * meta_bytes, cudaMemcpyHostToDevice, st), "stage H2D"))
    return false;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  if (!cu(cudaStreamSynchronize(st), "burst pre-sync")) return false;
  cudaEventRecord(e0, st);
  bool ok = true;
  for (int i = 0; ok && i < K; ++i) {
    ok = cu(cudaMemcpyAsync(meta_d.ptr, static_cast<char*>(staged.ptr) + (size_t)i * meta_bytes, meta_bytes,
                            cudaMemcpyDeviceToDevice, st), "meta D2D");
    if (ok && i > 0)  // feed the previous arg-max back as this step's token id
      ok = cu(cudaMemcpyAsync(meta_d.i32() + mo.token_ids, sample_out.ptr, 4, cudaMemcpyDeviceToDevice, st), "tok D2D");
    ok = ok && run_step_kernels(1, false);
    ok = ok && cu(cudaMemcpyAsync(toks.i32() + i, sample_out.ptr, 4, cudaMemcpyDeviceToDevice, st), "tok save");
  }
  cudaEventRecord(e1, st);
  ok = ok && cu(cudaStreamSynchronize(st), "burst sync");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_total) *ms_total = ms;
  if (ok && tokens_out) ok = cu(cudaMemcpy(tokens_out, toks.ptr, (size_t)K * 4, cudaMemcpyDeviceToHost), "tok D2H");
  return ok;
}

// The GEMV launches of one fused decode token (36 x {qkv, o, gate_up, down} + lm_head) back to back,
// `iters` times, timed with CUDA events: the roofline leg of bench.py (weights >> L2, so every pass is
// cold).  Returns ms per pass and the number of GEMV launches per pass.
bool Qwen3Model::bench_gemv_pass(int iters, float* ms_per_pass, int* launches_per_pass) {
  if (!finalized || rt.mode < 1 || tp.is_sharded()) return fail("bench_gemv_pass needs the fused single-GPU path");
  const Config& c = config;
  const int H = c.hidden_size, qd = local_q_dim(), kd = local_kv_dim(), I = local_inter();
  cudaStream_t st = ctx.stream;
  auto gemv = [&](const pk_bf16* W, const pk_bf16* X, int M, int K, pk_bf16* y, int xm, const pk_bf16* res,
                  const pk_bf16* nw, pk_bf16* hout, int epi) {
    pk_b200_gemv_args g{};
    g.W = W; g.X = X; g.Y[0] = y; g.seg_rows[0] = M; g.M = M; g.N = 1; g.K = K;
    g.x_mode = xm; g.residual = res; g.norm_w = nw; g.eps = c.rms_norm_eps; g.hidden_out = hout; g.epi = epi;
    return k.pk_b200_gemv_fused(&g, st) == 0;
  };
  int n = 0;
  auto pass = [&]() {
    bool ok = true;
    n = 0;
    for (int li = 0; ok && li < c.num_hidden_layers; ++li) {
      TransformerBlock& L = layers[li];
      ok = gemv(L.attention.qkv_proj.data.bf(), hidden.data.bf(), qd + 2 * kd, H, q.data.bf(), 1, zero_residual.bf(),
                L.input_layernorm.data.bf(), hidden_b.data.bf(), 0) &&
           gemv(L.attention.o_proj.data.bf(), attn_out.data.bf(), H, qd, attn_proj.data.bf(), 0, nullptr, nullptr, nullptr, 0) &&
           gemv(L.mlp.gate_up_proj.data.bf(), hidden_b.data.bf(), I, H, mlp_act.data.bf(), 1, zero_residual.bf(),
                L.post_attention_layernorm.data.bf(), hidden.data.bf(), 1) &&
           gemv(L.mlp.down_proj.data.bf(), mlp_act.data.bf(), H, I, mlp_out.data.bf(), 0, nullptr, nullptr, nullptr, 0);
      n += 4;
    }
    ok = ok && gemv(output_projection_full().data.bf(), hidden.data.bf(), c.vocab_size, H, logits.data.bf(), 1,
                    zero_residual.bf(), norm.data.bf(), hidden_b.data.bf(), 0);
    n += 1;
    return ok;
  };
  if (!pass()) return fail("gemv pass failed");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  bool ok = true;
  for (int i = 0; ok && i < iters; ++i) ok = pass();
  cudaEventRecord(e1, st);
  ok = ok && cu(cudaStreamSynchronize(st), "gemv pass sync");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_pass = ms / (float)iters;
  *launches_per_pass = n;
  return ok;
}

bool Qwen3Model::sample_greedy(const pk_bf16* lg, int* out) {
  k.flashinfer_top1_cuda(lg, static_cast<pk_bf16*>(top1_val.ptr), static_cast<uint8_t*>(top1_states.ptr),
                         sample_out.i32(), local_vocab(), ctx.stream);
  if (vocab_sharded() &&  // `lg` is this rank's vocabulary shard: agree on the global winner (every rank makes this call)
      k.pk_tp_top1_exchange(tp_comm, static_cast<pk_bf16*>(top1_val.ptr), sample_out.i32(), 1, vocab_offset(), nullptr,
                            0x80000000u | (++host_sample_seq & 0x7fffffffu), ctx.stream) != 0)
    return fail("pk_tp_top1_exchange failed");
  if (!cu(cudaMemcpyAsync(sample_h, sample_out.ptr, 4, cudaMemcpyDeviceToHost, ctx.stream), "sample D2H")) return false;
  if (!cu(cudaStreamSynchronize(ctx.stream), "sample sync")) return false;
  *out = sample_h[0];
  return true;
}

}  // namespace pq

// ===================================================================== C API (ctypes / FFI)
using pq::Qwen3Model;

extern "C" {

struct pq_config {
  int hidden_size, intermediate_size, num_hidden_layers, num_attention_heads, num_key_value_heads, head_dim,
      vocab_size;
  float rms_norm_eps, rope_theta;
  int tie_word_embeddings;
};
struct pq_runtime {
  int device_ordinal, tp_rank, tp_world, enable_cuda_graph, mode, num_pages, max_batch, enable_pdl;
};

static thread_local std::string g_create_err;

__attribute__((visibility("default"))) const char* pq_create_error() { return g_create_err.c_str(); }

__attribute__((visibility("default"))) void* pq_model_create(const pq_config* c, const pq_runtime* r,
                                                             const char* kernel_lib_path, void* tp_comm) {
  auto* m = new Qwen3Model();
  m->config = pq::Config{c->hidden_size, c->intermediate_size, c->num_hidden_layers, c->num_attention_heads,
                         c->num_key_value_heads, c->head_dim, c->vocab_size, c->rms_norm_eps, c->rope_theta,
                         c->tie_word_embeddings != 0};
  m->rt.device_ordinal = r->device_ordinal;
  m->rt.tp_rank = r->tp_rank;
  m->rt.tp_world = r->tp_world;
  m->rt.enable_cuda_graph = r->enable_cuda_graph;
  m->rt.mode = r->mode;
  m->rt.num_pages = r->num_pages;
  m->rt.max_batch = r->max_batch;
  m->rt.enable_pdl = r->enable_pdl;
  m->tp.rank = r->tp_rank;
  m->tp.world_size = r->tp_world;
  m->tp_comm = static_cast<pk_tp_comm*>(tp_comm);
  auto bail = [&](const std::string& e) -> void* {
    g_create_err = e;
    delete m;
    return nullptr;
  };
  std::string e = m->tp.validate_for(m->config);
  if (!e.empty()) return bail(e);
  if (c->head_dim != 128) return bail("head_dim must be 128 (HEAD_DIM is hard-coded in the reference kernels too)");
  e = m->k.load(kernel_lib_path);
  if (!e.empty()) return bail(e);
  if (m->rt.mode >= 1 && !m->k.has_extensions())
    return bail("fused mode needs the B200 extensions; this kernel library only has the reference ABI");
  // DeviceContext::new_with_device (tensor.rs:23-59)
  if (m->k.cuda_set_device(r->device_ordinal) != 0) return bail("cuda_set_device failed (no CUDA device?)");
  m->ctx.device = r->device_ordinal;
  if (cudaStreamCreateWithFlags(&m->ctx.stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(std::string("stream creation failed: ") + cudaGetErrorString(cudaGetLastError()));
  m->k.cublas_init();
  if (m->k.pk_b200_set_pdl) m->k.pk_b200_set_pdl(r->enable_pdl);
  g_create_err.clear();
  return m;
}

#define PQ_M static_cast<Qwen3Model*>(mp)
__attribute__((visibility("default"))) void pq_model_destroy(void* mp) { delete PQ_M; }
__attribute__((visibility("default"))) const char* pq_last_error(void* mp) { return PQ_M->err.c_str(); }
__attribute__((visibility("default"))) int pq_model_load_tensor(void* mp, const char* name, const void* data,
                                                                int rows, int cols) {
  return PQ_M->load_tensor(name, data, rows, cols) ? 0 : -1;
}
__attribute__((visibility("default"))) int pq_model_finalize(void* mp) { return PQ_M->finalize() ? 0 : -1; }
__attribute__((visibility("default"))) void* pq_stream(void* mp) { return PQ_M->ctx.stream; }

__attribute__((visibility("default"))) int pq_kv_alloc(void* mp) {
  auto& v = PQ_M->kv_states;
  for (size_t i = 0; i < v.size(); ++i)
    if (!v[i].live) {
      v[i] = pq::KvState();
      v[i].live = true;
      return (int)i;
    }
  v.emplace_back();
  v.back().live = true;
  return (int)v.size() - 1;
}
__attribute__((visibility("default"))) void pq_kv_free(void* mp, int id) {
  auto& v = PQ_M->kv_states;
  if (id < 0 || id >= (int)v.size() || !v[id].live) return;
  PQ_M->pool.release(v[id].pages);  // permit drop returns pages (kv_pool.rs, page_pool.rs)
  v[id] = pq::KvState();
}
__attribute__((visibility("default"))) int pq_kv_seq_len(void* mp, int id) {
  auto& v = PQ_M->kv_states;
  return (id < 0 || id >= (int)v.size()) ? -1 : v[id].seq_len;
}
__attribute__((visibility("default"))) int pq_available_pages(void* mp) { return (int)PQ_M->pool.free_list.size(); }

__attribute__((visibility("default"))) int pq_prefill(void* mp, int n_req, const uint32_t* tokens, const int* lens,
                                                      const int* kv_ids, void** logits_out) {
  return PQ_M->prefill(n_req, tokens, lens, kv_ids, logits_out) ? 0 : -1;
}
// unified_forward.rs:78: prompts and decode tokens in one forward pass; logits pointers per prompt / per decode row
__attribute__((visibility("default"))) int pq_unified_step(void* mp, int n_prefill, const uint32_t* tokens, const int* lens,
                                                           const int* prefill_kv_ids, int n_decode, const uint32_t* decode_tokens,
                                                           const int* decode_kv_ids, void** prefill_logits, void** decode_logits) {
  if (!PQ_M->finalized) return -1;
  return PQ_M->unified_step(n_prefill, tokens, lens, prefill_kv_ids, n_decode, decode_tokens, decode_kv_ids, prefill_logits,
                            decode_logits) ? 0 : -1;
}
__attribute__((visibility("default"))) int pq_decode(void* mp, int bs, const uint32_t* tokens, const int* kv_ids,
                                                     void** logits_out, int* sampled) {
  return PQ_M->decode(bs, tokens, kv_ids, logits_out, sampled) ? 0 : -1;
}
__attribute__((visibility("default"))) int pq_sample_greedy(void* mp, const void* logits, int* out) {
  return PQ_M->sample_greedy(static_cast<const pk_bf16*>(logits), out) ? 0 : -1;
}
__attribute__((visibility("default"))) int pq_sync(void* mp) {
  return cudaStreamSynchronize(PQ_M->ctx.stream) == cudaSuccess ? 0 : -1;
}
// copy a device buffer produced on the model's stream into caller-owned memory (device or host)
__attribute__((visibility("default"))) int pq_copy_out(void* mp, void* dst, const void* src, int64_t bytes) {
  if (cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, PQ_M->ctx.stream) != cudaSuccess) return -1;
  return cudaStreamSynchronize(PQ_M->ctx.stream) == cudaSuccess ? 0 : -1;
}
__attribute__((visibility("default"))) int64_t pq_launch_count(void* mp, int reset) {
  return PQ_M->k.pk_b200_launch_count ? PQ_M->k.pk_b200_launch_count(reset) : -1;
}
__attribute__((visibility("default"))) void* pq_debug_buffer(void* mp, const char* name) {
  Qwen3Model* m = PQ_M;
  const std::string n = name;
  if (n == "hidden") return m->hidden.data.ptr;
  if (n == "hidden_b") return m->hidden_b.data.ptr;
  if (n == "normed") return m->normed.data.ptr;
  if (n == "logits") return m->logits.data.ptr;
  if (n == "kv") return m->kv_buffer.ptr;
  if (n == "attn_out") return m->attn_out.data.ptr;
  return nullptr;
}

__attribute__((visibility("default"))) int pq_decode_burst(void* mp, int kv_id, uint32_t first_token, int K,
                                                           uint32_t* tokens_out, float* ms_total) {
  return PQ_M->decode_burst(kv_id, first_token, K, tokens_out, ms_total) ? 0 : -1;
}
__attribute__((visibility("default"))) int pq_bench_gemv_pass(void* mp, int iters, float* ms_per_pass, int* n) {
  return PQ_M->bench_gemv_pass(iters, ms_per_pass, n) ? 0 : -1;
}
__attribute__((visibility("default"))) int64_t pq_launches_per_step(void* mp) { return PQ_M->launches_per_step; }
// width of the logits rows this rank returns: vocab_size, or its vocabulary shard on the fused TP path (offset in *off)
__attribute__((visibility("default"))) int pq_logits_cols(void* mp, int* off) {
  if (off) *off = PQ_M->vocab_offset();
  return PQ_M->local_vocab();
}
__attribute__((visibility("default"))) int64_t pq_meta_bytes(void* mp) { return (int64_t)PQ_M->meta_bytes; }
// record a CUDA event pair around caller-driven work on the model's stream (bench.py e2e leg)
__attribute__((visibility("default"))) void* pq_event_record(void* mp) {
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  cudaEventRecord(e, PQ_M->ctx.stream);
  return e;
}
__attribute__((visibility("default"))) float pq_event_elapsed_ms(void* e0, void* e1) {
  float ms = -1.f;
  cudaEventSynchronize(static_cast<cudaEvent_t>(e1));
  cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(e0), static_cast<cudaEvent_t>(e1));
  cudaEventDestroy(static_cast<cudaEvent_t>(e0));
  cudaEventDestroy(static_cast<cudaEvent_t>(e1));
  return ms;
}

// Token log-probabilities on the host (executor.rs:400-436 `compute_logprobs_from_cpu`): log-softmax of the row in
// f32 with the max subtracted, the sampled token's logprob and the `top_k` largest (descending; among equal values the
// lower index first).  `logits_bf16` is a HOST copy of one bf16 logits row.  Needs no model / GPU.
__attribute__((visibility("default"))) int pq_logprobs_host(const uint16_t* logits_bf16, int n, int sampled_token, int top_k,
                                                            float* sampled_logprob, int* top_ids, float* top_logprobs) {
  if (!logits_bf16 || n <= 0 || sampled_token < 0 || sampled_token >= n || top_k < 0) return -1;
  std::vector<float> x((size_t)n);
  float mx = -INFINITY;
  for (int i = 0; i < n; ++i) {
    const uint32_t u = (uint32_t)logits_bf16[i] << 16;
    memcpy(&x[i], &u, 4);
    mx = std::max(mx, x[i]);
  }
  float sum = 0.f;
  for (int i = 0; i < n; ++i) sum += expf(x[i] - mx);
  const float lse = mx + logf(sum);
  if (sampled_logprob) *sampled_logprob = x[sampled_token] - lse;
  const int k = std::min(top_k, n);
  std::vector<std::pair<int, float>> best;  // sorted descending by value; insertion after equal values
  best.reserve((size_t)k + 1);
  for (int i = 0; i < n && k > 0; ++i) {
    if ((int)best.size() < k || x[i] > best.back().second) {
      size_t pos = 0;
      while (pos < best.size() && best[pos].second >= x[i]) ++pos;
      best.insert(best.begin() + (long)pos, {i, x[i]});
      if ((int)best.size() > k) best.pop_back();
    }
  }
  for (int j = 0; j < (int)best.size(); ++j) {
    if (top_ids) top_ids[j] = best[j].first;
    if (top_logprobs) top_logprobs[j] = best[j].second - lse;
  }
  return (int)best.size();
}

// Greedy generation of one request, host token ids in -> host token ids out, timed like
// pegainfer-server/src/bin/bench_serving.rs:856-892 (TTFT = submit -> first token; the gaps
// between later tokens are the decode steps).  max_tokens = n -> 1 prefill + (n-1) decode steps
// (scheduler.rs:192-200); ignore_eos semantics.
__attribute__((visibility("default"))) int pq_generate(void* mp, const uint32_t* prompt, int n_prompt,
                                                       int max_tokens, uint32_t* out_tokens, double* ttft_ms,
                                                       double* step_ms) {
  Qwen3Model* m = PQ_M;
  using clk = std::chrono::steady_clock;
  const int kv = pq_kv_alloc(mp);
  const auto t0 = clk::now();
  void* lg = nullptr;
  int tok = 0;
  bool ok = m->prefill(1, prompt, &n_prompt, &kv, &lg) && m->sample_greedy(static_cast<const pk_bf16*>(lg), &tok);
  auto tprev = clk::now();
  if (ttft_ms) *ttft_ms = std::chrono::duration<double, std::milli>(tprev - t0).count();
  if (ok && max_tokens > 0) out_tokens[0] = (uint32_t)tok;
  for (int i = 1; ok && i < max_tokens; ++i) {
    const uint32_t cur = (uint32_t)tok;
    ok = m->decode(1, &cur, &kv, nullptr, &tok);
    const auto tn = clk::now();
    if (step_ms) step_ms[i - 1] = std::chrono::duration<double, std::milli>(tn - tprev).count();
    tprev = tn;
    if (ok) out_tokens[i] = (uint32_t)tok;
  }
  pq_kv_free(mp, kv);
  return ok ? 0 : -1;
}

}  // extern "C"
