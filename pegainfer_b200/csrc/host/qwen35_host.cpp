// C++ mirror of the reference's Rust host side for the Qwen3.5 hybrid model (BASELINE config 4):
//   config / layer kinds       pegainfer-qwen35-4b/src/config.rs:4-155
//   weights                    pegainfer-qwen35-4b/src/weights.rs:25-390 (HF `qwen3_5` text-model tensor names)
//   recurrent state            pegainfer-qwen35-4b/src/recurrent_state.rs:9-55  (per request: conv window + fp32 S per linear layer)
//   prefill DAG                pegainfer-qwen35-4b/src/prefill.rs:24-449
//   decode DAG + CUDA graph    pegainfer-qwen35-4b/src/{batch_decode.rs:194-364, batch_decode_graph.rs, decode_buffers.rs}
// 24 linear-attention (gated delta rule) + 8 full-attention (gated, head dim 256, partial RoPE) layers for Qwen3.5-4B.
// Kernels are reached ONLY through the pegainfer-kernels C ABI (include/pegainfer_kernels.h), resolved with dlopen.
//
// What differs from the reference's launch sequence (same arithmetic, same rounding points):
//   * the projections that read the same activation are stacked at load and run as ONE launch with several outputs
//     (linear layers: in_proj_qkv | in_proj_z | in_proj_b;  full layers: q_proj | k_proj | v_proj;  MLP: gate | up)
//   * prefill of the gated delta rule runs the recurrent sequence kernel (one launch per layer, state in registers)
//     where the reference runs its 7-kernel chunk-wise pipeline -- algebraically the same recurrence
//   * one packed metadata block per decode step (token, position, page table), one H2D copy.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>

#include "qwen3_host.hpp"

namespace pq35 {

using pq::DeviceBuf;

struct Config {
  int hidden_size, intermediate_size, num_hidden_layers, vocab_size;
  int num_attention_heads, num_key_value_heads, head_dim;
  int linear_num_key_heads, linear_key_head_dim, linear_num_value_heads, linear_value_head_dim, linear_conv_kernel_dim;
  float rms_norm_eps, rope_theta;
  int rotary_dim;
};

struct Lib {
  void* handle = nullptr;
#define Q35_FN(name) decltype(&::name) name = nullptr;
  Q35_FN(cuda_set_device) Q35_FN(cublas_init) Q35_FN(cublas_destroy) Q35_FN(embedding_batched_cuda) Q35_FN(add_cuda)
  Q35_FN(gemm_cuda) Q35_FN(gemm_graphsafe_cuda) Q35_FN(rms_norm_offset_cuda) Q35_FN(rms_norm_batched_offset_cuda) Q35_FN(rms_norm_gated_cuda)
  Q35_FN(conv1d_prefill_cuda) Q35_FN(gated_delta_rule_decode_cuda) Q35_FN(pk_b200_gated_delta_rule_prefill_recurrent)
  Q35_FN(prefill_attention_hd256_prep_cuda) Q35_FN(attention_gate_batch_hd256_cuda) Q35_FN(qk_norm_partial_rope_batched_decode_hd256_cuda)
  Q35_FN(paged_kv_scatter_cuda) Q35_FN(paged_attention_decode_cuda_hd256) Q35_FN(batch_prefill_paged_cuda_hd256)
  Q35_FN(silu_mul_triton_aot_cuda) Q35_FN(flashinfer_top1_cuda) Q35_FN(pk_b200_gemv_fused) Q35_FN(pk_b200_gemm_segments)
  Q35_FN(pk_b200_set_pdl) Q35_FN(pk_b200_launch_count)
#undef Q35_FN
  std::string load(const std::string& p) {
    handle = dlopen(p.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!handle) return std::string("dlopen failed: ") + dlerror();
    std::string missing;
#define Q35_REQ(name)                                              \
  name = reinterpret_cast<decltype(name)>(dlsym(handle, #name));   \
  if (!name) missing += std::string(" ") + #name;
    Q35_REQ(cuda_set_device) Q35_REQ(cublas_init) Q35_REQ(cublas_destroy) Q35_REQ(embedding_batched_cuda) Q35_REQ(add_cuda)
    Q35_REQ(gemm_cuda) Q35_REQ(gemm_graphsafe_cuda) Q35_REQ(rms_norm_offset_cuda) Q35_REQ(rms_norm_batched_offset_cuda)
    Q35_REQ(rms_norm_gated_cuda) Q35_REQ(conv1d_prefill_cuda) Q35_REQ(gated_delta_rule_decode_cuda)
    Q35_REQ(pk_b200_gated_delta_rule_prefill_recurrent) Q35_REQ(prefill_attention_hd256_prep_cuda)
    Q35_REQ(attention_gate_batch_hd256_cuda) Q35_REQ(qk_norm_partial_rope_batched_decode_hd256_cuda) Q35_REQ(paged_kv_scatter_cuda)
    Q35_REQ(paged_attention_decode_cuda_hd256) Q35_REQ(batch_prefill_paged_cuda_hd256) Q35_REQ(silu_mul_triton_aot_cuda)
    Q35_REQ(flashinfer_top1_cuda) Q35_REQ(pk_b200_gemv_fused) Q35_REQ(pk_b200_gemm_segments) Q35_REQ(pk_b200_set_pdl)
    Q35_REQ(pk_b200_launch_count)
#undef Q35_REQ
    return missing.empty() ? "" : "kernel library " + p + " lacks:" + missing;
  }
  ~Lib() {
    if (handle) dlclose(handle);
  }
};

struct Mat {  // row-major [rows, cols] bf16
  DeviceBuf d;
  int rows = 0, cols = 0;
  pk_bf16* p() const { return d.bf(); }
};
struct Layer {
  bool full = false;
  DeviceBuf in_ln, post_ln;        // [H] bf16
  Mat gate_up, down;               // [2I, H], [H, I]
  // full attention: stacked [q_proj (2 nq hd) ; k_proj ; v_proj], o_proj, norms
  Mat qkv, o;
  DeviceBuf q_norm, k_norm;        // [hd] bf16
  // linear attention: stacked in-projections, conv, dt_bias, A_log (f32), norm (f32), out_proj
  Mat in_qzba, conv_w, out_proj;  // in_qzba = [in_proj_qkv ; in_proj_z ; in_proj_b ; in_proj_a]
  DeviceBuf dt_bias, A_log, gnorm;
  int loaded = 0;
};
struct Request {
  bool live = false;
  std::vector<int> pages;
  int seq_len = 0;
  std::vector<DeviceBuf> conv, S;  // per linear layer
};

static const int kPage = 16, kMaxSeq = 4096;

struct Model {
  Lib k;
  Config c{};
  std::vector<int> layer_full;  // 1 = full attention
  std::string err;
  int device = 0;
  cudaStream_t st = nullptr;
  Mat embed;
  DeviceBuf norm, cosc, sinc;
  std::vector<Layer> layers;
  int n_full = 0, n_lin = 0;
  // KV pool over the full-attention layers only
  DeviceBuf kv;
  int64_t block = 0, layer_stride = 0, page_stride = 0;
  pq::PagePool pool;
  std::vector<Request> reqs;
  bool finalized = false;
  // scratch (decode: 1 token; prefill: grown to T)
  int cap = 0;
  DeviceBuf h, h2, x, big0, big1, big2, conv_o, heads, normed, attn_o, gate, up, act, mo, mo2, zero_res, kc, vc, q_prep, a_seq;
  DeviceBuf logits, meta_d, sample_out, top1_val, top1_states;
  int* meta_h = nullptr;
  int* sample_h = nullptr;
  int meta_ints = 0;
  std::map<int, cudaGraphExec_t> graphs;
  int64_t launches_per_step = 0;
  bool use_graph = true;

  int qkv_dim() const { return 2 * c.linear_num_key_heads * c.linear_key_head_dim + c.linear_num_value_heads * c.linear_value_head_dim; }
  int z_dim() const { return c.linear_num_value_heads * c.linear_value_head_dim; }
  int q_dim() const { return c.num_attention_heads * c.head_dim; }
  int kv_dim() const { return c.num_key_value_heads * c.head_dim; }
  bool fail(const std::string& m) {
    err = m;
    return false;
  }
  bool cu(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
  }
  ~Model() {
    if (st) cudaStreamSynchronize(st);
    for (auto& g : graphs)
      if (g.second) cudaGraphExecDestroy(g.second);
    if (meta_h) cudaFreeHost(meta_h);
    if (sample_h) cudaFreeHost(sample_h);
    if (k.cublas_destroy) k.cublas_destroy();
    if (st) cudaStreamDestroy(st);
  }
  bool load_tensor(const std::string& name, const void* data, int rows, int cols, int is_f32);
  bool finalize(int num_pages);
  bool ensure_scratch(int T);
  bool ensure_pages(Request& r, int tokens);
  bool layer_tail(Layer& L, int T, pk_bf16* hcur, pk_bf16* attn_out_proj, pk_bf16* hnext);
  bool prefill(int rid, const uint32_t* tokens, int n, void** logits_out);
  bool decode_body(Request& r);
  bool decode_body_fused(Request& r);
  bool gemv_f(const pk_bf16* W, const pk_bf16* X, int M, int K, pk_bf16* y0, pk_bf16* y1, pk_bf16* y2, int s0, int s1, int s2,
              const pk_bf16* residual, const pk_bf16* norm_w, pk_bf16* hidden_out, int epi);
  bool decode(int rid, uint32_t token, void** logits_out, int* sampled);
  bool gemv(const pk_bf16* W, const pk_bf16* X, int M, int K, pk_bf16* y0, pk_bf16* y1, pk_bf16* y2, int s0, int s1, int s2);
};

static bool up_rows(Mat& m, int row0, const void* src, int rows, int cols) {
  return cudaMemcpy(m.p() + (size_t)row0 * m.cols, src, (size_t)rows * cols * 2, cudaMemcpyDefault) == cudaSuccess;
}
static bool alloc_mat(Mat& m, int rows, int cols) {
  if (m.d.ptr && m.rows == rows && m.cols == cols) return true;
  m.rows = rows;
  m.cols = cols;
  return m.d.alloc_zeros((size_t)rows * cols * 2);
}
static bool up_vec(DeviceBuf& b, const void* src, size_t bytes) {
  return b.alloc_zeros(bytes) && cudaMemcpy(b.ptr, src, bytes, cudaMemcpyDefault) == cudaSuccess;
}

bool Model::load_tensor(const std::string& name, const void* data, int rows, int cols, int is_f32) {
  const int H = c.hidden_size, I = c.intermediate_size, hd = c.head_dim;
  if ((int)layers.size() != c.num_hidden_layers) {
    layers.resize(c.num_hidden_layers);
    for (int i = 0; i < c.num_hidden_layers; ++i) layers[i].full = layer_full[i] != 0;
  }
  auto want = [&](int er, int ec, int f32) -> bool {
    if (rows == er && cols == ec && is_f32 == f32) return true;
    char b[240];
    snprintf(b, sizeof b, "%s: shape [%d, %d]%s does not match the expected [%d, %d]%s", name.c_str(), rows, cols, is_f32 ? " f32" : "", er, ec,
             f32 ? " f32" : "");
    return fail(b);
  };
  bool ok = true;
  if (name == "model.embed_tokens.weight") {
    if (!want(c.vocab_size, H, 0)) return false;
    ok = alloc_mat(embed, rows, cols) && up_rows(embed, 0, data, rows, cols);
  } else if (name == "lm_head.weight") {
    return true;  // tied (prefill.rs:108-110)
  } else if (name == "model.norm.weight") {
    if (!want(1, H, 0)) return false;
    ok = up_vec(norm, data, (size_t)H * 2);
  } else if (name.rfind("model.layers.", 0) == 0) {
    const size_t dot = name.find('.', 13);
    if (dot == std::string::npos) return fail("unknown tensor " + name);
    const int li = atoi(name.substr(13, dot - 13).c_str());
    if (li < 0 || li >= c.num_hidden_layers) return fail("layer index out of range: " + name);
    Layer& L = layers[li];
    const std::string sub = name.substr(dot + 1);
    const int qf = 2 * q_dim(), kd = kv_dim(), qkvd = qkv_dim(), zd = z_dim(), nv = c.linear_num_value_heads;
    if (sub == "input_layernorm.weight") ok = want(1, H, 0) && up_vec(L.in_ln, data, (size_t)H * 2);
    else if (sub == "post_attention_layernorm.weight") ok = want(1, H, 0) && up_vec(L.post_ln, data, (size_t)H * 2);
    else if (sub == "mlp.gate_proj.weight" || sub == "mlp.up_proj.weight")
      ok = want(I, H, 0) && alloc_mat(L.gate_up, 2 * I, H) && up_rows(L.gate_up, sub[4] == 'g' ? 0 : I, data, I, H);
    else if (sub == "mlp.down_proj.weight") ok = want(H, I, 0) && alloc_mat(L.down, H, I) && up_rows(L.down, 0, data, H, I);
    else if (L.full && sub == "self_attn.q_proj.weight") ok = want(qf, H, 0) && alloc_mat(L.qkv, qf + 2 * kd, H) && up_rows(L.qkv, 0, data, qf, H);
    else if (L.full && sub == "self_attn.k_proj.weight") ok = want(kd, H, 0) && alloc_mat(L.qkv, qf + 2 * kd, H) && up_rows(L.qkv, qf, data, kd, H);
    else if (L.full && sub == "self_attn.v_proj.weight") ok = want(kd, H, 0) && alloc_mat(L.qkv, qf + 2 * kd, H) && up_rows(L.qkv, qf + kd, data, kd, H);
    else if (L.full && sub == "self_attn.o_proj.weight") ok = want(H, q_dim(), 0) && alloc_mat(L.o, H, q_dim()) && up_rows(L.o, 0, data, H, q_dim());
    else if (L.full && sub == "self_attn.q_norm.weight") ok = want(1, hd, 0) && up_vec(L.q_norm, data, (size_t)hd * 2);
    else if (L.full && sub == "self_attn.k_norm.weight") ok = want(1, hd, 0) && up_vec(L.k_norm, data, (size_t)hd * 2);
    else if (!L.full && sub == "linear_attn.in_proj_qkv.weight")
      ok = want(qkvd, H, 0) && alloc_mat(L.in_qzba, qkvd + zd + 2 * nv, H) && up_rows(L.in_qzba, 0, data, qkvd, H);
    else if (!L.full && sub == "linear_attn.in_proj_z.weight")
      ok = want(zd, H, 0) && alloc_mat(L.in_qzba, qkvd + zd + 2 * nv, H) && up_rows(L.in_qzba, qkvd, data, zd, H);
    else if (!L.full && sub == "linear_attn.in_proj_b.weight")
      ok = want(nv, H, 0) && alloc_mat(L.in_qzba, qkvd + zd + 2 * nv, H) && up_rows(L.in_qzba, qkvd + zd, data, nv, H);
    else if (!L.full && sub == "linear_attn.in_proj_a.weight")
      ok = want(nv, H, 0) && alloc_mat(L.in_qzba, qkvd + zd + 2 * nv, H) && up_rows(L.in_qzba, qkvd + zd + nv, data, nv, H);
    else if (!L.full && sub == "linear_attn.conv1d.weight")
      ok = want(qkvd, c.linear_conv_kernel_dim, 0) && alloc_mat(L.conv_w, qkvd, c.linear_conv_kernel_dim) &&
           up_rows(L.conv_w, 0, data, qkvd, c.linear_conv_kernel_dim);
    else if (!L.full && sub == "linear_attn.dt_bias") ok = want(1, nv, 0) && up_vec(L.dt_bias, data, (size_t)nv * 2);
    else if (!L.full && sub == "linear_attn.A_log") ok = want(1, nv, 1) && up_vec(L.A_log, data, (size_t)nv * 4);
    else if (!L.full && sub == "linear_attn.norm.weight")
      ok = want(1, c.linear_value_head_dim, 1) && up_vec(L.gnorm, data, (size_t)c.linear_value_head_dim * 4);
    else if (!L.full && sub == "linear_attn.out_proj.weight") ok = want(H, zd, 0) && alloc_mat(L.out_proj, H, zd) && up_rows(L.out_proj, 0, data, H, zd);
    else return fail("unknown tensor " + name + (L.full ? " (full-attention layer)" : " (linear-attention layer)"));
    if (ok) L.loaded++;
  } else {
    return fail("unknown tensor " + name);
  }
  if (!ok && err.find("does not match") == std::string::npos)
    return fail("upload failed for " + name + ": " + cudaGetErrorString(cudaGetLastError()));
  return ok;
}

static uint16_t f2bf_h(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fff;
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

bool Model::finalize(int num_pages) {
  for (int i = 0; i < c.num_hidden_layers; ++i) {
    const int need = layers[i].full ? 11 : 14;
    if (layers[i].loaded != need) return fail("layer " + std::to_string(i) + " is missing tensors (" + std::to_string(layers[i].loaded) + "/" + std::to_string(need) + ")");
  }
  if (!embed.d.ptr || !norm.ptr) return fail("embed_tokens / norm missing");
  // RoPE tables over the rotary slice (weights.rs: precompute_rope(rotary_dim, 4096, theta)), bf16, half-split duplicated
  {
    const int rd = c.rotary_dim, half = rd / 2;
    std::vector<uint16_t> cs((size_t)kMaxSeq * rd), sn((size_t)kMaxSeq * rd);
    for (int pos = 0; pos < kMaxSeq; ++pos)
      for (int i = 0; i < half; ++i) {
        const float inv = 1.0f / powf(c.rope_theta, (float)i * 2.0f / (float)rd);
        const float fr = (float)pos * inv;
        const uint16_t cv = f2bf_h(cosf(fr)), sv = f2bf_h(sinf(fr));
        cs[(size_t)pos * rd + i] = cs[(size_t)pos * rd + i + half] = cv;
        sn[(size_t)pos * rd + i] = sn[(size_t)pos * rd + i + half] = sv;
      }
    if (!up_vec(cosc, cs.data(), cs.size() * 2) || !up_vec(sinc, sn.data(), sn.size() * 2)) return fail("rope upload failed");
  }
  n_full = n_lin = 0;
  for (int f : layer_full) (f ? n_full : n_lin)++;
  block = (int64_t)kPage * c.num_key_value_heads * c.head_dim;
  layer_stride = 2 * block;
  page_stride = (int64_t)std::max(1, n_full) * layer_stride;
  if (num_pages <= 0) num_pages = 2048;
  if (!kv.alloc_zeros((size_t)num_pages * page_stride * 2)) return fail("KV pool allocation failed");
  pool.init(num_pages);
  std::vector<int> pad;
  pool.acquire(1, &pad);
  meta_ints = 4 + kMaxSeq / kPage + 8;  // [0] token, [1] position, [2..3] page_indptr, [4] last_page_len, [5] zero, then the page table
  if (!meta_d.alloc_zeros((size_t)meta_ints * 4) || cudaMallocHost((void**)&meta_h, (size_t)meta_ints * 4) != cudaSuccess ||
      cudaMallocHost((void**)&sample_h, 64) != cudaSuccess)
    return fail("metadata allocation failed");
  if (!logits.alloc_zeros((size_t)c.vocab_size * 2) || !sample_out.alloc_zeros(64) || !top1_val.alloc_zeros(64) || !top1_states.alloc_zeros(1 << 16) ||
      !kc.alloc_zeros((size_t)c.num_key_value_heads * kMaxSeq * c.head_dim * 2) ||
      !vc.alloc_zeros((size_t)c.num_key_value_heads * kMaxSeq * c.head_dim * 2))
    return fail("buffer allocation failed");
  if (!ensure_scratch(1)) return false;
  finalized = true;
  return cu(cudaStreamSynchronize(st), "finalize sync");
}

bool Model::ensure_scratch(int T) {
  if (T <= cap) return true;
  cudaStreamSynchronize(st);
  const int n = ((T + 63) / 64) * 64;
  const size_t H = c.hidden_size, I = c.intermediate_size;
  const size_t wide = std::max(std::max((size_t)2 * q_dim(), (size_t)qkv_dim()), I);
  bool ok = h.alloc_uninit(H * n * 2) && h2.alloc_uninit(H * n * 2) && x.alloc_uninit(H * n * 2) && big0.alloc_uninit(wide * n * 2) &&
            big1.alloc_uninit(std::max<size_t>(z_dim(), kv_dim()) * n * 2) && big2.alloc_uninit(std::max<size_t>(kv_dim(), 64) * n * 2) &&
            conv_o.alloc_uninit((size_t)qkv_dim() * n * 2) && heads.alloc_uninit((size_t)z_dim() * n * 2) &&
            normed.alloc_uninit((size_t)z_dim() * n * 2) && attn_o.alloc_uninit((size_t)q_dim() * n * 2) && gate.alloc_uninit(I * n * 2) &&
            up.alloc_uninit(I * n * 2) && act.alloc_uninit(I * n * 2) && mo.alloc_uninit(H * n * 2) && mo2.alloc_uninit(H * n * 2) &&
            zero_res.alloc_zeros(H * 2) && q_prep.alloc_uninit((size_t)q_dim() * n * 2) &&
            a_seq.alloc_uninit((size_t)std::max(64, c.linear_num_value_heads) * n * 2);
  if (!ok) return fail("scratch allocation failed");
  cap = n;
  for (auto& g : graphs)  // captured decode graphs hold the old scratch pointers
    if (g.second) cudaGraphExecDestroy(g.second);
  graphs.clear();
  return true;
}

bool Model::ensure_pages(Request& r, int tokens) {
  const int need = (tokens + kPage - 1) / kPage - (int)r.pages.size();
  if (need <= 0) return true;
  if (!pool.acquire(need, &r.pages)) return fail("Qwen3.5 KvState: out of pages");
  return true;
}

bool Model::gemv(const pk_bf16* W, const pk_bf16* X, int M, int K, pk_bf16* y0, pk_bf16* y1, pk_bf16* y2, int s0, int s1, int s2) {
  pk_b200_gemv_args g{};
  g.W = W; g.X = X;
  g.Y[0] = y0; g.Y[1] = y1; g.Y[2] = y2;
  g.seg_rows[0] = s0; g.seg_rows[1] = s1; g.seg_rows[2] = s2;
  g.M = M; g.N = 1; g.K = K;
  if (k.pk_b200_gemv_fused(&g, st) != 0) {  // shapes the streaming kernel does not take: one GEMV per segment
    const int segs[3] = {s0, s1, s2};
    pk_bf16* ys[3] = {y0, y1, y2};
    int r0 = 0;
    for (int i = 0; i < 3; ++i) {
      if (segs[i] > 0) k.gemm_graphsafe_cuda(W + (size_t)r0 * K, X, ys[i], segs[i], 1, K, st);
      r0 += segs[i];
    }
  }
  return true;
}

// post-attention half of a layer (prefill.rs:160-188 / batch_decode.rs:300-364): h' = h + attn; x = norm(h'); MLP; h'' = h' + mlp
bool Model::layer_tail(Layer& L, int T, pk_bf16* hcur, pk_bf16* attn, pk_bf16* hnext) {
  const int H = c.hidden_size, I = c.intermediate_size;
  k.add_cuda(hcur, attn, hnext, H * T, st);
  k.rms_norm_batched_offset_cuda(hnext, L.post_ln.bf(), x.bf(), H, T, c.rms_norm_eps, st);
  if (T == 1) {
    gemv(L.gate_up.p(), x.bf(), 2 * I, H, gate.bf(), up.bf(), nullptr, I, I, 0);
  } else {
    pk_bf16* outs[3] = {gate.bf(), up.bf(), up.bf()};
    const int segs[3] = {I, I, 0};
    if (k.pk_b200_gemm_segments(L.gate_up.p(), x.bf(), outs, segs, 2 * I, T, H, st) != 0) return fail("gate_up GEMM failed");
  }
  if (k.silu_mul_triton_aot_cuda(gate.bf(), up.bf(), act.bf(), I * T, st) != 0) return fail("silu_mul failed");  // SiLU rounded to bf16 first
  if (T == 1) k.gemm_graphsafe_cuda(L.down.p(), act.bf(), mo.bf(), H, 1, I, st);
  else k.gemm_cuda(L.down.p(), act.bf(), mo.bf(), H, T, I, st);
  k.add_cuda(hnext, mo.bf(), hcur, H * T, st);  // result back in hcur
  return true;
}

bool Model::prefill(int rid, const uint32_t* tokens, int n, void** logits_out) {
  if (!finalized) return fail("model not finalized");
  if (rid < 0 || rid >= (int)reqs.size() || !reqs[rid].live) return fail("bad request id");
  if (n <= 0) return fail("empty prompt");
  Request& r = reqs[rid];
  if (r.seq_len + n > kMaxSeq) return fail("position beyond the 4096-entry RoPE table");
  const int old_pages = (int)r.pages.size();
  if (!ensure_pages(r, r.seq_len + n)) return false;
  if (!ensure_scratch(n)) {
    std::vector<int> extra(r.pages.begin() + old_pages, r.pages.end());
    pool.release(extra);
    r.pages.resize(old_pages);
    return false;
  }
  const int T = n, start = r.seq_len;
  const int H = c.hidden_size, hd = c.head_dim, nq = c.num_attention_heads, nkv = c.num_key_value_heads;
  const int nk = c.linear_num_key_heads, nv = c.linear_num_value_heads, dk = c.linear_key_head_dim, dv = c.linear_value_head_dim;
  const int qf = 2 * q_dim(), kd = kv_dim(), qkvd = qkv_dim(), zd = z_dim();
  const float eps = c.rms_norm_eps;
  // plan block: [tokens T][positions T][batch idx T (zeros)][page_indptr 2][last_page_len 1][q_indptr 2][start 1][total rows 1][pages]
  std::vector<int> plan;
  plan.insert(plan.end(), reinterpret_cast<const int*>(tokens), reinterpret_cast<const int*>(tokens) + T);
  const int o_pos = (int)plan.size();
  for (int t = 0; t < T; ++t) plan.push_back(start + t);
  const int o_bi = (int)plan.size();
  plan.insert(plan.end(), T, 0);
  const int o_ip = (int)plan.size();
  plan.push_back(0);
  plan.push_back((int)r.pages.size());
  const int o_lpl = (int)plan.size();
  plan.push_back(((start + T - 1) % kPage) + 1);
  const int o_qi = (int)plan.size();
  plan.push_back(0);
  plan.push_back(T);
  const int o_start = (int)plan.size();
  plan.push_back(start);
  const int o_tn = (int)plan.size();
  plan.push_back(T);
  const int o_zero = (int)plan.size();
  plan.insert(plan.end(), 8, 0);
  const int o_pi = (int)plan.size();
  plan.insert(plan.end(), r.pages.begin(), r.pages.end());
  DeviceBuf plan_d;
  if (!plan_d.alloc_uninit(plan.size() * 4) ||
      !cu(cudaMemcpyAsync(plan_d.ptr, plan.data(), plan.size() * 4, cudaMemcpyHostToDevice, st), "plan H2D"))
    return false;
  const int* P = plan_d.i32();
  k.embedding_batched_cuda(embed.p(), reinterpret_cast<const uint32_t*>(P), h.bf(), H, T, st);
  int fi = 0, lin = 0;
  for (int li = 0; li < c.num_hidden_layers; ++li) {
    Layer& L = layers[li];
    k.rms_norm_batched_offset_cuda(h.bf(), L.in_ln.bf(), x.bf(), H, T, eps, st);
    pk_bf16* attn = nullptr;
    if (L.full) {
      // q_full | k | v in one launch (prefill.rs:204-206), then prep -> scatter -> paged HD-256 attention -> sigmoid gate -> o_proj
      pk_bf16* outs[3] = {big0.bf(), big1.bf(), big2.bf()};
      const int segs[3] = {qf, kd, kd};
      if (T == 1) gemv(L.qkv.p(), x.bf(), qf + 2 * kd, H, outs[0], outs[1], outs[2], qf, kd, kd);
      else if (k.pk_b200_gemm_segments(L.qkv.p(), x.bf(), outs, segs, qf + 2 * kd, T, H, st) != 0) return fail("qkv GEMM failed");
      k.prefill_attention_hd256_prep_cuda(big0.bf(), big1.bf(), big2.bf(), L.q_norm.bf(), L.k_norm.bf(), cosc.bf(), sinc.bf(), q_prep.bf(),
                                          kc.bf(), vc.bf(), nq, nkv, T, P + o_start, c.rotary_dim, eps, kMaxSeq, st);
      const int64_t k_off = (int64_t)fi * layer_stride, v_off = k_off + block;
      if (k.paged_kv_scatter_cuda(kv.bf(), k_off, v_off, P + o_pi, P + o_ip, P + o_lpl, kc.bf() + (size_t)start * hd, vc.bf() + (size_t)start * hd,
                                  P + o_bi, P + o_pos, T, nkv, hd, kPage, page_stride, hd, (int64_t)kMaxSeq * hd, st) != 0)
        return fail("paged_kv_scatter_cuda failed");
      if (k.batch_prefill_paged_cuda_hd256(q_prep.bf(), attn_o.bf(), kv.bf(), k_off, v_off, P + o_pi, P + o_ip, P + o_lpl, P + o_qi, P + o_zero,
                                           P + o_zero, P + o_zero, P + o_zero, reinterpret_cast<const uint32_t*>(P + o_tn), nq, nkv, hd, kPage, T, 1, 1,
                                           page_stride, 1.0f / sqrtf((float)hd), st) != 0)
        return fail("batch_prefill_paged_cuda_hd256 failed");
      k.attention_gate_batch_hd256_cuda(big0.bf(), attn_o.bf(), nq, T, st);
      if (T == 1) k.gemm_graphsafe_cuda(L.o.p(), attn_o.bf(), mo.bf(), H, 1, q_dim(), st);
      else k.gemm_cuda(L.o.p(), attn_o.bf(), mo.bf(), H, T, q_dim(), st);
      attn = mo.bf();
      ++fi;
    } else {
      // in_proj_qkv | in_proj_z | in_proj_b in one launch + in_proj_a (prefill.rs: prefill_linear_attention), conv1d + SiLU,
      // gated delta rule over the sequence, gated RMSNorm, out_proj
      pk_bf16* outs[3] = {big0.bf(), big1.bf(), big2.bf()};
      const int segs[3] = {qkvd, zd, nv};
      const pk_bf16* w_a = L.in_qzba.p() + (size_t)(qkvd + zd + nv) * H;  // the in_proj_a rows of the stacked matrix
      if (T == 1) {
        gemv(L.in_qzba.p(), x.bf(), qkvd + zd + nv, H, outs[0], outs[1], outs[2], qkvd, zd, nv);
        k.gemm_graphsafe_cuda(w_a, x.bf(), a_seq.bf(), nv, 1, H, st);
      } else {
        if (k.pk_b200_gemm_segments(L.in_qzba.p(), x.bf(), outs, segs, qkvd + zd + nv, T, H, st) != 0) return fail("in_proj GEMM failed");
        k.gemm_cuda(w_a, x.bf(), a_seq.bf(), nv, T, H, st);
      }
      k.conv1d_prefill_cuda(big0.bf(), L.conv_w.p(), r.conv[lin].bf(), conv_o.bf(), qkvd, T, c.linear_conv_kernel_dim, st);
      if (k.pk_b200_gated_delta_rule_prefill_recurrent(conv_o.bf(), big2.bf(), a_seq.bf(), L.dt_bias.bf(), static_cast<const float*>(L.A_log.ptr),
                                                       static_cast<float*>(r.S[lin].ptr), heads.bf(), nk, nv, dk, dv, T, st) != 0)
        return fail("gated delta rule prefill failed");
      k.rms_norm_gated_cuda(heads.bf(), static_cast<const float*>(L.gnorm.ptr), big1.bf(), normed.bf(), nv * T, dv, eps, st);
      if (T == 1) k.gemm_graphsafe_cuda(L.out_proj.p(), normed.bf(), mo.bf(), H, 1, zd, st);
      else k.gemm_cuda(L.out_proj.p(), normed.bf(), mo.bf(), H, T, zd, st);
      attn = mo.bf();
      ++lin;
    }
    // layer_tail reads `attn` (= mo) before it overwrites mo with the MLP output
    if (!layer_tail(L, T, h.bf(), attn, h2.bf())) return false;
  }
  k.rms_norm_offset_cuda(h.bf() + (size_t)(T - 1) * H, norm.bf(), x.bf(), H, eps, st);
  k.gemm_graphsafe_cuda(embed.p(), x.bf(), logits.bf(), c.vocab_size, 1, H, st);
  if (!cu(cudaStreamSynchronize(st), "prefill sync")) {
    return false;
  }
  r.seq_len += T;
  if (logits_out) *logits_out = logits.ptr;
  return true;
}

// one decode token of one request; launched under capture or directly (batch_decode.rs:194-364 for batch size 1)
bool Model::decode_body(Request& r) {
  const int H = c.hidden_size, hd = c.head_dim, nq = c.num_attention_heads, nkv = c.num_key_value_heads;
  const int nk = c.linear_num_key_heads, nv = c.linear_num_value_heads, dk = c.linear_key_head_dim, dv = c.linear_value_head_dim;
  const int qf = 2 * q_dim(), kd = kv_dim(), qkvd = qkv_dim(), zd = z_dim();
  const float eps = c.rms_norm_eps;
  const int* M = meta_d.i32();
  k.embedding_batched_cuda(embed.p(), reinterpret_cast<const uint32_t*>(M), h.bf(), H, 1, st);
  int fi = 0, lin = 0;
  for (int li = 0; li < c.num_hidden_layers; ++li) {
    Layer& L = layers[li];
    k.rms_norm_offset_cuda(h.bf(), L.in_ln.bf(), x.bf(), H, eps, st);
    if (L.full) {
      gemv(L.qkv.p(), x.bf(), qf + 2 * kd, H, big0.bf(), big1.bf(), big2.bf(), qf, kd, kd);
      k.qk_norm_partial_rope_batched_decode_hd256_cuda(big0.bf(), big1.bf(), L.q_norm.bf(), L.k_norm.bf(), cosc.bf(), sinc.bf(), M + 1, q_prep.bf(),
                                                       nq, nkv, 1, c.rotary_dim, eps, st);
      const int64_t k_off = (int64_t)fi * layer_stride, v_off = k_off + block;
      if (k.paged_kv_scatter_cuda(kv.bf(), k_off, v_off, M + 6, M + 2, M + 4, big1.bf(), big2.bf(), M + 5, M + 1, 1, nkv, hd, kPage, page_stride,
                                  (int64_t)nkv * hd, hd, st) != 0)
        return fail("paged_kv_scatter_cuda failed");
      if (k.paged_attention_decode_cuda_hd256(q_prep.bf(), attn_o.bf(), kv.bf(), k_off, v_off, M + 6, M + 2, M + 4, M + 5, M + 5, M + 5, nq, nkv, hd,
                                              kPage, 1, page_stride, 1.0f / sqrtf((float)hd), st) != 0)
        return fail("paged_attention_decode_cuda_hd256 failed");
      k.attention_gate_batch_hd256_cuda(big0.bf(), attn_o.bf(), nq, 1, st);
      k.gemm_graphsafe_cuda(L.o.p(), attn_o.bf(), mo.bf(), H, 1, q_dim(), st);
      ++fi;
    } else {
      gemv(L.in_qzba.p(), x.bf(), qkvd + zd + nv, H, big0.bf(), big1.bf(), big2.bf(), qkvd, zd, nv);
      k.gemm_graphsafe_cuda(L.in_qzba.p() + (size_t)(qkvd + zd + nv) * H, x.bf(), a_seq.bf(), nv, 1, H, st);
      k.conv1d_prefill_cuda(big0.bf(), L.conv_w.p(), r.conv[lin].bf(), conv_o.bf(), qkvd, 1, c.linear_conv_kernel_dim, st);
      k.gated_delta_rule_decode_cuda(conv_o.bf(), big2.bf(), a_seq.bf(), L.dt_bias.bf(), static_cast<const float*>(L.A_log.ptr),
                                     static_cast<float*>(r.S[lin].ptr), heads.bf(), nk, nv, dk, dv, st);
      k.rms_norm_gated_cuda(heads.bf(), static_cast<const float*>(L.gnorm.ptr), big1.bf(), normed.bf(), nv, dv, eps, st);
      k.gemm_graphsafe_cuda(L.out_proj.p(), normed.bf(), mo.bf(), H, 1, zd, st);
      ++lin;
    }
    if (!layer_tail(L, 1, h.bf(), mo.bf(), h2.bf())) return false;
  }
  k.rms_norm_offset_cuda(h.bf(), norm.bf(), x.bf(), H, eps, st);
  k.gemm_graphsafe_cuda(embed.p(), x.bf(), logits.bf(), c.vocab_size, 1, H, st);
  k.flashinfer_top1_cuda(logits.bf(), static_cast<pk_bf16*>(top1_val.ptr), static_cast<uint8_t*>(top1_states.ptr), sample_out.i32(), c.vocab_size, st);
  return true;
}

// GEMV with the Qwen3.5 prologue (x_mode 3: hidden_out = bf16(X + residual), x = (1 + w) RMSNorm of that rounded sum) and,
// for epi 4, the rounded-SiLU SwiGLU epilogue
bool Model::gemv_f(const pk_bf16* W, const pk_bf16* X, int M, int K, pk_bf16* y0, pk_bf16* y1, pk_bf16* y2, int s0, int s1, int s2,
                   const pk_bf16* residual, const pk_bf16* norm_w, pk_bf16* hidden_out, int epi) {
  pk_b200_gemv_args g{};
  g.W = W; g.X = X;
  g.Y[0] = y0; g.Y[1] = y1; g.Y[2] = y2;
  g.seg_rows[0] = s0; g.seg_rows[1] = s1; g.seg_rows[2] = s2;
  g.M = M; g.N = 1; g.K = K;
  g.x_mode = 3; g.residual = residual; g.norm_w = norm_w; g.eps = c.rms_norm_eps; g.hidden_out = hidden_out; g.epi = epi;
  if (k.pk_b200_gemv_fused(&g, st) != 0) return fail("pk_b200_gemv_fused rejected its arguments");
  return true;
}

// Fused decode token (default): the two adds + two (1+w) norms of a layer ride in the prologues of the in-projection
// and gate_up GEMVs, SiLU-mul in the gate_up epilogue, b and a come out of the in-projection launch: 8 launches per
// layer instead of 14, same rounding points (PK_Q35_FUSED=0 keeps the launch-per-op sequence).
bool Model::decode_body_fused(Request& r) {
  const int H = c.hidden_size, I = c.intermediate_size, hd = c.head_dim, nq = c.num_attention_heads, nkv = c.num_key_value_heads;
  const int nk = c.linear_num_key_heads, nv = c.linear_num_value_heads, dk = c.linear_key_head_dim, dv = c.linear_value_head_dim;
  const int qf = 2 * q_dim(), kd = kv_dim(), qkvd = qkv_dim(), zd = z_dim();
  const float eps = c.rms_norm_eps;
  const int* M = meta_d.i32();
  pk_bf16 *Ha = h.bf(), *Hb = h2.bf();
  k.embedding_batched_cuda(embed.p(), reinterpret_cast<const uint32_t*>(M), Ha, H, 1, st);
  const pk_bf16* prev = zero_res.bf();  // layer 0: hidden + 0
  int fi = 0, lin = 0;
  for (int li = 0; li < c.num_hidden_layers; ++li) {
    Layer& L = layers[li];
    if (L.full) {
      // Hb = Ha + prev; q_full | k | v = W . norm(Hb)
      if (!gemv_f(L.qkv.p(), Ha, qf + 2 * kd, H, big0.bf(), big1.bf(), big2.bf(), qf, kd, kd, prev, L.in_ln.bf(), Hb, 0)) return false;
      k.qk_norm_partial_rope_batched_decode_hd256_cuda(big0.bf(), big1.bf(), L.q_norm.bf(), L.k_norm.bf(), cosc.bf(), sinc.bf(), M + 1, q_prep.bf(),
                                                       nq, nkv, 1, c.rotary_dim, eps, st);
      const int64_t k_off = (int64_t)fi * layer_stride, v_off = k_off + block;
      if (k.paged_kv_scatter_cuda(kv.bf(), k_off, v_off, M + 6, M + 2, M + 4, big1.bf(), big2.bf(), M + 5, M + 1, 1, nkv, hd, kPage, page_stride,
                                  (int64_t)nkv * hd, hd, st) != 0)
        return fail("paged_kv_scatter_cuda failed");
      if (k.paged_attention_decode_cuda_hd256(q_prep.bf(), attn_o.bf(), kv.bf(), k_off, v_off, M + 6, M + 2, M + 4, M + 5, M + 5, M + 5, nq, nkv, hd,
                                              kPage, 1, page_stride, 1.0f / sqrtf((float)hd), st) != 0)
        return fail("paged_attention_decode_cuda_hd256 failed");
      k.attention_gate_batch_hd256_cuda(big0.bf(), attn_o.bf(), nq, 1, st);
      k.gemm_graphsafe_cuda(L.o.p(), attn_o.bf(), mo.bf(), H, 1, q_dim(), st);
      ++fi;
    } else {
      // Hb = Ha + prev; qkv | z | [b; a] = W . norm(Hb)
      if (!gemv_f(L.in_qzba.p(), Ha, qkvd + zd + 2 * nv, H, big0.bf(), big1.bf(), big2.bf(), qkvd, zd, 2 * nv, prev, L.in_ln.bf(), Hb, 0)) return false;
      k.conv1d_prefill_cuda(big0.bf(), L.conv_w.p(), r.conv[lin].bf(), conv_o.bf(), qkvd, 1, c.linear_conv_kernel_dim, st);
      k.gated_delta_rule_decode_cuda(conv_o.bf(), big2.bf(), big2.bf() + nv, L.dt_bias.bf(), static_cast<const float*>(L.A_log.ptr),
                                     static_cast<float*>(r.S[lin].ptr), heads.bf(), nk, nv, dk, dv, st);
      k.rms_norm_gated_cuda(heads.bf(), static_cast<const float*>(L.gnorm.ptr), big1.bf(), normed.bf(), nv, dv, eps, st);
      k.gemm_graphsafe_cuda(L.out_proj.p(), normed.bf(), mo.bf(), H, 1, zd, st);
      ++lin;
    }
    // Ha = Hb + attn; act = bf16(bf16(silu(gate)) * up) with gate | up = W . norm(Ha)
    if (!gemv_f(L.gate_up.p(), Hb, I, H, act.bf(), nullptr, nullptr, I, 0, 0, mo.bf(), L.post_ln.bf(), Ha, 4)) return false;
    k.gemm_graphsafe_cuda(L.down.p(), act.bf(), mo2.bf(), H, 1, I, st);
    prev = mo2.bf();
  }
  // logits = embed . norm(Ha + mlp)
  if (!gemv_f(embed.p(), Ha, c.vocab_size, H, logits.bf(), nullptr, nullptr, c.vocab_size, 0, 0, prev, norm.bf(), Hb, 0)) return false;
  k.flashinfer_top1_cuda(logits.bf(), static_cast<pk_bf16*>(top1_val.ptr), static_cast<uint8_t*>(top1_states.ptr), sample_out.i32(), c.vocab_size, st);
  return true;
}

bool Model::decode(int rid, uint32_t token, void** logits_out, int* sampled) {
  if (!finalized) return fail("model not finalized");
  if (rid < 0 || rid >= (int)reqs.size() || !reqs[rid].live) return fail("bad request id");
  Request& r = reqs[rid];
  if (r.seq_len + 1 > kMaxSeq) return fail("position beyond the 4096-entry RoPE table");
  if (!ensure_pages(r, r.seq_len + 1)) return false;
  const int pos = r.seq_len;
  memset(meta_h, 0, (size_t)meta_ints * 4);
  meta_h[0] = (int)token;
  meta_h[1] = pos;
  meta_h[2] = 0;
  meta_h[3] = (int)r.pages.size();
  meta_h[4] = (pos % kPage) + 1;
  meta_h[5] = 0;
  memcpy(meta_h + 6, r.pages.data(), r.pages.size() * 4);
  if (!cu(cudaMemcpyAsync(meta_d.ptr, meta_h, (size_t)meta_ints * 4, cudaMemcpyHostToDevice, st), "meta H2D")) return false;
  const char* fe = getenv("PK_Q35_FUSED");  // read per call: the tests flip it between models
  const bool fused = !(fe && atoi(fe) == 0);
  auto body = [&]() { return fused ? decode_body_fused(r) : decode_body(r); };
  if (!use_graph) {
    if (!body()) return false;
  } else {
    cudaGraphExec_t& g = graphs[rid];  // the graph bakes this request's recurrent-state pointers
    if (!g) {
      const int64_t before = k.pk_b200_launch_count(0);
      if (!cu(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal), "begin capture")) return false;
      const bool ok = body();
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (!ok) {
        if (graph) cudaGraphDestroy(graph);
        return false;
      }
      if (!cu(e, "end capture")) return false;
      e = cudaGraphInstantiate(&g, graph, 0);
      cudaGraphDestroy(graph);
      if (!cu(e, "graph instantiate")) return false;
      launches_per_step = k.pk_b200_launch_count(0) - before;
    }
    if (!cu(cudaGraphLaunch(g, st), "graph launch")) return false;
  }
  r.seq_len += 1;
  if (logits_out) *logits_out = logits.ptr;
  if (sampled) {
    if (!cu(cudaMemcpyAsync(sample_h, sample_out.ptr, 4, cudaMemcpyDeviceToHost, st), "sample D2H")) return false;
    if (!cu(cudaStreamSynchronize(st), "decode sync")) return false;
    *sampled = sample_h[0];
  }
  return true;
}

}  // namespace pq35

// ===================================================================== C API
extern "C" {

struct pq35_config {
  int hidden_size, intermediate_size, num_hidden_layers, vocab_size, num_attention_heads, num_key_value_heads, head_dim;
  int linear_num_key_heads, linear_key_head_dim, linear_num_value_heads, linear_value_head_dim, linear_conv_kernel_dim;
  float rms_norm_eps, rope_theta;
  int rotary_dim, enable_cuda_graph, enable_pdl, device_ordinal;
};
static thread_local std::string g35_err;
#define V35 __attribute__((visibility("default")))
#define M35 static_cast<pq35::Model*>(mp)

V35 const char* pq35_create_error() { return g35_err.c_str(); }
V35 void* pq35_create(const pq35_config* c, const int* layer_is_full, const char* kernel_lib) {
  auto* m = new pq35::Model();
  auto bail = [&](const std::string& e) -> void* {
    g35_err = e;
    delete m;
    return nullptr;
  };
  m->c = pq35::Config{c->hidden_size, c->intermediate_size, c->num_hidden_layers, c->vocab_size, c->num_attention_heads, c->num_key_value_heads,
                      c->head_dim, c->linear_num_key_heads, c->linear_key_head_dim, c->linear_num_value_heads, c->linear_value_head_dim,
                      c->linear_conv_kernel_dim, c->rms_norm_eps, c->rope_theta, c->rotary_dim};
  m->layer_full.assign(layer_is_full, layer_is_full + c->num_hidden_layers);
  m->use_graph = c->enable_cuda_graph != 0;
  if (c->head_dim != 256 || c->linear_key_head_dim != 128 || c->linear_value_head_dim != 128)
    return bail("Qwen3.5 kernels are instantiated for head_dim 256 and 128 x 128 linear heads (as the reference's)");
  if (c->num_attention_heads != 4 * c->num_key_value_heads) return bail("HD-256 decode attention is instantiated for GQA group 4");
  std::string e = m->k.load(kernel_lib);
  if (!e.empty()) return bail(e);
  if (m->k.cuda_set_device(c->device_ordinal) != 0) return bail("cuda_set_device failed (no CUDA device?)");
  m->device = c->device_ordinal;
  if (cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking) != cudaSuccess) return bail("stream creation failed");
  m->k.cublas_init();
  m->k.pk_b200_set_pdl(c->enable_pdl);
  g35_err.clear();
  return m;
}
V35 void pq35_destroy(void* mp) { delete M35; }
V35 const char* pq35_last_error(void* mp) { return M35->err.c_str(); }
V35 int pq35_load_tensor(void* mp, const char* name, const void* data, int rows, int cols, int is_f32) {
  return M35->load_tensor(name, data, rows, cols, is_f32) ? 0 : -1;
}
V35 int pq35_finalize(void* mp, int num_pages) { return M35->finalize(num_pages) ? 0 : -1; }
V35 int pq35_request_alloc(void* mp) {
  pq35::Model* m = M35;
  int id = -1;
  for (size_t i = 0; i < m->reqs.size(); ++i)
    if (!m->reqs[i].live) {
      id = (int)i;
      break;
    }
  if (id < 0) {
    m->reqs.emplace_back();
    id = (int)m->reqs.size() - 1;
  }
  pq35::Request& r = m->reqs[id];
  r = pq35::Request();
  r.live = true;
  const size_t conv_bytes = (size_t)m->qkv_dim() * (m->c.linear_conv_kernel_dim - 1) * 2;
  const size_t s_bytes = (size_t)m->c.linear_num_value_heads * m->c.linear_key_head_dim * m->c.linear_value_head_dim * 4;
  r.conv.resize(m->n_lin);
  r.S.resize(m->n_lin);
  for (int i = 0; i < m->n_lin; ++i)
    if (!r.conv[i].alloc_zeros(conv_bytes) || !r.S[i].alloc_zeros(s_bytes)) {
      r.live = false;
      m->err = "recurrent state allocation failed";
      return -1;
    }
  auto it = m->graphs.find(id);  // a recycled slot has new state buffers: its captured graph is stale
  if (it != m->graphs.end()) {
    if (it->second) cudaGraphExecDestroy(it->second);
    m->graphs.erase(it);
  }
  return id;
}
V35 void pq35_request_free(void* mp, int id) {
  pq35::Model* m = M35;
  if (id < 0 || id >= (int)m->reqs.size() || !m->reqs[id].live) return;
  cudaStreamSynchronize(m->st);
  m->pool.release(m->reqs[id].pages);
  m->reqs[id] = pq35::Request();
}
V35 int pq35_seq_len(void* mp, int id) { return (id < 0 || id >= (int)M35->reqs.size()) ? -1 : M35->reqs[id].seq_len; }
V35 int pq35_prefill(void* mp, int id, const uint32_t* tokens, int n, void** logits_out) { return M35->prefill(id, tokens, n, logits_out) ? 0 : -1; }
V35 int pq35_decode(void* mp, int id, uint32_t token, void** logits_out, int* sampled) { return M35->decode(id, token, logits_out, sampled) ? 0 : -1; }
V35 int pq35_copy_out(void* mp, void* dst, const void* src, int64_t bytes) {
  if (cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, M35->st) != cudaSuccess) return -1;
  return cudaStreamSynchronize(M35->st) == cudaSuccess ? 0 : -1;
}
V35 int64_t pq35_launches_per_step(void* mp) { return M35->launches_per_step; }
// Greedy generation timed like bench_serving.rs (TTFT = submit -> first token, then the gaps between tokens).
V35 int pq35_generate(void* mp, const uint32_t* prompt, int n_prompt, int max_tokens, uint32_t* out_tokens, double* ttft_ms, double* step_ms) {
  pq35::Model* m = M35;
  using clk = std::chrono::steady_clock;
  const int id = pq35_request_alloc(mp);
  if (id < 0) return -1;
  const auto t0 = clk::now();
  void* lg = nullptr;
  bool ok = m->prefill(id, prompt, n_prompt, &lg);
  int tok = 0;
  if (ok) {
    m->k.flashinfer_top1_cuda(static_cast<const pk_bf16*>(lg), static_cast<pk_bf16*>(m->top1_val.ptr), static_cast<uint8_t*>(m->top1_states.ptr),
                              m->sample_out.i32(), m->c.vocab_size, m->st);
    ok = cudaMemcpyAsync(m->sample_h, m->sample_out.ptr, 4, cudaMemcpyDeviceToHost, m->st) == cudaSuccess &&
         cudaStreamSynchronize(m->st) == cudaSuccess;
    tok = m->sample_h[0];
  }
  auto tprev = clk::now();
  if (ttft_ms) *ttft_ms = std::chrono::duration<double, std::milli>(tprev - t0).count();
  if (ok && max_tokens > 0) out_tokens[0] = (uint32_t)tok;
  for (int i = 1; ok && i < max_tokens; ++i) {
    ok = m->decode(id, (uint32_t)tok, nullptr, &tok);
    const auto tn = clk::now();
    if (step_ms) step_ms[i - 1] = std::chrono::duration<double, std::milli>(tn - tprev).count();
    tprev = tn;
    if (ok) out_tokens[i] = (uint32_t)tok;
  }
  pq35_request_free(mp, id);
  return ok ? 0 : -1;
}

}  // extern "C"
