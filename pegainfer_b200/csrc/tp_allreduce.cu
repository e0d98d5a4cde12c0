// Tensor-parallel all-reduce over NVLink peer memory, optionally fused with the residual add +
// RMSNorm that follows it in the layer.  This is the hook north_star asks pegainfer-comm to gain;
// it replaces `Qwen3Model::all_reduce_hidden` (pegainfer-qwen3-4b/src/weights.rs:396-405, an
// ncclAllReduce on the compute stream) and the fused_add_rms_norm launch after it
// (batch_decode.rs:266-277,292 / prefill.rs:154-163).
//
// Decode messages are 5-8 KiB (H x 1 bf16): pure latency, so the algorithm is ONE-SHOT and
// PUSH based: every rank stores its partial straight into every peer's staging slot (posted
// NVLink writes, no round trip), publishes a sequence flag with release semantics, polls only
// LOCAL flags, then reduces the `world` partials from local memory in rank order (fp32, one
// rounding -- every rank computes the identical sum, so TP ranks never diverge).  Two staging
// slots alternate by sequence parity: a peer can be at most one collective ahead.  The sequence
// counter lives in device memory and is advanced by the kernel, so the launch is CUDA-graph
// replayable.  One process per GPU: peers map each other's staging with cudaIpc handles.
//
// Second protocol (PK_TP_PROTO=ll, opt-in until measured on hardware): the flag travels WITH the data.  Every
// 16-byte line carries {2 bf16, seq, 2 bf16, seq}; a line is stored and loaded with one volatile 16-byte access,
// so the receiver simply polls each line until both sequence words match -- no system-scope fence, no separate
// flag store, no block barrier between push and reduce.  Same rank-ordered fp32 sum, bit-identical results.
// The lines live in the tail (`slot_bytes - raw_bytes`) of every staging region, never shared with plain rows.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace pk {

struct TpArgs {
  TpDev d;
  const bf16* partial;  // this rank's partial [T, dim]
  bf16* hidden;         // mode 0: output of the sum (may alias partial); mode 1: residual stream
  const bf16* weight;
  bf16* out;
  int dim, T, mode;
  float eps;
};

__device__ __forceinline__ void st_volatile_v4(void* p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

template <bool LL>
__global__ void __launch_bounds__(kTpThreads) tp_allreduce_kernel(const TpArgs a) {
  extern __shared__ float tp_row[];  // mode 1: dim floats + 40
  const int me = a.d.rank, W = a.d.world;
  const int G = gridDim.x, c = blockIdx.x;
  const int nv = a.dim >> 3;
  uint32_t* my_flags = a.d.flags[me];
  uint32_t* ctl = my_flags + 2 * kTpMaxCtas * kTpMaxWorld;  // [0]=seq, [1]=done counter
  pdl_wait();
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(ctl) + 1u;
  const int slot = (int)(seq & 1u);

  // ---- push my partial into every peer's staging region [slot][me] ----
  for (int t = c; t < a.T; t += G) {
    const uint4* src = reinterpret_cast<const uint4*>(a.partial + (size_t)t * a.dim);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      const uint4 v = src[i];
      for (int p = 0; p < W; ++p) {
        if (p == me) continue;
        uint8_t* region = a.d.stage[p] + (size_t)(slot * W + me) * a.d.slot_bytes;
        if (LL) {
          uint8_t* line = region + a.d.ll_off + ((size_t)t * nv + i) * 32;
          st_volatile_v4(line, v.x, seq, v.y, seq);
          st_volatile_v4(line + 16, v.z, seq, v.w, seq);
        } else {
          reinterpret_cast<uint4*>(region + (size_t)t * a.dim * 2)[i] = v;
        }
      }
    }
  }
  if (!LL) {
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < W && (int)threadIdx.x != me) {
      const int p = threadIdx.x;
      st_release_sys(a.d.flags[p] + (size_t)(slot * kTpMaxCtas + c) * kTpMaxWorld + me, seq);
      const uint32_t* f = my_flags + (size_t)(slot * kTpMaxCtas + c) * kTpMaxWorld + p;
      for (uint32_t spins = 0; ld_acquire_sys(f) != seq; ++spins)
        if (spins > (1u << 27)) __trap();  // a lost peer becomes a launch failure instead of a wedged GPU
    }
    __syncthreads();
  }

  // ---- reduce in rank order from local memory ----
  const uint8_t* local = a.d.stage[me] + (size_t)slot * W * a.d.slot_bytes;
  float* red = tp_row + a.dim;
  for (int t = c; t < a.T; t += G) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int r = 0; r < W; ++r) {
        uint4 v;
        if (r == me) {
          v = reinterpret_cast<const uint4*>(a.partial + (size_t)t * a.dim)[i];
        } else if (LL) {
          const uint8_t* line = local + (size_t)r * a.d.slot_bytes + a.d.ll_off + ((size_t)t * nv + i) * 32;
          uint4 l0, l1;
          uint32_t spins = 0;
          do { l0 = ld_volatile_v4(line); if (++spins > (1u << 27)) __trap(); } while (l0.y != seq || l0.w != seq);
          do { l1 = ld_volatile_v4(line + 16); if (++spins > (1u << 27)) __trap(); } while (l1.y != seq || l1.w != seq);
          v = make_uint4(l0.x, l0.z, l1.x, l1.z);
        } else {
          v = __ldcg(reinterpret_cast<const uint4*>(local + (size_t)r * a.d.slot_bytes +
                                                    (size_t)t * a.dim * 2) + i);
        }
        acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
        acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
      }
      if (a.mode == 0) {
        uint4 o;
        o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
        o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
        reinterpret_cast<uint4*>(a.hidden + (size_t)t * a.dim)[i] = o;
      } else {
        // the collective's result is bf16 (NCCL in place on a bf16 buffer), then
        // fused_add_rms_norm: x = f32(hidden) + f32(sum); hidden = bf16(x); norm on unrounded x
        const uint4 h = reinterpret_cast<const uint4*>(a.hidden + (size_t)t * a.dim)[i];
        float x[8] = {bf16_lo(h.x) + round_bf16(acc[0]), bf16_hi(h.x) + round_bf16(acc[1]),
                      bf16_lo(h.y) + round_bf16(acc[2]), bf16_hi(h.y) + round_bf16(acc[3]),
                      bf16_lo(h.z) + round_bf16(acc[4]), bf16_hi(h.z) + round_bf16(acc[5]),
                      bf16_lo(h.w) + round_bf16(acc[6]), bf16_hi(h.w) + round_bf16(acc[7])};
        uint4 hs;
        hs.x = pack_bf16(x[0], x[1]); hs.y = pack_bf16(x[2], x[3]);
        hs.z = pack_bf16(x[4], x[5]); hs.w = pack_bf16(x[6], x[7]);
        reinterpret_cast<uint4*>(a.hidden + (size_t)t * a.dim)[i] = hs;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ss = fmaf(x[j], x[j], ss);
          tp_row[i * 8 + j] = x[j];
        }
      }
    }
    if (a.mode == 1) {
      const float tot = block_sum(ss, red);
      const float rinv = rsqrtf(tot / (float)a.dim + a.eps);
      for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const uint4 g = reinterpret_cast<const uint4*>(a.weight)[i];
        const float* x = tp_row + i * 8;
        uint4 o;
        o.x = pack_bf16(x[0] * rinv * bf16_lo(g.x), x[1] * rinv * bf16_hi(g.x));
        o.y = pack_bf16(x[2] * rinv * bf16_lo(g.y), x[3] * rinv * bf16_hi(g.y));
        o.z = pack_bf16(x[4] * rinv * bf16_lo(g.z), x[5] * rinv * bf16_hi(g.z));
        o.w = pack_bf16(x[6] * rinv * bf16_lo(g.w), x[7] * rinv * bf16_hi(g.w));
        reinterpret_cast<uint4*>(a.out + (size_t)t * a.dim)[i] = o;
      }
      __syncthreads();
    }
  }
  // ---- advance the sequence once per launch (last CTA) ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(ctl + 1, 1u);
    if (done == (uint32_t)G - 1) {
      ctl[1] = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(ctl) = seq;
    }
  }
}

// Vocab-sharded lm_head under tensor parallelism: every rank holds the (max logit, GLOBAL index) of its vocabulary
// shard; one 16-byte {value, seq, index, seq} line per request is pushed to every peer (posted NVLink stores; EACH
// 8-byte half carries the sequence number, because a 16-byte store may cross the link as two 8-byte transactions and
// a reader that validated only one half would pair a fresh sequence with a stale value -- seen on hardware) and every rank picks the same winner: highest value, lowest index on ties (the
// single-GPU top-1 rule).  Replaces a replicated lm_head GEMV (1.24 GB per token per rank for Qwen3-8B).
struct TpTop1Args {
  TpDev d;
  const bf16* vals;   // [bs] local maxima
  int* idx;           // [bs] in: local index inside the shard, out: global winner
  int bs, vocab_offset, entry0;
  const uint32_t* step;  // device step counter (graph path) or null
  uint32_t seq_arg;      // op index (with `step`) or the absolute sequence number
};

__global__ void __launch_bounds__(64) tp_top1_exchange_kernel(const TpTop1Args a) {
  pdl_wait();
  const int t = threadIdx.x;
  if (t >= a.bs) return;
  const int me = a.d.rank, W = a.d.world;
  const uint32_t seq = a.step ? (*a.step) * 256u + a.seq_arg + 1u : a.seq_arg;
  const float v = bf2f(a.vals[t]);
  const int gi = a.idx[t] + a.vocab_offset;
  const int64_t off = a.d.gll_off + a.d.gll_bytes - 4096 + (int64_t)(a.entry0 + t) * 16;
  for (int p = 0; p < W; ++p)
    if (p != me) {
      uint8_t* line = a.d.stage[p] + (size_t)me * a.d.slot_bytes + off;
      asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line), "r"(__float_as_uint(v)), "r"(seq),
                   "r"((uint32_t)gi), "r"(seq)
                   : "memory");
    }
  float bv = v;
  int bi = gi;
  for (int q = 0; q < W; ++q) {
    if (q == me) continue;
    const uint8_t* line = a.d.stage[me] + (size_t)q * a.d.slot_bytes + off;
    uint4 l;
    uint32_t spins = 0;
    do {
      asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(l.x), "=r"(l.y), "=r"(l.z), "=r"(l.w) : "l"(line) : "memory");
      if (++spins > (1u << 26)) __trap();
    } while (l.y != seq || l.w != seq);
    const float ov = __uint_as_float(l.x);
    const int oi = (int)l.z;
    if (ov > bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  a.idx[t] = bi;
}

}  // namespace pk


using namespace pk;

static bool tp_use_ll() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PK_TP_PROTO");
    v = (e && strcmp(e, "ll") == 0) ? 1 : 0;
  }
  return v == 1;
}

extern "C" {

pk_tp_comm* pk_tp_comm_create(int rank, int world, void* const* staging_ptrs,
                              void* const* flag_ptrs, int64_t staging_bytes) {
  if (world < 1 || world > kTpMaxWorld || rank < 0 || rank >= world) return nullptr;
  pk_tp_comm* c = new pk_tp_comm();
  memset(c, 0, sizeof(*c));
  c->d.rank = rank;
  c->d.world = world;
  for (int p = 0; p < world; ++p) {
    c->d.stage[p] = static_cast<uint8_t*>(staging_ptrs[p]);
    c->d.flags[p] = static_cast<uint32_t*>(flag_ptrs[p]);
  }
  c->staging_bytes = staging_bytes;
  c->d.slot_bytes = (staging_bytes / (2 * world)) & ~(int64_t)15;
  // region = [plain rows: 3/4][LL lines of the standalone kernel: 1/8][LL lines of the GEMV-fused all-reduce: 1/8]
  const int64_t eighth = (c->d.slot_bytes / 8) & ~(int64_t)15;
  c->d.raw_bytes = c->d.slot_bytes - 2 * eighth;
  c->d.ll_off = c->d.raw_bytes;
  c->d.ll_bytes = eighth;
  c->d.gll_off = c->d.raw_bytes + eighth;
  c->d.gll_bytes = eighth;
  return c;
}

void pk_tp_comm_destroy(pk_tp_comm* comm) { delete comm; }

int64_t pk_tp_flag_bytes(void) { return (2 * kTpMaxCtas * kTpMaxWorld + 16) * sizeof(uint32_t); }

static int tp_launch(pk_tp_comm* comm, const pk_bf16* partial, pk_bf16* hidden,
                     const pk_bf16* weight, pk_bf16* out, int dim, int T, float eps, int mode,
                     pk_stream stream) {
  if (!comm || dim % 8 != 0 || T <= 0) return -1;
  if ((int64_t)T * dim * 2 > comm->d.raw_bytes) return -2;  // caller chunks larger messages
  TpArgs a{};
  a.d = comm->d;
  a.partial = (const bf16*)partial;
  a.hidden = (bf16*)hidden;
  a.weight = (const bf16*)weight;
  a.out = (bf16*)out;
  a.dim = dim; a.T = T; a.mode = mode; a.eps = eps;
  const int grid = T < kTpMaxCtas ? T : kTpMaxCtas;
  const size_t smem = mode == 1 ? sizeof(float) * ((size_t)dim + 40) : sizeof(float) * 40;
  // LL lines need 4 bytes per element; messages that do not fit the LL area use the flag protocol
  const bool ll = tp_use_ll() && (int64_t)T * dim * 4 <= comm->d.ll_bytes;
  if (ll) return (int)launch(tp_allreduce_kernel<true>, dim3(grid), dim3(kTpThreads), smem, stream, true, a);
  return (int)launch(tp_allreduce_kernel<false>, dim3(grid), dim3(kTpThreads), smem, stream, true, a);
}

int pk_tp_all_reduce(pk_tp_comm* comm, pk_bf16* hidden, int64_t n, pk_stream stream) {
  // in-place SUM over n bf16 elements, treated as rows of <= 4096 elements
  if (!comm || n <= 0 || n % 8 != 0) return -1;
  int dim = (int)n, T = 1;
  if (n > 8192) {
    for (int cand = 8192; cand >= 8; cand -= 8)
      if (n % cand == 0) { dim = cand; T = (int)(n / cand); break; }
  }
  return tp_launch(comm, hidden, hidden, nullptr, nullptr, dim, T, 0.f, 0, stream);
}

int pk_tp_all_reduce_rows(pk_tp_comm* comm, pk_bf16* hidden, int hidden_dim, int rows,
                          pk_stream stream) {
  return tp_launch(comm, hidden, hidden, nullptr, nullptr, hidden_dim, rows, 0.f, 0, stream);
}

int pk_tp_all_reduce_add_rms_norm(pk_tp_comm* comm, pk_bf16* hidden, const pk_bf16* partial,
                                  const pk_bf16* weight, pk_bf16* out, int hidden_dim,
                                  int batch_size, float eps, pk_stream stream) {
  if (hidden_dim > 11000) return -1;  // row parked in shared memory as fp32
  return tp_launch(comm, partial, hidden, weight, out, hidden_dim, batch_size, eps, 1, stream);
}

int pk_tp_top1_exchange(pk_tp_comm* comm, const pk_bf16* local_max, int* index_inout, int batch_size, int vocab_offset,
                        const uint32_t* step_counter, uint32_t seq_or_op, pk_stream stream) {
  if (!comm || batch_size <= 0 || batch_size > 64 || comm->d.gll_bytes < 8192) return -1;
  TpTop1Args a{};
  a.d = comm->d;
  a.vals = (const bf16*)local_max;
  a.idx = index_inout;
  a.bs = batch_size;
  a.vocab_offset = vocab_offset;
  a.entry0 = step_counter ? 0 : 64;  // graph-path and host-path calls never share a line
  a.step = step_counter;
  a.seq_arg = seq_or_op;
  return (int)launch(tp_top1_exchange_kernel, dim3(1), dim3(64), 0, stream, true, a);
}

int64_t pk_tp_max_rows(pk_tp_comm* comm, int hidden_dim) {
  if (!comm || hidden_dim <= 0) return 0;
  return comm->d.raw_bytes / ((int64_t)hidden_dim * 2);
}

// cudaIpc plumbing for the one-process-per-GPU model (handles travel over torch.distributed)
int pk_tp_ipc_export(void* dev_ptr, void* handle_out_64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, dev_ptr);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle_out_64, &h, sizeof(h));
  return 0;
}
int pk_tp_ipc_open(const void* handle_64, void** dev_ptr_out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
}
int pk_tp_ipc_close(void* dev_ptr) { return (int)cudaIpcCloseMemHandle(dev_ptr); }

}  // extern "C"
