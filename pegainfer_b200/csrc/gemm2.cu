// Prefill GEMM on CTA PAIRS: Y[M, N] = W[M, K] x X[K, N] with tcgen05.mma.cta_group::2 (round 2).
//
// Why: the round-1 kernel (gemm.cu, cta_group::1, M128 x N256 x K16 per instruction) reads 12 KB of operands from
// shared memory per 128-cycle instruction and TMA writes another 12 KB: ~190 B/clk against a 128 B/clk shared-memory
// port.  ncu (profiles/r1_v3_gemm_full_details.txt) shows the tensor pipe 61-64 % active and cuBLAS' sm_100 kernels
// 1.27-1.53 PFLOP/s against our 0.77-1.20 on the same shapes.  A CTA pair (two SMs of one TPC) computes a
// 256-token x BN-feature tile: each CTA stages its OWN 128 tokens of A and only HALF of B (BN/2 features), the
// pair's tensor cores share the B halves, so per SM the operand traffic is 8 KB read + 8 KB written per instruction.
//
//   cluster (2,1,1): rank 0 = leader.   A = X [tokens, K] (UMMA M = 256 tokens, 128 per CTA), B = W [features, K]
//   (UMMA N = BN features, BN/2 per CTA), D in TMEM: each CTA holds its 128 token rows x BN columns (fp32), double buffered.
//   warp 0  TMA producer (both CTAs): A tile + B half per stage, complete_tx on the LEADER's full barrier
//           (cp.async.bulk.tensor ... .cta_group::2); the leader arms it with the bytes of both CTAs
//   warp 1  MMA issuer (leader only): tcgen05.mma.cta_group::2 M256 x BN x K16, tcgen05.commit ... multicast::cluster
//           frees the stage in BOTH CTAs / publishes the accumulator to BOTH epilogues
//   warps 2-5 epilogue (both CTAs): tcgen05.ld 32x32b.x32 -> bf16 -> 64-byte row stores; tmem-empty arrives on the
//           leader's barrier (remote mbarrier arrive for the peer)
// SwiGLU mode (gate_up projection): the pair's B halves are the SAME 128 features of the gate block (leader CTA) and
// of the up block (peer CTA) of the stacked [gate; up] weight, so columns [0,128) / [128,256) of every accumulator row
// are gate / up of one feature: the epilogue writes bf16(silu(bf16(g)) * bf16(u)) (csrc/fused_proj.cu:44-63 rounding)
// and the separate SiLU-mul pass over 2*I*T elements disappears.
// Every barrier wait is bounded (trap instead of hang).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tcgen05.cuh"

namespace pk {

namespace {

constexpr int G2_BM = 128;       // tokens per CTA (UMMA M = 256 per pair)
constexpr int G2_BK = 64;        // K elements per stage (one 128-byte swizzle span)
constexpr int G2_THREADS = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

__device__ __forceinline__ uint32_t g2_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t g2_map(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void g2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void g2_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 26)) __trap();
}
// TMA tile load whose completion bytes land on a barrier given as a shared::cluster address (the leader's)
__device__ __forceinline__ void g2_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
      : "memory");
}
__device__ __forceinline__ void g2_tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void g2_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void g2_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void g2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void g2_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

struct G2Args {
  bf16 *Y, *Y1, *Y2;
  int e0, e1;     // output segments (the fused q|k|v projection): features [0,e0) -> Y, [e0,e1) -> Y1, rest -> Y2
  int M, N, K;    // features, tokens, reduction
  int swiglu;     // 1: W = [gate (M rows); up (M rows)], Y[tok][M] = silu(gate) * up
};

// A pair tile is 256 tokens x (NSUB * BN) features: NSUB accumulators of BN columns side by side in TMEM, one
// tcgen05.mma per accumulator per K step (NSUB = 2, BN = 160: the 320-wide tile that turns the 2560-feature
// projections of Qwen3-4B into ONE wave of 64 tiles on 74 pairs instead of two waves of 256-wide tiles).
template <int BN, int NSUB, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const G2Args a) {
  constexpr int HB = BN / 2;          // B rows (features) staged per CTA per accumulator
  constexpr int TN = NSUB * BN;       // features per tile
  constexpr int A_BYTES = G2_BM * G2_BK * 2, B_SUB = HB * G2_BK * 2, B_BYTES = NSUB * B_SUB, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int NBUF = (2 * TN <= 512) ? 2 : 1;  // TMEM accumulator buffers
  constexpr uint32_t TMEM_COLS = (NBUF * TN <= 256) ? 256 : 512;
  static_assert(B_SUB % 1024 == 0, "each B sub-tile must start on a swizzle-atom boundary");
  static_assert(NSUB == 1 || TN <= 512, "accumulators exceed TMEM");
  extern __shared__ uint8_t g2_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(g2_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);  // used in the leader
  uint64_t* empty = full + STAGES;   // one per CTA (multicast commit)
  uint64_t* tfull = empty + STAGES;  // [2] one per CTA
  uint64_t* tempty = tfull + 2;      // [2] used in the leader (8 arrivals: 4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = g2_rank();
  const bool leader = rank == 0;
  const int M = a.M, N = a.N, K = a.K;
  const int m_tiles = (N + 2 * G2_BM - 1) / (2 * G2_BM);  // pair tiles along tokens
  const int n_tiles = a.swiglu ? (M + HB - 1) / HB : (M + TN - 1) / TN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + G2_BK - 1) / G2_BK;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull + s, 1);
      mbar_init(tempty + s, 8);
    }
    mbar_fence_init();
  }
  if (warp == 1) g2_tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  g2_cluster_sync();  // both CTAs' barriers are initialised and both TMEM allocations are done
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int mb = tile % m_tiles, nb = tile / m_tiles;
        const int tok0 = mb * 2 * G2_BM + (int)rank * G2_BM;
        // this CTA's B rows: its half of the BN features, or (SwiGLU) the gate block (leader) / up block (peer)
        const int wrow0 = a.swiglu ? (int)rank * M + nb * HB : nb * TN + (int)rank * HB;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          g2_wait(empty + s, (uint32_t)(((it / STAGES) & 1) ^ 1));
          const uint32_t fb = g2_map(smem_u32(full + s), 0);
          if (leader) mbar_expect_tx(full + s, 2 * STAGE_BYTES);
          uint8_t* st = smem + (size_t)s * STAGE_BYTES;
          g2_tma_2d(st, &map_x, kb * G2_BK, tok0, fb);
#pragma unroll
          for (int j = 0; j < NSUB; ++j) g2_tma_2d(st + A_BYTES + j * B_SUB, &map_w, kb * G2_BK, wrow0 + j * BN, fb);
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major both, N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int it = 0, lt = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
        const int as = lt % NBUF;
        g2_wait(tempty + as, (uint32_t)(((lt / NBUF) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * TN);
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          g2_wait(full + s, (uint32_t)((it / STAGES) & 1));
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_addr = smem_u32(smem + (size_t)s * STAGE_BYTES);
            const uint64_t adesc = make_sw128_desc(a_addr);
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k)
#pragma unroll
              for (int j = 0; j < NSUB; ++j)
                g2_umma(d_tmem + (uint32_t)(j * BN), adesc + (uint64_t)(k * 2),
                        make_sw128_desc(a_addr + A_BYTES + j * B_SUB) + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            g2_commit_both(empty + s);
            if (kb == k_blocks - 1) g2_commit_both(tfull + as);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const bool vec_ok = (a.e0 % 32 == 0) && (a.e1 % 32 == 0) && (M % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.Y) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(a.Y1) & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.Y2) & 15) == 0);
    const uint32_t tempty_leader0 = g2_map(smem_u32(tempty), 0);
    int lt = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
      const int mb = tile % m_tiles, nb = tile / m_tiles;
      const int as = lt % NBUF;
      g2_wait(tfull + as, (uint32_t)((lt / NBUF) & 1));
      tc_fence_after();
      const int tok = mb * 2 * G2_BM + (int)rank * G2_BM + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TN);
      if (a.swiglu) {
        // columns [0, HB) = gate, [HB, BN) = up of features nb*HB + c
#pragma unroll 1
        for (int c = 0; c < HB; c += 32) {
          uint32_t gv[32], uv[32];
          tmem_ld32_nowait(t_row + (uint32_t)c, gv);
          tmem_ld32(t_row + (uint32_t)(HB + c), uv);
          const int f0 = nb * HB + c;
          if (tok < N && f0 < M) {
            bf16* dst = a.Y + (size_t)tok * M + f0;
            float r[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float g = round_bf16(__uint_as_float(gv[j]));  // gate_up_out is bf16 in the reference
              const float u = round_bf16(__uint_as_float(uv[j]));
              r[j] = g / (1.0f + expf(-g)) * u;
            }
            if (M % 8 == 0 && f0 + 32 <= M && (reinterpret_cast<uintptr_t>(a.Y) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 o;
                o.x = pack_bf16(r[j * 8 + 0], r[j * 8 + 1]);
                o.y = pack_bf16(r[j * 8 + 2], r[j * 8 + 3]);
                o.z = pack_bf16(r[j * 8 + 4], r[j * 8 + 5]);
                o.w = pack_bf16(r[j * 8 + 6], r[j * 8 + 7]);
                reinterpret_cast<uint4*>(dst)[j] = o;
              }
            } else {
              for (int j = 0; j < 32 && f0 + j < M; ++j) dst[j] = f2bf(r[j]);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < TN; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)c, v);
          const int f0 = nb * TN + c;
          if (tok < N && f0 < M) {
            bf16* dst;
            if (f0 < a.e0) dst = a.Y + (size_t)tok * a.e0 + f0;
            else if (f0 < a.e1) dst = a.Y1 + (size_t)tok * (a.e1 - a.e0) + (f0 - a.e0);
            else dst = a.Y2 + (size_t)tok * (M - a.e1) + (f0 - a.e1);
            if (vec_ok && f0 + 32 <= M) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 o;
                o.x = pack_bf16(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
                o.y = pack_bf16(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
                o.z = pack_bf16(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
                o.w = pack_bf16(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
                reinterpret_cast<uint4*>(dst)[j] = o;
              }
            } else {
              for (int j = 0; j < 32 && f0 + j < M; ++j) {
                const int f = f0 + j;
                bf16* d1 = f < a.e0 ? a.Y + (size_t)tok * a.e0 + f
                                    : (f < a.e1 ? a.Y1 + (size_t)tok * (a.e1 - a.e0) + (f - a.e0) : a.Y2 + (size_t)tok * (M - a.e1) + (f - a.e1));
                *d1 = f2bf(__uint_as_float(v[j]));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) g2_arrive_cluster(tempty_leader0 + (uint32_t)(as * 8));
    }
  }
  tc_fence_before();
  __syncthreads();
  g2_cluster_sync();  // the peer's MMAs / loads may still target this CTA's shared memory and TMEM
  if (warp == 1) {
    tc_fence_after();
    g2_tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

bool g2_make_map(CUtensorMap* map, const void* base, int rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)G2_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int NSUB, int STAGES>
cudaError_t g2_launch(const CUtensorMap& mx, const CUtensorMap& mw, const G2Args& a, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (G2_BM * G2_BK * 2 + NSUB * (BN / 2) * G2_BK * 2) + 1024 + 256;
  auto kern = gemm_tc2_kernel<BN, NSUB, STAGES>;
  static thread_local bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return cudaErrorInvalidValue;
    configured = true;
  }
  const int m_tiles = (a.N + 2 * G2_BM - 1) / (2 * G2_BM);
  const int n_tiles = a.swiglu ? (a.M + BN / 2 - 1) / (BN / 2) : (a.M + NSUB * BN - 1) / (NSUB * BN);
  const int tiles = m_tiles * n_tiles;
  int pairs = sm_count() / 2;
  if (pairs > tiles) pairs = tiles;
  if (pairs < 1) pairs = 1;
  return launch(kern, dim3(2 * pairs), dim3(G2_THREADS), smem, stream, true, mx, mw, a);
}

}  // namespace

// Tile choice by wave count (cost ~ waves x (tile width + epilogue)), in the units of gemm.cu's model: a pair tile of 256
// tokens on sms/2 pairs costs what a 128-token tile of the same width costs on sms single CTAs.  Returns the candidate
// index (0: 256-wide, 1: 2 x 160-wide, 2: 128-wide) or -2 when the single-CTA kernel of gemm.cu should take the shape.
// Pure host arithmetic (exported as pk_b200_gemm_plan so the calibration is pinned by a CPU test).
static int g2_plan(int M, int N, int K, int swiglu, int sms) {
  if (K % 8 != 0 || M <= 0 || sms < 2) return -2;
  if (N <= G2_BM) return -2;  // one token tile: nothing for the second CTA of a pair to do
  const int pairs = sms / 2;
  const long m_tiles = (N + 2 * G2_BM - 1) / (2 * G2_BM);
  const int bn[3] = {256, 160, 128}, nsub[3] = {1, 2, 1};
  int best = -1;
  long best_cost = -1;
  for (int ci = 0; ci < (swiglu ? 1 : 3); ++ci) {
    const int tn = bn[ci] * nsub[ci];
    const long nt = swiglu ? (M + bn[ci] / 2 - 1) / (bn[ci] / 2) : (M + tn - 1) / tn;
    const long waves = (m_tiles * nt + pairs - 1) / pairs;
    // a 320-wide tile has ONE accumulator buffer in TMEM: its epilogue is not hidden behind the next tile's MMAs
    // (measured: gate_up 209 us with 320-wide vs 165 us with 256-wide tiles), so it only pays as a single wave
    const long cost = waves * (nsub[ci] == 2 ? tn + tn / 3 + 32 : tn + 32);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = ci;
    }
  }
  if (!swiglu) {  // is the single-CTA kernel (gemm.cu: 128-token tiles of width 256 / 160 / 128 on every SM) strictly better?
    // One cost unit of the single-CTA kernel takes 1.75-2.3 x as long as one of the pair kernel (it is bound by its own
    // shared-memory port, see the header): at 2048 tokens o_proj 80.8 us (2 waves of 160-wide tiles, 384 units) vs 46.5 us
    // for the pair's one wave of 320-wide tiles (458 units), down_proj 173 vs 88 us (profiles/README.md round 2).  Without
    // the factor the model sent both back to the single-CTA kernel (r2_v3_prefill_decode_launches.csv: 2 x 89 us per layer).
    const long m1 = (N + G2_BM - 1) / G2_BM;
    const int w1[3] = {256, 160, 128};
    for (int i = 0; i < 3; ++i) {
      const long waves = (m1 * ((M + w1[i] - 1) / w1[i]) + sms - 1) / sms;
      if (waves * (w1[i] + 32) * 7 / 4 < best_cost) return -2;
    }
  }
  return best;
}

// Returns 0 when launched, -2 when the shape / alignment is not for this kernel (caller uses the 1-CTA kernel).
int launch_gemm_pair(const bf16* W, const bf16* X, bf16* Y, bf16* Y1, bf16* Y2, int e0, int e1, int M, int N, int K, int swiglu,
                     cudaStream_t stream) {
  if (K % 8 != 0 || (reinterpret_cast<uintptr_t>(W) & 15) != 0 || (reinterpret_cast<uintptr_t>(X) & 15) != 0) return -2;
  const int best = g2_plan(M, N, K, swiglu, sm_count());
  if (best < 0) return -2;
  G2Args a{Y, Y1, Y2, e0, e1, M, N, K, swiglu};
  CUtensorMap mx, mw;
  if (!g2_make_map(&mx, X, N, K, G2_BM)) return -2;
  const int bn[3] = {256, 160, 128};
  const int rows_w = swiglu ? 2 * M : M;
  if (!g2_make_map(&mw, W, rows_w, K, bn[best] / 2)) return -2;
  cudaError_t e;
  if (best == 0) e = g2_launch<256, 1, 6>(mx, mw, a, stream);
  else if (best == 1) e = g2_launch<160, 2, 5>(mx, mw, a, stream);
  else e = g2_launch<128, 1, 8>(mx, mw, a, stream);
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace pk

// Which prefill GEMM kernel a [M features] x [N tokens] x [K] problem gets on a GPU with `sms` SMs: 256 / 320 / 128 = tile
// width of the CTA-pair kernel, -2 = the single-CTA kernel (gemm.cu; split-K when skinny).  No CUDA call.
extern "C" int pk_b200_gemm_plan(int M, int N, int K, int swiglu, int sms) {
  const int widths[3] = {256, 320, 128};
  const int best = pk::g2_plan(M, N, K, swiglu, sms);
  return best < 0 ? -2 : widths[best];
}
