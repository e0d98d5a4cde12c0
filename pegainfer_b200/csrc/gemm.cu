// Prefill GEMM: Y[M, N] = W[M, K] x X[K, N] on the 5th-generation tensor cores.
// Replaces `gemm_cuda` (csrc/linear.cu:48-62, a cublasGemmEx call in the reference).
// bf16 x bf16 -> fp32 accumulate (TMEM) -> one bf16 rounding, as CUBLAS_COMPUTE_32F.
//
// Mapping (both operands are K-major in memory, so no transposes anywhere):
//   A = X  : [tokens, K]   -> UMMA M dimension = 128 tokens  (TMEM lane  = token)
//   B = W  : [features, K] -> UMMA N dimension = BN features (TMEM column = feature)
//   D[token, feature] lands so that one thread owns one token row: the epilogue converts 32
//   consecutive features to bf16 and stores 64 contiguous bytes of HiddenStates[features, tokens].
// Pipeline (warp-specialised, persistent, one CTA per SM):
//   warp 0  : TMA producer   -- cp.async.bulk.tensor 2-D tiles, 128-B swizzle, mbarrier tx
//   warp 1  : MMA issuer     -- one elected lane issues tcgen05.mma (M128 x BN x K16), commits
//                               free the smem stage / publish the TMEM accumulator
//   warps 2-5: epilogue      -- tcgen05.ld 32x32b.x32 -> bf16 -> global; double-buffered TMEM
//                               accumulators so the epilogue of tile i overlaps the MMAs of i+1
// Tensor-bound: 2*M*N*K flops per call; roofline against MEASURED_PEAKS.json bf16_tflops.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tcgen05.cuh"

namespace pk {

// ------------------------------------------------------------------ SIMT fallback
// Shapes TMA cannot take (K % 8 != 0, unaligned bases) and the PK_GEMM_IMPL=simt debug switch.
__global__ void gemm_simt_kernel(const bf16* __restrict__ W, const bf16* __restrict__ X,
                                 bf16* __restrict__ Y, int M, int N, int K) {
  constexpr int T = 64, KT = 32;
  __shared__ float ws[KT][T + 1], xs[KT][T + 1];
  pdl_wait();
  const int m0 = blockIdx.x * T, n0 = blockIdx.y * T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 outputs each
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += KT) {
    for (int i = threadIdx.x; i < T * KT; i += 256) {
      const int r = i / KT, c = i % KT;
      ws[c][r] = (m0 + r < M && k0 + c < K) ? bf2f(W[(size_t)(m0 + r) * K + k0 + c]) : 0.f;
      xs[c][r] = (n0 + r < N && k0 + c < K) ? bf2f(X[(size_t)(n0 + r) * K + k0 + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < KT; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = ws[k][tx * 4 + i];
        b[i] = xs[k][ty * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(a[i], b[j], acc[j][i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + tx * 4 + i, n = n0 + ty * 4 + j;
      if (m < M && n < N) Y[(size_t)n * M + m] = f2bf(acc[j][i]);
    }
}

// ------------------------------------------------------------------ tcgen05 path
constexpr int BM = 128;  // tokens per tile (UMMA M)
constexpr int BK = 64;   // K elements per stage = one 128-byte swizzle span
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 192;

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
               bf16* __restrict__ Y, bf16* __restrict__ Y1, bf16* __restrict__ Y2, int e0, int e1,
               int M /*features*/, int N /*tokens*/, int K, int splits, float* __restrict__ part,
               unsigned int* __restrict__ cnt) {
  // splits > 1 (skinny problems: one 128-token tile, few feature tiles, deep K -- prefill of a short prompt, decode
  // buckets 5..64): the grid is tiles x splits CTAs, all co-resident; CTA (tile, sp) accumulates K blocks
  // [sp * kb / splits, (sp + 1) * kb / splits), parks its fp32 partial in `part` ([work][feature][token]: coalesced
  // both ways), and after the tile's `splits` partials have landed (ticket in cnt[2 * tile]) reduces ITS slice of the
  // tile's features in split order -- a fixed summation order, so the result is deterministic -- and stores bf16.
  // Output features [0,e0) go to Y (row length e0), [e0,e1) to Y1, [e1,M) to Y2: the fused q|k|v
  // projection writes three HiddenStates buffers from one launch (e0 == e1 == M: plain GEMM).
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;  // power of two >= 2 accumulator stages
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (N + BM - 1) / BM, n_tiles = (M + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles * splits;  // work items (tile, split)
  const int k_blocks_all = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull + s, 1);
      mbar_init(tempty + s, 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      int it = 0;
      for (int work = blockIdx.x; work < num_tiles; work += gridDim.x) {
        const int tile = work / splits, sp = work - tile * splits;
        const int mb = tile % m_tiles, nb = tile / m_tiles;
        const int kb0 = (int)(((int64_t)sp * k_blocks_all) / splits), kb1 = (int)(((int64_t)(sp + 1) * k_blocks_all) / splits);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty + s, (uint32_t)(((it / STAGES) & 1) ^ 1));
          mbar_expect_tx(full + s, STAGE_BYTES);
          uint8_t* st = smem + (size_t)s * STAGE_BYTES;
          tma_load_2d(st, &map_x, kb * BK, mb * BM, full + s);
          tma_load_2d(st + A_BYTES, &map_w, kb * BK, nb * BN, full + s);
        }
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major both,
    // N>>3 at [17,23), M>>4 at [24,29)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(BM >> 4) << 24);
    int it = 0, lt = 0;
    for (int work = blockIdx.x; work < num_tiles; work += gridDim.x, ++lt) {
      const int sp = work % splits;
      const int kb0 = (int)(((int64_t)sp * k_blocks_all) / splits), kb1 = (int)(((int64_t)(sp + 1) * k_blocks_all) / splits);
      const int k_blocks = kb1 - kb0;
      const int as = lt & 1;
      mbar_wait(tempty + as, (uint32_t)(((lt >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(full + s, (uint32_t)((it / STAGES) & 1));
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint64_t adesc = make_sw128_desc(a_addr);
          const uint64_t bdesc = make_sw128_desc(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance the start address by k*32 bytes inside the 128-byte swizzle span
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty + s);                       // smem stage free once these MMAs retire
          if (kb == k_blocks - 1) umma_commit(tfull + as);  // accumulator ready for the epilogue
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const bool vec_ok = (e0 % 32 == 0) && (e1 % 32 == 0) && (M % 8 == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(Y1) & 15) == 0) && ((reinterpret_cast<uintptr_t>(Y2) & 15) == 0);
    int lt = 0;
    for (int work = blockIdx.x; work < num_tiles; work += gridDim.x, ++lt) {
      const int tile = work / splits, sp = work - tile * splits;
      const int mb = tile % m_tiles, nb = tile / m_tiles;
      const int as = lt & 1;
      mbar_wait(tfull + as, (uint32_t)((lt >> 1) & 1));
      tc_fence_after();
      const int tok = mb * BM + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      if (splits > 1) {
        // ---- park the fp32 partial: part[work][feature][token] (a warp's 32 tokens are 128 contiguous bytes) ----
        float* mine = part + (size_t)work * BN * BM;
        const int trow = q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)c, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) mine[(size_t)(c + j) * BM + trow] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty + as);
        __threadfence();
        asm volatile("bar.sync 2, 128;" ::: "memory");  // the four epilogue warps
        if (warp == 2 && lane == 0) {
          atomicAdd(cnt + 2 * tile, 1u);
          uint32_t spins = 0;
          while (*reinterpret_cast<volatile unsigned int*>(cnt + 2 * tile) < (unsigned)splits)
            if (++spins > (1u << 26)) __trap();
          __threadfence();
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        // ---- reduce this CTA's slice of the tile's features (8-feature groups dealt round-robin to the splits) ----
        const float* base = part + (size_t)tile * splits * BN * BM;
        for (int g8 = sp; g8 < BN / 8; g8 += splits) {
          const int f0 = nb * BN + g8 * 8;
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
          for (int s2 = 0; s2 < splits; ++s2) {
            const float* src = base + ((size_t)s2 * BN + g8 * 8) * BM + trow;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += __ldcg(src + (size_t)j * BM);
          }
          if (tok < N && f0 < M) {
            for (int j = 0; j < 8 && f0 + j < M; ++j) {
              const int f = f0 + j;
              bf16* d1 = f < e0 ? Y + (size_t)tok * e0 + f
                                : (f < e1 ? Y1 + (size_t)tok * (e1 - e0) + (f - e0) : Y2 + (size_t)tok * (M - e1) + (f - e1));
              *d1 = f2bf(acc[j]);
            }
          }
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
          __threadfence();
          if (atomicAdd(cnt + 2 * tile + 1, 1u) == (unsigned)splits - 1) {  // last reader: reset for the next launch / replay
            cnt[2 * tile] = 0;
            cnt[2 * tile + 1] = 0;
            __threadfence();
          }
        }
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)c, v);
        const int f0 = nb * BN + c;
        if (tok < N && f0 < M) {
          bf16* dst;
          if (f0 < e0) dst = Y + (size_t)tok * e0 + f0;
          else if (f0 < e1) dst = Y1 + (size_t)tok * (e1 - e0) + (f0 - e0);
          else dst = Y2 + (size_t)tok * (M - e1) + (f0 - e1);
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
              bf16* d1 = f < e0 ? Y + (size_t)tok * e0 + f
                                : (f < e1 ? Y1 + (size_t)tok * (e1 - e0) + (f - e0) : Y2 + (size_t)tok * (M - e1) + (f - e1));
              *d1 = f2bf(__uint_as_float(v[j]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side ----
// rows x K bf16 row-major matrix, box = [box_rows, 64], 128-byte swizzle, OOB reads as zero
static bool make_map(CUtensorMap* map, const void* base, int rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool gemm_splitk_enabled() {  // PK_GEMM_SPLITK=0: A/B switch
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PK_GEMM_SPLITK");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

template <int BN, int STAGES>
static cudaError_t launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, bf16* Y, bf16* Y1, bf16* Y2, int e0,
                             int e1, int M, int N, int K, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 + 256;
  auto kern = gemm_tc_kernel<BN, STAGES>;
  static thread_local bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  const int tiles = ((N + BM - 1) / BM) * ((M + BN - 1) / BN);
  int grid = tiles < sm_count() ? tiles : sm_count();
  // Skinny problem (few tiles, deep K): split K across otherwise idle SMs.  Every (tile, split) CTA must be resident at
  // once (they wait for each other's partials): grid = tiles * splits <= #SMs, one CTA per SM by shared-memory size.
  ThreadState& ts = tls();
  const int k_blocks = (K + BK - 1) / BK;
  int splits = 1;
  if (gemm_splitk_enabled() && ts.gemm_part && tiles * 2 <= sm_count() && tiles <= 1024) {
    splits = sm_count() / tiles;
    if (splits > k_blocks / 4) splits = k_blocks / 4;  // >= 4 K blocks (256 elements) per split
    if (splits > BN / 8) splits = BN / 8;
    while (splits > 1 && (size_t)tiles * splits * BM * BN * 4 > ts.gemm_part_bytes) --splits;
    if (splits < 2) splits = 1;
  }
  if (splits > 1) grid = tiles * splits;
  return launch(kern, dim3(grid), dim3(kGemmThreads), smem, stream, true, mx, mw, Y, Y1, Y2, e0, e1, M, N, K, splits,
                ts.gemm_part, ts.gemm_cnt);
}

static int gemm_impl_mode() {  // 0 tcgen05 (default), 1 simt
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("PK_GEMM_IMPL");
    mode = (e && strcmp(e, "simt") == 0) ? 1 : 0;
  }
  return mode;
}

void launch_gemm_seg(const bf16* W, const bf16* X, bf16* Y, bf16* Y1, bf16* Y2, int e0, int e1, int M, int N, int K,
                     cudaStream_t stream);
// gemm2.cu: the CTA-pair kernel (tcgen05.mma.cta_group::2); -2 = not for this shape
int launch_gemm_pair(const bf16* W, const bf16* X, bf16* Y, bf16* Y1, bf16* Y2, int e0, int e1, int M, int N, int K, int swiglu,
                     cudaStream_t stream);
static bool gemm_pair_enabled() {  // PK_GEMM_PAIR=0 keeps the round-1 single-CTA kernel for A/B runs
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PK_GEMM_PAIR");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

void launch_gemm(const bf16* W, const bf16* X, bf16* Y, int M, int N, int K, cudaStream_t stream) {
  launch_gemm_seg(W, X, Y, Y, Y, M, M, M, N, K, stream);
}

void launch_gemm_seg(const bf16* W, const bf16* X, bf16* Y, bf16* Y1, bf16* Y2, int e0, int e1, int M, int N, int K,
                     cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return;
  const bool tma_ok = K % 8 == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(X) & 15) == 0;
  if (tma_ok && gemm_impl_mode() == 0 && gemm_pair_enabled() &&
      launch_gemm_pair(W, X, Y, Y1, Y2, e0, e1, M, N, K, 0, stream) == 0)
    return;
  if (tma_ok && gemm_impl_mode() == 0) {
    // Tile width by wave count: tiles = ceil(N/128) * ceil(M/BN) on `sms` persistent CTAs; cost ~ waves * (BN + c).
    // M = 2560 outputs (o_proj / down_proj of Qwen3-4B) at 2048 tokens: BN 256 -> 160 tiles = 2 waves of 256-wide
    // tiles, BN 160 -> 256 tiles = 2 waves of 160-wide tiles (profiles/README.md, r1 v3: tile quantisation).
    const int m_tiles = (N + BM - 1) / BM;
    const int sms = sm_count();
    int best_bn = 256;
    long best_cost = -1;
    const int cands[3] = {256, 160, 128};
    for (int ci = 0; ci < 3; ++ci) {
      const int bnc = cands[ci];
      const long tiles = (long)m_tiles * ((M + bnc - 1) / bnc);
      const long waves = (tiles + sms - 1) / sms;
      const long cost = waves * (bnc + 32);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_bn = bnc;
      }
    }
    CUtensorMap mx, mw;
    if (make_map(&mx, X, N, K, BM) && make_map(&mw, W, M, K, best_bn)) {
      if (best_bn == 128)
        launch_tc<128, 6>(mx, mw, Y, Y1, Y2, e0, e1, M, N, K, stream);
      else if (best_bn == 160)
        launch_tc<160, 5>(mx, mw, Y, Y1, Y2, e0, e1, M, N, K, stream);
      else
        launch_tc<256, 4>(mx, mw, Y, Y1, Y2, e0, e1, M, N, K, stream);
      return;
    }
  }
  // SIMT fallback: one launch per segment (row slices of W are contiguous)
  const int starts[3] = {0, e0, e1}, ends[3] = {e0, e1, M};
  bf16* outs[3] = {Y, Y1, Y2};
  for (int sgi = 0; sgi < 3; ++sgi) {
    const int m = ends[sgi] - starts[sgi];
    if (m <= 0) continue;
    launch(gemm_simt_kernel, dim3((m + 63) / 64, (N + 63) / 64), dim3(256), 0, stream, true,
           W + (size_t)starts[sgi] * K, X, outs[sgi], m, N, K);
  }
}

// defined in gemv.cu
struct GemvArgs;
bool gemv_stream_supported(const void* W, const void* X, int N, int K);
void launch_gemv_generic(const bf16* W, const bf16* X, bf16* Y, int M, int N, int K,
                         cudaStream_t stream);

}  // namespace pk

// Fused multi-output projection (q|k|v from the stacked qkv_proj): rows [0,seg_rows[0]) -> Y[0], ...
extern "C" int pk_b200_gemm_segments(const pk_bf16* W, const pk_bf16* X, pk_bf16* const* Y, const int* seg_rows, int M,
                                     int N, int K, pk_stream stream) {
  if (!W || !X || !Y || !seg_rows || seg_rows[0] + seg_rows[1] + seg_rows[2] != M) return -1;
  pk::launch_gemm_seg((const pk::bf16*)W, (const pk::bf16*)X, (pk::bf16*)Y[0], (pk::bf16*)(Y[1] ? Y[1] : Y[0]),
                      (pk::bf16*)(Y[2] ? Y[2] : Y[0]), seg_rows[0], seg_rows[0] + seg_rows[1], M, N, K, stream);
  return 0;
}

// gate_up projection with the SwiGLU activation folded into the epilogue: W = [gate (M rows); up (M rows)] x [K],
// Y[tok][M] = bf16(silu(bf16(gate.x)) * bf16(up.x)) -- gemm + silu_mul_fused_cuda of the reference in one launch, same
// rounding points.  Returns -2 when the CTA-pair kernel cannot take the shape (caller runs the two-kernel sequence).
extern "C" int pk_b200_gemm_swiglu(const pk_bf16* W, const pk_bf16* X, pk_bf16* Y, int M, int N, int K, pk_stream stream) {
  if (!W || !X || !Y || M <= 0 || N <= 0 || K <= 0) return -1;
  if (!pk::gemm_pair_enabled() || pk::gemm_impl_mode() != 0) return -2;
  return pk::launch_gemm_pair((const pk::bf16*)W, (const pk::bf16*)X, (pk::bf16*)Y, (pk::bf16*)Y, (pk::bf16*)Y, M, M, M, N, K, 1, stream);
}

// Same dispatch as gemm_graphsafe_cuda (gemv.cu): N <= 4 streams the weights (HBM-bound GEMV), larger N runs on the
// tensor cores.  Neither path allocates or synchronises, so both ABI names are capture-safe here.
extern "C" void gemm_cuda(const pk_bf16* W, const pk_bf16* X, pk_bf16* Y, int M, int N, int K,
                          pk_stream stream) {
  gemm_graphsafe_cuda(W, X, Y, M, N, K, stream);
}
