// Per-thread runtime state: device binding, scratch, launch counter.
// Replaces csrc/linear.cu:7-42 of the reference (thread-local cuBLAS handles + workspace):
// same entry points, same "once per rank thread" contract, no cuBLAS.
#include "common.cuh"

namespace pk {
ThreadState& tls() {
  static thread_local ThreadState s;
  return s;
}
int sm_count() {
  ThreadState& ts = tls();
  if (ts.sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&ts.sm_count, cudaDevAttrMultiProcessorCount, dev);
    ts.device = dev;
    if (ts.sm_count <= 0) ts.sm_count = 148;
  }
  return ts.sm_count;
}
}  // namespace pk

extern "C" {

int cuda_set_device(int device_ordinal) {
  pk::ThreadState& ts = pk::tls();
  ts.sm_count = 0;
  return static_cast<int>(cudaSetDevice(device_ordinal));
}

// Scratch for the internally split non-partition decode attention and the two-stage top-1.
static const size_t kScratchBytes = 16u << 20;

void cublas_init() {
  pk::ThreadState& ts = pk::tls();
  if (ts.scratch == nullptr) {
    if (cudaMalloc(&ts.scratch, kScratchBytes) == cudaSuccess) {
      ts.scratch_bytes = kScratchBytes;
      cudaMemset(ts.scratch, 0, kScratchBytes);
    } else {
      ts.scratch = nullptr;
      ts.scratch_bytes = 0;
    }
  }
  if (ts.gemm_part == nullptr) {  // split-K partial tiles: up to one 128 x 256 fp32 tile per SM, + tickets
    const size_t bytes = (size_t)160 * 128 * 256 * 4;
    void* p = nullptr;
    void* c = nullptr;
    if (cudaMalloc(&p, bytes) == cudaSuccess && cudaMalloc(&c, 4096) == cudaSuccess) {
      cudaMemset(c, 0, 4096);
      ts.gemm_part = static_cast<float*>(p);
      ts.gemm_cnt = static_cast<unsigned int*>(c);
      ts.gemm_part_bytes = bytes;
    } else {
      if (p) cudaFree(p);
      cudaGetLastError();
    }
  }
  (void)pk::sm_count();
}

void cublas_destroy() {
  pk::ThreadState& ts = pk::tls();
  if (ts.scratch != nullptr) {
    cudaFree(ts.scratch);
    ts.scratch = nullptr;
    ts.scratch_bytes = 0;
  }
  if (ts.gemm_part != nullptr) {
    cudaFree(ts.gemm_part);
    cudaFree(ts.gemm_cnt);
    ts.gemm_part = nullptr;
    ts.gemm_cnt = nullptr;
    ts.gemm_part_bytes = 0;
  }
}

const char* pk_b200_version(void) {
  return "pegainfer-kernels-b200 0.1 (sm_100a; TMA-bulk GEMV, tcgen05 GEMM, no cuBLAS/FlashInfer)";
}

int64_t pk_b200_launch_count(int reset) {
  pk::ThreadState& ts = pk::tls();
  int64_t v = ts.launches;
  if (reset) ts.launches = 0;
  return v;
}

void pk_b200_set_pdl(int enable) { pk::tls().pdl = enable != 0; }

}  // extern "C"
