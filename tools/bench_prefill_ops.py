"""CUDA-event timing of the prefill building blocks at Qwen3-4B shapes (tensor-bound leg)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pegainfer_b200 import ffi  # noqa: E402

lib = ffi.lib()
torch.zeros(1, device="cuda")
lib.cuda_set_device(0)
lib.cublas_init()
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for T in (128, 2048):
    print(f"--- T = {T} tokens")
    for name, M, K in (("q", 4096, 2560), ("kv", 1024, 2560), ("o", 2560, 4096), ("gate_up", 19456, 2560), ("down", 2560, 9728)):
        W = (torch.randn((M, K), device="cuda") * 0.02).to(torch.bfloat16)
        X = torch.randn((T, K), device="cuda").to(torch.bfloat16)
        Y = torch.empty((T, M), device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: lib.gemm_cuda(W.data_ptr(), X.data_ptr(), Y.data_ptr(), M, T, K, st))
        ref = timeit(lambda: torch.matmul(X, W.t()))
        fl = 2.0 * M * T * K
        print(f"gemm {name:8s} M={M:6d} K={K:5d}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s   (torch/cuBLAS {ref*1e3:8.1f} us {fl/ref/1e9:7.1f})")
    # gate_up + SwiGLU: one launch (pair kernel epilogue) vs gemm + silu_mul_fused
    I, H = 9728, 2560
    W = (torch.randn((2 * I, H), device="cuda") * 0.02).to(torch.bfloat16)
    X = torch.randn((T, H), device="cuda").to(torch.bfloat16)
    gu = torch.empty((T, 2 * I), device="cuda", dtype=torch.bfloat16)
    act = torch.empty((T, I), device="cuda", dtype=torch.bfloat16)
    if lib.pk_b200_gemm_swiglu(W.data_ptr(), X.data_ptr(), act.data_ptr(), I, T, H, st) == 0:
        ms1 = timeit(lambda: lib.pk_b200_gemm_swiglu(W.data_ptr(), X.data_ptr(), act.data_ptr(), I, T, H, st))
    else:
        ms1 = float("nan")
    def two():
        lib.gemm_cuda(W.data_ptr(), X.data_ptr(), gu.data_ptr(), 2 * I, T, H, st)
        lib.silu_mul_fused_cuda(gu.data_ptr(), act.data_ptr(), I, T, st)
    ms2 = timeit(two)
    fl = 2.0 * 2 * I * T * H
    print(f"gate_up+SwiGLU fused: {ms1*1e3:8.1f} us {fl/ms1/1e9:7.1f} TFLOP/s   gemm + silu_mul: {ms2*1e3:8.1f} us")
    del W, X, gu, act
    # attention: one layer, nq 32 nkv 8
    nq, nkv, hd = 32, 8, 128
    pages = T // 16 + 1
    stride = 2 * 16 * nkv * hd
    kv = torch.randn(((pages + 2) * stride,), device="cuda").to(torch.bfloat16)
    q = torch.randn((T, nq * hd), device="cuda").to(torch.bfloat16)
    out = torch.empty_like(q)
    i32 = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
    pi, ip, lpl = i32(list(range(1, pages + 1))), i32([0, pages]), i32([((T - 1) % 16) + 1])
    qi, z, kc, tn = i32([0, T]), i32([0] * 4096), i32([T]), i32([T])
    fn = lambda: lib.batch_prefill_paged_cuda_with_cta_tile_q(q.data_ptr(), out.data_ptr(), kv.data_ptr(), 0, 16 * nkv * hd, pi.data_ptr(),
        ip.data_ptr(), lpl.data_ptr(), qi.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), kc.data_ptr(), tn.data_ptr(), nq, nkv, hd, 16, T, 1,
        1, stride, 1 / math.sqrt(hd), 64, st)
    ms = timeit(fn)
    fl = 4.0 * nq * hd * T * (T + 1) / 2
    print(f"prefill attention T={T}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s (causal flops)")
