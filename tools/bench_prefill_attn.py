"""CUDA-event timing of the paged prefill attention kernels (Qwen3-4B heads: 32 q / 8 kv, head dim 128) per
implementation (PK_PREFILL_ATTN is read per call): tc2 = two query tiles per CTA, O and P in TMEM; tc = one tile per CTA;
legacy = mma.sync."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pegainfer_b200 import ffi  # noqa: E402

lib = ffi.lib()
torch.zeros(1, device="cuda")
lib.cuda_set_device(0)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


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


i32 = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
nq, nkv, hd = 32, 8, 128
impls = sys.argv[1:] or ["tc2", "tc", "legacy"]
for T in [int(x) for x in os.environ.get("PK_T", "128,512,2048,8192").split(",")]:
    pages = T // 16 + 1
    stride = 2 * 16 * nkv * hd
    kv = torch.randn(((pages + 2) * stride,), device="cuda").to(torch.bfloat16)
    q = torch.randn((T, nq * hd), device="cuda").to(torch.bfloat16)
    outs = {}
    pi, ip, lpl = i32(list(range(1, pages + 1))), i32([0, pages]), i32([((T - 1) % 16) + 1])
    qi, z, kc, tn = i32([0, T]), i32([0] * 4096), i32([T]), i32([T])
    fl = 4.0 * nq * hd * T * (T + 1) / 2
    for impl in impls:
        os.environ["PK_PREFILL_ATTN"] = impl
        out = torch.zeros_like(q)
        fn = lambda: lib.batch_prefill_paged_cuda_with_cta_tile_q(q.data_ptr(), out.data_ptr(), kv.data_ptr(), 0, 16 * nkv * hd, pi.data_ptr(),
            ip.data_ptr(), lpl.data_ptr(), qi.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), kc.data_ptr(), tn.data_ptr(), nq, nkv, hd, 16, T, 1,
            1, stride, 1 / math.sqrt(hd), 64, st)
        assert fn() == 0
        torch.cuda.synchronize()
        ms = timeit(fn)
        outs[impl] = out.float()
        d = (outs[impl] - outs[impls[0]]).abs().max().item()
        print(f"prefill attention T={T:5d} {impl:7s}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s (causal flops)   max |diff vs {impls[0]}| {d:.4f}  finite {bool(torch.isfinite(outs[impl]).all())}")
