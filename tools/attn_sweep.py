"""Fused decode-attention launch cost vs cluster size / ring depth (one process per setting: the library reads
PK_ATTN_CLUSTER / PK_ATTN_SLOTS once).  Train protocol of tests/tools/bench_decode_micro.py: N launches over distinct cold
KV pools in one CUDA graph, per-launch time = total / N, with and without programmatic dependent launch.

    PK_ATTN_CLUSTER=8 python tools/attn_sweep.py            # prints one line per context length
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pegainfer_b200 import ffi  # noqa: E402

lib = ffi.lib()
lib.cuda_set_device(0)
lib.cublas_init()
torch.zeros(1, device="cuda")
flush = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")
sink = torch.zeros((), dtype=torch.int64, device="cuda")
nq, nkv, hd, ps = 32, 8, 128, 16
sm = 1 / math.sqrt(hd)
i32 = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
env = {k: v for k, v in os.environ.items() if k.startswith("PK_")}
for seq in [int(s) for s in (sys.argv[1:] or ["1", "128", "2304", "4096"])]:
    pages = (seq + ps - 1) // ps
    stride = 2 * ps * nkv * hd
    nc = 32
    kvs = [(torch.randn((pages + 1) * stride, device="cuda") * 0.1).to(torch.bfloat16) for _ in range(nc)]
    q = torch.randn(nq * hd, device="cuda").to(torch.bfloat16)
    out = torch.empty_like(q)
    k1 = torch.randn(nkv * hd, device="cuda").to(torch.bfloat16)
    v1 = torch.randn(nkv * hd, device="cuda").to(torch.bfloat16)
    qn = torch.ones(hd, device="cuda", dtype=torch.bfloat16)
    cos = torch.ones((8192, hd), device="cuda", dtype=torch.bfloat16)
    sin = torch.zeros((8192, hd), device="cuda", dtype=torch.bfloat16)
    pi, ip, lpl, pos = i32(list(range(pages))), i32([0, pages]), i32([((seq - 1) % ps) + 1]), i32([seq - 1])
    partial = torch.zeros(64 * nq * (hd + 2) * 2, device="cuda", dtype=torch.float32)
    counters = torch.zeros(64, device="cuda", dtype=torch.int32)

    def launch(i, s_):
        rc = lib.pk_b200_decode_attention_fused(q.data_ptr(), k1.data_ptr(), v1.data_ptr(), out.data_ptr(), kvs[i % nc].data_ptr(), 0,
                                                ps * nkv * hd, pi.data_ptr(), ip.data_ptr(), lpl.data_ptr(), pos.data_ptr(), qn.data_ptr(),
                                                qn.data_ptr(), cos.data_ptr(), sin.data_ptr(), 1e-6, partial.data_ptr(), counters.data_ptr(),
                                                64, 37, nq, nkv, hd, ps, 1, stride, sm, s_)
        assert rc == 0, rc

    res = {}
    for pdl in (0, 1):
        lib.pk_b200_set_pdl(pdl)
        st = torch.cuda.current_stream().cuda_stream
        for i in range(nc):
            launch(i, st)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for i in range(nc):
                launch(i, side.cuda_stream)
        ts = []
        for _ in range(7):
            sink.copy_(flush.view(torch.int64).sum())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / nc)
        ts.sort()
        res["pdl" if pdl else "nopdl"] = round(ts[len(ts) // 2], 2)
    print(f"ATTN seq={seq} us/launch {res} env={env}", flush=True)
    del kvs
