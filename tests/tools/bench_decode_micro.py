"""BASELINE config 5: decode GEMV + GQA decode-attention microbench, cold L2, CUDA events.

Protocol after the reference's kernel bench (pegainfer-qwen3-4b/src/kernel_bench.rs:236-282,569-586): a streaming
sweep of a buffer >= 2 x L2 before every timed launch, one untimed pre-launch, N timed launches each bracketed by its
own event pair.  GEMV: every Qwen3-4B / 8B projection shape, bytes = 2*M*K + 2K + 2M.
Attention: the reference's own case (1 layer, nq 32, nkv 8, hd 128, page 16, bs 1, patterned q / kv), bytes =
4096*seq + 16 KiB, seq in {1, 128, 1024, 4096}.  When oracle/_ref/libkernels_ref.so is present the reference's own
kernels (cuBLAS GEMV, FlashInfer decode) are timed beside ours under the same protocol.

Round 2 adds the `train` protocol next to the single-launch one (single CUDA-event launches quantise to ~2 us and
cannot resolve the small shapes): N >= 32 launches over N DISTINCT cold buffers (together >= 2 x L2, after an L2
sweep), captured into ONE CUDA graph so no host launch gap sits between them, one event pair around the replay;
per-launch time = total / N.  For our library it is reported with programmatic dependent launch off and on (on = what
the decode graph runs: the next launch's weight / KV requests overlap this one's tail).

Prints one JSON object; `python tests/tools/bench_decode_micro.py > gpurun_out/micro.json`.
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from pegainfer_b200 import ffi  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ITERS = int(os.environ.get("MICRO_ITERS", "32"))
torch.zeros(1, device="cuda")
st = torch.cuda.current_stream().cuda_stream
flush = torch.zeros(512 << 20, dtype=torch.uint8, device="cuda")  # 4 x the 126 MB L2
flush_sink = torch.zeros((), dtype=torch.int64, device="cuda")


def cold_time(fn, iters=ITERS):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_sink.copy_(flush.view(torch.int64).sum())  # read-only sweep: L2 ends up full of clean lines
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def train_time(launch, n, reps=5):
    """`launch(i, stream)` enqueues the kernel on buffer copy i.  n launches in one CUDA graph, cold L2 before each
    replay; returns (median, min) microseconds PER LAUNCH."""
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    for i in range(n):
        launch(i, st)  # warm-up outside capture (lazy attribute setup, tensor maps)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=side):
        for i in range(n):
            launch(i, side.cuda_stream)
    ts = []
    for _ in range(reps):
        flush_sink.copy_(flush.view(torch.int64).sum())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def copies_for(nbytes):
    return int(max(32, min(64, math.ceil((256 << 20) / max(nbytes, 1))))) if nbytes < (64 << 20) else (4 if nbytes < (1 << 29) else 2)


def stat(by, med, best, peak, digits=3):
    return {"us": round(med, 2), "us_min": round(best, 2), "gbs": round(by / med / 1e3, 1), "frac": round(by / med / 1e3 / peak, digits)}


def load_libs():
    libs = {"b200": ffi.lib()}
    ref = os.path.join(ROOT, "oracle", "_ref", "libkernels_ref.so")
    if os.path.exists(ref):
        libs["reference"] = ffi.load(ref, extensions=False)
    for lib in libs.values():
        lib.cuda_set_device(0)
        lib.cublas_init()
    return libs


def pattern(n, scale):
    i = torch.arange(n, device="cuda", dtype=torch.int64)
    return (((i % 251) - 125).to(torch.float32) * scale).to(torch.bfloat16)


def bench_gemv(libs, peak):
    shapes = [("4b.q", 4096, 2560), ("4b.kv", 1024, 2560), ("4b.qkv", 6144, 2560), ("4b.o", 2560, 4096), ("4b.gate_up", 19456, 2560),
              ("4b.down", 2560, 9728), ("4b.lm_head", 151936, 2560), ("8b.qkv", 6144, 4096), ("8b.o", 4096, 4096),
              ("8b.gate_up", 24576, 4096), ("8b.down", 4096, 12288), ("8b.lm_head", 151936, 4096)]
    rows = []
    for name, M, K in shapes:
        W = (torch.randn((M, K), device="cuda") * 0.02).to(torch.bfloat16)
        X = torch.randn((K,), device="cuda").to(torch.bfloat16)
        Y = torch.empty((M,), device="cuda", dtype=torch.bfloat16)
        by = 2.0 * M * K + 2 * K + 2 * M
        row = {"shape": name, "M": M, "K": K, "bytes": by}
        for tag, lib in libs.items():
            med, best = cold_time(lambda: lib.gemm_cuda(W.data_ptr(), X.data_ptr(), Y.data_ptr(), M, 1, K, st))
            row[tag] = stat(by, med, best, peak)
        nc = copies_for(by)
        Ws = [W] + [W.clone() for _ in range(nc - 1)]
        row["train_launches"] = max(32, nc)
        for tag, lib in libs.items():
            for pdl in ((0, 1) if tag == "b200" else (0,)):
                if tag == "b200":
                    lib.pk_b200_set_pdl(pdl)
                med, best = train_time(lambda i, s_: lib.gemm_graphsafe_cuda(Ws[i % nc].data_ptr(), X.data_ptr(), Y.data_ptr(), M, 1, K, s_),
                                       max(32, nc))
                row[tag + (".train_pdl" if pdl else ".train")] = stat(by, med, best, peak)
            if tag == "b200":
                lib.pk_b200_set_pdl(0)
        rows.append(row)
        del W, Ws
    return rows


def bench_attention(libs, peak):
    nq, nkv, hd, ps = 32, 8, 128, 16
    sm = 1 / math.sqrt(hd)
    rows = []
    i32 = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
    b200 = libs["b200"]
    for seq in (1, 128, 1024, 4096):
        pages = (seq + ps - 1) // ps
        stride = 2 * ps * nkv * hd  # one layer: [K|V][16][nkv][hd]
        kv = pattern((pages + 1) * stride, 0.001)
        q = pattern(nq * hd, 0.01)
        out = torch.empty_like(q)
        pi, ip, lpl = i32(list(range(pages))), i32([0, pages]), i32([((seq - 1) % ps) + 1])
        by = 4096.0 * seq + 16384
        row = {"seq": seq, "bytes": by}
        # ABI entries (reference plan: non-partition below 1024 tokens, split-KV above; batch_decode_buffers.rs:281-287)
        csz = max(256, -(-seq // 64))
        n = max(1, -(-seq // csz))
        req, tile = i32([0] * 64), i32(list(range(n)) + [0] * (64 - n))
        mask = torch.tensor([1] * n + [0] * (64 - n), dtype=torch.uint8, device="cuda")
        oip, csz_d, full = i32([0, n]), i32([csz]), i32([seq])
        tmp_v = torch.zeros((64, nq * hd), dtype=torch.bfloat16, device="cuda")
        tmp_s = torch.zeros((64, nq), dtype=torch.float32, device="cuda")
        nc = copies_for(by)
        kvs = [kv] + [kv.clone() for _ in range(nc - 1)]
        row["train_launches"] = nc
        for tag, lib in libs.items():
            if seq < 1024:
                fn = lambda i=0, s_=st: lib.paged_attention_decode_cuda(q.data_ptr(), out.data_ptr(), kvs[i % nc].data_ptr(), 0, ps * nkv * hd, pi.data_ptr(),
                    ip.data_ptr(), lpl.data_ptr(), req.data_ptr(), tile.data_ptr(), full.data_ptr(), nq, nkv, hd, ps, 1, stride, sm, s_)
            else:
                fn = lambda i=0, s_=st: lib.paged_attention_decode_split_kv_cuda(q.data_ptr(), out.data_ptr(), kvs[i % nc].data_ptr(), 0, ps * nkv * hd,
                    pi.data_ptr(), ip.data_ptr(), lpl.data_ptr(), req.data_ptr(), tile.data_ptr(), csz_d.data_ptr(), oip.data_ptr(),
                    mask.data_ptr(), tmp_v.data_ptr(), tmp_s.data_ptr(), nq, nkv, hd, ps, 1, 64, stride, sm, s_)
            med, best = cold_time(fn)
            row[tag + ".abi"] = stat(by, med, best, peak, 4)
            med, best = train_time(fn, nc)
            row[tag + ".abi.train"] = stat(by, med, best, peak, 4)
        # the fused B200 entry (QK-norm + RoPE + KV append + attention + merge in one launch)
        k1, v1 = pattern(nkv * hd, 0.001), pattern(nkv * hd, 0.001)
        qn, kn = torch.ones(hd, device="cuda", dtype=torch.bfloat16), torch.ones(hd, device="cuda", dtype=torch.bfloat16)
        cos, sin = torch.ones((8192, hd), device="cuda", dtype=torch.bfloat16), torch.zeros((8192, hd), device="cuda", dtype=torch.bfloat16)
        pos = i32([seq - 1])
        max_chunks = min(64, (2 * torch.cuda.get_device_properties(0).multi_processor_count + nkv - 1) // nkv)  # as the host sizes it
        partial = torch.zeros(64 * nq * (hd + 2) * 2, device="cuda", dtype=torch.float32)
        counters = torch.zeros(64, device="cuda", dtype=torch.int32)
        fn = lambda i=0, s_=st: b200.pk_b200_decode_attention_fused(q.data_ptr(), k1.data_ptr(), v1.data_ptr(), out.data_ptr(), kvs[i % nc].data_ptr(), 0,
            ps * nkv * hd, pi.data_ptr(), ip.data_ptr(), lpl.data_ptr(), pos.data_ptr(), qn.data_ptr(), kn.data_ptr(), cos.data_ptr(),
            sin.data_ptr(), 1e-6, partial.data_ptr(), counters.data_ptr(), 64, max_chunks, nq, nkv, hd, ps, 1, stride, sm, s_)
        med, best = cold_time(fn)
        row["b200.fused"] = stat(by, med, best, peak, 4)
        for pdl in (0, 1):
            b200.pk_b200_set_pdl(pdl)
            med, best = train_time(fn, nc)
            row["b200.fused.train_pdl" if pdl else "b200.fused.train"] = stat(by, med, best, peak, 4)
        b200.pk_b200_set_pdl(0)
        del kvs
        rows.append(row)
    return rows


def main():
    peak = 6573.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    libs = load_libs()
    res = {"protocol": f"single: cold L2 (512 MiB read sweep before each launch), median of {ITERS} single launches, CUDA events; train: >= 32 launches over distinct cold buffers in one CUDA graph, per-launch = total / N, median of 5 replays", "hbm_peak_gbs": peak,
           "gemv": bench_gemv(libs, peak), "decode_attention": bench_attention(libs, peak)}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
