#!/usr/bin/env python
"""Headline benchmark: Qwen3 greedy decode tok/s (+ TTFT) on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1 : BASELINE config[1] -- Qwen3-4B bf16, 2048-token prefill + decode at ctx 2048.., CUDA Graph on,
        single request.  A "step" is one decode step (one generated token).
N > 1 : BASELINE config[2] -- Qwen3-8B bf16, tensor parallel over N ranks (one process per GPU,
        NVLink peer-memory all-reduce), 128-token prompt, decode steps.  scaling = "strong".
Random-init weights of the exact architecture (N(0, 0.02), seed 0) and synthetic prompt ids
((i % 1000) + 100, bench_serving.rs:761-763); there is no network for checkpoints.

Keys of the JSON line (see the task contract):
  value     decode tok/s with inputs resident in HBM (per-step metadata pre-staged on the device, token
            fed back on the device, no host sync inside the timed region; CUDA events on the stream)
  e2e       the same metric through the public API with HOST token ids: per step one H2D metadata copy
            from pinned memory, the graph launch, a 4-byte D2H of the sampled token and a stream sync
  ttft_tensor  TTFT against the tensor roofline: prompt GEMM + causal attention flops / TTFT vs the measured sustained bf16 peak
  roofline  the dominant kernel (the HBM-streaming decode GEMV): algorithmic weight bytes per token /
            CUDA-event time of the token's 145 GEMV launches, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (port of the reference forward) timed on the host cores, bounded sample
`--impl reference` times that CPU port alone (the reference has no CPU path of its own and its Rust
host cannot be built here; oracle/_ref holds its CUDA kernels, which are GPU code, not a CPU arm).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------ helpers
def weight_bytes_per_token(cfg, world=1):
    """SURVEY.md 8d / BASELINE.md 2: decode weight bytes per token per GPU."""
    H, I, L, V = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.vocab_size
    per_layer = 2 * (H * (cfg.q_dim + 2 * cfg.kv_dim) + H * cfg.q_dim + 3 * H * I)
    return L * per_layer // world + 2 * V * H


def kv_bytes_per_ctx_token(cfg, world=1):
    return cfg.num_hidden_layers * 2 * cfg.kv_dim * 2 // world


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def prefill_tensor_summary(cfg, prompt_len, ttft_ms, world):
    """TTFT against the tensor roofline: dense GEMM flops of the prompt + causal attention + one lm_head row,
    per GPU, over the measured TTFT (which also holds the host-side planning and the first sampling)."""
    try:
        h, i, l, v = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.vocab_size
        per_tok = 2.0 * l * (h * (cfg.q_dim + 2 * cfg.kv_dim) + cfg.q_dim * h + 3 * h * i)
        attn = 4.0 * cfg.num_attention_heads * cfg.head_dim * prompt_len * (prompt_len + 1) / 2 * l
        flop = (per_tok * prompt_len + attn) / world + 2.0 * h * v
        peak, src = 1416.7, "fallback"
        pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pth):
            d = json.load(open(pth))
            peak, src = float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", peak))), "measured (MEASURED_PEAKS.json, sustained)"
        ach = flop / (ttft_ms * 1e-3) / 1e12
        return {"bound": "tensor", "flop_per_gpu": flop, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "peak_source": src, "note": "end-to-end TTFT, not kernel time"}
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_oracle(cfg, weights_np, num_pages):
    from oracle import qwen3_oracle as O
    oc = O.OracleConfig(cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                        cfg.num_key_value_heads, cfg.head_dim, cfg.vocab_size, cfg.rms_norm_eps, cfg.rope_theta,
                        cfg.tie_word_embeddings)
    return O, O.OracleQwen3(oc, weights_np, num_pages=num_pages)


def cpu_decode_rate(cfg, weights_np, steps, budget_s, ctx=16):
    """Time `steps` full-depth greedy decode steps of the CPU oracle (OpenMP, all host cores).
    Context is `ctx` tokens (the weight stream is > 96 % of the bytes at the metric's ctx anyway)."""
    from pegainfer_b200.synthetic import synthetic_prompt
    O, orc = make_oracle(cfg, weights_np, num_pages=(ctx + steps + 2) // 16 + 4)  # whole run fits: prompt + warm-up + steps
    kv = orc.alloc_kv()
    lg = orc.prefill([synthetic_prompt(ctx)], [kv])[0]
    tok = O.argmax(lg)
    t_warm = time.perf_counter()
    tok = O.argmax(orc.decode([tok], [kv])[0])  # warm-up step
    t_step = time.perf_counter() - t_warm
    n = max(1, min(steps, int(budget_s / max(t_step, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        tok = O.argmax(orc.decode([tok], [kv])[0])
    dt = time.perf_counter() - t0
    return n / dt, n, dt


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args, cfg, rank, world):
    """CPU port of the reference forward on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import torch
    from pegainfer_b200.synthetic import random_weights, to_numpy_bits
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    w = to_numpy_bits(random_weights(cfg, seed=0, device="cpu"))
    gen_s = time.perf_counter() - t0
    rate, n, dt = cpu_decode_rate(cfg, w, args.steps, budget_s=150.0)
    sample = (f"{n} of the requested {args.steps} full-depth {cfg.name} greedy decode steps on the CPU oracle "
              f"(ctx 17.., bs 1), OpenMP over {cores} host threads; weights generated on CPU in {gen_s:.0f}s")
    line = {"impl": "reference", "metric": "decode_tok_s", "value": rate, "unit": "tok/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / rate, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": bench_workload_name(cfg, world)},
            "cpu_baseline": {"value": rate, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bench_workload_name(cfg, world):
    if world == 1:
        return f"{cfg.name} bf16 random-init, 2048-token prefill + greedy decode at ctx 2048.., bs 1, CUDA Graph on"
    return f"{cfg.name} bf16 random-init, tensor-parallel x{world} (NVLink peer all-reduce), 128-token prompt + greedy decode, bs 1"


# ------------------------------------------------------------------------------------ TP plumbing
def make_tp_comm(rank, world, dist, max_tokens, hidden):
    """Symmetric staging + flags via cudaIpc handles exchanged over torch.distributed."""
    import ctypes as C
    import torch
    from pegainfer_b200 import ffi
    lib = ffi.lib()
    cudart = C.CDLL("libcudart.so.12")
    staging_bytes = 2 * world * max(64 * 1024, max_tokens * hidden * 2)
    flag_bytes = int(lib.pk_tp_flag_bytes())

    def dev_alloc(n):
        p = C.c_void_p()
        assert cudart.cudaMalloc(C.byref(p), C.c_size_t(n)) == 0
        assert cudart.cudaMemset(p, 0, C.c_size_t(n)) == 0
        return p.value

    my_stage, my_flags = dev_alloc(staging_bytes), dev_alloc(flag_bytes)
    hs, hf = (C.c_char * 64)(), (C.c_char * 64)()
    assert lib.pk_tp_ipc_export(my_stage, hs) == 0 and lib.pk_tp_ipc_export(my_flags, hf) == 0
    gathered = [None] * world
    dist.all_gather_object(gathered, (bytes(hs), bytes(hf)))
    stage_ptrs, flag_ptrs = (C.c_void_p * world)(), (C.c_void_p * world)()
    for p, (bs_, bf_) in enumerate(gathered):
        if p == rank:
            stage_ptrs[p], flag_ptrs[p] = my_stage, my_flags
        else:
            a, b = C.c_void_p(), C.c_void_p()
            assert lib.pk_tp_ipc_open(bs_, C.byref(a)) == 0, "cudaIpcOpenMemHandle failed (staging)"
            assert lib.pk_tp_ipc_open(bf_, C.byref(b)) == 0, "cudaIpcOpenMemHandle failed (flags)"
            stage_ptrs[p], flag_ptrs[p] = a.value, b.value
    torch.cuda.synchronize()
    dist.barrier()
    comm = lib.pk_tp_comm_create(rank, world, stage_ptrs, flag_ptrs, staging_bytes)
    assert comm, "pk_tp_comm_create failed"
    return comm


# ------------------------------------------------------------------------------------ our arm
def run_ours(args, cfg, rank, world, dist):
    import torch
    from pegainfer_b200.config import TensorParallelConfig
    from pegainfer_b200.model import ModelRuntimeConfig, Qwen3Model
    from pegainfer_b200.synthetic import iter_random_weights, synthetic_prompt

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    prompt_len = 2048 if world == 1 else 128
    K, W = args.steps, args.warmup
    ctx_max = prompt_len + 2 * (K + W) + 64
    if ctx_max > 4096:
        raise SystemExit(f"prompt {prompt_len} + steps exceed the reference's 4096-position RoPE table")
    pages = 3 * (ctx_max // 16 + 2) + 8
    tp_comm = make_tp_comm(rank, world, dist, max_tokens=256, hidden=cfg.hidden_size) if world > 1 else None
    persistent = world == 1 and os.environ.get("PK_DECODE", "fused") == "persistent"
    rt = ModelRuntimeConfig(enable_cuda_graph=True, tensor_parallel=TensorParallelConfig(rank, world),
                            device_ordinal=local_rank, fused=True, persistent=persistent, num_pages=pages,
                            max_batch=1, enable_pdl=True)
    t0 = time.perf_counter()
    model = Qwen3Model(cfg, iter_random_weights(cfg, seed=0, device="cuda"), rt, tp_comm=tp_comm)
    load_s = time.perf_counter() - t0
    prompt = synthetic_prompt(prompt_len)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- warm-up: graph capture, allocator, clocks ----
    model.generate(prompt, 4)
    # ---- TTFT: prompt submit -> first token (host ids in, token out), median of 3 ----
    ttfts = []
    for _ in range(3):
        barrier()
        _, ttft, _ = model.generate(prompt, 1)
        ttfts.append(ttft)
    ttft_ms = statistics.median(ttfts)

    clocks = ClockSampler(local_rank).start() if rank == 0 else None
    # ---- e2e: public API, host token ids, per-step H2D + D2H + sync ----
    kv = model.alloc_kv()
    tok = model.sample_greedy(model.prefill([prompt], [kv])[0])
    for _ in range(W):
        _, s = model.decode([tok], [kv], want_logits=False)
        tok = s[0]
    barrier()
    e0 = model.event_record()
    for _ in range(K):
        _, s = model.decode([tok], [kv], want_logits=False)
        tok = s[0]
    e1 = model.event_record()
    e2e_ms = model.event_elapsed_ms(e0, e1)
    barrier()
    # ---- device-resident: K more steps, metadata staged in HBM, token fed back on the device ----
    burst_tokens, burst_ms = model.decode_burst(kv, tok, K)
    ctx_lo = prompt_len + 1 + W + K
    model.drop_request(kv)
    barrier()
    clk = clocks.stop() if clocks else None

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e2e_ms, burst_ms, ttft_ms = max_over_ranks(e2e_ms), max_over_ranks(burst_ms), max_over_ranks(ttft_ms)
    launches_per_step = model.launches_per_step()

    # ---- roofline of the dominant kernel (decode GEMV), live CUDA-event timing ----
    peak, peak_src = measured_peaks()
    wbytes = weight_bytes_per_token(cfg, world)
    roof = None
    if world == 1:
        ms_pass, n_gemv = model.bench_gemv_pass(20)
        ach = wbytes / (ms_pass * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "gemv_stream_kernel (145 launches/token: 36 x {qkv, o, gate_up+SwiGLU, down} + lm_head)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                "launches_per_pass": n_gemv, "ms_per_pass": ms_pass,
                "algorithmic_bytes_per_launch": wbytes / n_gemv, "traffic": profile_traffic()}
    step_bytes = wbytes + kv_bytes_per_ctx_token(cfg, world) * (ctx_lo + K // 2)
    step_ms = burst_ms / K
    step_gbs = step_bytes / (step_ms * 1e-3) / 1e9

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(cfg)

    if rank == 0:
        value = K / (burst_ms * 1e-3)
        line = {"metric": "decode_tok_s", "value": value, "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": bench_workload_name(cfg, world), "prompt_len": prompt_len,
                           "decode_ctx": [ctx_lo, ctx_lo + K], "l2": "inputs (weights 8-15 GB/token) >> 126 MB L2, no flush needed",
                           "parallelism": f"tp{world}", "cuda_graph": True, "pdl": True,
                           "decode_impl": "persistent single-launch step" if persistent else "fused multi-kernel graph",
                           "launches_per_step": launches_per_step, "model_load_s": round(load_s, 1)},
                "ttft_ms": ttft_ms, "ttft_prompt_len": prompt_len,
                "e2e": {"value": K / (e2e_ms * 1e-3), "unit": "tok/s", "ms_per_step": e2e_ms / K,
                        "h2d_bytes_per_step": model.meta_bytes(), "d2h_bytes_per_step": 4},
                "gpu_launches": launches_per_step * K,
                "step_hbm": {"bytes_per_step": step_bytes, "achieved_gbs": step_gbs, "frac_of_peak": step_gbs / peak},
                "clocks": clk}
        if roof:
            line["roofline"] = roof
        else:
            line["roofline"] = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                                "frac": step_gbs / peak, "peak_source": peak_src, "traffic": None,
                                "note": "whole decode step per GPU (weights + KV bytes / device time); per-kernel leg runs at N=1"}
        if cpu:
            line["cpu_baseline"] = cpu
        pts = prefill_tensor_summary(cfg, prompt_len, ttft_ms, world)
        if pts:
            line["ttft_tensor"] = pts
        print(json.dumps(line), flush=True)
    model.close()


def profile_traffic():
    """dram bytes per GEMV launch from the committed ncu summary, if present (profiles/)."""
    p = os.path.join(ROOT, "profiles", "gemv_traffic.json")
    try:
        return json.load(open(p))["dram_bytes_per_launch"]
    except Exception:
        return None


def cpu_baseline_leg(cfg):
    import torch
    from pegainfer_b200.synthetic import random_weights, to_numpy_bits
    cores = os.cpu_count() or 1
    w = to_numpy_bits(random_weights(cfg, seed=0, device="cpu"))
    rate, n, dt = cpu_decode_rate(cfg, w, steps=64, budget_s=20.0)
    return {"value": rate, "unit": "tok/s", "cores": cores, "kind": "port",
            "sample": f"{n} full-depth {cfg.name} greedy decode steps ({dt:.1f} s) of the CPU oracle at ctx 17.., "
                      f"OpenMP over {cores} host threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default=None, help="override: qwen3-4b | qwen3-8b | qwen3-small | qwen3-tiny")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    from pegainfer_b200.config import PRESETS
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    cfg = PRESETS[args.model] if args.model else PRESETS["qwen3-4b" if world == 1 else "qwen3-8b"]
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist_mod.init_process_group("nccl")
        dist = dist_mod
    run_ours(args, cfg, rank, world, dist)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
