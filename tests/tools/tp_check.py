"""Multi-GPU check under torchrun: (1) the one-shot NVLink all-reduce (plain and fused with add+RMSNorm)
against torch / the CPU oracle, (2) Qwen3 TP-N prefill + decode logits against the CPU oracle's TP-N model.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/tools/tp_check.py [--model qwen3-small|qwen3-8b] [--prompt 40] [--steps 6]

`--model qwen3-8b` is BASELINE config 3 at full size (36 layers, V = 151,936, untied lm_head): rank 0 generates the
seed-0 CPU checkpoint (the one bench.py loads), runs the TP-N oracle on the host cores and broadcasts every tensor to
the other ranks over NCCL; the GPU side is teacher-forced with the oracle's tokens (SURVEY 8c rule).
tests/test_tp_gpu.py runs this file under pytest when the box has >= 2 GPUs.
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import make_tp_comm  # noqa: E402
from oracle import qwen3_oracle as O  # noqa: E402
from pegainfer_b200 import ffi  # noqa: E402
from pegainfer_b200.config import PRESETS, TensorParallelConfig  # noqa: E402
from pegainfer_b200.model import ModelRuntimeConfig, Qwen3Model  # noqa: E402
from pegainfer_b200.synthetic import iter_random_weights, synthetic_prompt, to_numpy_bits, weight_shapes  # noqa: E402
from tests.helpers import bits, logits_agree, oracle_cfg  # noqa: E402


def broadcast_weights(cfg, rank, keep_cpu):
    """Rank 0 generates the seed-0 CPU checkpoint tensor by tensor and broadcasts it; yields (name, cuda tensor) on
    every rank.  With `keep_cpu` (rank 0) the CPU tensors are collected into that dict for the oracle."""
    gen = iter_random_weights(cfg, seed=0, device="cpu", norm_jitter=0.1 if cfg.num_hidden_layers <= 4 else 0.0) if rank == 0 else None
    for name, shape in weight_shapes(cfg).items():
        if rank == 0:
            n2, t = next(gen)
            assert n2 == name
            if keep_cpu is not None:
                keep_cpu[name] = t
            d = t.cuda()
        else:
            d = torch.empty(shape, dtype=torch.bfloat16, device="cuda")
        dist.broadcast(d, 0)
        yield name, d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="qwen3-small")
    ap.add_argument("--prompt", type=int, default=40)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--tol", type=float, default=None)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    lib = ffi.lib()
    lib.cuda_set_device(local)
    lib.cublas_init()
    H = 4096
    comm = make_tp_comm(rank, world, dist, max_tokens=256, hidden=H)
    st = torch.cuda.current_stream().cuda_stream
    ok = True
    # ---- (1a) in-place all-reduce, several sizes, repeated (sequence / slot reuse) ----
    t_max = min(200, int(lib.pk_tp_max_rows(comm, H)))  # the LL protocol (PK_TP_PROTO=ll) reserves part of each region
    for T in (1, 3, 64, t_max):
        for rep in range(3):
            g = torch.Generator(device="cuda").manual_seed(1000 * rep + 10 * T + rank)
            x = torch.randn((T, H), generator=g, device="cuda").to(torch.bfloat16)
            gathered = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(gathered, x)
            want = torch.stack([t.float() for t in gathered]).sum(0).to(torch.bfloat16)
            y = x.clone()
            rc = lib.pk_tp_all_reduce_rows(comm, y.data_ptr(), H, T, st)
            torch.cuda.synchronize()
            same = bool((y.view(torch.int16) == want.view(torch.int16)).all())
            ok &= rc == 0 and same
            if not (rc == 0 and same):
                print(f"[rank {rank}] all_reduce T={T} rep={rep}: rc={rc} same={same}", flush=True)
            if rank == 0 and not same:
                print(f"all_reduce T={T} rep={rep}: MISMATCH max|d|={(y.float() - want.float()).abs().max().item()}")
    # ---- (1b) fused all-reduce + residual add + RMSNorm vs the oracle ----
    for T in (1, 4):
        g = torch.Generator(device="cuda").manual_seed(77 + rank)
        part = torch.randn((T, H), generator=g, device="cuda").to(torch.bfloat16)
        g2 = torch.Generator(device="cuda").manual_seed(5)
        hidden = torch.randn((T, H), generator=g2, device="cuda").to(torch.bfloat16)
        w = (torch.randn((H,), generator=g2, device="cuda") * 0.1 + 1).to(torch.bfloat16)
        gathered = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(gathered, part)
        red = O.all_reduce_sum([bits(t) for t in gathered])
        h_np = bits(hidden).copy()
        want = O.fused_add_rms_norm(h_np, red, bits(w), 1e-6)
        out = torch.zeros_like(hidden)
        h_d = hidden.clone()
        rc = lib.pk_tp_all_reduce_add_rms_norm(comm, h_d.data_ptr(), part.data_ptr(), w.data_ptr(), out.data_ptr(), H, T,
                                               1e-6, st)
        torch.cuda.synchronize()
        e_h = bool((bits(h_d) == h_np).all())
        err = np.abs(O.bf16_to_f32(bits(out)) - O.bf16_to_f32(want)) / O.bf16_ulp(np.abs(O.bf16_to_f32(want)))
        ok &= rc == 0 and e_h and err.max() <= 1
        if not (rc == 0 and e_h and err.max() <= 1):
            print(f"[rank {rank}] fused allreduce+add+norm T={T}: rc={rc} hidden exact={e_h} err={err.max():.2f}", flush=True)
        if rank == 0:
            print(f"fused allreduce+add+norm T={T}: hidden exact={e_h} max out err={err.max():.2f} ulp")
    # ---- (2) model parity: TP-N on GPUs vs TP-N oracle ----
    cfg = PRESETS[args.model]
    tol = args.tol if args.tol is not None else (6 if cfg.num_hidden_layers <= 4 else 12)  # tests/test_fullsize_gpu.py TOL_ULP
    n_steps = args.steps
    prompt = [t % cfg.vocab_size for t in synthetic_prompt(args.prompt)]
    pages = (args.prompt + n_steps) // 16 + 4
    comm2 = make_tp_comm(rank, world, dist, max_tokens=256, hidden=cfg.hidden_size)
    rt = ModelRuntimeConfig(tensor_parallel=TensorParallelConfig(rank, world), device_ordinal=local, num_pages=pages,
                            max_batch=1)
    cpu_w = {} if rank == 0 else None
    m = Qwen3Model(cfg, broadcast_weights(cfg, rank, cpu_w), rt, tp_comm=comm2)
    kv = m.alloc_kv()
    got = [bits(m.gather_logits(m.prefill([prompt], [kv]), dist)[0])]
    if rank == 0:
        O.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
        orc = O.OracleQwen3(oracle_cfg(cfg), to_numpy_bits(cpu_w), tp_world=world, num_pages=pages)
        okv = orc.alloc_kv()
        want = [orc.prefill([prompt], [okv])[0]]
    toks = torch.zeros(n_steps, dtype=torch.int64, device="cuda")
    if rank == 0:
        for i in range(n_steps):
            toks[i] = O.argmax(want[-1])
            want.append(orc.decode([int(toks[i])], [okv])[0])
    dist.broadcast(toks, 0)
    for i in range(n_steps):
        lg, sampled = m.decode([int(toks[i])], [kv])
        full = m.gather_logits(lg, dist)
        # the vocab-sharded greedy token (max/index exchange) must be the arg-max of the gathered row, lowest index on ties
        row = full[0].float()
        first_max = int((row == row.max()).nonzero()[0])
        if int(sampled[0]) != first_max:
            print(f"[rank {rank}] step {i}: sampled {int(sampled[0])} != first arg-max {first_max} (value {float(row.max())})", flush=True)
        ok &= int(sampled[0]) == first_max
        got.append(bits(full[0]))
    if rank == 0:
        worst = 0.0
        for i, (g_, w_) in enumerate(zip(got, want)):
            good, info = logits_agree(g_, w_, tol)
            worst = max(worst, info["err"] / (info["tol"] / tol))
            ok &= good
            print(f"{cfg.name} TP{world} step {i}: ok={good} {info}")
        print(f"{cfg.name} TP{world}: worst |dlogit| = {worst:.2f} ulp(rowmax) over {len(got)} steps (tolerance {tol})")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("TP_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL")
    m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
