"""Multi-GPU check under torchrun: (1) the one-shot NVLink all-reduce (plain and fused with add+RMSNorm)
against torch / the CPU oracle, (2) Qwen3 TP-N prefill + decode logits against the CPU oracle's TP-N model.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/tp_check.py
"""
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
from pegainfer_b200.config import QWEN3_SMALL, TensorParallelConfig  # noqa: E402
from pegainfer_b200.model import ModelRuntimeConfig, Qwen3Model  # noqa: E402
from pegainfer_b200.synthetic import random_weights, synthetic_prompt, to_numpy_bits  # noqa: E402
from tests.helpers import bits, logits_agree, oracle_cfg  # noqa: E402


def main():
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
        if rank == 0:
            print(f"fused allreduce+add+norm T={T}: hidden exact={e_h} max out err={err.max():.2f} ulp")
    # ---- (2) model parity: TP-N on GPUs vs TP-N oracle ----
    cfg = QWEN3_SMALL
    w = random_weights(cfg, seed=0, norm_jitter=0.1)
    prompt = [t % cfg.vocab_size for t in synthetic_prompt(40)]
    comm2 = make_tp_comm(rank, world, dist, max_tokens=256, hidden=cfg.hidden_size)
    rt = ModelRuntimeConfig(tensor_parallel=TensorParallelConfig(rank, world), device_ordinal=local, num_pages=32,
                            max_batch=1)
    m = Qwen3Model(cfg, {k: v.cuda() for k, v in w.items()}, rt, tp_comm=comm2)
    kv = m.alloc_kv()
    got = [bits(m.prefill([prompt], [kv])[0])]
    if rank == 0:
        orc = O.OracleQwen3(oracle_cfg(cfg), to_numpy_bits(w), tp_world=world, num_pages=32)
        okv = orc.alloc_kv()
        want = [orc.prefill([prompt], [okv])[0]]
    toks = torch.zeros(6, dtype=torch.int64, device="cuda")
    if rank == 0:
        for i in range(6):
            toks[i] = O.argmax(want[-1])
            want.append(orc.decode([int(toks[i])], [okv])[0])
    dist.broadcast(toks, 0)
    for i in range(6):
        lg, _ = m.decode([int(toks[i])], [kv])
        got.append(bits(lg[0]))
    if rank == 0:
        for i, (g_, w_) in enumerate(zip(got, want)):
            good, info = logits_agree(g_, w_, 6)
            ok &= good
            print(f"TP{world} step {i}: ok={good} {info}")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("TP_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL")
    m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
