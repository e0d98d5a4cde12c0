"""Tensor-parallel host logic on CPU: shard plan (config.rs:150-153, weights.rs:121-291) and the
N>1 exchange step, with a real world_size-2 `gloo` process group (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import qwen3_oracle as O
from pegainfer_b200.config import QWEN3_4B, QWEN3_8B, QWEN3_TINY, Qwen3Config, TensorParallelConfig
from pegainfer_b200.synthetic import random_weights, synthetic_prompt, to_numpy_bits
from tests.helpers import oracle_cfg


def test_shard_ranges_tile_the_matrices():
    for cfg in (QWEN3_4B, QWEN3_8B):
        for world in (1, 2, 4, 8):
            covered_q, covered_i = [], []
            for rank in range(world):
                tp = TensorParallelConfig(rank, world)
                tp.validate_for(cfg)
                covered_q.append(tp.shard_range(cfg.q_dim))
                covered_i.append(tp.shard_range(cfg.intermediate_size))
            assert [o for o, _ in covered_q] == [r * cfg.q_dim // world for r in range(world)]
            assert sum(n for _, n in covered_q) == cfg.q_dim
            assert sum(n for _, n in covered_i) == cfg.intermediate_size


def test_validate_rejects_bad_worlds():
    with pytest.raises(ValueError, match="not divisible"):
        TensorParallelConfig(0, 3).validate_for(QWEN3_4B)
    with pytest.raises(ValueError, match="must be <"):
        TensorParallelConfig(2, 2).validate_for(QWEN3_4B)
    bad = Qwen3Config(256, 510, 2, 8, 4, 128, 1024)
    with pytest.raises(ValueError, match="intermediate_size"):
        TensorParallelConfig(0, 4).validate_for(bad)


def _rank_forward(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = QWEN3_TINY
    w = to_numpy_bits(random_weights(cfg, seed=0, norm_jitter=0.1))
    oc = oracle_cfg(cfg)
    rk = O._Rank(oc, w, rank, world, num_pages=8)  # this rank's weight shard + KV pool only
    cos, sin = O.precompute_rope(cfg.head_dim, oc.max_position, cfg.rope_theta)
    prompt = [t % cfg.vocab_size for t in synthetic_prompt(8)]
    T, hd = len(prompt), cfg.head_dim

    def all_reduce(part_bits):  # bf16 partial -> fp32 SUM over ranks -> one bf16 rounding
        t = torch.from_numpy(O.bf16_to_f32(part_bits).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return O.f32_to_bf16(t.numpy())

    hidden = O.embedding_batched(w["model.embed_tokens.weight"], np.array(prompt, np.uint32), cfg.hidden_size)
    pi, ip, lpl = np.array([1], np.int32), np.array([0, 1], np.int32), np.array([T], np.int32)
    pos, bidx = np.arange(T, dtype=np.int32), np.zeros(T, np.int32)
    q_indptr = np.array([0, T], np.int32)
    for li in range(cfg.num_hidden_layers):
        L = rk.layers[li]
        normed = O.rms_norm(hidden, L["input_ln"], cfg.rms_norm_eps)
        qd, kd = rk.nq * hd, rk.nkv * hd
        q, k, v = O.gemm(L["qkv"][:qd], normed), O.gemm(L["qkv"][qd:qd + kd], normed), O.gemm(L["qkv"][qd + kd:], normed)
        O.qk_norm_rope(q, k, L["q_norm"], L["k_norm"], cos, sin, rk.nq, rk.nkv, hd, cfg.rms_norm_eps, positions=pos)
        k_off = li * rk.layer_stride
        O.paged_kv_scatter(rk.kv, k_off, k_off + rk.kv_block_len, pi, ip, lpl, k, v, bidx, pos, rk.nkv, hd, 16, rk.page_stride)
        attn = O.batch_prefill_paged(q, rk.kv, k_off, k_off + rk.kv_block_len, pi, ip, lpl, q_indptr, rk.nq, rk.nkv, hd,
                                     16, rk.page_stride, 1 / np.sqrt(hd))
        o = all_reduce(O.gemm(L["o"], attn))
        normed = O.fused_add_rms_norm(hidden, o, L["post_ln"], cfg.rms_norm_eps)
        mlp = all_reduce(O.gemm(L["down"], O.silu_mul_fused(O.gemm(L["gate_up"], normed), rk.inter)))
        hidden = O.add(hidden, mlp)
    last = np.ascontiguousarray(hidden[-1:])
    logits = O.gemm(w["model.embed_tokens.weight"], O.rms_norm(last, w["model.norm.weight"], cfg.rms_norm_eps))[0]
    if rank == 0:
        np.save(out_path, logits)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_forward_matches_single_process_tp(tmp_path):
    """Each gloo rank holds only its shard; the exchanged result equals the in-process TP2 oracle bit for
    bit and stays within tolerance of TP1."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "logits.npy")
    mp.spawn(_rank_forward, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    cfg = QWEN3_TINY
    w = to_numpy_bits(random_weights(cfg, seed=0, norm_jitter=0.1))
    prompt = [t % cfg.vocab_size for t in synthetic_prompt(8)]
    ref2 = O.OracleQwen3(oracle_cfg(cfg), w, tp_world=2, num_pages=8)
    want2 = ref2.prefill([prompt], [ref2.alloc_kv()])[0]
    assert (got == want2).all()
    ref1 = O.OracleQwen3(oracle_cfg(cfg), w, tp_world=1, num_pages=8)
    want1 = O.bf16_to_f32(ref1.prefill([prompt], [ref1.alloc_kv()])[0])
    assert np.abs(O.bf16_to_f32(got) - want1).max() <= 4 * O.bf16_ulp(np.abs(want1).max())
