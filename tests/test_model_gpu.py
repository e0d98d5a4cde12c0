"""Model-level parity on the GPU: the C++ host mirror driving the sm_100a kernels through the C ABI,
teacher-forced against the CPU oracle on the same random-init weights and synthetic token ids.

Parity rule: at every step feed the ORACLE's token; require every logit within
`TOL_ULP` bf16 ulps of the row's max magnitude and arg-max equality unless the oracle's top-1/top-2
gap is inside that tolerance.  Runs the B200 fused path, the reference op sequence on our kernels
(`fused=False`) and -- when oracle/_ref exists -- the reference's own kernels under the same host."""
import os

import numpy as np
import pytest
import torch

from oracle import qwen3_oracle as O
from pegainfer_b200.config import QWEN3_SMALL, QWEN3_TINY, TensorParallelConfig
from pegainfer_b200.model import ModelRuntimeConfig, Qwen3Model
from pegainfer_b200.synthetic import random_weights, synthetic_prompt, to_numpy_bits
from tests.helpers import bits, f32, logits_agree, oracle_cfg

pytestmark = pytest.mark.gpu
REF_LIB = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libkernels_ref.so"))
TOL_ULP = 6  # logits, in bf16 ulps at the row's max |logit| (tiny/small configs, <= 4 layers)

VARIANTS = [("fused", dict(fused=True)), ("fused-nograph", dict(fused=True, enable_cuda_graph=False)),
            ("compat", dict(fused=False)),
            ("compat-nograph", dict(fused=False, enable_cuda_graph=False))]
if os.path.exists(REF_LIB):
    VARIANTS.append(("refkernels", dict(fused=False, kernel_lib=REF_LIB)))


def run_teacher_forced(cfg, variant_kw, prompt_len, n_decode, tp_world=1, num_pages=64):
    w = random_weights(cfg, seed=0, norm_jitter=0.1)
    orc = O.OracleQwen3(oracle_cfg(cfg), to_numpy_bits(w), tp_world=1, num_pages=num_pages)
    okv = orc.alloc_kv()
    prompt = [t % cfg.vocab_size for t in synthetic_prompt(prompt_len)]
    want = [orc.prefill([prompt], [okv])[0]]
    rt = ModelRuntimeConfig(num_pages=num_pages, **variant_kw)
    m = Qwen3Model(cfg, {k: v.cuda() for k, v in w.items()}, rt)
    kv = m.alloc_kv()
    got = [bits(m.prefill([prompt], [kv])[0])]
    toks = []
    for _ in range(n_decode):
        tok = O.argmax(want[-1])  # teacher forcing with the oracle's token
        toks.append(tok)
        want.append(orc.decode([tok], [okv])[0])
        lg, _ = m.decode([tok], [kv])
        got.append(bits(lg[0]))
    m.drop_request(kv)
    m.close()
    return got, want, orc


@pytest.mark.parametrize("name,kw", VARIANTS)
def test_tiny_prefill_decode_parity(name, kw):
    got, want, _ = run_teacher_forced(QWEN3_TINY, kw, prompt_len=24, n_decode=10)
    for step, (g, w) in enumerate(zip(got, want)):
        ok, info = logits_agree(g, w, TOL_ULP)
        assert ok, f"{name} step {step}: {info}"


@pytest.mark.parametrize("name,kw", [VARIANTS[0], VARIANTS[3]])
def test_small_config_parity(name, kw):
    """4 layers, GQA group 2, untied lm_head, prompt crossing several pages."""
    got, want, _ = run_teacher_forced(QWEN3_SMALL, kw, prompt_len=70, n_decode=6)
    for step, (g, w) in enumerate(zip(got, want)):
        ok, info = logits_agree(g, w, TOL_ULP)
        assert ok, f"{name} step {step}: {info}"


@pytest.mark.parametrize("name,kw", [VARIANTS[0], VARIANTS[1], VARIANTS[3]])
def test_long_context_split_kv_path(name, kw):
    """ctx >= 1024 at bs=1 takes the split-KV path in the reference (batch_decode_buffers.rs:281-287)."""
    got, want, orc = run_teacher_forced(QWEN3_TINY, kw, prompt_len=1030, n_decode=3, num_pages=128)
    assert orc.last_attention_path == "split_kv"
    for step, (g, w) in enumerate(zip(got, want)):
        ok, info = logits_agree(g, w, TOL_ULP)
        assert ok, f"{name} step {step}: {info}"


def test_batch_matches_sequential():
    """pegainfer-qwen3-4b/src/batch_decode.rs:505-606: batch prefill/decode == per-request, exact
    greedy-token equality, with CUDA Graph."""
    cfg = QWEN3_TINY
    w = {k: v.cuda() for k, v in random_weights(cfg, seed=0, norm_jitter=0.1).items()}
    prompts = [[t % cfg.vocab_size for t in synthetic_prompt(n, off)] for n, off in ((9, 0), (21, 50))]
    m = Qwen3Model(cfg, w, ModelRuntimeConfig(num_pages=64, fused=False))
    seq_tokens = []
    for p in prompts:
        kv = m.alloc_kv()
        lg = m.prefill([p], [kv])
        toks = [m.sample_greedy(lg[0])]
        for _ in range(6):
            _, s = m.decode([toks[-1]], [kv], want_logits=False)
            toks.append(s[0])
        seq_tokens.append(toks)
        m.drop_request(kv)
    kvs = [m.alloc_kv(), m.alloc_kv()]
    lg = m.prefill(prompts, kvs)
    cur = [m.sample_greedy(lg[i]) for i in range(2)]
    batch_tokens = [[cur[0]], [cur[1]]]
    for _ in range(6):
        _, cur = m.decode(cur, kvs, want_logits=False)
        batch_tokens[0].append(cur[0]); batch_tokens[1].append(cur[1])
    assert batch_tokens == seq_tokens
    m.close()


def test_determinism_and_page_lifecycle():
    """tests/paged_attention.rs:100-126 (same greedy tokens twice) + kv_pool.rs page return on drop."""
    cfg = QWEN3_TINY
    w = {k: v.cuda() for k, v in random_weights(cfg, seed=0).items()}
    m = Qwen3Model(cfg, w, ModelRuntimeConfig(num_pages=32))
    free0 = m.available_pages()
    prompt = [t % cfg.vocab_size for t in synthetic_prompt(40)]
    a, _, _ = m.generate(prompt, 10)
    assert m.available_pages() == free0
    b, _, _ = m.generate(prompt, 10)
    assert a == b
    kv = m.alloc_kv()
    with pytest.raises(RuntimeError, match="out of pages"):
        m.prefill([[1] * (16 * 40)], [kv])
    m.close()


def test_generate_matches_oracle_free_running():
    """Free-running greedy sequence against the oracle (first divergence index reported)."""
    cfg = QWEN3_TINY
    w = random_weights(cfg, seed=0, norm_jitter=0.1)
    orc = O.OracleQwen3(oracle_cfg(cfg), to_numpy_bits(w), num_pages=32)
    okv = orc.alloc_kv()
    prompt = [t % cfg.vocab_size for t in synthetic_prompt(16)]
    want = [O.argmax(orc.prefill([prompt], [okv])[0])]
    for _ in range(11):
        want.append(O.argmax(orc.decode([want[-1]], [okv])[0]))
    m = Qwen3Model(cfg, {k: v.cuda() for k, v in w.items()}, ModelRuntimeConfig(num_pages=32))
    got, ttft, steps = m.generate(prompt, 12)
    m.close()
    div = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), None)
    assert div is None or div >= 4, f"free-running sequences diverge at {div}: {got} vs {want}"
    assert ttft > 0 and len(steps) == 11


@pytest.mark.parametrize("name,kw", [VARIANTS[0], VARIANTS[3]])
def test_decode_buckets_above_four(name, kw):
    """The reference's decode buckets 8..64 (batch_decode_buffers.rs:12): 6 requests of different lengths -> bucket 8.
    Fused: tensor-core skinny GEMMs + one fused attention launch over all requests + in-graph top-1 per row."""
    cfg = QWEN3_TINY
    w = random_weights(cfg, seed=0, norm_jitter=0.1)
    orc = O.OracleQwen3(oracle_cfg(cfg), to_numpy_bits(w), num_pages=96)
    m = Qwen3Model(cfg, {k: v.cuda() for k, v in w.items()}, ModelRuntimeConfig(num_pages=96, max_batch=8, **kw))
    prompts = [[t % cfg.vocab_size for t in synthetic_prompt(n, off)] for n, off in ((5, 0), (17, 30), (33, 7), (16, 90), (70, 11), (2, 400))]
    okvs, kvs = [orc.alloc_kv() for _ in prompts], [m.alloc_kv() for _ in prompts]
    want = orc.prefill(prompts, okvs)
    got = m.prefill(prompts, kvs)
    for i in range(len(prompts)):
        ok, info = logits_agree(bits(got[i]), want[i], TOL_ULP)
        assert ok, f"{name} batched prefill row {i}: {info}"
    for step in range(4):
        toks = [O.argmax(r) for r in want]
        want = list(orc.decode(toks, okvs))
        lg, sampled = m.decode(toks, kvs)
        for i in range(len(prompts)):
            ok, info = logits_agree(bits(lg[i]), want[i], TOL_ULP)
            assert ok, f"{name} step {step} row {i}: {info}"
            row = lg[i].float()
            assert float(row[sampled[i]]) == float(row.max()), "in-graph top-1 is not an arg-max of its row"
    m.close()


@pytest.mark.parametrize("name,kw", [VARIANTS[0], VARIANTS[3]] + ([VARIANTS[-1]] if VARIANTS[-1][0] == "refkernels" else []))
def test_unified_step_prefill_plus_decode(name, kw):
    """unified_forward.rs:78-567: two requests mid-decode and two new prompts in ONE forward pass, against the oracle's
    separate prefill + decode on the same states."""
    cfg = QWEN3_TINY
    w = random_weights(cfg, seed=0, norm_jitter=0.1)
    orc = O.OracleQwen3(oracle_cfg(cfg), to_numpy_bits(w), num_pages=96)
    m = Qwen3Model(cfg, {k: v.cuda() for k, v in w.items()}, ModelRuntimeConfig(num_pages=96, max_batch=8, **kw))
    old = [[t % cfg.vocab_size for t in synthetic_prompt(n, off)] for n, off in ((21, 0), (40, 55))]
    new = [[t % cfg.vocab_size for t in synthetic_prompt(n, off)] for n, off in ((9, 300), (35, 77))]
    okv_old, kv_old = [orc.alloc_kv() for _ in old], [m.alloc_kv() for _ in old]
    w0 = orc.prefill(old, okv_old)
    m.prefill(old, kv_old)
    dec_toks = [O.argmax(r) for r in w0]
    okv_new, kv_new = [orc.alloc_kv() for _ in new], [m.alloc_kv() for _ in new]
    want_p = orc.prefill(new, okv_new)
    want_d = orc.decode(dec_toks, okv_old)
    got_p, got_d = m.unified_step(new, kv_new, dec_toks, kv_old)
    for i in range(2):
        ok, info = logits_agree(bits(got_p[i]), want_p[i], TOL_ULP)
        assert ok, f"{name} unified prefill row {i}: {info}"
        ok, info = logits_agree(bits(got_d[i]), want_d[i], TOL_ULP)
        assert ok, f"{name} unified decode row {i}: {info}"
    assert [m.kv_seq_len(k) for k in kv_old] == [22, 41] and [m.kv_seq_len(k) for k in kv_new] == [9, 35]
    # the states keep working through the ordinary entry points: one more decode step over all four requests
    toks = [O.argmax(r) for r in list(want_d) + list(want_p)]
    want = orc.decode(toks, okv_old + okv_new)
    lg, _ = m.decode(toks, kv_old + kv_new)
    for i in range(4):
        ok, info = logits_agree(bits(lg[i]), want[i], TOL_ULP)
        assert ok, f"{name} decode after the unified step, row {i}: {info}"
    # a request cannot be on both sides of one step, and a rejected step leaves every state untouched
    with pytest.raises(RuntimeError, match="same step"):
        m.unified_step([[1, 2, 3]], [kv_old[0]], [5], [kv_old[0]])
    assert m.kv_seq_len(kv_old[0]) == 23
    m.close()
