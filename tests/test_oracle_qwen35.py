"""Pins oracle/qwen35_oracle.py (groundwork for SURVEY 8(f)-1, the Qwen3.5 hybrid layers) to HF transformers through the
committed fixture tests/golden/hf_qwen35_tiny.npz (generator: tests/golden/make_hf_qwen35_fixtures.py)."""
import os

import numpy as np

from oracle import qwen3_oracle as O
from oracle.qwen35_oracle import OracleQwen35, Qwen35Config, gated_delta_rule_step

FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "hf_qwen35_tiny.npz"))


def _model():
    kv = dict(zip(FX["cfg_keys"].tolist(), FX["cfg_vals"].tolist()))
    cfg = Qwen35Config(kv["hidden_size"], kv["intermediate_size"], kv["num_hidden_layers"], kv["num_attention_heads"],
                       kv["num_key_value_heads"], kv["head_dim"], kv["vocab_size"], kv["linear_num_key_heads"],
                       kv["linear_num_value_heads"], kv["linear_key_head_dim"], kv["linear_value_head_dim"],
                       kv["linear_conv_kernel_dim"], FX["layer_types"].tolist(), 1e-6, float(FX["theta"]),
                       float(FX["partial_rotary_factor"]))
    w = {k[2:]: FX[k] for k in FX.files if k.startswith("w:")}
    return OracleQwen35(cfg, w)


def _err_ulp_rowmax(a, b):
    return float((np.abs(a - b) / O.bf16_ulp(np.full_like(b, np.abs(b).max()))).max())


def test_hybrid_model_matches_hf_prefill_and_decode():
    toks = FX["tokens"].tolist()
    hf16, hf32 = FX["logits_bf16"], FX["logits_f32"]
    m = _model()
    got = {9: m.prefill(toks[:10])}
    for i in range(10, len(toks)):
        got[i] = m.decode(toks[i])
    for i, lg in got.items():
        assert lg.argmax() == hf16[i].argmax() == hf32[i].argmax(), i
        # bf16 noise floor of the model itself: HF-bf16 vs HF-fp32 on this fixture is 2.6 - 6.5 ulp of the row maximum
        assert _err_ulp_rowmax(lg, hf16[i]) <= 10, (i, _err_ulp_rowmax(lg, hf16[i]))
        assert _err_ulp_rowmax(lg, hf32[i]) <= 10, (i, _err_ulp_rowmax(lg, hf32[i]))


def test_prefill_then_decode_equals_longer_prefill_within_p_rounding():
    """The recurrent state hand-over (conv window, delta-rule state, KV rows) is consistent: decoding token 10 after a
    10-token prefill gives the logits of an 11-token prefill up to the bf16 rounding of P in the prefill attention."""
    toks = FX["tokens"].tolist()
    a, b = _model(), _model()
    a.prefill(toks[:10])
    la = a.decode(toks[10])
    lb = b.prefill(toks[:11])
    assert la.argmax() == lb.argmax() and _err_ulp_rowmax(la, lb) <= 2


def test_gated_delta_rule_step_matches_hf_recurrence():
    q, k, v, a, b = FX["gdr_q"], FX["gdr_k"], FX["gdr_v"], FX["gdr_a"], FX["gdr_b"]
    S = np.zeros((v.shape[1], q.shape[2], v.shape[2]), np.float32)
    for t in range(q.shape[0]):
        out = gated_delta_rule_step(q[t], k[t], v[t], a[t], b[t], FX["gdr_dt_bias"], FX["gdr_a_log"], S)
        # HF normalises with eps 1e-6 (reference kernel: 1e-12) -> identical to ~1e-6 relative
        np.testing.assert_allclose(out, FX["gdr_out"][t], rtol=2e-4, atol=2e-5)


def test_chunkwise_form_equals_the_recurrence():
    """oracle/qwen35_chunkwise.py (the matrix form the B200 prefill kernel will implement) against the per-token rule."""
    from oracle.qwen35_chunkwise import gated_delta_rule_chunkwise
    rng = np.random.RandomState(3)
    T, nk, nv, dk, dv = 150, 2, 4, 32, 16
    q, k = rng.randn(T, nk, dk).astype(np.float32), rng.randn(T, nk, dk).astype(np.float32)
    v, a, b = rng.randn(T, nv, dv).astype(np.float32), rng.randn(T, nv).astype(np.float32), rng.randn(T, nv).astype(np.float32)
    dt_bias, a_log = (rng.randn(nv) * 0.5).astype(np.float32), (rng.randn(nv) * 0.5).astype(np.float32)
    S = (rng.randn(nv, dk, dv) * 0.1).astype(np.float32)
    S_rec = S.copy()
    want = np.stack([gated_delta_rule_step(q[t], k[t], v[t], a[t], b[t], dt_bias, a_log, S_rec) for t in range(T)])
    for h in range(nv):
        kh = h * nk // nv
        qn = q[:, kh] / np.sqrt((q[:, kh] ** 2).sum(-1, keepdims=True) + 1e-12) / np.sqrt(dk)
        kn = k[:, kh] / np.sqrt((k[:, kh] ** 2).sum(-1, keepdims=True) + 1e-12)
        x = a[:, h] + dt_bias[h]
        g = -np.exp(a_log[h]) * np.where(x > 20, x, np.log1p(np.exp(x)))
        beta = 1 / (1 + np.exp(-b[:, h]))
        for chunk in (64, 37):
            O, SC = gated_delta_rule_chunkwise(qn, kn, v[:, h], g, beta, S[h], chunk=chunk)
            np.testing.assert_allclose(O, want[:, h], rtol=2e-4, atol=2e-5)
            np.testing.assert_allclose(SC, S_rec[h], rtol=2e-4, atol=2e-5)
