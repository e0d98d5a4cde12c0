"""CPU restatement of the reference's Qwen3.5 hybrid forward (BASELINE config 4).

TEST INFRASTRUCTURE ONLY, groundwork for the next scope row: nothing under pegainfer_b200/ may import it.  It restates,
with the reference's rounding points, the per-token recurrent form of every Qwen3.5 op; the reference's prefill uses
chunk-wise kernels for the gated delta rule (Triton AOT, 7 kernels) that are algebraically the same recurrence.
Pinned against HF transformers' `Qwen3_5ForCausalLM` (the reference's declared external truth,
scripts/generate_test_data.py) by tests/test_oracle_qwen35.py: tiny random-init hybrid model, logits within a few bf16
ulps of the row maximum, arg-max identical.

Reference rules followed (paths relative to /root/reference):
  * (1+w) RMSNorm, one rounding             pegainfer-kernels/csrc/flashinfer_norm.cu:108-133 (FlashInfer GemmaRMSNorm)
  * full attention: q_proj rows per head = [q(hd) | gate(hd)], per-head (1+w) norm -> bf16, partial NeoX RoPE on the
    first rotary_dim dims with bf16 cos/sin, gate = sigmoid in fp32, one rounding each
                                             csrc/prefill_attention_hd256.cu:7-113,134-157
  * causal conv1d (k = 4) -> bf16 -> SiLU -> bf16                                  csrc/conv1d.cu:19-63
  * gated delta rule, fp32 state [key, val], L2-normalised q/k (eps 1e-12), q * rsqrt(dk), g = -exp(A_log) *
    softplus(a + dt_bias), beta = sigmoid(b), output rounded to bf16              csrc/gated_delta_rule.cu:27-166
  * gated per-head RMSNorm: x * rsqrt(mean + eps) * w(f32) * silu(z), one rounding csrc/norm.cu:17-61
  * MLP: separate gate / up GEMMs, SiLU rounded to bf16 BEFORE the multiply        csrc/elementwise.cu:26-41
  * residual adds round to bf16; tied lm_head                                      pegainfer-qwen35-4b/src/prefill.rs:112-188
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from oracle.qwen3_oracle import bf16_to_f32, f32_to_bf16, precompute_rope


def rb(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32."""
    return bf16_to_f32(f32_to_bf16(np.asarray(x, dtype=np.float32)))


@dataclass
class Qwen35Config:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    vocab_size: int
    linear_num_key_heads: int
    linear_num_value_heads: int
    linear_key_head_dim: int
    linear_value_head_dim: int
    linear_conv_kernel_dim: int
    layer_types: list = field(default_factory=list)  # "full_attention" | "linear_attention"
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e7
    partial_rotary_factor: float = 0.25

    @property
    def rotary_dim(self) -> int:  # config.rs:99
        return int(self.head_dim * self.partial_rotary_factor)


def _f(w):  # bf16 bit pattern (uint16) or fp32 array -> fp32
    w = np.asarray(w)
    return bf16_to_f32(w) if w.dtype == np.uint16 else w.astype(np.float32)


def gemm(W, X):
    """Y[T, M] = X[T, K] W[M, K]^T, bf16 operands (already fp32-valued), fp32 accumulate, one bf16 rounding."""
    return rb(X.astype(np.float32) @ W.astype(np.float32).T)


def rms_norm_offset(x, w, eps):
    inv = 1.0 / np.sqrt((x * x).mean(axis=-1, keepdims=True, dtype=np.float32) + np.float32(eps))
    return rb(x * inv * (1.0 + w))


def silu(x):
    return x / (1.0 + np.exp(-x, dtype=np.float32))


def gated_delta_rule_step(q, k, v, a, b, dt_bias, a_log, S):
    """One recurrent step for all value heads (csrc/gated_delta_rule.cu:27-166), fp32 throughout; updates the state
    S[nv, dk, dv] in place and returns the un-rounded outputs [nv, dv].  q, k: [nk, dk]; v: [nv, dv]; a, b: [nv]."""
    nk, dk = q.shape
    nv = v.shape[0]
    out = np.zeros_like(v, dtype=np.float32)
    for h in range(nv):
        kh = h * nk // nv
        qn = q[kh] * np.float32(1.0 / np.sqrt((q[kh] * q[kh]).sum(dtype=np.float32) + np.float32(1e-12)))
        kn = k[kh] * np.float32(1.0 / np.sqrt((k[kh] * k[kh]).sum(dtype=np.float32) + np.float32(1e-12)))
        qn = qn * np.float32(1.0 / np.sqrt(dk))
        x = np.float32(a[h] + dt_bias[h])
        softplus = x if x > 20.0 else np.log1p(np.exp(x, dtype=np.float32), dtype=np.float32)
        g = -np.exp(a_log[h], dtype=np.float32) * softplus
        beta = np.float32(1.0 / (1.0 + np.exp(-b[h], dtype=np.float32)))
        Sh = S[h] * np.exp(g, dtype=np.float32)  # decay, [dk, dv]
        delta = (v[h] - kn @ Sh) * beta          # delta rule: correct the memory's prediction of v
        Sh = Sh + np.outer(kn, delta)
        S[h] = Sh
        out[h] = qn @ Sh
    return out


class _LinearState:
    def __init__(self, cfg: Qwen35Config):
        c = cfg
        self.conv = np.zeros((2 * c.linear_num_key_heads * c.linear_key_head_dim + c.linear_num_value_heads * c.linear_value_head_dim,
                              c.linear_conv_kernel_dim - 1), np.float32)  # bf16-valued history, oldest first
        self.S = np.zeros((c.linear_num_value_heads, c.linear_key_head_dim, c.linear_value_head_dim), np.float32)


class OracleQwen35:
    """Single-request model: `prefill(tokens)` then `decode(token)`; both return the last token's logits as fp32
    values that are exactly representable in bf16."""

    def __init__(self, cfg: Qwen35Config, weights: dict):
        self.c = cfg
        self.w = {k: _f(v) for k, v in weights.items()}
        cos, sin = precompute_rope(cfg.rotary_dim, 4096, cfg.rope_theta)
        self.cos = bf16_to_f32(cos).reshape(4096, cfg.rotary_dim)
        self.sin = bf16_to_f32(sin).reshape(4096, cfg.rotary_dim)
        self.pos = 0
        self.kcache = {i: [] for i, t in enumerate(cfg.layer_types) if t == "full_attention"}
        self.vcache = {i: [] for i in self.kcache}
        self.lin = {i: _LinearState(cfg) for i, t in enumerate(cfg.layer_types) if t == "linear_attention"}

    # ---------------------------------------------------------------- full attention (one token)
    def _full_attention(self, li, x, prefill):
        c, w, p = self.c, self.w, f"model.layers.{li}.self_attn."
        hd, nq, nkv, rd = c.head_dim, c.num_attention_heads, c.num_key_value_heads, c.rotary_dim
        qf = gemm(w[p + "q_proj.weight"], x[None])[0].reshape(nq, 2, hd)
        q, gate = qf[:, 0], qf[:, 1]
        k = gemm(w[p + "k_proj.weight"], x[None])[0].reshape(nkv, hd)
        v = gemm(w[p + "v_proj.weight"], x[None])[0].reshape(nkv, hd)

        def norm_rope(h, nw):
            inv = 1.0 / np.sqrt((h * h).mean(axis=-1, keepdims=True, dtype=np.float32) + np.float32(c.rms_norm_eps))
            n = rb(h * inv * (1.0 + nw))
            lo, hi = n[:, : rd // 2].copy(), n[:, rd // 2: rd].copy()
            cs, sn = self.cos[self.pos, : rd // 2], self.sin[self.pos, : rd // 2]
            n[:, : rd // 2] = rb(lo * cs - hi * sn)
            n[:, rd // 2: rd] = rb(lo * sn + hi * cs)
            return n

        q = norm_rope(q, w[p + "q_norm.weight"])
        k = norm_rope(k, w[p + "k_norm.weight"])
        self.kcache[li].append(k)
        self.vcache[li].append(v)
        K = np.stack(self.kcache[li])  # [ctx, nkv, hd]
        V = np.stack(self.vcache[li])
        grp = nq // nkv
        out = np.zeros((nq, hd), np.float32)
        scale = np.float32(1.0 / np.sqrt(hd))
        for h in range(nq):
            s = (K[:, h // grp] @ q[h]).astype(np.float32) * scale
            pr = np.exp(s - s.max(), dtype=np.float32)
            if prefill:  # FA2 prefill rounds P to bf16 and sums the rounded P; the decode kernel keeps fp32 (a6 / a8)
                pr = rb(pr)
            out[h] = rb((pr @ V[:, h // grp]) / pr.sum(dtype=np.float32))
        out = rb(out * (1.0 / (1.0 + np.exp(-gate, dtype=np.float32))))
        return gemm(w[p + "o_proj.weight"], out.reshape(1, -1))[0]

    # ---------------------------------------------------------------- gated delta net (one token)
    def _linear_attention(self, li, x):
        c, w, p, st = self.c, self.w, f"model.layers.{li}.linear_attn.", self.lin[li]
        nk, nv, dk, dv = c.linear_num_key_heads, c.linear_num_value_heads, c.linear_key_head_dim, c.linear_value_head_dim
        qkv = gemm(w[p + "in_proj_qkv.weight"], x[None])[0]
        z = gemm(w[p + "in_proj_z.weight"], x[None])[0].reshape(nv, dv)
        b = gemm(w[p + "in_proj_b.weight"], x[None])[0]
        a = gemm(w[p + "in_proj_a.weight"], x[None])[0]
        # causal depthwise conv over [history | current], fp32 sum -> bf16 -> SiLU -> bf16
        cw = w[p + "conv1d.weight"].reshape(qkv.shape[0], c.linear_conv_kernel_dim)
        window = np.concatenate([st.conv, qkv[:, None]], axis=1)
        acc = np.zeros(qkv.shape[0], np.float32)
        for kk in range(c.linear_conv_kernel_dim):
            acc = acc + window[:, kk] * cw[:, kk]
        conv = rb(silu(rb(acc)))
        st.conv = window[:, 1:]
        q = conv[: nk * dk].reshape(nk, dk)
        k = conv[nk * dk: 2 * nk * dk].reshape(nk, dk)
        v = conv[2 * nk * dk:].reshape(nv, dv)
        out = rb(gated_delta_rule_step(q, k, v, a, b, w[p + "dt_bias"], w[p + "A_log"], st.S))
        inv = 1.0 / np.sqrt((out * out).mean(axis=-1, keepdims=True, dtype=np.float32) + np.float32(c.rms_norm_eps))
        normed = rb(out * inv * w[p + "norm.weight"][None, :] * silu(z))
        return gemm(w[p + "out_proj.weight"], normed.reshape(1, -1))[0]

    # ---------------------------------------------------------------- one token through the stack
    def _token(self, tok, want_logits, prefill):
        c, w = self.c, self.w
        h = w["model.embed_tokens.weight"][tok].copy()
        for li, kind in enumerate(c.layer_types):
            p = f"model.layers.{li}."
            x = rms_norm_offset(h, w[p + "input_layernorm.weight"], c.rms_norm_eps)
            attn = self._full_attention(li, x, prefill) if kind == "full_attention" else self._linear_attention(li, x)
            h = rb(h + attn)
            x = rms_norm_offset(h, w[p + "post_attention_layernorm.weight"], c.rms_norm_eps)
            gate = gemm(w[p + "mlp.gate_proj.weight"], x[None])[0]
            up = gemm(w[p + "mlp.up_proj.weight"], x[None])[0]
            act = rb(rb(silu(gate)) * up)
            h = rb(h + gemm(w[p + "mlp.down_proj.weight"], act[None])[0])
        self.pos += 1
        if not want_logits:
            return None
        x = rms_norm_offset(h, w["model.norm.weight"], c.rms_norm_eps)
        return gemm(w["model.embed_tokens.weight"], x[None])[0]

    def prefill(self, tokens):
        out = None
        for i, t in enumerate(tokens):
            out = self._token(int(t), i == len(tokens) - 1, True)
        return out

    def decode(self, token):
        return self._token(int(token), True, False)
