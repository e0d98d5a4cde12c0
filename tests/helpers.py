"""Shared test helpers (tests only): oracle config bridge, bf16 views, tolerances."""
import numpy as np
import torch

from oracle import qwen3_oracle as O


def oracle_cfg(c):
    return O.OracleConfig(c.hidden_size, c.intermediate_size, c.num_hidden_layers,
                          c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.vocab_size,
                          c.rms_norm_eps, c.rope_theta, c.tie_word_embeddings)


def bits(t: torch.Tensor) -> np.ndarray:
    """bf16 torch tensor (any device) -> numpy uint16 bits."""
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def from_bits(a: np.ndarray, device="cpu") -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(torch.bfloat16).to(device)


def bf16_bits_of(xs) -> np.ndarray:
    return O.f32_to_bf16(np.asarray(xs, dtype=np.float32))


def f32(a: np.ndarray) -> np.ndarray:
    return O.bf16_to_f32(a)


def ulp_err(got_bits: np.ndarray, want_bits: np.ndarray, floor: float = 0.0) -> np.ndarray:
    """|got - want| in units of bf16 ulp(max(|want|, floor))."""
    g, w = f32(got_bits), f32(want_bits)
    return np.abs(g - w) / O.bf16_ulp(np.maximum(np.abs(w), floor))


def assert_bf16_close(got_bits, want_bits, max_ulp, floor=0.0, frac_exact=None, what=""):
    e = ulp_err(np.asarray(got_bits).ravel(), np.asarray(want_bits).ravel(), floor)
    assert np.isfinite(e).all(), f"{what}: non-finite"
    assert e.max() <= max_ulp, f"{what}: max err {e.max():.2f} ulp > {max_ulp} (at {e.argmax()})"
    if frac_exact is not None:
        fe = float((np.asarray(got_bits).ravel() == np.asarray(want_bits).ravel()).mean())
        assert fe >= frac_exact, f"{what}: only {fe:.4f} bit-exact < {frac_exact}"


def logits_agree(got_bits, want_bits, max_ulp_rowmax):
    """SURVEY 8c parity rule: |dlogit| <= N ulp at the row's max magnitude, and arg-max
    equality unless the oracle's top-1/top-2 gap is within that tolerance."""
    g, w = f32(got_bits), f32(want_bits)
    tol = max_ulp_rowmax * float(O.bf16_ulp(np.abs(w).max()))
    err = float(np.abs(g - w).max())
    top = np.sort(w)[::-1]
    margin = float(top[0] - top[1])
    same = int(g.argmax()) == int(w.argmax())
    ok = err <= tol and (same or margin <= 2 * tol)
    return ok, dict(err=err, tol=tol, margin=margin, same_argmax=same)
