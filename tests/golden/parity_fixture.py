"""Compact oracle fixtures for the full-size configurations (numpy only -- no oracle import, so bench.py's
`parity` key and the GPU tests can read them without executing anything under oracle/).

A fixture is the CPU oracle's output for one (model, tp_world, prompt_len) run on the seed-0 CPU-generated
random-init checkpoint (pegainfer_b200/synthetic.py), teacher-forced with the oracle's own greedy tokens:
per step (prefill + n decode steps) the 256 largest logits (index + bf16 bits), every 16th logit of the row,
the row's max |logit| (the ulp scale of the SURVEY 8c rule) and the top-1/top-2 margin.  ~20 KB per step instead of
the 300 KB full row; written by tests/golden/make_parity_fixtures.py (committed generator).
"""
from __future__ import annotations

import json
import os
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TOPK = 256
STRIDE = 16


def fixture_path(model: str, prompt_len: int, tp_world: int) -> str:
    return os.path.join(HERE, f"parity_{model}_p{prompt_len}_tp{tp_world}.npz")


def _f32(bits: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(bits, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def bf16_ulp(x: float) -> float:
    a = max(abs(float(x)), 2.0 ** -126)
    return float(2.0 ** (np.floor(np.log2(a)) - 7))


def weights_crc(get_tensor_bytes) -> int:
    """CRC32 over the leading MiB of a few tensors.  `get_tensor_bytes(name)` -> bytes-like of the bf16 tensor."""
    crc = 0
    for name in ("model.embed_tokens.weight", "model.layers.0.self_attn.q_proj.weight",
                 "model.layers.1.mlp.down_proj.weight", "model.norm.weight"):
        b = memoryview(get_tensor_bytes(name)).cast("B")
        crc = zlib.crc32(b[: 1 << 20], crc)
    return crc & 0xFFFFFFFF


def torch_weights_crc(weights: dict) -> int:
    import torch

    def get(name):
        t = weights[name].detach().cpu().contiguous().view(torch.int16).numpy()
        return t.reshape(-1).view(np.uint8)[: 1 << 20].tobytes()
    return weights_crc(get)


def numpy_weights_crc(weights_np: dict) -> int:
    return weights_crc(lambda name: np.ascontiguousarray(weights_np[name]).reshape(-1).view(np.uint8)[: 1 << 20].tobytes())


def pack_row(bits: np.ndarray) -> dict:
    w = _f32(bits)
    order = np.argsort(-w, kind="stable")[:TOPK].astype(np.int32)
    return dict(idx_top=order, val_top=np.asarray(bits, np.uint16)[order], val_str=np.asarray(bits, np.uint16)[::STRIDE].copy(),
                rowmax=np.float32(np.abs(w).max()), margin=np.float32(w[order[0]] - w[order[1]]))


def save(path: str, meta: dict, tokens, rows: list[dict]) -> None:
    np.savez_compressed(
        path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), tokens=np.asarray(tokens, np.int32),
        idx_top=np.stack([r["idx_top"] for r in rows]), val_top=np.stack([r["val_top"] for r in rows]),
        val_str=np.stack([r["val_str"] for r in rows]), rowmax=np.array([r["rowmax"] for r in rows], np.float32),
        margin=np.array([r["margin"] for r in rows], np.float32))


class Fixture:
    def __init__(self, path: str):
        z = np.load(path)
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.tokens = [int(t) for t in z["tokens"]]
        self.idx_top, self.val_top, self.val_str = z["idx_top"], z["val_top"], z["val_str"]
        self.rowmax, self.margin = z["rowmax"], z["margin"]
        self.steps = self.idx_top.shape[0]

    def oracle_argmax(self, step: int) -> int:
        return int(self.idx_top[step][0])

    def compare(self, step: int, got_bits: np.ndarray, max_ulp_rowmax: float) -> tuple[bool, dict]:
        """SURVEY 8c rule on the sampled row: |dlogit| <= N bf16 ulps at the row's max magnitude on the top-256 and
        the strided sample, arg-max equal unless the oracle's top-1/top-2 gap is inside the tolerance."""
        g = _f32(np.asarray(got_bits, np.uint16).reshape(-1))
        tol = max_ulp_rowmax * bf16_ulp(self.rowmax[step])
        e_top = np.abs(g[self.idx_top[step]] - _f32(self.val_top[step])).max()
        e_str = np.abs(g[::STRIDE] - _f32(self.val_str[step])).max()
        err = float(max(e_top, e_str))
        same = int(g.argmax()) == self.oracle_argmax(step)
        ok = bool(np.isfinite(g).all()) and err <= tol and (same or float(self.margin[step]) <= 2 * tol)
        return ok, dict(err=err, tol=tol, err_ulp_rowmax=err / bf16_ulp(self.rowmax[step]), margin=float(self.margin[step]),
                        same_argmax=same)
