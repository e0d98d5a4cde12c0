"""CPU oracle for the pegainfer Qwen3 hot path -- TEST INFRASTRUCTURE ONLY.

numpy/ctypes front end of ``qwen3_oracle.c`` plus a model-level restatement of
the reference's prefill / decode call sequences:

* ``OracleQwen3.prefill``  follows pegainfer-qwen3-4b/src/prefill.rs:73-188,220-285
* ``OracleQwen3.decode``   follows pegainfer-qwen3-4b/src/batch_decode.rs:17-295 and
  the paged / split-KV metadata of batch_decode_buffers.rs:177-287
* TP sharding follows pegainfer-qwen3-4b/src/weights.rs:121-291 and
  config.rs:150-153 (``shard_range``); the all-reduce is weights.rs:396-405.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product
(``pegainfer_b200``) never does.

All activations / weights are numpy ``uint16`` arrays holding bf16 bits.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqwen3_oracle.so")
_lib = None

_u16p = C.POINTER(C.c_uint16)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    """Compile qwen3_oracle.c with the committed Makefile (gcc, OpenMP)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "qwen3_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        try:
            build()
            _lib = C.CDLL(_LIB_PATH)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a: np.ndarray | None, ty):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle needs contiguous arrays"
    return a.ctypes.data_as(ty)


def set_num_threads(n: int) -> int:
    """Pin the oracle's OpenMP team size (torchrun exports OMP_NUM_THREADS=1); returns the size in effect."""
    lib().orc_set_num_threads(int(n))
    return int(lib().orc_get_max_threads())


def get_max_threads() -> int:
    return int(lib().orc_get_max_threads())


def spread_rows(a: np.ndarray) -> np.ndarray:
    """Copy of a 2-D matrix whose pages are first touched with orc_gemm's static row partition (NUMA locality)."""
    a = np.ascontiguousarray(a)
    if a.ndim != 2:
        return a
    out = np.empty_like(a)
    lib().orc_spread_rows(_p(out, _u16p), _p(a, _u16p), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]))
    return out


# ---------------------------------------------------------------- bf16 helpers
def f32_to_bf16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    lib().orc_f32_to_bf16(_p(x, _f32p), _p(out, _u16p), C.c_int64(x.size))
    return out


def bf16_to_f32(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.uint16)
    return (x.astype(np.uint32) << 16).view(np.float32)


def bf16_ulp(x: np.ndarray) -> np.ndarray:
    """Size of one bf16 ulp at |x| (fp32 array in, fp32 out)."""
    a = np.abs(np.asarray(x, dtype=np.float32))
    e = np.floor(np.log2(np.maximum(a, 2.0 ** -126)))
    return np.exp2(e - 7).astype(np.float32)


# ---------------------------------------------------------------- op wrappers
def embedding_batched(embed, ids, hidden):
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    out = np.empty((len(ids), hidden), dtype=np.uint16)
    lib().orc_embedding_batched(_p(embed, _u16p), _p(ids, _u32p), _p(out, _u16p), hidden, len(ids))
    return out


def embedding_batched_vocab_shard(embed, ids, hidden, vocab_start, part_vocab):
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    out = np.empty((len(ids), hidden), dtype=np.uint16)
    lib().orc_embedding_batched_vocab_shard(_p(embed, _u16p), _p(ids, _u32p), _p(out, _u16p),
                                            hidden, len(ids), C.c_uint32(vocab_start),
                                            C.c_uint32(part_vocab))
    return out


def rms_norm(x, w, eps):
    x = np.ascontiguousarray(x).reshape(-1, w.shape[0])
    out = np.empty_like(x)
    lib().orc_rms_norm_batched(_p(x, _u16p), _p(w, _u16p), _p(out, _u16p), x.shape[1], x.shape[0],
                               C.c_float(eps))
    return out


def fused_add_rms_norm(hidden, residual, w, eps):
    """In place on ``hidden``; returns the normed output."""
    h = hidden.reshape(-1, w.shape[0])
    out = np.empty_like(h)
    lib().orc_fused_add_rms_norm_batched(_p(h, _u16p), _p(np.ascontiguousarray(residual), _u16p),
                                         _p(w, _u16p), _p(out, _u16p), h.shape[1], h.shape[0],
                                         C.c_float(eps))
    return out


def add(a, b):
    out = np.empty_like(a)
    lib().orc_add(_p(a, _u16p), _p(b, _u16p), _p(out, _u16p), C.c_int64(a.size))
    return out


def silu_mul_fused(gate_up, inter):
    gu = np.ascontiguousarray(gate_up).reshape(-1, 2 * inter)
    out = np.empty((gu.shape[0], inter), dtype=np.uint16)
    lib().orc_silu_mul_fused(_p(gu, _u16p), _p(out, _u16p), inter, gu.shape[0])
    return out


def silu_mul(gate, up):
    out = np.empty_like(gate)
    lib().orc_silu_mul(_p(gate, _u16p), _p(up, _u16p), _p(out, _u16p), C.c_int64(gate.size))
    return out


def gemm(W, X):
    """Y[N, M] = X[N, K] @ W[M, K]^T  (reference: Y[M,N] col-major = W @ X)."""
    W = np.ascontiguousarray(W)
    X = np.ascontiguousarray(X).reshape(-1, W.shape[1])
    M, K = W.shape
    Y = np.empty((X.shape[0], M), dtype=np.uint16)
    lib().orc_gemm(_p(W, _u16p), _p(X, _u16p), _p(Y, _u16p), M, X.shape[0], K)
    return Y


def precompute_rope(head_dim, max_pos, theta):
    cos = np.empty(max_pos * head_dim, dtype=np.uint16)
    sin = np.empty(max_pos * head_dim, dtype=np.uint16)
    lib().orc_precompute_rope(_p(cos, _u16p), _p(sin, _u16p), head_dim, max_pos, C.c_float(theta))
    return cos, sin


def qk_norm_rope(q, k, qw, kw, cos, sin, nq, nkv, hd, eps, positions=None, start_pos=0):
    """In place on q [T, nq*hd] and k [T, nkv*hd]."""
    T = q.shape[0]
    pos = None if positions is None else np.ascontiguousarray(positions, dtype=np.int32)
    lib().orc_qk_norm_rope(_p(q, _u16p), _p(k, _u16p), _p(qw, _u16p), _p(kw, _u16p),
                           _p(cos, _u16p), _p(sin, _u16p), _p(pos, _i32p), start_pos, nq, nkv, hd,
                           T, C.c_float(eps))


def paged_kv_scatter(kv, k_off, v_off, page_indices, page_indptr, last_page_len, src_k, src_v,
                     batch_indices, positions, nkv, hd, page_size, stride_page):
    nnz = len(positions)
    return lib().orc_paged_kv_scatter(
        _p(kv, _u16p), C.c_int64(k_off), C.c_int64(v_off), _p(page_indices, _i32p),
        _p(page_indptr, _i32p), _p(last_page_len, _i32p), _p(src_k, _u16p), _p(src_v, _u16p),
        _p(batch_indices, _i32p), _p(positions, _i32p), nnz, nkv, hd, page_size,
        C.c_int64(stride_page), C.c_int64(nkv * hd), C.c_int64(hd))


def paged_attention_decode(q, kv, k_off, v_off, page_indices, page_indptr, last_page_len,
                           request_indices, kv_tile_indices, kv_chunk_size, nq, nkv, hd, page_size,
                           stride_page, sm_scale):
    bs = len(request_indices)
    out = np.empty((bs, nq * hd), dtype=np.uint16)
    rc = lib().orc_paged_attention_decode(
        _p(q, _u16p), _p(out, _u16p), _p(kv, _u16p), C.c_int64(k_off), C.c_int64(v_off),
        _p(page_indices, _i32p), _p(page_indptr, _i32p), _p(last_page_len, _i32p),
        _p(request_indices, _i32p), _p(kv_tile_indices, _i32p), _p(kv_chunk_size, _i32p), nq, nkv,
        hd, page_size, bs, C.c_int64(stride_page), C.c_float(sm_scale))
    assert rc == 0
    return out


def paged_attention_decode_split_kv(q, kv, k_off, v_off, page_indices, page_indptr, last_page_len,
                                    request_indices, kv_tile_indices, kv_chunk_size, o_indptr,
                                    block_valid_mask, nq, nkv, hd, page_size, bs, stride_page,
                                    sm_scale, return_tmp=False):
    slots = len(request_indices)
    out = np.empty((bs, nq * hd), dtype=np.uint16)
    tmp_v = np.zeros((slots, nq * hd), dtype=np.uint16)
    tmp_s = np.zeros((slots, nq), dtype=np.float32)
    rc = lib().orc_paged_attention_decode_split_kv(
        _p(q, _u16p), _p(out, _u16p), _p(kv, _u16p), C.c_int64(k_off), C.c_int64(v_off),
        _p(page_indices, _i32p), _p(page_indptr, _i32p), _p(last_page_len, _i32p),
        _p(request_indices, _i32p), _p(kv_tile_indices, _i32p), _p(kv_chunk_size, _i32p),
        _p(o_indptr, _i32p), _p(block_valid_mask, _u8p), _p(tmp_v, _u16p), _p(tmp_s, _f32p), nq,
        nkv, hd, page_size, bs, slots, C.c_int64(stride_page), C.c_float(sm_scale))
    assert rc == 0
    return (out, tmp_v, tmp_s) if return_tmp else out


def batch_prefill_paged(q, kv, k_off, v_off, page_indices, page_indptr, last_page_len, q_indptr,
                        nq, nkv, hd, page_size, stride_page, sm_scale):
    out = np.empty_like(q)
    rc = lib().orc_batch_prefill_paged(
        _p(q, _u16p), _p(out, _u16p), _p(kv, _u16p), C.c_int64(k_off), C.c_int64(v_off),
        _p(page_indices, _i32p), _p(page_indptr, _i32p), _p(last_page_len, _i32p),
        _p(q_indptr, _i32p), nq, nkv, hd, page_size, len(q_indptr) - 1, C.c_int64(stride_page),
        C.c_float(sm_scale))
    assert rc == 0
    return out


def argmax(x):
    x = np.ascontiguousarray(x)
    out = np.zeros(1, dtype=np.int32)
    lib().orc_argmax(_p(x, _u16p), _p(out, _i32p), x.size)
    return int(out[0])


def all_reduce_sum(parts):
    parts = [np.ascontiguousarray(p) for p in parts]
    arr = (_u16p * len(parts))(*[_p(p, _u16p) for p in parts])
    out = np.empty_like(parts[0])
    lib().orc_all_reduce_sum(arr, len(parts), _p(out, _u16p), C.c_int64(out.size))
    return out


# ---------------------------------------------------------------- model level
@dataclass
class OracleConfig:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    vocab_size: int
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e6
    tie_word_embeddings: bool = True
    max_position: int = 4096  # weights.rs:300 precompute_rope(head_dim, 4096, theta)


PAGE_SIZE = 16  # weights.rs:309
SPLIT_KV_CHUNK_TOKENS = 256  # batch_decode_buffers.rs:14-17
SPLIT_KV_MAX_CHUNKS = 64
SPLIT_KV_MAX_BS = 2
SPLIT_KV_MIN_SEQ = 1024


class _Rank:
    """Rank-local weights + KV pool (weights.rs:121-334)."""

    def __init__(self, cfg: OracleConfig, w: dict, rank: int, world: int, num_pages: int):
        c = cfg
        self.nq = c.num_attention_heads // world
        self.nkv = c.num_key_value_heads // world
        self.inter = c.intermediate_size // world
        hd = c.head_dim
        qs, ql = rank * self.nq * hd, self.nq * hd
        ks, kl = rank * self.nkv * hd, self.nkv * hd
        is_, il = rank * self.inter, self.inter
        self.layers = []
        for i in range(c.num_hidden_layers):
            p = f"model.layers.{i}."
            qkv = np.concatenate([w[p + "self_attn.q_proj.weight"][qs:qs + ql],
                                  w[p + "self_attn.k_proj.weight"][ks:ks + kl],
                                  w[p + "self_attn.v_proj.weight"][ks:ks + kl]], axis=0)
            gate_up = np.concatenate([w[p + "mlp.gate_proj.weight"][is_:is_ + il],
                                      w[p + "mlp.up_proj.weight"][is_:is_ + il]], axis=0)
            self.layers.append(dict(
                input_ln=w[p + "input_layernorm.weight"], post_ln=w[p + "post_attention_layernorm.weight"],
                qkv=np.ascontiguousarray(qkv), gate_up=np.ascontiguousarray(gate_up),
                o=np.ascontiguousarray(w[p + "self_attn.o_proj.weight"][:, qs:qs + ql]),
                down=np.ascontiguousarray(w[p + "mlp.down_proj.weight"][:, is_:is_ + il]),
                q_norm=w[p + "self_attn.q_norm.weight"], k_norm=w[p + "self_attn.k_norm.weight"]))
        # KvLayout::new (kv_pool.rs:28-42)
        self.kv_block_len = PAGE_SIZE * self.nkv * hd
        self.layer_stride = 2 * self.kv_block_len
        self.page_stride = c.num_hidden_layers * self.layer_stride
        self.kv = np.zeros(num_pages * self.page_stride, dtype=np.uint16)


class OracleKv:
    """KvState (kv_pool.rs:130-230): per-request page list + seq_len."""

    def __init__(self):
        self.pages: list[int] = []
        self.seq_len = 0

    def last_page_len(self):
        if self.seq_len == 0:
            return 0
        r = self.seq_len % PAGE_SIZE
        return PAGE_SIZE if r == 0 else r


class OracleQwen3:
    def __init__(self, cfg: OracleConfig, weights: dict, tp_world: int = 1, num_pages: int = 512):
        self.cfg = cfg
        self.world = tp_world
        self.embed = weights["model.embed_tokens.weight"]
        self.lm_head = self.embed if cfg.tie_word_embeddings else weights["lm_head.weight"]
        self.norm = weights["model.norm.weight"]
        self.ranks = [_Rank(cfg, weights, r, tp_world, num_pages) for r in range(tp_world)]
        self.cos, self.sin = precompute_rope(cfg.head_dim, cfg.max_position, cfg.rope_theta)
        # PagePool hands out ascending ids; page 0 is the padding page (kv_pool.rs:99-102)
        self.free = list(range(num_pages - 1, 0, -1))
        self.sm_scale = 1.0 / math.sqrt(cfg.head_dim)
        self.last_attention_path = None

    def rehome_weights(self):
        """Timing hygiene for bench.py's CPU arm: re-home every streamed matrix with spread_rows (first touch by the
        thread that reads it).  No arithmetic changes."""
        tied = self.lm_head is self.embed
        self.lm_head = spread_rows(self.lm_head)
        if tied:
            self.embed = self.lm_head
        for rk in self.ranks:
            for L in rk.layers:
                for k in ("qkv", "gate_up", "o", "down"):
                    L[k] = spread_rows(L[k])

    def fill_context(self, kv: "OracleKv", tokens: int, seed: int = 0):
        """Timing hygiene: give `kv` a context of `tokens` entries of small random K/V WITHOUT running a prefill (a
        2048-token CPU prefill is ~16 TFLOP); decode timing at that context does not depend on the values."""
        self._ensure(kv, tokens)
        kv.seq_len = tokens
        rng = np.random.default_rng(seed)
        for rk in self.ranks:
            for pg in kv.pages:
                lo = pg * rk.page_stride
                rk.kv[lo:lo + rk.page_stride] = rng.integers(0x3c00, 0x3c80, rk.page_stride, dtype=np.uint16)

    # -- paging --------------------------------------------------------------
    def alloc_kv(self) -> OracleKv:
        return OracleKv()

    def _ensure(self, kv: OracleKv, tokens: int):
        need = -(-tokens // PAGE_SIZE)
        while len(kv.pages) < need:
            kv.pages.append(self.free.pop())

    def _all_reduce(self, parts):
        return parts[0] if self.world == 1 else all_reduce_sum(parts)

    def _meta(self, kvs):
        pi, ip, lpl = [], [0], []
        for kv in kvs:
            pi += kv.pages
            ip.append(len(pi))
            lpl.append(kv.last_page_len())
        return (np.array(pi, np.int32), np.array(ip, np.int32), np.array(lpl, np.int32))

    # -- prefill (prefill.rs) --------------------------------------------------
    def prefill(self, prompts, kvs):
        c = self.cfg
        seq_lens = [len(p) for p in prompts]
        starts = [kv.seq_len for kv in kvs]
        ids = np.concatenate([np.asarray(p, np.uint32) for p in prompts])
        hidden = embedding_batched(self.embed, ids, c.hidden_size)
        for kv, s, n in zip(kvs, starts, seq_lens):
            self._ensure(kv, s + n)
            kv.seq_len += n
        pi, ip, lpl = self._meta(kvs)
        bidx = np.concatenate([np.full(n, b, np.int32) for b, n in enumerate(seq_lens)])
        pos = np.concatenate([np.arange(s, s + n, dtype=np.int32) for s, n in zip(starts, seq_lens)])
        q_indptr = np.concatenate([[0], np.cumsum(seq_lens)]).astype(np.int32)
        hd = c.head_dim
        for li in range(c.num_hidden_layers):
            o_parts, mlp_parts = [], []
            normed = rms_norm(hidden, self.ranks[0].layers[li]["input_ln"], c.rms_norm_eps)
            for rk in self.ranks:
                L = rk.layers[li]
                qd, kd = rk.nq * hd, rk.nkv * hd
                q = gemm(L["qkv"][:qd], normed)
                k = gemm(L["qkv"][qd:qd + kd], normed)
                v = gemm(L["qkv"][qd + kd:], normed)
                qk_norm_rope(q, k, L["q_norm"], L["k_norm"], self.cos, self.sin, rk.nq, rk.nkv, hd,
                             c.rms_norm_eps, positions=pos)
                k_off = li * rk.layer_stride
                v_off = k_off + rk.kv_block_len
                paged_kv_scatter(rk.kv, k_off, v_off, pi, ip, lpl, k, v, bidx, pos, rk.nkv, hd,
                                 PAGE_SIZE, rk.page_stride)
                attn = batch_prefill_paged(q, rk.kv, k_off, v_off, pi, ip, lpl, q_indptr, rk.nq,
                                           rk.nkv, hd, PAGE_SIZE, rk.page_stride, self.sm_scale)
                o_parts.append(gemm(L["o"], attn))
            o = self._all_reduce(o_parts)
            normed = fused_add_rms_norm(hidden, o, self.ranks[0].layers[li]["post_ln"], c.rms_norm_eps)
            for rk in self.ranks:
                L = rk.layers[li]
                act = silu_mul_fused(gemm(L["gate_up"], normed), rk.inter)
                mlp_parts.append(gemm(L["down"], act))
            mlp = self._all_reduce(mlp_parts)
            hidden = add(hidden, mlp)  # prefill.rs:183 rounds the residual sum to bf16
        logits = []
        off = 0
        for n in seq_lens:
            last = np.ascontiguousarray(hidden[off + n - 1:off + n])
            logits.append(gemm(self.lm_head, rms_norm(last, self.norm, c.rms_norm_eps))[0])
            off += n
        self.last_hidden = hidden
        return logits

    # -- decode (batch_decode.rs) ----------------------------------------------
    def decode(self, token_ids, kvs):
        c = self.cfg
        hd = c.head_dim
        bs = len(token_ids)
        positions = []
        for kv in kvs:
            positions.append(kv.seq_len)
            self._ensure(kv, kv.seq_len + 1)
            kv.seq_len += 1
        pos = np.array(positions, np.int32)
        pi, ip, lpl = self._meta(kvs)
        req = np.arange(bs, dtype=np.int32)
        tile0 = np.zeros(bs, np.int32)
        chunk = np.array([kv.seq_len for kv in kvs], np.int32)
        max_len = max(kv.seq_len for kv in kvs)
        split = bs <= SPLIT_KV_MAX_BS and max_len >= SPLIT_KV_MIN_SEQ
        self.last_attention_path = "split_kv" if split else "non_partition"
        if split:  # batch_decode_buffers.rs:229-279
            csz = max(SPLIT_KV_CHUNK_TOKENS, -(-max_len // SPLIT_KV_MAX_CHUNKS))
            sreq, stile, mask, oip = [], [], [], [0]
            for b, kv in enumerate(kvs):
                n = max(1, -(-kv.seq_len // csz))
                sreq += [b] * n
                stile += list(range(n))
                mask += [1] * n
                oip.append(len(sreq))
            pad = bs * SPLIT_KV_MAX_CHUNKS - len(sreq)
            sreq += [0] * pad
            stile += [0] * pad
            mask += [0] * pad
            sreq, stile = np.array(sreq, np.int32), np.array(stile, np.int32)
            mask, oip = np.array(mask, np.uint8), np.array(oip, np.int32)
            csz_a = np.array([csz], np.int32)
        hidden = embedding_batched(self.embed, np.asarray(token_ids, np.uint32), c.hidden_size)
        normed = rms_norm(hidden, self.ranks[0].layers[0]["input_ln"], c.rms_norm_eps)
        for li in range(c.num_hidden_layers):
            o_parts, mlp_parts = [], []
            for rk in self.ranks:
                L = rk.layers[li]
                qd, kd = rk.nq * hd, rk.nkv * hd
                q = gemm(L["qkv"][:qd], normed)
                k = gemm(L["qkv"][qd:qd + kd], normed)
                v = gemm(L["qkv"][qd + kd:], normed)
                qk_norm_rope(q, k, L["q_norm"], L["k_norm"], self.cos, self.sin, rk.nq, rk.nkv, hd,
                             c.rms_norm_eps, positions=pos)
                k_off = li * rk.layer_stride
                v_off = k_off + rk.kv_block_len
                paged_kv_scatter(rk.kv, k_off, v_off, pi, ip, lpl, k, v, req, pos, rk.nkv, hd,
                                 PAGE_SIZE, rk.page_stride)
                if split:
                    attn = paged_attention_decode_split_kv(
                        q, rk.kv, k_off, v_off, pi, ip, lpl, sreq, stile, csz_a, oip, mask, rk.nq,
                        rk.nkv, hd, PAGE_SIZE, bs, rk.page_stride, self.sm_scale)
                else:
                    attn = paged_attention_decode(q, rk.kv, k_off, v_off, pi, ip, lpl, req, tile0,
                                                  chunk, rk.nq, rk.nkv, hd, PAGE_SIZE,
                                                  rk.page_stride, self.sm_scale)
                o_parts.append(gemm(L["o"], attn))
            o = self._all_reduce(o_parts)
            normed = fused_add_rms_norm(hidden, o, self.ranks[0].layers[li]["post_ln"], c.rms_norm_eps)
            for rk in self.ranks:
                L = rk.layers[li]
                act = silu_mul_fused(gemm(L["gate_up"], normed), rk.inter)
                mlp_parts.append(gemm(L["down"], act))
            mlp = self._all_reduce(mlp_parts)
            nw = (self.ranks[0].layers[li + 1]["input_ln"] if li + 1 < c.num_hidden_layers
                  else self.norm)
            # batch_decode.rs:121-133: norm of the UNROUNDED fp32 residual sum
            normed = fused_add_rms_norm(hidden, mlp, nw, c.rms_norm_eps)
        self.last_hidden = hidden
        return gemm(self.lm_head, normed)  # [bs, V]
