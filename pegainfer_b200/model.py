"""Python front end of the C++ host mirror (``libpegainfer_qwen3_host.so``).

Mirrors the reference's executor-facing surface for the hot path
(pegainfer-qwen3-4b/src/executor.rs:541-640 ``Qwen3Executor::{from_runtime, execute_prefill,
execute_decode, drop_request}``): create a model from a checkpoint dict, allocate per-request KV
state, run prefill / decode steps, sample greedily.  All compute happens in the sm_100a kernels
behind the pegainfer-kernels C ABI; there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import ffi
from .config import Qwen3Config, TensorParallelConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libpegainfer_qwen3_host.so")


class _PqConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("hidden_size", "intermediate_size", "num_hidden_layers",
                                       "num_attention_heads", "num_key_value_heads", "head_dim", "vocab_size")] + \
               [("rms_norm_eps", C.c_float), ("rope_theta", C.c_float), ("tie_word_embeddings", C.c_int)]


class _PqRuntime(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("device_ordinal", "tp_rank", "tp_world", "enable_cuda_graph", "mode",
                                       "num_pages", "max_batch", "enable_pdl")]


_host = None


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise OSError(f"{HOST_LIB_PATH} not found: run `python -m pegainfer_b200.build`")
        h = C.CDLL(HOST_LIB_PATH)
        vp, i32 = C.c_void_p, C.c_int
        h.pq_create_error.restype = C.c_char_p
        h.pq_model_create.restype = vp
        h.pq_model_create.argtypes = [C.POINTER(_PqConfig), C.POINTER(_PqRuntime), C.c_char_p, vp]
        h.pq_model_destroy.argtypes = [vp]
        h.pq_last_error.restype = C.c_char_p
        h.pq_last_error.argtypes = [vp]
        h.pq_model_load_tensor.argtypes = [vp, C.c_char_p, vp, i32, i32]
        h.pq_model_finalize.argtypes = [vp]
        h.pq_stream.restype = vp
        h.pq_stream.argtypes = [vp]
        h.pq_kv_alloc.argtypes = [vp]
        h.pq_kv_free.argtypes = [vp, i32]
        h.pq_kv_seq_len.argtypes = [vp, i32]
        h.pq_available_pages.argtypes = [vp]
        h.pq_prefill.argtypes = [vp, i32, vp, vp, vp, vp]
        h.pq_decode.argtypes = [vp, i32, vp, vp, vp, vp]
        h.pq_unified_step.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, vp]
        h.pq_logprobs_host.argtypes = [vp, i32, i32, i32, vp, vp, vp]
        h.pq_sample_greedy.argtypes = [vp, vp, vp]
        h.pq_sync.argtypes = [vp]
        h.pq_copy_out.argtypes = [vp, vp, vp, C.c_int64]
        h.pq_launch_count.restype = C.c_int64
        h.pq_launch_count.argtypes = [vp, i32]
        h.pq_debug_buffer.restype = vp
        h.pq_debug_buffer.argtypes = [vp, C.c_char_p]
        h.pq_generate.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        h.pq_decode_burst.argtypes = [vp, i32, C.c_uint32, i32, vp, vp]
        h.pq_bench_gemv_pass.argtypes = [vp, i32, vp, vp]
        h.pq_launches_per_step.restype = C.c_int64
        h.pq_launches_per_step.argtypes = [vp]
        h.pq_logits_cols.argtypes = [vp, vp]
        h.pq_meta_bytes.restype = C.c_int64
        h.pq_meta_bytes.argtypes = [vp]
        h.pq_event_record.restype = vp
        h.pq_event_record.argtypes = [vp]
        h.pq_event_elapsed_ms.restype = C.c_float
        h.pq_event_elapsed_ms.argtypes = [vp, vp]
        _host = h
    return _host


@dataclass
class ModelRuntimeConfig:
    """weights.rs:13-29 ``ModelRuntimeConfig`` plus the B200 switches."""
    enable_cuda_graph: bool = True
    tensor_parallel: TensorParallelConfig = TensorParallelConfig()
    device_ordinal: int = 0
    fused: bool = True          # False: the reference's op sequence through the ffi.rs ABI only
    num_pages: int = 0          # 0: 85 % of free memory
    max_batch: int = 4
    enable_pdl: bool = True
    kernel_lib: str | None = None  # default: libpegainfer_kernels_b200.so


def token_logprobs(logits_row: torch.Tensor, token: int, top_k: int = 0):
    """executor.rs:400-436 `compute_logprobs_from_cpu`: (logprob of `token`, [(id, logprob)] of the top_k largest) from one
    bf16 logits row (any device; computed on the host in f32 like the reference)."""
    row = logits_row.detach().to("cpu").contiguous().view(torch.int16)
    lp = C.c_float()
    ids = (C.c_int * max(top_k, 1))()
    vals = (C.c_float * max(top_k, 1))()
    n = host_lib().pq_logprobs_host(row.data_ptr(), row.numel(), int(token), int(top_k), C.byref(lp), ids, vals)
    if n < 0:
        raise ValueError("token out of range / empty logits")
    return lp.value, [(ids[i], vals[i]) for i in range(n)]


class Qwen3Model:
    def __init__(self, cfg: Qwen3Config, weights,
                 runtime: ModelRuntimeConfig | None = None, tp_comm: int | None = None):
        """``weights``: dict or iterable of (HF tensor name, bf16 tensor on host or device); every
        tensor is the FULL unsharded matrix, the loader takes this rank's shard (weights.rs:121-291).
        ``None``: stream tensors in with ``load_tensor`` and call ``finalize()``."""
        rt = runtime or ModelRuntimeConfig()
        rt.tensor_parallel.validate_for(cfg)
        if not torch.cuda.is_available():
            raise RuntimeError("pegainfer_b200 needs a CUDA device (there is no CPU path)")
        self.cfg, self.rt = cfg, rt
        self._h = host_lib()
        pc = _PqConfig(cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                       cfg.num_key_value_heads, cfg.head_dim, cfg.vocab_size, cfg.rms_norm_eps, cfg.rope_theta,
                       int(cfg.tie_word_embeddings))
        pr = _PqRuntime(rt.device_ordinal, rt.tensor_parallel.rank, rt.tensor_parallel.world_size,
                        int(rt.enable_cuda_graph), 1 if rt.fused else 0, rt.num_pages, rt.max_batch,
                        int(rt.enable_pdl))
        lib_path = rt.kernel_lib or ffi.KERNEL_LIB_PATH
        torch.cuda.set_device(rt.device_ordinal)
        self._m = self._h.pq_model_create(C.byref(pc), C.byref(pr), lib_path.encode(), tp_comm)
        if not self._m:
            raise RuntimeError("pq_model_create: " + self._h.pq_create_error().decode())
        self._finalized = False
        if weights is not None:
            for name, t in (weights.items() if isinstance(weights, dict) else weights):
                self.load_tensor(name, t)
            self.finalize()

    def load_tensor(self, name: str, t: torch.Tensor) -> None:
        """Feed one FULL (unsharded) HF tensor; the loader keeps this rank's shard.  For callers that stream a
        checkpoint into several models at once (``weights=None`` at construction, then ``finalize()``)."""
        if name == "lm_head.weight" and self.cfg.tie_word_embeddings:
            return
        if t.is_cuda:
            torch.cuda.current_stream().synchronize()
        assert t.dtype == torch.bfloat16 and t.is_contiguous(), name
        rows, cols = (t.shape[0], t.shape[1]) if t.dim() == 2 else (1, t.shape[0])
        self._ck(self._h.pq_model_load_tensor(self._m, name.encode(), t.data_ptr(), rows, cols))

    def finalize(self) -> None:
        self._ck(self._h.pq_model_finalize(self._m))
        self._finalized = True

    @classmethod
    def from_safetensors(cls, model_path: str, runtime: ModelRuntimeConfig | None = None, tp_comm: int | None = None):
        """``Qwen3Model::from_safetensors_with_runtime`` (weights.rs:83-125): ``config.json`` + ``model.safetensors`` (or the
        sharded index) of an HF Qwen3 checkpoint in bf16.  ``stop_token_ids`` of the checkpoint end up on the instance."""
        from .weights import iter_safetensors, load_config
        cfg, stop = load_config(model_path)
        m = cls(cfg, iter_safetensors(model_path, cfg), runtime, tp_comm)
        m.stop_token_ids = stop
        return m

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self._h.pq_last_error(self._m).decode())

    def close(self):
        if getattr(self, "_m", None):
            self._h.pq_model_destroy(self._m)
            self._m = None

    __del__ = close

    # -- request state (executor.rs: alloc on first prefill, drop_request) --
    def alloc_kv(self) -> int:
        return self._h.pq_kv_alloc(self._m)

    def drop_request(self, kv_id: int) -> None:
        self._h.pq_kv_free(self._m, kv_id)

    def kv_seq_len(self, kv_id: int) -> int:
        return self._h.pq_kv_seq_len(self._m, kv_id)

    def available_pages(self) -> int:
        return self._h.pq_available_pages(self._m)

    def logits_shard(self) -> tuple[int, int]:
        """(columns, first vocabulary id) of the logits rows this rank returns: the whole vocabulary, or -- on the
        fused tensor-parallel path, where lm_head is vocab-sharded -- this rank's slice of it."""
        off = C.c_int()
        cols = self._h.pq_logits_cols(self._m, C.byref(off))
        return int(cols), int(off.value)

    def gather_logits(self, local: torch.Tensor, dist) -> torch.Tensor:
        """Full-vocabulary rows from every rank's shard (test / parity plumbing over torch.distributed; the decode
        path itself only exchanges each shard's (max, index))."""
        cols, _ = self.logits_shard()
        if cols == self.cfg.vocab_size or dist is None:
            return local
        parts = [torch.empty_like(local) for _ in range(self.rt.tensor_parallel.world_size)]
        dist.all_gather(parts, local.contiguous())
        return torch.cat(parts, dim=1)

    def _logits_view(self, ptr: int, rows: int) -> torch.Tensor:
        out = torch.empty((rows, self.logits_shard()[0]), dtype=torch.bfloat16, device="cuda")
        torch.cuda.current_stream().synchronize()
        self._ck(self._h.pq_copy_out(self._m, out.data_ptr(), ptr, out.numel() * 2))
        return out

    # -- execute_prefill: last-token logits per request, bf16 [n_req, vocab] --
    def prefill(self, prompts: list[list[int]], kv_ids: list[int]) -> torch.Tensor:
        n = len(prompts)
        flat = [t for p in prompts for t in p]
        toks = (C.c_uint32 * len(flat))(*flat)
        lens = (C.c_int * n)(*[len(p) for p in prompts])
        ids = (C.c_int * n)(*kv_ids)
        outs = (C.c_void_p * n)()
        self._ck(self._h.pq_prefill(self._m, n, toks, lens, ids, outs))
        return torch.cat([self._logits_view(outs[i], 1) for i in range(n)], dim=0)

    # -- unified step (unified_forward.rs:78-567): prompts and decode tokens in one forward pass --
    def unified_step(self, prompts: list[list[int]], prefill_kv_ids: list[int], decode_tokens: list[int], decode_kv_ids: list[int]):
        """Returns (prefill last-token logits [n_prompts, vocab], decode logits [n_decode, vocab])."""
        n, nd = len(prompts), len(decode_tokens)
        flat = [t for p in prompts for t in p]
        toks = (C.c_uint32 * len(flat))(*flat)
        lens = (C.c_int * n)(*[len(p) for p in prompts])
        ids = (C.c_int * n)(*prefill_kv_ids)
        dt = (C.c_uint32 * max(nd, 1))(*decode_tokens)
        did = (C.c_int * max(nd, 1))(*decode_kv_ids)
        outs, douts = (C.c_void_p * n)(), (C.c_void_p * max(nd, 1))()
        self._ck(self._h.pq_unified_step(self._m, n, toks, lens, ids, nd, dt, did, outs, douts))
        pl = torch.cat([self._logits_view(outs[i], 1) for i in range(n)], dim=0)
        dl = torch.cat([self._logits_view(douts[i], 1) for i in range(nd)], dim=0) if nd else pl[:0]
        return pl, dl

    # -- execute_decode: one token per request; returns (logits [bs, vocab], greedy tokens) --
    def decode(self, tokens: list[int], kv_ids: list[int], want_logits: bool = True):
        bs = len(tokens)
        toks = (C.c_uint32 * bs)(*tokens)
        ids = (C.c_int * bs)(*kv_ids)
        lg = C.c_void_p()
        sampled = (C.c_int * bs)()
        self._ck(self._h.pq_decode(self._m, bs, toks, ids, C.byref(lg), sampled))
        logits = self._logits_view(lg.value, bs) if want_logits else None
        return logits, list(sampled)

    def sample_greedy(self, logits_row: torch.Tensor) -> int:
        out = C.c_int()
        self._ck(self._h.pq_sample_greedy(self._m, logits_row.data_ptr(), C.byref(out)))
        return out.value

    def generate(self, prompt: list[int], max_tokens: int):
        """Greedy generation timed like bench_serving.rs: returns (tokens, ttft_ms, step_ms[])."""
        toks = (C.c_uint32 * len(prompt))(*prompt)
        out = (C.c_uint32 * max_tokens)()
        ttft = C.c_double()
        steps = (C.c_double * max(1, max_tokens - 1))()
        self._ck(self._h.pq_generate(self._m, toks, len(prompt), max_tokens, out, C.byref(ttft), steps))
        return list(out), ttft.value, list(steps)[:max_tokens - 1]

    def decode_burst(self, kv_id: int, first_token: int, steps: int):
        """`steps` greedy decode steps with no host round trip (bench.py's device-resident leg)."""
        out = (C.c_uint32 * steps)()
        ms = C.c_float()
        self._ck(self._h.pq_decode_burst(self._m, kv_id, first_token, steps, out, C.byref(ms)))
        return list(out), ms.value

    def bench_gemv_pass(self, iters: int):
        ms, n = C.c_float(), C.c_int()
        self._ck(self._h.pq_bench_gemv_pass(self._m, iters, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launches_per_step(self) -> int:
        return int(self._h.pq_launches_per_step(self._m))

    def meta_bytes(self) -> int:
        return int(self._h.pq_meta_bytes(self._m))

    def event_record(self):
        return self._h.pq_event_record(self._m)

    def event_elapsed_ms(self, e0, e1) -> float:
        return float(self._h.pq_event_elapsed_ms(e0, e1))

    def launch_count(self, reset: bool = False) -> int:
        return int(self._h.pq_launch_count(self._m, int(reset)))

    def sync(self):
        self._ck(self._h.pq_sync(self._m))

    @property
    def stream(self) -> int:
        return self._h.pq_stream(self._m)

    def debug_buffer(self, name: str, numel: int) -> torch.Tensor:
        ptr = self._h.pq_debug_buffer(self._m, name.encode())
        out = torch.empty(numel, dtype=torch.bfloat16, device="cuda")
        torch.cuda.current_stream().synchronize()
        self._ck(self._h.pq_copy_out(self._m, out.data_ptr(), ptr, numel * 2))
        return out
