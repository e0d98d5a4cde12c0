"""Python front end of the C++ Qwen3.5 hybrid host (``libpegainfer_qwen3_host.so``, csrc/host/qwen35_host.cpp).

Mirrors the reference's Qwen3.5 executor surface for one request at a time (pegainfer-qwen35-4b/src/lib.rs,
prefill.rs, batch_decode.rs): create a model from HF ``qwen3_5`` text-model tensors, allocate a request (paged KV for
the full-attention layers + conv / delta-rule state for the linear layers), prefill, decode, greedy generation.
All compute runs in the sm_100a kernels behind the pegainfer-kernels C ABI; there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import torch

from . import ffi
from .model import host_lib


@dataclass(frozen=True)
class Qwen35Config:
    """pegainfer-qwen35-4b/src/config.rs:42-155 (HF ``text_config`` keys)."""
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
    layer_types: tuple = field(default_factory=tuple)  # "full_attention" | "linear_attention"
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e7
    partial_rotary_factor: float = 0.25
    name: str = "custom"

    @property
    def rotary_dim(self) -> int:
        return int(self.head_dim * self.partial_rotary_factor)


def _layer_types(n: int, full_every: int = 4) -> tuple:
    return tuple("full_attention" if (i + 1) % full_every == 0 else "linear_attention" for i in range(n))


# Qwen3.5-4B text model (docs/models/qwen35/optimization.md:57-77): 24 linear + 8 full-attention layers
QWEN35_4B = Qwen35Config(2560, 9216, 32, 16, 4, 256, 248320, 16, 32, 128, 128, 4, _layer_types(32), 1e-6, 1e7, 0.25, "qwen3.5-4b")
# small hybrid stack for parity tests the numpy oracle finishes in seconds (the kernels' head sizes are fixed: 256 / 128 x 128)
QWEN35_TINY = Qwen35Config(256, 512, 4, 4, 1, 256, 512, 2, 4, 128, 128, 4,
                           ("linear_attention", "linear_attention", "full_attention", "linear_attention"), 1e-6, 1e7, 0.25, "qwen3.5-tiny")


class _Cfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("hidden_size", "intermediate_size", "num_hidden_layers", "vocab_size", "num_attention_heads",
                                       "num_key_value_heads", "head_dim", "linear_num_key_heads", "linear_key_head_dim",
                                       "linear_num_value_heads", "linear_value_head_dim", "linear_conv_kernel_dim")] + \
               [("rms_norm_eps", C.c_float), ("rope_theta", C.c_float)] + \
               [(n, C.c_int) for n in ("rotary_dim", "enable_cuda_graph", "enable_pdl", "device_ordinal")]


def weight_shapes(c: Qwen35Config) -> dict:
    """HF tensor name -> (shape, dtype) of the text model (weights.rs:25-390)."""
    s = {"model.embed_tokens.weight": ((c.vocab_size, c.hidden_size), torch.bfloat16), "model.norm.weight": ((c.hidden_size,), torch.bfloat16)}
    qkv = 2 * c.linear_num_key_heads * c.linear_key_head_dim + c.linear_num_value_heads * c.linear_value_head_dim
    z = c.linear_num_value_heads * c.linear_value_head_dim
    for i, kind in enumerate(c.layer_types):
        p = f"model.layers.{i}."
        s[p + "input_layernorm.weight"] = ((c.hidden_size,), torch.bfloat16)
        s[p + "post_attention_layernorm.weight"] = ((c.hidden_size,), torch.bfloat16)
        s[p + "mlp.gate_proj.weight"] = ((c.intermediate_size, c.hidden_size), torch.bfloat16)
        s[p + "mlp.up_proj.weight"] = ((c.intermediate_size, c.hidden_size), torch.bfloat16)
        s[p + "mlp.down_proj.weight"] = ((c.hidden_size, c.intermediate_size), torch.bfloat16)
        if kind == "full_attention":
            a = p + "self_attn."
            s[a + "q_proj.weight"] = ((c.num_attention_heads * c.head_dim * 2, c.hidden_size), torch.bfloat16)
            s[a + "k_proj.weight"] = ((c.num_key_value_heads * c.head_dim, c.hidden_size), torch.bfloat16)
            s[a + "v_proj.weight"] = ((c.num_key_value_heads * c.head_dim, c.hidden_size), torch.bfloat16)
            s[a + "o_proj.weight"] = ((c.hidden_size, c.num_attention_heads * c.head_dim), torch.bfloat16)
            s[a + "q_norm.weight"] = ((c.head_dim,), torch.bfloat16)
            s[a + "k_norm.weight"] = ((c.head_dim,), torch.bfloat16)
        else:
            a = p + "linear_attn."
            s[a + "in_proj_qkv.weight"] = ((qkv, c.hidden_size), torch.bfloat16)
            s[a + "in_proj_z.weight"] = ((z, c.hidden_size), torch.bfloat16)
            s[a + "in_proj_b.weight"] = ((c.linear_num_value_heads, c.hidden_size), torch.bfloat16)
            s[a + "in_proj_a.weight"] = ((c.linear_num_value_heads, c.hidden_size), torch.bfloat16)
            s[a + "conv1d.weight"] = ((qkv, c.linear_conv_kernel_dim), torch.bfloat16)
            s[a + "dt_bias"] = ((c.linear_num_value_heads,), torch.bfloat16)
            s[a + "A_log"] = ((c.linear_num_value_heads,), torch.float32)
            s[a + "norm.weight"] = ((c.linear_value_head_dim,), torch.float32)
            s[a + "out_proj.weight"] = ((c.hidden_size, z), torch.bfloat16)
    return s


def iter_random_weights(c: Qwen35Config, seed: int = 0, device: str = "cpu"):
    """Random-init hybrid checkpoint (no network): N(0, s) projections with s chosen so activations stay O(1), norm
    offsets / gated-norm weights near their trained ranges.  Deterministic per (seed, device type)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    for name, (shape, dt) in weight_shapes(c).items():
        if name.endswith("A_log"):
            t = torch.randn(shape, generator=g, device=device) * 0.5
        elif name.endswith("linear_attn.norm.weight"):
            t = torch.randn(shape, generator=g, device=device) * 0.2 + 1
        elif name.endswith("dt_bias"):
            t = torch.randn(shape, generator=g, device=device) * 0.5
        elif len(shape) == 1:
            t = torch.randn(shape, generator=g, device=device) * 0.1  # (1 + w) norms: w near 0
        elif name.endswith("conv1d.weight"):
            t = torch.randn(shape, generator=g, device=device) * 0.4
        elif name.endswith("embed_tokens.weight"):
            t = torch.empty(shape, dtype=torch.bfloat16, device=device)
            rows = max(1, (1 << 26) // shape[1])
            for r0 in range(0, shape[0], rows):
                r1 = min(shape[0], r0 + rows)
                t[r0:r1] = (torch.randn((r1 - r0, shape[1]), generator=g, device=device) * 0.05).to(torch.bfloat16)
        else:
            t = torch.randn(shape, generator=g, device=device) * (0.5 / shape[1] ** 0.5)
        yield name, t.to(dt)


class Qwen35Model:
    def __init__(self, cfg: Qwen35Config, weights, num_pages: int = 0, enable_cuda_graph: bool = True, enable_pdl: bool = True,
                 device_ordinal: int = 0, kernel_lib: str | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("pegainfer_b200 needs a CUDA device (there is no CPU path)")
        self.cfg, self._h = cfg, host_lib()
        h, vp, i32 = self._h, C.c_void_p, C.c_int
        if not getattr(h, "_pq35_typed", False):
            h.pq35_create_error.restype = C.c_char_p
            h.pq35_create.restype = vp
            h.pq35_create.argtypes = [C.POINTER(_Cfg), vp, C.c_char_p]
            h.pq35_destroy.argtypes = [vp]
            h.pq35_last_error.restype = C.c_char_p
            h.pq35_last_error.argtypes = [vp]
            h.pq35_load_tensor.argtypes = [vp, C.c_char_p, vp, i32, i32, i32]
            h.pq35_finalize.argtypes = [vp, i32]
            h.pq35_request_alloc.argtypes = [vp]
            h.pq35_request_free.argtypes = [vp, i32]
            h.pq35_seq_len.argtypes = [vp, i32]
            h.pq35_prefill.argtypes = [vp, i32, vp, i32, vp]
            h.pq35_decode.argtypes = [vp, i32, C.c_uint32, vp, vp]
            h.pq35_copy_out.argtypes = [vp, vp, vp, C.c_int64]
            h.pq35_launches_per_step.restype = C.c_int64
            h.pq35_launches_per_step.argtypes = [vp]
            h.pq35_generate.argtypes = [vp, vp, i32, i32, vp, vp, vp]
            h._pq35_typed = True
        c = cfg
        pc = _Cfg(c.hidden_size, c.intermediate_size, c.num_hidden_layers, c.vocab_size, c.num_attention_heads, c.num_key_value_heads,
                  c.head_dim, c.linear_num_key_heads, c.linear_key_head_dim, c.linear_num_value_heads, c.linear_value_head_dim,
                  c.linear_conv_kernel_dim, c.rms_norm_eps, c.rope_theta, c.rotary_dim, int(enable_cuda_graph), int(enable_pdl), device_ordinal)
        kinds = (C.c_int * c.num_hidden_layers)(*[1 if t == "full_attention" else 0 for t in c.layer_types])
        torch.cuda.set_device(device_ordinal)
        self._m = h.pq35_create(C.byref(pc), kinds, (kernel_lib or ffi.KERNEL_LIB_PATH).encode())
        if not self._m:
            raise RuntimeError("pq35_create: " + h.pq35_create_error().decode())
        for name, t in (weights.items() if isinstance(weights, dict) else weights):
            if t.is_cuda:
                torch.cuda.current_stream().synchronize()
            if t.dim() == 3:  # HF stores the depthwise conv as [channels, 1, k]
                t = t.reshape(t.shape[0], t.shape[2])
            t = t.contiguous()
            assert t.dtype in (torch.bfloat16, torch.float32), name
            rows, cols = (t.shape[0], t.shape[1]) if t.dim() == 2 else (1, t.shape[0])
            self._ck(h.pq35_load_tensor(self._m, name.encode(), t.data_ptr(), rows, cols, int(t.dtype == torch.float32)))
        self._ck(h.pq35_finalize(self._m, num_pages))

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(self._h.pq35_last_error(self._m).decode())

    def close(self):
        if getattr(self, "_m", None):
            self._h.pq35_destroy(self._m)
            self._m = None

    __del__ = close

    def alloc_request(self) -> int:
        rid = self._h.pq35_request_alloc(self._m)
        if rid < 0:
            raise RuntimeError(self._h.pq35_last_error(self._m).decode())
        return rid

    def drop_request(self, rid: int) -> None:
        self._h.pq35_request_free(self._m, rid)

    def seq_len(self, rid: int) -> int:
        return self._h.pq35_seq_len(self._m, rid)

    def _logits(self, ptr) -> torch.Tensor:
        out = torch.empty(self.cfg.vocab_size, dtype=torch.bfloat16, device="cuda")
        torch.cuda.current_stream().synchronize()
        self._ck(self._h.pq35_copy_out(self._m, out.data_ptr(), ptr, out.numel() * 2))
        return out

    def prefill(self, rid: int, tokens: list[int]) -> torch.Tensor:
        """Last-token logits [vocab] (prefill.rs:24-110); may be called again on the same request (chunked prompt)."""
        toks = (C.c_uint32 * len(tokens))(*tokens)
        lg = C.c_void_p()
        self._ck(self._h.pq35_prefill(self._m, rid, toks, len(tokens), C.byref(lg)))
        return self._logits(lg.value)

    def decode(self, rid: int, token: int, want_logits: bool = True):
        """One token (batch_decode.rs:194-364 at batch 1, CUDA graph): (logits [vocab] or None, greedy token)."""
        lg, sampled = C.c_void_p(), C.c_int()
        self._ck(self._h.pq35_decode(self._m, rid, token, C.byref(lg), C.byref(sampled)))
        return (self._logits(lg.value) if want_logits else None), sampled.value

    def generate(self, prompt: list[int], max_tokens: int):
        """Greedy generation timed like bench_serving.rs: (tokens, ttft_ms, step_ms[])."""
        toks = (C.c_uint32 * len(prompt))(*prompt)
        out = (C.c_uint32 * max_tokens)()
        ttft = C.c_double()
        steps = (C.c_double * max(1, max_tokens - 1))()
        self._ck(self._h.pq35_generate(self._m, toks, len(prompt), max_tokens, out, C.byref(ttft), steps))
        return list(out), ttft.value, list(steps)[:max_tokens - 1]

    def launches_per_step(self) -> int:
        return int(self._h.pq35_launches_per_step(self._m))
