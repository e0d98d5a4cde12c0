"""Random-init Qwen3 checkpoints and synthetic prompts (no network, no real weights).

Weights follow HF ``Qwen3ForCausalLM`` default init (Linear / Embedding N(0, 0.02),
norm weights 1.0 -- SURVEY.md section 8d) under the HF tensor names the reference's
loader reads (pegainfer-qwen3-4b/src/weights.rs:100-291).  Prompt ids follow
pegainfer-server/src/bin/bench_serving.rs:761-763: ``(i % 1000) + 100``.
"""
from __future__ import annotations

import torch

from .config import Qwen3Config


def synthetic_prompt(n: int, offset: int = 0) -> list[int]:
    return [((i + offset) % 1000) + 100 for i in range(n)]


def weight_shapes(cfg: Qwen3Config) -> dict[str, tuple[int, ...]]:
    c = cfg
    s: dict[str, tuple[int, ...]] = {"model.embed_tokens.weight": (c.vocab_size, c.hidden_size)}
    for i in range(c.num_hidden_layers):
        p = f"model.layers.{i}."
        s[p + "input_layernorm.weight"] = (c.hidden_size,)
        s[p + "self_attn.q_proj.weight"] = (c.q_dim, c.hidden_size)
        s[p + "self_attn.k_proj.weight"] = (c.kv_dim, c.hidden_size)
        s[p + "self_attn.v_proj.weight"] = (c.kv_dim, c.hidden_size)
        s[p + "self_attn.o_proj.weight"] = (c.hidden_size, c.q_dim)
        s[p + "self_attn.q_norm.weight"] = (c.head_dim,)
        s[p + "self_attn.k_norm.weight"] = (c.head_dim,)
        s[p + "post_attention_layernorm.weight"] = (c.hidden_size,)
        s[p + "mlp.gate_proj.weight"] = (c.intermediate_size, c.hidden_size)
        s[p + "mlp.up_proj.weight"] = (c.intermediate_size, c.hidden_size)
        s[p + "mlp.down_proj.weight"] = (c.hidden_size, c.intermediate_size)
    s["model.norm.weight"] = (c.hidden_size,)
    if not c.tie_word_embeddings:
        s["lm_head.weight"] = (c.vocab_size, c.hidden_size)
    return s


_CHUNK_ELEMS = 1 << 25  # elements per independently seeded block of a matrix (fp32 staging: 128 MB per worker)


def _block_seed(seed: int, tensor_index: int, block: int) -> int:
    return (seed * 1000003 + tensor_index) * 4099 + block


def _fill_block(t: torch.Tensor, r0: int, r1: int, seed: int) -> None:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    t[r0:r1] = (torch.randn((r1 - r0, t.shape[1]), generator=g, dtype=torch.float32) * 0.02).to(torch.bfloat16)


def _cpu_tensor_jobs(name, shape, idx, seed, norm_jitter):
    """(tensor, [callables that fill it]) -- every block of a matrix has its own generator seeded by
    (seed, tensor index, block index), so the blocks can be filled by a thread pool in any order and the checkpoint is
    the same on every machine and for every worker count."""
    if len(shape) == 1:
        t = torch.ones(shape, dtype=torch.float32)
        if norm_jitter:
            g = torch.Generator(device="cpu")
            g.manual_seed(_block_seed(seed, idx, 0))
            t += norm_jitter * torch.randn(shape, generator=g, dtype=torch.float32)
        return t.to(torch.bfloat16), []
    t = torch.empty(shape, dtype=torch.bfloat16)
    rows = max(1, _CHUNK_ELEMS // shape[1])
    jobs = [(lambda r0=r0, b=b: _fill_block(t, r0, min(shape[0], r0 + rows), _block_seed(seed, idx, b)))
            for b, r0 in enumerate(range(0, shape[0], rows))]
    return t, jobs


def iter_random_weights(cfg: Qwen3Config, seed: int = 0, device: str = "cpu", norm_jitter: float = 0.0, workers: int | None = None):
    """Yield (HF name, bf16 tensor) one at a time, in `weight_shapes` order.

    device "cpu" (the checkpoint of the parity tests, the oracle fixtures and bench.py): every <= 32 Mi-element block
    of every matrix is drawn from its own torch CPU generator seeded by (seed, tensor index, block index) and the
    blocks are filled by a thread pool (torch releases the GIL inside randn) -- deterministic regardless of the number
    of workers, ~10 x faster than one sequential stream on a many-core host.  device "cuda": one sequential CUDA
    generator (tuning runs only; a different checkpoint).

    ``norm_jitter`` > 0 perturbs norm weights away from 1.0 (tests only) so a dropped weight
    multiply cannot hide.
    """
    shapes = weight_shapes(cfg)
    if device != "cpu":
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        for name, shape in shapes.items():
            if len(shape) == 1:
                t = torch.ones(shape, dtype=torch.float32, device=device)
                if norm_jitter:
                    t += norm_jitter * torch.randn(shape, generator=g, dtype=torch.float32, device=device)
                yield name, t.to(torch.bfloat16)
            else:
                t = torch.empty(shape, dtype=torch.bfloat16, device=device)
                rows = max(1, (1 << 26) // shape[1])  # fp32 staging in <= 256 MB pieces
                for r0 in range(0, shape[0], rows):
                    r1 = min(shape[0], r0 + rows)
                    t[r0:r1] = (torch.randn((r1 - r0, shape[1]), generator=g, dtype=torch.float32, device=device)
                                * 0.02).to(torch.bfloat16)
                yield name, t
        return
    import os
    from concurrent.futures import ThreadPoolExecutor
    try:
        ncpu = len(os.sched_getaffinity(0))
    except Exception:
        ncpu = os.cpu_count() or 1
    workers = workers or max(1, min(32, ncpu))
    items = list(shapes.items())
    window = 24  # tensors in flight ahead of the consumer (bounds host memory: the consumer may drop what it has used)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        pending = []  # (name, tensor, futures)

        def submit(i):
            name, shape = items[i]
            t, jobs = _cpu_tensor_jobs(name, shape, i, seed, norm_jitter)
            pending.append((name, t, [ex.submit(j) for j in jobs]))

        nxt = 0
        while nxt < len(items) and nxt < window:
            submit(nxt)
            nxt += 1
        while pending:
            name, t, futs = pending.pop(0)
            for f in futs:
                f.result()
            if nxt < len(items):
                submit(nxt)
                nxt += 1
            yield name, t


def random_weights(cfg: Qwen3Config, seed: int = 0, device: str = "cpu",
                   norm_jitter: float = 0.0) -> dict[str, torch.Tensor]:
    """bf16 tensors by HF name (see iter_random_weights)."""
    return dict(iter_random_weights(cfg, seed, device, norm_jitter))


def to_numpy_bits(weights: dict[str, torch.Tensor]):
    """bf16 torch tensors -> numpy uint16 bit patterns (what the CPU oracle eats)."""
    return {k: v.detach().cpu().contiguous().view(torch.int16).numpy().view("uint16")
            for k, v in weights.items()}
