"""Qwen3 model configuration (mirror of pegainfer-qwen3-4b/src/config.rs:23-158).

Field names follow the HF ``config.json`` keys the reference deserialises.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

PREFILL_ATTENTION_CTA_TILE_Q = 64  # config.rs:5
PAGE_SIZE = 16  # weights.rs:309
ROPE_TABLE_POSITIONS = 4096  # weights.rs:300


@dataclass(frozen=True)
class Qwen3Config:
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
    name: str = "custom"

    @property
    def q_dim(self) -> int:
        return self.num_attention_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.num_key_value_heads * self.head_dim

    def to_dict(self):
        return asdict(self)


@dataclass(frozen=True)
class TensorParallelConfig:
    """config.rs:7-19,114-158."""
    rank: int = 0
    world_size: int = 1

    def validate_for(self, c: Qwen3Config) -> None:
        if self.world_size < 1:
            raise ValueError("tensor_parallel.world_size must be >= 1")
        if not (0 <= self.rank < self.world_size):
            raise ValueError(f"tensor_parallel.rank {self.rank} must be < world_size {self.world_size}")
        for nm, v in (("num_attention_heads", c.num_attention_heads),
                      ("num_key_value_heads", c.num_key_value_heads),
                      ("intermediate_size", c.intermediate_size)):
            if v % self.world_size:
                raise ValueError(f"{nm}={v} not divisible by tp world_size={self.world_size}")

    def shard_range(self, total: int) -> tuple[int, int]:
        n = total // self.world_size
        return self.rank * n, n

    @property
    def is_sharded(self) -> bool:
        return self.world_size > 1


QWEN3_4B = Qwen3Config(2560, 9728, 36, 32, 8, 128, 151936, 1e-6, 1e6, True, "qwen3-4b")
QWEN3_8B = Qwen3Config(4096, 12288, 36, 32, 8, 128, 151936, 1e-6, 1e6, False, "qwen3-8b")
# Small shapes for parity tests the CPU oracle finishes in seconds (head_dim stays 128:
# the reference hard-codes HEAD_DIM 128, csrc/prefill_attention.cu:3).
QWEN3_TINY = Qwen3Config(256, 512, 2, 8, 2, 128, 1024, 1e-6, 1e6, True, "qwen3-tiny")
QWEN3_SMALL = Qwen3Config(1024, 3072, 4, 16, 8, 128, 8192, 1e-6, 1e6, False, "qwen3-small")

PRESETS = {c.name: c for c in (QWEN3_4B, QWEN3_8B, QWEN3_TINY, QWEN3_SMALL)}
