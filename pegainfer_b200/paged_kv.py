"""Page-first KV geometry (mirror of pegainfer-kernels/src/paged_kv.rs:5-34 and
pegainfer-core/src/kv_pool.rs:14-52).  Pure value type: no GPU, no allocation.

Pool layout (bf16 elements): ``[page][layer][K block | V block]`` with each block
``[page_size slots][num_kv_heads][head_dim]`` (NHD).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class PagedKvLayout:
    page_size: int
    num_layers: int
    num_kv_heads: int
    head_dim: int
    kv_block_len: int   # page_size * num_kv_heads * head_dim
    layer_stride: int   # 2 * kv_block_len (K then V)
    page_stride: int    # num_layers * layer_stride

    @staticmethod
    def new(num_layers: int, num_kv_heads: int, head_dim: int, page_size: int) -> "PagedKvLayout":
        kv_block_len = page_size * num_kv_heads * head_dim
        layer_stride = 2 * kv_block_len
        return PagedKvLayout(page_size, num_layers, num_kv_heads, head_dim, kv_block_len,
                             layer_stride, num_layers * layer_stride)

    def k_offset(self, layer: int) -> int:
        return layer * self.layer_stride

    def v_offset(self, layer: int) -> int:
        return layer * self.layer_stride + self.kv_block_len
