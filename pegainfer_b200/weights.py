"""Checkpoint loading for the Qwen3 path: ``config.json`` + safetensors, as the reference accepts them.

Mirror of ``pegainfer-qwen3-4b/src/config.rs:61-112`` (``Config::from_file`` with the ``generation_config.json`` stop
tokens) and ``pegainfer-core/src/weight_loader.rs:15-48`` (``load_shard_info``: a single ``model.safetensors`` wins,
otherwise ``model.safetensors.index.json`` maps tensor names to shard files).  Tensors are handed to the host mirror
one at a time (``Qwen3Model`` slices its tensor-parallel shard and uploads it, ``weights.rs:121-291``), so a rank never
holds more than one full tensor on the host; the files are memory-mapped by the ``safetensors`` package.
"""
from __future__ import annotations

import json
import os
from typing import Iterator

import torch

from .config import Qwen3Config
from .synthetic import weight_shapes

_CONFIG_KEYS = ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads",
                "head_dim", "vocab_size", "rms_norm_eps", "rope_theta", "tie_word_embeddings")


def load_config(model_path: str) -> tuple[Qwen3Config, list[int]]:
    """``(config, stop_token_ids)`` from ``<model_path>/config.json`` (+ ``generation_config.json`` when present).

    Every key the reference deserialises is required (serde would fail on a missing field, config.rs:23-37)."""
    with open(os.path.join(model_path, "config.json")) as f:
        raw = json.load(f)
    missing = [k for k in _CONFIG_KEYS + ("eos_token_id",) if k not in raw]
    if missing:
        raise KeyError(f"config.json misses {missing}")
    cfg = Qwen3Config(**{k: raw[k] for k in _CONFIG_KEYS}, name=os.path.basename(os.path.normpath(model_path)) or "custom")
    stop = [int(raw["eos_token_id"])]
    gen_path = os.path.join(model_path, "generation_config.json")
    if os.path.exists(gen_path):  # config.rs:97-111: int or list, consecutive duplicates dropped
        with open(gen_path) as f:
            eos = json.load(f)["eos_token_id"]
        ids = [int(eos)] if isinstance(eos, int) else [int(t) for t in eos]
        stop = [t for i, t in enumerate(ids) if i == 0 or t != ids[i - 1]]
    return cfg, stop


def load_shard_info(model_path: str) -> tuple[list[str], dict[str, int]]:
    """``(shard_files, tensor_name -> shard index)``; the map is empty for a single-file checkpoint."""
    single = os.path.join(model_path, "model.safetensors")
    if os.path.exists(single):
        return [single], {}
    with open(os.path.join(model_path, "model.safetensors.index.json")) as f:
        index = json.load(f)
    if not isinstance(index.get("weight_map"), dict):
        raise ValueError("Invalid index.json: missing weight_map")
    files: list[str] = []
    file_idx: dict[str, int] = {}
    weight_map: dict[str, int] = {}
    for name, shard in index["weight_map"].items():
        if shard not in file_idx:
            file_idx[shard] = len(files)
            files.append(os.path.join(model_path, shard))
        weight_map[name] = file_idx[shard]
    return files, weight_map


def iter_safetensors(model_path: str, cfg: Qwen3Config) -> Iterator[tuple[str, torch.Tensor]]:
    """Yield ``(hf_name, bf16 tensor)`` for every tensor the Qwen3 path reads, in the order of ``weight_shapes(cfg)``.

    Shapes and dtype are checked here (the reference asserts them while uploading, weight_loader.rs:130-206); a tied
    checkpoint needs no ``lm_head.weight`` (config.rs:68-74)."""
    from safetensors import safe_open
    files, weight_map = load_shard_info(model_path)
    handles = [safe_open(p, framework="pt", device="cpu") for p in files]
    try:
        names_in = [set(h.keys()) for h in handles]
        for name, shape in weight_shapes(cfg).items():
            if weight_map:
                if name not in weight_map:
                    raise KeyError(f"{name} is not in model.safetensors.index.json")
                si = weight_map[name]
            else:
                si = 0
            if name not in names_in[si]:
                raise KeyError(f"{name} is not in {files[si]}")
            t = handles[si].get_tensor(name)
            if t.dtype != torch.bfloat16:
                raise TypeError(f"{name}: dtype {t.dtype}, the path computes in bf16 (convert the checkpoint)")
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: shape {tuple(t.shape)} != {tuple(shape)} from config.json")
            yield name, t
    finally:
        handles.clear()
