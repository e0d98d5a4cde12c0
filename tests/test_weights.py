"""Checkpoint loading (pegainfer_b200/weights.py) against the layouts the reference accepts
(pegainfer-core/src/weight_loader.rs:15-48, pegainfer-qwen3-4b/src/config.rs:61-112)."""
import json
import os

import pytest
import torch
from safetensors.torch import save_file

from pegainfer_b200.config import PRESETS
from pegainfer_b200.synthetic import random_weights
from pegainfer_b200.weights import iter_safetensors, load_config, load_shard_info

CFG = PRESETS["qwen3-tiny"]


def write_config(d, cfg, **extra):
    raw = {k: getattr(cfg, k) for k in ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
                                       "num_key_value_heads", "head_dim", "vocab_size", "rms_norm_eps", "rope_theta",
                                       "tie_word_embeddings")}
    raw.update({"eos_token_id": 7, "model_type": "qwen3", "torch_dtype": "bfloat16"}, **extra)
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(raw, f)


def test_single_file_round_trip(tmp_path):
    w = random_weights(CFG, seed=3)
    d = str(tmp_path)
    write_config(d, CFG)
    save_file({k: v.contiguous() for k, v in w.items()}, os.path.join(d, "model.safetensors"))
    cfg, stop = load_config(d)
    assert cfg.to_dict() == {**CFG.to_dict(), "name": cfg.name} and stop == [7]
    files, wmap = load_shard_info(d)
    assert len(files) == 1 and wmap == {}
    got = dict(iter_safetensors(d, cfg))
    assert set(got) == set(w)
    for k in w:
        assert got[k].dtype == torch.bfloat16 and torch.equal(got[k].view(torch.int16), w[k].view(torch.int16)), k


def test_sharded_index_and_generation_config(tmp_path):
    w = random_weights(CFG, seed=4)
    d = str(tmp_path)
    write_config(d, CFG)
    names = list(w)
    shards = {"model-00001-of-00002.safetensors": names[::2], "model-00002-of-00002.safetensors": names[1::2]}
    weight_map = {}
    for fn, ns in shards.items():
        save_file({n: w[n].contiguous() for n in ns}, os.path.join(d, fn))
        weight_map.update({n: fn for n in ns})
    with open(os.path.join(d, "model.safetensors.index.json"), "w") as f:
        json.dump({"metadata": {}, "weight_map": weight_map}, f)
    with open(os.path.join(d, "generation_config.json"), "w") as f:
        json.dump({"eos_token_id": [5, 5, 9]}, f)
    cfg, stop = load_config(d)
    assert stop == [5, 9]  # consecutive duplicates dropped, as Vec::dedup
    files, wmap = load_shard_info(d)
    assert len(files) == 2 and set(wmap) == set(names)
    got = dict(iter_safetensors(d, cfg))
    assert all(torch.equal(got[k].view(torch.int16), w[k].view(torch.int16)) for k in w)


def test_rejects_wrong_dtype_shape_and_missing(tmp_path):
    w = random_weights(CFG, seed=5)
    d = str(tmp_path)
    write_config(d, CFG)
    bad = {k: v.contiguous() for k, v in w.items()}
    bad["model.norm.weight"] = bad["model.norm.weight"].float()
    save_file(bad, os.path.join(d, "model.safetensors"))
    cfg, _ = load_config(d)
    with pytest.raises(TypeError):
        dict(iter_safetensors(d, cfg))
    bad["model.norm.weight"] = w["model.norm.weight"][:-1].contiguous()
    save_file(bad, os.path.join(d, "model.safetensors"))
    with pytest.raises(ValueError):
        dict(iter_safetensors(d, cfg))
    del bad["model.norm.weight"]
    save_file(bad, os.path.join(d, "model.safetensors"))
    with pytest.raises(KeyError):
        dict(iter_safetensors(d, cfg))
    os.remove(os.path.join(d, "config.json"))
    write_config(d, CFG)
    raw = json.load(open(os.path.join(d, "config.json")))
    del raw["head_dim"]
    json.dump(raw, open(os.path.join(d, "config.json"), "w"))
    with pytest.raises(KeyError):
        load_config(d)


@pytest.mark.gpu
def test_model_from_safetensors_matches_oracle(tmp_path):
    """weights.rs:83-125: the checkpoint on disk drives the GPU model; logits against the oracle on the same weights."""
    from oracle import qwen3_oracle as O
    from pegainfer_b200.model import ModelRuntimeConfig, Qwen3Model
    from pegainfer_b200.synthetic import synthetic_prompt, to_numpy_bits
    from tests.helpers import bits, logits_agree, oracle_cfg
    w = random_weights(CFG, seed=0, norm_jitter=0.1)
    d = str(tmp_path)
    write_config(d, CFG)
    save_file({k: v.contiguous() for k, v in w.items()}, os.path.join(d, "model.safetensors"))
    m = Qwen3Model.from_safetensors(d, ModelRuntimeConfig(num_pages=64))
    assert m.stop_token_ids == [7]
    orc = O.OracleQwen3(oracle_cfg(CFG), to_numpy_bits(w), tp_world=1, num_pages=64)
    prompt = [t % CFG.vocab_size for t in synthetic_prompt(24)]
    okv, kv = orc.alloc_kv(), m.alloc_kv()
    want = orc.prefill([prompt], [okv])[0]
    got = bits(m.prefill([prompt], [kv])[0])
    ok, info = logits_agree(got, want, 6)
    m.drop_request(kv)
    m.close()
    assert ok, info


def test_token_logprobs_match_log_softmax():
    """compute_logprobs_from_cpu (executor.rs:400-436): host-side log-softmax of a bf16 logits row, top-k descending."""
    import numpy as np
    import torch
    from pegainfer_b200.model import token_logprobs
    g = torch.Generator().manual_seed(3)
    row = (torch.randn(5000, generator=g) * 3).to(torch.bfloat16)
    row[17] = row[4000] = row.float().max() + 1  # a tie at the top: lower index first
    ref = torch.log_softmax(row.float(), dim=0)
    lp, top = token_logprobs(row, 123, 5)
    assert abs(lp - float(ref[123])) < 1e-4
    assert [i for i, _ in top[:2]] == [17, 4000]
    order = sorted(range(5000), key=lambda i: (-float(row[i]), i))[:5]
    assert [i for i, _ in top] == order
    assert np.allclose([v for _, v in top], [float(ref[i]) for i in order], atol=1e-4)
    assert token_logprobs(row, 0, 0)[1] == []
