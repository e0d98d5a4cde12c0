"""Generate HF-transformers golden vectors for the oracle (run in the build container).

The reference names HF ``AutoModelForCausalLM`` bf16 greedy as its external truth
(scripts/generate_test_data.py:41-52).  ``transformers`` cannot travel to the GPU
box, so its outputs on a tiny random-init Qwen3 are committed as a small fixture:

    python tests/golden/make_hf_fixtures.py   ->  tests/golden/hf_qwen3_tiny.npz

Contents: the config, the weight seed (weights are re-derived with
``pegainfer_b200.synthetic.random_weights``), the prompt, and HF's fp32 view of the
bf16 logits for the prompt's last token and 8 decode steps teacher-forced with the
synthetic continuation ids (random-init tied-embedding models greedy-repeat one token,
which would exercise nothing), plus HF's arg-max at each step.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from pegainfer_b200.config import QWEN3_TINY  # noqa: E402
from pegainfer_b200.synthetic import random_weights, synthetic_prompt  # noqa: E402


def main():
    from transformers import Qwen3Config, Qwen3ForCausalLM

    c = QWEN3_TINY
    seed, jitter, n_prompt, n_decode = 0, 0.1, 24, 8
    hf_cfg = Qwen3Config(
        hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
        num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
        num_key_value_heads=c.num_key_value_heads, head_dim=c.head_dim, vocab_size=c.vocab_size,
        rms_norm_eps=c.rms_norm_eps, rope_theta=c.rope_theta, tie_word_embeddings=True,
        max_position_embeddings=4096, attention_bias=False, use_sliding_window=False)
    hf_cfg._attn_implementation = "eager"
    model = Qwen3ForCausalLM(hf_cfg).to(torch.bfloat16).eval()
    w = random_weights(c, seed=seed, norm_jitter=jitter)
    sd = {k: v for k, v in w.items()}
    sd["lm_head.weight"] = w["model.embed_tokens.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in m or "inv_freq" in m for m in missing), missing

    full = [t % c.vocab_size for t in synthetic_prompt(n_prompt + n_decode)]
    prompt, forced = full[:n_prompt], full[n_prompt:]
    logits, tokens = [], []
    with torch.no_grad():
        out = model(torch.tensor([prompt]), use_cache=True)
        past = out.past_key_values
        lg = out.logits[0, -1]
        for step in range(n_decode + 1):
            logits.append(lg.float().numpy().copy())
            tokens.append(int(torch.argmax(lg)))
            if step == n_decode:
                break
            out = model(torch.tensor([[forced[step]]]), past_key_values=past, use_cache=True)
            past = out.past_key_values
            lg = out.logits[0, -1]
    path = os.path.join(os.path.dirname(__file__), "hf_qwen3_tiny.npz")
    np.savez_compressed(path, logits=np.stack(logits), tokens=np.array(tokens, np.int32),
                        prompt=np.array(prompt, np.int32), forced=np.array(forced, np.int32), seed=seed, norm_jitter=jitter,
                        transformers_version=np.array(__import__("transformers").__version__))
    print("wrote", path, np.stack(logits).shape, tokens)


if __name__ == "__main__":
    main()
