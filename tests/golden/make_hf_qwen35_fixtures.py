"""Generates tests/golden/hf_qwen35_tiny.npz: a tiny random-init Qwen3.5 hybrid model (3 gated-delta-net layers + 1 gated
full-attention layer) from HF transformers -- the reference's declared external truth (scripts/generate_test_data.py) --
with its bf16 and fp32 logits on a fixed prompt, plus a few fp32 steps of HF's `torch_recurrent_gated_delta_rule`.
Run in the build container (transformers >= 5.5, CPU):  python tests/golden/make_hf_qwen35_fixtures.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from transformers.models.qwen3_5 import modeling_qwen3_5 as m  # noqa: E402
from transformers.models.qwen3_5.configuration_qwen3_5 import Qwen3_5TextConfig  # noqa: E402

LT = ["linear_attention", "linear_attention", "linear_attention", "full_attention"]
CFG = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=64,
           vocab_size=512, linear_conv_kernel_dim=4, linear_key_head_dim=32, linear_value_head_dim=32, linear_num_key_heads=2,
           linear_num_value_heads=4)
THETA, PRF = 10000.0, 0.25


def main():
    hf = Qwen3_5TextConfig(**CFG, layer_types=LT, tie_word_embeddings=True,
                           rope_parameters={"rope_theta": THETA, "partial_rotary_factor": PRF, "rope_type": "default"})
    torch.manual_seed(0)
    model = m.Qwen3_5ForCausalLM(hf)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("layernorm.weight") or n.endswith("model.norm.weight") or "q_norm" in n or "k_norm" in n:
                p.add_(torch.randn_like(p) * 0.1)
            elif p.dim() >= 2:
                p.mul_(4.0)  # livelier activations than the std-0.02 init
    model = model.to(torch.bfloat16).eval()
    sd = model.state_dict()
    toks = [(7 * i + 3) % CFG["vocab_size"] for i in range(14)]
    with torch.no_grad():
        lg_bf16 = model(torch.tensor([toks])).logits[0].float().numpy()
    m32 = m.Qwen3_5ForCausalLM(hf)
    m32.load_state_dict({k: v.float() for k, v in sd.items()})
    with torch.no_grad():
        lg_f32 = m32.float().eval()(torch.tensor([toks])).logits[0].numpy()
    out = {"tokens": np.array(toks, np.int32), "logits_bf16": lg_bf16, "logits_f32": lg_f32,
           "layer_types": np.array(LT), "theta": np.float32(THETA), "partial_rotary_factor": np.float32(PRF),
           "cfg_keys": np.array(list(CFG)), "cfg_vals": np.array(list(CFG.values()), np.int64)}
    for k, v in sd.items():
        if k != "lm_head.weight":
            out["w:" + k] = v.view(torch.int16).numpy().view(np.uint16)  # every tensor is bf16 in the checkpoint
    # a few fp32 steps of HF's recurrent gated delta rule (B=1, T=3, heads repeated to nv as HF does)
    g = torch.Generator().manual_seed(1)
    nk, nv, dk, dv, T = 2, 4, 32, 32, 3
    q = torch.randn((1, T, nk, dk), generator=g)
    k = torch.randn((1, T, nk, dk), generator=g)
    v = torch.randn((1, T, nv, dv), generator=g)
    a = torch.randn((1, T, nv), generator=g)
    b = torch.randn((1, T, nv), generator=g)
    dt_bias, a_log = torch.randn(nv, generator=g) * 0.5, torch.randn(nv, generator=g) * 0.5
    beta = b.sigmoid()
    gdec = -a_log.exp() * torch.nn.functional.softplus(a + dt_bias)
    core, _ = m.torch_recurrent_gated_delta_rule(q.repeat_interleave(nv // nk, dim=2), k.repeat_interleave(nv // nk, dim=2), v, g=gdec,
                                                 beta=beta, initial_state=None, output_final_state=True, use_qk_l2norm_in_kernel=True)
    out.update({"gdr_q": q[0].numpy(), "gdr_k": k[0].numpy(), "gdr_v": v[0].numpy(), "gdr_a": a[0].numpy(), "gdr_b": b[0].numpy(),
                "gdr_dt_bias": dt_bias.numpy(), "gdr_a_log": a_log.numpy(), "gdr_out": core[0].float().numpy()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hf_qwen35_tiny.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
