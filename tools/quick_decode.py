"""Quick decode-step timing for tuning sweeps: Qwen3-4B, 2048-token prompt, `--steps` device-resident decode steps.

Prints one line `ms_per_step tok_s first_tokens`; knobs come from the environment (PK_PF_O, PK_PF_GU, PK_PF_Y,
PK_ATTN, ...), which the libraries read once per process -- run one process per setting.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pegainfer_b200.config import PRESETS, TensorParallelConfig  # noqa: E402
from pegainfer_b200.model import ModelRuntimeConfig, Qwen3Model  # noqa: E402
from pegainfer_b200.synthetic import iter_random_weights, synthetic_prompt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="qwen3-4b")
ap.add_argument("--prompt", type=int, default=2048)
ap.add_argument("--steps", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--lib", default=None, help="alternative kernel library (tools/build_variant.sh)")
a = ap.parse_args()
cfg = PRESETS[a.model]
pages = 3 * ((a.prompt + a.reps * a.steps + 128) // 16 + 2) + 8
rt = ModelRuntimeConfig(enable_cuda_graph=True, tensor_parallel=TensorParallelConfig(0, 1), device_ordinal=0, fused=True,
                        num_pages=pages, max_batch=1, enable_pdl=True,
                        kernel_lib=os.path.abspath(a.lib) if a.lib else None)
model = Qwen3Model(cfg, iter_random_weights(cfg, seed=0, device="cuda"), rt)
prompt = synthetic_prompt(a.prompt)
model.generate(prompt, 4)
kv = model.alloc_kv()
tok = model.sample_greedy(model.prefill([prompt], [kv])[0])
toks, _ = model.decode_burst(kv, tok, 8)
best, first = 1e9, list(toks)
for _ in range(a.reps):
    torch.cuda.synchronize()
    out, ms = model.decode_burst(kv, toks[-1], a.steps)
    toks = out
    best = min(best, ms / a.steps)
env = {k: v for k, v in os.environ.items() if k.startswith("PK_")}
print(f"QUICK {best:.4f} ms/step {1e3 / best:.1f} tok/s first={first[:6]} lib={a.lib} env={env}", flush=True)
