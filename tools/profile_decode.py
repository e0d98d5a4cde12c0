"""Small driver for ncu: Qwen3-4B, short prefill, a few fused decode steps WITHOUT CUDA graph capture
(plain launches, so every kernel is an individual ncu record).  Usage: profile_decode.py [ctx] [steps] [pdl]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pegainfer_b200.config import QWEN3_4B  # noqa: E402
from pegainfer_b200.model import ModelRuntimeConfig, Qwen3Model  # noqa: E402
from pegainfer_b200.synthetic import iter_random_weights, synthetic_prompt  # noqa: E402

ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pdl = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
m = Qwen3Model(QWEN3_4B, iter_random_weights(QWEN3_4B, 0, "cuda"),
               ModelRuntimeConfig(enable_cuda_graph=False, num_pages=ctx // 16 + 64, max_batch=1, enable_pdl=pdl))
kv = m.alloc_kv()
torch.cuda.synchronize()
torch.cuda.profiler.start()  # with `ncu --profile-from-start off` the weight generation stays out of the launch list
tok = m.sample_greedy(m.prefill([synthetic_prompt(ctx)], [kv])[0])
torch.cuda.synchronize()
for _ in range(steps):
    _, s = m.decode([tok], [kv], want_logits=False)
    tok = s[0]
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", tok)
