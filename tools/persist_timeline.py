"""Phase timeline of the persistent decode kernel (CTA 0 globaltimer stamps, layers 0..3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pegainfer_b200.config import QWEN3_4B  # noqa: E402
from pegainfer_b200.model import ModelRuntimeConfig, Qwen3Model  # noqa: E402
from pegainfer_b200.synthetic import iter_random_weights, synthetic_prompt  # noqa: E402

ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = Qwen3Model(QWEN3_4B, iter_random_weights(QWEN3_4B, 0, "cuda"),
               ModelRuntimeConfig(enable_cuda_graph=True, num_pages=ctx // 16 + 64, max_batch=1, persistent=True))
kv = m.alloc_kv()
tok = m.sample_greedy(m.prefill([synthetic_prompt(ctx)], [kv])[0])
for _ in range(5):
    _, s = m.decode([tok], [kv], want_logits=False)
    tok = s[0]
d = m.debug_buffer("persist_dbg", 4 * 16 * 4).view(torch.int64).cpu().tolist()
names = ["start", "x_norm", "qkv", "bar", "attn", "bar", "x_o", "o", "bar", "x_norm2", "gate_up", "bar", "x_down", "down", "bar"]
for li in range(4):
    row = d[li * 16:li * 16 + 15]
    print(f"layer {li}: total {(row[14] - row[0]) / 1000:.1f} us :: " +
          " ".join(f"{names[i]}={(row[i] - row[i - 1]) / 1000:.1f}" for i in range(1, 15)))
