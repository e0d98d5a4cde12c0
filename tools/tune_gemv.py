"""GEMV ring-depth / CTAs-per-SM sweep on the real Qwen3-4B weight set (CUDA-event timed GEMV passes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pegainfer_b200 import ffi  # noqa: E402
from pegainfer_b200.config import QWEN3_4B  # noqa: E402
from pegainfer_b200.model import ModelRuntimeConfig, Qwen3Model  # noqa: E402
from pegainfer_b200.synthetic import iter_random_weights  # noqa: E402

m = Qwen3Model(QWEN3_4B, iter_random_weights(QWEN3_4B, 0, "cuda"),
               ModelRuntimeConfig(enable_cuda_graph=False, num_pages=64, max_batch=1, enable_pdl=True))
lib = ffi.lib()
W = 8044544000
for pdl in (1, 0):
    lib.pk_b200_set_pdl(pdl)
    for ctas in (1, 2):
        for kc in (1024, 2048, 4096):
            for stages in (2, 3, 4, 6, 8):
                lib.pk_b200_set_gemv_tuning(stages, ctas, kc)
                ms, n = m.bench_gemv_pass(10)
                print(f"pdl={pdl} ctas/SM={ctas} kc={kc} stages={stages}: {ms:.3f} ms/pass  {W / ms / 1e6:.0f} GB/s", flush=True)
