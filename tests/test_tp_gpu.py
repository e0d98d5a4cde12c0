"""Tensor-parallel parity on real GPUs (SURVEY 8 row a15 / BASELINE config 3): runs tests/tools/tp_check.py under
torchrun when the box has >= 2 GPUs (`gpurun --gpus N`); skipped on a single-GPU box.

* all-reduce kernels (plain, fused with add + RMSNorm) vs torch / the CPU oracle, bit-exact rank-ordered sums
* Qwen3-small TP-N and the FULL Qwen3-8B TP-N (36 layers, V = 151,936) prefill + decode logits, teacher-forced
  against the CPU oracle's TP-N model on the seed-0 CPU checkpoint (SURVEY 8c rule).
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(world, extra, port, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "tools", "tp_check.py"), *extra]
    env = {k: v for k, v in os.environ.items() if k not in ("OMP_PROC_BIND", "OMP_PLACES")}  # ranks must not share CPU 0
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "TP_CHECK PASS" in r.stdout, tail
    return r.stdout


@pytest.mark.skipif(NGPU < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [w for w in (2, 4, 8) if w <= max(NGPU, 2)])
def test_tp_small_model_and_collectives(world):
    if world > NGPU:
        pytest.skip(f"{world} GPUs not present")
    _torchrun(world, ["--model", "qwen3-small"], 29520 + world, 600)


@pytest.mark.skipif(NGPU < 2 or os.environ.get("PK_SKIP_FULLSIZE") == "1", reason="needs >= 2 GPUs (full-size run)")
@pytest.mark.parametrize("world", [w for w in (2, 8) if w <= max(NGPU, 2)])
def test_tp_qwen3_8b_full_size(world):
    if world > NGPU:
        pytest.skip(f"{world} GPUs not present")
    out = _torchrun(world, ["--model", "qwen3-8b", "--prompt", "128", "--steps", "8"], 29540 + world, 1800)
    print(out[-600:])
