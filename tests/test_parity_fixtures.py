"""The committed full-size oracle fixtures (tests/golden/parity_*.npz) are self-consistent and belong to the checkpoint
the generator produces: what the GPU parity tests and bench.py's `parity` key compare against."""
import glob
import os

import numpy as np
import pytest

from pegainfer_b200.config import PRESETS
from tests.golden import parity_fixture as F

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(F.__file__), "parity_*.npz")))


def _row_from_fixture(fx, step, vocab):
    """A full logits row that agrees with the fixture on every sampled position (elsewhere: far below the top-256)."""
    floor = F._f32(fx.val_top[step]).min() - 4.0
    row = np.full(vocab, floor, np.float32)
    row[::F.STRIDE] = np.minimum(F._f32(fx.val_str[step]), floor)  # strided sample, kept below the top-256 unless listed
    row[::F.STRIDE] = F._f32(fx.val_str[step])
    row[fx.idx_top[step]] = F._f32(fx.val_top[step])
    bits = (row.view(np.uint32) >> 16).astype(np.uint16)
    return bits


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_fixture_is_well_formed_and_accepts_itself(path):
    fx = F.Fixture(path)
    cfg = PRESETS[fx.meta["model"]]
    assert fx.steps == len(fx.tokens) + 1 >= 2  # prefill + one row per teacher-forced token
    assert fx.idx_top.shape == (fx.steps, F.TOPK) and fx.val_str.shape[1] == (cfg.vocab_size + F.STRIDE - 1) // F.STRIDE
    assert all(0 <= t < cfg.vocab_size for t in fx.tokens)
    for step in range(fx.steps):
        top = F._f32(fx.val_top[step])
        assert (np.diff(top) <= 0).all()  # sorted, largest first
        assert abs(float(fx.margin[step]) - float(top[0] - top[1])) < 1e-6
        assert float(fx.rowmax[step]) >= float(np.abs(top).max())
        if step < len(fx.tokens):
            assert fx.tokens[step] == fx.oracle_argmax(step)  # teacher forcing used the oracle's own greedy tokens
        bits = _row_from_fixture(fx, step, cfg.vocab_size)
        ok, info = fx.compare(step, bits, 0.0)
        assert ok and info["err"] == 0.0 and info["same_argmax"], info
        # a logit moved by more than the tolerance is rejected
        bad = bits.copy()
        i = int(fx.idx_top[step][5])
        bad[i] = (np.float32(F._f32(bad[i:i + 1])[0] + 64 * F.bf16_ulp(fx.rowmax[step])).view(np.uint32) >> 16).astype(np.uint16)
        assert not fx.compare(step, bad, 12.0)[0]


def test_fixture_checkpoint_is_the_generators():
    """The CRC stored in the Qwen3-4B fixtures is the CRC of the seed-0 CPU checkpoint as generated today (per-block seeded,
    independent of the worker count): the GPU tests and bench.py refuse a checkpoint whose CRC differs."""
    from pegainfer_b200.synthetic import iter_random_weights
    fx = F.Fixture(F.fixture_path("qwen3-4b", 128, 1))
    wanted = {"model.embed_tokens.weight", "model.layers.0.self_attn.q_proj.weight", "model.layers.1.mlp.down_proj.weight",
              "model.norm.weight"}
    got = {}
    for name, t in iter_random_weights(PRESETS["qwen3-4b"], seed=0, device="cpu"):
        if name in wanted:
            got[name] = t
        if len(got) == len(wanted):
            break
    assert F.torch_weights_crc(got) == fx.meta["weights_crc"] == F.Fixture(F.fixture_path("qwen3-4b", 2048, 1)).meta["weights_crc"]
