"""Pin the CPU oracle against every weight-free golden vector the reference holds
for the hot path and against the committed HF fixture.  CPU only."""
import math
import os

import numpy as np
import pytest

from oracle import qwen3_oracle as O
from pegainfer_b200.config import QWEN3_TINY
from pegainfer_b200.synthetic import random_weights, to_numpy_bits
from tests.helpers import bf16_bits_of, f32, oracle_cfg

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rms_norm_reference(x_bits, w_bits, eps, offset=False):
    """pegainfer-server/src/ops/tests.rs:12-34 restated (the reference's CPU checker)."""
    x, w = f32(x_bits), f32(w_bits)
    inv = np.float32(1.0) / np.sqrt(np.float32((x * x).sum(dtype=np.float32) / np.float32(len(x)) + np.float32(eps)))
    normed = f32(O.f32_to_bf16(x * inv))
    scale = (1.0 + w) if offset else w
    return f32(O.f32_to_bf16(normed * scale))


def test_gemv_known_answer():  # ops/tests.rs:50-77
    a = bf16_bits_of([1, 2, 3, 4, 5, 6]).reshape(2, 3)
    x = bf16_bits_of([1, 2, 3]).reshape(1, 3)
    y = f32(O.gemm(a, x))[0]
    assert abs(y[0] - 14.0) < 0.1 and abs(y[1] - 32.0) < 0.1


def test_argmax_known_answer():  # ops/tests.rs:79-86
    assert O.argmax(bf16_bits_of([1.0, 9.0, 3.0, 8.0])) == 1


def test_argmax_lowest_index_wins_ties():  # csrc/argmax.cu:18
    assert O.argmax(bf16_bits_of([1.0, 9.0, 9.0, 8.0])) == 1


def test_rms_norm_known_answer():  # ops/tests.rs:88-103, tol 0.01
    x, w = bf16_bits_of([1, 2, 3, 4]), bf16_bits_of([1, 1, 1, 1])
    got = f32(O.rms_norm(x, w, 1e-6))[0]
    assert np.abs(got - rms_norm_reference(x, w, 1e-6)).max() <= 0.01


def test_rms_norm_batch_multi_tile():  # ops/tests.rs:105-152, hidden 260 x 2 rows, tol 0.02
    hidden, seq = 260, 2
    x = bf16_bits_of([((i % 17) - 8.0) * 0.25 for i in range(hidden * seq)]).reshape(seq, hidden)
    w = bf16_bits_of([0.5 + (i % 11) * 0.0625 for i in range(hidden)])
    got = f32(O.rms_norm(x, w, 1e-6))
    for r in range(seq):
        assert np.abs(got[r] - rms_norm_reference(x[r], w, 1e-6)).max() <= 0.02


def test_embedding_variants():  # ops/tests.rs:172-226
    embed = bf16_bits_of(np.arange(1, 13)).reshape(3, 4)
    dec = f32(O.embedding_batched(embed, [1], 4))[0]
    assert abs(dec[0] - 5.0) < 0.01 and abs(dec[3] - 8.0) < 0.01
    b = f32(O.embedding_batched(embed, [2, 0], 4)).ravel()
    assert [b[0], b[3], b[4], b[7]] == [9.0, 12.0, 1.0, 4.0]


def test_embedding_vocab_shard_masks_non_local():  # pegainfer-kernels/src/ops/embedding.rs:99-128
    embed = bf16_bits_of([10, 11, 12, 20, 21, 22]).reshape(2, 3)
    out = f32(O.embedding_batched_vocab_shard(embed, [4, 5, 1, 4], 3, 4, 2)).ravel()
    assert out.tolist() == [10, 11, 12, 20, 21, 22, 0, 0, 0, 10, 11, 12]


def test_kv_layout_stride_geometry():  # pegainfer-core/src/kv_pool.rs:290-310
    from pegainfer_b200.paged_kv import PagedKvLayout
    l35 = PagedKvLayout.new(8, 4, 256, 16)
    assert (l35.kv_block_len, l35.layer_stride, l35.page_stride) == (16384, 32768, 262144)
    l4b = PagedKvLayout.new(36, 4, 128, 16)  # page_stride == 589_824
    assert l4b.page_stride == 589_824
    assert PagedKvLayout.new(36, 8, 128, 16).page_stride == 1_179_648  # 2.25 MiB / page


def test_rope_table_layout():  # weight_loader.rs:210-244: half-split duplicated layout
    cos, sin = O.precompute_rope(128, 8, 1e6)
    cos, sin = f32(cos).reshape(8, 128), f32(sin).reshape(8, 128)
    assert (cos[:, :64] == cos[:, 64:]).all() and (sin[:, :64] == sin[:, 64:]).all()
    assert (cos[0] == 1).all() and (sin[0] == 0).all()
    assert abs(cos[3, 0] - math.cos(3.0)) < 4e-3 and abs(sin[3, 0] - math.sin(3.0)) < 4e-3


def test_silu_rounding_points_differ():  # fused_proj.cu:44-63 vs elementwise.cu:27-42
    g, u = bf16_bits_of([0.3, -1.7, 2.9, 0.011]), bf16_bits_of([1.3, 0.7, -0.9, 3.0])
    fused = O.silu_mul_fused(np.concatenate([g, u]).reshape(1, 8), 4)[0]
    gf, uf = f32(g), f32(u)
    want = O.f32_to_bf16(gf / (1 + np.exp(-gf, dtype=np.float32)) * uf)
    assert np.abs(f32(fused) - f32(want)).max() <= 2 ** -8 * np.abs(f32(want)).max()
    assert O.silu_mul(g, u).shape == g.shape


def test_fused_add_norm_uses_unrounded_sum():  # norm.cuh:419-424,467
    h = bf16_bits_of([1.0, 256.0]).reshape(1, 2).copy()
    r = bf16_bits_of([0.00390625 * 0.75, 0.75]).reshape(1, 2)  # sums are not bf16-representable
    w = bf16_bits_of([1.0, 1.0])
    out = f32(O.fused_add_rms_norm(h, r, w, 0.0))[0]
    x = np.array([1.0 + 0.00390625 * 0.75, 256.75], np.float32)
    want = x / np.sqrt((x * x).mean())
    assert np.abs(out - want).max() <= 2 ** -8
    assert f32(h).ravel().tolist() == f32(O.f32_to_bf16(x)).tolist()


def test_oracle_matches_hf_fixture():
    """HF transformers is the reference's declared external truth
    (scripts/generate_test_data.py:41-52).  HF rounds the normalised value before the
    weight multiply and uses unfused adds, so this is a tolerance check: every logit
    within 4 bf16 ulp of the row's max magnitude, arg-max identical on all 9 steps,
    at TP1 and TP2."""
    g = np.load(os.path.join(GOLD, "hf_qwen3_tiny.npz"))
    c = QWEN3_TINY
    w = to_numpy_bits(random_weights(c, seed=int(g["seed"]), norm_jitter=float(g["norm_jitter"])))
    for world in (1, 2):
        m = O.OracleQwen3(oracle_cfg(c), w, tp_world=world, num_pages=16)
        kv = m.alloc_kv()
        lg = [f32(m.prefill([g["prompt"].tolist()], [kv])[0])]
        for t in g["forced"]:
            lg.append(f32(m.decode([int(t)], [kv])[0]))
        lg, ref = np.stack(lg), g["logits"]
        tol = 4 * O.bf16_ulp(np.abs(ref).max(axis=1))
        assert (np.abs(lg - ref).max(axis=1) <= tol).all(), np.abs(lg - ref).max(axis=1) / tol
        assert (lg.argmax(1) == g["tokens"]).all()
