"""The C-ABI library builds, loads and exports every symbol include/pegainfer_kernels.h declares,
and the ctypes table in pegainfer_b200/ffi.py covers the same set.  No compute (no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pegainfer_kernels.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\(", src)
    return sorted({n for n in names if n not in ("defined",)})


@pytest.fixture(scope="module")
def built():
    from pegainfer_b200 import build
    build.build()
    return build.KERNEL_LIB


def test_header_declares_reference_ffi_subset():
    want = {"rms_norm_cuda", "rms_norm_batched_cuda", "fused_add_rms_norm_cuda",
            "fused_add_rms_norm_batched_cuda", "add_cuda", "embedding_batched_cuda",
            "embedding_decode_cuda", "embedding_batched_vocab_shard_cuda", "silu_mul_triton_aot_cuda",
            "silu_mul_fused_cuda", "gemm_cuda", "gemm_graphsafe_cuda", "argmax_cuda",
            "flashinfer_top1_cuda", "gpu_sample_flashinfer_cuda", "prefill_qk_norm_rope_only_cuda", "qk_norm_rope_batched_decode_cuda",
            "cublas_init", "cublas_destroy", "cuda_set_device", "paged_kv_scatter_cuda",
            "batch_prefill_paged_num_tiles", "batch_prefill_paged_num_tiles_with_cta_tile_q",
            "batch_prefill_cta_tile_q", "batch_prefill_cta_tile_q_with_override",
            "batch_prefill_paged_cuda", "batch_prefill_paged_cuda_with_cta_tile_q", "single_prefill_cuda",
            "paged_attention_decode_cuda", "paged_attention_decode_split_kv_cuda"}
    assert want <= set(declared_symbols())


def test_library_exports_every_declared_symbol(built):
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_table_matches_header(built):
    from pegainfer_b200 import ffi
    table = set(ffi.SIGNATURES) | set(ffi.EXT_SIGNATURES) | set(ffi.QWEN35_SIGNATURES)
    assert set(declared_symbols()) == table
    lib = ffi.load(built)  # dlopen + every symbol typed; no launches
    assert b"sm_100a" in lib.pk_b200_version()


def test_sass_is_blackwell_native(built):
    """tcgen05 / TMA evidence in the shipped SASS (B200_PROFILING.md table)."""
    sass = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass, "no tcgen05.mma in SASS"
    assert "UTMALDG" in sass, "no TMA tensor load in SASS"
    assert "UBLKCP" in sass, "no cp.async.bulk in SASS"
    assert "LDTM" in sass, "no tcgen05.ld in SASS"


def test_prefill_gemm_plan_matches_the_measured_calibration():
    """pk_b200_gemm_plan (host arithmetic of gemm2.cu, no CUDA call) on a 148-SM GPU: the choices that were measured on B200
    (profiles/README.md round 2, profiles/r2_v3_prefill_ops.txt).  o_proj / down_proj at 2048 tokens must take the pair
    kernel's one-wave 320-wide tile (46 / 89 us; the single-CTA kernel needs 81 / 173 us) -- a cost-model slip that sent
    them back to the single-CTA kernel cost 1.4 ms of TTFT(2048) until it was calibrated."""
    from pegainfer_b200 import ffi
    plan = ffi.lib().pk_b200_gemm_plan
    H, I, QKV, Q = 2560, 9728, 6144, 4096  # Qwen3-4B
    assert plan(Q, 2048, H, 0, 148) == 256 and plan(QKV, 2048, H, 0, 148) == 256   # q / fused qkv
    assert plan(H, 2048, Q, 0, 148) == 320 and plan(H, 2048, I, 0, 148) == 320       # o_proj, down_proj
    assert plan(I, 2048, H, 1, 148) == 256                                            # gate_up + SwiGLU: 256-wide only
    assert plan(2 * I, 2048, H, 0, 148) == 256                                        # plain gate_up: never the single-buffered 320
    for m, k in ((Q, H), (H, Q), (H, I), (2 * I, H)):
        assert plan(m, 128, k, 0, 148) == -2                                          # one token tile: single-CTA kernel (split-K)
    assert plan(I, 128, H, 1, 148) == -2 and plan(Q, 2048, H + 4, 0, 148) == -2       # SwiGLU at 128 tokens, K % 8 != 0
    assert plan(4096, 2048, 4096, 0, 148) in (256, 320, 128)                          # Qwen3-8B q: some pair tile
