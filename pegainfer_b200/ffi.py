"""ctypes binding of ``libpegainfer_kernels_b200.so`` -- the Python twin of the reference's
``pegainfer-kernels/src/ffi.rs`` (same symbol names, same argument order).

There is no fallback of any kind: if the library is missing or cannot be loaded this module
raises, and every launch needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
KERNEL_LIB_PATH = os.path.join(_HERE, "libpegainfer_kernels_b200.so")

vp, i32, i64, f32, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32


class GemvArgs(C.Structure):
    """pk_b200_gemv_args (include/pegainfer_kernels.h)."""
    _fields_ = [("W", vp), ("X", vp), ("Y", vp * 3), ("seg_rows", i32 * 3), ("M", i32), ("N", i32),
                ("K", i32), ("x_mode", i32), ("residual", vp), ("norm_w", vp), ("eps", f32),
                ("hidden_out", vp), ("normed_out", vp), ("epi", i32), ("tp_comm", vp), ("tp_step", vp), ("tp_op", i32)]


class PrefetchSpan(C.Structure):
    """pk_b200_prefetch_span (include/pegainfer_kernels.h)."""
    _fields_ = [("base", vp), ("rows", i32), ("row_bytes", i32), ("slices", i32), ("prefetch_rows", i32)]


# name -> (restype, argtypes); order follows include/pegainfer_kernels.h
SIGNATURES = {
    "cuda_set_device": (i32, [i32]),
    "cublas_init": (None, []),
    "cublas_destroy": (None, []),
    "embedding_batched_cuda": (i32, [vp, vp, vp, i32, i32, vp]),
    "embedding_decode_cuda": (i32, [vp, vp, vp, i32, vp]),
    "embedding_batched_vocab_shard_cuda": (i32, [vp, vp, vp, i32, i32, u32, u32, vp]),
    "rms_norm_cuda": (None, [vp, vp, vp, i32, f32, vp]),
    "rms_norm_batched_cuda": (None, [vp, vp, vp, i32, i32, f32, vp]),
    "fused_add_rms_norm_cuda": (None, [vp, vp, vp, vp, i32, f32, vp]),
    "fused_add_rms_norm_batched_cuda": (None, [vp, vp, vp, vp, i32, i32, f32, vp]),
    "add_cuda": (i32, [vp, vp, vp, i32, vp]),
    "silu_mul_triton_aot_cuda": (i32, [vp, vp, vp, i32, vp]),
    "silu_mul_fused_cuda": (None, [vp, vp, i32, i32, vp]),
    "gemm_cuda": (None, [vp, vp, vp, i32, i32, i32, vp]),
    "gemm_graphsafe_cuda": (None, [vp, vp, vp, i32, i32, i32, vp]),
    "prefill_qk_norm_rope_only_cuda": (None, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp]),
    "qk_norm_rope_batched_decode_cuda": (None, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp]),
    "paged_kv_scatter_cuda": (i32, [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, i64, i64, vp]),
    "batch_prefill_paged_num_tiles": (i32, [i32, i32, i32, i32]),
    "batch_prefill_paged_num_tiles_with_cta_tile_q": (i32, [i32, i32, i32, i32, i32]),
    "batch_prefill_cta_tile_q": (i32, [i32, i32, i32, i32]),
    "batch_prefill_cta_tile_q_with_override": (i32, [i32, i32, i32, i32, i32]),
    "batch_prefill_paged_cuda": (i32, [vp, vp, vp, i64, i64] + [vp] * 9 + [i32] * 7 + [i64, f32, vp]),
    "batch_prefill_paged_cuda_with_cta_tile_q": (i32, [vp, vp, vp, i64, i64] + [vp] * 9 + [i32] * 7 + [i64, f32, i32, vp]),
    "single_prefill_cuda": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, vp]),
    "paged_attention_decode_cuda": (i32, [vp, vp, vp, i64, i64] + [vp] * 6 + [i32] * 5 + [i64, f32, vp]),
    "paged_attention_decode_split_kv_cuda": (i32, [vp, vp, vp, i64, i64] + [vp] * 10 + [i32] * 6 + [i64, f32, vp]),
    "argmax_cuda": (None, [vp, vp, i32, vp]),
    "flashinfer_top1_cuda": (None, [vp, vp, vp, vp, i32, vp]),
    "gpu_sample_flashinfer_cuda": (None, [vp, vp, vp, vp, i32, f32, i32, f32, C.c_uint64, vp]),
}

# Qwen3.5 hybrid-layer ops (ffi.rs:181-226,981-1039): present in our library and in the reference's full build, but not
# in the Qwen3-only oracle/_ref library -> typed only when the symbol exists
QWEN35_SIGNATURES = {
    "rms_norm_batched_offset_cuda": (None, [vp, vp, vp, i32, i32, f32, vp]),
    "rms_norm_offset_cuda": (None, [vp, vp, vp, i32, f32, vp]),
    "rms_norm_gated_cuda": (None, [vp, vp, vp, vp, i32, i32, f32, vp]),
    "gated_delta_rule_decode_cuda": (None, [vp] * 7 + [i32] * 4 + [vp]),
    "conv1d_prefill_cuda": (None, [vp, vp, vp, vp, i32, i32, i32, vp]),
    "pk_b200_gated_delta_rule_prefill_recurrent": (i32, [vp] * 7 + [i32] * 5 + [vp]),
    "prefill_attention_hd256_prep_cuda": (None, [vp] * 10 + [i32, i32, i32, vp, i32, f32, i32, vp]),
    "attention_gate_batch_hd256_cuda": (None, [vp, vp, i32, i32, vp]),
    "qk_norm_partial_rope_batched_decode_hd256_cuda": (None, [vp] * 8 + [i32, i32, i32, i32, f32, vp]),
    "batch_prefill_paged_cuda_hd256": (i32, [vp, vp, vp, i64, i64] + [vp] * 9 + [i32] * 7 + [i64, f32, vp]),
    "paged_attention_decode_cuda_hd256": (i32, [vp, vp, vp, i64, i64] + [vp] * 6 + [i32] * 5 + [i64, f32, vp]),
}

# B200 extensions (absent from the reference's library)
EXT_SIGNATURES = {
    "pk_b200_version": (C.c_char_p, []),
    "pk_b200_launch_count": (i64, [i32]),
    "pk_b200_set_pdl": (None, [i32]),
    "pk_b200_gemm_swiglu": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "pk_b200_gemm_segments": (i32, [vp, vp, C.POINTER(vp), C.POINTER(i32), i32, i32, i32, vp]),
    "pk_b200_gemv_fused": (i32, [C.POINTER(GemvArgs), vp]),
    "pk_b200_set_gemv_tuning": (None, [i32, i32, i32]),
    "pk_b200_decode_attention_fused": (i32, [vp, vp, vp, vp, vp, i64, i64] + [vp] * 8 + [f32, vp, vp] + [i32] * 7 + [i64, f32, vp]),
    "pk_b200_decode_attention_fused_prefetch": (i32, [vp, vp, vp, vp, vp, i64, i64] + [vp] * 8 + [f32, vp, vp] + [i32] * 7
                                                + [i64, f32, C.POINTER(PrefetchSpan), i32, vp]),
    "pk_b200_gemv_grid": (i32, [i32, i32]),
    "pk_b200_prefill_attention_tc": (i32, [vp, vp, vp, i64, i64, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i64, f32, vp]),
    "pk_b200_fa2_trace_copy": (i32, [vp, i32]),
    "pk_b200_gemm_plan": (i32, [i32, i32, i32, i32, i32]),
    "pk_tp_flag_bytes": (i64, []),
    "pk_tp_comm_create": (vp, [i32, i32, C.POINTER(vp), C.POINTER(vp), i64]),
    "pk_tp_comm_destroy": (None, [vp]),
    "pk_tp_max_rows": (i64, [vp, i32]),
    "pk_tp_top1_exchange": (i32, [vp, vp, vp, i32, i32, vp, C.c_uint32, vp]),
    "pk_tp_all_reduce": (i32, [vp, vp, i64, vp]),
    "pk_tp_all_reduce_rows": (i32, [vp, vp, i32, i32, vp]),
    "pk_tp_all_reduce_add_rms_norm": (i32, [vp, vp, vp, vp, vp, i32, i32, f32, vp]),
    "pk_tp_ipc_export": (i32, [vp, vp]),
    "pk_tp_ipc_open": (i32, [vp, C.POINTER(vp)]),
    "pk_tp_ipc_close": (i32, [vp]),
}


def load(path: str | None = None, extensions: bool = True) -> C.CDLL:
    """dlopen a kernel library behind the pegainfer-kernels C ABI and type its symbols.

    ``path`` may also point at the reference's own kernels built by ``oracle/Makefile``
    (``oracle/_ref/libkernels_ref.so``) -- same ABI, used by the tests as the on-GPU A/B
    baseline.  Raises OSError / AttributeError if the library or a symbol is missing.
    """
    path = path or KERNEL_LIB_PATH
    if not os.path.exists(path):
        raise OSError(f"{path} not found: build it with `python -m pegainfer_b200.build` "
                      "(there is no fallback path)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL if extensions else C.RTLD_LOCAL)
    sigs = dict(SIGNATURES)
    if extensions:
        sigs.update(EXT_SIGNATURES)
        sigs.update(QWEN35_SIGNATURES)
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB
