"""Bring-up / regression check of the tcgen05 prefill attention kernel against the CPU oracle, one process per run.

    python tests/tools/check_prefill_tc.py [case ...]     cases: small ragged offset mid long batch (default: all)

Prints per case the max error in bf16 ulps (floor = max|want| / 32) and a timing for the long case.
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import qwen3_oracle as O  # noqa: E402
from pegainfer_b200 import ffi  # noqa: E402
from pegainfer_b200.paged_kv import PagedKvLayout  # noqa: E402
from tests.helpers import bits, f32, ulp_err  # noqa: E402

CASES = {
    "small": ([0], [128], 32, 8),
    "ragged": ([0, 0, 0], [33, 100, 5], 32, 8),
    "offset": ([40], [60], 32, 8),
    "mid": ([0], [300], 4, 1),
    "long": ([0], [2048], 32, 8),
    "batch": ([100, 0], [700, 129], 32, 8),
}


def main():
    names = sys.argv[1:] or list(CASES)
    lib = ffi.lib()
    torch.zeros(1, device="cuda")
    lib.cuda_set_device(0)
    lib.cublas_init()
    st = torch.cuda.current_stream().cuda_stream
    ok_all = True
    for name in names:
        starts, lens, nq, nkv = CASES[name]
        hd, layer, bs = 128, 1, len(lens)
        kv_lens = [s + n for s, n in zip(starts, lens)]
        L = PagedKvLayout.new(2, nkv, hd, 16)
        rng = np.random.RandomState(5)
        need = [-(-s // 16) for s in kv_lens]
        ids = rng.permutation(np.arange(1, sum(need) + 4))
        pi, ip, lpl, off = [], [0], [], 0
        for s, n in zip(kv_lens, need):
            pi += ids[off:off + n].tolist(); off += n
            ip.append(len(pi)); lpl.append(((s - 1) % 16) + 1)
        g = torch.Generator().manual_seed(11)
        kv = torch.randn((sum(need) + 5) * L.page_stride, generator=g).to(torch.bfloat16)
        T = sum(lens)
        q = torch.randn((T, nq * hd), generator=g).to(torch.bfloat16)
        q_indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        dv = lambda a: torch.tensor(np.asarray(a, np.int32), device="cuda")
        q_d, kv_d = q.cuda(), kv.cuda()
        out = torch.zeros((T, nq * hd), dtype=torch.bfloat16, device="cuda")
        pi_d, ip_d, lpl_d, qi_d = dv(pi), dv(ip), dv(lpl), dv(q_indptr)
        sm = 1 / math.sqrt(hd)
        call = lambda: lib.pk_b200_prefill_attention_tc(q_d.data_ptr(), out.data_ptr(), kv_d.data_ptr(), L.k_offset(layer), L.v_offset(layer),
                                                        pi_d.data_ptr(), ip_d.data_ptr(), lpl_d.data_ptr(), qi_d.data_ptr(), nq, nkv, hd, 16, T, bs,
                                                        L.page_stride, sm, st)
        rc = call()
        torch.cuda.synchronize()
        want = O.batch_prefill_paged(bits(q), bits(kv), L.k_offset(layer), L.v_offset(layer), np.array(pi, np.int32), np.array(ip, np.int32),
                                     np.array(lpl, np.int32), q_indptr, nq, nkv, hd, 16, L.page_stride, sm)
        e = ulp_err(bits(out).ravel(), np.asarray(want).ravel(), float(np.abs(f32(want)).max()) / 32)
        ok = rc == 0 and np.isfinite(e).all() and e.max() <= 8
        ok_all &= bool(ok)
        msg = f"TC_ATTN case={name} rc={rc} max_ulp={e.max():.2f} mean_ulp={e.mean():.3f} {'OK' if ok else 'FAIL'}"
        if name in ("long", "batch"):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            for _ in range(3):
                call()
            e0.record()
            for _ in range(20):
                call()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            fl = sum(4.0 * nq * hd * (n * (n + 1) / 2 + n * s) for s, n in zip(starts, lens))
            msg += f" time={us:.1f}us {fl / us / 1e6:.1f} TFLOP/s"
        print(msg, flush=True)
    print("TC_ATTN_ALL", "PASS" if ok_all else "FAIL", flush=True)


if __name__ == "__main__":
    main()
