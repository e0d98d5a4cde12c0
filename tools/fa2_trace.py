"""Per-block hand-over timeline of the two-tile prefill attention kernel (PK_FA2_DBG=32): clock64 stamps of tile 0 in the
heaviest CTA of head 0 -- softmax (S seen, max done, exps done, P signalled), MMA issuer (P seen, O-scaled seen, V seen,
P V issued, K seen, S issued), correction (alpha seen, P V done seen, O scaled)."""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PK_FA2_DBG"] = os.environ.get("PK_FA2_DBG", "32")
os.environ["PK_PREFILL_ATTN"] = "tc2"
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pegainfer_b200 import ffi  # noqa: E402

lib = ffi.lib()
torch.zeros(1, device="cuda")
lib.cuda_set_device(0)
st = torch.cuda.current_stream().cuda_stream
raw = lib
T = int(os.environ.get("PK_T", "4096"))
nq, nkv, hd = 32, 8, 128
pages = T // 16
stride = 2 * 16 * nkv * hd
kv = torch.randn(((pages + 2) * stride,), device="cuda").to(torch.bfloat16)
q = torch.randn((T, nq * hd), device="cuda").to(torch.bfloat16)
out = torch.zeros_like(q)
i32 = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
pi, ip, lpl = i32(list(range(1, pages + 1))), i32([0, pages]), i32([16])
qi, z, kc, tn = i32([0, T]), i32([0] * 4096), i32([T]), i32([T])
fn = lambda: lib.batch_prefill_paged_cuda_with_cta_tile_q(q.data_ptr(), out.data_ptr(), kv.data_ptr(), 0, 16 * nkv * hd, pi.data_ptr(),
    ip.data_ptr(), lpl.data_ptr(), qi.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), kc.data_ptr(), tn.data_ptr(), nq, nkv, hd, 16, T, 1,
    1, stride, 1 / math.sqrt(hd), 64, st)
for _ in range(3):
    assert fn() == 0
torch.cuda.synchronize()
buf = np.zeros(3 * 128 * 8, dtype=np.uint64)
f = raw.pk_b200_fa2_trace_copy
assert f(buf.ctypes.data, buf.nbytes) == 0
t = buf.reshape(3, 128, 8).astype(np.int64)
base = t[0, 0, 0]
nblk = T // 128 - 1
print(f"T={T}: tile 0 of the heaviest pair, {nblk} blocks; cycles relative to the first S seen")
print("blk | softmax: S_seen max_done exps_done P_signalled | mma(tile 0): VK_seen PO_seen PV_issued PV_committed S_issued(j+1) S_committed(j+1) | corr: alpha_seen pvdone_seen scaled")
for j in range(2, min(nblk, 14)):
    sm = [int(x - base) for x in t[0, j, :4]]
    mm = [int(x - base) for x in t[1, j, :4]] + [int(x - base) for x in t[1, j + 1, 4:6]]
    cc = [int(x - base) for x in t[2, j, :3]]
    print(j, "|", *sm, "|", *mm, "|", *cc)
per = (t[0, 12, 0] - t[0, 4, 0]) / 8
print("period per block (softmax S_seen to S_seen):", per)
