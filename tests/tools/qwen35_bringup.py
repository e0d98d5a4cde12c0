"""Bring-up harness for the Qwen3.5 hybrid layers: a single-request forward driven from Python, one token at a time,
through the C ABI only (GEMV, (1+w) / gated norms, conv1d, delta-rule decode step, HD-256 prep + paged decode attention +
gate, scatter, SiLU-mul, add), checked step by step against oracle/qwen35_oracle.py on a tiny random hybrid model.
It is the launch order a C++ `Qwen35Model` will have to reproduce (pegainfer-qwen35-4b/src/{prefill,batch_decode}.rs);
prefill is run as repeated decode steps until the chunk-wise kernels exist.

    python tests/tools/qwen35_bringup.py          # needs a GPU; prints QWEN35_BRINGUP PASS / FAIL
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import qwen3_oracle as O  # noqa: E402
from oracle.qwen35_oracle import OracleQwen35, Qwen35Config  # noqa: E402
from pegainfer_b200 import ffi  # noqa: E402

LT = ["linear_attention", "linear_attention", "full_attention", "linear_attention"]
CFG = Qwen35Config(hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=1, head_dim=256,
                   vocab_size=512, linear_num_key_heads=2, linear_num_value_heads=4, linear_key_head_dim=128, linear_value_head_dim=128,
                   linear_conv_kernel_dim=4, layer_types=LT, rms_norm_eps=1e-6, rope_theta=1e7, partial_rotary_factor=0.25)


def random_weights(c: Qwen35Config, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, scale=0.08: (torch.randn(s, generator=g) * scale).to(torch.bfloat16)
    w = {"model.embed_tokens.weight": r(c.vocab_size, c.hidden_size, scale=0.15), "model.norm.weight": r(c.hidden_size, scale=0.1)}
    qkv_dim = 2 * c.linear_num_key_heads * c.linear_key_head_dim + c.linear_num_value_heads * c.linear_value_head_dim
    z_dim = c.linear_num_value_heads * c.linear_value_head_dim
    for i, kind in enumerate(c.layer_types):
        p = f"model.layers.{i}."
        w[p + "input_layernorm.weight"], w[p + "post_attention_layernorm.weight"] = r(c.hidden_size, scale=0.1), r(c.hidden_size, scale=0.1)
        w[p + "mlp.gate_proj.weight"], w[p + "mlp.up_proj.weight"] = r(c.intermediate_size, c.hidden_size), r(c.intermediate_size, c.hidden_size)
        w[p + "mlp.down_proj.weight"] = r(c.hidden_size, c.intermediate_size)
        if kind == "full_attention":
            a = p + "self_attn."
            w[a + "q_proj.weight"] = r(c.num_attention_heads * c.head_dim * 2, c.hidden_size)
            w[a + "k_proj.weight"], w[a + "v_proj.weight"] = r(c.num_key_value_heads * c.head_dim, c.hidden_size), r(c.num_key_value_heads * c.head_dim, c.hidden_size)
            w[a + "o_proj.weight"] = r(c.hidden_size, c.num_attention_heads * c.head_dim)
            w[a + "q_norm.weight"], w[a + "k_norm.weight"] = r(c.head_dim, scale=0.1), r(c.head_dim, scale=0.1)
        else:
            a = p + "linear_attn."
            w[a + "in_proj_qkv.weight"], w[a + "in_proj_z.weight"] = r(qkv_dim, c.hidden_size), r(z_dim, c.hidden_size)
            w[a + "in_proj_b.weight"], w[a + "in_proj_a.weight"] = r(c.linear_num_value_heads, c.hidden_size), r(c.linear_num_value_heads, c.hidden_size)
            w[a + "conv1d.weight"] = r(qkv_dim, c.linear_conv_kernel_dim, scale=0.4)
            w[a + "dt_bias"] = r(c.linear_num_value_heads, scale=0.5)
            w[a + "A_log"] = (torch.randn(c.linear_num_value_heads, generator=g) * 0.5).float()       # f32 in the checkpoint
            w[a + "norm.weight"] = (torch.randn(c.linear_value_head_dim, generator=g) * 0.2 + 1).float()  # f32 in the checkpoint
            w[a + "out_proj.weight"] = r(c.hidden_size, z_dim)
    return w


class Qwen35Gpu:
    def __init__(self, c: Qwen35Config, w: dict, max_tokens=64):
        self.c, self.lib = c, ffi.lib()
        torch.zeros(1, device="cuda")
        self.lib.cuda_set_device(0)
        self.lib.cublas_init()
        self.st = torch.cuda.current_stream().cuda_stream
        self.w = {k: v.cuda().contiguous() for k, v in w.items()}
        cos, sin = O.precompute_rope(c.rotary_dim, 4096, c.rope_theta)
        self.cos = torch.from_numpy(cos.view(np.int16)).view(torch.bfloat16).cuda()
        self.sin = torch.from_numpy(sin.view(np.int16)).view(torch.bfloat16).cuda()
        self.full = [i for i, t in enumerate(c.layer_types) if t == "full_attention"]
        nkv, hd = c.num_key_value_heads, c.head_dim
        self.block = 16 * nkv * hd
        self.layer_stride, self.page_stride = 2 * self.block, len(self.full) * 2 * self.block
        pages = max_tokens // 16 + 2
        self.pool = torch.zeros((pages + 1) * self.page_stride, dtype=torch.bfloat16, device="cuda")
        self.page_ids = torch.arange(1, pages + 1, dtype=torch.int32, device="cuda")
        qkv_dim = 2 * c.linear_num_key_heads * c.linear_key_head_dim + c.linear_num_value_heads * c.linear_value_head_dim
        self.conv = {i: torch.zeros((qkv_dim, c.linear_conv_kernel_dim - 1), dtype=torch.bfloat16, device="cuda")
                     for i, t in enumerate(c.layer_types) if t == "linear_attention"}
        self.S = {i: torch.zeros((c.linear_num_value_heads, c.linear_key_head_dim, c.linear_value_head_dim), dtype=torch.float32, device="cuda")
                  for i in self.conv}
        self.pos = 0

    def _gemv(self, W, x):
        y = torch.empty(W.shape[0], dtype=torch.bfloat16, device="cuda")
        self.lib.gemm_cuda(W.data_ptr(), x.data_ptr(), y.data_ptr(), W.shape[0], 1, W.shape[1], self.st)
        return y

    def _norm(self, x, w):
        y = torch.empty_like(x)
        self.lib.rms_norm_offset_cuda(x.data_ptr(), w.data_ptr(), y.data_ptr(), x.numel(), self.c.rms_norm_eps, self.st)
        return y

    def _add(self, a, b):
        y = torch.empty_like(a)
        self.lib.add_cuda(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), self.st)
        return y

    def _linear(self, li, x):
        c, w, lib, p = self.c, self.w, self.lib, f"model.layers.{li}.linear_attn."
        nk, nv, dk, dv = c.linear_num_key_heads, c.linear_num_value_heads, c.linear_key_head_dim, c.linear_value_head_dim
        qkv, z = self._gemv(w[p + "in_proj_qkv.weight"], x), self._gemv(w[p + "in_proj_z.weight"], x)
        b, a = self._gemv(w[p + "in_proj_b.weight"], x), self._gemv(w[p + "in_proj_a.weight"], x)
        conv = torch.empty_like(qkv)
        lib.conv1d_prefill_cuda(qkv.data_ptr(), w[p + "conv1d.weight"].data_ptr(), self.conv[li].data_ptr(), conv.data_ptr(), qkv.numel(), 1,
                                c.linear_conv_kernel_dim, self.st)
        heads = torch.empty(nv * dv, dtype=torch.bfloat16, device="cuda")
        lib.gated_delta_rule_decode_cuda(conv.data_ptr(), b.data_ptr(), a.data_ptr(), w[p + "dt_bias"].data_ptr(), w[p + "A_log"].data_ptr(),
                                         self.S[li].data_ptr(), heads.data_ptr(), nk, nv, dk, dv, self.st)
        normed = torch.empty_like(heads)
        lib.rms_norm_gated_cuda(heads.data_ptr(), w[p + "norm.weight"].data_ptr(), z.data_ptr(), normed.data_ptr(), nv, dv, c.rms_norm_eps, self.st)
        return self._gemv(w[p + "out_proj.weight"], normed)

    def _full(self, li, x):
        c, w, lib, p = self.c, self.w, self.lib, f"model.layers.{li}.self_attn."
        nq, nkv, hd = c.num_attention_heads, c.num_key_value_heads, c.head_dim
        qf, k, v = self._gemv(w[p + "q_proj.weight"], x), self._gemv(w[p + "k_proj.weight"], x), self._gemv(w[p + "v_proj.weight"], x)
        q = torch.empty(nq * hd, dtype=torch.bfloat16, device="cuda")
        pos_d = torch.tensor([self.pos], dtype=torch.int32, device="cuda")
        lib.qk_norm_partial_rope_batched_decode_hd256_cuda(qf.data_ptr(), k.data_ptr(), w[p + "q_norm.weight"].data_ptr(), w[p + "k_norm.weight"].data_ptr(),
                                                           self.cos.data_ptr(), self.sin.data_ptr(), pos_d.data_ptr(), q.data_ptr(), nq, nkv, 1,
                                                           c.rotary_dim, c.rms_norm_eps, self.st)
        fi = self.full.index(li)
        k_off, v_off = fi * self.layer_stride, fi * self.layer_stride + self.block
        n_pages = self.pos // 16 + 1
        ip = torch.tensor([0, n_pages], dtype=torch.int32, device="cuda")
        lpl = torch.tensor([self.pos % 16 + 1], dtype=torch.int32, device="cuda")
        zero = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = lib.paged_kv_scatter_cuda(self.pool.data_ptr(), k_off, v_off, self.page_ids.data_ptr(), ip.data_ptr(), lpl.data_ptr(), k.data_ptr(),
                                       v.data_ptr(), zero.data_ptr(), pos_d.data_ptr(), 1, nkv, hd, 16, self.page_stride, nkv * hd, hd, self.st)
        assert rc == 0
        out = torch.empty(nq * hd, dtype=torch.bfloat16, device="cuda")
        rc = lib.paged_attention_decode_cuda_hd256(q.data_ptr(), out.data_ptr(), self.pool.data_ptr(), k_off, v_off, self.page_ids.data_ptr(),
                                                   ip.data_ptr(), lpl.data_ptr(), zero.data_ptr(), zero.data_ptr(), zero.data_ptr(), nq, nkv, hd, 16, 1,
                                                   self.page_stride, 1 / math.sqrt(hd), self.st)
        assert rc == 0
        lib.attention_gate_batch_hd256_cuda(qf.data_ptr(), out.data_ptr(), nq, 1, self.st)
        y = self._gemv(w[p + "o_proj.weight"], out)
        torch.cuda.synchronize()  # the small metadata tensors above must outlive the launches
        return y

    def step(self, tok):
        c, w = self.c, self.w
        tok_d = torch.tensor([tok], dtype=torch.int32, device="cuda")
        h = torch.empty(c.hidden_size, dtype=torch.bfloat16, device="cuda")
        self.lib.embedding_decode_cuda(w["model.embed_tokens.weight"].data_ptr(), tok_d.data_ptr(), h.data_ptr(), c.hidden_size, self.st)
        for li, kind in enumerate(c.layer_types):
            p = f"model.layers.{li}."
            x = self._norm(h, w[p + "input_layernorm.weight"])
            h = self._add(h, self._full(li, x) if kind == "full_attention" else self._linear(li, x))
            x = self._norm(h, w[p + "post_attention_layernorm.weight"])
            gate, up = self._gemv(w[p + "mlp.gate_proj.weight"], x), self._gemv(w[p + "mlp.up_proj.weight"], x)
            act = torch.empty_like(gate)
            self.lib.silu_mul_triton_aot_cuda(gate.data_ptr(), up.data_ptr(), act.data_ptr(), gate.numel(), self.st)
            h = self._add(h, self._gemv(w[p + "mlp.down_proj.weight"], act))
        logits = self._gemv(w["model.embed_tokens.weight"], self._norm(h, w["model.norm.weight"]))
        torch.cuda.synchronize()
        self.pos += 1
        return logits


def to_oracle_weights(w):
    return {k: (v.view(torch.int16).numpy().view(np.uint16) if v.dtype == torch.bfloat16 else v.numpy()) for k, v in w.items()}


def run(n_tokens=24, tol_ulp=8.0):
    w = random_weights(CFG)
    gpu, orc = Qwen35Gpu(CFG, w), OracleQwen35(CFG, to_oracle_weights(w))
    toks = [(7 * i + 3) % CFG.vocab_size for i in range(n_tokens)]
    worst, ok = 0.0, True
    for i, t in enumerate(toks):
        got = O.bf16_to_f32(gpu.step(t).view(torch.int16).cpu().numpy().view(np.uint16))
        want = orc.decode(t)
        err = float((np.abs(got - want) / O.bf16_ulp(np.full_like(want, np.abs(want).max()))).max())
        margin = np.sort(want)[-1] - np.sort(want)[-2]
        same = got.argmax() == want.argmax() or margin <= tol_ulp * float(O.bf16_ulp(np.array([np.abs(want).max()], np.float32))[0])
        worst = max(worst, err)
        ok &= err <= tol_ulp and bool(same)
        print(f"step {i:2d} tok {t:3d}: max err {err:5.2f} ulp(rowmax)  argmax {got.argmax()} / {want.argmax()}", flush=True)
    print("QWEN35_BRINGUP", "PASS" if ok else "FAIL", f"worst {worst:.2f} ulp", flush=True)
    return ok


if __name__ == "__main__":
    sys.exit(0 if run() else 1)
