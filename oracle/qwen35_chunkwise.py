"""Chunk-wise (matrix) form of the gated delta rule -- the algorithm the B200 prefill kernel for the Qwen3.5 linear
layers is specified by (the reference runs a 7-kernel Triton-AOT pipeline with chunk 64,
tools/triton/gated_delta_rule_chunkwise_kernels.py; this file is OUR restatement of the mathematics, checked on the CPU
against the per-token recurrence of oracle/qwen35_oracle.py::gated_delta_rule_step by tests/test_oracle_qwen35.py).

TEST INFRASTRUCTURE / design spec only.

Per value head, with L2-normalised k_t, scaled q_t, alpha_t = exp(g_t), beta_t and state S (dk x dv):
    S_t = alpha_t S_{t-1} + k_t u_t^T,      u_t = beta_t (v_t - (alpha_t S_{t-1})^T k_t),      o_t = S_t^T q_t
Inside a chunk of C tokens starting from S_0, with G_t = prod_{i<=t} alpha_i (cumulative decay from the chunk start):
    (I + A) U = diag(beta) (V - diag(G) K S_0),   A[t,i] = beta_t (G_t / G_i) (k_t . k_i) for i < t, else 0
    O        = diag(G) Q S_0 + (M * (Q K^T)) U,   M[t,i] = G_t / G_i for i <= t, else 0
    S_C      = G_C S_0 + K^T diag(G_C / G) U
Everything except the C x C unit-lower-triangular solve is a GEMM (tensor-core shaped: C x dk x dv, C x C x dv); the
solve is a forward substitution over C rows (or a blocked inverse).  Ratios G_t / G_i are formed as exp(cumsum(g)_t -
cumsum(g)_i) so long chunks do not underflow.
"""
from __future__ import annotations

import numpy as np


def gated_delta_rule_chunk(q, k, v, g, beta, S0):
    """One chunk, one value head.  q, k: [C, dk] (k L2-normalised, q normalised and scaled); v: [C, dv]; g, beta: [C];
    S0: [dk, dv].  Returns (O [C, dv], S_C [dk, dv]); float64 inside for a clean comparison."""
    q, k, v, g, beta, S0 = (np.asarray(x, np.float64) for x in (q, k, v, g, beta, S0))
    C = q.shape[0]
    cg = np.cumsum(g)                                  # log G_t
    ratio = np.exp(cg[:, None] - cg[None, :])          # G_t / G_i
    lower = np.tril(np.ones((C, C)), -1)
    A = beta[:, None] * ratio * (k @ k.T) * lower
    rhs = beta[:, None] * (v - np.exp(cg)[:, None] * (k @ S0))
    U = np.zeros_like(rhs)
    for t in range(C):                                 # forward substitution of the unit lower-triangular system
        U[t] = rhs[t] - A[t, :t] @ U[:t]
    M = ratio * np.tril(np.ones((C, C)))
    O = np.exp(cg)[:, None] * (q @ S0) + (M * (q @ k.T)) @ U
    SC = np.exp(cg[-1]) * S0 + k.T @ (np.exp(cg[-1] - cg)[:, None] * U)
    return O, SC


def gated_delta_rule_chunkwise(q, k, v, g, beta, S0, chunk=64):
    """Whole sequence [T, ...] for one value head, processed in chunks."""
    T = q.shape[0]
    outs, S = [], np.asarray(S0, np.float64)
    for t0 in range(0, T, chunk):
        sl = slice(t0, min(T, t0 + chunk))
        O, S = gated_delta_rule_chunk(q[sl], k[sl], v[sl], g[sl], beta[sl], S)
        outs.append(O)
    return np.concatenate(outs), S
