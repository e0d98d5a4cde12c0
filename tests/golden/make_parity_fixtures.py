#!/usr/bin/env python
"""Generate the full-size oracle fixtures (tests/golden/parity_<model>_p<prompt>_tp<N>.npz).

    python tests/golden/make_parity_fixtures.py --model qwen3-4b --prompt 2048 --steps 8 --tp 1
    python tests/golden/make_parity_fixtures.py --model qwen3-8b --prompt 128 --steps 8 --tp 1,2,4,8

Runs the CPU oracle (oracle/qwen3_oracle.{c,py}) on the seed-0 CPU-generated random-init checkpoint of
pegainfer_b200/synthetic.py with prompt ids (i % 1000) + 100, teacher-forced with its own greedy tokens.
Minutes per run on 8 cores (the 2048-token Qwen3-4B prefill is ~16 TFLOP in fp32).
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import qwen3_oracle as O  # noqa: E402
from pegainfer_b200.config import PRESETS  # noqa: E402
from pegainfer_b200.synthetic import random_weights, synthetic_prompt, to_numpy_bits  # noqa: E402
from tests.golden import parity_fixture as F  # noqa: E402
from tests.helpers import oracle_cfg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="qwen3-4b")
    ap.add_argument("--prompt", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--tp", default="1")
    args = ap.parse_args()
    cfg = PRESETS[args.model]
    t0 = time.time()
    w = to_numpy_bits(random_weights(cfg, seed=0, device="cpu"))
    crc = F.numpy_weights_crc(w)
    print(f"weights generated in {time.time() - t0:.0f}s, crc {crc:#010x}", flush=True)
    prompt = synthetic_prompt(args.prompt)
    for world in [int(x) for x in args.tp.split(",")]:
        t0 = time.time()
        orc = O.OracleQwen3(oracle_cfg(cfg), w, tp_world=world, num_pages=(args.prompt + args.steps) // 16 + 4)
        kv = orc.alloc_kv()
        rows = [F.pack_row(orc.prefill([prompt], [kv])[0])]
        print(f"tp{world}: prefill {time.time() - t0:.0f}s", flush=True)
        tokens, paths = [], []
        for _ in range(args.steps):
            tok = int(rows[-1]["idx_top"][0])
            tokens.append(tok)
            rows.append(F.pack_row(orc.decode([tok], [kv])[0]))
            paths.append(orc.last_attention_path)
        meta = dict(model=cfg.name, tp_world=world, prompt_len=args.prompt, n_decode=args.steps, seed=0,
                    weights_crc=crc, vocab=cfg.vocab_size, decode_attention_path=paths[-1] if paths else None,
                    generator=f"tests/golden/make_parity_fixtures.py --model {cfg.name} --prompt {args.prompt} "
                              f"--steps {args.steps} --tp {world}",
                    prompt_ids="(i % 1000) + 100")
        path = F.fixture_path(cfg.name, args.prompt, world)
        F.save(path, meta, tokens, rows)
        print(f"tp{world}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB) in {time.time() - t0:.0f}s; tokens {tokens}",
              flush=True)
        del orc


if __name__ == "__main__":
    main()
