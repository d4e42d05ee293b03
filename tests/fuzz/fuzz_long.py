#!/usr/bin/env python3
"""Randomised parity sweep of the long-pair path (ga_align_long: band-pipelined fill + linear-space
traceback, parallel and sequential forms) against the C oracle's full-table alignment.
Usage: fuzz_long.py [seconds] [seed]"""
import os
import random
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "fuzz"))
from globalign_b200 import _native  # noqa: E402
from oracle import c_oracle  # noqa: E402
from fuzz_parity import dense_simple, dense_table  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    ctx = _native.Context(0)
    t_end = time.time() + budget
    runs = 0
    modes = {}
    while time.time() < t_end:
        simple = rng.random() < 0.6
        A = rng.choice([2, 4, 4, 5] if simple else [3, 4, 20])
        dense = dense_simple(rng, A) if simple else dense_table(rng, A)
        m = rng.choice([1, 2, 127, 128, 129, 191, 192, 193, 255, 256, 257, 383, 384, 385, rng.randint(1, 700), rng.randint(1, 3000), rng.randint(1, 3000)])
        n = rng.choice([1, 2, 15, 16, 17, 127, 128, 129, rng.randint(1, 700), rng.randint(1, 3000), rng.randint(1, 3000)])
        a = np.array([rng.randrange(A) for _ in range(m)], dtype=np.uint8)
        kind = rng.random()
        if kind < 0.5:       # related sequences (indels and substitutions)
            b = []
            for x in a:
                u = rng.random()
                if u < 0.08:
                    b.append(rng.randrange(A))
                elif u < 0.11:
                    continue
                elif u < 0.14:
                    b += [int(x)] + [rng.randrange(A) for _ in range(rng.randint(1, 40))]
                else:
                    b.append(int(x))
            b = np.array(b[:3000] or [0], dtype=np.uint8)
        else:
            b = np.array([rng.randrange(A) for _ in range(n)], dtype=np.uint8)
        flags = _native.GA_F_SERIAL if rng.random() < 0.25 else 0
        os.environ["GA_LONG_SEQ_BANDS"] = rng.choice(["0", "0", "6"])      # band-parallel traceback also for small pairs
        ctx.set_scheme(dense)
        desc = dict(A=A, simple=simple, m=len(a), n=len(b), flags=flags, o=dense.gap_open, seed=seed, run=runs)
        cost, ops, status, info = ctx.align_long(a, b, flags=flags)
        modes[info["trace_mode"]] = modes.get(info["trace_mode"], 0) + 1
        d = c_oracle.Dense.__new__(c_oracle.Dense)
        d.A, d.sub, d.gap_h, d.gap_v, d.o, d.max_cost = dense.A, dense.sub, dense.gap_h, dense.gap_v, dense.gap_open, dense.max_cost
        o_ops, o_cost, o_rc = c_oracle.align_codes(a, b, d)
        ok = cost == o_cost and status == (1 if o_rc == 1 else 0)
        if ok and o_rc == 0:
            ok = np.array_equal(ops, o_ops)
        if not ok:
            print("MISMATCH", desc, cost, o_cost, status, o_rc, flush=True)
            sys.exit(1)
        runs += 1
    print(f"fuzz_long ok: {runs} pairs, seed {seed}; trace modes used: {modes}")


if __name__ == "__main__":
    main()
