#!/usr/bin/env python3
"""Randomised parity sweep on the GPU against the C oracle (test infrastructure, not a product path):
random schemes (simple / table-driven), alphabets, length distributions (ragged, fixed, lengths that
hit every short-last-pass variant), kernel configurations, flags (serial, keep-order, cost-only,
packed letter streams), scratch limits and entry points.  Usage: fuzz_parity.py [seconds] [seed]"""
import os
import random
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from globalign_b200 import _native, lowering  # noqa: E402
from oracle import c_oracle  # noqa: E402

PACKED = ["fill_p16_r7_g8", "fill_p16_r10_g8", "fill_p16_r13_g8", "fill_p16_r16_g8", "fill_p16_r13_g16", "fill_p16_r19_g16",
               "fill_p16_r5_g32", "fill_p16_r10_g16", "fill_p16_r15_g10", "fill_p16_r19_g8", "fill_p16_r16_g16",
          "fill_p16_r8_g32", "fill_p16_r16_g32"]
I32S = ["fill_i32s_r8_g8", "fill_i32s_r8_g16", "fill_i32s_r8_g32"]
I32T = ["fill_i32t_r8_g8", "fill_i32t_r8_g16", "fill_i32t_r8_g32"]
P16T = ["fill_p16t_r8_g8", "fill_p16t_r8_g16", "fill_p16t_r8_g32", "fill_p16t_r12_g32", "fill_p16t_r16_g32",
             "fill_p16b_r16_g32", "fill_p16h_r16_g32"]


def dense_simple(rng, A):
    letters = [chr(65 + k) for k in range(A)]
    mm, gh, gv = rng.randint(1, 9), rng.randint(1, 5), rng.randint(1, 5)
    cm = {x: {y: (0 if x == y else mm) for y in letters} for x in letters}
    for x in letters:
        cm[x]["-"] = gv
    cm["-"] = {y: gh for y in letters}
    cm["-"]["-"] = 0
    return lowering.DenseScheme(cm, rng.choice([0, 1, 3, 10, 17]))


def dense_table(rng, A):
    letters = [chr(65 + k) for k in range(A)]
    cm = {x: {} for x in letters + ["-"]}
    for i, x in enumerate(letters):
        for j, y in enumerate(letters):
            cm[x][y] = cm[y][x] if j < i else (0 if (x == y and rng.random() < 0.8) else rng.randint(0, 12))
        cm[x]["-"] = rng.randint(1, 9)
        cm["-"][x] = rng.randint(1, 9)
    cm["-"]["-"] = 0
    return lowering.DenseScheme(cm, rng.choice([0, 2, 8, 12]))


def batch(rng, n_pairs, A, max_m, max_n, fixed, similar):
    s1, s2, o1, o2 = [], [], [0], [0]
    fm, fn = rng.randint(1, max_m), rng.randint(1, max_n)
    for _ in range(n_pairs):
        m = fm if fixed else (rng.randint(1, max_m) if rng.random() < 0.85 else rng.randint(1, 3))
        a = [rng.randrange(A) for _ in range(m)]
        if rng.random() < similar and not fixed:
            b = []
            for x in a:
                u = rng.random()
                if u < 0.08:
                    b.append(rng.randrange(A))
                elif u < 0.12:
                    continue
                elif u < 0.16:
                    b += [x, rng.randrange(A)]
                else:
                    b.append(x)
            b = b[:max_n] or [rng.randrange(A)]
        else:
            n = fn if fixed else rng.randint(1, max_n)
            b = [(a[k] if k < len(a) and rng.random() < 0.85 else rng.randrange(A)) for k in range(n)]
        s1 += a
        s2 += b
        o1.append(len(s1))
        o2.append(len(s2))
    return (np.array(s1, dtype=np.uint8), np.array(o1, dtype=np.int64),
            np.array(s2, dtype=np.uint8), np.array(o2, dtype=np.int64))


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    ctx = _native.Context(0)
    t_end = time.time() + budget
    runs = pairs = 0
    kernels = {}
    while time.time() < t_end:
        simple = rng.random() < 0.5
        A = rng.choice([2, 3, 4, 4, 5] if simple else [3, 4, 5, 20, 24])
        dense = dense_simple(rng, A) if simple else dense_table(rng, A)
        cfgs = (PACKED + I32S) if simple else (I32T + P16T)
        cfg = rng.choice(cfgs + [None, None])
        if cfg:
            R, G = int(cfg.split("_r")[1].split("_")[0]), int(cfg.split("_g")[1])
            band = R * G
        else:
            band = rng.choice([64, 150, 300, 700])
        max_m = max(1, min(int(band * rng.choice([0.3, 1.0, 1.0, 1.6, 2.4])), 900))
        max_n = rng.choice([5, 60, 150, 260, 400])
        fixed = rng.random() < 0.3
        n_pairs = rng.choice([1, 2, 7, 64, 301, 1200])
        data = batch(rng, n_pairs, A, max_m, max_n, fixed, rng.random())
        os.environ.pop("GA_FORCE_CFG", None)
        if cfg:
            os.environ["GA_FORCE_CFG"] = cfg
        os.environ["GA_CODE_PAIRS"] = rng.choice(["1", "1", "1", "0"])
        flags = 0
        if rng.random() < 0.2:
            flags |= _native.GA_F_SERIAL
        if rng.random() < 0.2:
            flags |= _native.GA_F_KEEP_ORDER
        bits = 8
        if A <= 4 and rng.random() < 0.4:
            bits = rng.choice([2, 4, 5])
        elif A <= 16 and rng.random() < 0.3:
            bits = rng.choice([4, 5])
        elif A <= 32 and rng.random() < 0.3:
            bits = 5
        flags |= _native.seq_flag(bits)
        ctx.set_scheme(dense)
        ctx.set_scratch_limit(rng.choice([64 << 30, 64 << 30, 32 << 20, 8 << 20]))
        p1, p2 = _native.pack_letters(data[0], bits), _native.pack_letters(data[2], bits)
        desc = dict(cfg=cfg, A=A, simple=simple, n_pairs=n_pairs, max_m=max_m, max_n=max_n, fixed=fixed, flags=flags,
                    bits=bits, o=dense.gap_open, seed=seed, run=runs)
        try:
            if fixed and n_pairs > 1 and rng.random() < 0.6:
                m0, n0 = int(data[1][1]), int(data[3][1])
                cost, ops, stride, ops_len, status = ctx.align_batch_uniform(p1, m0, p2, n0, n_pairs, flags=flags)
                ops_off = np.arange(n_pairs + 1, dtype=np.int64) * stride
            else:
                cost, ops, ops_off, ops_len, status = ctx.align_batch(p1, data[1], p2, data[3], flags=flags)
        except RuntimeError as e:
            msg = str(e)
            # legitimate refusals: range / scratch / forced kernel not applicable
            if any(k in msg for k in ("no usable kernel", "scratch limit", "overflow", "too long")):
                continue
            print("FAILED CALL", desc, msg, flush=True)
            raise
        kernels[ctx.plan_info()["fill_kernel"]] = kernels.get(ctx.plan_info()["fill_kernel"], 0) + 1
        d = c_oracle.Dense.__new__(c_oracle.Dense)
        d.A, d.sub, d.gap_h, d.gap_v, d.o, d.max_cost = dense.A, dense.sub, dense.gap_h, dense.gap_v, dense.gap_open, dense.max_cost
        o_ops, o_off, o_n, o_cost, o_status = c_oracle.align_batch(*data, d)
        ok = np.array_equal(cost, o_cost) and np.array_equal(status, o_status)
        if ok:
            for k in range(n_pairs):
                if o_status[k] == 0:
                    got = _native.unpack_ops(ops, int(ops_off[k]), int(ops_len[k]))
                    if not np.array_equal(got, o_ops[o_off[k]:o_off[k] + o_n[k]]):
                        ok = False
                        desc["pair"] = k
                        break
        if not ok:
            print("MISMATCH", desc, flush=True)
            sys.exit(1)
        runs += 1
        pairs += n_pairs
    print(f"fuzz ok: {runs} batches, {pairs} pairs, seed {seed}; kernels used: {kernels}")


if __name__ == "__main__":
    main()
