"""The long-pair fill's schedule and shifted-domain arithmetic (tests/long_model.py, a statement-
by-statement model of ga_long.cuh's long_fill_kernel) against the C oracle's full table.  CPU only."""
import random

import numpy as np

import long_model as LM
from oracle import c_oracle


def _dense(rng, A, simple):
    d = c_oracle.Dense.__new__(c_oracle.Dense)
    d.A = A
    if simple:
        mm, gh, gv = rng.randint(1, 6), rng.randint(0, 4), rng.randint(0, 4)
        d.sub = np.array([[0 if a == b else mm for b in range(A)] for a in range(A)], dtype=np.int32).reshape(-1)
        d.gap_h = np.full(A, gh, dtype=np.int32)
        d.gap_v = np.full(A, gv, dtype=np.int32)
    else:
        sub = np.array([[rng.randint(0, 12) for _ in range(A)] for _ in range(A)], dtype=np.int32)
        d.sub = np.minimum(sub, sub.T).reshape(-1).astype(np.int32)
        d.gap_h = np.array([rng.randint(0, 9) for _ in range(A)], dtype=np.int32)
        d.gap_v = np.array([rng.randint(0, 9) for _ in range(A)], dtype=np.int32)
    d.o = rng.randint(0, 9)
    d.max_cost = int(max(d.sub.max(), d.gap_h.max(), d.gap_v.max()))
    return d


def test_model_equals_oracle_table():
    rng = random.Random(5)
    for trial in range(120):
        A = rng.choice([2, 4, 5])
        d = _dense(rng, A, trial % 2 == 0)
        m, n = rng.randint(1, 70), rng.randint(1, 70)
        s1 = np.array([rng.randrange(A) for _ in range(m)], dtype=np.uint8)
        s2 = np.array([rng.randrange(A) for _ in range(n)], dtype=np.uint8)
        ELL, NC = rng.choice([(1, 1), (2, 1), (1, 2), (2, 2), (3, 2), (2, 3), (4, 1)])
        LANES, LSYNC = rng.choice([2, 4]), rng.choice([2, 4])
        LCC = rng.choice([c for c in (4, 8, 16) if c >= LANES * NC])
        sc = LM.Scheme(A, d.sub.tolist(), d.gap_h.tolist(), d.gap_v.tolist(), d.o, d.max_cost)
        cost, brow, vchk, (rowsum, colsum, LBAND, LR) = LM.fill_long(s1.tolist(), s2.tolist(), sc, ELL, NC, LANES, LSYNC, LCC)
        t = np.zeros((m + 1, n + 1, 3), dtype=np.int32)
        c_oracle.lib().gao_fill_table(m, n, s1, s2, *d.args(), 0, t.reshape(-1))
        o = d.o
        assert cost == int(t[m, n].min())
        for band in range(len(brow)):
            i = min((band + 1) * LBAND, m) if (band + 1) * LBAND <= m else None
            if i is not None:
                for j in range(1, n + 1):
                    h = int(t[i, j].min())
                    U = min(int(t[i, j, 0]) + o, int(t[i, j, 1]) + o, int(t[i, j, 2]))
                    got = brow[band][j]
                    assert LM.unshift(got[0], i, j, rowsum, colsum, o, m) == 4 * h
                    assert LM.unshift(got[1], i, j, rowsum, colsum, o, m) == 4 * U
            for c in range(1, n // LCC + 1):
                j = c * LCC
                for g in range(LANES):
                    for r in range(LR):
                        i = band * LBAND + g * LR + r + 1
                        if i > m:
                            continue
                        h = int(t[i, j].min())
                        L = min(int(t[i, j, 0]) + o, int(t[i, j, 1]), int(t[i, j, 2]) + o)
                        got = vchk[band][c][g][r]
                        assert LM.unshift(got[0], i, j, rowsum, colsum, o, m) == 4 * h
                        assert LM.unshift(got[1], i, j, rowsum, colsum, o, m) == 4 * L
