"""The lane-level model of the CUDA kernels (tests/kernel_model.py) against the C oracle."""
import random

import numpy as np
import pytest

from oracle import c_oracle
import kernel_model as km


def _random_scheme(rng, A, simple):
    if simple:
        mm, gh, gv = rng.randint(1, 6), rng.randint(1, 4), rng.randint(1, 4)
        sub = [0 if a == b else mm for a in range(A) for b in range(A)]
        return sub, [gh] * A, [gv] * A
    sub = [0] * (A * A)
    for a in range(A):
        for b in range(a, A):
            sub[a * A + b] = sub[b * A + a] = 0 if a == b and rng.random() < 0.7 else rng.randint(0, 9)
    return sub, [rng.randint(1, 6) for _ in range(A)], [rng.randint(1, 6) for _ in range(A)]


@pytest.mark.parametrize("R,G,RPW", [(4, 8, 16), (5, 4, 8), (3, 5, 8), (10, 3, 8), (8, 32, 16), (1, 2, 16)])
def test_model_matches_oracle(R, G, RPW):
    rng = random.Random(R * 1000 + G * 10 + RPW)
    for case in range(120):
        A = rng.choice([2, 4, 5])
        simple = rng.random() < 0.5
        sub, gh, gv = _random_scheme(rng, A, simple)
        o = rng.randint(0, 7)
        m, n = rng.randint(1, 3 * G * R + 3), rng.randint(1, 40)
        if rng.random() < 0.3:
            m = rng.randint(1, 4)
        s1 = [rng.randrange(A) for _ in range(m)]
        s2 = [rng.randrange(A) for _ in range(n)]
        max_cost = max(sub + gh + gv)
        d = c_oracle.Dense.__new__(c_oracle.Dense)
        d.A, d.o, d.max_cost = A, o, max_cost
        d.sub = np.array(sub, dtype=np.int32)
        d.gap_h = np.array(gh, dtype=np.int32)
        d.gap_v = np.array(gv, dtype=np.int32)
        ops, cost, rc = c_oracle.align_codes(np.array(s1, dtype=np.uint8), np.array(s2, dtype=np.uint8), d)
        mcost, codes, L = km.fill_model(s1, s2, sub, gh, gv, o, max_cost, R=R, G=G, rows_per_word=RPW)
        assert mcost == cost, (case, m, n)
        mops, status = km.walk_model(codes, L, s1, s2, gh, gv, o, max_cost)
        assert status == rc, (case, m, n, o)
        if rc == 0:
            assert mops == list(ops), (case, m, n)
