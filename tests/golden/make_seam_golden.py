#!/usr/bin/env python3
"""Golden vectors for the traceback ON A CALLER-PROVIDED TABLE (run in the build container only).

    python tests/golden/make_seam_golden.py

The reference's dp_array_backward (globaligner.py:395-593) walks whatever dp_array it is
handed; its own test (tests/globaligner_test.py:6-37) builds one with hand-written,
non-standard boundary values.  This script takes that table and a few hundred random ones
whose boundary row/column were perturbed by hand-like edits, fills them with the reference's
dp_array_forward, walks them with the reference's dp_array_backward (canonical tie order:
``random.choice`` patched to "first listed option") and freezes inputs + outputs in
``tests/golden/seam_cases.json``.  Where the reference raises (its walk leaves the table) the
exception type is recorded instead.  It also asserts that oracle/py_oracle.py (fill_in_place +
walk on the same table) reproduces every case -- this pins the oracle's walk on non-standard
tables.  /root/reference is never read at test time.
"""

from __future__ import annotations

import copy
import json
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF / "src"))
sys.path.insert(0, str(REF))

import globalign.globaligner as G          # noqa: E402  (the reference)
import globalign.start as S                # noqa: E402
from oracle import py_oracle               # noqa: E402

G.random.choice = lambda seq: seq[0]       # canonical tie order


def boundary_only(dp, m, n):
    return {"row0": [list(dp[0][j]) for j in range(n + 1)], "col0": [list(dp[i][0]) for i in range(m + 1)]}


def run_case(s1, s2, cm, o, dp):
    """dp: boundary set, interior None.  -> record"""
    m, n = len(s1), len(s2)
    rec = {"seq_1": s1, "seq_2": s2, "cm": cm, "o": o, **boundary_only(dp, m, n)}
    ref_dp = copy.deepcopy(dp)
    G.dp_array_forward(ref_dp, s1, s2, cm, o)
    rec["table"] = [[list(c) for c in row] for row in ref_dp]
    try:
        a1, mid, a2, cost = G.dp_array_backward(ref_dp, s1, s2, cm, o)
        rec["out"] = [a1, mid, a2, cost]
    except Exception as exc:                 # the reference's walk left the table
        rec["error"] = type(exc).__name__
    # pin the oracle on the same table
    tb = copy.deepcopy(dp)
    py_oracle.fill_in_place(tb, s1, s2, cm, o)
    assert [[list(c) for c in row] for row in tb] == rec["table"]
    try:
        ops, cost = py_oracle.walk(tb, s1, s2, o)
        got = [*py_oracle.render(ops, s1, s2), cost]
    except (ValueError, AssertionError):
        got = None
    if "error" in rec:
        assert got is None, (s1, s2, got)    # the oracle refuses where the reference raises
    if "out" in rec:
        # where the reference wrapped to negative indices and still returned something, the
        # oracle (and the GPU path) refuse: those cases are recorded as "wrapped"
        if got is None:
            rec["wrapped"] = True
        else:
            assert got == rec["out"], (s1, s2, got, rec["out"])
    return rec


def main():
    rng = random.Random(20261019)
    cases = []
    # (1) the reference's own hand-bounded table
    import tests.globaligner_test as T
    fwd = None
    for mark in T.test_dp_array_forward.pytestmark:
        if mark.name == "parametrize":
            fwd = (mark.args[1] if len(mark.args) > 1 else mark.kwargs["argvalues"])[0]
    dp_array, s1, s2, cm, o, _expected = fwd
    cases.append(run_case(s1, s2, cm, o, copy.deepcopy(dp_array)))

    # (2) random tables: make_dp_array's boundary, then hand-like edits
    n_wrapped = n_err = 0
    while len(cases) < 400:
        alpha = rng.choice(["ACGT", "AG", "ACGTN"])
        m, n = rng.randint(1, 14), rng.randint(1, 14)
        s1 = "".join(rng.choice(alpha) for _ in range(m))
        s2 = "".join(rng.choice(alpha) for _ in range(n)) if rng.random() < 0.5 else \
            "".join(ch if rng.random() < 0.8 else rng.choice(alpha) for ch in (s1 * 15)[:n])
        letters = sorted(set(s1 + s2))
        mm, gh, gv = rng.randint(1, 6), rng.randint(0, 4), rng.randint(0, 4)
        cm = {x: {y: (0 if x == y else mm) for y in letters + ["-"]} for x in letters + ["-"]}
        for x in letters:
            cm["-"][x] = gh
            cm[x]["-"] = gv
        o = rng.randint(0, 5)
        dp = G.make_dp_array(s1, s2, cm, S.get_max_val(cm), o)
        kind = rng.random()
        if kind < 0.25:
            pass                                         # standard boundary
        elif kind < 0.5:                                 # origin like the reference's test: (0, k, k)
            k = rng.randint(1, 9)
            dp[0][0] = (0, k, k)
        else:                                            # arbitrary edits of boundary cells
            for _ in range(rng.randint(1, 5)):
                if rng.random() < 0.5:
                    j = rng.randint(0, n)
                    dp[0][j] = tuple(max(0, v + rng.randint(-4, 6)) for v in dp[0][j])
                else:
                    i = rng.randint(0, m)
                    dp[i][0] = tuple(max(0, v + rng.randint(-4, 6)) for v in dp[i][0])
        rec = run_case(s1, s2, cm, o, dp)
        n_wrapped += bool(rec.get("wrapped"))
        n_err += "error" in rec
        cases.append(rec)
    (HERE / "seam_cases.json").write_text(json.dumps(cases, separators=(",", ":")))
    print(f"{len(cases)} cases, {n_err} where the reference raises, {n_wrapped} where it wraps to negative indices")


if __name__ == "__main__":
    main()
