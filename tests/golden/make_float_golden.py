"""Golden vectors for float-valued penalties -> float_cases.json.

The reference's entry point int()-coerces every score/cost, so floats reach the DP only when
its inner functions are called directly (type hints globaligner.py:371,761).  This script
imports the reference from /root/reference/src (read-only, never copied), calls
make_dp_array / dp_array_forward / dp_array_backward with float costing matrices under the
canonical tie patch (random.choice -> first option) and records, per case, the inputs, the
alignment, the cost, and the full value table as a SHA-256 of its IEEE-754 bytes plus a few
sampled cells.  It also checks oracle/py_oracle.py against the reference on every case.
json round-trips doubles exactly (repr).   Run:  python tests/golden/make_float_golden.py
"""

from __future__ import annotations

import hashlib
import json
import random
import struct
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/src")

import globalign.globaligner as G          # noqa: E402  (the reference)
import globalign.start as S                # noqa: E402
from oracle import py_oracle               # noqa: E402

G.random.choice = lambda seq: seq[0]       # canonical tie order


def table_digest(dp):
    h = hashlib.sha256()
    for row in dp:
        for cell in row:
            h.update(struct.pack("<3d", *(float(v) for v in cell)))
    return h.hexdigest()


def random_cm(rng, alpha, kind):
    """kind 'quarter': multiples of 0.25 (exact sums, many ties); 'tenth': multiples of 0.1
    (inexact sums, rounding-dependent ranks); 'uniform': arbitrary doubles (no ties)."""
    def draw(lo, hi):
        if kind == "quarter":
            return rng.randint(int(lo * 4), int(hi * 4)) / 4.0
        if kind == "tenth":
            return rng.randint(int(lo * 10), int(hi * 10)) / 10.0
        return rng.uniform(lo, hi)
    keys = list(alpha) + ["-"]
    cm = {x: {y: 0.0 for y in keys} for x in keys}
    per_letter = rng.random() < 0.5
    gh, gv = draw(0.5, 3), draw(0.5, 3)
    for x in alpha:
        for y in alpha:
            cm[x][y] = 0.0 if x == y else draw(0.5, 5)
        cm["-"][x] = draw(0.5, 3) if per_letter else gh
        cm[x]["-"] = draw(0.5, 3) if per_letter else gv
    cm["-"]["-"] = max(gh, gv)
    return cm, draw(0.0, 4)


def main():
    rng = random.Random(20260218)
    cases = []
    for k in range(240):
        alpha = rng.choice(["AC", "ACGT", "ACGT", "ACDEFGHIK"])
        kind = ("quarter", "tenth", "uniform")[k % 3]
        cm, o = random_cm(rng, alpha, kind)
        hi = 60 if k % 8 else 200
        m, n = rng.randint(1, hi), rng.randint(1, hi)
        s1 = "".join(rng.choice(alpha) for _ in range(m))
        if rng.random() < 0.6:
            s2 = "".join(ch if rng.random() < 0.85 else rng.choice(alpha) for ch in s1)[:max(1, n)]
            if len(s2) > 3 and rng.random() < 0.5:
                cut = rng.randrange(len(s2) - 1)
                s2 = s2[:cut] + s2[cut + rng.randint(1, 2):]
            s2 = s2 or rng.choice(alpha)
        else:
            s2 = "".join(rng.choice(alpha) for _ in range(n))
        max_cost = S.get_max_val(cm)
        rec = {"seq_1": s1, "seq_2": s2, "cm": cm, "o": o, "max_cost": max_cost, "kind": kind}
        dp = G.make_dp_array(s1, s2, cm, max_cost, o)
        G.dp_array_forward(dp, s1, s2, cm, o)
        rec["table_sha256"] = table_digest(dp)
        pts = [(len(s1), len(s2))] + [(rng.randint(0, len(s1)), rng.randint(0, len(s2))) for _ in range(6)]
        rec["cells"] = [[i, j, [float(v) for v in dp[i][j]]] for i, j in pts]
        try:
            a1, mid, a2, cost = G.dp_array_backward(dp, s1, s2, cm, o)
            rec["out"] = [a1, mid, a2, float(cost)]
        except IndexError:
            rec["exc"] = "IndexError"
        # oracle pin: same table, same walk
        tab = py_oracle.fill(s1, s2, cm, o)
        assert table_digest(tab) == rec["table_sha256"], "py_oracle table differs from the reference"
        try:
            ops, ocost = py_oracle.walk(tab, s1, s2, o, cm=cm)
            got = list(py_oracle.render(ops, s1, s2)) + [float(ocost)]
        except ValueError:
            got = None
        if "out" in rec:
            # where the reference wraps to negative indices and still returns, the oracle flags it
            assert got is None or got == rec["out"], (k, got, rec["out"])
            rec["oracle_defined"] = got is not None
        cases.append(rec)
    (HERE / "float_cases.json").write_text(json.dumps(cases))
    n_def = sum(1 for r in cases if r.get("oracle_defined"))
    print(f"float_cases.json: {len(cases)} cases, {n_def} with a defined walk")


if __name__ == "__main__":
    main()
