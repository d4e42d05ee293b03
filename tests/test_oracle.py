"""The oracle (oracle/py_oracle.py, oracle/ga_oracle.c) against vectors frozen from the REAL
reference by tests/golden/make_golden.py.  CPU only."""
import random

import numpy as np
import pytest

from globalign_b200 import start, synth
from oracle import c_oracle, py_oracle
from util import DenseView, fix_paths, golden, rescore


def _scheme(kw):
    a = start.validate_and_transform_args(**fix_paths(kw))
    return a[0], a[1], a[2], a[3], a[5]          # s1, s2, scoring_mat, costing_mat, gap_open_cost


def _check_case(rec, use_py=True):
    s1, s2, sm, cm, o = _scheme(rec["kw"])
    exp = rec["out"]
    ops, cost, rc = c_oracle.align(s1, s2, cm, o)
    assert rc == 0 and cost == exp[3]
    assert list(py_oracle.render(list(ops), s1, s2)) == exp[:3]
    assert py_oracle.cost_to_score(cost, len(s1), len(s2), py_oracle.max_entry(sm)) == exp[4]
    d = c_oracle.Dense(cm, o)
    e1, e2 = d.encode(s1), d.encode(s2)
    assert c_oracle.verify_path(e1, e2, d, ops) == (0, cost)
    assert c_oracle.cost_only(e1, e2, d) == cost
    assert rescore(ops, e1, e2, DenseView(d)) == cost
    if use_py:
        a1, mid, a2, pcost, pops = py_oracle.align(s1, s2, cm, o)
        assert [a1, mid, a2, pcost] == exp[:4] and pops == list(ops)


def test_reference_pytest_cases():
    for rec in golden("kat.json")["pytest_cases"]:
        _check_case(rec)
        assert rec["out"][3] == rec["expected_cost"] and rec["out"][4] == rec["expected_score"]


def test_tutorial_blosum_custom():
    kat = golden("kat.json")
    for key in ("tutorial", "blosum", "custom"):
        for rec in kat[key]:
            _check_case(rec)
    # the two alignments printed in reference/tutorial.qmd:15-22 and :141-148
    assert kat["tutorial"][0]["out"][:3] == ["ACGT", "| ||", "A-GT"]
    assert kat["tutorial"][1]["out"][:3] == ["CCTGAA", "*|*| |", "GCCG-A"]
    assert kat["tutorial"][0]["out"][3:] == [7, 0] and kat["tutorial"][1]["out"][3:] == [8, -2]


def test_dp_array_forward_table():
    f = golden("kat.json")["dp_array_forward"]
    exp = f["expected"]
    t = c_oracle.fill_table(f["seq_1"], f["seq_2"], f["costing_mat"], f["gap_open_cost"], boundary=exp)
    assert t.tolist() == exp
    tb = [[tuple(exp[0][j]) for j in range(3)], [tuple(exp[1][0]), None, None], [tuple(exp[2][0]), None, None]]
    py_oracle.fill_in_place(tb, f["seq_1"], f["seq_2"], f["costing_mat"], f["gap_open_cost"])
    assert [[list(c) for c in row] for row in tb] == exp


def test_random_cases():
    cases = golden("random_cases.json")
    assert len(cases) >= 2000
    for k, rec in enumerate(cases):
        _check_case(rec, use_py=(k % 4 == 0))


def test_medium_cases():
    for rec in golden("medium_cases.json"):
        _check_case(rec, use_py=False)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg5"])
def test_workload_prefix(name):
    rows = golden("workload_prefix.json")[name]
    batch = synth.workload(name, n_pairs=len(rows))
    for k, row in enumerate(rows):
        s1, s2 = batch.pair(k)
        _check_case({"kw": dict(seq_1=s1, seq_2=s2, **synth.SCHEMES[name]), "out": row["out"]},
                    use_py=(name in ("cfg2", "cfg5") and k < 4))


def test_seeded_replay_of_unpatched_reference():
    """Reproduce seeded runs of the randomly tie-breaking reference."""
    for rec in golden("seeded_cases.json"):
        s1, s2, sm, cm, o = _scheme(rec["kw"])
        table = py_oracle.fill(s1, s2, cm, o)
        ops, cost = py_oracle.walk(table, s1, s2, o, py_oracle.SeededChooser(random.Random(rec["seed"]), s1, s2))
        assert list(py_oracle.render(ops, s1, s2)) + [cost] == rec["out"][:4]
        assert py_oracle.admissible(ops, s1, s2, cm, o, table)
        # the canonical alignment is admissible too
        cops, _, _ = c_oracle.align(s1, s2, cm, o)
        assert py_oracle.admissible(list(cops), s1, s2, cm, o, table)


def test_out_of_domain_matches_reference_or_flags():
    """max(m,n) <= gap_open_cost: the reference raises IndexError, wraps around, or still works.
    The oracle must flag the first two and reproduce the third."""
    same = flagged = 0
    for rec in golden("out_of_domain.json"):
        kw = rec["kw"]
        s1, s2, sm, cm, o = _scheme(kw)
        ops, cost, rc = c_oracle.align(s1, s2, cm, o)
        if "exc" in rec:
            assert rc == 1
            flagged += 1
        elif rc == 0:
            assert list(py_oracle.render(list(ops), s1, s2)) + [cost] == rec["ok"][:4]
            same += 1
        else:
            flagged += 1            # reference wrapped to negative indices and returned garbage
    assert same > 200 and flagged > 50


def test_verify_path_rejects_non_canonical():
    rng = random.Random(5)
    hits = 0
    for _ in range(200):
        s1 = "".join(rng.choice("AC") for _ in range(rng.randint(3, 12)))
        s2 = "".join(rng.choice("AC") for _ in range(rng.randint(3, 12)))
        cm = start.create_costing_mat(["A", "C"], 1, 1)
        table = py_oracle.fill(s1, s2, cm, 0)
        ops, cost = py_oracle.walk(table, s1, s2, 0, lambda opts: opts[-1])   # last-listed ties
        cops, _, _ = c_oracle.align(s1, s2, cm, 0)
        d = c_oracle.Dense(cm, 0)
        rc, _ = c_oracle.verify_path(d.encode(s1), d.encode(s2), d, np.array(ops, dtype=np.uint8))
        assert (rc == 0) == (ops == list(cops))
        hits += rc != 0
    assert hits > 50


def test_float_penalty_golden_vectors():
    """float costs reach the reference's inner functions only; the oracle
    reproduces the reference's doubles bit for bit (SHA-256 of the table's IEEE-754 bytes) and its
    alignments, ranking the reference's exact sums (a common addend can change float ranks)."""
    import hashlib
    import struct

    cases = golden("float_cases.json")
    assert len(cases) >= 200
    for r in cases:
        tab = py_oracle.fill(r["seq_1"], r["seq_2"], r["cm"], r["o"])
        h = hashlib.sha256()
        for row in tab:
            for cell in row:
                h.update(struct.pack("<3d", *(float(v) for v in cell)))
        assert h.hexdigest() == r["table_sha256"]
        for i, j, vals in r["cells"]:
            assert [float(v) for v in tab[i][j]] == vals          # tolerance: none (bit-exact)
        ops, cost = py_oracle.walk(tab, r["seq_1"], r["seq_2"], r["o"], cm=r["cm"])
        assert list(py_oracle.render(ops, r["seq_1"], r["seq_2"])) + [float(cost)] == r["out"]


def test_walk_on_caller_provided_tables():
    """tests/golden/seam_cases.json: tables with hand-written / edited boundary values, filled and
    walked by the reference itself (make_seam_golden.py), incl. the table of the reference's own
    test (tests/globaligner_test.py:6-37).  The oracle's fill + value-comparing walk reproduce
    them; where the reference raises, the oracle refuses."""
    cases = golden("seam_cases.json")
    assert cases[0]["seq_1"] == "AG" and cases[0]["row0"][0] == [0, 7, 7]
    n_nonstandard = 0
    for r in cases:
        s1, s2, cm, o = r["seq_1"], r["seq_2"], r["cm"], r["o"]
        m, n = len(s1), len(s2)
        tb = [[None] * (n + 1) for _ in range(m + 1)]
        for j in range(n + 1):
            tb[0][j] = tuple(r["row0"][j])
        for i in range(m + 1):
            tb[i][0] = tuple(r["col0"][i])
        std = py_oracle.boundary_table(s1, s2, cm, o)
        n_nonstandard += any(std[0][j] != tb[0][j] for j in range(n + 1)) or \
            any(std[i][0] != tb[i][0] for i in range(m + 1))
        py_oracle.fill_in_place(tb, s1, s2, cm, o)
        assert [[list(c) for c in row] for row in tb] == r["table"]
        if "error" in r:
            with pytest.raises(ValueError):
                py_oracle.walk(tb, s1, s2, o)
            continue
        ops, cost = py_oracle.walk(tb, s1, s2, o)
        assert [*py_oracle.render(ops, s1, s2), cost] == r["out"]
    assert n_nonstandard > 200
