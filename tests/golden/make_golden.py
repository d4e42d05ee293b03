#!/usr/bin/env python3
"""Freeze golden vectors from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py

imports globalign from /root/reference/src (read-only, never copied), runs it on
 * the reference's own pytest known-answer cases (read from its test module),
 * the tutorial inputs (reference/tutorial.qmd),
 * BLOSUM and custom-matrix cases,
 * a few thousand random small cases over every scheme branch,
 * prefixes of the BASELINE.json synthetic workloads (globalign_b200.synth),
 * seeded runs of the *unpatched* (randomly tie-breaking) reference,
 * argument-error cases,
with ``globaligner.random.choice`` patched to "first listed option" (the
canonical tie order) unless stated otherwise, and writes
``tests/golden/*.json``.  While doing so it asserts that oracle/py_oracle.py and
oracle/ga_oracle.c reproduce the reference on every case: this is what pins the
oracle.  /root/reference does not exist on the GPU box; the tests read only the
JSON written here.
"""

from __future__ import annotations

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
from oracle import c_oracle, py_oracle     # noqa: E402
from globalign_b200 import synth           # noqa: E402

_REAL_CHOICE = G.random.choice


def canonical(on: bool):
    G.random.choice = (lambda seq: seq[0]) if on else _REAL_CHOICE


KW = ("input_fasta", "output", "seq_1", "seq_2", "scoring_mat_name", "scoring_mat_path",
      "match_score", "mismatch_score", "mismatch_cost", "gap_open_score", "gap_open_cost",
      "gap_extension_score", "gap_extension_cost")


def run_ref(kw: dict):
    r = G.find_global_alignment(**kw)
    return r


def out_of(r):
    return [r.seq_1_aligned, r.middle_part, r.seq_2_aligned, r.cost, r.score]


def check_oracles(kw, r):
    """Both restatements must reproduce the reference's canonical result."""
    s1 = kw["seq_1"].upper()
    s2 = kw["seq_2"].upper()
    cm, o = r.costing_mat, r.gap_open_cost
    a1, mid, a2, cost, ops = py_oracle.align(s1, s2, cm, o)
    assert (a1, mid, a2, cost) == (r.seq_1_aligned, r.middle_part, r.seq_2_aligned, r.cost), kw
    assert py_oracle.cost_to_score(cost, len(s1), len(s2), py_oracle.max_entry(r.scoring_mat)) == r.score
    cops, ccost, rc = c_oracle.align(s1, s2, cm, o)
    assert rc == 0 and ccost == r.cost and list(cops) == ops, kw
    d = c_oracle.Dense(cm, o)
    vrc, vcost = c_oracle.verify_path(d.encode(s1), d.encode(s2), d, cops)
    assert vrc == 0 and vcost == r.cost, (kw, vrc)
    return ops


def portable(kw):
    """kwargs with paths made relative to tests/golden (fixtures travel)."""
    kw = dict(kw)
    for k in ("scoring_mat_path",):
        if kw.get(k) is not None:
            kw[k] = Path(kw[k]).name
    return {k: v for k, v in kw.items() if v is not None}


# ------------------------------------------------------------------ KATs
def reference_pytest_cases():
    import tests.globaligner_test as T     # the reference's own test module
    cases = []
    for mark in T.test_find_global_alignment.pytestmark:
        if mark.name == "parametrize":
            for row in mark.args[1] if len(mark.args) > 1 else mark.kwargs["argvalues"]:
                kw = dict(zip(KW, row[:13]))
                cases.append((kw, row[13], row[14]))
    fwd = None
    for mark in T.test_dp_array_forward.pytestmark:
        if mark.name == "parametrize":
            fwd = (mark.args[1] if len(mark.args) > 1 else mark.kwargs["argvalues"])[0]
    return cases, fwd


CUSTOM_MTX = """   A  C  G  T  N  -
A  5 -4 -1 -4 -2 -3
C -4  5 -4 -1 -2 -2
G -1 -4  5 -4 -2 -3
T -4 -1 -4  5 -2 -1
N -2 -2 -2 -2  5 -2
- -3 -2 -3 -1 -2  1
"""


def main():
    random.seed(20261018)
    canonical(True)
    (HERE / "custom_dna.mtx").write_text(CUSTOM_MTX)
    custom_path = str(HERE / "custom_dna.mtx")

    kat = {"pytest_cases": [], "tutorial": [], "blosum": [], "custom": []}
    cases, fwd = reference_pytest_cases()
    for kw, exp_score, exp_cost in cases:
        r = run_ref(kw)
        assert (r.score, r.cost) == (exp_score, exp_cost)
        check_oracles(kw, r)
        kat["pytest_cases"].append({"kw": portable(kw), "expected_score": exp_score,
                                    "expected_cost": exp_cost, "out": out_of(r)})
    dp_array, s1, s2, cm, o, expected = fwd
    G.dp_array_forward(dp_array, s1, s2, cm, o)
    assert dp_array == expected
    t = c_oracle.fill_table(s1, s2, cm, o, boundary=expected)
    assert t.tolist() == [[list(c) for c in row] for row in expected]
    tb = [[tuple(c) if c is not None else None for c in row] for row in
          [[expected[0][0], expected[0][1], expected[0][2]],
           [expected[1][0], None, None], [expected[2][0], None, None]]]
    py_oracle.fill_in_place(tb, s1, s2, cm, o)
    assert tb == [[tuple(c) for c in row] for row in expected]
    kat["dp_array_forward"] = {"seq_1": s1, "seq_2": s2, "costing_mat": cm, "gap_open_cost": o,
                               "expected": [[list(c) for c in row] for row in expected]}

    tutorial_inputs = [
        dict(seq_1="ACGT", seq_2="AGT"),
        dict(seq_1="CCTGAA", seq_2="GCCGA", match_score=1, mismatch_score=-1, gap_open_score=-2,
             gap_extension_score=-1),
    ]
    for kw in tutorial_inputs:
        r = run_ref(kw)
        check_oracles(kw, r)
        kat["tutorial"].append({"kw": kw, "out": out_of(r), "str": str(r),
                                "scoring_mat": r.scoring_mat, "costing_mat": r.costing_mat,
                                "gap_open_score": r.gap_open_score, "gap_open_cost": r.gap_open_cost})

    blosum_inputs = [
        dict(seq_1="HEAGAWGHEE", seq_2="PAWHEAE", scoring_mat_name="BLOSUM62", gap_open_cost=10),
        dict(seq_1="HEAGAWGHEE", seq_2="PAWHEAE", scoring_mat_name="BLOSUM50"),
        dict(seq_1="heagawghee", seq_2="pawheae", scoring_mat_name="BLOSUM62", gap_open_score=-3),
        dict(seq_1="MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ", seq_2="MKTAYIAKQRISFVKSHFSRQEERLGLIEVQBZXJ",
             scoring_mat_name="BLOSUM62", gap_open_cost=11),
    ]
    for kw in blosum_inputs:
        r = run_ref(kw)
        check_oracles(kw, r)
        kat["blosum"].append({"kw": kw, "out": out_of(r),
                              "max_cost": S.get_max_val(r.costing_mat),
                              "max_score": S.get_max_val(r.scoring_mat),
                              "gap_open_score": r.gap_open_score, "gap_open_cost": r.gap_open_cost})
    for kw in [dict(seq_1="ACGTNNACGT", seq_2="ACTTNAGT", scoring_mat_path=custom_path),
               dict(seq_1="GATTACAGATTACA", seq_2="GCATGCTGCATACA", scoring_mat_path=custom_path,
                    gap_open_cost=3)]:
        r = run_ref(kw)
        check_oracles(kw, r)
        kat["custom"].append({"kw": portable(kw), "out": out_of(r), "costing_mat": r.costing_mat,
                              "scoring_mat": r.scoring_mat})

    mats = {}
    for name in ("BLOSUM50", "BLOSUM62", "nucleotide"):
        mats[name] = S.read_scoring_mat(REF / "src/globalign/data/scoring_matrices" / f"{name}.mtx")
    kat["scoring_matrices"] = mats
    (HERE / "kat.json").write_text(json.dumps(kat, indent=1))

    # -------------------------------------------------------------- random cases
    aminos = "ARNDCQEGHILKMFPSTWYVBJZX"
    rnd = []

    def rseq(alpha, lo, hi):
        return "".join(random.choice(alpha) for _ in range(random.randint(lo, hi)))

    def mutate(s, alpha):
        out = []
        for ch in s:
            u = random.random()
            if u < 0.12:
                out.append(random.choice(alpha))
            elif u < 0.18:
                continue
            elif u < 0.24:
                out.append(ch)
                out.append(random.choice(alpha))
            else:
                out.append(ch)
        return "".join(out) or random.choice(alpha)

    def gen_case():
        branch = random.choice(["default", "scores", "scores", "costs", "blosum", "custom"])
        hi = random.choice([4, 8, 16, 30, 48])
        if branch == "blosum":
            alpha = aminos
        elif branch == "custom":
            alpha = "ACGTN"
        else:
            alpha = random.choice(["AC", "ACGT", "ACGT", "acgtACGT", "ACGTN", "XYZéβ"])
        s1 = rseq(alpha, 1, hi)
        s2 = mutate(s1, alpha) if random.random() < 0.7 else rseq(alpha, 1, hi)
        kw = dict(seq_1=s1, seq_2=s2)
        if branch == "scores":
            for k, lo_, hi_ in (("match_score", 1, 5), ("mismatch_score", -6, -1),
                                ("gap_open_score", -12, 0), ("gap_extension_score", -4, -1)):
                if random.random() < 0.8:
                    v = random.randint(lo_, hi_)
                    kw[k] = str(v) if random.random() < 0.1 else v
        elif branch == "costs":
            for k, lo_, hi_ in (("mismatch_cost", 1, 6), ("gap_open_cost", 0, 8),
                                ("gap_extension_cost", 1, 4)):
                if random.random() < 0.8:
                    kw[k] = random.randint(lo_, hi_)
            if len(kw) == 2:
                kw["mismatch_cost"] = 4
        elif branch == "blosum":
            kw["scoring_mat_name"] = random.choice(["BLOSUM62", "BLOSUM50"])
            u = random.random()
            if u < 0.5:
                kw["gap_open_cost"] = random.randint(0, 12)
            elif u < 0.8:
                kw["gap_open_score"] = -random.randint(0, 12)
        elif branch == "custom":
            kw["scoring_mat_path"] = custom_path
            if random.random() < 0.6:
                kw["gap_open_cost"] = random.randint(0, 6)
        return kw

    n_dom = 0
    while n_dom < 2500:
        kw = gen_case()
        try:
            r = run_ref(kw)
        except Exception:
            continue
        if max(len(kw["seq_1"]), len(kw["seq_2"])) <= r.gap_open_cost:
            continue            # outside the parity domain
        check_oracles(kw, r)
        rnd.append({"kw": portable(kw), "out": out_of(r)})
        n_dom += 1
    (HERE / "random_cases.json").write_text(json.dumps(rnd, ensure_ascii=True))

    # -------------------------------------------------------------- medium cases
    medium = []
    for (m, n, alpha, extra) in [(300, 280, "ACGT", dict(match_score=2, mismatch_score=-3,
                                                          gap_open_score=-10, gap_extension_score=-1)),
                                 (513, 257, "ACGT", dict()),
                                 (257, 600, "AC", dict(gap_open_cost=0, mismatch_cost=1,
                                                       gap_extension_cost=1)),
                                 (220, 200, aminos[:20], dict(scoring_mat_name="BLOSUM62",
                                                              gap_open_cost=10)),
                                 (129, 400, aminos[:20], dict(scoring_mat_name="BLOSUM50"))]:
        s1 = rseq(alpha, m, m)
        s2 = (mutate(s1, alpha) + rseq(alpha, n, n))[:n]
        kw = dict(seq_1=s1, seq_2=s2, **extra)
        r = run_ref(kw)
        check_oracles(kw, r)
        medium.append({"kw": kw, "out": out_of(r)})
    (HERE / "medium_cases.json").write_text(json.dumps(medium))

    # -------------------------------------------------------------- workload prefixes
    wl = {}
    for name, count in (("cfg1", 1), ("cfg2", 24), ("cfg3", 3), ("cfg5", 8)):
        batch = synth.workload(name, n_pairs=max(count, 1))
        rows = []
        for k in range(count):
            s1, s2 = batch.pair(k)
            kw = dict(seq_1=s1, seq_2=s2, **synth.SCHEMES[name])
            r = run_ref(kw)
            check_oracles(kw, r)
            rows.append({"out": out_of(r)})
            print(name, k, r.cost, r.score, flush=True)
        wl[name] = rows
    (HERE / "workload_prefix.json").write_text(json.dumps(wl))

    # -------------------------------------------------------------- seeded, unpatched reference
    canonical(False)
    seeded = []
    for k in range(300):
        kw = gen_case()
        seed = random.randint(0, 10**6)
        state = random.getstate()
        try:
            random.seed(seed)
            r = run_ref(kw)
            next_random = random.random()      # fingerprint of the generator state the reference leaves
        except Exception:
            random.setstate(state)
            continue
        random.setstate(state)
        if max(len(kw["seq_1"]), len(kw["seq_2"])) <= r.gap_open_cost:
            continue
        s1, s2 = kw["seq_1"].upper(), kw["seq_2"].upper()
        table = py_oracle.fill(s1, s2, r.costing_mat, r.gap_open_cost)
        chooser = py_oracle.SeededChooser(random.Random(seed), s1, s2)
        ops, cost = py_oracle.walk(table, s1, s2, r.gap_open_cost, chooser)
        got = py_oracle.render(ops, s1, s2)
        assert got == (r.seq_1_aligned, r.middle_part, r.seq_2_aligned) and cost == r.cost, kw
        assert py_oracle.admissible(ops, s1, s2, r.costing_mat, r.gap_open_cost, table)
        seeded.append({"kw": portable(kw), "seed": seed, "out": out_of(r), "next_random": next_random})
    canonical(True)
    (HERE / "seeded_cases.json").write_text(json.dumps(seeded, ensure_ascii=True))

    # -------------------------------------------------------------- error / edge behaviour
    errs = []
    tmp_existing = HERE / "custom_dna.mtx"
    for kw in [
        dict(seq_1="ACGT"), dict(seq_2="ACGT"), dict(), dict(seq_1="", seq_2="ACGT"),
        dict(seq_1="AC-GT", seq_2="ACGT"), dict(seq_1="ACGT", seq_2="ACGT", match_score=0),
        dict(seq_1="ACGT", seq_2="ACGT", mismatch_score=1),
        dict(seq_1="ACGT", seq_2="ACGT", gap_open_score=1),
        dict(seq_1="ACGT", seq_2="ACGT", gap_extension_score=0),
        dict(seq_1="ACGT", seq_2="ACGT", match_score="x"),
        dict(seq_1="ACGT", seq_2="ACGT", mismatch_cost=0),
        dict(seq_1="ACGT", seq_2="ACGT", gap_open_cost=-1),
        dict(seq_1="ACGT", seq_2="ACGT", gap_extension_cost=0),
        dict(seq_1="ACGT", seq_2="ACGT", mismatch_cost="5"),
        dict(seq_1="ACGT", seq_2="ACGT", match_score=1, mismatch_cost=2),
        dict(seq_1="ACGT", seq_2="ACGT", scoring_mat_name="BLOSUM62", match_score=1),
        dict(seq_1="ACGT", seq_2="ACGT", scoring_mat_name="BLOSUM62", scoring_mat_path="x.mtx"),
        dict(seq_1="ACGU", seq_2="ACGT", scoring_mat_name="BLOSUM62"),
        dict(seq_1="ACGT", seq_2="ACGT", scoring_mat_path="does_not_exist.mtx"),
        dict(seq_1="ACGT", seq_2="ACGT", scoring_mat_path=custom_path, gap_extension_cost=2),
        dict(seq_1="ACGT", seq_2="ACGT", output=str(tmp_existing)),
        dict(seq_1="ACGT", seq_2="ACGT", output="/nonexistent_dir_xyz/out.txt"),
        dict(seq_1="ACGT", seq_2="ACGT", input_fasta="nope.fa"),
        dict(input_fasta="nope.fa"),
        dict(seq_1="A" * 5000, seq_2="C" * 4000),
        dict(seq_1="ACGT", seq_2="ACGT", gap_open_score=-3, gap_open_cost=3),
        dict(seq_1="ACGT", seq_2="ACGT", scoring_mat_name="BLOSUM62", gap_open_score=-3, gap_open_cost=3),
    ]:
        try:
            r = run_ref(kw)
            errs.append({"kw": portable({k: (Path(v).name if k == "output" and "custom" in str(v) else v)
                                         for k, v in kw.items()}), "ok": out_of(r)})
        except BaseException as e:          # RuntimeWarning is raised as an exception
            errs.append({"kw": portable({k: (Path(v).name if k == "output" and "custom" in str(v) else v)
                                         for k, v in kw.items()}), "exc": type(e).__name__})
    (HERE / "error_cases.json").write_text(json.dumps(errs, indent=1))

    # out-of-domain: max(m, n) <= gap_open_cost -- record what the reference does
    ood = []
    for _ in range(400):
        alpha = random.choice(["AC", "ACGT"])
        s1, s2 = rseq(alpha, 1, 6), rseq(alpha, 1, 6)
        o = random.randint(max(len(s1), len(s2)), 40)
        kw = dict(seq_1=s1, seq_2=s2, gap_open_cost=o, mismatch_cost=random.randint(1, 5),
                  gap_extension_cost=random.randint(1, 3))
        try:
            r = run_ref(kw)
            rec = {"kw": kw, "ok": out_of(r)}
        except BaseException as e:
            rec = {"kw": kw, "exc": type(e).__name__}
        ood.append(rec)
    (HERE / "out_of_domain.json").write_text(json.dumps(ood))
    print("golden vectors written:", sorted(p.name for p in HERE.glob("*.json")))


if __name__ == "__main__":
    main()
