"""The long-pair path (ga_align_long: whole-chip band pipeline + checkpointed linear-space
traceback) against the oracle.  Bit-exact costs and canonical alignments."""
import os
import random

import numpy as np
import pytest

from globalign_b200 import _native, batch, start, synth
from globalign_b200 import globaligner
from globalign_b200.globaligner import find_global_alignment
from oracle import c_oracle
from test_gpu_parity import _oracle_dense, _simple_dense, _table_dense
from util import DenseView, golden, rescore

pytestmark = pytest.mark.gpu


def _check(ctx, dense, s1, s2, full_oracle=True):
    cost, ops, status, info = ctx.align_long(s1, s2)
    # the sequential traceback (GA_F_SERIAL) and the band-parallel one must agree exactly
    cost_s, ops_s, status_s, info_s = ctx.align_long(s1, s2, flags=_native.GA_F_SERIAL)
    assert (cost, status) == (cost_s, status_s) and info_s["trace_mode"] == 0
    if status == 0:
        assert np.array_equal(ops, ops_s)
    if len(s1) >= 2 and len(s2) >= 2:
        # pairs of a few bands take the sequential form by default: force the band-parallel one too
        os.environ["GA_LONG_SEQ_BANDS"] = "0"
        try:
            cost_p, ops_p, status_p, info_p = ctx.align_long(s1, s2)
        finally:
            os.environ.pop("GA_LONG_SEQ_BANDS", None)
        assert info_p["trace_mode"] in (1, 2) and (cost_p, status_p) == (cost, status)
        if status == 0:
            assert np.array_equal(ops_p, ops)
        if len(s1) > 6 * info["tile_rows"]:
            assert info["trace_mode"] in (1, 2)
    d = _oracle_dense(dense)
    if full_oracle:
        o_ops, o_cost, rc = c_oracle.align_codes(s1, s2, d)
        assert cost == o_cost and status == rc
        if rc == 0:
            assert np.array_equal(ops, o_ops)
    else:
        rc, vcost = c_oracle.verify_path(s1, s2, d, ops)
        assert rc == 0 and vcost == cost and status == 0
    return cost, ops, info


def test_long_path_random_sizes_simple_and_table():
    rng = random.Random(77)
    ctx = _native.Context(0)
    for trial in range(90):
        table = trial % 3 == 2
        if table:
            A = rng.choice([4, 20, 24])
            dense = _table_dense(rng, A, rng.randint(0, 11))
        else:
            A = rng.choice([2, 4, 5])
            dense = _simple_dense(A, rng.randint(1, 6), rng.randint(1, 3), rng.randint(1, 3), rng.randint(0, 12))
        ctx.set_scheme(dense)
        # around the band edges of every tuning build (128-, 192-, 256-, 288-row bands)
        m = rng.choice([1, 2, 7, 127, 128, 129, 191, 192, 193, 255, 256, 257, 300, 383, 384, 385, 511, 513, 577, 1000, 1300])
        n = rng.choice([1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 700, 1100])
        s1 = np.array([rng.randrange(A) for _ in range(m)], dtype=np.uint8)
        if rng.random() < 0.6 and m > 8:
            keep = np.array([rng.random() > 0.08 for _ in range(m)])
            s2 = s1[keep][:n].copy()
            flip = np.array([rng.random() < 0.1 for _ in range(len(s2))])
            s2[flip] = [rng.randrange(A) for _ in range(int(flip.sum()))]
            if len(s2) == 0:
                s2 = np.array([0], dtype=np.uint8)
        else:
            s2 = np.array([rng.randrange(A) for _ in range(n)], dtype=np.uint8)
        _check(ctx, dense, s1, s2)
    ctx.close()


def test_long_path_large_alphabet_simple_scheme_and_per_letter_gaps():
    """The fill's two forms: shifted substitution table in shared memory (A <= 64, any scheme,
    per-letter gap costs absorbed by the shift) and compare-and-select (simple schemes over more
    letters than the table holds)."""
    rng = random.Random(78)
    ctx = _native.Context(0)
    for A, dense in ((100, _simple_dense(100, 4, 1, 2, 3)), (64, _simple_dense(64, 2, 3, 1, 0)),
                     (24, _table_dense(rng, 24, 9)), (3, _table_dense(rng, 3, 0))):
        ctx.set_scheme(dense)
        for m, n in ((700, 650), (193, 4000), (1500, 129)):
            s1 = np.array([rng.randrange(A) for _ in range(m)], dtype=np.uint8)
            s2 = np.array([(int(s1[k % m]) if rng.random() < 0.85 else rng.randrange(A)) for k in range(n)], dtype=np.uint8)
            _check(ctx, dense, s1, s2)
    ctx.close()


def test_long_path_equals_batched_kernel_and_golden_cfg1():
    row = golden("workload_prefix.json")["cfg1"][0]
    b = synth.workload("cfg1")
    scheme = batch.Scheme(**synth.SCHEMES["cfg1"])
    dense = scheme.dense_for(list("ACGT"))
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    cost, ops, info = _check(ctx, dense, b.s1, b.s2)
    s1, s2 = b.pair(0)
    assert list(batch.render(ops, s1, s2)) + [cost] == row["out"][:4]
    bc, bops, boff, blen, bst = ctx.align_batch(b.s1, b.off1, b.s2, b.off2)
    assert bc[0] == cost and np.array_equal(_native.unpack_ops(bops, 0, int(blen[0])), ops)
    assert info["trace_mode"] == 1 and info["retraced_tiles"] >= 4      # 1 kb x 1 kb: six bands, band-parallel traceback
    os.environ["GA_LONG_SEQ_BANDS"] = "6"                               # ... and the sequential form on the same pair
    try:
        cost_p, ops_p, status_p, info_p = ctx.align_long(b.s1, b.s2)
    finally:
        os.environ.pop("GA_LONG_SEQ_BANDS", None)
    assert info_p["trace_mode"] == 0 and cost_p == cost and np.array_equal(ops_p, ops)
    ctx.close()


def test_find_global_alignment_routes_long_pairs(monkeypatch):
    rec = golden("medium_cases.json")[1]            # 513 x 257 DNA
    monkeypatch.setattr(globaligner, "LONG_PAIR_CELLS", 1000)
    r = find_global_alignment(**rec["kw"])
    assert [r.seq_1_aligned, r.middle_part, r.seq_2_aligned, r.cost, r.score] == rec["out"]
    prot = golden("medium_cases.json")[3]           # protein, BLOSUM62
    r = find_global_alignment(**prot["kw"])
    assert [r.seq_1_aligned, r.middle_part, r.seq_2_aligned, r.cost, r.score] == prot["out"]


def test_long_pair_20kb_canonical_path():
    b = synth.dna_single(20000, 7001, 7002)
    dense = batch.Scheme(**synth.SCHEMES["cfg4"]).dense_for(list("ACGT"))
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    cost, ops, info = _check(ctx, dense, b.s1, b.s2[:18000].copy(), full_oracle=False)
    assert rescore(ops, b.s1, b.s2[:18000], DenseView(dense)) == cost
    only, _, _, _ = ctx.align_long(b.s1, b.s2[:18000].copy(), flags=_native.GA_F_COST_ONLY)
    assert only == cost
    ctx.close()


def test_config4_100kb_linear_space():
    """BASELINE configs[3]: one DNA pair 100 kb x 100 kb.  The reference cannot run it
    (start.py:213 guard; ~1.6 TB of tuples); parity is against the linear-memory C oracle,
    which re-decides every step of the returned path on forward values (canonical ties)."""
    b = synth.workload("cfg4")
    dense = batch.Scheme(**synth.SCHEMES["cfg4"]).dense_for(list("ACGT"))
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    cost, ops, status, info = ctx.align_long(b.s1, b.s2)
    assert status == 0 and info["cells"] == 10**10 and info["trace_mode"] == 1
    assert info["checkpoint_bytes"] < 2.2e9            # vs 5 GB of stored direction codes
    assert rescore(ops, b.s1, b.s2, DenseView(dense)) == cost
    rc, vcost = c_oracle.verify_path(b.s1, b.s2, _oracle_dense(dense), ops)
    assert rc == 0 and vcost == cost
    # through the reference-facing entry point, with the length guard lifted
    s1, s2 = b.pair(0)
    with pytest.raises(RuntimeError):
        find_global_alignment(seq_1=s1, seq_2=s2, **synth.SCHEMES["cfg4"])
    with start.lifted_length_guard():
        r = find_global_alignment(seq_1=s1, seq_2=s2, **synth.SCHEMES["cfg4"])
    assert r.cost == cost and (r.seq_1_aligned, r.middle_part, r.seq_2_aligned) == batch.render(ops, s1, s2)
    ctx.close()


def test_second_corridor_attempt_catches_a_path_off_the_narrow_corridor():
    """The band-parallel traceback tries a narrow corridor around the straight diagonal first and a
    wide one only if the path leaves it (both on the device, no host round trip).  A 350-letter
    insertion pushes the path out of the first (192 + (m+n)/1024 columns) but not the second."""
    rng = np.random.default_rng(6)
    ctx = _native.Context(0)
    dense = _simple_dense(4, 3, 1, 1, 2)
    ctx.set_scheme(dense)
    s1 = rng.integers(0, 4, 6000).astype(np.uint8)
    s2 = np.concatenate([rng.integers(0, 4, 350).astype(np.uint8), s1[:5650]])
    cost, ops, status, info = ctx.align_long(s1, s2)
    assert status == 0 and info["trace_mode"] == 1
    rc, vcost = c_oracle.verify_path(s1, s2, _oracle_dense(dense), ops)
    assert rc == 0 and vcost == cost
    os.environ["GA_LONG_CORRIDORS"] = "1"                  # the wide corridor alone gives the same path
    try:
        cost_w, ops_w, status_w, info_w = ctx.align_long(s1, s2)
    finally:
        os.environ.pop("GA_LONG_CORRIDORS", None)
    assert (cost_w, status_w, info_w["trace_mode"]) == (cost, 0, 1) and np.array_equal(ops_w, ops)
    ctx.close()


def test_parallel_trace_falls_back_outside_the_corridor():
    """Unrelated sequences of very different composition push the path far off the diagonal;
    whichever form runs, the result is the canonical alignment."""
    rng = np.random.default_rng(5)
    ctx = _native.Context(0)
    dense = _simple_dense(4, 3, 1, 1, 2)
    ctx.set_scheme(dense)
    modes = set()
    for trial in range(6):
        m, n = int(rng.integers(3000, 6000)), int(rng.integers(3000, 6000))
        s1 = rng.integers(0, 4, m).astype(np.uint8)
        s2 = rng.integers(0, 4, n).astype(np.uint8)
        if trial % 2 == 0:          # a long insertion at one end: the path hugs an edge
            s2 = np.concatenate([rng.integers(0, 4, 2500).astype(np.uint8), s1[: n - 2500]])
        cost, ops, status, info = ctx.align_long(s1, s2)
        modes.add(info["trace_mode"])
        rc, vcost = c_oracle.verify_path(s1, s2, _oracle_dense(dense), ops)
        assert rc == 0 and vcost == cost and status == 0
    assert modes <= {1, 2}
    ctx.close()
