"""Parity of the CUDA path (through the C-ABI) with the oracle and with the golden vectors frozen
from the reference.  Bit-exact: costs, scores, the three alignment strings, tie order."""
import os
import random
import zlib

import numpy as np
import pytest

from globalign_b200 import _native, batch, globaligner, lowering, start, synth
from globalign_b200.globaligner import find_global_alignment
from oracle import c_oracle
from util import DenseView, fix_paths, golden, rescore

pytestmark = pytest.mark.gpu

PACKED_CFGS = ["fill_p16_r7_g8", "fill_p16_r10_g8", "fill_p16_r13_g8", "fill_p16_r16_g8", "fill_p16_r13_g16", "fill_p16_r19_g16",
               "fill_p16_r5_g32", "fill_p16_r10_g16", "fill_p16_r15_g10", "fill_p16_r19_g8",
               "fill_p16_r16_g16", "fill_p16_r8_g32", "fill_p16_r16_g32"]
I32S_CFGS = ["fill_i32s_r8_g8", "fill_i32s_r8_g16", "fill_i32s_r8_g32"]
I32T_CFGS = ["fill_i32t_r8_g8", "fill_i32t_r8_g16", "fill_i32t_r8_g32"]
P16T_CFGS = ["fill_p16t_r8_g8", "fill_p16t_r8_g16", "fill_p16t_r8_g32", "fill_p16t_r12_g32", "fill_p16t_r16_g32",
             "fill_p16b_r16_g32", "fill_p16h_r16_g32"]


@pytest.fixture(autouse=True)
def _no_forced_cfg():
    os.environ.pop("GA_FORCE_CFG", None)
    yield
    os.environ.pop("GA_FORCE_CFG", None)


def _oracle_dense(dense):
    d = c_oracle.Dense.__new__(c_oracle.Dense)
    d.A, d.sub, d.gap_h, d.gap_v, d.o, d.max_cost = (dense.A, dense.sub, dense.gap_h, dense.gap_v,
                                                     dense.gap_open, dense.max_cost)
    return d


def _compare_with_oracle(s1, off1, s2, off2, dense, out, sample=None):
    cost, ops, ops_off, ops_len, status = out
    o_ops, o_off, o_n, o_cost, o_status = c_oracle.align_batch(s1, off1, s2, off2, _oracle_dense(dense))
    assert np.array_equal(cost, o_cost)
    assert np.array_equal(status, o_status)
    idx = range(len(cost)) if sample is None else sample
    for k in idx:
        if o_status[k] == 0:
            got = _native.unpack_ops(ops, int(ops_off[k]), int(ops_len[k]))
            assert np.array_equal(got, o_ops[o_off[k]:o_off[k] + o_n[k]]), k


def _check_golden(rec):
    r = find_global_alignment(**fix_paths(rec["kw"]))
    assert [r.seq_1_aligned, r.middle_part, r.seq_2_aligned, r.cost, r.score] == rec["out"], rec["kw"]
    return r


# ------------------------------------------------------------------ golden vectors of the reference
def test_reference_known_answers():
    kat = golden("kat.json")
    for rec in kat["pytest_cases"]:
        r = _check_golden(rec)
        assert (r.score, r.cost) == (rec["expected_score"], rec["expected_cost"])
    for rec in kat["tutorial"]:
        r = _check_golden(rec)
        assert str(r) == rec["str"]
        assert r.scoring_mat == rec["scoring_mat"] and r.costing_mat == rec["costing_mat"]
    for key in ("blosum", "custom"):
        for rec in kat[key]:
            _check_golden(rec)


def test_random_golden_cases():
    for rec in golden("random_cases.json"):
        _check_golden(rec)


def test_medium_golden_cases():
    for rec in golden("medium_cases.json"):
        _check_golden(rec)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg5"])
def test_workload_prefix_golden(name):
    rows = golden("workload_prefix.json")[name]
    b = synth.workload(name, n_pairs=len(rows))
    cost, ops, ops_off, ops_len, status, scheme, dense = batch.align_encoded(b, synth.SCHEMES[name])
    m, n = b.lens()
    for k, row in enumerate(rows):
        s1, s2 = b.pair(k)
        got = batch.render(_native.unpack_ops(ops, int(ops_off[k]), int(ops_len[k])), s1, s2)
        score = int(n[k]) * (scheme.max_score // 2) + int(m[k]) * ((scheme.max_score + 1) // 2) - int(cost[k])
        assert list(got) + [int(cost[k]), score] == row["out"], (name, k)


def test_batched_string_api():
    recs = [r for r in golden("random_cases.json") if set(r["kw"]) == {"seq_1", "seq_2"}][:200]
    res = batch.find_global_alignments([r["kw"]["seq_1"] for r in recs], [r["kw"]["seq_2"] for r in recs])
    assert len(res) == len(recs)
    for k, rec in enumerate(recs):
        a = res[k]
        assert [a.seq_1_aligned, a.middle_part, a.seq_2_aligned, a.cost, a.score] == rec["out"]
        one = find_global_alignment(**rec["kw"])
        assert one.scoring_mat == a.scoring_mat and one.costing_mat == a.costing_mat
    # a fixed-length, mixed-case batch goes through the entry point without offset arrays
    rng = random.Random(3)
    s1 = ["".join(rng.choice("ACGTacgt") for _ in range(60)) for _ in range(300)]
    s2 = [x[:20] + "".join(rng.choice("ACGT") for _ in range(8)) + x[28:] for x in s1]
    kw = dict(match_score=2, mismatch_score=-3, gap_open_score=-10, gap_extension_score=-1)
    res = batch.find_global_alignments(s1, s2, **kw)
    for k in range(0, 300, 17):
        one = find_global_alignment(seq_1=s1[k], seq_2=s2[k], **kw)
        a = res[k]
        assert (a.seq_1_aligned, a.middle_part, a.seq_2_aligned, a.cost, a.score) == \
            (one.seq_1_aligned, one.middle_part, one.seq_2_aligned, one.cost, one.score)


# ------------------------------------------------------------------ the seam functions
def test_dp_array_forward_reference_table():
    f = golden("kat.json")["dp_array_forward"]
    exp = [[tuple(c) for c in row] for row in f["expected"]]
    dp = [[exp[0][0], exp[0][1], exp[0][2]], [exp[1][0], None, None], [exp[2][0], None, None]]
    globaligner.dp_array_forward(dp, f["seq_1"], f["seq_2"], f["costing_mat"], f["gap_open_cost"])
    assert dp == exp


def test_seam_functions_equal_oracle_table():
    rng = random.Random(11)
    for _ in range(20):
        s1 = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 60)))
        s2 = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 60)))
        a = start.validate_and_transform_args(seq_1=s1, seq_2=s2, gap_open_score=-rng.randint(0, 6))
        cm, o = a[3], a[5]
        dp = globaligner.make_dp_array(s1, s2, cm, start.get_max_val(cm), o)
        globaligner.dp_array_forward(dp, s1, s2, cm, o)
        t = c_oracle.fill_table(s1, s2, cm, o)
        assert [[list(c) for c in row] for row in dp] == t.tolist()
        if max(len(s1), len(s2)) > o:
            top, mid, bot, cost = globaligner.dp_array_backward(dp, s1, s2, cm, o)
            ops, ocost, rc = c_oracle.align(s1, s2, cm, o)
            assert rc == 0 and cost == ocost == min(dp[len(s1)][len(s2)])
            assert (top, mid, bot) == batch.render(ops, s1, s2)


# ------------------------------------------------------------------ every kernel configuration
def _random_batch(rng, n_pairs, A, max_m, max_n, similar=0.6):
    s1, s2, o1, o2 = [], [], [0], [0]
    for _ in range(n_pairs):
        m = rng.randint(1, max_m) if rng.random() < 0.8 else rng.randint(1, 3)
        a = [rng.randrange(A) for _ in range(m)]
        if rng.random() < similar:
            b = []
            for x in a:
                u = rng.random()
                if u < 0.1:
                    b.append(rng.randrange(A))
                elif u < 0.15:
                    continue
                elif u < 0.2:
                    b += [x, rng.randrange(A)]
                else:
                    b.append(x)
            b = b[:max_n] or [rng.randrange(A)]
        else:
            b = [rng.randrange(A) for _ in range(rng.randint(1, max_n))]
        s1 += a
        s2 += b
        o1.append(len(s1))
        o2.append(len(s2))
    return (np.array(s1, dtype=np.uint8), np.array(o1, dtype=np.int64),
            np.array(s2, dtype=np.uint8), np.array(o2, dtype=np.int64))


def _simple_dense(A, mm, gh, gv, o):
    letters = [chr(65 + k) for k in range(A)]
    cm = {x: {y: (0 if x == y else mm) for y in letters} for x in letters}
    for x in letters:
        cm[x]["-"] = gv
    cm["-"] = {y: gh for y in letters}
    cm["-"]["-"] = 0
    return lowering.DenseScheme(cm, o)


def _table_dense(rng, A, o):
    letters = [chr(65 + k) for k in range(A)]
    cm = {x: {} for x in letters + ["-"]}
    for i, x in enumerate(letters):
        for j, y in enumerate(letters):
            if j < i:
                cm[x][y] = cm[y][x]
            else:
                cm[x][y] = 0 if (x == y and rng.random() < 0.8) else rng.randint(0, 12)
        cm[x]["-"] = rng.randint(1, 9)
        cm["-"][x] = rng.randint(1, 9)
    cm["-"]["-"] = 0
    return lowering.DenseScheme(cm, o)


@pytest.mark.parametrize("cfg", PACKED_CFGS + I32S_CFGS)
def test_simple_scheme_kernels(cfg):
    rng = random.Random(zlib.crc32(cfg.encode()))
    R, G = int(cfg.split("_r")[1].split("_")[0]), int(cfg.split("_g")[1])
    band = R * G
    os.environ["GA_FORCE_CFG"] = cfg
    ctx = _native.Context(0)
    for trial, (A, mm, gh, gv, o) in enumerate([(4, 5, 2, 2, 10), (2, 1, 1, 1, 0), (4, 4, 1, 2, 3), (5, 3, 2, 1, 1)]):
        dense = _simple_dense(A, mm, gh, gv, o)
        ctx.set_scheme(dense)
        max_m = min(int(band * (2.4 if trial == 1 else 1.0)), 420)
        data = _random_batch(rng, 257 if trial else 515, A, max_m, 200)
        out = ctx.align_batch(*data)
        assert ctx.plan_info()["fill_kernel"] == cfg
        _compare_with_oracle(*data, dense, out)
    ctx.close()


@pytest.mark.parametrize("cfg", I32T_CFGS + P16T_CFGS)
def test_table_scheme_kernels(cfg):
    rng = random.Random(zlib.crc32(cfg.encode()))
    R, G = int(cfg.split("_r")[1].split("_")[0]), int(cfg.split("_g")[1])
    os.environ["GA_FORCE_CFG"] = cfg
    ctx = _native.Context(0)
    for trial in range(4):
        A = rng.choice([3, 5, 20, 24])
        dense = _table_dense(rng, A, rng.randint(0, 12))
        ctx.set_scheme(dense)
        data = _random_batch(rng, 301, A, min(int(R * G * (2.3 if trial % 2 else 1.0)), 900), 260)
        out = ctx.align_batch(*data)
        assert ctx.plan_info()["fill_kernel"] == cfg
        _compare_with_oracle(*data, dense, out)
    ctx.close()


def test_automatic_kernel_choice():
    ctx = _native.Context(0)
    dna = synth.dna_fixed(64, 150, 2000)
    sch = batch.Scheme(**synth.SCHEMES["cfg2"])
    ctx.set_scheme(sch.dense_for(list("ACGT")))
    assert ctx.plan_batch(dna.off1, dna.off2)["packed16"] is True
    assert ctx.plan_batch(dna.off1, dna.off2, flags=_native.GA_F_FORCE_GENERIC)["packed16"] is False
    long_dna = synth.dna_fixed(4, 3000, 7)          # values no longer fit 16 bits
    assert ctx.plan_batch(long_dna.off1, long_dna.off2)["packed16"] is False
    prot = synth.protein_varlen(8, 3000)
    ctx.set_scheme(batch.Scheme(**synth.SCHEMES["cfg3"]).dense_for(sorted(synth.AMINO)))
    # proteins up to 1000 residues under BLOSUM62 fit the unsigned packed table kernels ...
    # (20 letters present: the lane-replicated table still leaves three blocks per SM)
    assert ctx.plan_batch(prot.off1, prot.off2)["fill_kernel"] == "fill_p16b_r16_g32"
    assert ctx.plan_batch(prot.off1, prot.off2, flags=_native.GA_F_FORCE_GENERIC)["fill_kernel"].startswith("fill_i32t")
    # ... longer ones do not: 4 * 16 * 1100 > 65535
    off = np.array([0, 1100, 2200], dtype=np.int64)
    assert ctx.plan_batch(off, off)["fill_kernel"].startswith("fill_i32t")
    ctx.close()


def test_packed_table_kernel_at_the_edge_of_its_value_range():
    """1000 x 1000 under BLOSUM62 (gap open 10): 4 * (16000 + 50) + 3 = 64203 of 65535.  Worst
    cases for the magnitude of the values: unrelated sequences, all-W vs all-C-like mismatches
    at the highest cost, and one-letter sequences against long ones (sentinel-derived values)."""
    rng = random.Random(99)
    dense = batch.Scheme(**synth.SCHEMES["cfg3"]).dense_for(sorted(synth.AMINO))
    A = dense.A
    worst = int(np.argmax(dense.sub.reshape(A, A).max(axis=1)))
    worst_b = int(np.argmax(dense.sub.reshape(A, A)[worst]))
    seqs = [([worst] * 1000, [worst_b] * 1000),
            ([rng.randrange(A) for _ in range(1000)], [rng.randrange(A) for _ in range(1000)]),
            ([rng.randrange(A)], [rng.randrange(A) for _ in range(1000)]),
            ([rng.randrange(A) for _ in range(1000)], [rng.randrange(A)]),
            ([worst] * 1000, [worst_b] * 11),
            ([worst] * 997, [worst] * 1000),
            ([rng.randrange(A) for _ in range(513)], [rng.randrange(A) for _ in range(1000)])]
    s1, s2, o1, o2 = [], [], [0], [0]
    for a, b in seqs:
        s1 += a
        s2 += b
        o1.append(len(s1))
        o2.append(len(s2))
    data = (np.array(s1, dtype=np.uint8), np.array(o1, dtype=np.int64),
            np.array(s2, dtype=np.uint8), np.array(o2, dtype=np.int64))
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    out = ctx.align_batch(*data)
    assert ctx.plan_info()["fill_kernel"] == "fill_p16b_r16_g32"
    _compare_with_oracle(*data, dense, out)
    # the same through the plain-table kernel (what alphabets above ~21 letters get)
    os.environ["GA_FORCE_CFG"] = "fill_p16t_r16_g32"
    out = ctx.align_batch(*data)
    assert ctx.plan_info()["fill_kernel"] == "fill_p16t_r16_g32"
    _compare_with_oracle(*data, dense, out)
    ctx.close()


def test_chunked_scratch_gives_identical_results():
    b = synth.dna_fixed(3000, 150, 2000)
    dense = batch.Scheme(**synth.SCHEMES["cfg2"]).dense_for(list("ACGT"))
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    full = ctx.align_batch(b.s1, b.off1, b.s2, b.off2)
    assert ctx.plan_info()["n_chunks"] == 1
    ctx.set_scratch_limit(8 << 20)
    parts = ctx.align_batch(b.s1, b.off1, b.s2, b.off2)
    assert ctx.plan_info()["n_chunks"] > 3
    for x, y in zip(full, parts):
        assert np.array_equal(x, y)
    ctx.close()


def test_cost_only_flag():
    b = synth.dna_fixed(500, 150, 2000)
    dense = batch.Scheme(**synth.SCHEMES["cfg2"]).dense_for(list("ACGT"))
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    full = ctx.align_batch(b.s1, b.off1, b.s2, b.off2)
    only = ctx.align_batch(b.s1, b.off1, b.s2, b.off2, flags=_native.GA_F_COST_ONLY)
    assert np.array_equal(full[0], only[0]) and only[1] is None
    ctx.close()


# ------------------------------------------------------------------ edges
def test_out_of_domain_inputs_are_flagged_like_the_reference_fails():
    n_raise = n_ok = 0
    for rec in golden("out_of_domain.json")[:200]:
        kw = rec["kw"]
        a = start.validate_and_transform_args(**kw)
        _, _, rc = c_oracle.align(a[0], a[1], a[3], a[5])
        if rc == 1:
            with pytest.raises(batch.ReferenceUndefinedError):
                find_global_alignment(**kw)
            n_raise += 1
        else:
            r = find_global_alignment(**kw)
            assert [r.seq_1_aligned, r.middle_part, r.seq_2_aligned, r.cost, r.score] == rec["ok"]
            n_ok += 1
    assert n_raise > 20 and n_ok > 100


def test_argument_errors_surface_before_the_gpu():
    for rec in golden("error_cases.json"):
        if "exc" in rec:
            with pytest.raises(BaseException) as ei:
                find_global_alignment(**fix_paths(rec["kw"]))
            assert type(ei.value).__name__ == rec["exc"]


def test_native_argument_checks():
    ctx = _native.Context(0)
    off = np.array([0, 4], dtype=np.int64)
    with pytest.raises(RuntimeError):
        ctx.plan_batch(off, off)                      # no scheme yet
    ctx.set_scheme(_simple_dense(4, 5, 2, 2, 10))
    with pytest.raises(RuntimeError):
        ctx.plan_batch(np.array([0, 0], dtype=np.int64), off)      # empty sequence
    huge = _simple_dense(4, 1 << 20, 1 << 20, 1 << 20, 5)
    ctx.set_scheme(huge)
    with pytest.raises(OverflowError):
        ctx.plan_batch(np.array([0, 3000], dtype=np.int64), np.array([0, 3000], dtype=np.int64))
    ctx.close()


def test_single_letter_and_unicode_and_case():
    r = find_global_alignment(seq_1="a", seq_2="A")
    assert (r.seq_1_aligned, r.middle_part, r.seq_2_aligned, r.cost) == ("A", "|", "A", 0)
    r = find_global_alignment(seq_1="éβéβ", seq_2="éβ", gap_open_cost=1, mismatch_cost=2)
    ops, cost, rc = c_oracle.align("ÉΒÉΒ", "ÉΒ", r.costing_mat, 1)
    assert rc == 0 and r.cost == cost and r.seq_1_aligned == "ÉΒÉΒ"


# ------------------------------------------------------------------ full-size properties (config 2)
def test_config2_full_size_properties():
    n_pairs = 1_000_000
    b = synth.workload("cfg2", n_pairs=n_pairs)
    dense = batch.Scheme(**synth.SCHEMES["cfg2"]).dense_for(list("ACGT"))
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    packed = ctx.align_batch(b.s1, b.off1, b.s2, b.off2)
    assert ctx.plan_info()["packed16"] and ctx.plan_info()["cells"] == 150 * 150 * n_pairs
    generic = ctx.align_batch(b.s1, b.off1, b.s2, b.off2, flags=_native.GA_F_FORCE_GENERIC)
    assert not ctx.plan_info()["packed16"]
    for x, y in zip(packed, generic):          # two independent kernels, every pair, every op
        assert np.array_equal(x, y)
    cost, ops, ops_off, ops_len, status = packed
    assert not status.any() and ops_len.min() >= 150 and ops_len.max() <= 300
    # sample against the oracle + independent re-scoring of the emitted alignment
    rng = np.random.default_rng(1)
    sample = np.concatenate([np.arange(2000), rng.integers(0, n_pairs, 2000)])
    sub = synth.PairBatch(b.alphabet, b.s1, b.off1, b.s2, b.off2)
    o_ops, o_off, o_n, o_cost, o_st = c_oracle.align_batch(sub.s1[:2000 * 150], sub.off1[:2001], sub.s2[:2000 * 150],
                                                         sub.off2[:2001], _oracle_dense(dense))
    assert np.array_equal(cost[:2000], o_cost)
    dv = DenseView(dense)
    for k in sample:
        k = int(k)
        got = _native.unpack_ops(ops, int(ops_off[k]), int(ops_len[k]))
        if k < 2000:
            assert np.array_equal(got, o_ops[o_off[k]:o_off[k] + o_n[k]])
        a = b.s1[b.off1[k]:b.off1[k + 1]]
        c = b.s2[b.off2[k]:b.off2[k + 1]]
        assert rescore(got, a, c, dv) == cost[k]
    ctx.close()


@pytest.mark.parametrize("name,n_pairs", [("cfg3", 10_000), ("cfg5", 12_000)])
def test_config3_and_config5_shapes_10k_pairs_against_the_oracle(name, n_pairs):
    """BASELINE configs[2] (protein pairs 300-1000 aa, BLOSUM62, asymmetric 9/10 gap costs from the
    score -> cost transform, start.py:500-557) and configs[4] (250 bp DNA pairs): >= 10,000 pairs of the
    workload's own shape through the kernel the planner picks, every cost and every traceback op
    against the C oracle; the compact (RLE) output of the same call against the fixed-slice one."""
    b = synth.workload(name, n_pairs=n_pairs)
    scheme = batch.Scheme(**synth.SCHEMES[name])
    dense = scheme.dense_for(sorted(b.alphabet))
    remap = np.array([dense.index[ch] for ch in b.alphabet], dtype=np.uint8)
    s1, s2 = remap[b.s1], remap[b.s2]
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    out = ctx.align_batch(s1, b.off1, s2, b.off2)
    info = ctx.plan_info()
    assert info["packed16"] and info["fill_kernel"] == {"cfg3": "fill_p16b_r16_g32", "cfg5": "fill_p16_r16_g16"}[name]
    _compare_with_oracle(s1, b.off1, s2, b.off2, dense, out)
    res, rle, roff = ctx.align_batch_rle(s1, b.off1, s2, b.off2)
    assert np.array_equal(res["cost"], out[0]) and not (res["nops_status"] >> 31).any()
    for k in range(0, n_pairs, 7):
        assert np.array_equal(_native.unpack_rle(rle, int(roff[k]), int(roff[k + 1])),
                              _native.unpack_ops(out[1], int(out[2][k]), int(out[3][k]))), k
    ctx.close()


# ------------------------------------------------------------------ f1: seeded replay of random ties
def test_seeded_replay_reproduces_unpatched_reference(monkeypatch):
    """With TIE_BREAKING = 'reference_random', random.seed(s) + find_global_alignment gives what the
    UNPATCHED (randomly tie-breaking) reference gave for the same seed, and leaves the global
    generator in the same state (fingerprint: the next random.random())."""
    monkeypatch.setattr(globaligner, "TIE_BREAKING", "reference_random")
    cases = golden("seeded_cases.json")
    differs_from_canonical = 0
    for rec in cases:
        random.seed(rec["seed"])
        r = find_global_alignment(**fix_paths(rec["kw"]))
        assert [r.seq_1_aligned, r.middle_part, r.seq_2_aligned, r.cost, r.score] == rec["out"], rec["kw"]
        assert random.random() == rec["next_random"]
    monkeypatch.setattr(globaligner, "TIE_BREAKING", "canonical")
    for rec in cases[:60]:
        r = find_global_alignment(**fix_paths(rec["kw"]))
        assert r.cost == rec["out"][3]
        differs_from_canonical += [r.seq_1_aligned, r.seq_2_aligned] != [rec["out"][0], rec["out"][2]]
    assert differs_from_canonical > 3          # the seeded runs really exercise random ties


# ------------------------------------------------------------------ f3 / f4: FASTA batches, CLI, batch output
def test_cli_prints_the_reference_printout(capsys):
    rec = golden("kat.json")["tutorial"][0]            # globaligner --seq_1 ACGT --seq_2 AGT (tutorial.qmd:9-47)
    globaligner.main(["--seq_1", "ACGT", "--seq_2", "AGT"])
    assert capsys.readouterr().out == rec["str"] + "\n"
    rec = golden("kat.json")["tutorial"][1]
    globaligner.main(["--seq_1", "CCTGAA", "--seq_2", "GCCGA", "--match_score", "1", "--mismatch_score", "-1",
                      "--gap_open_score", "-2", "--gap_extension_score", "-1"])
    assert capsys.readouterr().out == rec["str"] + "\n"
    with pytest.raises(TypeError):                       # the reference's str-vs-int comparison (start.py:138)
        globaligner.main(["--seq_1", "ACGT", "--seq_2", "AGT", "--mismatch_cost", "5"])
    with pytest.raises(SystemExit):
        globaligner.main(["--version"])


def test_fasta_batch_and_batch_writer(tmp_path):
    recs = [r for r in golden("random_cases.json") if set(r["kw"]) == {"seq_1", "seq_2"}
            and r["kw"]["seq_1"].isascii() and r["kw"]["seq_2"].isascii()][:12]
    fa = tmp_path / "in.fa"
    fa.write_text("".join(f">a{k}\n{r['kw']['seq_1']}\n>b{k}\n{r['kw']['seq_2']}\n" for k, r in enumerate(recs)))
    descs, res = batch.find_global_alignments_from_fasta(fa)
    assert len(res) == len(recs)
    for k, rec in enumerate(recs):
        a = res[k]
        assert [a.seq_1_aligned, a.middle_part, a.seq_2_aligned, a.cost, a.score] == rec["out"]
    out = tmp_path / "all.txt"
    res.write(out, descs)
    text = out.read_text()
    one = find_global_alignment(input_fasta=str(fa))      # the reference-style call reads the first two records
    assert text.startswith(one.__str__(desc_1=">a0", desc_2=">b0"))
    assert text.count("\nscore: ") == len(recs)
    # the CLI's batch mode writes the same file; FASTQ input gives the same alignments
    out2 = tmp_path / "cli.txt"
    globaligner.main(["--batch", "-i", str(fa), "-o", str(out2)])
    assert out2.read_text() == text
    fq = tmp_path / "in.fastq"
    fq.write_text("".join(f"@a{k}\n{r['kw']['seq_1']}\n+\n{'I' * len(r['kw']['seq_1'])}\n"
                          f"@b{k}\n{r['kw']['seq_2']}\n+\n{'I' * len(r['kw']['seq_2'])}\n" for k, r in enumerate(recs)))
    out3 = tmp_path / "cli_fq.txt"
    globaligner.main(["--input_fastq", str(fq), "-o", str(out3)])
    assert out3.read_text() == text.replace(">a", "@a").replace(">b", "@b")


# ------------------------------------------------------------------ bit-packed sequence input
@pytest.mark.parametrize("bits,A", [(2, 4), (4, 4), (4, 13), (2, 3), (5, 24), (5, 32), (5, 4)])
def test_packed_sequence_streams_give_identical_results(bits, A):
    """GA_F_SEQ_2BIT / GA_F_SEQ_4BIT / GA_F_SEQ_5BIT: the same batch as little-endian 2- / 4- / 5-bit
    letter streams (what packed nucleotide files hold; 5 bits for amino acids, letters straddling
    bytes) must give byte-identical outputs."""
    rng = random.Random(bits * 100 + A)
    ctx = _native.Context(0)
    for dense in (_simple_dense(A, 4, 2, 1, 6), _table_dense(rng, A, 3)):
        ctx.set_scheme(dense)
        s1, off1, s2, off2 = _random_batch(rng, 333, A, 170, 190)
        plain = ctx.align_batch(s1, off1, s2, off2)
        packed = ctx.align_batch(_native.pack_letters(s1, bits), off1, _native.pack_letters(s2, bits), off2,
                                 flags=_native.seq_flag(bits))
        for x, y in zip(plain, packed):
            assert np.array_equal(x, y)
    with pytest.raises(RuntimeError):                 # 5 letters do not fit 2 bits
        ctx.set_scheme(_simple_dense(5, 4, 2, 1, 6))
        ctx.plan_batch(off1, off2, flags=_native.GA_F_SEQ_2BIT)
    ctx.close()


# ------------------------------------------------------------------ more edges
def test_long_seq2_in_batched_kernels_and_auto_routing():
    """seq_2 up to 32768 letters is staged in shared memory by the batched kernels; beyond that
    (or beyond 4M cells) find_global_alignment routes to the long-pair path."""
    rng = random.Random(9)
    ctx = _native.Context(0)
    for dense, A in ((_simple_dense(4, 5, 2, 2, 10), 4), (_table_dense(rng, 20, 7), 20)):
        ctx.set_scheme(dense)
        s1 = np.array([rng.randrange(A) for _ in range(300 + 90)], dtype=np.uint8)
        s2 = np.array([rng.randrange(A) for _ in range(20000 + 5000)], dtype=np.uint8)
        off1 = np.array([0, 300, 390], dtype=np.int64)
        off2 = np.array([0, 20000, 25000], dtype=np.int64)
        out = ctx.align_batch(s1, off1, s2, off2)
        _compare_with_oracle(s1, off1, s2, off2, dense, out)
    with pytest.raises(RuntimeError):
        ctx.plan_batch(np.array([0, 10], dtype=np.int64), np.array([0, 40000], dtype=np.int64))
    ctx.close()
    a = "ACGT" * 3
    b = "".join(rng.choice("ACGT") for _ in range(33000))
    with start.lifted_length_guard():
        r = find_global_alignment(seq_1=a, seq_2=b)
    ops, cost, rc = c_oracle.align(a, b, r.costing_mat, r.gap_open_cost)
    assert rc == 0 and r.cost == cost and (r.seq_1_aligned, r.middle_part, r.seq_2_aligned) == batch.render(ops, a, b)


def test_short_seq1_against_long_seq2_picks_a_kernel_that_fits():
    """A short seq_1 selects a few-lanes-per-group kernel by its rows; its shared-memory staging of
    seq_2 grows with n, so for long seq_2 the planner must move to wider groups / the int32 family
    instead of failing (header: n <= 32768).  12 x 20000 and 50 x 3500 with unit costs, through
    the batch call and through find_global_alignment (m * n below the long-pair threshold)."""
    rng = random.Random(21)
    ctx = _native.Context(0)
    for dense, A in ((_simple_dense(4, 1, 1, 1, 1), 4), (_simple_dense(4, 5, 2, 2, 10), 4), (_table_dense(rng, 20, 7), 20)):
        ctx.set_scheme(dense)
        for m, n in ((12, 20000), (50, 3500), (15, 16000), (3, 32768), (40, 13000)):
            s1 = np.array([rng.randrange(A) for _ in range(m + 7)], dtype=np.uint8)
            s2 = np.array([rng.randrange(A) for _ in range(n + 900)], dtype=np.uint8)
            off1 = np.array([0, m, m + 7], dtype=np.int64)
            off2 = np.array([0, n, n + 900], dtype=np.int64)
            out = ctx.align_batch(s1, off1, s2, off2)
            _compare_with_oracle(s1, off1, s2, off2, dense, out)
    ctx.close()
    a = "ACGTTGCAACGT"
    b = "".join(rng.choice("ACGT") for _ in range(20000))
    r = find_global_alignment(seq_1=a, seq_2=b, mismatch_cost=1, gap_open_cost=1, gap_extension_cost=1)
    ops, cost, rc = c_oracle.align(a, b, r.costing_mat, r.gap_open_cost)
    assert rc == 0 and r.cost == cost and (r.seq_1_aligned, r.middle_part, r.seq_2_aligned) == batch.render(ops, a, b)


def test_long_pair_call_leaves_a_packed_plan_intact():
    """ga_align_long takes one letter per byte; it must not change how a batch plan made for
    2-bit / 4-bit packed streams reads its input (the plan stays valid across the call)."""
    rng = random.Random(22)
    dense = _simple_dense(4, 5, 2, 2, 10)
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    import torch
    n = 300
    s1 = np.array([rng.randrange(4) for _ in range(n * 100)], dtype=np.uint8)
    s2 = np.array([rng.randrange(4) for _ in range(n * 100)], dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.int64) * 100
    ref = ctx.align_batch(s1, off, s2, off)
    for bits in (2, 4):
        ops_off = _native.ops_offsets(off, off)
        ctx.plan_batch(off, off, ops_off, flags=_native.seq_flag(bits))
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        d1, d2, do, doo = dev(_native.pack_letters(s1, bits)), dev(_native.pack_letters(s2, bits)), dev(off), dev(ops_off)
        dc = torch.empty(n, dtype=torch.int32, device="cuda")
        dl, dst = torch.empty_like(dc), torch.empty_like(dc)
        dops = torch.empty(int(ops_off[-1]) + 16, dtype=torch.uint8, device="cuda")
        run = lambda: ctx.run_batch(d1.data_ptr(), do.data_ptr(), d2.data_ptr(), do.data_ptr(), dc.data_ptr(),
                                    dops.data_ptr(), doo.data_ptr(), dl.data_ptr(), dst.data_ptr())
        run()
        ctx.sync()
        assert np.array_equal(dc.cpu().numpy(), ref[0])
        lp = synth.dna_single(700, 5, 6)
        ctx.align_long(lp.s1, lp.s2)                       # same context, between two runs of the plan
        dc.zero_()
        run()
        ctx.sync()
        assert np.array_equal(dc.cpu().numpy(), ref[0]) and np.array_equal(dl.cpu().numpy(), ref[3])
        got = dops.cpu().numpy()
        for k in range(n):
            assert np.array_equal(_native.unpack_ops(got, int(ops_off[k]), int(ref[3][k])),
                                  _native.unpack_ops(ref[1], int(ref[2][k]), int(ref[3][k]))), k
    ctx.close()


def test_out_of_range_letters_are_refused_or_harmless():
    """Letters >= A are a caller error.  Single-pair entry points refuse them; the batch kernels
    clamp them (results for that pair are meaningless, but nothing is read out of bounds and the
    context stays usable)."""
    rng = random.Random(23)
    dense = _table_dense(rng, 5, 3)
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    good = np.array([0, 1, 2, 3, 4, 1], dtype=np.uint8)
    bad = np.array([0, 1, 200, 3, 255, 1], dtype=np.uint8)
    table = np.zeros((7, 7, 3), dtype=np.int32)
    with pytest.raises(RuntimeError):
        ctx.fill_values(bad, good, table, keep_boundary=False)
    with pytest.raises(RuntimeError):
        ctx.align_replay(good, bad, np.zeros(12, dtype=np.uint16))
    with pytest.raises(RuntimeError):
        ctx.align_long(bad, good)
    off = np.array([0, 6, 12], dtype=np.int64)
    s1 = np.concatenate([good, bad])
    s2 = np.concatenate([bad, good])
    for cfg in (None, "fill_i32t_r8_g8", "fill_p16t_r8_g8", "fill_p16b_r16_g32"):
        if cfg:
            os.environ["GA_FORCE_CFG"] = cfg
        ctx.align_batch(s1, off, s2, off)                  # must not fault
        os.environ.pop("GA_FORCE_CFG", None)
    gs1 = np.concatenate([good, good[::-1]])
    out = ctx.align_batch(gs1, off, gs1[::-1].copy(), off)   # the context still works
    _compare_with_oracle(gs1, off, gs1[::-1].copy(), off, dense, out)
    ctx.close()


def test_timing_is_invalid_after_a_new_plan():
    dense = _simple_dense(4, 5, 2, 2, 10)
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    off = np.arange(65, dtype=np.int64) * 50
    s = np.zeros(64 * 50, dtype=np.uint8)
    ctx.align_batch(s, off, s, off)
    assert ctx.timing()["total"] > 0
    ctx.plan_batch(off[:9], off[:9], _native.ops_offsets(off[:9], off[:9]))
    with pytest.raises(RuntimeError):
        ctx.timing()                                       # the new plan has not run
    ctx.close()


def test_dp_array_backward_walks_the_table_it_is_given():
    """tests/golden/seam_cases.json (frozen from the reference by make_seam_golden.py): tables
    with hand-written or edited boundary values -- the first one is the reference's own
    (tests/globaligner_test.py:6-37).  dp_array_forward fills them on the GPU, dp_array_backward
    walks THOSE values (ga_walk_values), and both equal the reference."""
    from globalign_b200 import globaligner as GA
    cases = golden("seam_cases.json")
    assert cases[0]["seq_1"] == "AG" and cases[0]["row0"][0] == [0, 7, 7]
    differs_from_recomputed = 0
    for r in cases:
        s1, s2, cm, o = r["seq_1"], r["seq_2"], r["cm"], r["o"]
        m, n = len(s1), len(s2)
        dp = [[None] * (n + 1) for _ in range(m + 1)]
        for j in range(n + 1):
            dp[0][j] = tuple(r["row0"][j])
        for i in range(m + 1):
            dp[i][0] = tuple(r["col0"][i])
        GA.dp_array_forward(dp, s1, s2, cm, o)
        assert [[list(c) for c in row] for row in dp] == r["table"]
        if "error" in r:
            with pytest.raises(batch.ReferenceUndefinedError):
                GA.dp_array_backward(dp, s1, s2, cm, o)
            continue
        got = GA.dp_array_backward(dp, s1, s2, cm, o)
        assert list(got) == r["out"]
        if max(m, n) > o:
            try:
                differs_from_recomputed += list(GA.dp_array_backward(None, s1, s2, cm, o)) != r["out"]
            except batch.ReferenceUndefinedError:
                differs_from_recomputed += 1
    assert differs_from_recomputed > 0       # the fixture does exercise tables the sequences alone do not give


def test_batch_agrees_with_single_pair_call_on_one_letter_pairs():
    """Simple schemes build the costing matrix over the letters of the PAIR (start.py:319-343): a
    pair of one repeated letter has no mismatch entry, hence a smaller sentinel.  It only matters
    for max(m, n) <= gap_open_cost; there the batch call must still agree with the single-pair
    call ('AA' vs 'AAA', mismatch 3 / extension 1 / open 8 -> cost 6, not 9)."""
    kw = dict(mismatch_cost=3, gap_extension_cost=1, gap_open_cost=8)
    s1 = ["AA", "AAA", "ACGTACGTACGT", "CCCC", "A", "GGGGGGGGG", "TT", "AC"]
    s2 = ["AAA", "AA", "ACGAACGTACT", "CC", "AAAA", "GGGGGGGGGGG", "TTTTTTT", "AAA"]
    res = batch.find_global_alignments(s1, s2, **kw)
    for k, (a, b) in enumerate(zip(s1, s2)):
        try:
            one = find_global_alignment(seq_1=a, seq_2=b, **kw)
        except batch.ReferenceUndefinedError:
            assert res.status[k] != 0, k
            continue
        assert res.status[k] == 0, k
        r = res[k]
        assert (r.seq_1_aligned, r.middle_part, r.seq_2_aligned, r.cost, r.score) == \
            (one.seq_1_aligned, one.middle_part, one.seq_2_aligned, one.cost, one.score), k
    assert find_global_alignment(seq_1="AA", seq_2="AAA", **kw).cost == 6


def test_compact_rle_output_equals_fixed_slice_output():
    """ga_align_batch_rle: the same alignments as ga_align_batch, as one record per pair plus
    run-length encoded ops packed back to back.  Ragged and fixed-length batches, several chunk
    counts, simple and table schemes, out-of-domain pairs (status 1), and a buffer that is too small."""
    rng = random.Random(31)
    ctx = _native.Context(0)
    cases = []
    for A, dense in ((4, _simple_dense(4, 5, 2, 2, 10)), (4, _simple_dense(4, 3, 1, 1, 40)), (20, _table_dense(rng, 20, 6))):
        s1, o1, s2, o2 = _random_batch(rng, 3000, A, 180, 170)
        cases.append((dense, s1, o1, s2, o2, None))
        n = 2500
        u1 = np.array([rng.randrange(A) for _ in range(n * 100)], dtype=np.uint8)
        u2 = u1.copy()
        flip = np.array([rng.random() < 0.12 for _ in range(len(u2))])
        u2[flip] = [rng.randrange(A) for _ in range(int(flip.sum()))]
        off = np.arange(n + 1, dtype=np.int64) * 100
        cases.append((dense, u1, off, u2, off.copy(), (100, 100, n)))
    for dense, s1, o1, s2, o2, uni in cases:
        ctx.set_scheme(dense)
        for chunks in (None, "1", "3"):
            if chunks:
                os.environ["GA_CHUNKS"] = chunks
            else:
                os.environ.pop("GA_CHUNKS", None)
            cost, ops, ops_off, ops_len, status = ctx.align_batch(s1, o1, s2, o2)
            if uni:
                res, rle, roff = ctx.align_batch_rle(s1, None, s2, None, uni[0], uni[1], uni[2])
            else:
                res, rle, roff = ctx.align_batch_rle(s1, o1, s2, o2)
            assert np.array_equal(res["cost"], cost)
            assert np.array_equal(res["nops_status"] >> 31, (status != 0).astype(np.uint32))
            ok = status == 0
            assert np.array_equal((res["nops_status"] & 0x7FFFFFFF)[ok], ops_len[ok].astype(np.uint32))
            assert roff[0] == 0 and np.all(np.diff(roff) >= 0)
            for k in range(len(cost)):
                got = _native.unpack_rle(rle, int(roff[k]), int(roff[k + 1]))
                if ok[k]:
                    assert np.array_equal(got, _native.unpack_ops(ops, int(ops_off[k]), int(ops_len[k]))), k
                else:
                    assert len(got) == 0
            # runs are maximal: no two neighbouring bytes of a pair with the same op unless the first is a full run of 64
            k = int(np.argmax(np.diff(roff)))
            b = rle[int(roff[k]): int(roff[k + 1])]
            same = (b[1:] & 3) == (b[:-1] & 3)
            assert np.all((b[:-1][same] >> 2) == 63)
        os.environ.pop("GA_CHUNKS", None)
        total = int(roff[-1])
        small = (np.empty(len(cost), dtype=_native.PAIR_RESULT), np.empty(max(1, total // 2), dtype=np.uint8),
                 np.empty(len(cost) + 1, dtype=np.int64))
        with pytest.raises(MemoryError):
            if uni:
                ctx.align_batch_rle(s1, None, s2, None, uni[0], uni[1], uni[2], out=small)
            else:
                ctx.align_batch_rle(s1, o1, s2, o2, out=small)
        exact = (small[0], np.empty(total, dtype=np.uint8), small[2])
        r2 = ctx.align_batch_rle(s1, o1, s2, o2, out=exact) if not uni else \
            ctx.align_batch_rle(s1, None, s2, None, uni[0], uni[1], uni[2], out=exact)
        assert np.array_equal(r2[1][:total], rle[:total])
    ctx.close()


def test_one_letter_and_large_alphabets():
    r = find_global_alignment(seq_1="AAAAAAA", seq_2="AAAA")
    ops, cost, rc = c_oracle.align("AAAAAAA", "AAAA", r.costing_mat, r.gap_open_cost)
    assert r.cost == cost and r.seq_1_aligned == "AAAAAAA" and r.seq_2_aligned.count("-") == 3
    letters = "".join(chr(0x100 + k) for k in range(120))          # 120 distinct non-ASCII letters
    rng = random.Random(3)
    s1 = "".join(rng.choice(letters) for _ in range(200))
    s2 = "".join(rng.choice(letters) if rng.random() < 0.3 else c for c in s1[5:180])
    r = find_global_alignment(seq_1=s1, seq_2=s2, mismatch_cost=3, gap_open_cost=2, gap_extension_cost=1)
    ops, cost, rc = c_oracle.align(s1.upper(), s2.upper(), r.costing_mat, r.gap_open_cost)
    assert rc == 0 and r.cost == cost
    assert (r.seq_1_aligned, r.middle_part, r.seq_2_aligned) == batch.render(ops, s1.upper(), s2.upper())
    ctx = _native.Context(0)
    big = _table_dense(rng, 70, 2)                                   # table scheme with > 64 letters
    ctx.set_scheme(big)
    with pytest.raises(RuntimeError):
        ctx.plan_batch(np.array([0, 10], dtype=np.int64), np.array([0, 10], dtype=np.int64))
    ctx.close()


def test_pipeline_variants_on_mixed_length_batches():
    """Non-uniform batches cut into many chunks (uploaded tile descriptors, chunk-relative code
    offsets, ping-pong scratch), the serial plan, and cost-only must all agree with the oracle."""
    rng = random.Random(21)
    ctx = _native.Context(0)
    for dense, A in ((_simple_dense(4, 5, 2, 2, 10), 4), (_table_dense(rng, 20, 8), 20)):
        ctx.set_scheme(dense)
        data = _random_batch(rng, 1500, A, 260, 240)
        ctx.set_scratch_limit(64 << 30)
        whole = ctx.align_batch(*data)
        assert ctx.plan_info()["n_chunks"] == 1
        _compare_with_oracle(*data, dense, whole, sample=range(0, 1500, 7))
        ctx.set_scratch_limit(6 << 20)
        chunked = ctx.align_batch(*data)
        assert ctx.plan_info()["n_chunks"] >= 5
        serial = ctx.align_batch(*data, flags=_native.GA_F_SERIAL)
        only = ctx.align_batch(*data, flags=_native.GA_F_COST_ONLY)
        kept = ctx.align_batch(*data, flags=_native.GA_F_KEEP_ORDER)     # tiles in the caller's order
        for x, y, z, k in zip(whole, chunked, serial, kept):
            assert np.array_equal(x, y) and np.array_equal(x, z) and np.array_equal(x, k)
        assert np.array_equal(only[0], whole[0])
    ctx.close()


@pytest.mark.parametrize("m,n,bits", [(150, 150, 8), (150, 150, 2), (37, 91, 4), (300, 280, 2), (1, 5, 8), (700, 40, 8)])
def test_fixed_length_entry_point_equals_offset_array_call(m, n, bits):
    """ga_align_batch_uniform (implicit offsets) returns byte-identical costs, ops, lengths and
    status to ga_align_batch on the same fixed-length batch, also when the scratch limit forces
    several chunks."""
    rng = np.random.default_rng(m * 1000 + n)
    n_pairs = 3001
    s1 = rng.integers(0, 4, size=m * n_pairs, dtype=np.uint8)
    s2 = rng.integers(0, 4, size=n * n_pairs, dtype=np.uint8)
    # make most pairs similar so the walks are not trivial
    if m == n:
        keep = rng.random(m * n_pairs) < 0.9
        s2 = np.where(keep, s1, s2).astype(np.uint8)
    off1 = np.arange(n_pairs + 1, dtype=np.int64) * m
    off2 = np.arange(n_pairs + 1, dtype=np.int64) * n
    dense = batch.Scheme(mismatch_cost=3, gap_extension_cost=2, gap_open_cost=4).dense_for(list("ACGT"))
    ctx = _native.Context(0)
    ctx.set_scheme(dense)
    p1, p2 = _native.pack_letters(s1, bits), _native.pack_letters(s2, bits)
    fl = _native.seq_flag(bits)
    for limit in (64 << 30, 24 << 20):
        ctx.set_scratch_limit(limit)
        cost, ops, ops_off, ops_len, status = ctx.align_batch(p1, off1, p2, off2, flags=fl)
        ucost, uops, stride, ulen, ustatus = ctx.align_batch_uniform(p1, m, p2, n, n_pairs, flags=fl)
        assert np.array_equal(ops_off, np.arange(n_pairs + 1) * stride)
        assert np.array_equal(cost, ucost) and np.array_equal(ops_len, ulen) and np.array_equal(status, ustatus)
        for p in (0, 1, 1500, n_pairs - 1):
            assert np.array_equal(_native.unpack_ops(ops, int(ops_off[p]), int(ops_len[p])),
                                  _native.unpack_ops(uops, p * stride, int(ulen[p])))
        nb = (ops_len + 3) // 4
        a = ops[:len(uops)].reshape(n_pairs, stride)
        b = uops.reshape(n_pairs, stride)
        mask = np.arange(stride)[None, :] < nb[:, None]
        assert np.array_equal(a[mask], b[mask])
    # cost-only form
    c2 = ctx.align_batch_uniform(p1, m, p2, n, n_pairs, flags=fl | _native.GA_F_COST_ONLY)[0]
    assert np.array_equal(c2, cost)
    # against the oracle on a few pairs
    k = 64
    ref = c_oracle.align_batch(s1[:k * m], off1[:k + 1], s2[:k * n], off2[:k + 1], _oracle_dense(dense), 4)
    assert np.array_equal(ref[3], cost[:k])
    for p in range(k):
        if ref[4][p] == 0:
            assert np.array_equal(ref[0][ref[1][p]:ref[1][p] + ref[2][p]],
                                  _native.unpack_ops(uops, p * stride, int(ulen[p])))


def test_float_penalties_through_the_seam_functions():
    """SURVEY.md 8f, f2 / north_star: float-valued penalties.  The tolerance north_star allows is
    1e-6 relative on scores with identical alignments where there are no ties; the f64 kernels
    add in the reference's order, so this test demands MORE: bit-identical doubles (SHA-256 of
    the table bytes) and identical alignments including ties, against vectors generated by the
    reference's own make_dp_array / dp_array_forward / dp_array_backward
    (tests/golden/make_float_golden.py)."""
    import hashlib
    import struct
    from globalign_b200 import globaligner as GA

    cases = golden("float_cases.json")
    for r in cases:
        s1, s2, cm, o = r["seq_1"], r["seq_2"], r["cm"], r["o"]
        dp = GA.make_dp_array(s1, s2, cm, r["max_cost"], o)
        GA.dp_array_forward(dp, s1, s2, cm, o)
        h = hashlib.sha256()
        for row in dp:
            for cell in row:
                h.update(struct.pack("<3d", *(float(v) for v in cell)))
        assert h.hexdigest() == r["table_sha256"]
        a1, mid, a2, cost = GA.dp_array_backward(dp, s1, s2, cm, o)
        assert [a1, mid, a2, float(cost)] == r["out"]
        assert abs(cost - r["out"][3]) <= 1e-6 * max(1.0, abs(r["out"][3]))   # the stated tolerance


def test_float_walk_with_replayed_draws_and_argument_checks():
    from globalign_b200 import lowering
    cm = {x: {y: (0.0 if x == y else 1.5) for y in "AC-"} for x in "AC-"}
    cm["-"]["A"] = cm["-"]["C"] = cm["A"]["-"] = cm["C"]["-"] = 0.75
    dense = lowering.DenseSchemeF64(cm, 0.5)
    ctx = _native.default_context(0)
    s1, s2 = dense.encode("ACCA"), dense.encode("ACA")
    c0, ops0, ne0, st0 = ctx.align_f64(dense, s1, s2)
    assert st0 == 0 and c0 == 1.25
    # all-zero draws = first option everywhere = canonical
    c1, ops1, ne1, st1 = ctx.align_f64(dense, s1, s2, draws=np.zeros(len(s1) + len(s2), dtype=np.uint16))
    assert (c1, list(ops1), ne1, st1) == (c0, list(ops0), ne0, st0)
    # draws choosing option 1 at every tie = the oracle walking with that rule
    from oracle import py_oracle
    c2, ops2, _, st2 = ctx.align_f64(dense, s1, s2, draws=np.full(len(s1) + len(s2), 0x5555, dtype=np.uint16))
    tab = py_oracle.fill("ACCA", "ACA", cm, 0.5)
    want, wcost = py_oracle.walk(tab, "ACCA", "ACA", 0.5, choose=lambda opts: opts[min(1, len(opts) - 1)], cm=cm)
    assert st2 == 0 and c2 == wcost == c0 and list(ops2) == list(want) and list(ops2) != list(ops0)
    bad = lowering.DenseSchemeF64(cm, float("nan"))
    with pytest.raises(RuntimeError):
        ctx.align_f64(bad, s1, s2)


def test_two_gpu_sharding():
    """find_global_alignments(devices=[0, 1]): one host thread + context per GPU, equal-cell
    shards, merged in pair order -- identical to the single-GPU call."""
    if _native.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = random.Random(8)
    s1 = ["".join(rng.choice("ACGT") for _ in range(rng.randint(20, 300))) for _ in range(2000)]
    s2 = ["".join(ch if rng.random() < 0.9 else rng.choice("ACGT") for ch in x)[:rng.randint(15, 300)] or "A" for x in s1]
    one = batch.find_global_alignments(s1, s2, gap_open_cost=5)
    two = batch.find_global_alignments(s1, s2, gap_open_cost=5, devices=[0, 1])
    assert np.array_equal(one.cost, two.cost) and np.array_equal(one.score, two.score)
    assert np.array_equal(one.ops_len, two.ops_len) and np.array_equal(one.status, two.status)
    for k in range(0, 2000, 37):
        assert np.array_equal(one.ops_of(k), two.ops_of(k))
        a, b = one[k], two[k]
        assert (a.seq_1_aligned, a.middle_part, a.seq_2_aligned) == (b.seq_1_aligned, b.middle_part, b.seq_2_aligned)
    # fixed-length shards take the entry point without offset arrays on every GPU
    u1 = ["".join(rng.choice("ACGT") for _ in range(100)) for _ in range(1001)]
    u2 = [x[:50] + x[52:] + "AC" for x in u1]
    one = batch.find_global_alignments(u1, u2)
    two = batch.find_global_alignments(u1, u2, devices=[0, 1])
    assert np.array_equal(one.cost, two.cost)
    assert all(np.array_equal(one.ops_of(k), two.ops_of(k)) for k in range(0, 1001, 50))


def test_integration_md_ctypes_stub_runs_as_written():
    """The binding INTEGRATION.md shows a maintainer of the reference is executable: take the code
    block verbatim (only the library path is made absolute), run the reference's tutorial vectors
    and a few random golden cases through it."""
    from pathlib import Path
    import re

    root = Path(__file__).resolve().parents[1]
    text = (root / "INTEGRATION.md").read_text()
    block = re.search(r"## 2\. ctypes stub.*?```python\n(.*?)```", text, re.S).group(1)
    lib = root / "globalign_b200" / "libglobalign_b200.so"
    block = block.replace('"libglobalign_b200.so"', repr(str(lib)))
    ns = {"get_max_val": start.get_max_val}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    gpu_dp = ns["_gpu_dp"]
    cases = [r for r in golden("random_cases.json") if set(r["kw"]) == {"seq_1", "seq_2"}][:40]
    cases += [golden("kat.json")["tutorial"][0]]
    for rec in cases:
        kw = rec["kw"]
        s1, s2 = kw["seq_1"].upper(), kw["seq_2"].upper()
        ref = find_global_alignment(**kw)
        got = gpu_dp(s1, s2, ref.costing_mat, ref.gap_open_cost)
        assert list(got) == rec["out"][:4]


@pytest.mark.gpu
def test_guard_band_mode_detects_a_stray_write_and_passes_a_clean_run():
    """GA_DEBUG_GUARD=1 puts every scratch buffer of a context between two pattern-filled bands and
    checks them after each call (compute-sanitizer is not available on the pool).  A clean batch
    passes; with GA_DEBUG_GUARD=2 the library plants one stray byte behind a buffer and the same
    call must fail with the guard's message."""
    import subprocess
    import sys
    from pathlib import Path
    code = (
        "import numpy as np\n"
        "from globalign_b200 import batch\n"
        "r = batch.find_global_alignments(['ACGTACGTAC', 'TTTT'], ['ACGTTCGTAC', 'TTAT'])\n"
        "print('costs', r.cost.tolist())\n")
    root = str(Path(__file__).resolve().parents[1])
    outs = {}
    for mode in ("1", "2"):
        env = dict(os.environ, GA_DEBUG_GUARD=mode, PYTHONPATH=root)
        outs[mode] = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300, cwd=root)
    assert outs["1"].returncode == 0 and "costs" in outs["1"].stdout, outs["1"].stderr[-500:]
    assert outs["2"].returncode != 0 and "GA_DEBUG_GUARD" in outs["2"].stderr, outs["2"].stderr[-500:]
