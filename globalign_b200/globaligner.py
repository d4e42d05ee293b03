"""Drop-in for the reference's ``globalign.globaligner`` on the GPU hot path.

``find_global_alignment`` keeps the reference's name, its 13 keyword arguments,
defaults, exceptions and return type (globaligner.py:132-314).  What changes is
the seam at :256-293: instead of make_dp_array -> dp_array_forward ->
dp_array_backward over a Python list of tuples, the pair goes through the C-ABI
(``ga_align_batch``) to the sm_100a kernels.  The three seam functions are also
exported with the reference's signatures for callers/tests that use them
directly; they run on the GPU as well (``ga_fill_values`` / ``ga_align_batch``).
There is no CPU implementation in this package.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

from . import _native
from .batch import ReferenceUndefinedError, render
from .conclude import AlignmentResults, final_cost_to_score
from .lowering import DenseScheme, DenseSchemeF64, is_integral
from .start import get_max_val, make_matrix, validate_and_transform_args

__all__ = ["find_global_alignment", "make_dp_array", "dp_array_forward", "dp_array_backward", "main"]


#: "canonical" (default): ties between co-optimal alignments go to the first option listed in
#: the reference's table (match/mismatch, gap in seq_1, gap in seq_2).  "reference_random":
#: reproduce the reference's random.choice tie-breaking draw for draw, consuming the global
#: ``random`` state exactly as it does -- after ``random.seed(s)`` the result equals the
#: unpatched reference's (single pairs; slower: full value table on the device).
TIE_BREAKING = "canonical"

_DRAW_SIZES = (3, 2, 2, 2, 3, 2, 2, 2, 3)      # random.choice sizes in dict-literal order (:599-634)


def _reference_draws(n_steps: int):
    """Emulate the 18 random.choice calls cost_ranks_dispatcher makes per step
    (globaligner.py:599-671) on the global ``random`` state.  CPython: choice(seq) =
    seq[_randbelow(len(seq))], _randbelow(n) = rejection sampling of getrandbits(n.bit_length())
    -- 2 bits for both n = 2 and n = 3 -- one Mersenne-Twister word per getrandbits call.
    -> (uint16 draws per step, cumulative MT words consumed after each step)."""
    import random
    grb = random.getrandbits
    draws = np.zeros(n_steps, dtype=np.uint16)
    words = np.zeros(n_steps + 1, dtype=np.int64)
    used = 0
    for s in range(n_steps):
        packed = 0
        for half in range(2):
            for k, size in enumerate(_DRAW_SIZES):
                r = grb(2)
                used += 1
                while r >= size:
                    r = grb(2)
                    used += 1
                if k < 4:
                    packed |= r << (8 * half + 2 * k)
        draws[s] = packed
        words[s + 1] = used
    return draws, words


def _align_one_replay(seq_1: str, seq_2: str, costing_mat: dict, gap_open_cost: int, device: int = 0):
    import random
    dense = DenseScheme(costing_mat, gap_open_cost)
    ctx = _native.default_context(device)
    ctx.set_scheme(dense)
    s1, s2 = dense.encode(seq_1), dense.encode(seq_2)
    state = random.getstate()
    draws, words = _reference_draws(len(s1) + len(s2))
    cost, ops, n_eval, status = ctx.align_replay(s1, s2, draws)
    # leave the global generator exactly where the reference would have left it
    random.setstate(state)
    if words[n_eval]:
        random.getrandbits(32 * int(words[n_eval]))
    if status != 0:
        raise ReferenceUndefinedError(
            "the reference's traceback leaves the DP table for this input "
            "(max(len(seq_1), len(seq_2)) <= gap_open_cost)")
    return ops, cost


#: a single pair with at least this many DP cells goes to the whole-chip band pipeline with
#: linear-space traceback (ga_align_long) instead of the one-warp batched kernel.  Measured
#: latency per call (tools/probe_single_pair.py): 300 x 300 0.32 vs 0.34 ms, 1000 x 1000 0.67 vs
#: 1.54 ms, 2000 x 2000 0.91 vs 5.2 ms -- the band pipeline wins from ~250 k cells on.
LONG_PAIR_CELLS = 1 << 18


def _align_one_f64(seq_1: str, seq_2: str, costing_mat: dict, gap_open_cost, device: int = 0):
    """Float-valued penalties: f64 value table + value-ranking walk on the GPU."""
    dense = DenseSchemeF64(costing_mat, gap_open_cost)
    ctx = _native.default_context(device)
    cost, ops, _n_eval, status = ctx.align_f64(dense, dense.encode(seq_1), dense.encode(seq_2))
    if status != 0:
        raise ReferenceUndefinedError(
            "the reference's traceback leaves the DP table for this input: it raises IndexError "
            "or wraps to negative indices there")
    return ops, cost


def _align_one(seq_1: str, seq_2: str, costing_mat: dict, gap_open_cost: int, device: int = 0):
    if not is_integral(costing_mat, gap_open_cost):
        return _align_one_f64(seq_1, seq_2, costing_mat, gap_open_cost, device)
    dense = DenseScheme(costing_mat, gap_open_cost)
    ctx = _native.default_context(device)
    ctx.set_scheme(dense)
    s1, s2 = dense.encode(seq_1), dense.encode(seq_2)
    m, n = len(s1), len(s2)
    if m * n >= LONG_PAIR_CELLS or n > 32768:
        cost1, ops1, status1, _info = ctx.align_long(s1, s2)
        cost, status = [cost1], [status1]
    else:
        cost, ops, ops_off, ops_len, status = ctx.align_batch(
            s1, np.array([0, m], dtype=np.int64), s2, np.array([0, n], dtype=np.int64))
        ops1 = None
    if status[0] != 0:
        raise ReferenceUndefinedError(
            "the reference's traceback leaves the DP table for this input "
            "(max(len(seq_1), len(seq_2)) <= gap_open_cost): it raises IndexError or wraps "
            "to negative indices there")
    if ops1 is not None:
        return ops1, int(cost[0])
    return _native.unpack_ops(ops, 0, int(ops_len[0])), int(cost[0])


def find_global_alignment(
    input_fasta: str | Path = None,
    output: str | Path = None,
    seq_1: str = None,
    seq_2: str = None,
    scoring_mat_name: str = None,
    scoring_mat_path: str | Path = None,
    match_score: str | int = None,
    mismatch_score: str | int = None,
    mismatch_cost: str | int = None,
    gap_open_score: str | int = None,
    gap_open_cost: str | int = None,
    gap_extension_score: str | int = None,
    gap_extension_cost: str | int = None,
) -> AlignmentResults:
    """Optimal global alignment of two sequences (affine gaps, three-level DP).

    Arguments, defaults and errors are those of the reference
    (globaligner.py:147-215); ties between co-optimal alignments are broken in
    the reference's canonical order (match/mismatch, then gap in seq_1, then gap
    in seq_2).
    """
    (seq_1, seq_2, scoring_mat, costing_mat, gap_open_score, gap_open_cost,
     output) = validate_and_transform_args(
        input_fasta, output, seq_1, seq_2, scoring_mat_name, scoring_mat_path, match_score,
        mismatch_score, mismatch_cost, gap_open_score, gap_open_cost, gap_extension_score,
        gap_extension_cost)
    if TIE_BREAKING == "reference_random":
        ops, cost = _align_one_replay(seq_1, seq_2, costing_mat, gap_open_cost)
    elif TIE_BREAKING == "canonical":
        ops, cost = _align_one(seq_1, seq_2, costing_mat, gap_open_cost)
    else:
        raise ValueError("TIE_BREAKING must be 'canonical' or 'reference_random'")
    seq_1_aligned, middle_part, seq_2_aligned = render(ops, seq_1, seq_2)
    score = final_cost_to_score(cost=cost, m=len(seq_1), n=len(seq_2),
                                max_score=get_max_val(scoring_mat))
    return AlignmentResults(seq_1_aligned, middle_part, seq_2_aligned, cost, score, scoring_mat,
                            costing_mat, gap_open_score, gap_open_cost, output)


# ---------------------------------------------------------------------------
# the seam functions, reference signatures (globaligner.py:756, :366, :395)
# ---------------------------------------------------------------------------
def make_dp_array(seq_1: str, seq_2: str, costing_mat: dict, max_cost: int, gap_open_cost: int):
    """(m+1) x (n+1) nested list, boundary row/column set, interior ``None``."""
    m, n = len(seq_1), len(seq_2)
    dp = make_matrix(num_rows=m + 1, num_cols=n + 1, fill_val=None)
    big = (max_cost + 1) * max(m, n)
    dp[0][0] = (0, 0, 0)
    run = gap_open_cost
    for j in range(1, n + 1):
        run += costing_mat["-"][seq_2[j - 1]]
        dp[0][j] = (big, run, big)
    if n == 0:
        return dp
    run = gap_open_cost
    for i in range(1, m + 1):
        run += costing_mat[seq_1[i - 1]]["-"]
        dp[i][0] = (big, big, run)
    return dp


def dp_array_forward(dp_array, seq_1: str, seq_2: str, costing_mat: dict, gap_open_cost):
    """Fill ``dp_array`` in place on the GPU, starting from the boundary row and
    column already present in it (as the reference's own test does)."""
    m, n = len(seq_1), len(seq_2)
    ctx = _native.default_context(0)
    boundary = [v for j in range(n + 1) for v in dp_array[0][j]] + [v for i in range(m + 1) for v in dp_array[i][0]]
    if not is_integral(costing_mat, gap_open_cost) or any(int(v) != v for v in boundary):
        dense = DenseSchemeF64(costing_mat, gap_open_cost)
        table = np.zeros((m + 1, n + 1, 3), dtype=np.float64)
        for j in range(n + 1):
            table[0, j] = dp_array[0][j]
        for i in range(m + 1):
            table[i, 0] = dp_array[i][0]
        ctx.fill_values_f64(dense, dense.encode(seq_1), dense.encode(seq_2), table, keep_boundary=True)
        for i in range(1, m + 1):
            row = dp_array[i]
            for j in range(1, n + 1):
                row[j] = tuple(float(v) for v in table[i, j])
        return None
    dense = DenseScheme(costing_mat, gap_open_cost)
    ctx.set_scheme(dense)
    table = np.zeros((m + 1, n + 1, 3), dtype=np.int32)
    for j in range(n + 1):
        table[0, j] = dp_array[0][j]
    for i in range(m + 1):
        table[i, 0] = dp_array[i][0]
    ctx.fill_values(dense.encode(seq_1), dense.encode(seq_2), table, keep_boundary=True)
    for i in range(1, m + 1):
        row = dp_array[i]
        for j in range(1, n + 1):
            row[j] = tuple(int(v) for v in table[i, j])
    return None


def dp_array_backward(dp_array, seq_1: str, seq_2: str, costing_mat: dict, gap_open_cost: int):
    """-> (seq_1_aligned, middle_part, seq_2_aligned, cost).  Like the reference's
    (globaligner.py:395-593) this walks the ``dp_array`` it is handed -- whatever boundary values
    it was built with (the reference's own test writes them by hand,
    tests/globaligner_test.py:6-37): the table is uploaded and one GPU thread ranks the three
    values of every cell on the path (``ga_walk_values``).  ``dp_array=None`` recomputes the
    table from the sequences (the whole hot path)."""
    if dp_array is None:
        ops, cost = _align_one(seq_1, seq_2, costing_mat, gap_open_cost)
        top, mid, bot = render(ops, seq_1, seq_2)
        return top, mid, bot, cost
    m, n = len(seq_1), len(seq_2)
    if m < 1 or n < 1:
        raise RuntimeError("dp_array_backward needs two non-empty sequences")
    try:
        values = np.array([[dp_array[i][j] for j in range(n + 1)] for i in range(m + 1)], dtype=np.float64)
    except (TypeError, ValueError) as exc:
        raise TypeError("dp_array must hold a 3-tuple of numbers in every cell "
                        "(fill it with dp_array_forward first)") from exc
    if values.shape != (m + 1, n + 1, 3):
        raise TypeError("dp_array must be (len(seq_1)+1) x (len(seq_2)+1) cells of 3 values")
    ctx = _native.default_context(0)
    integral = is_integral(costing_mat, gap_open_cost) and bool(np.all(values == np.rint(values))) \
        and float(np.abs(values).max()) < 2.0 ** 30
    if TIE_BREAKING == "reference_random":
        import random
        state = random.getstate()
        draws, words = _reference_draws(m + n)
    elif TIE_BREAKING == "canonical":
        draws = words = None
    else:
        raise ValueError("TIE_BREAKING must be 'canonical' or 'reference_random'")
    if integral:
        dense = DenseScheme(costing_mat, gap_open_cost)
        ctx.set_scheme(dense)
        cost, ops, n_eval, status = ctx.walk_values(dense.encode(seq_1), dense.encode(seq_2),
                                                    np.ascontiguousarray(values, dtype=np.int32), draws)
    else:
        dense = DenseSchemeF64(costing_mat, gap_open_cost)
        cost, ops, n_eval, status = ctx.walk_values_f64(dense, dense.encode(seq_1), dense.encode(seq_2),
                                                        np.ascontiguousarray(values), draws)
    if draws is not None:        # leave the global generator where the reference would have left it
        random.setstate(state)
        if words[n_eval]:
            random.getrandbits(32 * int(words[n_eval]))
    if status != 0:
        raise ReferenceUndefinedError(
            "the reference's traceback leaves the DP table on this dp_array: it raises IndexError "
            "or wraps to negative indices there")
    top, mid, bot = render(ops, seq_1, seq_2)
    if integral:
        # hand back the caller's own number type for the corner minimum (:425)
        cost = min(dp_array[m][n])
    return top, mid, bot, cost


def _version() -> str:
    """The reference asks importlib.metadata for 'globalign' eagerly and crashes when the
    package is not pip-installed (globaligner.py:35); here a missing install is not fatal."""
    try:
        from importlib.metadata import version
        return version("globalign_b200")
    except Exception:
        return "0+b200"


def main(argv=None):
    """Command line with the reference's flags (globaligner.py:23-129)."""
    import argparse

    p = argparse.ArgumentParser(
        description="Perform optimal global alignment of two nucleotide or amino acid sequences.")
    p.add_argument("--version", action="version", version=_version(), help="Prints the version and exits.")
    p.add_argument("-i", "--input_fasta", required=False)
    p.add_argument("-o", "--output", required=False)
    p.add_argument("--seq_1", required=False)
    p.add_argument("--seq_2", required=False)
    p.add_argument("--scoring_mat_name", required=False, choices=["BLOSUM50", "BLOSUM62"])
    p.add_argument("--scoring_mat_path", required=False)
    for name in ("match_score", "mismatch_score", "mismatch_cost", "gap_open_score", "gap_open_cost",
                 "gap_extension_score", "gap_extension_cost"):
        p.add_argument(f"--{name}", required=False)
    # batch mode (not in the reference, whose CLI aligns the first two FASTA records only):
    # every consecutive record pair of the FASTA / FASTQ input, one scheme, one GPU batch
    p.add_argument("--batch", action="store_true",
                   help="Align every consecutive pair of records of --input_fasta (or of --input_fastq) "
                        "and write all alignments to --output.")
    p.add_argument("--input_fastq", required=False)
    p.add_argument("--input_fastq_2", required=False, help="Second FASTQ file (R2): record k pairs with record k of --input_fastq.")
    args = p.parse_args(argv)
    kw = vars(args)
    batch_mode, fq1, fq2 = kw.pop("batch"), kw.pop("input_fastq"), kw.pop("input_fastq_2")
    if not (batch_mode or fq1):
        find_global_alignment(**kw).write()
        return None
    from . import batch
    scheme_kw = {k: v for k, v in kw.items() if k not in ("input_fasta", "output", "seq_1", "seq_2")}
    if fq1:
        descs, res = batch.find_global_alignments_from_fastq(fq1, fq2, **scheme_kw)
    elif kw["input_fasta"]:
        descs, res = batch.find_global_alignments_from_fasta(kw["input_fasta"], **scheme_kw)
    else:
        raise RuntimeError("--batch needs --input_fasta or --input_fastq")
    if kw["output"] is None:
        for k in range(len(res)):
            print(res[k].__str__(desc_1=descs[k][0], desc_2=descs[k][1]))
    else:
        res.write(kw["output"], descs=descs)
    return None


if __name__ == "__main__":
    sys.exit(main())
