"""Argument validation and scoring/costing scheme handling (host side).

Mirrors the public behaviour of the reference's ``globalign/start.py`` -- same
function names, argument meaning, defaults, return shapes and exception types --
so that ``find_global_alignment`` is a drop-in.  Citations are to
/root/reference/src/globalign/start.py.

  validate_and_transform_args        :150-353
  SimpleScoringSettings / SimpleCostingSettings   :10-147
  get_common_alphabet :355  check_seq_lengths :361  read_scoring_mat :378
  create_scoring_mat :431  create_costing_mat :451  validate_scoring_mat_keys :471
  get_max_val :488  scoring_mat_to_costing_mat :500  costing_mat_to_scoring_mat :559
  read_seq_from_fasta :614  read_first_2_seqs_from_fasta :666
  make_matrix :869  check_symmetric :883  check_big_main_diag :915

The scheme arithmetic is integer and tiny (alphabet <= a few dozen letters); it
stays in Python.  The DP itself never runs here: see ``globaligner.py``.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from pathlib import Path

from . import _matrices

GAP = "-"

#: The reference refuses pairs with len(seq_1) * len(seq_2) >= 20,000,000
#: (start.py:213).  The GPU path has no such limit; raise this module attribute
#: (or use ``lifted_length_guard``) to align longer pairs through the same entry
#: point.  The default keeps reference behaviour.
MAX_SEQ_LEN_PRODUCT = 20_000_000


class lifted_length_guard:
    """``with lifted_length_guard(): find_global_alignment(...)`` for long pairs."""

    def __init__(self, limit: int = 1 << 62):
        self.limit = limit

    def __enter__(self):
        global MAX_SEQ_LEN_PRODUCT
        self._saved = MAX_SEQ_LEN_PRODUCT
        MAX_SEQ_LEN_PRODUCT = self.limit
        return self

    def __exit__(self, *exc):
        global MAX_SEQ_LEN_PRODUCT
        MAX_SEQ_LEN_PRODUCT = self._saved
        return False


def _to_int(value, name):
    try:
        return int(value)
    except (TypeError, ValueError):
        print(f"{name} must be convertible to an integer.")
        raise


@dataclass
class SimpleScoringSettings:
    """match 2 / mismatch -3 / gap open -4 / gap extension -2 unless given (:16-19)."""
    match_score: int = 2
    mismatch_score: int = -3
    gap_open_score: int = -4
    gap_extension_score: int = -2

    def __post_init__(self):
        defaults = (("match_score", 2), ("mismatch_score", -3), ("gap_open_score", -4),
                    ("gap_extension_score", -2))
        vals = {}
        for name, dflt in defaults:
            given = getattr(self, name)
            vals[name] = _to_int(dflt if given is None else given, name)
        # sign rules are applied to the coerced integers (:74-84)
        if vals["match_score"] <= 0 or vals["mismatch_score"] >= 0:
            raise ValueError
        if vals["gap_open_score"] > 0 or vals["gap_extension_score"] >= 0:
            raise ValueError
        for name, v in vals.items():
            setattr(self, name, v)


@dataclass
class SimpleCostingSettings:
    """mismatch 5 / gap open 4 / gap extension 3 unless given (:95-97)."""
    mismatch_cost: int = 5
    gap_open_cost: int = 4
    gap_extension_cost: int = 3

    def __post_init__(self):
        defaults = (("mismatch_cost", 5), ("gap_open_cost", 4), ("gap_extension_cost", 3))
        raw = {}
        for name, dflt in defaults:
            given = getattr(self, name)
            raw[name] = dflt if given is None else given
            setattr(self, name, _to_int(raw[name], name))
        # The reference compares the value *as passed* (:138-145), so a numeric
        # string raises TypeError here exactly as it does there.
        if raw["mismatch_cost"] <= 0:
            raise ValueError
        if raw["gap_open_cost"] < 0:
            raise ValueError
        if raw["gap_extension_cost"] <= 0:
            raise ValueError


def get_common_alphabet(seq_1, seq_2):
    return sorted(set(seq_1) | set(seq_2))


def check_seq_lengths(seq_1, seq_2, max_seq_len_prod):
    m, n = len(seq_1), len(seq_2)
    if not m * n < max_seq_len_prod:
        raise RuntimeError(
            f"Your sequences are too long.  The product of their lengths should be less than "
            f"{max_seq_len_prod}.  They have lengths of {m} and {n}")
    if m * n == 0:
        raise RuntimeError("Detected a sequence of length 0.")


def get_max_val(m):
    """Largest entry of a dict-of-dict matrix."""
    best = -math.inf
    for row in m.values():
        best = max(best, max(row.values()))
    return best


def _deltas(max_score, delta_d, delta_i):
    if delta_d is None:
        delta_d = math.floor(max_score / 2)
    if delta_i is None:
        delta_i = math.ceil(max_score / 2)
    return delta_d, delta_i


def scoring_mat_to_costing_mat(scoring_mat, max_score, delta_d=None, delta_i=None):
    """Match-bonus scores -> non-negative costs (:500-557).

    With b = max_score, dd = floor(b/2), di = ceil(b/2):
    cost("-", y) = dd - s (horizontal step), cost(x, "-") = di - s (vertical step),
    everything else (including "-","-") = dd + di - s.
    """
    dd, di = _deltas(max_score, delta_d, delta_i)
    costing = {}
    for x, row in scoring_mat.items():
        costing[x] = {}
        for y, s in row.items():
            if x == GAP and y != GAP:
                costing[x][y] = dd - s
            elif y == GAP and x != GAP:
                costing[x][y] = di - s
            else:
                costing[x][y] = dd + di - s
    return costing


def costing_mat_to_scoring_mat(costing_mat, max_score, delta_d=None, delta_i=None):
    """Inverse transform (:559-612)."""
    dd, di = _deltas(max_score, delta_d, delta_i)
    scoring = {}
    for x, row in costing_mat.items():
        scoring[x] = {}
        for y, c in row.items():
            if x == GAP and y != GAP:
                scoring[x][y] = dd - c
            elif y == GAP and x != GAP:
                scoring[x][y] = di - c
            else:
                scoring[x][y] = dd + di - c
    return scoring


def _simple_mat(common_alphabet, diag, off_diag, gap):
    common_alphabet.append(GAP)          # the reference mutates the caller's list too
    return {x: {y: (diag if x == y else gap if GAP in (x, y) else off_diag)
                for y in common_alphabet} for x in common_alphabet}


def create_scoring_mat(common_alphabet, match_score, mismatch_score, gap_extension_score):
    return _simple_mat(common_alphabet, match_score, mismatch_score, gap_extension_score)


def create_costing_mat(common_alphabet, mismatch_cost, gap_extension_cost):
    return _simple_mat(common_alphabet, 0, mismatch_cost, gap_extension_cost)


def validate_scoring_mat_keys(scoring_mat_keys, common_alphabet):
    common_alphabet.append(GAP)
    missing = set(common_alphabet).difference(scoring_mat_keys)
    if missing:
        raise RuntimeError(
            f"common_alphabet contains values not in scoring_mat_keys, e.g. {missing}.  "
            f"Please check your sequences and your scoring matrix.")


def _parse_scoring_table(lines):
    """Whitespace table: header of single letters, then one row per letter."""
    it = iter(lines)
    letters = next(it, "").upper().split()
    if not all(len(x) == 1 for x in letters):
        raise RuntimeError("The header row did not have single letters spaced apart.")
    mat = dict.fromkeys(letters)
    for row_no, line in enumerate(it):
        cells = line.split()
        if cells[0] != letters[row_no]:
            raise RuntimeError("Row headers do not match column headers.")
        mat[cells[0]] = {y: int(cells[k + 1]) for k, y in enumerate(letters)}
    return mat


def read_scoring_mat(scoring_mat_path: Path):
    if not Path(scoring_mat_path).is_file():
        raise FileNotFoundError("scoring_mat_path does not point to a valid file.")
    with Path(scoring_mat_path).open() as fh:
        return _parse_scoring_table(fh)


def check_symmetric(mat):
    try:
        keys = list(mat.keys())
    except AttributeError:
        print("The check_symmetric function expected a nested dictionary.")
        raise
    for x in keys:
        for y in keys:
            try:
                if mat[x][y] != mat[y][x]:
                    return False
            except KeyError:
                return False
    return True


def check_big_main_diag(mat):
    ok = True            # empty matrix: the reference would fail; nothing sensible to mirror
    for x in mat.keys():
        try:
            ok = mat[x][x] == max(mat[x].values())
        except KeyError:
            raise RuntimeError("mat is not a proper nested dict representation of a matrix.")
        if not ok:
            return False
    return ok


def read_seq_from_fasta(fasta_path: Path):
    """Yield (description, SEQUENCE) records (:614-663)."""
    with Path(fasta_path).open() as fh:
        desc = fh.readline().strip()
        if not desc.startswith(">"):
            raise RuntimeError("Invalid FASTA format. Expected the first line to start with '>'.")
        chunks = []

        def finish():
            seq = "".join(chunks).upper()
            if not seq:
                raise RuntimeError("Empty sequence detected in FASTA.")
            return seq

        for line in fh:
            line = line.strip()
            if line.startswith(">"):
                yield desc, finish()
                desc = line
                chunks = []
            elif line:
                chunks.append(line)
        yield desc, finish()


def _open_text(path):
    """Plain or gzip-compressed text file (by the .gz suffix)."""
    path = Path(path)
    if path.suffix == ".gz":
        import gzip
        return gzip.open(path, "rt")
    return path.open()


def read_seq_from_fastq(fastq_path: Path):
    """Yield (description, SEQUENCE) from a four-line-per-record FASTQ file (plain or .gz).  Not in
    the reference (it reads FASTA only, :614-688); SURVEY.md 8f, f3 -- sequencing reads, the
    batched path's natural input, come as FASTQ.  Quality strings are checked for length and
    dropped (the aligner has no use for them); sequences are upper-cased like the FASTA reader's."""
    with _open_text(fastq_path) as fh:
        n = 0
        while True:
            head = fh.readline()
            if not head:
                break
            if not head.strip():
                continue
            seq, plus, qual = fh.readline().strip(), fh.readline().strip(), fh.readline().strip()
            n += 1
            if not head.startswith("@") or not plus.startswith("+"):
                raise RuntimeError(f"Invalid FASTQ format in record {n}: expected '@' and '+' lines.")
            if not seq:
                raise RuntimeError("Empty sequence detected in FASTQ.")
            if len(qual) != len(seq):
                raise RuntimeError(f"Invalid FASTQ format in record {n}: quality and sequence lengths differ.")
            yield head.strip(), seq.upper()
        if n == 0:
            raise RuntimeError("Invalid FASTQ format. No records found.")


def read_first_2_seqs_from_fasta(fasta_path: Path):
    seqs = []
    for _desc, seq in read_seq_from_fasta(fasta_path):
        seqs.append(seq)
        if len(seqs) == 2:
            break
    if len(seqs) < 2:
        raise RuntimeError("Two sequences could not be read from the FASTA file.")
    return seqs[0], seqs[1]


def make_matrix(num_rows, num_cols, fill_val):
    return [[fill_val] * num_cols for _ in range(num_rows)]


# --------------------------------------------------------------------------
# scheme resolution shared by the single-pair and the batched entry points
# --------------------------------------------------------------------------
_COST_OPTS = ("mismatch_cost", "gap_open_cost", "gap_extension_cost")


def resolve_scheme_options(scoring_mat_name, scoring_mat_path, match_score, mismatch_score,
                           mismatch_cost, gap_open_score, gap_open_cost, gap_extension_score,
                           gap_extension_cost):
    """Mutual-exclusion checks + settings objects (:227-262).

    Returns (branch, scoring_settings, costing_settings) with
    branch in {"name", "path", "costs", "scores"}.
    """
    others = (match_score, mismatch_score, mismatch_cost, gap_extension_score, gap_extension_cost)
    if scoring_mat_name is not None and any(
            x is not None for x in (scoring_mat_path,) + others):
        raise RuntimeError(
            "The scoring_mat_name should not be specified if any of the other options with scores "
            "or costs are specified, except for the gap_open options.")
    if scoring_mat_path is not None and any(
            x is not None for x in (scoring_mat_name,) + others):
        raise RuntimeError(
            "The scoring_mat_path should not be specified if any of the other options with scores "
            "or costs are specified, except for the gap_open options.")
    any_score = any(x is not None for x in
                    (match_score, mismatch_score, gap_open_score, gap_extension_score))
    any_cost = any(x is not None for x in (mismatch_cost, gap_open_cost, gap_extension_cost))
    if any_score and any_cost:
        raise RuntimeError("Scoring and costing options should not both be set.")

    scoring = SimpleScoringSettings(match_score, mismatch_score, gap_open_score,
                                    gap_extension_score)
    costing = SimpleCostingSettings(mismatch_cost, gap_open_cost, gap_extension_cost)
    # the two gap-open settings are always opposites of each other (:251-262)
    if gap_open_score is not None:
        costing.gap_open_cost = -scoring.gap_open_score
    else:
        scoring.gap_open_score = -costing.gap_open_cost

    if scoring_mat_name is not None:
        branch = "name"
    elif scoring_mat_path is not None:
        branch = "path"
    elif any_cost:
        branch = "costs"
    else:
        branch = "scores"
    return branch, scoring, costing


def load_named_scoring_mat(scoring_mat_name):
    """BLOSUM50 / BLOSUM62 / nucleotide, as the reference's packaged .mtx files
    parse (:265-273).  Unknown names fail like a missing file does there."""
    mat = _matrices.named(str(scoring_mat_name))
    if mat is None:
        raise FileNotFoundError("scoring_mat_path does not point to a valid file.")
    return mat


def load_matrix_scheme(branch, scoring_mat_name, scoring_mat_path):
    """Scoring + costing matrices of a matrix-based scheme (no sequence checks)."""
    if branch == "name":
        scoring_mat = load_named_scoring_mat(scoring_mat_name)
    else:
        scoring_mat = read_scoring_mat(Path(scoring_mat_path))
        if not check_symmetric(mat=scoring_mat):
            raise RuntimeError("The scoring matrix is not symmetric.")
        if not check_big_main_diag(mat=scoring_mat):
            raise RuntimeError(
                "The scoring matrix does not make sense because the maximum for each row does "
                "not occur on the main diagonal.")
    return scoring_mat


def simple_scheme_matrices(branch, common_alphabet, scoring, costing):
    """(scoring_mat, costing_mat) of a simple scheme over `common_alphabet`
    (which gets "-" appended, as in the reference) (:319-343)."""
    if branch == "costs":
        costing_mat = create_costing_mat(common_alphabet, costing.mismatch_cost,
                                         costing.gap_extension_cost)
        scoring_mat = costing_mat_to_scoring_mat(costing_mat, max_score=scoring.match_score)
    else:
        scoring_mat = create_scoring_mat(common_alphabet, scoring.match_score,
                                         scoring.mismatch_score, scoring.gap_extension_score)
        costing_mat = scoring_mat_to_costing_mat(scoring_mat, max_score=scoring.match_score)
    return scoring_mat, costing_mat


def validate_and_transform_args(
    input_fasta=None, output=None, seq_1=None, seq_2=None, scoring_mat_name=None,
    scoring_mat_path=None, match_score=None, mismatch_score=None, mismatch_cost=None,
    gap_open_score=None, gap_open_cost=None, gap_extension_score=None, gap_extension_cost=None,
):
    """-> (seq_1, seq_2, scoring_mat, costing_mat, gap_open_score, gap_open_cost, output)."""
    # ---- output (:184-194)
    if output is not None:
        output_validated = Path(output)
        if output_validated.is_file():
            raise RuntimeWarning(f"Overwriting {output_validated}")
        if not output_validated.parent.exists():
            raise FileNotFoundError("The parent directory of output does not exist.")
    else:
        output_validated = None

    # ---- sequences (:201-220)
    if input_fasta is not None and seq_1 is None and seq_2 is None:
        try:
            seq_1, seq_2 = read_first_2_seqs_from_fasta(Path(input_fasta))
        except FileNotFoundError:
            print("input_fasta does not point to a valid file.  Please make sure it is in the "
                  "correct FASTA format.  Note that reading from standard input is not supported "
                  "at this time.")
            raise
    elif ((input_fasta is None and seq_2 is None) or (input_fasta is not None and seq_1 is not None)
          or (seq_1 is None and seq_2 is not None)):
        raise RuntimeError(
            "The combination of arguments for input_fasta, seq_1, and seq_2 does not make sense.")
    check_seq_lengths(seq_1, seq_2, MAX_SEQ_LEN_PRODUCT)
    if GAP in seq_1 or GAP in seq_2:
        raise RuntimeError(
            "The current implementation does not allow for '-' characters in the sequences because "
            "they are used internally for gaps.  Please replace this character in your sequences.")
    seq_1 = seq_1.upper()
    seq_2 = seq_2.upper()

    # ---- scheme (:227-343)
    branch, scoring, costing = resolve_scheme_options(
        scoring_mat_name, scoring_mat_path, match_score, mismatch_score, mismatch_cost,
        gap_open_score, gap_open_cost, gap_extension_score, gap_extension_cost)
    common_alphabet = get_common_alphabet(seq_1, seq_2)
    if branch in ("name", "path"):
        scoring_mat = load_matrix_scheme(branch, scoring_mat_name, scoring_mat_path)
        validate_scoring_mat_keys(scoring_mat_keys=scoring_mat.keys(),
                                  common_alphabet=common_alphabet)
        costing_mat = scoring_mat_to_costing_mat(scoring_mat, max_score=get_max_val(scoring_mat))
    else:
        scoring_mat, costing_mat = simple_scheme_matrices(branch, common_alphabet, scoring, costing)

    return (seq_1, seq_2, scoring_mat, costing_mat, scoring.gap_open_score,
            costing.gap_open_cost, output_validated)
