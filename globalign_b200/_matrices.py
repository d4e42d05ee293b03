"""Substitution-score tables shipped with the package (BLOSUM50, BLOSUM62, nucleotide).

The values are the standard published matrices extended, as in the reference's
packaged tables (/root/reference/src/globalign/data/scoring_matrices/), by a gap
column/row "-" and a "-","-" corner of 1.  All three tables are symmetric, so
only the lower triangle (row-major, diagonal included) is stored;
tests/test_scheme.py checks them entry by entry against the golden dump of the
reference's own reader.
"""

_TABLES = {
    "BLOSUM50": ("ARNDCQEGHILKMFPSTWYVBJZX-", (
        "5",
        "-2 7",
        "-1 -1 7",
        "-2 -2 2 8",
        "-1 -4 -2 -4 13",
        "-1 1 0 0 -3 7",
        "-1 0 0 2 -3 2 6",
        "0 -3 0 -1 -3 -2 -3 8",
        "-2 0 1 -1 -3 1 0 -2 10",
        "-1 -4 -3 -4 -2 -3 -4 -4 -4 5",
        "-2 -3 -4 -4 -2 -2 -3 -4 -3 2 5",
        "-1 3 0 -1 -3 2 1 -2 0 -3 -3 6",
        "-1 -2 -2 -4 -2 0 -2 -3 -1 2 3 -2 7",
        "-3 -3 -4 -5 -2 -4 -3 -4 -1 0 1 -4 0 8",
        "-1 -3 -2 -1 -4 -1 -1 -2 -2 -3 -4 -1 -3 -4 10",
        "1 -1 1 0 -1 0 -1 0 -1 -3 -3 0 -2 -3 -1 5",
        "0 -1 0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1 2 5",
        "-3 -3 -4 -5 -5 -1 -3 -3 -3 -3 -2 -3 -1 1 -4 -4 -3 15",
        "-2 -1 -2 -3 -3 -1 -2 -3 2 -1 -1 -2 0 4 -3 -2 -2 2 8",
        "0 -3 -3 -4 -1 -3 -3 -4 -4 4 1 -3 1 -1 -3 -2 0 -3 -1 5",
        "-2 -1 5 6 -3 0 1 -1 0 -4 -4 0 -3 -4 -2 0 0 -5 -3 -3 6",
        "-2 -3 -4 -4 -2 -3 -3 -4 -3 4 4 -3 2 1 -3 -3 -1 -2 -1 2 -4 4",
        "-1 0 0 1 -3 4 5 -2 0 -3 -3 1 -1 -4 -1 0 -1 -2 -2 -3 1 -3 5",
        "-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1",
        "-5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 1",
    )),
    "BLOSUM62": ("ARNDCQEGHILKMFPSTWYVBJZX-", (
        "4",
        "-1 5",
        "-2 0 6",
        "-2 -2 1 6",
        "0 -3 -3 -3 9",
        "-1 1 0 0 -3 5",
        "-1 0 0 2 -4 2 5",
        "0 -2 0 -1 -3 -2 -2 6",
        "-2 0 1 -1 -3 0 0 -2 8",
        "-1 -3 -3 -3 -1 -3 -3 -4 -3 4",
        "-1 -2 -3 -4 -1 -2 -3 -4 -3 2 4",
        "-1 2 0 -1 -3 1 1 -2 -1 -3 -2 5",
        "-1 -1 -2 -3 -1 0 -2 -3 -2 1 2 -1 5",
        "-2 -3 -3 -3 -2 -3 -3 -3 -1 0 0 -3 0 6",
        "-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4 7",
        "1 -1 1 0 -1 0 0 0 -1 -2 -2 0 -1 -2 -1 4",
        "0 -1 0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1 1 5",
        "-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1 1 -4 -3 -2 11",
        "-2 -2 -2 -3 -2 -1 -2 -3 2 -1 -1 -2 -1 3 -3 -2 -2 2 7",
        "0 -3 -3 -3 -1 -2 -2 -3 -3 3 1 -2 1 -1 -2 -2 0 -3 -1 4",
        "-2 -1 4 4 -3 0 1 -1 0 -3 -4 0 -3 -3 -2 0 -1 -4 -3 -3 4",
        "-1 -2 -3 -3 -1 -2 -3 -4 -3 3 3 -3 2 0 -3 -2 -1 -2 -1 2 -3 3",
        "-1 0 0 1 -3 4 4 -2 0 -3 -3 1 -1 -3 -1 0 -1 -2 -2 -2 0 -3 4",
        "-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1",
        "-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 1",
    )),
    "nucleotide": ("ACGT-", (
        "1",
        "-1 1",
        "-1 -1 1",
        "-1 -1 -1 1",
        "-2 -2 -2 -2 1",
    )),
}


def named(name):
    """dict-of-dict scoring matrix (letters in table order, "-" last) or None."""
    entry = _TABLES.get(name)
    if entry is None:
        return None
    letters, rows = entry
    tri = [[int(v) for v in row.split()] for row in rows]
    return {x: {y: (tri[i][j] if j <= i else tri[j][i]) for j, y in enumerate(letters)}
            for i, x in enumerate(letters)}
