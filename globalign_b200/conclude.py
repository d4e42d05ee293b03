"""Result object and cost <-> score conversion (host side).

Same public surface as the reference's ``globalign/conclude.py``
(/root/reference/src/globalign/conclude.py): ``AlignmentResults`` (:7-151, ten
fields, ``__str__`` / ``print`` / ``write`` with 70-column blocks),
``final_cost_to_score`` (:154-177), ``final_score_to_cost`` (:179-202),
``prettify_mat`` (:252-310).
"""

from __future__ import annotations

import math
from pathlib import Path
from typing import NamedTuple

_BANNER = "###########################################\n# Settings\n###########################################\n"


class AlignmentResults(NamedTuple):
    seq_1_aligned: str
    middle_part: str
    seq_2_aligned: str
    cost: int
    score: int
    scoring_mat: dict
    costing_mat: dict
    gap_open_score: int
    gap_open_cost: int
    output: Path

    def __str__(self, desc_1: str = "seq_1", desc_2: str = "seq_2", chars_per_line: int = 70):
        total = len(self.middle_part)
        n_blocks = math.ceil(total / chars_per_line)
        parts = [desc_1, "\n", desc_2]
        for b in range(n_blocks):
            lo = b * chars_per_line
            hi = total if n_blocks == 1 else lo + chars_per_line
            parts += ["\n\n", self.seq_1_aligned[lo:hi], "\n", self.middle_part[lo:hi], "\n",
                      self.seq_2_aligned[lo:hi]]
        parts += [
            "\n\n",
            f"score: {self.score}\n",
            f"cost: {self.cost}\n",
            _BANNER,
            "scoring_mat:\n", prettify_mat(self.scoring_mat),
            f"\n\ngap_open_score: {self.gap_open_score}\n",
            "\ncosting_mat:\n", prettify_mat(self.costing_mat),
            f"\n\ngap_open_cost: {self.gap_open_cost}\n",
        ]
        return "".join(parts)

    def print(self, desc_1: str = "seq_1", desc_2: str = "seq_2", chars_per_line: int = 70) -> None:
        print(self.__str__(desc_1=desc_1, desc_2=desc_2, chars_per_line=chars_per_line))

    def write(self, file=None, desc_1: str = "seq_1", desc_2: str = "seq_2",
              chars_per_line: int = 70):
        """Write the printout to `file`, else to ``self.output``, else to stdout
        (``file="stdout"`` forces stdout)."""
        if (file is None and self.output is None) or file == "stdout":
            self.print(desc_1=desc_1, desc_2=desc_2, chars_per_line=chars_per_line)
            return None
        target = self.output if file is None else file
        text = self.__str__(desc_1=desc_1, desc_2=desc_2, chars_per_line=chars_per_line)
        with open(file=target, mode="w+") as fh:
            fh.write(text)
        return None


def _deltas(max_score, delta_d, delta_i):
    if delta_d is None:
        delta_d = math.floor(max_score / 2)
    if delta_i is None:
        delta_i = math.ceil(max_score / 2)
    return delta_d, delta_i


def final_cost_to_score(cost, m, n, max_score, delta_d=None, delta_i=None):
    """score = n*floor(b/2) + m*ceil(b/2) - cost, b = max_score (m = len(seq_1))."""
    dd, di = _deltas(max_score, delta_d, delta_i)
    return n * dd + m * di - cost


def final_score_to_cost(score, m, n, max_score, delta_d=None, delta_i=None):
    dd, di = _deltas(max_score, delta_d, delta_i)
    return -score + n * dd + m * di


def prettify_mat(mat: dict) -> str:
    """Right-aligned text table of a dict-of-dict matrix, column widths fitted
    to the widest cell/header, row headers left-aligned."""
    try:
        cols = list(next(iter(mat.values())).keys())
    except Exception:
        print("mat does not appear to represent a matrix as a nested dictionary.")
        raise
    widths = [max([len(str(c))] + [len(str(mat[r][c])) for r in mat]) for c in cols]
    head_w = max(len(str(c)) for c in cols) + 1
    lines = [" " * head_w + "".join(f"{str(c):>{w + 1}}" for c, w in zip(cols, widths))]
    for r in mat:
        lines.append(f"{str(r):<{head_w}}" + "".join(f"{str(mat[r][c]):>{w + 1}}"
                                                      for c, w in zip(cols, widths)))
    return "\n".join(lines)
