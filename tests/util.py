"""Shared helpers of the test-suite."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def golden(name: str):
    return json.loads((GOLDEN / name).read_text())


def fix_paths(kw: dict) -> dict:
    """Golden kwargs store file names relative to tests/golden."""
    kw = dict(kw)
    for k in ("scoring_mat_path",):
        if kw.get(k) is not None and not Path(kw[k]).is_absolute():
            kw[k] = str(GOLDEN / kw[k])
    if kw.get("output") == "custom_dna.mtx":
        kw["output"] = str(GOLDEN / "custom_dna.mtx")
    return kw


def rescore(ops, s1, s2, dense) -> int:
    """Cost of the alignment described by walk-order ops (independent re-scoring):
    substitution / gap-extension cost of every column plus gap_open per gap run."""
    A = dense.A
    i, j = len(s1), len(s2)
    total = 0
    prev = 0
    for k in ops:
        if k == 0:
            i -= 1
            j -= 1
            total += int(dense.sub[int(s1[i]) * A + int(s2[j])])
        elif k == 1:
            j -= 1
            total += int(dense.gap_h[int(s2[j])]) + (dense.o_ if prev != 1 else 0)
        else:
            i -= 1
            total += int(dense.gap_v[int(s1[i])]) + (dense.o_ if prev != 2 else 0)
        prev = k
    assert i == 0 and j == 0
    return total


class DenseView:
    """Uniform view over oracle.c_oracle.Dense / globalign_b200.lowering.DenseScheme."""

    def __init__(self, d):
        self.A, self.sub, self.gap_h, self.gap_v = d.A, d.sub, d.gap_h, d.gap_v
        self.o_ = getattr(d, "o", None)
        if self.o_ is None:
            self.o_ = d.gap_open
