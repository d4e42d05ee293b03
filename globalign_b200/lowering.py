"""Lowering of a scheme (dict-of-dict costing matrix) to the dense tables of the C-ABI.

``costing_mat`` is exactly what the reference passes to make_dp_array /
dp_array_forward / dp_array_backward (globaligner.py:260-287): a nested dict over
the alphabet plus ``"-"``.  ``ga_set_scheme`` wants
    sub[a*A+b] = costing_mat[a][b],  gap_h[b] = costing_mat["-"][b],
    gap_v[a]   = costing_mat[a]["-"], gap_open, max_cost = get_max_val(costing_mat)
with letters replaced by indices 0..A-1 in the dict's own key order.
"""

from __future__ import annotations

import numpy as np

GAP = "-"


def is_integral(costing_mat: dict, gap_open_cost) -> bool:
    """True when every cost is a whole number -- what the reference's entry point always produces
    (start.py:51-69,121-133,423); False selects the f64 path."""
    vals = [v for row in costing_mat.values() for v in row.values()] + [gap_open_cost]
    return all(int(v) == v for v in vals)


class DenseScheme:
    def __init__(self, costing_mat: dict, gap_open_cost: int, letters=None):
        self.letters = [k for k in costing_mat if k != GAP] if letters is None else list(letters)
        A = len(self.letters)
        if A < 1 or A > 250:
            raise RuntimeError(f"the GPU path supports alphabets of 1..250 letters, got {A}")
        self.A = A
        self.index = {ch: k for k, ch in enumerate(self.letters)}
        vals = [[costing_mat[x][y] for y in self.letters] for x in self.letters]
        gh = [costing_mat[GAP][y] for y in self.letters]
        gv = [costing_mat[x][GAP] for x in self.letters]
        for v in [c for row in vals for c in row] + gh + gv:
            if int(v) != v:
                raise TypeError("the GPU path is integer-only (the reference coerces every "
                                "score/cost with int(); start.py:51-69,121-133,423)")
        self.sub = np.array(vals, dtype=np.int32).reshape(A * A)
        self.gap_h = np.array(gh, dtype=np.int32)
        self.gap_v = np.array(gv, dtype=np.int32)
        self.gap_open = int(gap_open_cost)
        self.max_cost = int(max(max(row.values()) for row in costing_mat.values()))
        self._lut = None

    def key(self):
        return (self.A, self.sub.tobytes(), self.gap_h.tobytes(), self.gap_v.tobytes(),
                self.gap_open, self.max_cost)

    def encode(self, seq: str) -> np.ndarray:
        """str -> uint8 letter indices (KeyError for letters outside the scheme)."""
        idx = self.index
        try:
            if self._lut is None and all(ord(ch) < 256 for ch in self.letters):
                lut = np.full(256, 255, dtype=np.uint8)
                for ch, k in idx.items():
                    lut[ord(ch)] = k
                self._lut = lut
            if self._lut is not None:
                raw = np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)
                out = self._lut[raw]
                if out.size and out.max() == 255:
                    raise KeyError(next(ch for ch in seq if ch not in idx))
                return out
        except UnicodeEncodeError:
            pass
        return np.fromiter((idx[ch] for ch in seq), dtype=np.uint8, count=len(seq))


class DenseSchemeF64(DenseScheme):
    """The same lowering in float64, for costing matrices with fractional entries (reachable in
    the reference only through its inner functions; globaligner.py:371,761)."""

    def __init__(self, costing_mat: dict, gap_open_cost, letters=None):
        self.letters = [k for k in costing_mat if k != GAP] if letters is None else list(letters)
        A = len(self.letters)
        if A < 1 or A > 250:
            raise RuntimeError(f"the GPU path supports alphabets of 1..250 letters, got {A}")
        self.A = A
        self.index = {ch: k for k, ch in enumerate(self.letters)}
        self.sub = np.array([[costing_mat[x][y] for y in self.letters] for x in self.letters],
                            dtype=np.float64).reshape(A * A)
        self.gap_h = np.array([costing_mat[GAP][y] for y in self.letters], dtype=np.float64)
        self.gap_v = np.array([costing_mat[x][GAP] for x in self.letters], dtype=np.float64)
        self.gap_open = float(gap_open_cost)
        self.max_cost = float(max(max(row.values()) for row in costing_mat.values()))
        self._lut = None
