"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/ga_oracle.c.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
import this.  It carries its own lowering of a dict-of-dict costing matrix to
dense tables, written independently of ``globalign_b200.lowering`` so that the
product's lowering is itself checked by the parity tests.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libga_oracle.so"
_lib = None

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> Path:
    """Compile the C oracle with gcc (seconds)."""
    src = _HERE / "ga_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s", "--always-make"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        scheme = [C.c_int, _i32p, _i32p, _i32p, C.c_int32, C.c_int32]
        L.gao_fill_table.argtypes = [C.c_int64, C.c_int64, _u8p, _u8p] + scheme + [C.c_int, _i32p]
        L.gao_align.argtypes = [C.c_int64, C.c_int64, _u8p, _u8p] + scheme + [
            _u8p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.gao_align_batch.argtypes = [C.c_int64, _u8p, _i64p, _u8p, _i64p] + scheme + [
            _u8p, _i64p, _i64p, _i32p, _i32p, C.c_int]
        L.gao_verify_path.argtypes = [C.c_int64, C.c_int64, _u8p, _u8p] + scheme + [
            _u8p, C.c_int64, C.POINTER(C.c_int32)]
        L.gao_cost_only.argtypes = [C.c_int64, C.c_int64, _u8p, _u8p] + scheme + [
            C.POINTER(C.c_int32)]
        for f in (L.gao_fill_table, L.gao_align, L.gao_align_batch, L.gao_verify_path,
                  L.gao_cost_only):
            f.restype = C.c_int
        _lib = L
    return _lib


class Dense:
    """Dense form of a dict-of-dict costing matrix (letters other than '-')."""

    def __init__(self, cm: dict, gap_open_cost: int):
        self.letters = [k for k in cm if k != "-"]
        self.index = {ch: k for k, ch in enumerate(self.letters)}
        A = len(self.letters)
        self.A = A
        self.sub = np.array([[cm[x][y] for y in self.letters] for x in self.letters],
                            dtype=np.int32).reshape(A * A)
        self.gap_h = np.array([cm["-"][y] for y in self.letters], dtype=np.int32)
        self.gap_v = np.array([cm[x]["-"] for x in self.letters], dtype=np.int32)
        self.o = int(gap_open_cost)
        self.max_cost = int(max(max(r.values()) for r in cm.values()))

    def encode(self, seq: str) -> np.ndarray:
        return np.fromiter((self.index[ch] for ch in seq), dtype=np.uint8, count=len(seq))

    def args(self):
        return (self.A, self.sub, self.gap_h, self.gap_v, self.o, self.max_cost)


def align_codes(s1: np.ndarray, s2: np.ndarray, d: Dense):
    """-> (ops uint8 walk order, cost, status)."""
    m, n = len(s1), len(s2)
    ops = np.zeros(m + n, dtype=np.uint8)
    nops = C.c_int64(0)
    cost = C.c_int32(0)
    rc = lib().gao_align(m, n, np.ascontiguousarray(s1), np.ascontiguousarray(s2), *d.args(),
                         ops, C.byref(nops), C.byref(cost))
    if rc < 0:
        raise MemoryError("oracle table allocation failed")
    return ops[: nops.value].copy(), cost.value, rc


def align(seq_1: str, seq_2: str, cm: dict, gap_open_cost: int):
    d = Dense(cm, gap_open_cost)
    return align_codes(d.encode(seq_1), d.encode(seq_2), d)


def fill_table(seq_1: str, seq_2: str, cm: dict, gap_open_cost: int, boundary=None):
    """Full (m+1, n+1, 3) int32 table.  `boundary`: optional nested list whose
    row 0 / column 0 are taken as they are (reference test globaligner_test.py:6-37)."""
    d = Dense(cm, gap_open_cost)
    m, n = len(seq_1), len(seq_2)
    t = np.zeros((m + 1, n + 1, 3), dtype=np.int32)
    keep = 0
    if boundary is not None:
        keep = 1
        for j in range(n + 1):
            t[0, j] = boundary[0][j]
        for i in range(m + 1):
            t[i, 0] = boundary[i][0]
    lib().gao_fill_table(m, n, d.encode(seq_1), d.encode(seq_2), *d.args(), keep, t.reshape(-1))
    return t


def align_batch(s1, off1, s2, off2, d: Dense, n_threads: int | None = None):
    """Batch over pthreads -> (ops, ops_off, nops, cost, status)."""
    n_pairs = len(off1) - 1
    lens = (off1[1:] - off1[:-1]) + (off2[1:] - off2[:-1])
    ops_off = np.zeros(n_pairs + 1, dtype=np.int64)
    np.cumsum(lens, out=ops_off[1:])
    ops = np.zeros(int(ops_off[-1]), dtype=np.uint8)
    nops = np.zeros(n_pairs, dtype=np.int64)
    cost = np.zeros(n_pairs, dtype=np.int32)
    status = np.zeros(n_pairs, dtype=np.int32)
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    rc = lib().gao_align_batch(n_pairs, s1, off1, s2, off2, *d.args(), ops, ops_off, nops, cost,
                               status, int(n_threads))
    if rc != 0:
        raise RuntimeError("gao_align_batch failed")
    return ops, ops_off, nops, cost, status


def verify_path(s1, s2, d: Dense, ops):
    """0 = `ops` is the canonical walk; 2 = differs; 1 = reference-undefined.  -> (rc, cost)"""
    cost = C.c_int32(0)
    ops = np.ascontiguousarray(ops, dtype=np.uint8)
    rc = lib().gao_verify_path(len(s1), len(s2), np.ascontiguousarray(s1), np.ascontiguousarray(s2),
                               *d.args(), ops, len(ops), C.byref(cost))
    return rc, cost.value


def cost_only(s1, s2, d: Dense):
    cost = C.c_int32(0)
    rc = lib().gao_cost_only(len(s1), len(s2), np.ascontiguousarray(s1), np.ascontiguousarray(s2),
                             *d.args(), C.byref(cost))
    if rc != 0:
        raise MemoryError
    return cost.value
