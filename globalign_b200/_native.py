"""ctypes binding of libglobalign_b200.so (the C-ABI in include/globalign_b200.h).

There is no CPU fallback: importing this module is cheap, but the first call
that needs the library raises ``NativeLibraryError`` if the shared object has
not been built (``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C globalign_b200/csrc``), and any entry point that computes raises if no
CUDA device is present.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import os

#: GA_B200_LIB: another build of the library (tuning variants of tools/tune_long.py); default: in-tree
LIB_PATH = Path(os.environ.get("GA_B200_LIB") or Path(__file__).resolve().parent / "libglobalign_b200.so")

GA_F_DEFAULT = 0
GA_F_FORCE_GENERIC = 1
GA_F_COST_ONLY = 2
GA_F_SERIAL = 4
GA_F_SEQ_4BIT = 8
GA_F_SEQ_2BIT = 16
GA_F_SEQ_5BIT = 64
GA_F_KEEP_ORDER = 32
T_H2D, T_FILL, T_WALK, T_D2H, T_TOTAL, T_COUNT = 0, 1, 2, 3, 4, 5

#: every symbol include/globalign_b200.h declares (checked by the CPU tests)
EXPORTS = (
    "ga_version", "ga_device_count", "ga_ctx_create", "ga_ctx_destroy", "ga_last_error",
    "ga_ctx_set_stream", "ga_ctx_set_scratch_limit", "ga_set_scheme", "ga_align_batch",
    "ga_plan_batch", "ga_run_batch", "ga_sync", "ga_get_plan_info", "ga_get_timing",
    "ga_fill_values", "ga_alu_peak", "ga_align_long", "ga_align_replay", "ga_align_batch_uniform",
    "ga_fill_values_f64", "ga_align_f64", "ga_walk_values", "ga_walk_values_f64", "ga_align_batch_rle",
)


class NativeLibraryError(RuntimeError):
    pass


class PlanInfo(C.Structure):
    _fields_ = [("fill_kernel", C.c_char * 64), ("rows_per_lane", C.c_int32),
                ("lanes_per_group", C.c_int32), ("packed16", C.c_int32),
                ("launches_per_run", C.c_int32), ("code_bytes", C.c_int64),
                ("cells", C.c_int64), ("n_chunks", C.c_int64)]


class LongInfo(C.Structure):
    _fields_ = [("fill_ms", C.c_float), ("trace_ms", C.c_float), ("launches", C.c_int32),
                ("tile_rows", C.c_int32), ("tile_cols", C.c_int32), ("retraced_tiles", C.c_int32),
                ("checkpoint_bytes", C.c_int64), ("cells", C.c_int64), ("trace_mode", C.c_int32),
                ("reserved", C.c_int32)]


_lib = None


def lib():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            f"(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, i64, i32, u32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32
    L.ga_version.restype = C.c_int
    L.ga_device_count.argtypes = [C.POINTER(C.c_int)]
    L.ga_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.ga_ctx_destroy.argtypes = [vp]
    L.ga_last_error.argtypes = [vp]
    L.ga_last_error.restype = C.c_char_p
    L.ga_ctx_set_stream.argtypes = [vp, vp]
    L.ga_ctx_set_scratch_limit.argtypes = [vp, i64]
    L.ga_set_scheme.argtypes = [vp, C.c_int, vp, vp, vp, i32, i32]
    L.ga_align_batch.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, u32]
    L.ga_align_batch_uniform.argtypes = [vp, vp, i64, vp, i64, i64, vp, vp, vp, vp, u32]
    L.ga_align_batch_rle.argtypes = [vp, vp, vp, i64, vp, vp, i64, i64, vp, vp, i64, vp, u32]
    L.ga_plan_batch.argtypes = [vp, vp, vp, i64, vp, u32]
    L.ga_run_batch.argtypes = [vp] + [vp] * 9
    L.ga_sync.argtypes = [vp]
    L.ga_get_plan_info.argtypes = [vp, C.POINTER(PlanInfo)]
    L.ga_get_timing.argtypes = [vp, C.POINTER(C.c_float)]
    L.ga_fill_values.argtypes = [vp, vp, i64, vp, i64, C.c_int, vp]
    L.ga_alu_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.ga_align_long.argtypes = [vp, vp, i64, vp, i64, C.POINTER(i32), vp, C.POINTER(i64),
                                C.POINTER(i32), u32, C.POINTER(LongInfo)]
    f64 = C.c_double
    L.ga_fill_values_f64.argtypes = [vp, i32, vp, vp, vp, f64, f64, vp, i64, vp, i64, C.c_int, vp]
    L.ga_align_f64.argtypes = [vp, i32, vp, vp, vp, f64, f64, vp, i64, vp, i64, vp, i64, C.POINTER(f64), vp,
                               C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    L.ga_align_replay.argtypes = [vp, vp, i64, vp, i64, vp, i64, C.POINTER(i32), vp, C.POINTER(i64),
                                  C.POINTER(i32), C.POINTER(i32)]
    L.ga_walk_values.argtypes = [vp, vp, i64, vp, i64, vp, vp, i64, C.POINTER(i32), vp, C.POINTER(i64),
                                 C.POINTER(i32), C.POINTER(i32)]
    L.ga_walk_values_f64.argtypes = [vp, i32, vp, vp, vp, f64, f64, vp, i64, vp, i64, vp, vp, i64,
                                     C.POINTER(f64), vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    for name in EXPORTS:
        if name not in ("ga_last_error",):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    """CUDA devices the library sees (0 when there is none)."""
    n = C.c_int(0)
    rc = lib().ga_device_count(C.byref(n))
    return n.value if rc == 0 else 0


class Context:
    """One ga_ctx (device + stream + scratch).  Not thread-safe by design."""

    def __init__(self, device: int = 0):
        self._L = lib()
        n = C.c_int(0)
        rc = self._L.ga_device_count(C.byref(n))
        if rc != 0 or n.value <= 0:
            raise NativeLibraryError("no CUDA device available: globalign_b200 has no CPU fallback")
        h = C.c_void_p()
        rc = self._L.ga_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise NativeLibraryError(f"ga_ctx_create(device={device}) failed with {rc}")
        self._h = h
        self.device = int(device)
        self._scheme_key = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.ga_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._L.ga_last_error(self._h).decode("utf-8", "replace")
            exc = {-3: OverflowError, -4: MemoryError}.get(rc, RuntimeError)
            raise exc(f"globalign_b200 native error {rc}: {msg}")

    # -- configuration
    def set_stream(self, cuda_stream: int | None):
        self._check(self._L.ga_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_scratch_limit(self, nbytes: int):
        self._check(self._L.ga_ctx_set_scratch_limit(self._h, int(nbytes)))

    def set_scheme(self, dense):
        """`dense`: a lowering.DenseScheme."""
        key = dense.key()
        if key == self._scheme_key:
            return
        self._check(self._L.ga_set_scheme(self._h, dense.A, _ptr(dense.sub), _ptr(dense.gap_h),
                                          _ptr(dense.gap_v), dense.gap_open, dense.max_cost))
        self._scheme_key = key

    # -- host-buffer entry point
    def align_batch(self, s1, off1, s2, off2, flags: int = 0, out=None):
        """-> (cost int32[n], ops uint8[...], ops_off int64[n+1], ops_len int32[n], status int32[n]).

        `out`: optional preallocated (cost, ops, ops_off, ops_len, status) arrays
        (e.g. views of pinned memory) to write into instead of allocating."""
        n = len(off1) - 1
        s1 = np.ascontiguousarray(s1, dtype=np.uint8)
        s2 = np.ascontiguousarray(s2, dtype=np.uint8)
        off1 = np.ascontiguousarray(off1, dtype=np.int64)
        off2 = np.ascontiguousarray(off2, dtype=np.int64)
        if flags & GA_F_COST_ONLY:
            cost = np.empty(n, dtype=np.int32) if out is None else out[0]
            self._check(self._L.ga_align_batch(self._h, _ptr(s1), _ptr(off1), _ptr(s2), _ptr(off2), n,
                                               _ptr(cost), None, None, None, None, flags))
            return cost, None, None, None, None
        if out is None:
            ops_off = ops_offsets(off1, off2)
            cost = np.empty(n, dtype=np.int32)
            ops = np.empty(int(ops_off[-1]), dtype=np.uint8)
            ops_len = np.empty(n, dtype=np.int32)
            status = np.empty(n, dtype=np.int32)
        else:
            cost, ops, ops_off, ops_len, status = out
        self._check(self._L.ga_align_batch(self._h, _ptr(s1), _ptr(off1), _ptr(s2), _ptr(off2), n,
                                           _ptr(cost), _ptr(ops), _ptr(ops_off), _ptr(ops_len),
                                           _ptr(status), flags))
        return cost, ops, ops_off, ops_len, status

    def align_batch_uniform(self, s1, m: int, s2, n: int, n_pairs: int, flags: int = 0, out=None):
        """Fixed-length batch (every pair m x n), no offset arrays.
        -> (cost, ops, ops_stride, ops_len, status); pair p's ops start at byte p * ops_stride."""
        s1 = np.ascontiguousarray(s1, dtype=np.uint8)
        s2 = np.ascontiguousarray(s2, dtype=np.uint8)
        stride = (((m + n + 3) // 4) + 3) // 4 * 4
        if flags & GA_F_COST_ONLY:
            cost = np.empty(n_pairs, dtype=np.int32) if out is None else out[0]
            self._check(self._L.ga_align_batch_uniform(self._h, _ptr(s1), m, _ptr(s2), n, n_pairs, _ptr(cost),
                                                       None, None, None, flags))
            return cost, None, stride, None, None
        if out is None:
            out = (np.empty(n_pairs, dtype=np.int32), np.empty(stride * n_pairs, dtype=np.uint8),
                   np.empty(n_pairs, dtype=np.int32), np.empty(n_pairs, dtype=np.int32))
        cost, ops, ops_len, status = out
        self._check(self._L.ga_align_batch_uniform(self._h, _ptr(s1), m, _ptr(s2), n, n_pairs, _ptr(cost),
                                                   _ptr(ops), _ptr(ops_len), _ptr(status), flags))
        return cost, ops, stride, ops_len, status

    def align_batch_rle(self, s1, off1, s2, off2, m: int = 0, n: int = 0, n_pairs: int | None = None,
                        flags: int = 0, out=None, capacity: int | None = None):
        """Compact result stream (ga_align_batch_rle): -> (res, rle, rle_off) with
        res = structured array (cost int32, nops_status uint32: bit 31 = status), rle = run-length
        encoded ops of all pairs back to back, rle_off = int64[n_pairs + 1].
        off1 = off2 = None: fixed-length batch, every pair m x n (n_pairs given).
        `out`: optional preallocated (res, rle, rle_off), e.g. views of pinned memory."""
        s1 = np.ascontiguousarray(s1, dtype=np.uint8)
        s2 = np.ascontiguousarray(s2, dtype=np.uint8)
        if off1 is not None:
            off1 = np.ascontiguousarray(off1, dtype=np.int64)
            off2 = np.ascontiguousarray(off2, dtype=np.int64)
            n_pairs = len(off1) - 1
            worst = int(off1[-1] + off2[-1])
        else:
            worst = (m + n) * n_pairs
        if out is None:
            cap = worst if capacity is None else int(capacity)         # untouched pages cost nothing
            out = (np.empty(n_pairs, dtype=PAIR_RESULT), np.empty(max(cap, 1), dtype=np.uint8),
                   np.empty(n_pairs + 1, dtype=np.int64))
        res, rle, rle_off = out
        self._check(self._L.ga_align_batch_rle(self._h, _ptr(s1), _ptr(off1), m, _ptr(s2), _ptr(off2), n, n_pairs,
                                               _ptr(res), _ptr(rle), len(rle), _ptr(rle_off), flags))
        return res, rle, rle_off

    # -- device-buffer entry points (pointers are ints, e.g. torch.Tensor.data_ptr())
    def plan_batch(self, off1, off2, ops_off=None, flags: int = 0):
        off1 = np.ascontiguousarray(off1, dtype=np.int64)
        off2 = np.ascontiguousarray(off2, dtype=np.int64)
        self._check(self._L.ga_plan_batch(self._h, _ptr(off1), _ptr(off2), len(off1) - 1,
                                          _ptr(ops_off), flags))
        return self.plan_info()

    def run_batch(self, d_s1, d_off1, d_s2, d_off2, d_cost, d_ops, d_ops_off, d_len, d_status):
        self._check(self._L.ga_run_batch(self._h, *[C.c_void_p(int(x or 0)) for x in
                                                    (d_s1, d_off1, d_s2, d_off2, d_cost, d_ops,
                                                     d_ops_off, d_len, d_status)]))

    def sync(self):
        self._check(self._L.ga_sync(self._h))

    def plan_info(self) -> dict:
        info = PlanInfo()
        self._check(self._L.ga_get_plan_info(self._h, C.byref(info)))
        return dict(fill_kernel=info.fill_kernel.decode(), rows_per_lane=info.rows_per_lane,
                    lanes_per_group=info.lanes_per_group, packed16=bool(info.packed16),
                    launches_per_run=info.launches_per_run, code_bytes=info.code_bytes,
                    cells=info.cells, n_chunks=info.n_chunks)

    def timing(self) -> dict:
        ms = (C.c_float * T_COUNT)()
        self._check(self._L.ga_get_timing(self._h, ms))
        return dict(h2d=ms[T_H2D], fill=ms[T_FILL], walk=ms[T_WALK], d2h=ms[T_D2H], total=ms[T_TOTAL])

    def fill_values(self, s1, s2, table, keep_boundary: bool):
        s1 = np.ascontiguousarray(s1, dtype=np.uint8)
        s2 = np.ascontiguousarray(s2, dtype=np.uint8)
        assert table.dtype == np.int32 and table.flags.c_contiguous
        self._check(self._L.ga_fill_values(self._h, _ptr(s1), len(s1), _ptr(s2), len(s2),
                                           1 if keep_boundary else 0, _ptr(table)))
        return table

    def fill_values_f64(self, dense, s1, s2, table, keep_boundary: bool):
        """dense: lowering.DenseSchemeF64.  table: float64 (m+1, n+1, 3), filled in place."""
        s1 = np.ascontiguousarray(s1, dtype=np.uint8)
        s2 = np.ascontiguousarray(s2, dtype=np.uint8)
        assert table.dtype == np.float64 and table.flags.c_contiguous
        self._check(self._L.ga_fill_values_f64(self._h, dense.A, _ptr(dense.sub), _ptr(dense.gap_h),
                                               _ptr(dense.gap_v), dense.gap_open, dense.max_cost, _ptr(s1),
                                               len(s1), _ptr(s2), len(s2), 1 if keep_boundary else 0, _ptr(table)))
        return table

    def align_f64(self, dense, s1, s2, draws=None):
        """One pair with float-valued penalties -> (cost float, ops walk order, n_eval, status)."""
        s1 = np.ascontiguousarray(s1, dtype=np.uint8)
        s2 = np.ascontiguousarray(s2, dtype=np.uint8)
        m, n = len(s1), len(s2)
        if draws is not None:
            draws = np.ascontiguousarray(draws, dtype=np.uint16)
        packed = np.zeros(4 * ((m + n + 15) // 16) + 16, dtype=np.uint8)
        cost, nops, neval, status = C.c_double(0.0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
        self._check(self._L.ga_align_f64(self._h, dense.A, _ptr(dense.sub), _ptr(dense.gap_h), _ptr(dense.gap_v),
                                         dense.gap_open, dense.max_cost, _ptr(s1), m, _ptr(s2), n,
                                         None if draws is None else _ptr(draws), 0 if draws is None else len(draws),
                                         C.byref(cost), _ptr(packed), C.byref(nops), C.byref(neval),
                                         C.byref(status)))
        return cost.value, unpack_ops(packed, 0, nops.value), neval.value, status.value

    def walk_values(self, s1, s2, table, draws=None):
        """The traceback alone on the caller's int32 (m+1, n+1, 3) value table
        -> (cost, ops walk order, n_eval, status)."""
        s1 = np.ascontiguousarray(s1, dtype=np.uint8)
        s2 = np.ascontiguousarray(s2, dtype=np.uint8)
        m, n = len(s1), len(s2)
        assert table.dtype == np.int32 and table.flags.c_contiguous and table.shape == (m + 1, n + 1, 3)
        if draws is not None:
            draws = np.ascontiguousarray(draws, dtype=np.uint16)
        packed = np.zeros(4 * ((m + n + 15) // 16) + 16, dtype=np.uint8)
        cost, nops, neval, status = C.c_int32(0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
        self._check(self._L.ga_walk_values(self._h, _ptr(s1), m, _ptr(s2), n, _ptr(table),
                                           None if draws is None else _ptr(draws),
                                           0 if draws is None else len(draws), C.byref(cost), _ptr(packed),
                                           C.byref(nops), C.byref(neval), C.byref(status)))
        return cost.value, unpack_ops(packed, 0, nops.value), neval.value, status.value

    def walk_values_f64(self, dense, s1, s2, table, draws=None):
        """The same for a float64 table and float-valued penalties (dense: DenseSchemeF64)."""
        s1 = np.ascontiguousarray(s1, dtype=np.uint8)
        s2 = np.ascontiguousarray(s2, dtype=np.uint8)
        m, n = len(s1), len(s2)
        assert table.dtype == np.float64 and table.flags.c_contiguous and table.shape == (m + 1, n + 1, 3)
        if draws is not None:
            draws = np.ascontiguousarray(draws, dtype=np.uint16)
        packed = np.zeros(4 * ((m + n + 15) // 16) + 16, dtype=np.uint8)
        cost, nops, neval, status = C.c_double(0.0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
        self._check(self._L.ga_walk_values_f64(self._h, dense.A, _ptr(dense.sub), _ptr(dense.gap_h),
                                               _ptr(dense.gap_v), dense.gap_open, dense.max_cost, _ptr(s1), m,
                                               _ptr(s2), n, _ptr(table), None if draws is None else _ptr(draws),
                                               0 if draws is None else len(draws), C.byref(cost), _ptr(packed),
                                               C.byref(nops), C.byref(neval), C.byref(status)))
        return cost.value, unpack_ops(packed, 0, nops.value), neval.value, status.value

    def alu_peak(self, mode: int) -> float:
        g = C.c_double(0.0)
        self._check(self._L.ga_alu_peak(self._h, int(mode), C.byref(g)))
        return g.value


    def align_long(self, s1, s2, flags: int = 0):
        """One long pair, linear-space traceback -> (cost, ops uint8 walk order | None, status, info)."""
        s1 = np.ascontiguousarray(s1, dtype=np.uint8)
        s2 = np.ascontiguousarray(s2, dtype=np.uint8)
        m, n = len(s1), len(s2)
        packed = np.zeros(4 * ((m + n + 15) // 16) + 16, dtype=np.uint8)
        cost, nops, status = C.c_int32(0), C.c_int64(0), C.c_int32(0)
        info = LongInfo()
        self._check(self._L.ga_align_long(self._h, _ptr(s1), m, _ptr(s2), n, C.byref(cost), _ptr(packed),
                                          C.byref(nops), C.byref(status), flags, C.byref(info)))
        d = {k: getattr(info, k) for k, _ in LongInfo._fields_}
        if flags & GA_F_COST_ONLY:
            return cost.value, None, 0, d
        return cost.value, unpack_ops(packed, 0, nops.value), status.value, d


    def align_replay(self, s1, s2, draws):
        """One pair with the reference's random tie-breaking replayed from `draws`
        (uint16 per traceback step) -> (cost, ops walk order, n_eval, status)."""
        s1 = np.ascontiguousarray(s1, dtype=np.uint8)
        s2 = np.ascontiguousarray(s2, dtype=np.uint8)
        draws = np.ascontiguousarray(draws, dtype=np.uint16)
        m, n = len(s1), len(s2)
        packed = np.zeros(4 * ((m + n + 15) // 16) + 16, dtype=np.uint8)
        cost, nops, neval, status = C.c_int32(0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
        self._check(self._L.ga_align_replay(self._h, _ptr(s1), m, _ptr(s2), n, _ptr(draws), len(draws),
                                            C.byref(cost), _ptr(packed), C.byref(nops), C.byref(neval),
                                            C.byref(status)))
        return cost.value, unpack_ops(packed, 0, nops.value), neval.value, status.value


def pack_letters(codes: np.ndarray, bits: int) -> np.ndarray:
    """uint8 letter indices -> little-endian `bits`-bit stream (bits = 2, 4 or 5), the layout
    GA_F_SEQ_2BIT / GA_F_SEQ_4BIT / GA_F_SEQ_5BIT expect (letter k in bits [k*bits, (k+1)*bits))."""
    if bits == 8:
        return np.ascontiguousarray(codes, dtype=np.uint8)
    if bits == 5:                      # 8 letters -> 5 bytes
        n = len(codes)
        padded = np.zeros(-(-n // 8) * 8, dtype=np.uint64)
        padded[:n] = codes
        word = np.zeros(len(padded) // 8, dtype=np.uint64)
        for k in range(8):
            word |= padded[k::8] << np.uint64(5 * k)
        out = np.zeros((len(word), 5), dtype=np.uint8)
        for b in range(5):
            out[:, b] = (word >> np.uint64(8 * b)) & np.uint64(0xFF)
        return out.reshape(-1)[: (n * 5 + 7) // 8 + 1]
    per = 8 // bits
    n = len(codes)
    padded = np.zeros(-(-n // per) * per, dtype=np.uint8)
    padded[:n] = codes
    lanes = padded.reshape(-1, per)
    out = np.zeros(len(lanes), dtype=np.uint8)
    for k in range(per):
        out |= (lanes[:, k] << (bits * k)).astype(np.uint8)
    return out


def seq_flag(bits: int) -> int:
    return {8: 0, 5: GA_F_SEQ_5BIT, 4: GA_F_SEQ_4BIT, 2: GA_F_SEQ_2BIT}[bits]


def ops_offsets(off1, off2) -> np.ndarray:
    """Byte offsets of each pair's packed-ops slice: ceil((m+n)/4) rounded up to 4."""
    lens = (np.diff(off1) + np.diff(off2)).astype(np.int64)
    nbytes = (((lens + 3) // 4) + 3) // 4 * 4
    out = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(nbytes, out=out[1:])
    return out


#: one record of ga_align_batch_rle (ga_pair_result)
PAIR_RESULT = np.dtype([("cost", "<i4"), ("nops_status", "<u4")])


def unpack_rle(rle: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """Run-length encoded ops of one pair (bytes [lo, hi) of the compact stream) as a uint8 array
    of 0/1/2 in walk order -- the same array unpack_ops gives for the fixed-slice form."""
    b = rle[lo:hi]
    return np.repeat(b & 3, (b >> 2).astype(np.int64) + 1)


def unpack_ops(ops: np.ndarray, start: int, count: int) -> np.ndarray:
    """2-bit ops of one pair (walk order) as a uint8 array of 0/1/2."""
    nbytes = (count + 3) // 4
    b = ops[start:start + nbytes]
    out = np.empty(nbytes * 4, dtype=np.uint8)
    out[0::4] = b & 3
    out[1::4] = (b >> 2) & 3
    out[2::4] = (b >> 4) & 3
    out[3::4] = (b >> 6) & 3
    return out[:count]


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
