"""Batched alignment: many pairs, one scheme, one GPU call.

The reference aligns exactly one pair per call (globaligner.py:132-314); the
BASELINE workloads (10^5 .. 10^7 pairs) need a batched entry point.  It keeps the
reference's scheme keyword arguments and returns results that materialise the
reference's ``AlignmentResults`` lazily (building 10^6 nested dicts eagerly would
cost far more than the alignment itself).
"""

from __future__ import annotations

import threading

import numpy as np

from . import _native, start
from .conclude import AlignmentResults, final_cost_to_score
from .lowering import DenseScheme

GAP = "-"


class ReferenceUndefinedError(RuntimeError):
    """The reference's traceback would leave the DP table for this input (it
    raises IndexError or silently wraps to negative indices).  Only possible when
    max(len(seq_1), len(seq_2)) <= gap_open_cost."""


def render(ops: np.ndarray, seq_1: str, seq_2: str):
    """Walk-order ops (0 diag, 1 gap in seq_1, 2 gap in seq_2) -> the three
    strings of the reference's emitters (globaligner.py:688-753, reversed :584-593)."""
    fwd = ops[::-1]
    take1 = fwd != 1                     # columns that consume a seq_1 letter
    take2 = fwd != 2
    a = np.frombuffer(seq_1.encode("utf-32-le"), dtype=np.uint32)
    b = np.frombuffer(seq_2.encode("utf-32-le"), dtype=np.uint32)
    top = np.full(len(fwd), ord(GAP), dtype=np.uint32)
    bot = np.full(len(fwd), ord(GAP), dtype=np.uint32)
    top[take1] = a
    bot[take2] = b
    mid = np.full(len(fwd), ord(" "), dtype=np.uint32)
    diag = fwd == 0
    mid[diag] = np.where(top[diag] == bot[diag], ord("|"), ord("*"))
    dec = lambda x: x.tobytes().decode("utf-32-le")
    return dec(top), dec(mid), dec(bot)


class BatchAlignment:
    """Packed results of a batch; ``batch[k]`` builds the k-th AlignmentResults."""

    def __init__(self, seqs_1, seqs_2, cost, ops, ops_off, ops_len, status, scheme, rle: bool = False, lens=None):
        """`rle`: ops / ops_off hold the compact run-length stream of ga_align_batch_rle instead of
        fixed 2-bit slices (what find_global_alignments asks the library for)."""
        self.seqs_1, self.seqs_2 = seqs_1, seqs_2
        self.cost, self.ops, self.ops_off, self.ops_len, self.status = cost, ops, ops_off, ops_len, status
        self.rle = rle
        self._patched_ops = {}       # pairs re-aligned one by one (_realign_single_letter_pairs)
        self._scheme = scheme
        # (ASCII upper-casing keeps lengths: measure the caller's strings, not the lazy view)
        if lens is not None:
            m, n = lens
        else:
            m = np.fromiter(map(len, getattr(seqs_1, "_seqs", seqs_1)), dtype=np.int64, count=len(seqs_1))
            n = np.fromiter(map(len, getattr(seqs_2, "_seqs", seqs_2)), dtype=np.int64, count=len(seqs_2))
        self.score = final_cost_to_score(cost.astype(np.int64), m, n, scheme.max_score)

    def __len__(self):
        return len(self.cost)

    def ops_of(self, k: int) -> np.ndarray:
        if k in self._patched_ops:
            return self._patched_ops[k]
        if self.rle:
            return _native.unpack_rle(self.ops, int(self.ops_off[k]), int(self.ops_off[k + 1]))
        return _native.unpack_ops(self.ops, int(self.ops_off[k]), int(self.ops_len[k]))

    def __getitem__(self, k: int) -> AlignmentResults:
        if self.status[k] != 0:
            raise ReferenceUndefinedError(
                f"pair {k}: the reference's traceback leaves the DP table for this input "
                f"(max(m, n) <= gap_open_cost)")
        s1, s2 = self.seqs_1[k], self.seqs_2[k]
        top, mid, bot = render(self.ops_of(k), s1, s2)
        scoring_mat, costing_mat = self._scheme.matrices_for(s1, s2)
        return AlignmentResults(top, mid, bot, int(self.cost[k]), int(self.score[k]), scoring_mat,
                                costing_mat, self._scheme.gap_open_score, self._scheme.gap_open_cost,
                                None)

    def __iter__(self):
        return (self[k] for k in range(len(self)))

    def write(self, file, descs=None, chars_per_line: int = 70):
        """Write every alignment, one after the other, in the reference's printout format
        (AlignmentResults.__str__, conclude.py:19-89).  `descs`: optional list of
        (desc_1, desc_2) per pair (e.g. FASTA header lines)."""
        with open(file, mode="w+") as fh:
            for k in range(len(self)):
                d1, d2 = descs[k] if descs is not None else (f"seq_1[{k}]", f"seq_2[{k}]")
                fh.write(self[k].__str__(desc_1=d1, desc_2=d2, chars_per_line=chars_per_line))
                fh.write("\n")


class Scheme:
    """A resolved scoring/costing scheme, independent of any particular pair."""

    def __init__(self, scoring_mat_name=None, scoring_mat_path=None, match_score=None,
                 mismatch_score=None, mismatch_cost=None, gap_open_score=None, gap_open_cost=None,
                 gap_extension_score=None, gap_extension_cost=None):
        self.branch, self.scoring, self.costing = start.resolve_scheme_options(
            scoring_mat_name, scoring_mat_path, match_score, mismatch_score, mismatch_cost,
            gap_open_score, gap_open_cost, gap_extension_score, gap_extension_cost)
        self.gap_open_score = self.scoring.gap_open_score
        self.gap_open_cost = self.costing.gap_open_cost
        if self.branch in ("name", "path"):
            self.scoring_mat = start.load_matrix_scheme(self.branch, scoring_mat_name, scoring_mat_path)
            self.costing_mat = start.scoring_mat_to_costing_mat(
                self.scoring_mat, max_score=start.get_max_val(self.scoring_mat))
            self.max_score = start.get_max_val(self.scoring_mat)
        else:
            self.scoring_mat = self.costing_mat = None
            # simple schemes: max of the scoring matrix is the match score
            # (the "-","-" corner and the diagonal; start.py:431-449, 559-612)
            self.max_score = self.scoring.match_score

    def matrices_for(self, seq_1: str, seq_2: str):
        """The (scoring_mat, costing_mat) the reference would report for this pair:
        matrix schemes are pair-independent, simple schemes are built over the
        letters of the pair (start.py:319-343)."""
        if self.scoring_mat is not None:
            return self.scoring_mat, self.costing_mat
        return start.simple_scheme_matrices(self.branch, start.get_common_alphabet(seq_1, seq_2),
                                            self.scoring, self.costing)

    def dense_for(self, alphabet) -> DenseScheme:
        """Dense tables over `alphabet` (sorted letters of the whole batch)."""
        if self.scoring_mat is not None:
            missing = set(alphabet).difference(self.scoring_mat.keys())
            if missing:
                raise RuntimeError(
                    f"common_alphabet contains values not in scoring_mat_keys, e.g. {missing}.  "
                    f"Please check your sequences and your scoring matrix.")
            # tables over the letters that occur (in the matrix's own order): BLOSUM62 has 24 rows,
            # protein reads use 20 -- smaller tables keep the bank-conflict-free kernel at three
            # blocks per SM.  max_cost stays the maximum of the WHOLE matrix (globaligner.py:258).
            present = set(alphabet)
            letters = [k for k in self.costing_mat if k != GAP and k in present]
            return DenseScheme(self.costing_mat, self.gap_open_cost, letters)
        _, cm = start.simple_scheme_matrices(self.branch, list(alphabet), self.scoring, self.costing)
        return DenseScheme(cm, self.gap_open_cost)


_GAP_MSG = ("The current implementation does not allow for '-' characters in the sequences "
            "because they are used internally for gaps.  Please replace this character in "
            "your sequences.")


def _lower_batch(seqs_1, seqs_2, limit, scheme):
    """Pair by pair: the checks of validate_and_transform_args (start.py:209-232), upper-casing
    and encoding.  Handles any str (non-ASCII letters, case mappings that change length)."""
    up1, up2 = [], []
    for s1, s2 in zip(seqs_1, seqs_2):
        start.check_seq_lengths(s1, s2, limit)
        if GAP in s1 or GAP in s2:
            raise RuntimeError(_GAP_MSG)
        up1.append(s1.upper())
        up2.append(s2.upper())
    alphabet = sorted(set().union(*map(set, up1), *map(set, up2)))
    dense = scheme.dense_for(alphabet)
    enc1 = [dense.encode(s) for s in up1]
    enc2 = [dense.encode(s) for s in up2]
    off1 = np.zeros(len(up1) + 1, dtype=np.int64)
    off2 = np.zeros(len(up2) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in enc1], out=off1[1:])
    np.cumsum([len(x) for x in enc2], out=off2[1:])
    return up1, up2, dense, np.concatenate(enc1), off1, np.concatenate(enc2), off2


_PINNED = threading.local()


def _pinned_bytes(role: str, nbytes: int) -> np.ndarray:
    """A reusable page-locked host buffer (per thread and role) as a uint8 array: the letters are
    written once, by the lowering threads, where the copy engine reads them from.  Plain memory
    when there is no CUDA device (CPU-only tests of the lowering)."""
    pool = _PINNED.__dict__.setdefault("pool", {})
    t = pool.get(role)
    if t is None or t.numel() < nbytes:
        import torch
        cap = max(nbytes, 1 << 20) * 5 // 4
        t = torch.empty(cap, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        pool[role] = t
    return t.numpy()[:nbytes]


def _letters_of(cat: np.ndarray, off: np.ndarray, k: int, bits: int) -> np.ndarray:
    """Letter indices of sequence k of a lowered batch (bits = 8: one byte per letter; 2 / 4: the
    C-ABI's packed letter streams)."""
    lo, hi = int(off[k]), int(off[k + 1])
    if bits == 8:
        return cat[lo:hi]
    bit = np.arange(lo, hi, dtype=np.int64) * bits
    two = cat[bit >> 3].astype(np.uint16) | (cat[np.minimum((bit >> 3) + 1, len(cat) - 1)].astype(np.uint16) << 8)
    return ((two >> (bit & 7).astype(np.uint16)) & ((1 << bits) - 1)).astype(np.uint8)


def _lower_native(seqs_1, seqs_2, limit, scheme, pack=False):
    """The all-ASCII case through the `_strings` extension (csrc/ga_strings.c): threaded passes
    over the strings' own buffers -- headers, which byte values occur, then every letter mapped
    through one 256-entry table (case folding + letter -> index) straight into page-locked
    memory -- instead of join / encode / translate in Python.  `pack`: alphabets of up to 4 / 16 /
    32 letters are written as the C-ABI's 2- / 4- / 5-bit letter streams (a quarter / half / five
    eighths of the upload).  -> (..., bits) or None when the extension is not built, the inputs are not lists of
    ASCII str, or anything needs the per-pair path's exact error."""
    try:
        from . import _strings
    except ImportError:
        return None
    if not (isinstance(seqs_1, list) and isinstance(seqs_2, list)):
        return None
    sc1 = _strings.scan(seqs_1)
    sc2 = _strings.scan(seqs_2) if sc1 is not None else None
    if sc1 is None or sc2 is None:
        return None
    (h1, _t1, l1, seen1), (h2, _t2, l2, seen2) = sc1, sc2
    m_len = np.frombuffer(l1, dtype=np.int64)
    n_len = np.frombuffer(l2, dtype=np.int64)
    seen = np.frombuffer(seen1, dtype=np.uint8) | np.frombuffer(seen2, dtype=np.uint8)
    if seen[ord(GAP)] or bool((~(m_len * n_len < limit) | (m_len * n_len == 0)).any()):
        return None                      # the general path raises the reference's error for the first offending pair
    known = {chr(c).upper() for c in np.flatnonzero(seen)}
    if any(len(ch) != 1 or ord(ch) > 127 for ch in known):
        return None
    dense = scheme.dense_for(sorted(known))
    table = bytearray(256)
    for ch, k in dense.index.items():
        if ord(ch) < 128:
            table[ord(ch)] = k
            table[ord(ch.lower())] = k
    table = bytes(table)
    bits = 8
    if pack:
        bits = 2 if dense.A <= 4 else 4 if dense.A <= 16 else 5 if dense.A <= 32 else 8
    cat1 = _pinned_bytes("s1", (_t1 * bits + 7) // 8 + 16)
    cat2 = _pinned_bytes("s2", (_t2 * bits + 7) // 8 + 16)
    _strings.translate_into(h1, table, cat1, bits)
    _strings.translate_into(h2, table, cat2, bits)
    n = len(seqs_1)
    off1 = np.zeros(n + 1, dtype=np.int64)
    off2 = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(m_len, out=off1[1:])
    np.cumsum(n_len, out=off2[1:])
    return _LazyUpper(seqs_1), _LazyUpper(seqs_2), dense, cat1, off1, cat2, off2, bits


def _lower_ascii_batch(seqs_1, seqs_2, limit, scheme):
    """The same for all-ASCII input without a Python-level loop over the pairs: the sequences are
    joined, upper-cased and mapped through one 256-entry table (large read sets spend their time
    here, not on the GPU).  -> None when a sequence is not ASCII (the caller then goes pair by
    pair); raises exactly what the per-pair path raises, for the first offending pair."""
    n = len(seqs_1)
    try:
        j1, j2 = "".join(seqs_1), "".join(seqs_2)
    except TypeError:
        return None
    if not (j1.isascii() and j2.isascii()):
        return None
    m_len = np.fromiter(map(len, seqs_1), dtype=np.int64, count=n)
    n_len = np.fromiter(map(len, seqs_2), dtype=np.int64, count=n)
    bad = ~(m_len * n_len < limit) | (m_len * n_len == 0)
    first_bad = int(np.argmax(bad)) if bad.any() else n
    if GAP in j1 or GAP in j2:
        for k in range(first_bad):           # only when there is a '-' somewhere
            if GAP in seqs_1[k] or GAP in seqs_2[k]:
                raise RuntimeError(_GAP_MSG)
    if first_bad < n:
        start.check_seq_lengths(seqs_1[first_bad], seqs_2[first_bad], limit)
    e1, e2 = j1.encode("ascii"), j2.encode("ascii")
    # the letters that occur (upper-cased), found with bytes.translate at C speed: delete what is
    # known, look at the start of what is left, repeat
    known = set()
    for e in (e1, e2):
        rest = e
        while rest:
            known |= set(rest[:4096].upper())
            both = bytes(known) + bytes(known).lower()
            rest = rest.translate(None, both)
    alphabet = [chr(c) for c in sorted(known)]                      # sorted by code point, like sorted(set)
    dense = scheme.dense_for(alphabet)
    table = bytearray(256)
    for ch, k in dense.index.items():
        if ord(ch) < 128:
            table[ord(ch)] = k
            table[ord(ch.lower())] = k                               # case folding and encoding in one pass
    table = bytes(table)
    cat1 = np.frombuffer(e1.translate(table), dtype=np.uint8)
    cat2 = np.frombuffer(e2.translate(table), dtype=np.uint8)
    off1 = np.zeros(n + 1, dtype=np.int64)
    off2 = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(m_len, out=off1[1:])
    np.cumsum(n_len, out=off2[1:])
    up1 = _LazyUpper(seqs_1)
    up2 = _LazyUpper(seqs_2)
    return up1, up2, dense, cat1, off1, cat2, off2


class _LazyUpper:
    """Sequence view that upper-cases on access (BatchAlignment renders pairs lazily)."""

    def __init__(self, seqs):
        self._seqs = seqs

    def __len__(self):
        return len(self._seqs)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [s.upper() for s in self._seqs[k]]
        return self._seqs[k].upper()

    def __iter__(self):
        return (s.upper() for s in self._seqs)


def find_global_alignments(seqs_1, seqs_2, *, device: int = 0, devices=None, max_seq_len_prod=None,
                           scoring_mat_name=None, scoring_mat_path=None, match_score=None,
                           mismatch_score=None, mismatch_cost=None, gap_open_score=None,
                           gap_open_cost=None, gap_extension_score=None, gap_extension_cost=None
                           ) -> BatchAlignment:
    """Align seqs_1[k] with seqs_2[k] for every k under one scheme.

    Scheme keywords mean exactly what they mean for ``find_global_alignment``.
    ``max_seq_len_prod`` defaults to the reference's guard (start.py:213); pass a
    larger number (or ``float('inf')``) to lift it.  ``devices``: a list of GPU indices to
    shard the batch over (equal DP cells per GPU, one host thread and context each, results
    merged in pair order); default: the single GPU ``device``.
    """
    if len(seqs_1) != len(seqs_2) or len(seqs_1) == 0:
        raise RuntimeError("seqs_1 and seqs_2 must be non-empty and of equal length")
    limit = start.MAX_SEQ_LEN_PRODUCT if max_seq_len_prod is None else max_seq_len_prod
    scheme = Scheme(scoring_mat_name, scoring_mat_path, match_score, mismatch_score, mismatch_cost,
                    gap_open_score, gap_open_cost, gap_extension_score, gap_extension_cost)
    one_gpu = devices is None or len(devices) <= 1
    bits = 8
    lowered = _lower_native(seqs_1, seqs_2, limit, scheme, pack=one_gpu)    # (shards are cut at letter offsets: bytes)
    if lowered is not None:
        bits, lowered = lowered[-1], lowered[:-1]
    if lowered is None:
        lowered = _lower_ascii_batch(seqs_1, seqs_2, limit, scheme)
    if lowered is None:
        lowered = _lower_batch(seqs_1, seqs_2, limit, scheme)
    up1, up2, dense, cat1, off1, cat2, off2 = lowered
    seq_flags = _native.seq_flag(bits)
    def on_device(dev, c1, o1, c2, o2):
        """-> (cost, rle, rle_off, ops_len, status): the compact result stream of one GPU."""
        ctx = _native.default_context(dev)
        ctx.set_scheme(dense)
        n_pairs = len(o1) - 1
        m0, n0 = int(o1[1]), int(o2[1])
        if n_pairs > 1 and o1[-1] == m0 * n_pairs and o2[-1] == n0 * n_pairs and \
                np.array_equal(o1, np.arange(n_pairs + 1) * m0) and np.array_equal(o2, np.arange(n_pairs + 1) * n0):
            # fixed-length batch (reads of one length): the entry point without offset arrays
            res, rle, rle_off = ctx.align_batch_rle(c1, None, c2, None, m0, n0, n_pairs, flags=seq_flags)
        else:
            res, rle, rle_off = ctx.align_batch_rle(c1, o1, c2, o2, flags=seq_flags)
        nops = res["nops_status"]
        return (res["cost"].copy(), rle[: int(rle_off[-1])], rle_off, (nops & 0x7FFFFFFF).astype(np.int32),
                (nops >> 31).astype(np.int32))

    if one_gpu:
        cost, ops, ops_off, ops_len, status = on_device(device if not devices else devices[0], cat1, off1, cat2, off2)
    else:
        from .parallel import align_on_devices
        cost, ops, ops_off, ops_len, status = align_on_devices(on_device, cat1, off1, cat2, off2, list(devices))
    res = BatchAlignment(up1, up2, cost, ops, ops_off, ops_len, status, scheme, rle=True,
                         lens=(np.diff(off1), np.diff(off2)))
    _realign_single_letter_pairs(res, dense, cat1, off1, cat2, off2, scheme, device if not devices else devices[0], bits)
    return res


def _realign_single_letter_pairs(res: "BatchAlignment", dense, cat1, off1, cat2, off2, scheme, device, bits=8):
    """Simple schemes build their matrices over the letters of the PAIR (start.py:319-343), so a
    pair made of one repeated letter has no mismatch entry and a smaller max_cost -- hence a
    smaller boundary sentinel (globaligner.py:777) -- than the batch-wide tables give it.  The
    sentinel only matters when max(m, n) <= gap_open_cost, where the
    reference's result is sentinel-dependent (or undefined): those few pairs are re-aligned one
    by one with their own matrices so that the batch call agrees with find_global_alignment."""
    if scheme.scoring_mat is not None or dense.A < 2:
        return
    m_len, n_len = np.diff(off1), np.diff(off2)
    short = np.flatnonzero(np.maximum(m_len, n_len) <= scheme.gap_open_cost)
    if short.size == 0:
        return
    ctx = _native.default_context(device)
    patched = {}
    for k in short:
        a = _letters_of(cat1, off1, int(k), bits)
        b = _letters_of(cat2, off2, int(k), bits)
        if not (a.min() == a.max() == b.min() == b.max()):
            continue
        s1, s2 = res.seqs_1[int(k)], res.seqs_2[int(k)]
        _, cm = scheme.matrices_for(s1, s2)
        own = DenseScheme(cm, scheme.gap_open_cost)
        ctx.set_scheme(own)
        c1, o1, _, l1, st1 = ctx.align_batch(own.encode(s1), np.array([0, len(s1)], dtype=np.int64),
                                             own.encode(s2), np.array([0, len(s2)], dtype=np.int64))
        res.cost[k], res.ops_len[k], res.status[k] = c1[0], l1[0], st1[0]
        patched[int(k)] = _native.unpack_ops(o1, 0, int(l1[0]))
    ctx.set_scheme(dense)
    if patched:
        res.score = final_cost_to_score(res.cost.astype(np.int64), m_len, n_len, scheme.max_score)
        res._patched_ops = patched


def read_pairs_from_fasta(fasta_path, fasta_path_2=None):
    """Batched FASTA ingestion.  One file: consecutive records are paired
    (1st with 2nd, 3rd with 4th, ...; the reference reads only the first two,
    start.py:666-688).  Two files: record k of the first with record k of the second.
    -> (descs [(desc_1, desc_2)], seqs_1, seqs_2), sequences upper-cased as the reference's
    reader does (start.py:643)."""
    recs = list(start.read_seq_from_fasta(fasta_path))
    if fasta_path_2 is None:
        if len(recs) < 2 or len(recs) % 2:
            raise RuntimeError("the FASTA file must hold an even, non-zero number of records")
        left, right = recs[0::2], recs[1::2]
    else:
        left, right = recs, list(start.read_seq_from_fasta(fasta_path_2))
        if len(left) != len(right) or not left:
            raise RuntimeError("the two FASTA files must hold the same, non-zero number of records")
    descs = [(a[0], b[0]) for a, b in zip(left, right)]
    return descs, [a[1] for a in left], [b[1] for b in right]


def read_pairs_from_fastq(fastq_path, fastq_path_2=None):
    """The same for FASTQ (plain or .gz): one file with consecutive records paired (interleaved
    reads), or two files read in step (R1 / R2).  -> (descs, seqs_1, seqs_2)"""
    recs = list(start.read_seq_from_fastq(fastq_path))
    if fastq_path_2 is None:
        if len(recs) < 2 or len(recs) % 2:
            raise RuntimeError("the FASTQ file must hold an even, non-zero number of records")
        left, right = recs[0::2], recs[1::2]
    else:
        left, right = recs, list(start.read_seq_from_fastq(fastq_path_2))
        if len(left) != len(right) or not left:
            raise RuntimeError("the two FASTQ files must hold the same, non-zero number of records")
    descs = [(a[0], b[0]) for a, b in zip(left, right)]
    return descs, [a[1] for a in left], [b[1] for b in right]


def find_global_alignments_from_fastq(fastq_path, fastq_path_2=None, **kwargs):
    """-> (descs, BatchAlignment) for every record pair of the FASTQ file(s)."""
    descs, s1, s2 = read_pairs_from_fastq(fastq_path, fastq_path_2)
    return descs, find_global_alignments(s1, s2, **kwargs)


def find_global_alignments_from_fasta(fasta_path, fasta_path_2=None, **kwargs):
    """-> (descs, BatchAlignment) for every record pair of the file(s)."""
    descs, s1, s2 = read_pairs_from_fasta(fasta_path, fasta_path_2)
    return descs, find_global_alignments(s1, s2, **kwargs)


def align_encoded(batch, scheme_kwargs: dict, *, device: int = 0, flags: int = 0):
    """Fast path for already-encoded batches (``synth.PairBatch``-like objects with
    .alphabet/.s1/.off1/.s2/.off2): no Python strings are touched.
    -> (cost, ops, ops_off, ops_len, status, Scheme, DenseScheme)"""
    scheme = Scheme(**scheme_kwargs)
    dense_any = scheme.dense_for(sorted(batch.alphabet))
    # re-index the batch's letter codes into the scheme's letter order
    remap = np.array([dense_any.index[ch] for ch in batch.alphabet], dtype=np.uint8)
    identity = np.array_equal(remap, np.arange(len(remap), dtype=np.uint8))
    s1 = batch.s1 if identity else remap[batch.s1]
    s2 = batch.s2 if identity else remap[batch.s2]
    ctx = _native.default_context(device)
    ctx.set_scheme(dense_any)
    out = ctx.align_batch(s1, batch.off1, s2, batch.off2, flags)
    return (*out, scheme, dense_any)
