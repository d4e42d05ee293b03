"""Synthetic sequence workloads of the BASELINE.json shapes.

Everything is numpy-vectorised and seeded with ``numpy.random.default_rng`` so
the same pairs can be produced on the GPU box, in the build container (for the
golden vectors) and in the CPU tests.  Sequences are returned as uint8 letter
*indices* into the given alphabet plus int64 offsets (CSR layout), which is what
the C-ABI consumes; ``decode`` turns a row back into a ``str`` for the
reference-facing API.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

DNA = "ACGT"
AMINO = "ARNDCQEGHILKMFPSTWYV"          # the 20 standard residues (BLOSUM order)


@dataclass
class PairBatch:
    """CSR batch of pairs: letters are indices into `alphabet`."""
    alphabet: str
    s1: np.ndarray        # uint8, concatenated seq_1 letters
    off1: np.ndarray      # int64, n_pairs + 1
    s2: np.ndarray
    off2: np.ndarray

    @property
    def n_pairs(self) -> int:
        return len(self.off1) - 1

    def lens(self):
        return np.diff(self.off1), np.diff(self.off2)

    def cells(self) -> int:
        m, n = self.lens()
        return int(np.sum(m.astype(np.int64) * n.astype(np.int64)))

    def pair(self, k: int) -> tuple[str, str]:
        a = self.s1[self.off1[k]:self.off1[k + 1]]
        b = self.s2[self.off2[k]:self.off2[k + 1]]
        return decode(a, self.alphabet), decode(b, self.alphabet)

    def head(self, k: int) -> "PairBatch":
        return PairBatch(self.alphabet, self.s1[: self.off1[k]], self.off1[: k + 1].copy(),
                         self.s2[: self.off2[k]], self.off2[: k + 1].copy())

    def slice(self, lo: int, hi: int) -> "PairBatch":
        o1 = self.off1[lo:hi + 1] - self.off1[lo]
        o2 = self.off2[lo:hi + 1] - self.off2[lo]
        return PairBatch(self.alphabet, self.s1[self.off1[lo]:self.off1[hi]], o1.copy(),
                         self.s2[self.off2[lo]:self.off2[hi]], o2.copy())


def decode(codes: np.ndarray, alphabet: str) -> str:
    lut = np.frombuffer(alphabet.encode("ascii"), dtype=np.uint8)
    return lut[codes].tobytes().decode("ascii")


def _mutate_fixed(rng, s1: np.ndarray, n_letters: int, p_sub: float, p_ins: float,
                  p_del: float) -> np.ndarray:
    """Point mutations on an (N, L) matrix; the result is trimmed / padded (with
    fresh random letters) back to L columns."""
    N, L = s1.shape
    u = rng.random((N, L), dtype=np.float32)
    shift = rng.integers(1, n_letters, size=(N, L), dtype=np.uint8)
    base = np.where(u < p_sub, (s1 + shift) % n_letters, s1).astype(np.uint8)
    deleted = (u >= p_sub) & (u < p_sub + p_del)
    inserted = (u >= p_sub + p_del) & (u < p_sub + p_del + p_ins)
    emit = np.ones((N, L), dtype=np.int32)
    emit[deleted] = 0
    emit[inserted] = 2
    start = np.cumsum(emit, axis=1) - emit
    out = rng.integers(0, n_letters, size=(N, L), dtype=np.uint8)      # padding letters
    extra = rng.integers(0, n_letters, size=(N, L), dtype=np.uint8)    # inserted letters
    rows = np.broadcast_to(np.arange(N)[:, None], (N, L))
    keep = (emit >= 1) & (start < L)
    out[rows[keep], start[keep]] = base[keep]
    ins = (emit == 2) & (start + 1 < L)
    out[rows[ins], start[ins] + 1] = extra[ins]
    return out


_BLOCK = 4096      # pairs per independently seeded block (keeps prefixes stable)


def _for_blocks(fn, jobs):
    """Blocks are independent (own seed, own output rows): large workloads are drawn on a few
    host threads (numpy releases the GIL inside its generators and ufuncs)."""
    import os
    threads = min(len(jobs), max(1, min(16, (os.cpu_count() or 1))))
    if threads <= 1 or len(jobs) < 8:
        return [fn(j) for j in jobs]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=threads) as pool:
        return list(pool.map(fn, jobs))


def dna_fixed(n_pairs: int, length: int, seed: int, p_sub=0.08, p_ins=0.01, p_del=0.01,
              first_block: int = 0) -> PairBatch:
    """Configs 2 and 5: `n_pairs` DNA pairs, both of exactly `length` letters.

    Pairs are drawn in blocks of 4096 from ``default_rng([seed, block])`` so the
    first k pairs do not depend on n_pairs (CPU samples are prefixes).  `first_block`
    skips that many blocks: rank r of a sharded run generates only its own shard."""
    s1 = np.empty((n_pairs, length), dtype=np.uint8)
    s2 = np.empty((n_pairs, length), dtype=np.uint8)

    def one(job):
        blk, lo = job
        hi = min(n_pairs, lo + _BLOCK)
        rng = np.random.default_rng([seed, first_block + blk])
        a = rng.integers(0, 4, size=(_BLOCK, length), dtype=np.uint8)
        b = _mutate_fixed(rng, a, 4, p_sub, p_ins, p_del)
        s1[lo:hi] = a[: hi - lo]
        s2[lo:hi] = b[: hi - lo]

    _for_blocks(one, list(enumerate(range(0, n_pairs, _BLOCK))))
    off = np.arange(n_pairs + 1, dtype=np.int64) * length
    return PairBatch(DNA, s1.reshape(-1), off, s2.reshape(-1), off.copy())


def dna_single(length: int, seed_1: int, seed_2: int) -> PairBatch:
    """Configs 1 and 4: one DNA pair, both `length` letters (seeds 1001/1002, 4001/4002)."""
    a = np.random.default_rng(seed_1).integers(0, 4, size=(1, length), dtype=np.uint8)
    b = _mutate_fixed(np.random.default_rng(seed_2), a, 4, 0.08, 0.01, 0.01)
    off = np.array([0, length], dtype=np.int64)
    return PairBatch(DNA, a.reshape(-1), off, b.reshape(-1), off.copy())


def _protein_block(rng, n_pairs, lo, hi, p_sub, p_indel, mean_indel) -> PairBatch:
    m = rng.integers(lo, hi + 1, size=n_pairs, dtype=np.int64)
    off1 = np.zeros(n_pairs + 1, dtype=np.int64)
    np.cumsum(m, out=off1[1:])
    s1 = rng.integers(0, 20, size=int(off1[-1]), dtype=np.uint8)
    total = len(s1)
    u = rng.random(total, dtype=np.float32)
    shift = rng.integers(1, 20, size=total, dtype=np.uint8)
    base = np.where(u < p_sub, (s1 + shift) % 20, s1).astype(np.uint8)
    ev = (u >= p_sub) & (u < p_sub + p_indel)
    is_del = ev & (rng.random(total, dtype=np.float32) < 0.5)
    is_ins = ev & ~is_del
    glen = rng.geometric(1.0 / mean_indel, size=total).astype(np.int64)
    # a deletion removes `glen` letters starting here (difference-array marking)
    dmark = np.zeros(total + 1, dtype=np.int64)
    didx = np.nonzero(is_del)[0]
    np.add.at(dmark, didx, 1)
    np.add.at(dmark, np.minimum(didx + glen[didx], total), -1)
    removed = np.cumsum(dmark[:-1]) > 0
    emit = np.where(removed, 0, 1).astype(np.int64)
    emit = np.where(is_ins & ~removed, 1 + glen, emit)
    csum = np.concatenate([[0], np.cumsum(emit)])
    raw = csum[off1[1:]] - csum[off1[:-1]]
    n = np.clip(raw, lo, hi)                                           # clamp to [lo, hi]
    off2 = np.zeros(n_pairs + 1, dtype=np.int64)
    np.cumsum(n, out=off2[1:])
    s2 = rng.integers(0, 20, size=int(off2[-1]), dtype=np.uint8)       # padding / inserted letters
    pair_of = np.repeat(np.arange(n_pairs), m)
    local = csum[:-1] - csum[off1[:-1]][pair_of]                        # output position in own pair
    ok = (emit >= 1) & (local < n[pair_of])
    s2[off2[:-1][pair_of[ok]] + local[ok]] = base[ok]
    return PairBatch(AMINO, s1, off1, s2, off2)


def protein_varlen(n_pairs: int, seed: int, lo: int = 300, hi: int = 1000, p_sub=0.15,
                   p_indel=0.05, mean_indel=3.0) -> PairBatch:
    """Config 3: protein pairs, len(seq_1) ~ U{lo..hi}; seq_2 = seq_1 with 15 %
    substitutions and 5 % indel events of geometric length (mean 3), clamped to
    [lo, hi] letters.  Drawn in blocks of 4096 pairs (prefix-stable)."""
    def one(job):
        blk, start = job
        rng = np.random.default_rng([seed, blk])
        full = _protein_block(rng, _BLOCK, lo, hi, p_sub, p_indel, mean_indel)
        return full.head(min(_BLOCK, n_pairs - start))

    return concat(_for_blocks(one, list(enumerate(range(0, n_pairs, _BLOCK)))))


def concat(parts: list[PairBatch]) -> PairBatch:
    s1 = np.concatenate([p.s1 for p in parts])
    s2 = np.concatenate([p.s2 for p in parts])
    off1 = [np.zeros(1, dtype=np.int64)]
    off2 = [np.zeros(1, dtype=np.int64)]
    b1 = b2 = 0
    for p in parts:
        off1.append(p.off1[1:] + b1)
        off2.append(p.off2[1:] + b2)
        b1 += int(p.off1[-1])
        b2 += int(p.off2[-1])
    return PairBatch(parts[0].alphabet, s1, np.concatenate(off1), s2, np.concatenate(off2))


#: name -> (builder, human description).  Sizes are the BASELINE.json ones;
#: every builder accepts n_pairs / length overrides for bounded samples.
def workload(name: str, n_pairs: int | None = None, first_block: int = 0) -> PairBatch:
    if name == "cfg1":
        return dna_single(1000, 1001, 1002)
    if name == "cfg2":
        return dna_fixed(n_pairs or 1_000_000, 150, 2000, first_block=first_block)
    if name == "cfg3":
        return protein_varlen(n_pairs or 100_000, 3000)
    if name == "cfg4":
        return dna_single(100_000, 4001, 4002)
    if name == "cfg5":
        return dna_fixed(n_pairs or 10_000_000, 250, 5000, first_block=first_block)
    raise ValueError(f"unknown workload {name!r}")


#: scheme keyword arguments (reference `find_global_alignment` names) per config
SCHEMES = {
    "cfg1": dict(match_score=2, mismatch_score=-3, gap_open_score=-10, gap_extension_score=-1),
    "cfg2": dict(match_score=2, mismatch_score=-3, gap_open_score=-10, gap_extension_score=-1),
    "cfg3": dict(scoring_mat_name="BLOSUM62", gap_open_cost=10),
    "cfg4": dict(match_score=2, mismatch_score=-3, gap_open_score=-10, gap_extension_score=-1),
    "cfg5": dict(match_score=2, mismatch_score=-3, gap_open_score=-10, gap_extension_score=-1),
}
