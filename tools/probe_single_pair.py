#!/usr/bin/env python3
"""Latency of ONE pair through the batched entry point (one warp fills it) and through
ga_align_long (bands across the chip): where should find_global_alignment switch?"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from globalign_b200 import _native, synth
from globalign_b200.batch import Scheme
ctx = _native.Context(0)
ctx.set_scheme(Scheme(**synth.SCHEMES["cfg2"]).dense_for(list("ACGT")))
for L in (100, 200, 300, 500, 700, 1000, 1500, 2000):
    b = synth.dna_single(L, 1, 2)
    off = np.array([0, L], dtype=np.int64)
    res = {}
    for name, fn in (("batch", lambda: ctx.align_batch(b.s1, off, b.s2, off)), ("long", lambda: ctx.align_long(b.s1, b.s2))):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(20):
            fn()
        res[name] = (time.perf_counter() - t0) / 20 * 1e3
    print(L, {k: round(v, 3) for k, v in res.items()}, flush=True)
