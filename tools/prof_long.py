#!/usr/bin/env python3
"""One ga_align_long call on a 20 kb x 20 kb DNA pair (ncu target for the long-pair kernels)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from globalign_b200 import _native, synth  # noqa: E402
from globalign_b200.batch import Scheme  # noqa: E402
b = synth.dna_single(20000, 7001, 7002)
ctx = _native.Context(0)
ctx.set_scheme(Scheme(**synth.SCHEMES["cfg4"]).dense_for(list("ACGT")))
for _ in range(2):
    cost, ops, status, info = ctx.align_long(b.s1, b.s2)
print(cost, len(ops), status, info)
