#!/usr/bin/env python3
"""ga_align_long on a DNA pair of [length] x [length] letters (default 20 kb; ncu target for the long-pair kernels)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from globalign_b200 import _native, synth  # noqa: E402
from globalign_b200.batch import Scheme  # noqa: E402
L = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
b = synth.dna_single(L, 4001, 4002) if L == 100000 else synth.dna_single(L, 7001, 7002)
ctx = _native.Context(0)
ctx.set_scheme(Scheme(**synth.SCHEMES["cfg4"]).dense_for(list("ACGT")))
for _ in range(2):
    cost, ops, status, info = ctx.align_long(b.s1, b.s2)
print(cost, len(ops), status, info)
