#!/usr/bin/env python3
"""Small single pairs through ga_align_long: band-parallel vs sequential traceback (GA_LONG_SEQ_BANDS)."""
import json
import os
import subprocess
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
CHILD = r"""
import json, sys
sys.path.insert(0, %r)
from globalign_b200 import _native, synth
from globalign_b200.batch import Scheme
ctx = _native.Context(0)
ctx.set_scheme(Scheme(**synth.SCHEMES["cfg1"]).dense_for(list("ACGT")))
out = {}
for L in (500, 1000, 2000, 4000, 8000, 16000):
    b = synth.dna_single(L, 1001, 1002)
    best = None
    for it in range(6):
        cost, ops, status, info = ctx.align_long(b.s1, b.s2)
        t = info["fill_ms"] + info["trace_ms"]
        if it and (best is None or t < best[0]):
            best = (t, info["fill_ms"], info["trace_ms"], info["trace_mode"], cost, len(ops))
    out[L] = best
print(json.dumps(out))
""" % str(ROOT)
for sb in (0, 1000000):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, GA_LONG_SEQ_BANDS=str(sb)), capture_output=True, text=True)
    print("GA_LONG_SEQ_BANDS", sb, r.stdout.strip() or r.stderr[-500:], flush=True)
