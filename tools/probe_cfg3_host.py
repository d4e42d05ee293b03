#!/usr/bin/env python3
"""Host-side phases of ga_align_batch_rle on the cfg3 workload (GA_DEBUG_TIMING=1 prints them)."""
import os
import sys
import time
from pathlib import Path

os.environ["GA_DEBUG_TIMING"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
from globalign_b200 import _native, synth  # noqa: E402
from globalign_b200.batch import Scheme  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
b = synth.workload(name, n_pairs=100000 if name == "cfg3" else 1000000)
scheme = Scheme(**synth.SCHEMES[name])
dense = scheme.dense_for(sorted(b.alphabet))
remap = np.array([dense.index[ch] for ch in b.alphabet], dtype=np.uint8)
s1, s2 = remap[b.s1], remap[b.s2]
ctx = _native.Context(0)
ctx.set_scheme(dense)
for it in range(4):
    t0 = time.perf_counter()
    res, rle, off = ctx.align_batch_rle(s1, b.off1, s2, b.off2)
    print("call %d: %.2f ms" % (it, (time.perf_counter() - t0) * 1e3), flush=True)
