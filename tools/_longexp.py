import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from globalign_b200 import _native, synth
from globalign_b200.batch import Scheme
b = synth.workload("cfg4")
dense = Scheme(**synth.SCHEMES["cfg4"]).dense_for(list("ACGT"))
ctx = _native.Context(0); ctx.set_scheme(dense)
for n in (100000, 50000, 25000):
    for rep in range(2):
        c, _, _, info = ctx.align_long(b.s1[:n].copy(), b.s2[:n].copy(), flags=_native.GA_F_COST_ONLY)
    print(os.environ.get("GA_DEBUG_LONG_NOSYNC"), n, "fill_ms", round(info["fill_ms"], 3), "cost", c, flush=True)
