#!/usr/bin/env python3
"""End-to-end probe: ga_align_batch on pinned host buffers, host phase times (GA_DEBUG_TIMING=1)."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from globalign_b200 import _native, synth
from globalign_b200.batch import Scheme
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
b = synth.workload("cfg2", n_pairs=n)
ctx = _native.Context(0)
ctx.set_scheme(Scheme(**synth.SCHEMES["cfg2"]).dense_for(list("ACGT")))
ops_off = _native.ops_offsets(b.off1, b.off2)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
keep = [pin(x) for x in (b.s1, b.off1, b.s2, b.off2)]
hb = [t.numpy() for t in keep]
outs_t = [torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(int(ops_off[-1]), dtype=torch.uint8).pin_memory(),
          pin(ops_off), torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()]
outs = tuple(t.numpy() for t in outs_t)
for it in range(4):
    os.environ.pop("GA_DEBUG_TIMING", None)
    if it == 3:
        os.environ["GA_DEBUG_TIMING"] = "1"
    t0 = time.perf_counter()
    ctx.align_batch(hb[0], hb[1], hb[2], hb[3], out=outs)
    print("wall ms", round((time.perf_counter() - t0) * 1e3, 3), ctx.timing(), flush=True)
