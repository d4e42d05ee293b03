#!/usr/bin/env python3
"""Fill + walk time per read length (fixed-length DNA batches, serial plan): which kernel the planner
picks for 76 / 100 / 125 / 150 / 200 / 250-letter reads and what it achieves."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from globalign_b200 import _native, synth
from globalign_b200.batch import Scheme
dense = Scheme(**synth.SCHEMES["cfg2"]).dense_for(list("ACGT"))
dev = torch.device("cuda", 0)
ctx = _native.Context(0); ctx.set_scheme(dense)
for L in (36, 50, 76, 100, 125, 150, 200, 250, 300, 400):
    n = int(9e9 / (L * L))
    b = synth.dna_fixed(n, L, 77)
    oo = _native.ops_offsets(b.off1, b.off2)
    d = {k: torch.from_numpy(v).to(dev) for k, v in dict(s1=b.s1, s2=b.s2, off1=b.off1, off2=b.off2, oo=oo).items()}
    cost = torch.empty(n, dtype=torch.int32, device=dev); ln = torch.empty_like(cost); st = torch.empty_like(cost)
    ops = torch.empty(int(oo[-1]) + 16, dtype=torch.uint8, device=dev)
    info = ctx.plan_batch(b.off1, b.off2, oo, flags=_native.GA_F_SERIAL)
    best = None
    for _ in range(3):
        ctx.run_batch(d["s1"].data_ptr(), d["off1"].data_ptr(), d["s2"].data_ptr(), d["off2"].data_ptr(),
                      cost.data_ptr(), ops.data_ptr(), d["oo"].data_ptr(), ln.data_ptr(), st.data_ptr())
        t = ctx.timing()
        best = t if best is None or t["fill"] + t["walk"] < best["fill"] + best["walk"] else best
    cells = b.cells()
    print(L, n, info["fill_kernel"], "fill_gcups", round(cells / best["fill"] / 1e6, 1), "total_gcups",
          round(cells / (best["fill"] + best["walk"]) / 1e6, 1), flush=True)
    del d, cost, ln, st, ops
    torch.cuda.empty_cache()
