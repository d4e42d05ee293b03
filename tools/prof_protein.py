#!/usr/bin/env python3
"""cfg3 sample (protein, BLOSUM62) through the serial plan: ncu target + timing of the table kernels."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from globalign_b200 import _native, synth
from globalign_b200.batch import Scheme
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
b = synth.workload("cfg3", n_pairs=n)
dense = Scheme(**synth.SCHEMES["cfg3"]).dense_for(sorted(synth.AMINO))
if os.environ.get("PROT_LETTERS") == "20":      # tables over the 20 letters that occur, not the matrix's 24
    from globalign_b200 import lowering
    sch = Scheme(**synth.SCHEMES["cfg3"])
    dense = lowering.DenseScheme(sch.costing_mat, sch.gap_open_cost, letters=sorted(synth.AMINO))
    dense.max_cost = int(max(max(row.values()) for row in sch.costing_mat.values()))
remap = __import__("numpy").array([dense.index[ch] for ch in b.alphabet], dtype="uint8")
s1, s2 = remap[b.s1], remap[b.s2]
dev = torch.device("cuda", 0)
oo = _native.ops_offsets(b.off1, b.off2)
d = {k: torch.from_numpy(v).to(dev) for k, v in dict(s1=s1, s2=s2, off1=b.off1, off2=b.off2, oo=oo).items()}
cost = torch.empty(n, dtype=torch.int32, device=dev); ln = torch.empty_like(cost); st = torch.empty_like(cost)
ops = torch.empty(int(oo[-1]) + 16, dtype=torch.uint8, device=dev)
for cfg in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["fill_i32t_r8_g32", "fill_i32t_r8_g16", "fill_i32t_r8_g8"]):
    os.environ["GA_FORCE_CFG"] = cfg
    ctx = _native.Context(0); ctx.set_scheme(dense)
    info = ctx.plan_batch(b.off1, b.off2, oo, flags=_native.GA_F_SERIAL)
    best = None
    for _ in range(3):
        ctx.run_batch(d["s1"].data_ptr(), d["off1"].data_ptr(), d["s2"].data_ptr(), d["off2"].data_ptr(),
                      cost.data_ptr(), ops.data_ptr(), d["oo"].data_ptr(), ln.data_ptr(), st.data_ptr())
        t = ctx.timing()
        best = t if best is None or t["fill"] < best["fill"] else best
    print(cfg, "fill_ms", round(best["fill"], 3), "walk_ms", round(best["walk"], 3), "fill_gcups", round(b.cells() / best["fill"] / 1e6, 1), "sum", int(cost.sum()), flush=True)
    ctx.close()
