#!/usr/bin/env python3
"""Exercise every kernel once on tiny inputs (target for compute-sanitizer)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from globalign_b200 import _native, synth  # noqa: E402
from globalign_b200.batch import Scheme  # noqa: E402
from globalign_b200.globaligner import find_global_alignment  # noqa: E402

r = find_global_alignment(seq_1="ACGTTGCA", seq_2="ACTTGGA")
print("single", r.cost, r.score)
for cfg in ["fill_p16_r19_g8", "fill_p16_r15_g10", "fill_p16_r10_g16", "fill_p16_r16_g32", "fill_i32s_r8_g16"]:
    os.environ["GA_FORCE_CFG"] = cfg
    ctx = _native.Context(0)
    b = synth.dna_fixed(37, 150 if "g32" not in cfg else 600, 2000)
    ctx.set_scheme(Scheme(**synth.SCHEMES["cfg2"]).dense_for(list("ACGT")))
    ctx.set_scratch_limit(4 << 20)
    out = ctx.align_batch(b.s1, b.off1, b.s2, b.off2)
    print(cfg, int(out[0].sum()), ctx.plan_info()["n_chunks"])
    ctx.close()
os.environ.pop("GA_FORCE_CFG")
ctx = _native.Context(0)
p = synth.protein_varlen(9, 3000)
ctx.set_scheme(Scheme(**synth.SCHEMES["cfg3"]).dense_for(sorted(synth.AMINO)))
out = ctx.align_batch(p.s1, p.off1, p.s2, p.off2)
print("protein", int(out[0].sum()))
cost, ops, status, info = ctx.align_long(p.s1[:700].copy(), p.s2[:650].copy())
print("long table", cost, len(ops), status)
d = synth.dna_single(1500, 1, 2)
ctx.set_scheme(Scheme(**synth.SCHEMES["cfg4"]).dense_for(list("ACGT")))
cost, ops, status, info = ctx.align_long(d.s1, d.s2)
print("long simple", cost, len(ops), status, info["retraced_tiles"])
t = np.zeros((8, 9, 3), dtype=np.int32)
ctx.fill_values(d.s1[:7].copy(), d.s2[:8].copy(), t, keep_boundary=False)
print("values", int(t[7, 8].min()))
