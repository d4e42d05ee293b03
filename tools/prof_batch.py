#!/usr/bin/env python3
"""ncu / timing target for the batched kernels: a sample of a BASELINE workload through the serial
plan (one fill launch + one walk launch per run), three runs.
Usage: prof_batch.py <cfg2|cfg3|cfg5> <pairs> [GA_FORCE_CFG name]"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from globalign_b200 import _native, synth  # noqa: E402
from globalign_b200.batch import Scheme  # noqa: E402

name, n = sys.argv[1], int(sys.argv[2])
if len(sys.argv) > 3:
    os.environ["GA_FORCE_CFG"] = sys.argv[3]
b = synth.workload(name, n_pairs=n)
scheme = Scheme(**synth.SCHEMES[name])
dense = scheme.dense_for(sorted(b.alphabet))
remap = np.array([dense.index[ch] for ch in b.alphabet], dtype=np.uint8)
dev = torch.device("cuda", 0)
oo = _native.ops_offsets(b.off1, b.off2)
d = {k: torch.from_numpy(v).to(dev) for k, v in dict(s1=remap[b.s1], s2=remap[b.s2], off1=b.off1, off2=b.off2, oo=oo).items()}
cost = torch.empty(n, dtype=torch.int32, device=dev)
ln, st = torch.empty_like(cost), torch.empty_like(cost)
ops = torch.empty(int(oo[-1]) + 16, dtype=torch.uint8, device=dev)
ctx = _native.Context(0)
ctx.set_scheme(dense)
info = ctx.plan_batch(b.off1, b.off2, oo, flags=_native.GA_F_SERIAL)
for _ in range(3):
    ctx.run_batch(d["s1"].data_ptr(), d["off1"].data_ptr(), d["s2"].data_ptr(), d["off2"].data_ptr(),
                  cost.data_ptr(), ops.data_ptr(), d["oo"].data_ptr(), ln.data_ptr(), st.data_ptr())
    t = ctx.timing()
print(name, n, info["fill_kernel"], "chunks", info["n_chunks"], "cells", b.cells(), "code_bytes", info["code_bytes"],
      "fill_ms", round(t["fill"], 3), "walk_ms", round(t["walk"], 3), "fill_gcups", round(b.cells() / t["fill"] / 1e6, 1),
      "sum", int(cost.sum()))
