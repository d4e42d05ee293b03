#!/usr/bin/env python3
"""cfg2: fill residency (blocks per SM) and chunk count vs step time on device-resident buffers."""
import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from globalign_b200 import _native, synth
from globalign_b200.batch import Scheme
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
b = synth.workload("cfg2", n_pairs=n)
dense = Scheme(**synth.SCHEMES["cfg2"]).dense_for(list("ACGT"))
dev = torch.device("cuda", 0)
oo = _native.ops_offsets(b.off1, b.off2)
d = {k: torch.from_numpy(v).to(dev) for k, v in dict(s1=b.s1, s2=b.s2, off1=b.off1, off2=b.off2, oo=oo).items()}
cost = torch.empty(n, dtype=torch.int32, device=dev); ln = torch.empty_like(cost); st = torch.empty_like(cost)
ops = torch.empty(int(oo[-1]) + 16, dtype=torch.uint8, device=dev)
ctx = _native.Context(0); ctx.set_scheme(dense)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
import itertools
LAY = os.environ.get("PROBE_LAYOUTS")
combos = [(l, "4") for l in LAY.split(",")] * 3 if LAY else [("1", x) for x in ("4", "3", "2")]
SPLIT = os.environ.get("PROBE_SPLIT")
if SPLIT:
    combos = [("1", "4")] * 2
for lay, per_sm in combos:
    os.environ["GA_CODE_PAIRS"] = lay
    os.environ["GA_FILL_BLOCKS_PER_SM"] = per_sm
    for chunks, flags in ([("4", 0), ("4s", 0), ("8", 0), ("8s", 0), ("16", 0), ("16s", 0)] if SPLIT else (("serial", _native.GA_F_SERIAL), ("4", 0)) if LAY else (("serial", _native.GA_F_SERIAL), ("2", 0), ("4", 0), ("8", 0), ("16", 0))):
        os.environ["GA_SPLIT_FILL"] = "1" if chunks.endswith("s") else "0"
        if chunks != "serial":
            os.environ["GA_CHUNKS"] = chunks.rstrip("s")
        info = ctx.plan_batch(b.off1, b.off2, oo, flags=flags)
        run = lambda: ctx.run_batch(d["s1"].data_ptr(), d["off1"].data_ptr(), d["s2"].data_ptr(), d["off2"].data_ptr(),
                                    cost.data_ptr(), ops.data_ptr(), d["oo"].data_ptr(), ln.data_ptr(), st.data_ptr())
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            run()
        e1.record(stream)
        torch.cuda.synchronize()
        t = ctx.timing()
        print("code_pairs", lay, "blocks/SM", per_sm, "chunks", chunks, info["n_chunks"], "ms/step", round(e0.elapsed_time(e1) / 10, 3),
              "fill", round(t["fill"], 2), "walk", round(t["walk"], 2), flush=True)
