#!/usr/bin/env python3
"""cfg3 (ragged protein batch): pipeline chunk count vs step time on device-resident buffers, and
host phase times of the host-buffer call (GA_DEBUG_TIMING=1)."""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from globalign_b200 import _native, synth
from globalign_b200.batch import Scheme
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
b = synth.workload("cfg3", n_pairs=n)
dense = Scheme(**synth.SCHEMES["cfg3"]).dense_for(sorted(synth.AMINO))
remap = np.array([dense.index[ch] for ch in b.alphabet], dtype=np.uint8)
s1, s2 = remap[b.s1], remap[b.s2]
dev = torch.device("cuda", 0)
oo = _native.ops_offsets(b.off1, b.off2)
d = {k: torch.from_numpy(v).to(dev) for k, v in dict(s1=s1, s2=s2, off1=b.off1, off2=b.off2, oo=oo).items()}
cost = torch.empty(n, dtype=torch.int32, device=dev); ln = torch.empty_like(cost); st = torch.empty_like(cost)
ops = torch.empty(int(oo[-1]) + 16, dtype=torch.uint8, device=dev)
ctx = _native.Context(0); ctx.set_scheme(dense)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
for chunks, flags in (("1", 0), ("2", 0), ("3", 0), ("4", 0), ("8", 0), ("serial", _native.GA_F_SERIAL)):
    if chunks != "serial":
        os.environ["GA_CHUNKS"] = chunks
    info = ctx.plan_batch(b.off1, b.off2, oo, flags=flags)
    run = lambda: ctx.run_batch(d["s1"].data_ptr(), d["off1"].data_ptr(), d["s2"].data_ptr(), d["off2"].data_ptr(),
                                cost.data_ptr(), ops.data_ptr(), d["oo"].data_ptr(), ln.data_ptr(), st.data_ptr())
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    print("GA_CHUNKS", chunks, "n_chunks", info["n_chunks"], "ms/step", round(e0.elapsed_time(e1) / 5, 3), ctx.timing(), flush=True)
os.environ.pop("GA_CHUNKS", None)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
keep = [pin(x) for x in (s1, b.off1, s2, b.off2)]
hb = [t.numpy() for t in keep]
outs_t = [torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(int(oo[-1]), dtype=torch.uint8).pin_memory(),
          pin(oo), torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()]
outs = tuple(t.numpy() for t in outs_t)
for it in range(4):
    os.environ.pop("GA_DEBUG_TIMING", None)
    if it == 3:
        os.environ["GA_DEBUG_TIMING"] = "1"
    t0 = time.perf_counter()
    ctx.align_batch(hb[0], hb[1], hb[2], hb[3], out=outs)
    print("host-buffer call wall ms", round((time.perf_counter() - t0) * 1e3, 3), ctx.timing(), flush=True)
