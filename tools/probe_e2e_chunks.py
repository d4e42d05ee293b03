#!/usr/bin/env python3
"""cfg2 through the fixed-length host-buffer call (2-bit packed, pinned): wall time per batch for
different first/last chunk shares (GA_EDGE_CHUNK) and chunk counts (GA_CHUNKS)."""
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
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
p1, p2 = _native.pack_letters(b.s1, 2), _native.pack_letters(b.s2, 2)
keep = [pin(p1), pin(p2)]
h1, h2 = keep[0].numpy(), keep[1].numpy()
stride = (((150 + 150 + 3) // 4) + 3) // 4 * 4
outs_t = [torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(stride * n, dtype=torch.uint8).pin_memory(),
          torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()]
outs = tuple(t.numpy() for t in outs_t)
for edge, chunks in ((None, None), ("0.0625", None), ("0.04", None), ("0.0625", "8"), ("0.03", "10"), (None, None)):
    for k, v in (("GA_EDGE_CHUNK", edge), ("GA_CHUNKS", chunks)):
        os.environ.pop(k, None)
        if v:
            os.environ[k] = v
    for _ in range(3):
        ctx.align_batch_uniform(h1, 150, h2, 150, n, flags=_native.GA_F_SEQ_2BIT, out=outs)
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.align_batch_uniform(h1, 150, h2, 150, n, flags=_native.GA_F_SEQ_2BIT, out=outs)
    dt = (time.perf_counter() - t0) / 10
    print("edge", edge, "chunks", chunks, "n_chunks", ctx.plan_info()["n_chunks"], "ms/batch", round(dt * 1e3, 3), flush=True)

os.environ.pop("GA_EDGE_CHUNK", None); os.environ.pop("GA_CHUNKS", None)
os.environ["GA_DEBUG_TIMELINE"] = "1"
ctx.align_batch_uniform(h1, 150, h2, 150, n, flags=_native.GA_F_SEQ_2BIT, out=outs)
ctx.timing()
