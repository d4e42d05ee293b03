#!/usr/bin/env python3
"""End-to-end time of the compact-output call under different chunk policies (GA_EDGE_LAST, GA_EDGE_CHUNK, GA_CHUNKS).
Usage: probe_rle_chunks.py <cfg2|cfg3|cfg5> [pairs]"""
import os
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from globalign_b200 import _native, synth  # noqa: E402
from globalign_b200.batch import Scheme  # noqa: E402

name = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else {"cfg2": 1_000_000, "cfg3": 100_000, "cfg5": 1_000_000}[name]
b = synth.workload(name, n_pairs=n)
dense = Scheme(**synth.SCHEMES[name]).dense_for(sorted(b.alphabet))
remap = np.array([dense.index[ch] for ch in b.alphabet], dtype=np.uint8)
bits = 2 if dense.A <= 4 else 8
lens1, lens2 = np.diff(b.off1), np.diff(b.off2)
fixed = bool((lens1 == lens1[0]).all() and (lens2 == lens2[0]).all())


def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


keep = []
hb = []
for a in (_native.pack_letters(remap[b.s1], bits), b.off1, _native.pack_letters(remap[b.s2], bits), b.off2):
    t, v = pin(a)
    keep.append(t)
    hb.append(v)
cap = int(min((lens1 + lens2).sum(), 48 * n + (1 << 20)))
out = []
for shape, dt in ((n, _native.PAIR_RESULT), (cap, np.uint8), (n + 1, np.int64)):
    t = torch.empty(shape * np.dtype(dt).itemsize, dtype=torch.uint8).pin_memory()
    keep.append(t)
    out.append(t.numpy().view(dt))
ctx = _native.Context(0)
ctx.set_scheme(dense)
fl = _native.seq_flag(bits)
for env in ({}, {"GA_EDGE_LAST": "1"}, {"GA_EDGE_CHUNK": "0.06"}, {"GA_EDGE_CHUNK": "0.06", "GA_EDGE_LAST": "1"}, {"GA_EDGE_CHUNK": "0.2"}):
    for k in ("GA_EDGE_LAST", "GA_EDGE_CHUNK", "GA_CHUNKS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    call = (lambda: ctx.align_batch_rle(hb[0], None, hb[2], None, int(lens1[0]), int(lens2[0]), n, flags=fl, out=tuple(out))) if fixed \
        else (lambda: ctx.align_batch_rle(hb[0], hb[1], hb[2], hb[3], flags=fl, out=tuple(out)))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        call()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 8
    print(name, env, "ms", round(dt * 1e3, 3), "gcups", round(b.cells() / dt / 1e9, 1), "chunks", ctx.plan_info()["n_chunks"],
          {k: round(v, 2) for k, v in ctx.timing().items()}, flush=True)
