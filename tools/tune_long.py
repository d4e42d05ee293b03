#!/usr/bin/env python3
"""Sweep the long-pair fill over tuning builds (make -C globalign_b200/csrc variants: rows per
chain x chains per lane) and resident blocks per SM.  One subprocess per configuration (the
library is chosen at load time).  Usage: tune_long.py [length] [variants...]"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r"""
import json, os, sys
sys.path.insert(0, %r)
import numpy as np
from globalign_b200 import _native, synth
from globalign_b200.batch import Scheme
L = int(sys.argv[1])
b = synth.dna_single(L, 4001, 4002)
dense = Scheme(**synth.SCHEMES["cfg4"]).dense_for(list("ACGT"))
ctx = _native.Context(0)
ctx.set_scheme(dense)
best = None
for it in range(4):
    cost, ops, status, info = ctx.align_long(b.s1, b.s2)
    if it and (best is None or info["fill_ms"] < best["fill_ms"]):
        best = info
print(json.dumps(dict(cost=cost, n_ops=len(ops), status=status, fill_ms=best["fill_ms"], trace_ms=best["trace_ms"],
                      rows=best["tile_rows"], ckpt_mb=best["checkpoint_bytes"] / 1e6, mode=best["trace_mode"])))
""" % str(ROOT)


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    variants = sys.argv[2:] or sorted(p.name for p in (ROOT / "globalign_b200" / "_variants").glob("*.so"))
    ref = None
    for v in variants:
        for bps in (1, 2):
            env = dict(os.environ, GA_B200_LIB=str(ROOT / "globalign_b200" / "_variants" / v), GA_LONG_BLOCKS_PER_SM=str(bps))
            r = subprocess.run([sys.executable, "-c", CHILD, str(L)], env=env, capture_output=True, text=True, timeout=600)
            if r.returncode != 0:
                print(v, bps, "FAILED", r.stderr[-400:], flush=True)
                continue
            d = json.loads(r.stdout.strip().splitlines()[-1])
            if ref is None:
                ref = (d["cost"], d["n_ops"])
            ok = (d["cost"], d["n_ops"]) == ref and d["status"] == 0
            print(json.dumps(dict(variant=v, blocks_per_sm=bps, gcups_fill=L * L / d["fill_ms"] / 1e6, same=ok, **d)), flush=True)


if __name__ == "__main__":
    main()
