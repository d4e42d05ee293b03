#!/usr/bin/env python3
"""Where does the long-pair fill spend its time?  For each tuning build: cost-only fills of
(1, 2, 8, 64 bands) x 100k columns and of the full 100 kb x 100 kb pair.  One band gives the
step time of a lone warp (no hand-off); more bands add the band pipeline.
Usage: diag_long.py [variants...]   (default: all builds under globalign_b200/_variants + the in-tree one)"""
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
b = synth.dna_single(100_000, 4001, 4002)
dense = Scheme(**synth.SCHEMES["cfg4"]).dense_for(list("ACGT"))
ctx = _native.Context(0)
ctx.set_scheme(dense)
_, _, _, info = ctx.align_long(b.s1[:64], b.s2[:64])
rows = info["tile_rows"]
out = {"rows": rows}
for bands in (1, 2, 8, 64, 0):
    m = bands * rows if bands else 100_000
    best = 1e9
    for it in range(3):
        cost, _, _, info = ctx.align_long(b.s1[:m], b.s2, flags=_native.GA_F_COST_ONLY)
        best = min(best, info["fill_ms"])
    out[str(bands or "full")] = round(best, 3)
print(json.dumps(out))
""" % str(ROOT)


def main():
    vdir = ROOT / "globalign_b200" / "_variants"
    variants = sys.argv[1:] or ([""] + sorted(p.name for p in vdir.glob("*.so")))
    for v in variants:
        for bps in [int(x) for x in os.environ.get("DIAG_BPS", "1,2").split(",")]:
            env = dict(os.environ, GA_LONG_BLOCKS_PER_SM=str(bps))
            if v:
                env["GA_B200_LIB"] = str(vdir / v)
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
            if r.returncode != 0:
                print(v, bps, "FAILED", r.stderr[-400:], flush=True)
                continue
            d = json.loads(r.stdout.strip().splitlines()[-1])
            d.update(variant=v or "in-tree", blocks_per_sm=bps,
                     cycles_per_step_one_band=round(d["1"] * 1.965e6 / 100_000, 1),
                     gcups_full=round(1e10 / d["full"] / 1e6, 1))
            print(json.dumps(d), flush=True)


if __name__ == "__main__":
    main()
