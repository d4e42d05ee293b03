#!/usr/bin/env python3
"""find_global_alignments(list[str], list[str]) on 1M cfg2 read pairs: time per call and where it goes."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
from globalign_b200 import batch as B, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
b = synth.workload("cfg2", n_pairs=n)
lut = np.frombuffer(b.alphabet.encode("ascii"), dtype=np.uint8)
m, nn = int(b.off1[1]), int(b.off2[1])
t1 = lut[b.s1].tobytes().decode("ascii")
t2 = lut[b.s2].tobytes().decode("ascii")
s1 = [t1[k * m:(k + 1) * m] for k in range(n)]
s2 = [t2[k * nn:(k + 1) * nn] for k in range(n)]
B.find_global_alignments(s1[:1000], s2[:1000], **synth.SCHEMES["cfg2"])
for _ in range(4):
    t0 = time.perf_counter()
    res = B.find_global_alignments(s1, s2, **synth.SCHEMES["cfg2"])
    dt = time.perf_counter() - t0
    print("call %.1f ms  %.0f GCUPS  cost sum %d" % (dt * 1e3, n * m * nn / dt / 1e9, int(res.cost.sum())), flush=True)
cProfile.run('B.find_global_alignments(s1, s2, **synth.SCHEMES["cfg2"])', "/tmp/sprof")
pstats.Stats("/tmp/sprof").sort_stats("cumtime").print_stats(22)
