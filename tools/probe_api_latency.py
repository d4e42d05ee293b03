#!/usr/bin/env python3
"""Wall time of the reference-shaped call find_global_alignment(seq_1=..., seq_2=...) per pair size."""
import sys, time, random
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from globalign_b200.globaligner import find_global_alignment
rng = random.Random(1)
for L in (20, 150, 500, 1000, 3000):
    a = "".join(rng.choice("ACGT") for _ in range(L))
    b = "".join(ch if rng.random() < 0.9 else rng.choice("ACGT") for ch in a)
    for _ in range(3):
        find_global_alignment(seq_1=a, seq_2=b)
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        r = find_global_alignment(seq_1=a, seq_2=b)
    dt = (time.perf_counter() - t0) / n * 1e3
    print(L, "ms per call", round(dt, 3), "cost", r.cost, flush=True)
import cProfile, pstats
a = "".join(rng.choice("ACGT") for _ in range(150)); b = a[:70] + a[75:]
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    find_global_alignment(seq_1=a, seq_2=b)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
