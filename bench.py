#!/usr/bin/env python3
"""Benchmark of the hot path: GCUPS with traceback (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

A "step" is one pass of the hot path (fill + direction-code store + traceback walk)
over one batch of synthetic pairs of the BASELINE shape.  Default workload:
BASELINE.json configs[1] (1,000,000 DNA pairs, 150 x 150) per GPU; with N > 1 every
rank aligns its own shard of N x 1,000,000 pairs (pairs are independent: no
collective on the data path, "scaling": "weak").

`value`   device-resident inputs -> device-resident cost + packed ops, CUDA events on the
          launching stream, barrier + synchronize on both sides, max over ranks.
`e2e`     the same work through the public host-buffer call (ga_align_batch): pinned host
          inputs, H2D, planning, kernels, D2H of costs/ops/lengths/status, all timed.
`roofline`     fill kernel vs the integer-ALU roofline (BASELINE metric), from its CUDA-event
               duration (recorded by the library on the launching stream).
`roofline_hbm` direction-code store vs the measured HBM peak (MEASURED_PEAKS.json).
`cpu_baseline` the C oracle port on the host cores, on a bounded sample (rank 0, N = 1).
`--impl reference` times that CPU port alone (the reference is pure Python and cannot travel
          to the GPU box; its restatement can), with all host threads.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from globalign_b200 import synth  # noqa: E402

METRIC = "GCUPS w/ traceback"
UNIT = "GCUPS"
OPS_PER_CELL = 24          # SURVEY.md 8d contract figure: int32 ops per cell, fill with codes
CODE_BYTES_PER_CELL = 0.5   # 2 bits (level-0 argmin) + 1 + 1 (gap-level take bits), DESIGN.md
NOMINAL_INT32_TOPS = 148 * 64 * 1.965e9 / 1e12     # 18.6 Tops/s

DESCR = {
    "cfg1": "single DNA pair 1 kb x 1 kb, match 2 / mismatch -3 / gap open -10 / extend -1 (BASELINE configs[0])",
    "cfg4": "single long DNA pair 100 kb x 100 kb, band-pipeline fill + checkpointed linear-space traceback "
            "(BASELINE configs[3])",
    "cfg2": "1M synthetic DNA short-read pairs 150 bp x 150 bp with traceback (BASELINE configs[1])",
    "cfg3": "100k protein pairs 300-1000 aa, BLOSUM62, gap_open_cost=10 (BASELINE configs[2])",
    "cfg5": "10M DNA pairs 250 bp x 250 bp sharded across the GPUs (BASELINE configs[4])",
}
DEFAULT_PAIRS = {"cfg1": 1, "cfg4": 1, "cfg2": 1_000_000, "cfg3": 100_000, "cfg5": 10_000_000}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(DESCR))  # cfg1 cfg2 cfg3 cfg4 cfg5
    ap.add_argument("--pairs", type=int, default=None, help="pairs per GPU (default: the BASELINE size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic_r1.json)."""
    p = ROOT / "profiles" / "traffic_r1.json"
    try:
        e = json.loads(p.read_text()).get(kernel)
    except (OSError, ValueError):
        e = None
    return e


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_lo: float, t_hi: float) -> dict:
        """Median SM clock / active throttle reasons of the samples taken in [t_lo, t_hi]
        (the timed region); if the region was shorter than the sampling period, of the
        samples taken while the GPU was busy just before it (warm-up)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        good = [(t, r) for t, r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        win = [r for t, r in good if t_lo <= t <= t_hi + 0.05]
        scope = "timed region"
        if not win:
            win = [r for t, r in good if t_lo - 1.0 <= t <= t_hi + 0.2]
            scope = "timed region +-1 s"
        sm = [float(r[1]) for r in win]
        mx = [float(r[2]) for r in win]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in win for k in range(4) if r[4 + k].lower().startswith("active")})
        pw = [float(r[3]) for r in win if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "scope": scope,
                "reasons": reasons}


def oracle_dense(dense):
    from oracle import c_oracle
    d = c_oracle.Dense.__new__(c_oracle.Dense)
    d.A, d.sub, d.gap_h, d.gap_v, d.o, d.max_cost = (dense.A, dense.sub, dense.gap_h, dense.gap_v,
                                                     dense.gap_open, dense.max_cost)
    return d


def cpu_port_gcups(batch, dense, remap, target_s: float = 12.0, threads: int | None = None):
    """C oracle (port of the reference's path) on the host cores, bounded sample."""
    from oracle import c_oracle
    threads = threads or os.cpu_count() or 1
    d = oracle_dense(dense)
    n_probe = min(batch.n_pairs, 64 * threads)
    sub = batch.head(n_probe)
    t0 = time.perf_counter()
    c_oracle.align_batch(remap[sub.s1], sub.off1, remap[sub.s2], sub.off2, d, threads)
    dt = time.perf_counter() - t0
    rate = sub.cells() / max(dt, 1e-9)
    n = int(min(batch.n_pairs, max(n_probe, target_s * rate / (batch.cells() / batch.n_pairs))))
    sub = batch.head(n)
    t0 = time.perf_counter()
    c_oracle.align_batch(remap[sub.s1], sub.off1, remap[sub.s2], sub.off2, d, threads)
    dt = time.perf_counter() - t0
    return sub.cells() / dt / 1e9, n, dt, threads


def python_port_mcups(batch, scheme, n: int = 6):
    """The pure-Python restatement (same data structures as the reference), 1 core."""
    from oracle import py_oracle
    t0 = time.perf_counter()
    cells = 0
    for k in range(min(n, batch.n_pairs)):
        s1, s2 = batch.pair(k)
        sm, cm = scheme.matrices_for(s1, s2)
        py_oracle.align(s1, s2, cm, scheme.gap_open_cost)
        cells += len(s1) * len(s2)
    return cells / (time.perf_counter() - t0) / 1e6


def run_reference_long(args, name, scheme, threads):
    """cfg1: the C port on the pair itself (1 core: one pair does not parallelise in the
    reference's algorithm).  cfg4: a bounded 8 kb x 8 kb prefix of the pair (the full table the
    reference's algorithm needs, 12 B x 1e10 cells, does not fit in host memory)."""
    from oracle import c_oracle
    batch = synth.workload(name)
    k = 1000 if name == "cfg1" else 8000
    dense = scheme.dense_for(sorted(batch.alphabet))
    remap = np.array([dense.index[ch] for ch in batch.alphabet], dtype=np.uint8)
    s1, s2 = remap[batch.s1[:k]], remap[batch.s2[:k]]
    d = oracle_dense(dense)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        c_oracle.align_codes(s1, s2, d)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = k * k / (ms / 1e3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": DESCR[name], "sample": f"{k} x {k} prefix of the pair"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{k} x {k} prefix, C restatement (oracle/ga_oracle.c), one thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    if rank != 0:
        return
    from globalign_b200.batch import Scheme
    name = args.workload
    scheme = Scheme(**synth.SCHEMES[name])
    threads = os.cpu_count() or 1
    if name in ("cfg1", "cfg4"):
        return run_reference_long(args, name, scheme, threads)
    per_pair = {"cfg2": 22500, "cfg3": 450000, "cfg5": 62500}[name]
    n_sample = int(min(DEFAULT_PAIRS[name], max(256, 4e9 / per_pair)))       # ~4e9 cells per step
    batch = synth.workload(name, n_pairs=n_sample)
    dense = scheme.dense_for(sorted(batch.alphabet))
    remap = np.array([dense.index[ch] for ch in batch.alphabet], dtype=np.uint8)
    from oracle import c_oracle
    d = oracle_dense(dense)
    s1, s2 = remap[batch.s1], remap[batch.s2]
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        c_oracle.align_batch(s1, batch.off1, s2, batch.off2, d, threads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = batch.cells() / (ms / 1e3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": DESCR[name], "sample": f"first {n_sample} pairs per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"first {n_sample} pairs of {name}, C restatement of the reference's "
                                   f"fill + value-comparing traceback (oracle/ga_oracle.c), pthreads"},
        "python_port_mcups_per_core": python_port_mcups(batch, scheme),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- single long pair
def run_long(args, rank, world, local_rank):
    """cfg1 / cfg4: one pair per GPU through ga_align_long (every rank aligns the same pair:
    a single long pair stays on one GPU; N > 1 runs are replicas)."""
    import torch
    import torch.distributed as dist

    from globalign_b200 import _native
    from globalign_b200.batch import Scheme

    torch.cuda.set_device(local_rank)
    name = args.workload
    batch = synth.workload(name)
    scheme = Scheme(**synth.SCHEMES[name])
    dense = scheme.dense_for(sorted(batch.alphabet))
    ctx = _native.Context(local_rank)
    ctx.set_scheme(dense)
    cells = batch.cells()
    steps = max(1, min(args.steps, 10))
    for _ in range(max(1, min(args.warmup, 3))):
        ctx.align_long(batch.s1, batch.s2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    t_lo = time.perf_counter()
    dev_ms, fill_ms, trace_ms, info = 0.0, 0.0, 0.0, None
    for _ in range(steps):
        cost, ops, status, info = ctx.align_long(batch.s1, batch.s2)
        dev_ms += info["fill_ms"] + info["trace_ms"]
        fill_ms += info["fill_ms"]
        trace_ms += info["trace_ms"]
    torch.cuda.synchronize()
    t_hi = time.perf_counter()
    clocks = sampler.stop(t_lo, t_hi)
    vals = torch.tensor([dev_ms / steps, (t_hi - t_lo) * 1e3 / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    if rank != 0:
        return
    step_ms, e2e_ms = vals[0].item(), vals[1].item()
    total_cells = cells * world
    fill_gcups = cells / (fill_ms / steps / 1e3) / 1e9
    achieved = fill_gcups * 1e9 * OPS_PER_CELL / 1e12
    line = {
        "metric": METRIC, "value": total_cells / (step_ms / 1e3) / 1e9, "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": DESCR[name], "cells_per_gpu": cells, "replicas": world,
                   "l2": "checkpoints %.0f MB per step" % (info["checkpoint_bytes"] / 1e6),
                   "tile": [info["tile_rows"], info["tile_cols"]], "retraced_tiles": info["retraced_tiles"]},
        "clocks": clocks,
        "e2e": {"value": total_cells / (e2e_ms / 1e3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(batch.s1.nbytes + batch.s2.nbytes + 4 * (len(batch.s1) + len(batch.s2) + 2)),
                "d2h_bytes_per_step": int(len(ops) // 4 + 32)},
        "gpu_launches": info["launches"] * steps,
        "kernels_ms": {"fill": fill_ms / steps, "trace": trace_ms / steps},
        "roofline": {"bound": "int_alu", "kernel": "long_fill_kernel", "achieved": achieved,
                     "peak": NOMINAL_INT32_TOPS, "unit": "Tops/s", "frac": achieved / NOMINAL_INT32_TOPS,
                     "peak_source": "nominal 148 SM x 64 lanes x 1.965 GHz", "ops_per_cell": OPS_PER_CELL,
                     "fill_gcups": fill_gcups, "traffic": None},
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa(local_rank: int):
    """Pin this rank's host thread to the CPU cores NVML reports as local to its GPU, BEFORE any
    pinned buffer is allocated (first touch then places the pages on that NUMA node).  With eight
    ranks the end-to-end copies otherwise cross sockets.  Returns the core count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from globalign_b200 import _native
    from globalign_b200.batch import Scheme

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: globalign_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cores = bind_to_gpu_numa(local_rank) if world > 1 else None
    name = args.workload
    n_total = args.pairs * world if args.pairs else DEFAULT_PAIRS[name] * (world if name != "cfg5" else 1)
    scaling = "strong" if (name == "cfg5" and not args.pairs) else "weak"
    per_rank = n_total // world
    blocks_per_rank = -(-per_rank // synth._BLOCK)
    if name == "cfg3":
        batch = synth.protein_varlen(per_rank * world, 3000).slice(rank * per_rank, (rank + 1) * per_rank) \
            if world > 1 else synth.workload(name, n_pairs=per_rank)
    else:
        batch = synth.workload(name, n_pairs=per_rank, first_block=rank * blocks_per_rank)
    scheme = Scheme(**synth.SCHEMES[name])
    dense = scheme.dense_for(sorted(batch.alphabet))
    remap = np.array([dense.index[ch] for ch in batch.alphabet], dtype=np.uint8)
    s1h = remap[batch.s1]
    s2h = remap[batch.s2]
    cells = batch.cells()

    ctx = _native.Context(local_rank)
    ctx.set_scheme(dense)
    # an explicit (non-default) torch stream: the library launches on it and the
    # torch.cuda.Event pair below brackets exactly those launches
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ops_off_h = _native.ops_offsets(batch.off1, batch.off2)

    # ---- device-resident buffers
    d_s1 = torch.from_numpy(s1h).to(dev)
    d_s2 = torch.from_numpy(s2h).to(dev)
    d_off1 = torch.from_numpy(batch.off1).to(dev)
    d_off2 = torch.from_numpy(batch.off2).to(dev)
    d_opsoff = torch.from_numpy(ops_off_h).to(dev)
    n = batch.n_pairs
    d_cost = torch.empty(n, dtype=torch.int32, device=dev)
    d_len = torch.empty(n, dtype=torch.int32, device=dev)
    d_status = torch.empty(n, dtype=torch.int32, device=dev)
    d_ops = torch.empty(int(ops_off_h[-1]) + 16, dtype=torch.uint8, device=dev)
    info = ctx.plan_batch(batch.off1, batch.off2, ops_off_h)

    def step_device():
        ctx.run_batch(d_s1.data_ptr(), d_off1.data_ptr(), d_s2.data_ptr(), d_off2.data_ptr(),
                      d_cost.data_ptr(), d_ops.data_ptr(), d_opsoff.data_ptr(), d_len.data_ptr(),
                      d_status.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fill_ms, walk_ms = [], []
    barrier()
    t_lo = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    t_hi = time.perf_counter()
    clocks = sampler.stop(t_lo, t_hi)
    dev_ms = ev0.elapsed_time(ev1)
    wall_ms = (t_hi - t_lo) * 1e3
    dev_cost = d_cost.cpu().numpy()
    # Per-kernel durations for the roofline: the same work planned with GA_F_SERIAL (one
    # stream, no fill/walk overlap), library CUDA events on the launching stream.
    info_serial = ctx.plan_batch(batch.off1, batch.off2, ops_off_h, flags=_native.GA_F_SERIAL)
    for k in range(2 + min(5, args.steps)):
        step_device()
        t = ctx.timing()
        if k >= 2:
            fill_ms.append(t["fill"])
            walk_ms.append(t["walk"])
    assert np.array_equal(d_cost.cpu().numpy(), dev_cost), "serial and pipelined plans disagree"

    # ---- end to end through the host-buffer call, pinned host memory.  Nucleotide workloads are
    # fed as 2-bit packed letter streams (GA_F_SEQ_2BIT, the usual on-disk / in-memory form of
    # DNA reads; 4x fewer H2D bytes); the byte-per-letter form is timed beside it.
    def pinned(a):
        tt = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return tt, tt.numpy()
    keep = []
    outs = []
    for shape, dt in ((n, np.int32), (int(ops_off_h[-1]), np.uint8), (n + 1, np.int64), (n, np.int32), (n, np.int32)):
        tt = torch.empty(shape, dtype=getattr(torch, np.dtype(dt).name)).pin_memory()
        keep.append(tt)
        outs.append(tt.numpy())
    outs[2][:] = ops_off_h
    d2h = outs[0].nbytes + outs[1].nbytes + outs[3].nbytes + outs[4].nbytes
    e2e_steps = max(2, min(args.steps, 5))

    lens1, lens2 = np.diff(batch.off1), np.diff(batch.off2)
    fixed = bool((lens1 == lens1[0]).all() and (lens2 == lens2[0]).all())
    um, un = int(lens1[0]), int(lens2[0])

    def time_e2e(bits, uniform):
        """uniform: the fixed-length entry point (ga_align_batch_uniform, no offset arrays);
        otherwise ga_align_batch with its three offset arrays."""
        hb = []
        for a in (_native.pack_letters(s1h, bits), batch.off1, _native.pack_letters(s2h, bits), batch.off2):
            tt, v = pinned(a)
            keep.append(tt)
            hb.append(v)
        fl = _native.seq_flag(bits)
        if uniform:
            uo = (outs[0], outs[1], outs[3], outs[4])
            call = lambda: ctx.align_batch_uniform(hb[0], um, hb[2], un, n, flags=fl, out=uo)
            h2d_bytes = hb[0].nbytes + hb[2].nbytes
        else:
            call = lambda: ctx.align_batch(hb[0], hb[1], hb[2], hb[3], flags=fl, out=tuple(outs))
            h2d_bytes = hb[0].nbytes + hb[2].nbytes + 2 * batch.off1.nbytes + ops_off_h.nbytes
        for _ in range(2):
            call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call()
        torch.cuda.synchronize()
        dt_s = (time.perf_counter() - t0) / e2e_steps
        assert np.array_equal(outs[0], dev_cost), "device-resident and host-buffer paths disagree"
        return dt_s, h2d_bytes, ctx.timing()

    seq_bits = 2 if len(dense.letters) <= 4 else 8
    e2e_s, h2d, e2e_phase = time_e2e(seq_bits, fixed)
    e2e_offsets = time_e2e(seq_bits, False) if fixed else (e2e_s, h2d, e2e_phase)
    e2e_plain = time_e2e(8, False) if seq_bits != 8 else e2e_offsets

    # ---- reduce over ranks: max time, sum of cells
    vals = torch.tensor([dev_ms / args.steps, e2e_s * 1e3, float(cells), e2e_plain[0] * 1e3, e2e_offsets[0] * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        step_ms, e2e_ms, total_cells, e2e_plain_ms, e2e_off_ms = mx[0].item(), mx[1].item(), sm[2].item(), mx[3].item(), mx[4].item()
    else:
        step_ms, e2e_ms, total_cells, e2e_plain_ms, e2e_off_ms = vals[0].item(), vals[1].item(), float(cells), vals[3].item(), vals[4].item()
    if rank != 0:
        return

    value = total_cells / (step_ms / 1e3) / 1e9
    e2e_value = total_cells / (e2e_ms / 1e3) / 1e9
    fms = sum(fill_ms) / len(fill_ms)
    wms = sum(walk_ms) / len(walk_ms)
    fill_gcups = cells / (fms / 1e3) / 1e9
    hbm_peak, hbm_src = peaks()
    try:
        int32_measured = ctx.alu_peak(0) / 1e3          # Tops/s, IADD chains
        mix_measured = ctx.alu_peak(5) / 1e3            # Tinstr/s on the packed cell's own mix
        vimnmx3 = ctx.alu_peak(1) / 1e3
    except Exception:
        int32_measured = mix_measured = vimnmx3 = None
    achieved_tops = fill_gcups * 1e9 * OPS_PER_CELL / 1e12
    peak_tops = int32_measured or NOMINAL_INT32_TOPS
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "s16x2" if info["packed16"] else "int32", "data": "synthetic",
        "config": {"workload": DESCR[name], "pairs_per_gpu": n, "cells_per_gpu": cells,
                   "l2": "inputs + direction codes (%.1f GB per step) are far larger than the 126 MB L2"
                         % (info["code_bytes"] / 1e9),
                   "fill_kernel": info["fill_kernel"], "chunks": info["n_chunks"],
                   "pipeline": "walk of chunk c overlaps fill of chunk c+1 (2 fill + 2 walk streams); "
                               "kernels_ms / roofline are from a serial (non-overlapped) plan"},
        "clocks": clocks, "wall_ms_timed_region": wall_ms, "host_cores_bound_rank0": numa_cores,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms,
                "input": "%d-bit letters, pinned host memory, %s" % (
                    seq_bits, "fixed-length call (ga_align_batch_uniform)" if fixed else "offset arrays (ga_align_batch)"),
                "offset_array_input": {"value": total_cells / (e2e_off_ms / 1e3) / 1e9, "ms_per_step": e2e_off_ms,
                                       "h2d_bytes_per_step": int(e2e_offsets[1])},
                "phases_ms": {k: round(v, 3) for k, v in e2e_phase.items()},
                "byte_per_letter_input": {"value": total_cells / (e2e_plain_ms / 1e3) / 1e9, "ms_per_step": e2e_plain_ms,
                                          "h2d_bytes_per_step": int(e2e_plain[1])}},
        "gpu_launches": int(info["launches_per_run"]) * args.steps,
        "kernels_ms": {"fill": fms, "walk": wms},
        "roofline": {"bound": "int_alu", "kernel": info["fill_kernel"],
                     "achieved": achieved_tops, "peak": peak_tops, "unit": "Tops/s",
                     "frac": achieved_tops / peak_tops,
                     "peak_source": "measured in this run (ga_alu_peak mode 0: int32 IADD chains)"
                     if int32_measured else "nominal 148 SM x 64 lanes x 1.965 GHz",
                     "frac_of_nominal_18.6": achieved_tops / NOMINAL_INT32_TOPS,
                     "ops_per_cell": OPS_PER_CELL, "fill_gcups": fill_gcups,
                     "measured_mix_tinstr_s": mix_measured, "measured_vimnmx3_s16x2_tinstr_s": vimnmx3,
                     "traffic": (ncu_traffic(info["fill_kernel"]) or {}).get("dram_bytes"),
                     "traffic_capture": ncu_traffic(info["fill_kernel"])},
        "roofline_hbm": {"bound": "hbm", "kernel": info["fill_kernel"] + " (direction-code store)",
                         "achieved": cells * CODE_BYTES_PER_CELL / (fms / 1e3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": cells * CODE_BYTES_PER_CELL / (fms / 1e3) / 1e9 / hbm_peak,
                         "peak_source": hbm_src, "stored_bytes_per_step": info["code_bytes"],
                         "traffic": (ncu_traffic(info["fill_kernel"]) or {}).get("dram_bytes")},
    }
    if world == 1 and not args.no_cpu_baseline:
        g, ns, dt, th = cpu_port_gcups(batch, dense, remap)
        line["cpu_baseline"] = {
            "value": g, "unit": UNIT, "cores": th, "kind": "port",
            "sample": f"first {ns} pairs of {name} ({dt:.1f} s), C restatement of the reference's fill + "
                      f"value-comparing traceback (oracle/ga_oracle.c), pthreads"}
        line["python_port_mcups_per_core"] = python_port_mcups(batch, scheme)
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.workload in ("cfg1", "cfg4"):
            run_long(args, rank, world, local_rank)
        else:
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
