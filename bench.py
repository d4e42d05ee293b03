#!/usr/bin/env python3
"""Benchmark of the hot path: GCUPS with traceback (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--configs cfg3,cfg5,cfg4,cfg1|none]
                    [--workload cfg2] [--impl ours|reference]

A "step" is one pass of the hot path (fill + direction-code store + traceback walk) over one batch
of synthetic pairs of the BASELINE shape.  The headline workload is BASELINE.json configs[1]
(1,000,000 DNA pairs, 150 x 150) per GPU; with N > 1 every rank aligns its own shard of
N x 1,000,000 pairs (pairs are independent: no collective on the data path, "scaling": "weak").

After the headline's timed region the same run measures the other BASELINE configs and reports them
as sub-records under `configs` (same fields, fewer steps):
    N = 1:  cfg3 (100k protein pairs), cfg5 (10M DNA pairs 250 bp), cfg4 (100 kb x 100 kb), cfg1 (1 kb x 1 kb)
    N > 1:  cfg5_strong (the 10M pairs of configs[4] sharded over the N GPUs: strong scaling) and
            devices_api (the one-process multi-GPU call find_global_alignments(devices=[0..N-1]), rank 0).

Per record:
`value`   device-resident inputs -> device-resident cost + packed ops, CUDA events on the launching
          stream, barrier + synchronize on both sides, max over ranks.
`e2e`     the same work through the public host-buffer call: pinned host inputs, H2D, planning,
          kernels, D2H of costs / ops / lengths / status, all timed.
`parity`  the GPU results of the CPU-timed sample compared with the oracle's, pair by pair
          (cost, status, every op): {"pairs", "mismatches"}.
`roofline`     fill kernel vs the integer-ALU pipe: thread-instructions the kernel issues on that pipe
               per DP cell (from the committed ncu capture, profiles/traffic_r2.json) x cells/s, over the
               pipe's rate measured in this run (ga_alu_peak mode 1).  `frac_contract` is the SURVEY 8d
               figure (24 int32 ops per cell over the measured int32 IADD rate); it exceeds 1 for the
               packed kernels because one s16x2 instruction updates two cells.
`roofline_hbm` direction-code store vs the measured HBM peak (MEASURED_PEAKS.json).
`cpu_baseline` the C oracle port on the host cores, on a bounded sample (rank 0, N = 1).
`--impl reference` times that CPU port alone with all host threads and, beside it, the unmodified
          Python reference (oracle/_ref, copied there by build()) on a small sample.
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
CODE_BYTES_PER_CELL = 0.5   # 2 bits (level-0 argmin) + 2 bits (c1 ^ c2), DESIGN.md section 3
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
    ap.add_argument("--workload", default="cfg2", choices=sorted(DESCR))  # the headline record
    ap.add_argument("--configs", default=None,
                    help="comma list of further configs measured after the headline (default: all others at "
                         "N = 1, cfg5_strong + devices_api at N > 1; 'none' = headline only)")
    ap.add_argument("--pairs", type=int, default=None, help="pairs per GPU of the headline (default: BASELINE size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
def ncu_capture(kernel: str):
    """Per-kernel figures of the committed ncu captures (profiles/traffic_r2.json, else r1)."""
    for fn in ("traffic_r2.json", "traffic_r1.json"):
        try:
            e = json.loads((ROOT / "profiles" / fn).read_text()).get(kernel)
        except (OSError, ValueError):
            e = None
        if e:
            return dict(e, file="profiles/" + fn)
    return None


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

    def window(self, t_lo: float, t_hi: float) -> dict:
        """Median SM clock / active throttle reasons of the samples taken in [t_lo, t_hi]
        (a timed region); if the region was shorter than the sampling period, of the
        samples taken while the GPU was busy just around it."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        good = [(t, r) for t, r in list(self.rows) if len(r) >= 8 and r[1].replace(".", "").isdigit()]
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

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


def oracle_dense(dense):
    from oracle import c_oracle
    d = c_oracle.Dense.__new__(c_oracle.Dense)
    d.A, d.sub, d.gap_h, d.gap_v, d.o, d.max_cost = (dense.A, dense.sub, dense.gap_h, dense.gap_v,
                                                     dense.gap_open, dense.max_cost)
    return d


def cpu_port_sample(batch, dense, remap, target_s: float = 12.0, threads: int | None = None):
    """C oracle (port of the reference's path) on the host cores, bounded sample.
    -> (gcups, n_pairs, seconds, threads, oracle outputs of the sample)"""
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
    out = c_oracle.align_batch(remap[sub.s1], sub.off1, remap[sub.s2], sub.off2, d, threads)
    dt = time.perf_counter() - t0
    return sub.cells() / dt / 1e9, n, dt, threads, out


def count_mismatches(oracle_out, cost, ops, ops_off, ops_len, status) -> int:
    """Pairs of the oracle's sample whose GPU result differs (cost, status or any op).
    oracle_out = c_oracle.align_batch(...) of the first n pairs; GPU arrays cover at least those."""
    o_ops, o_off, o_n, o_cost, o_status = oracle_out
    n = len(o_cost)
    bad = (np.asarray(cost[:n]) != o_cost) | ((np.asarray(status[:n]) != 0) != (o_status != 0))
    live = o_status == 0
    bad |= live & (np.asarray(ops_len[:n]).astype(np.int64) != o_n)
    # pack the oracle's byte-per-op stream like the library does (2 bits per op, walk order,
    # pair p at byte ops_off[p]) and compare byte by byte
    ops_off = np.asarray(ops_off[: n + 1], dtype=np.int64)
    total = int(o_off[n])
    if total:
        pair = np.repeat(np.arange(n, dtype=np.int64), np.diff(o_off[: n + 1]))
        t = np.arange(total, dtype=np.int64) - o_off[:n][pair]
        keep = live[pair] & (t < o_n[pair])
        pair, t = pair[keep], t[keep]
        byte = ops_off[pair] - ops_off[0] + (t >> 2)
        want = np.zeros(int(ops_off[n] - ops_off[0]), dtype=np.uint8)
        np.bitwise_or.at(want, byte, (o_ops[:total][keep] << (2 * (t & 3)).astype(np.uint8)).astype(np.uint8))
        got = np.asarray(ops[int(ops_off[0]): int(ops_off[n])])
        nbytes = (o_n + 3) >> 2
        idx = np.arange(len(want), dtype=np.int64)
        owner = np.searchsorted(ops_off[1: n + 1] - ops_off[0], idx, side="right")
        inside = idx - (ops_off[:n] - ops_off[0])[owner] < nbytes[owner]
        diff = (want != got) & inside & live[owner]
        if diff.any():
            bad[np.unique(owner[diff])] = True
    return int(bad.sum())


def python_port_mcups(batch, scheme, n: int = 4):
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


# ----------------------------------------------------------------------------- reference arm
def _ref_worker(job):
    """One pair through the UNMODIFIED reference (oracle/_ref/globalign, copied from
    /root/reference/src by build()): find_global_alignment, its public entry point."""
    s1, s2, kw = job
    sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
    from globalign.globaligner import find_global_alignment as ref_align   # noqa: E402
    r = ref_align(seq_1=s1, seq_2=s2, **kw)
    return r.cost


def real_reference_sample(name, batch, target_s: float = 8.0):
    """The real (pure-Python) reference on a small sample, all host cores (one pair per task).
    -> dict or None when oracle/_ref is absent."""
    if not (ROOT / "oracle" / "_ref" / "globalign" / "globaligner.py").exists():
        return None
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    kw = synth.SCHEMES[name]
    per_pair = batch.cells() / batch.n_pairs
    n = int(max(1, min(batch.n_pairs, target_s * 0.45e6 * cores / per_pair)))   # ~0.45 MCUPS per core
    jobs = [(*batch.pair(k), kw) for k in range(n)]
    cells = sum(len(a) * len(b) for a, b, _ in jobs)
    ctx = mp.get_context("spawn")
    with ctx.Pool(min(cores, n)) as pool:
        pool.map(_ref_worker, jobs[: min(cores, n)])           # import + warm-up, untimed
        t0 = time.perf_counter()
        costs = pool.map(_ref_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return {"kind": "_ref", "value": cells / dt / 1e9, "unit": UNIT, "cores": min(cores, n),
            "mcups_per_core": cells / dt / 1e6 / min(cores, n), "seconds": dt,
            "sample": f"first {n} pairs of {name} through the unmodified globalign.find_global_alignment "
                      f"(multiprocessing.Pool, one pair per task)", "costs_head": [int(c) for c in costs[:4]]}


def run_reference_long(args, name, scheme):
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
    if name == "cfg1":
        line["reference_python"] = real_reference_sample(name, batch, target_s=3.0)
    print(json.dumps(line), flush=True)


def run_reference(args, rank):
    if rank != 0:
        return
    from globalign_b200.batch import Scheme
    name = args.workload
    scheme = Scheme(**synth.SCHEMES[name])
    threads = os.cpu_count() or 1
    if name in ("cfg1", "cfg4"):
        return run_reference_long(args, name, scheme)
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
        "config": {"workload": DESCR[name], "sample": f"first {n_sample} pairs per step",
                   "note": "a GCUPS rate on a bounded sample of the same workload (same shapes, same scheme); "
                           "the GPU arm's rate is over the full batch"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"first {n_sample} pairs of {name}, C restatement of the reference's "
                                   f"fill + value-comparing traceback (oracle/ga_oracle.c), pthreads"},
        "reference_python": real_reference_sample(name, batch),
        "python_port_mcups_per_core": python_port_mcups(batch, scheme),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- shared state of our arm
class Env:
    """torch / distributed plumbing of one rank."""

    def __init__(self, rank, world, local_rank):
        import torch
        self.torch = torch
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dev = torch.device("cuda", local_rank)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.host_pg = None
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
            self.host_pg = dist.new_group(backend="gloo")       # host-side waits (no spinning GPU kernel)
        # one nvidia-smi poller per node: with eight ranks, eight of them (NVML queries every 100 ms)
        # perturb the host-side copies that the end-to-end number times
        self.sampler = ClockSampler(local_rank)
        if rank == 0:
            self.sampler.start()
        self._alu = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.host_pg)

    def reduce(self, vals_max, vals_sum):
        """max over ranks of the first list, sum over ranks of the second."""
        torch = self.torch
        a = torch.tensor(vals_max, dtype=torch.float64, device=self.dev)
        b = torch.tensor(vals_sum, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(a, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(b, op=self.dist.ReduceOp.SUM)
        return a.tolist(), b.tolist()

    def alu_peaks(self, ctx):
        """Integer-pipe rates measured on this GPU (T thread-instructions / s), once per run."""
        if self._alu is None:
            try:
                self._alu = {"iadd": ctx.alu_peak(0) / 1e3, "alu_pipe": ctx.alu_peak(1) / 1e3,
                             "cell_mix": ctx.alu_peak(5) / 1e3}
            except Exception:
                self._alu = {"iadd": None, "alu_pipe": None, "cell_mix": None}
        return self._alu


def roofline_record(env, ctx, kernel, cells, fill_ms, lanes):
    """The pipe-aware integer roofline of one fill kernel (module docstring)."""
    pk = env.alu_peaks(ctx)
    cap = ncu_capture(kernel)
    fill_gcups = cells / (fill_ms / 1e3) / 1e9
    contract_tops = fill_gcups * 1e9 * OPS_PER_CELL / 1e12
    peak_pipe = pk["alu_pipe"] or NOMINAL_INT32_TOPS
    rec = {"bound": "int_alu", "kernel": kernel, "unit": "Tinstr/s", "lanes": lanes, "fill_gcups": fill_gcups,
           "peak": peak_pipe,
           "peak_source": "ALU-pipe rate measured in this run (ga_alu_peak mode 1: VIMNMX3 chains, thread-instructions/s)"
           if pk["alu_pipe"] else "nominal 148 SM x 64 lanes x 1.965 GHz",
           "achieved": None, "frac": None, "instr_per_cell": None,
           "frac_contract": contract_tops / (pk["iadd"] or NOMINAL_INT32_TOPS),
           "contract": {"ops_per_cell": OPS_PER_CELL, "achieved_tops": contract_tops,
                        "peak_tops": pk["iadd"] or NOMINAL_INT32_TOPS,
                        "peak_source": "int32 IADD rate measured in this run (ga_alu_peak mode 0, both pipes)"
                        if pk["iadd"] else "nominal", "frac_of_nominal_18.6": contract_tops / NOMINAL_INT32_TOPS},
           "measured_cell_mix_tinstr_s": pk["cell_mix"], "traffic": None}
    if cap:
        ipc = cap.get("alu_thread_inst_per_cell")
        if ipc:
            rec["instr_per_cell"] = ipc
            rec["achieved"] = fill_gcups * 1e9 * ipc / 1e12
            rec["frac"] = rec["achieved"] / peak_pipe
        rec["traffic"] = cap.get("dram_bytes")
        rec["capture"] = {k: cap.get(k) for k in ("file", "launch", "alu_pipe_pct", "issue_active_pct",
                                                   "duration_ms_under_ncu", "dram_bytes", "algorithmic_bytes") if k in cap}
    return rec


# ----------------------------------------------------------------------------- batched configs (cfg2, cfg3, cfg5)
def bench_batched(env, name, per_rank, steps, warmup, *, scaling="weak", first_rank_block=None,
                  cpu_target_s=12.0, want_cpu=True, e2e_variants=True, string_api=False):
    torch = env.torch
    from globalign_b200 import _native
    from globalign_b200.batch import Scheme
    rank, world, dev = env.rank, env.world, env.dev
    if name == "cfg3":
        batch = synth.protein_varlen(per_rank * world, 3000).slice(rank * per_rank, (rank + 1) * per_rank) \
            if world > 1 else synth.workload(name, n_pairs=per_rank)
    else:
        blocks_per_rank = -(-per_rank // synth._BLOCK)
        batch = synth.workload(name, n_pairs=per_rank, first_block=rank * blocks_per_rank)
    scheme = Scheme(**synth.SCHEMES[name])
    dense = scheme.dense_for(sorted(batch.alphabet))
    remap = np.array([dense.index[ch] for ch in batch.alphabet], dtype=np.uint8)
    s1h = remap[batch.s1]
    s2h = remap[batch.s2]
    cells = batch.cells()
    n = batch.n_pairs

    ctx = _native.Context(env.local_rank)
    ctx.set_scheme(dense)
    ctx.set_stream(env.stream.cuda_stream)     # the torch.cuda.Event pair below brackets exactly these launches
    ops_off_h = _native.ops_offsets(batch.off1, batch.off2)

    # ---- device-resident buffers
    d_s1 = torch.from_numpy(s1h).to(dev)
    d_s2 = torch.from_numpy(s2h).to(dev)
    d_off1 = torch.from_numpy(batch.off1).to(dev)
    d_off2 = torch.from_numpy(batch.off2).to(dev)
    d_opsoff = torch.from_numpy(ops_off_h).to(dev)
    d_cost = torch.empty(n, dtype=torch.int32, device=dev)
    d_len = torch.empty(n, dtype=torch.int32, device=dev)
    d_status = torch.empty(n, dtype=torch.int32, device=dev)
    d_ops = torch.empty(int(ops_off_h[-1]) + 16, dtype=torch.uint8, device=dev)
    info = ctx.plan_batch(batch.off1, batch.off2, ops_off_h)

    def step_device():
        ctx.run_batch(d_s1.data_ptr(), d_off1.data_ptr(), d_s2.data_ptr(), d_off2.data_ptr(),
                      d_cost.data_ptr(), d_ops.data_ptr(), d_opsoff.data_ptr(), d_len.data_ptr(),
                      d_status.data_ptr())

    for _ in range(max(warmup, 3)):
        step_device()
    env.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    t_lo = time.perf_counter()
    ev0.record(env.stream)
    for _ in range(steps):
        step_device()
    ev1.record(env.stream)
    env.barrier()
    t_hi = time.perf_counter()
    clocks = env.sampler.window(t_lo, t_hi)
    dev_ms = ev0.elapsed_time(ev1)
    dev_cost = d_cost.cpu().numpy()
    # the sample the CPU leg will use, taken from THIS (pipelined, timed) plan's outputs
    n_keep = min(n, 400_000)
    keep = {"cost": dev_cost[:n_keep], "len": d_len[:n_keep].cpu().numpy(), "status": d_status[:n_keep].cpu().numpy(),
            "ops": d_ops[: int(ops_off_h[n_keep])].cpu().numpy(), "off": ops_off_h[: n_keep + 1]}
    # Per-kernel durations for the roofline: the same work planned with GA_F_SERIAL (one
    # stream, no fill/walk overlap), library CUDA events on the launching stream.
    fill_ms, walk_ms = [], []
    ctx.plan_batch(batch.off1, batch.off2, ops_off_h, flags=_native.GA_F_SERIAL)
    for k in range(2 + min(5, steps)):
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
    keep_alive = []
    outs = []
    for shape, dt in ((n, np.int32), (int(ops_off_h[-1]), np.uint8), (n + 1, np.int64), (n, np.int32), (n, np.int32)):
        tt = torch.empty(shape, dtype=getattr(torch, np.dtype(dt).name)).pin_memory()
        keep_alive.append(tt)
        outs.append(tt.numpy())
    outs[2][:] = ops_off_h
    d2h = outs[0].nbytes + outs[1].nbytes + outs[3].nbytes + outs[4].nbytes
    e2e_steps = max(2, min(steps, 5))

    lens1, lens2 = np.diff(batch.off1), np.diff(batch.off2)
    fixed = bool((lens1 == lens1[0]).all() and (lens2 == lens2[0]).all())
    um, un = int(lens1[0]), int(lens2[0])

    def time_e2e(bits, uniform):
        """uniform: the fixed-length entry point (ga_align_batch_uniform, no offset arrays);
        otherwise ga_align_batch with its three offset arrays."""
        hb = []
        for a in (_native.pack_letters(s1h, bits), batch.off1, _native.pack_letters(s2h, bits), batch.off2):
            tt, v = pinned(a)
            keep_alive.append(tt)
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
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call()
        torch.cuda.synchronize()
        dt_s = (time.perf_counter() - t0) / e2e_steps
        assert np.array_equal(outs[0], dev_cost), "device-resident and host-buffer paths disagree"
        return dt_s, h2d_bytes, ctx.timing()

    # the compact-output call (ga_align_batch_rle): one 8-byte record per pair + run-length encoded
    # ops packed back to back -- what find_global_alignments uses; far fewer D2H bytes
    rle_cap = int(min((lens1 + lens2).sum(), 48 * n + (1 << 20)))
    rle_out = []
    for shape, dt in ((n, _native.PAIR_RESULT), (rle_cap, np.uint8), (n + 1, np.int64)):
        tt = torch.empty(shape * np.dtype(dt).itemsize, dtype=torch.uint8).pin_memory()
        keep_alive.append(tt)
        rle_out.append(tt.numpy().view(dt))

    def time_e2e_rle(bits):
        hb = []
        for a in (_native.pack_letters(s1h, bits), batch.off1, _native.pack_letters(s2h, bits), batch.off2):
            tt, v = pinned(a)
            keep_alive.append(tt)
            hb.append(v)
        fl = _native.seq_flag(bits)
        if fixed:
            call = lambda: ctx.align_batch_rle(hb[0], None, hb[2], None, um, un, n, flags=fl, out=tuple(rle_out))
            h2d_bytes = hb[0].nbytes + hb[2].nbytes
        else:
            call = lambda: ctx.align_batch_rle(hb[0], hb[1], hb[2], hb[3], flags=fl, out=tuple(rle_out))
            h2d_bytes = hb[0].nbytes + hb[2].nbytes + 2 * batch.off1.nbytes
        for _ in range(2):
            call()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call()
        torch.cuda.synchronize()
        dt_s = (time.perf_counter() - t0) / e2e_steps
        res, rle, roff = rle_out
        assert np.array_equal(res["cost"], dev_cost), "device-resident and compact host-buffer paths disagree"
        for k in range(0, n_keep, max(1, n_keep // 2000)):          # the compact stream decodes to the same ops
            if keep["status"][k] == 0:
                assert np.array_equal(_native.unpack_rle(rle, int(roff[k]), int(roff[k + 1])),
                                      _native.unpack_ops(keep["ops"], int(keep["off"][k]), int(keep["len"][k]))), k
        return dt_s, h2d_bytes, ctx.timing(), int(res.nbytes + roff.nbytes + int(roff[-1]))

    seq_bits = 2 if len(dense.letters) <= 4 else (4 if len(dense.letters) <= 16 else (5 if len(dense.letters) <= 32 else 8))
    e2e_fixed = time_e2e(seq_bits, fixed)
    e2e_s, h2d, e2e_phase, d2h_rle = time_e2e_rle(seq_bits)
    e2e_offsets = time_e2e(seq_bits, False) if (fixed and e2e_variants) else e2e_fixed
    e2e_plain = time_e2e(8, False) if (seq_bits != 8 and e2e_variants) else e2e_offsets

    # ---- reduce over ranks: max time, sum of cells
    (step_ms, e2e_ms, e2e_plain_ms, e2e_off_ms, e2e_fixed_ms), (total_cells,) = env.reduce(
        [dev_ms / steps, e2e_s * 1e3, e2e_plain[0] * 1e3, e2e_offsets[0] * 1e3, e2e_fixed[0] * 1e3], [float(cells)])
    fms = sum(fill_ms) / len(fill_ms)
    wms = sum(walk_ms) / len(walk_ms)
    hbm_peak, hbm_src = peaks()
    cap = ncu_capture(info["fill_kernel"]) or {}
    rec = {
        "metric": METRIC, "value": total_cells / (step_ms / 1e3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "s16x2" if info["packed16"] else "int32", "data": "synthetic",
        "config": {"workload": DESCR[name], "pairs_per_gpu": n, "cells_per_gpu": cells,
                   "l2": "inputs + direction codes (%.1f GB per step) are far larger than the 126 MB L2"
                         % (info["code_bytes"] / 1e9),
                   "fill_kernel": info["fill_kernel"], "chunks": info["n_chunks"],
                   "pipeline": "walk of chunk c overlaps fill of chunk c+1 (2 fill + 2 walk streams); "
                               "kernels_ms / roofline are from a serial (non-overlapped) plan"},
        "clocks": clocks,
        "e2e": {"value": total_cells / (e2e_ms / 1e3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h_rle), "ms_per_step": e2e_ms,
                "call": "ga_align_batch_rle: compact output (8-byte record per pair + run-length encoded ops)",
                "input": "%d-bit letters, pinned host memory, %s" % (
                    seq_bits, "fixed-length form (no offset arrays)" if fixed else "offset arrays"),
                "phases_ms": {k: round(v, 3) for k, v in e2e_phase.items()},
                "fixed_slice_output": {"value": total_cells / (e2e_fixed_ms / 1e3) / 1e9, "ms_per_step": e2e_fixed_ms,
                                       "d2h_bytes_per_step": int(d2h), "h2d_bytes_per_step": int(e2e_fixed[1]),
                                       "phases_ms": {k: round(v, 3) for k, v in e2e_fixed[2].items()},
                                       "call": "ga_align_batch_uniform" if fixed else "ga_align_batch"}},
        "gpu_launches": int(info["launches_per_run"]) * steps,
        "kernels_ms": {"fill": fms, "walk": wms},
        "roofline": roofline_record(env, ctx, info["fill_kernel"], cells, fms,
                                    "s16x2 (two cells per instruction lane)" if info["packed16"] else "int32"),
        "roofline_hbm": {"bound": "hbm", "kernel": info["fill_kernel"] + " (direction-code store)",
                         "achieved": cells * CODE_BYTES_PER_CELL / (fms / 1e3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": cells * CODE_BYTES_PER_CELL / (fms / 1e3) / 1e9 / hbm_peak,
                         "peak_source": hbm_src, "stored_bytes_per_step": info["code_bytes"],
                         "traffic": cap.get("dram_bytes")},
    }
    if e2e_variants and fixed:
        rec["e2e"]["offset_array_input"] = {"value": total_cells / (e2e_off_ms / 1e3) / 1e9, "ms_per_step": e2e_off_ms,
                                            "h2d_bytes_per_step": int(e2e_offsets[1])}
    if e2e_variants and seq_bits != 8:
        rec["e2e"]["byte_per_letter_input"] = {"value": total_cells / (e2e_plain_ms / 1e3) / 1e9,
                                               "ms_per_step": e2e_plain_ms, "h2d_bytes_per_step": int(e2e_plain[1])}

    # ---- CPU leg on rank 0: the port's rate (N = 1) and, always, the parity of its sample
    if rank == 0:
        if world == 1 and want_cpu:
            g, ns, dt, th, oout = cpu_port_sample(batch, dense, remap, target_s=cpu_target_s)
            rec["cpu_baseline"] = {
                "value": g, "unit": UNIT, "cores": th, "kind": "port",
                "sample": f"first {ns} pairs of {name} ({dt:.1f} s), C restatement of the reference's fill + "
                          f"value-comparing traceback (oracle/ga_oracle.c), pthreads"}
            rec["python_port_mcups_per_core"] = python_port_mcups(batch, scheme)
        else:
            _, ns, _, _, oout = cpu_port_sample(batch, dense, remap, target_s=1.0)
        ns = min(ns, n_keep)
        oout = tuple(x[: ns + 1] if i == 1 else (x if i == 0 else x[:ns]) for i, x in enumerate(oout))
        rec["parity"] = {"pairs": ns, "mismatches": count_mismatches(oout, keep["cost"], keep["ops"], keep["off"],
                                                                     keep["len"], keep["status"]),
                         "what": "cost, status and every traceback op of the timed run vs oracle/ga_oracle.c"}

    # ---- the string-level public API (N = 1 headline only)
    if string_api and rank == 0:
        rec["e2e"]["string_api"] = time_string_api(name, batch, dev_cost, cells)
    ctx.close()
    del d_s1, d_s2, d_ops, d_cost, d_len, d_status, keep_alive, outs
    torch.cuda.empty_cache()
    return rec


def time_string_api(name, batch, dev_cost, cells):
    """find_global_alignments(list[str], list[str]) -- what a user of the reference would call for
    many pairs: lowering of the Python strings, GPU, lazy results; then 1,000 AlignmentResults
    rendered (strings, per-pair matrices)."""
    from globalign_b200 import batch as B
    lut = np.frombuffer(batch.alphabet.encode("ascii"), dtype=np.uint8)
    m, n = int(batch.off1[1]), int(batch.off2[1])
    t1 = lut[batch.s1].tobytes().decode("ascii")
    t2 = lut[batch.s2].tobytes().decode("ascii")
    seqs_1 = [t1[k * m:(k + 1) * m] for k in range(batch.n_pairs)]
    seqs_2 = [t2[k * n:(k + 1) * n] for k in range(batch.n_pairs)]
    B.find_global_alignments(seqs_1[:1000], seqs_2[:1000], **synth.SCHEMES[name])       # warm-up
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        res = B.find_global_alignments(seqs_1, seqs_2, **synth.SCHEMES[name])
        t_call = time.perf_counter() - t0
        if best is None or t_call < best:
            best = t_call
    assert np.array_equal(res.cost, dev_cost)
    t0 = time.perf_counter()
    sample = [res[k] for k in range(0, batch.n_pairs, max(1, batch.n_pairs // 1000))][:1000]
    t_render = time.perf_counter() - t0
    assert all(len(r.seq_1_aligned) == len(r.seq_2_aligned) for r in sample)
    return {"value": cells / best / 1e9, "unit": UNIT, "ms_per_call": best * 1e3,
            "render_1000_results_ms": t_render * 1e3,
            "what": "find_global_alignments(list[str], list[str]): string lowering + GPU + lazy BatchAlignment"}


# ----------------------------------------------------------------------------- single long pair (cfg1, cfg4)
def bench_long(env, name, steps, warmup, want_cpu=True):
    """One pair per GPU through ga_align_long (a single long pair stays on one GPU; N > 1 = replicas)."""
    torch = env.torch
    from globalign_b200 import _native
    from globalign_b200.batch import Scheme
    from oracle import c_oracle
    batch = synth.workload(name)
    scheme = Scheme(**synth.SCHEMES[name])
    dense = scheme.dense_for(sorted(batch.alphabet))
    ctx = _native.Context(env.local_rank)
    ctx.set_scheme(dense)
    cells = batch.cells()
    steps = max(1, min(steps, 10))
    for _ in range(max(1, min(warmup, 3))):
        ctx.align_long(batch.s1, batch.s2)
    env.barrier()
    t_lo = time.perf_counter()
    dev_ms, fill_ms, trace_ms, info = 0.0, 0.0, 0.0, None
    for _ in range(steps):
        cost, ops, status, info = ctx.align_long(batch.s1, batch.s2)
        dev_ms += info["fill_ms"] + info["trace_ms"]
        fill_ms += info["fill_ms"]
        trace_ms += info["trace_ms"]
    torch.cuda.synchronize()
    t_hi = time.perf_counter()
    clocks = env.sampler.window(t_lo, t_hi)
    (step_ms, e2e_ms), _ = env.reduce([dev_ms / steps, (t_hi - t_lo) * 1e3 / steps], [0.0])
    total_cells = cells * env.world
    rec = {
        "metric": METRIC, "value": total_cells / (step_ms / 1e3) / 1e9, "unit": UNIT, "n_gpus": env.world,
        "steps": steps, "warmup": min(warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": DESCR[name], "cells_per_gpu": cells, "replicas": env.world,
                   "l2": "checkpoints %.0f MB per step" % (info["checkpoint_bytes"] / 1e6),
                   "tile": [info["tile_rows"], info["tile_cols"]], "retraced_tiles": info["retraced_tiles"],
                   "trace_mode": info["trace_mode"]},
        "clocks": clocks,
        "e2e": {"value": total_cells / (e2e_ms / 1e3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(batch.s1.nbytes + batch.s2.nbytes + 4 * (len(batch.s1) + len(batch.s2) + 2)),
                "d2h_bytes_per_step": int(len(ops) // 4 + 32)},
        "gpu_launches": info["launches"] * steps,
        "kernels_ms": {"fill": fill_ms / steps, "trace": trace_ms / steps},
        "roofline": roofline_record(env, ctx, "long_fill_kernel", cells, fill_ms / steps, "int32"),
    }
    if env.rank == 0:
        # parity: the whole returned path re-decided step by step on forward values by the
        # linear-memory oracle (cfg4), or the oracle's own full alignment (cfg1)
        d = oracle_dense(dense)
        t0 = time.perf_counter()
        if cells <= 4_000_000:
            o_ops, o_cost, rc = c_oracle.align_codes(batch.s1, batch.s2, d)
            bad = int(not (rc == 0 and status == 0 and cost == o_cost and np.array_equal(ops, o_ops)))
            what = "cost and every traceback op vs oracle/ga_oracle.c (full table)"
        else:
            rc, vcost = c_oracle.verify_path(batch.s1, batch.s2, d, ops)
            bad = int(not (rc == 0 and status == 0 and vcost == cost))
            what = "every step of the returned path re-decided on forward values by oracle/ga_oracle.c (gao_verify_path, O(n) memory)"
        dt = time.perf_counter() - t0
        rec["parity"] = {"pairs": 1, "mismatches": bad, "what": what}
        if env.world == 1 and want_cpu:
            rec["cpu_baseline"] = {"value": cells / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                                   "sample": f"the pair itself, {dt:.1f} s: " + what}
    ctx.close()
    return rec


def bench_devices_api(env, n_pairs=200_000):
    """The one-process multi-GPU call: find_global_alignments(..., devices=[0..N-1]) on rank 0
    while the other ranks wait on the host; checked against the single-GPU call."""
    if env.rank != 0:
        return None
    from globalign_b200 import batch as B
    b = synth.workload("cfg2", n_pairs=n_pairs)
    lut = np.frombuffer(b.alphabet.encode("ascii"), dtype=np.uint8)
    t1 = lut[b.s1].tobytes().decode("ascii")
    t2 = lut[b.s2].tobytes().decode("ascii")
    s1 = [t1[k * 150:(k + 1) * 150] for k in range(n_pairs)]
    s2 = [t2[k * 150:(k + 1) * 150] for k in range(n_pairs)]
    devices = list(range(env.world))
    one = B.find_global_alignments(s1, s2, **synth.SCHEMES["cfg2"])
    B.find_global_alignments(s1, s2, devices=devices, **synth.SCHEMES["cfg2"])            # contexts + scratch
    t0 = time.perf_counter()
    many = B.find_global_alignments(s1, s2, devices=devices, **synth.SCHEMES["cfg2"])
    dt_many = time.perf_counter() - t0
    t0 = time.perf_counter()
    B.find_global_alignments(s1, s2, **synth.SCHEMES["cfg2"])
    dt_one = time.perf_counter() - t0
    same = bool(np.array_equal(one.cost, many.cost) and np.array_equal(one.ops_len, many.ops_len)
                and np.array_equal(one.status, many.status)
                and all(np.array_equal(one.ops_of(k), many.ops_of(k)) for k in range(0, n_pairs, 97)))
    return {"devices": devices, "pairs": n_pairs, "identical_to_one_gpu": same,
            "value": b.cells() / dt_many / 1e9, "unit": UNIT, "ms_per_call": dt_many * 1e3,
            "one_gpu_ms_per_call": dt_one * 1e3,
            "what": "find_global_alignments(list[str], list[str], devices=[...]): one process, one host thread and "
                    "context per GPU, equal-cell shards (string lowering included)"}


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
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: globalign_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_cores = bind_to_gpu_numa(local_rank) if world > 1 else None
    env = Env(rank, world, local_rank)
    time.sleep(0.3)
    want_cpu = not args.no_cpu_baseline
    name = args.workload
    if args.configs is None:
        extra = [c for c in ("cfg3", "cfg5", "cfg4", "cfg1") if c != name] if world == 1 else ["cfg5_strong", "devices_api"]
    else:
        extra = [c for c in args.configs.split(",") if c and c != "none"]

    if name in ("cfg1", "cfg4"):
        line = bench_long(env, name, args.steps, args.warmup, want_cpu)
    else:
        per_rank = args.pairs or (DEFAULT_PAIRS[name] // (world if name == "cfg5" else 1))
        line = bench_batched(env, name, per_rank, args.steps, args.warmup,
                             scaling="strong" if (name == "cfg5" and not args.pairs) else "weak",
                             want_cpu=want_cpu, string_api=(world == 1 and name == "cfg2" and want_cpu))
    line["host_cores_bound_rank0"] = numa_cores
    configs = {}
    for c in extra:
        t0 = time.perf_counter()
        try:
            if c in ("cfg1", "cfg4"):
                configs[c] = bench_long(env, c, 5, 3, want_cpu)
            elif c == "cfg5_strong":
                configs[c] = bench_batched(env, "cfg5", DEFAULT_PAIRS["cfg5"] // world, 5, 3, scaling="strong",
                                           want_cpu=False, e2e_variants=False)
            elif c == "devices_api":
                env.host_barrier()
                configs[c] = bench_devices_api(env)
                env.host_barrier()
            elif c in ("cfg2", "cfg3", "cfg5"):
                per_rank = DEFAULT_PAIRS[c] // (world if c == "cfg5" else 1)
                configs[c] = bench_batched(env, c, per_rank, 10 if c != "cfg5" else 5, 3,
                                           scaling="strong" if c == "cfg5" else "weak", cpu_target_s=5.0,
                                           want_cpu=want_cpu, e2e_variants=False)
            else:
                raise ValueError(f"unknown config {c!r}")
        except Exception as exc:          # a failing side record must not lose the headline
            if world > 1:
                raise
            configs[c] = {"error": f"{type(exc).__name__}: {exc}"}
        if isinstance(configs.get(c), dict):
            configs[c]["bench_wall_s"] = round(time.perf_counter() - t0, 1)
    env.sampler.stop()
    if rank != 0:
        return
    if configs:
        line["configs"] = configs
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
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
