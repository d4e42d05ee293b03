#!/usr/bin/env python3
"""Time every fill-kernel configuration on a short-read workload (device-resident inputs).

    python tools/tune_cfgs.py [cfg2|cfg5] [n_pairs]

Prints fill / walk milliseconds and GCUPS per configuration plus the integer-ALU
microbenchmark (ga_alu_peak).  Tuning aid, not a benchmark of record."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from globalign_b200 import _native, synth  # noqa: E402
from globalign_b200.batch import Scheme  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
    b = synth.workload(name, n_pairs=n_pairs)
    dense = Scheme(**synth.SCHEMES[name]).dense_for(sorted(b.alphabet))
    dev = torch.device("cuda", 0)
    ops_off = _native.ops_offsets(b.off1, b.off2)
    d = {k: torch.from_numpy(v).to(dev) for k, v in dict(s1=b.s1, s2=b.s2, off1=b.off1, off2=b.off2, oo=ops_off).items()}
    cost = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    ln = torch.empty_like(cost)
    st = torch.empty_like(cost)
    ops = torch.empty(int(ops_off[-1]) + 16, dtype=torch.uint8, device=dev)
    ctx0 = _native.Context(0)
    res = {"alu_peak_gops": {m: ctx0.alu_peak(k) for k, m in enumerate(
        ["iadd", "vimnmx3_s16x2", "vimnmx3_s32", "lop3", "imad", "packed_cell_mix", "viaddmnmx_s16x2",
         "vimnmx_u16x2", "viadd_16x2", "viaddmnmx_s32", "chain_vimnmx3_s16x2", "chain_viaddmnmx_s16x2",
         "chain_iadd"])}}
    print(json.dumps(res), flush=True)
    ref_cost = None
    cfgs = sys.argv[3].split(",") if len(sys.argv) > 3 else [
        "fill_p16_r5_g32", "fill_p16_r10_g16", "fill_p16_r15_g10", "fill_p16_r19_g8", "fill_p16_r16_g16",
        "fill_p16_r8_g32", "fill_p16_r16_g32", "fill_i32s_r8_g32", "fill_i32s_r8_g16"]
    for cfg in cfgs:
        os.environ["GA_FORCE_CFG"] = cfg
        ctx = _native.Context(0)
        ctx.set_scheme(dense)
        try:
            info = ctx.plan_batch(b.off1, b.off2, ops_off, flags=_native.GA_F_SERIAL)
        except Exception as e:
            print(cfg, "skipped:", e, flush=True)
            continue
        best = None
        for _ in range(4):
            ctx.run_batch(d["s1"].data_ptr(), d["off1"].data_ptr(), d["s2"].data_ptr(), d["off2"].data_ptr(),
                          cost.data_ptr(), ops.data_ptr(), d["oo"].data_ptr(), ln.data_ptr(), st.data_ptr())
            t = ctx.timing()
            if best is None or t["fill"] < best["fill"]:
                best = t
        c = cost.cpu().numpy()
        if ref_cost is None:
            ref_cost = c
        print(json.dumps({"cfg": cfg, "fill_ms": round(best["fill"], 3), "walk_ms": round(best["walk"], 3),
                          "fill_gcups": round(b.cells() / best["fill"] / 1e6, 1),
                          "total_gcups": round(b.cells() / (best["fill"] + best["walk"]) / 1e6, 1),
                          "code_gb": round(info["code_bytes"] / 1e9, 2), "chunks": info["n_chunks"],
                          "same_costs": bool(np.array_equal(c, ref_cost))}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
