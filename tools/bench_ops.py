#!/usr/bin/env python
"""bench_ops.py — device-resident timing of every operator on the path at BASELINE.json sizes
(not the driver's bench.py contract; feeds DESIGN.md / profiles/).  One JSON line per operator."""
import argparse
import json
import sys
import time

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vega_b200 as vb
from vega_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=float, default=1e9)
ap.add_argument("--distinct", type=float, default=1e6)
ap.add_argument("--ops", default="reduce,group,sort,sortkv,join,count")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
N, D = int(args.rows), int(args.distinct)
sc = vb.Context(0, profile=True)
stream = sc.stream()
PEAK = 6570.6


def timed(fn, reps):
    fn()  # warm-up
    out = []
    for _ in range(reps):
        torch.cuda.synchronize(); sc.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = fn()
        e1.record(stream)
        sc.synchronize(); torch.cuda.synchronize()
        out.append((e0.elapsed_time(e1), st))
    return out


def run_shuffle(agg, rows=None, keys=None, vals=None, M=8, R=8, hint=0, n=None):
    sh = vb.Shuffle(sc, M, R, L.VB_U64, L.VB_U64, agg, hint=hint)
    src = vb.rdd._Col(rows, allow_rows=True) if rows is not None else vb.rdd._Col(keys)
    vcol = vb.rdd._Col(vals) if vals is not None else None
    starts = vb.slice_starts(src.n, M)
    for m in range(M):
        sh.map(m, src, vcol, int(starts[m]), int(starts[m + 1]))
    sh.seal()
    st = sh.stats()
    sh.free()
    return st


def report(name, n_rows, alg_bytes, res, extra=None):
    ms = sorted(r[0] for r in res)[len(res) // 2]
    st = res[-1][1]
    k = {kk: round(v["ms"], 3) for kk, v in st["kernels"].items() if v["launches"]}
    line = {"op": name, "rows": n_rows, "ms": round(ms, 3), "rows_per_s": n_rows / ms * 1e3,
            "alg_GBps": alg_bytes / ms / 1e6, "roofline_frac": alg_bytes / ms / 1e6 / PEAK, "kernel_ms": k,
            "launches": st["kernel_launches"], "table_slots": st["table_slots"], "restarts": st["table_restarts"]}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


ops = args.ops.split(",")
rows = torch.empty((N, 2), dtype=torch.int64, device="cuda")
sc.gen_pairs(out_rows=rows, first=0, n=N, mode="uniform", n_distinct=D)
if "reduce" in ops:
    report("reduce_by_key(sum)", N, 16.0 * N + 16.0 * D, timed(lambda: run_shuffle(L.VB_AGG_SUM, rows=rows, hint=D), args.reps))
if "count" in ops:
    report("count_by_key", N, 16.0 * N + 16.0 * D, timed(lambda: run_shuffle(L.VB_AGG_COUNT, rows=rows, hint=D), args.reps))
if "group" in ops:
    report("group_by_key", N, 24.0 * N + 16.0 * D, timed(lambda: run_shuffle(L.VB_AGG_GROUP, rows=rows, hint=D), args.reps))
if "sortkv" in ops:
    report("sort_by_key(k,v)", N, 32.0 * N, timed(lambda: run_shuffle(L.VB_AGG_SORT, rows=rows), max(1, args.reps - 1)))
del rows
torch.cuda.empty_cache()
if "zipf" in ops:
    rows = torch.empty((N, 2), dtype=torch.int64, device="cuda")
    sc.gen_pairs(out_rows=rows, first=0, n=N, mode="zipf", n_distinct=D, seed_k=5, zipf_s=1.1)
    report("reduce_by_key(sum) zipf(1.1)", N, 16.0 * N + 16.0 * D, timed(lambda: run_shuffle(L.VB_AGG_SUM, rows=rows), args.reps))
    if "zipfgroup" in ops:
        report("group_by_key zipf(1.1)", N, 24.0 * N + 16.0 * D, timed(lambda: run_shuffle(L.VB_AGG_GROUP, rows=rows), max(1, args.reps - 1)))
    del rows
    torch.cuda.empty_cache()
if "sort" in ops:
    keys = torch.empty(N, dtype=torch.int64, device="cuda")
    sc.gen_pairs(out_keys=keys, first=0, n=N, mode="unique", rank_base=3)
    report("sort_by_key(keys only)", N, 16.0 * N, timed(lambda: run_shuffle(L.VB_AGG_SORT, keys=keys), max(1, args.reps - 1)))
    del keys
    torch.cuda.empty_cache()
if "join" in ops:
    # BASELINE config 4, per-GPU share at 8 GPUs: 6.25e7 rows per side, each rank once per side,
    # 1.25e6 shared keys per GPU (1e7 / 8)
    n_side = int(min(N, 5e8) // 8)
    shared = n_side // 50
    a = torch.empty((n_side, 2), dtype=torch.int64, device="cuda")
    b = torch.empty((n_side, 2), dtype=torch.int64, device="cuda")
    sc.gen_pairs(out_rows=a, first=0, n=n_side, mode="unique", rank_base=0)
    sc.gen_pairs(out_rows=b, first=0, n=n_side, mode="unique", rank_base=n_side - shared)

    def join():
        sa = vb.Shuffle(sc, 8, 8, L.VB_U64, L.VB_U64, L.VB_AGG_COGROUP)
        sb = vb.Shuffle(sc, 8, 8, L.VB_U64, L.VB_U64, L.VB_AGG_COGROUP)
        for s_, t in ((sa, a), (sb, b)):
            col = vb.rdd._Col(t, allow_rows=True)
            st = vb.slice_starts(col.n, 8)
            for m in range(8):
                s_.map(m, col, None, int(st[m]), int(st[m + 1]))
            s_.seal()
        import ctypes
        tot = 0
        for r in range(8):
            nn = ctypes.c_uint64()
            L.check(sc._lib.vb_join_size(sa._h, sb._h, r, ctypes.byref(nn)))
            k, v, w = (torch.empty(nn.value, dtype=torch.int64, device="cuda") for _ in range(3))
            p = lambda t: ctypes.c_void_p(t.data_ptr()) if t.numel() else None
            L.check(sc._lib.vb_join(sa._h, sb._h, r, p(k), p(v), p(w), L.VB_DEVICE))
            tot += nn.value
        st = sa.stats()
        stb = sb.stats()
        for kk in st["kernels"]:
            st["kernels"][kk]["ms"] += stb["kernels"][kk]["ms"]; st["kernels"][kk]["launches"] += stb["kernels"][kk]["launches"]
        st["kernel_launches"] += stb["kernel_launches"]
        st["join_rows"] = tot
        sa.free(); sb.free()
        return st

    res = timed(join, max(1, args.reps - 1))
    assert res[-1][1]["join_rows"] == shared, (res[-1][1]["join_rows"], shared)
    report("join", 2 * n_side, 32.0 * n_side + 24.0 * shared, res, {"join_rows": shared})
sc.close()
