#!/usr/bin/env python
"""Where does a channel of the sweep spend its host time?  (LCS_TRACE_PLANS=1 prints plan build / eviction cost.)"""
import os, sys, time
import numpy as np
os.environ["LCS_TRACE_PLANS"] = "1"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "lte-cell-scanner_b200"))
import lcs_b200 as L
ctx = L.Context(0)
rng = np.random.default_rng(1)
cap = np.clip(np.round(127.5 + 20 * rng.standard_normal((153600, 2))), 0, 255).astype(np.uint8)
for i in range(14):
    fc = 730e6 + i * 100e3
    t0 = time.perf_counter()
    f = L.f_search_set(fc, 120.0)
    t1 = time.perf_counter()
    cells, peaks = ctx.cell_search(cap, f, fc, fc, 1.92e6)
    t2 = time.perf_counter()
    cells, peaks = ctx.cell_search(cap, f, fc, fc, 1.92e6)     # cached plan
    t3 = time.perf_counter()
    print("channel %2d: f_search_set %.3f ms, first search %.3f ms, repeat (cached plan) %.3f ms" % (i, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
