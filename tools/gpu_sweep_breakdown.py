#!/usr/bin/env python
"""Where the time of a small sweep goes (per-rank load of an 8-GPU 512-channel sweep is 64 channels)."""
import os, sys, time
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "lte-cell-scanner_b200"))
import lcs_b200 as L
import sweep as SW
import torch

def synth(seed, n_cap=153600):
    rng = np.random.default_rng(seed)
    return np.clip(np.round(127.5 + 20 * rng.standard_normal((n_cap, 2))), 0, 255).astype(np.uint8)

ctx = L.Context(0)
f = L.f_search_set(715e6, 120.0)
real = np.load(os.path.join(ROOT, "tests/golden/capbuf_0000.npz"))["cu8"].reshape(-1, 2)
base = [synth(i) for i in range(8)]
sw = L.Sweep(ctx, 153600)
for n_ch in (16, 64, 128, 512):
    fcs = 715e6 + 100e3 * np.arange(n_ch)
    iq = torch.empty((n_ch, 153600, 2), dtype=torch.uint8).pin_memory()
    for k in range(n_ch):
        iq.numpy()[k] = real if k == n_ch // 2 else np.roll(base[k % 8], 31 * k, axis=0)
    for _ in range(2):
        sw.search_cu8(None, fcs, f, host_ptr=iq.data_ptr(), max_cells=4)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        res = sw.search_cu8(None, fcs, f, host_ptr=iq.data_ptr(), max_cells=4)
        t1 = time.perf_counter()
        fin = SW.gather_dedup(list(zip(range(n_ch), res)), L.new_cell, L.dedup)
        t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1))
    a = np.median(np.array(ts), axis=0)
    print("n_ch %4d: search call %.3f ms (%.1f us/channel), pack+dedup %.3f ms, cells %s" % (n_ch, a[0] * 1e3, a[0] / n_ch * 1e6, a[1] * 1e3, sorted(c.n_id_cell() for c in fin)), flush=True)
    # the same channels without the real capture: no per-peak chain
    iq.numpy()[n_ch // 2] = base[0]
    t0 = time.perf_counter()
    for _ in range(5):
        sw.search_cu8(None, fcs, f, host_ptr=iq.data_ptr(), max_cells=4)
    print("          noise only: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3), flush=True)
