#!/usr/bin/env python
"""Bring-up check of the tensor-core correlator: small case first, then the bench shape, then timing."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "lte-cell-scanner_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lcs_b200 as L
import lcs_oracle as O


def synth(seed, n_cap=153600):
    rng = np.random.default_rng(seed)
    return np.clip(np.round(127.5 + 20 * rng.standard_normal((n_cap, 2))), 0, 255).astype(np.uint8)


ctx = L.Context(0)
cases = [("n_cap=29000 n_f=3", synth(1, 29000), np.array([-5000.0, 0.0, 5000.0]), 1),
         ("n_cap=29000 n_f=19 (16,4,1)", synth(5, 29000), np.arange(-9, 10) * 5000.0, 2),
         ("n_cap=153600 n_f=7", synth(2), O.f_search_set(739e6, 20.0), 1),
         ("n_cap=153600 n_f=31 batch3", synth(3), O.f_search_set(739e6, 100.0), 3),
         ("n_cap=153600 n_f=37", synth(4), O.f_search_set(739e6, 120.0), 1)]
for name, c, f, batch in cases:
    cap = ((c.astype(np.float64) - 127) / 128).view(np.complex128).reshape(-1)
    ref = O.xcorr_pss(cap, f, 2, 739e6, 739e6, 1.92e6)
    plan = ctx.plan(cap.size, f, 2, 739e6, 739e6, 1.92e6, max_batch=batch, kernel=L.KERNEL_TC)
    iq = np.stack([c] * batch)
    out = plan.run_host_np(iq, L.IQ_CU8)
    for b in range(batch):
        s = out["single"][b].transpose(0, 2, 1)
        es = np.abs(s - ref["single"]).max() / ref["single"].max()
        ep = np.abs(out["pow"][b] - ref["pow"]).max() / ref["pow"].max()
        mism = int((out["frq"][b] != ref["frq"]).sum())
        print("tc %-28s b=%d single %.3e pow %.3e frq mismatches %d  nonfinite %d zero %d" %
              (name, b, es, ep, mism, int((~np.isfinite(s)).sum()), int((s == 0).sum())), flush=True)
    if es > 1e-3:
        d = np.abs(s - ref["single"])
        t, i, ff = np.unravel_index(d.argmax(), d.shape)
        print("   worst at t=%d idx=%d f=%d: got %.6e ref %.6e; row0 got %s ref %s" %
              (t, i, ff, s[t, i, ff], ref["single"][t, i, ff], s[0, :4, 0], ref["single"][0, :4, 0]))
    plan.close()

# timing, bench shape
import torch
f = O.f_search_set(739e6, 100.0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
plan = ctx.plan(153600, f, 2, 739e6, 739e6, 1.92e6, max_batch=B, kernel=L.KERNEL_TC)
iq = torch.from_numpy(np.stack([synth(100 + i) for i in range(B)])).cuda()
single = torch.empty((B, 3, f.size, 9600), dtype=torch.float32, device="cuda")
pw = torch.empty((B, 3, 9600), dtype=torch.float64, device="cuda"); fq = torch.empty((B, 3, 9600), dtype=torch.int32, device="cuda")
spi = torch.empty((B, 9600), dtype=torch.float64, device="cuda")
plan.timing_enable(True)
for _ in range(3):
    plan.run_device(iq.data_ptr(), L.IQ_CU8, B, single.data_ptr(), pw.data_ptr(), fq.data_ptr(), spi.data_ptr())
torch.cuda.synchronize(); plan.timing_read()
t0 = time.perf_counter()
for _ in range(20):
    plan.run_device(iq.data_ptr(), L.IQ_CU8, B, single.data_ptr(), pw.data_ptr(), fq.data_ptr(), spi.data_ptr())
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ms, n = plan.timing_read()
print("tc timing: %.1f capbufs/s whole call; fold kernel %.3f ms per %d capbufs = %.1f us/capbuf -> %.1f alg TFLOP/s" %
      (20 * B / dt, ms / n, B, 1e3 * ms / n / B, 8 * 137 * 3 * f.size * 144000 * B / (ms / n * 1e-3) / 1e12))
