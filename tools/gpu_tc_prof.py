#!/usr/bin/env python
"""Run the TC kernel once on the bench shape with LCS_TC_PROF=1 (per-CTA pipeline cycle counters)."""
import os, sys, time
import numpy as np
os.environ["LCS_TC_PROF"] = "1"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "lte-cell-scanner_b200"))
import lcs_b200 as L
import torch
def synth(seed, n_cap=153600):
    rng = np.random.default_rng(seed)
    return np.clip(np.round(127.5 + 20 * rng.standard_normal((n_cap, 2))), 0, 255).astype(np.uint8)
ctx = L.Context(0)
f = L.f_search_set(739e6, 100.0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
plan = ctx.plan(153600, f, 2, 739e6, 739e6, 1.92e6, max_batch=B, kernel=L.KERNEL_TC)
iq = torch.from_numpy(np.stack([synth(100 + i) for i in range(B)])).cuda()
single = torch.empty((B, 3, f.size, 9600), dtype=torch.float32, device="cuda")
pw = torch.empty((B, 3, 9600), dtype=torch.float64, device="cuda"); fq = torch.empty((B, 3, 9600), dtype=torch.int32, device="cuda")
spi = torch.empty((B, 9600), dtype=torch.float64, device="cuda")
plan.timing_enable(True)
for _ in range(3):
    plan.run_device(iq.data_ptr(), L.IQ_CU8, B, single.data_ptr(), pw.data_ptr(), fq.data_ptr(), spi.data_ptr())
torch.cuda.synchronize()
ms, n = plan.timing_read()
print("fold kernel %.3f ms per %d capbufs = %.1f us/capbuf" % (ms / n, B, 1e3 * ms / n / B))
plan.close()
