#!/usr/bin/env python
"""Small end-to-end pass for compute-sanitizer: both correlators, device peak search, batched search, tracker cycle."""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "lte-cell-scanner_b200"))
import lcs_b200 as L
rng = np.random.default_rng(2)
def synth(n_cap):
    return np.clip(np.round(127.5 + 20 * rng.standard_normal((n_cap, 2))), 0, 255).astype(np.uint8)
ctx = L.Context(0)
for n_cap, f, kern in [(29000, np.array([-5000.0, 0.0, 5000.0, 10000.0]), L.KERNEL_TC), (29000, np.array([0.0]), L.KERNEL_FP32),
                       (29000, np.arange(-16, 17) * 2000.0, L.KERNEL_TC), (29000, np.arange(-5, 6) * 5000.0, L.KERNEL_FP32)]:
    plan = ctx.plan(n_cap, f, 2, 739e6, 739e6, 1.92e6, max_batch=3, kernel=kern)
    cu8 = np.stack([synth(n_cap) for _ in range(3)])
    out = plan.run_host_np(cu8, L.IQ_CU8)
    pk = plan.peaks_batch(cu8, L.IQ_CU8)
    cells = plan.cell_search_batch_cu8(cu8)
    print("n_f=%d kernel=%d pow max %.3e peaks %s cells %s" % (f.size, kern, out["pow"].max(), [len(p) for p in pk], [len(c) for c in cells]))
    plan.close()
g = np.load(os.path.join(ROOT, "tests/golden/capbuf_0000.npz"))
real = g["cu8"].reshape(-1, 2)
new = ctx.tracker_search_cu8(real, 35000.0, 739e6, 739e6, 1.92e6, 0.0)
print("tracker cells", [c.n_id_cell() for c, _ in new])
# multi-plan paths: frequency sweep (one plan per channel, tensor-core correlator) and multi-channel tracker search (FP32, n_f = 1)
sw = L.Sweep(ctx, 29000)
iq = np.stack([synth(29000) for _ in range(5)])
fcs = 739e6 + 100e3 * np.arange(5)
res = sw.search_cu8(iq, fcs, np.arange(-17, 18) * 5000.0)
trk = sw.track_cu8(iq, [100.0, -2000.0, 0.0, 35000.0, 5.0], fcs, late=[0.0] * 5, tracked=[[], [1], [], [], [2, 3]])
print("sweep cells %s tracker cells %s" % ([len(c) for c in res], [len(c) for c in trk]))
sw.close()
best, resid, n = ctx.kalibrate_cu8(synth(29000), 739e6, 739e6, 1.92e6, 120.0)
print("kalibrate on noise: %d cells" % n)
ctx.close()
print("sanity ok")
