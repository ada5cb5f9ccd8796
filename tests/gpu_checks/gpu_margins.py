#!/usr/bin/env python
"""Print the measured parity margins of the CUDA path vs the oracle (run on the GPU box)."""
import os
import sys

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
g = np.load(os.path.join(ROOT, "tests/golden/capbuf_0000.npz"))
cu8 = g["cu8"].reshape(-1, 2)
cases = [("capbuf_0000 n_f=37", cu8, O.f_search_set(739e6, 120.0)), ("synthetic n_f=7", synth(0xC0FFEE), O.f_search_set(739e6, 20.0)),
         ("synthetic n_f=31", synth(0xC0FFEF), O.f_search_set(739e6, 100.0))]
for kern, kname in [(L.KERNEL_FP32, "fp32"), (L.KERNEL_TC, "tc")]:
    for name, c, f in cases:
        cap = ((c.astype(np.float64) - 127) / 128).view(np.complex128).reshape(-1)
        ref = O.xcorr_pss(cap, f, 2, 739e6, 739e6, 1.92e6)
        try:
            plan = ctx.plan(cap.size, f, 2, 739e6, 739e6, 1.92e6, max_batch=1, kernel=kern)
            out = plan.run_host_np(c[None], L.IQ_CU8)
        except L.LcsError as e:
            print("%-5s %-22s unavailable: %s" % (kname, name, e))
            continue
        s = out["single"][0].transpose(0, 2, 1)
        es = np.abs(s - ref["single"]).max() / ref["single"].max()
        ep = np.abs(out["pow"][0] - ref["pow"]).max() / ref["pow"].max()
        mism = int((out["frq"][0] != ref["frq"]).sum())
        esp = np.abs(out["sp_incoherent"][0] / ref["sp_incoherent"] - 1).max()
        print("%-5s %-22s single %.3e  pow %.3e  frq mismatches %d/28800  sp_incoherent %.1e" % (kname, name, es, ep, mism, esp))
        plan.close()
