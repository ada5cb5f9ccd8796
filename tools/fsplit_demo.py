#!/usr/bin/env python
"""Latency mode (secondary partitioning): ONE capture buffer, the frequency hypotheses split across the GPUs,
xc_peak_freq finished by one NCCL all_reduce(MAX) over packed {power bits, ~f index} keys (sweep.xcorr_pss_fsplit).
Rank 0 checks that pow / frq are bit-identical to the whole grid on one GPU.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/fsplit_demo.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "lte-cell-scanner_b200"))
import torch
import torch.distributed as dist
import lcs_b200 as L
import sweep


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    d = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        d = dist
    cu8 = np.load(os.path.join(ROOT, "tests/golden/capbuf_0000.npz"))["cu8"].reshape(-1, 2)
    fc = 739e6
    f = L.f_search_set(fc, 120.0)
    ctx = L.Context(local)
    plans = {}

    def run_slice(fsub):
        key = (float(fsub[0]), len(fsub))
        if key not in plans:
            plans[key] = ctx.plan(cu8.shape[0], fsub, 2, fc, fc, 1.92e6, max_batch=1)
        o = plans[key].run_host_np(cu8[None], L.IQ_CU8, want_single=False)
        return dict(pow=o["pow"][0], frq=o["frq"][0], sp_incoherent=o["sp_incoherent"][0])

    sweep.xcorr_pss_fsplit(run_slice, f, dist=d, device=dev)          # warm-up: plan build, NCCL communicator
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        pw, frq, spi = sweep.xcorr_pss_fsplit(run_slice, f, dist=d, device=dev)
    dt = (time.perf_counter() - t0) / reps
    if rank == 0:
        full = run_slice(f)
        t0 = time.perf_counter()
        for _ in range(reps):
            full = run_slice(f)
        dt1 = (time.perf_counter() - t0) / reps
        ok = np.array_equal(pw, full["pow"]) and np.array_equal(frq, full["frq"])
        print("f-split over %d GPU(s): %.3f ms per buffer (whole grid on one GPU: %.3f ms); pow/frq identical to the single-GPU result: %s"
              % (world, dt * 1e3, dt1 * 1e3, ok))
        assert ok
    for p in plans.values():
        p.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
