"""Host-side multi-rank logic on CPU: 2 gloo ranks shard a synthetic sweep, gather the detected
cells and dedup on rank 0; the result must equal the single-process sweep.  The per-channel
search is a deterministic stand-in (no GPU here) - the real search function is exercised by the
gpu-marked tests."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

WORKER = textwrap.dedent('''
    import json, os, sys
    sys.path.insert(0, os.path.join(%(root)r, "lte-cell-scanner_b200"))
    sys.path.insert(0, os.path.join(%(root)r, "oracle"))
    import numpy as np, torch.distributed as dist
    import sweep, lcs_b200 as L

    def fake_search(fc, cap):
        # channel -> 0..2 cells; neighbouring channels re-detect the same cell with other power
        k = int(round((fc - 715e6) / 100e3))
        out = []
        if k %% 3 != 2:
            out.append(L.new_cell(fc_requested=fc, fc_programmed=fc, n_id_1=90 + (k // 3) %% 4, n_id_2=1,
                                  pss_pow=0.01 * (1 + (k * 7) %% 5), freq_superfine=1000.0 * (k %% 3), ind=k, sfn=k))
        if k %% 4 == 0:
            out.append(L.new_cell(fc_requested=fc, fc_programmed=fc, n_id_1=5, n_id_2=0, pss_pow=0.5 + k, freq_superfine=-2e3, ind=k))
        return out

    chans = [(i, 715e6 + i * 100e3, None) for i in range(23)]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    d = None
    if world > 1:
        dist.init_process_group("gloo")
        d = dist
    res = sweep.sweep(chans, fake_search, L.new_cell, L.dedup, dist=d)
    # the batched variant (all channels of a rank in one call, lcs_sweep_search_cu8) must give the same final list
    fcs = [c[1] for c in chans]
    res2 = sweep.sweep_batched(fcs, None, lambda iq, f: [fake_search(x, None) for x in f], L.new_cell, L.dedup, dist=d)
    if res is not None:
        key = lambda r: [[c.n_id_cell(), c.fc_requested, c.pss_pow, c.ind] for c in r]
        assert key(res) == key(res2)
        print("RESULT " + json.dumps(key(res)))
    if world > 1:
        dist.destroy_process_group()
''')


def _run(world):
    code = WORKER % {"root": ROOT}
    if world == 1:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    else:
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                              "--master-addr", "127.0.0.1", "--master-port", "29617", "--no-python", sys.executable, "-c", code],
                             capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert len(lines) == 1, out.stdout
    return lines[0]


def test_shard_and_pack_roundtrip():
    sys.path.insert(0, os.path.join(ROOT, "lte-cell-scanner_b200"))
    import sweep
    import lcs_b200 as L
    assert sweep.shard(10, 1, 4) == [1, 5, 9] and sweep.shard(3, 3, 4) == []
    cells = [L.new_cell(fc_requested=739e6, fc_programmed=739e6, pss_pow=0.0613795, ind=1410, freq=35000.0, n_id_2=1, n_id_1=92, cp_type=1,
                        frame_start=585.039, freq_fine=35265.2, freq_superfine=35228.4, n_ports=2, n_rb_dl=50,
                        phich_duration=1, phich_resource=3, sfn=74), L.new_cell()]
    back = sweep.array_to_cells(sweep.cells_to_array(cells, 8), L.new_cell)
    assert len(back) == 2
    for k, _ in L.Cell._fields_:
        a, b = getattr(cells[0], k), getattr(back[0], k)
        assert a == b
    assert back[1].ind == -1 and np.isnan(back[1].pss_pow)


def test_two_rank_sweep_equals_single_process():
    assert _run(2) == _run(1)


def test_fsplit_pack_orders_like_the_reference_argmax():
    """Latency mode: the packed {power bits, ~f index} keys reduce with MAX to the reference's xc_peak_freq result (largest
    power, FIRST maximum among equals) for any split of the hypothesis list."""
    sys.path.insert(0, os.path.join(ROOT, "lte-cell-scanner_b200"))
    import sweep
    rng = np.random.default_rng(4)
    n_f = 37
    inc = rng.random((3, 9600, n_f)).astype(np.float32)
    inc[:, ::7, 5] = inc[:, ::7, 20] = 2.0              # exact ties: the lower index must win
    inc[0, 3, :] = 0.0                                  # all-zero column: index 0
    ref_frq = inc.argmax(axis=2).astype(np.int32)       # numpy argmax returns the first maximum, like the strict '>' scan
    ref_pow = inc.max(axis=2).astype(np.float64)
    for world in (1, 2, 3, 8, 40):
        keys = np.full((3, 9600), -1, np.int64)
        for lo, hi in sweep.f_slices(n_f, world):
            if hi == lo:
                continue
            sl = inc[:, :, lo:hi]
            k = sweep.pack_pow_frq(sl.max(axis=2), sl.argmax(axis=2), lo)
            keys = np.maximum(keys, k)                  # what all_reduce(MAX) computes
        pw, frq = sweep.unpack_pow_frq(keys)
        assert np.array_equal(frq, ref_frq) and np.array_equal(pw, ref_pow), world


def test_fsplit_two_ranks_gloo():
    code = textwrap.dedent("""
        import os, sys
        sys.path.insert(0, os.path.join(%(root)r, "lte-cell-scanner_b200"))
        import numpy as np, torch.distributed as dist
        import sweep
        dist.init_process_group("gloo")
        rng = np.random.default_rng(9)
        inc = rng.random((3, 9600, 11)).astype(np.float32)
        f = np.arange(11) * 5000.0
        def run(fsub):
            lo = int(round(fsub[0] / 5000.0)); hi = lo + len(fsub)
            sl = inc[:, :, lo:hi]
            return dict(pow=sl.max(axis=2).astype(np.float64), frq=sl.argmax(axis=2).astype(np.int32), sp_incoherent=np.zeros(9600))
        pw, frq, _ = sweep.xcorr_pss_fsplit(run, f, dist=dist)
        assert np.array_equal(frq, inc.argmax(axis=2)) and np.array_equal(pw, inc.max(axis=2).astype(np.float64))
        if dist.get_rank() == 0:
            print("FSPLIT OK")
        dist.destroy_process_group()
    """) % {"root": ROOT}
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29619", "--no-python", sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "FSPLIT OK" in out.stdout, out.stderr[-2000:]
