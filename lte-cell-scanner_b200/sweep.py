"""Multi-GPU frequency sweep: the per-centre-frequency loop of CellSearch (reference
src/CellSearch.cpp:471-569) sharded over the ranks of one node.

Centre frequencies (one capture buffer each) are independent (CellSearch.cpp:467-469), so rank r
takes channels r, r+W, r+2W, ... and runs the whole chain for each on its own GPU - no data-path
collective.  The only exchange is the final gather of the fixed-size detected-cell records to
rank 0, which applies the reference's cross-frequency `dedup` (CellSearch.cpp:285-319; it needs
all cells).  With torch.distributed this is one all_gather of a padded [max_cells, 16] float64
tensor per rank (NCCL over NVLink on GPUs, gloo on CPU in the tests).
"""
import numpy as np

CELL_FIELDS = ["fc_requested", "fc_programmed", "pss_pow", "ind", "freq", "n_id_2", "n_id_1", "cp_type", "frame_start",
               "freq_fine", "freq_superfine", "n_ports", "n_rb_dl", "phich_duration", "phich_resource", "sfn"]


def shard(n_items, rank, world):
    """Round-robin assignment of channel indices (BASELINE config 4)."""
    return list(range(rank, n_items, world))


def cells_to_array(cells, max_cells):
    """Pack cells (objects with the lcs_cell fields) into a fixed [max_cells+1, 16] float64 array;
    row 0 holds the count.  Every field of lcs_cell is exactly representable in float64."""
    a = np.full((max_cells + 1, len(CELL_FIELDS)), np.nan)
    n = min(len(cells), max_cells)
    a[0, 0] = n
    for i in range(n):
        for j, k in enumerate(CELL_FIELDS):
            a[i + 1, j] = getattr(cells[i], k)
    return a


def array_to_cells(a, new_cell):
    out = []
    int_fields = {"ind", "n_id_2", "n_id_1", "cp_type", "n_ports", "n_rb_dl", "phich_duration", "phich_resource", "sfn"}
    for i in range(int(a[0, 0])):
        kw = {}
        for j, k in enumerate(CELL_FIELDS):
            v = a[i + 1, j]
            kw[k] = int(v) if k in int_fields else float(v)
        out.append(new_cell(**kw))
    return out


def gather_dedup(mine, new_cell, dedup_fn, dist=None, device=None, max_cells_per_rank=256):
    """mine: [(channel index, [cells])] of this rank.  One all_gather of the fixed-size cell records (NCCL over NVLink on
    GPUs, gloo on CPU), then the reference's cross-frequency dedup (CellSearch.cpp:285-319) on rank 0 in channel order.
    Returns the final list on rank 0, None elsewhere."""
    flat = [c for _, cs in mine for c in cs]
    order = [idx for idx, cs in mine for _ in cs]
    if dist is None:
        tagged = sorted(zip(order, range(len(flat))), key=lambda x: x[0])
        return dedup_fn([flat[i] for _, i in tagged])
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    arr = cells_to_array(flat, max_cells_per_rank)
    tag = np.full((max_cells_per_rank + 1, 1), -1.0)
    tag[1:1 + len(order), 0] = order[:max_cells_per_rank]
    t = torch.from_numpy(np.concatenate([arr, tag], axis=1))
    if device is not None:
        t = t.to(device)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    if rank != 0:
        return None
    tagged = []
    for g in gathered:
        g = g.cpu().numpy()
        cs = array_to_cells(g[:, :-1], new_cell)
        tagged += list(zip(g[1:1 + len(cs), -1].astype(int), cs))
    # restore the reference's visiting order (ascending centre frequency) before dedup: the result of
    # dedup depends on the order in which equal-power candidates are met
    tagged.sort(key=lambda x: x[0])
    return dedup_fn([c for _, c in tagged])


def sweep(channels, search_fn, new_cell, dedup_fn, dist=None, device=None, max_cells_per_rank=256):
    """channels: list of (channel_index, fc_requested, capbuf).  search_fn(fc, capbuf) -> list of cells (one channel at a
    time).  Returns the deduplicated list on rank 0 (None elsewhere).  `dist` is torch.distributed (initialised) or None
    for a single process."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    mine = []           # (channel index, cells) in channel order
    for pos in shard(len(channels), rank, world):
        idx, fc, cap = channels[pos]
        mine.append((idx, search_fn(fc, cap)))
    return gather_dedup(mine, new_cell, dedup_fn, dist, device, max_cells_per_rank)


def sweep_batched(fc_all, iq_mine, batch_search_fn, new_cell, dedup_fn, dist=None, device=None, max_cells_per_rank=256):
    """All channels of this rank in one call.  fc_all: centre frequencies of ALL channels (every rank); iq_mine: this
    rank's capture buffers in the order of shard(len(fc_all), rank, world); batch_search_fn(iq, fcs) -> list (per channel)
    of lists of cells (lcs_sweep_search_cu8)."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    idx = shard(len(fc_all), rank, world)
    res = batch_search_fn(iq_mine, [fc_all[i] for i in idx]) if idx else []
    return gather_dedup(list(zip(idx, res)), new_cell, dedup_fn, dist, device, max_cells_per_rank)


# ---------------------------------------------------------------------------------------------------------------------
# Latency mode (secondary partitioning): ONE capture buffer, the frequency hypotheses split across the ranks.
# Every rank runs xcorr_pss on its slice of f_search_set; xc_peak_freq (searcher.cpp:353-383: largest power over f, FIRST
# maximum wins) is finished by one all_reduce(MAX) over packed 64-bit keys {float bits of the power, ~global f index}:
# the power is a non-negative float, so its bit pattern orders like the value, and the complemented index makes the
# lowest f win among equal powers - exactly the strict '>' scan of the reference.
# ---------------------------------------------------------------------------------------------------------------------
def f_slices(n_f, world):
    """Contiguous, balanced slices of the hypothesis list (rank r gets [lo, hi))."""
    base, extra = divmod(n_f, world)
    lo = [r * base + min(r, extra) for r in range(world)]
    return [(lo[r], lo[r] + base + (1 if r < extra else 0)) for r in range(world)]


def pack_pow_frq(pow_f32, frq_local, f_lo):
    """pow_f32: float32 [3][9600] (the collapsed power is a float widened to double in the reference, :380);
    frq_local: int32 index inside this rank's slice.  Returns int64 keys."""
    bits = np.ascontiguousarray(pow_f32, np.float32).view(np.uint32).astype(np.int64)
    idx = (np.asarray(frq_local, np.int64) + f_lo)
    return (bits << 32) | (0xFFFFFFFF - idx)


def unpack_pow_frq(keys):
    keys = np.asarray(keys, np.int64)
    bits = (keys >> 32).astype(np.uint32)
    frq = (0xFFFFFFFF - (keys & 0xFFFFFFFF)).astype(np.int32)
    return bits.view(np.float32).astype(np.float64), frq


def xcorr_pss_fsplit(run_slice, f_set, dist=None, device=None):
    """run_slice(f_subset) -> dict(pow [3][9600] float64, frq [3][9600] int32, sp_incoherent) for this rank's
    hypotheses (an empty slice returns None).  Returns (pow, frq, sp_incoherent) of the whole grid on every rank."""
    import torch
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    lo, hi = f_slices(len(f_set), world)[rank]
    out = run_slice(f_set[lo:hi]) if hi > lo else None
    if out is None:
        keys = np.full((3, 9600), -1, np.int64)          # below every real key (powers are >= 0)
        spi = None
    else:
        keys = pack_pow_frq(out["pow"].astype(np.float32), out["frq"], lo)
        spi = out["sp_incoherent"]
    if dist is not None:
        t = torch.from_numpy(keys)
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        keys = t.cpu().numpy()
    pw, frq = unpack_pow_frq(keys)
    return pw, frq, spi
