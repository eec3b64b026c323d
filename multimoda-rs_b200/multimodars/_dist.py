"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the GPU box, gloo in CPU tests).

The path shards on independent units: cases / pullback pairs / frame pairs are dealt
to ranks in contiguous balanced blocks with NO data-path collective; only the tiny per-unit results
(index, distance) and the per-frame log rows are exchanged, with all_gather. A single huge unit can
instead be split along the candidate axis; its partial arg-mins are combined with the reference's
tie rule (lowest distance, then lowest global candidate index; process_utils.rs:69-74)."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous balanced block of `n` items for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def candidate_shard(n_cand: int, rank: int, world: int):
    """Contiguous candidate sub-range [lo, hi) of one unit, so that rank order == index order."""
    r = shard_range(n_cand, rank, world)
    return r.start, r.stop


def _dev():
    import torch
    import torch.distributed as dist

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def all_gather_rows(local: np.ndarray, group=None) -> np.ndarray:
    """all_gather of a (k_r, c) float64 array with rank-dependent k_r; returns the rows of all ranks in
    rank order. Two collectives: the row counts, then the padded payload."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    local = np.ascontiguousarray(local, dtype=np.float64)
    c = local.shape[1]
    dev = _dev()
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    sizes = [int(x.item()) for x in cnts]
    pad = max(max(sizes), 1)
    buf = torch.zeros(pad, c, dtype=torch.float64, device=dev)
    if local.shape[0]:
        buf[:local.shape[0]] = torch.from_numpy(local).to(dev)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return np.concatenate([o[:k].cpu().numpy() for o, k in zip(out, sizes)], axis=0)


def gather_unit_results(best_idx: np.ndarray, best_dist: np.ndarray, group=None):
    """Unit-sharded sweep: every rank contributes its units' (index, distance); all ranks get all units."""
    rows = all_gather_rows(np.stack([best_idx.astype(np.float64), best_dist.astype(np.float64)], axis=1), group)
    return rows[:, 0].astype(np.int64), rows[:, 1]


def combine_angle_sharded(local_best_idx: int, local_best_dist: float, cand_lo: int, group=None):
    """Angle-sharded sweep of ONE unit: each rank swept candidates [cand_lo, cand_hi) and holds its local
    leftmost arg-min. Returns the global (index, distance): minimum distance, ties -> lowest global index."""
    rows = all_gather_rows(np.array([[float(cand_lo + local_best_idx), float(local_best_dist),
                                      1.0 if local_best_idx >= 0 else 0.0]]), group)
    best = None
    for gi, d, ok in rows:
        if not ok:
            continue
        if best is None or d < best[1] or (d == best[1] and gi < best[0]):
            best = (gi, d)
    return (-1, 0.0) if best is None else (int(best[0]), float(best[1]))


def process_cases_sharded(mode, blobs, step_deg, range_deg, sample_size, smooth, bruteforce, run_local, group=None):
    """Deals whole cases to ranks, runs `run_local(mode, blobs_of_my_cases, ...)` (mmrs_process_cases on this
    rank's GPU) and all-gathers the per-frame log rows. Returns {case: [log arrays per pullback]} for ALL cases
    on every rank; geometries stay on the rank that computed them (returned as the second value)."""
    import torch.distributed as dist

    from ._native import N_IN

    n_in = N_IN[mode]
    n_cases = len(blobs) // n_in
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard_range(n_cases, rank, world)
    my_blobs = [b for c in mine for b in blobs[c * n_in:(c + 1) * n_in]]
    outs, logs = ([], [])
    if len(mine):
        outs, logs, _ = run_local(mode, my_blobs, step_deg, range_deg, sample_size, smooth, bruteforce)
    rows = []
    for k, c in enumerate(mine):
        for p in range(n_in):
            l = logs[k * n_in + p]
            rows.append(np.column_stack([np.full(len(l), float(c)), np.full(len(l), float(p)), l]))
    local = np.concatenate(rows, axis=0) if rows else np.zeros((0, 9))
    allrows = all_gather_rows(local, group)
    table = {c: [allrows[(allrows[:, 0] == c) & (allrows[:, 1] == p)][:, 2:] for p in range(n_in)]
             for c in range(n_cases)}
    return table, {c: outs[k * len(outs) // max(len(mine), 1):(k + 1) * len(outs) // max(len(mine), 1)]
                   for k, c in enumerate(mine)}


def process_cases_pipelined(device, mode, blobs, step_deg, range_deg, sample_size, smooth, bruteforce,
                            postprocessing=False, chunk_cases=1, workers=4):
    """A cohort on ONE GPU with the host work hidden behind the device work: the cases are cut into chunks and
    `workers` host threads, each with its own context (its own CUDA stream and device workspaces), pull chunks from
    a queue and run mmrs_process_cases on them. While one thread's sweep kernels occupy the GPU, the other decodes
    blobs, builds units, replays the frame chain and encodes results (ctypes releases the GIL for the whole call),
    so the decode / chain / post-step time of chunk k+1 overlaps the sweeps of chunk k. Results come back in case
    order: (out_blobs, logs, anomalous, stats) exactly like _native.process_cases plus the summed counters."""
    import queue
    import threading

    from ._native import N_IN, N_OUT, Context, process_cases

    n_in, n_out = N_IN[mode], N_OUT[mode]
    n_cases = len(blobs) // n_in
    chunks = [range(i, min(i + chunk_cases, n_cases)) for i in range(0, n_cases, chunk_cases)]
    if len(chunks) <= 1 or workers <= 1:
        ctx = Context(device)
        o, l, a = process_cases(ctx, mode, blobs, step_deg, range_deg, sample_size, smooth, bruteforce, postprocessing)
        st = ctx.process_stats()
        ctx.close()
        return o, l, a, st
    q = queue.Queue()
    for k in range(len(chunks)):
        q.put(k)
    res, errs = [None] * len(chunks), []
    stats = {}
    lock = threading.Lock()

    def run():
        ctx = Context(device)
        try:
            while True:
                try:
                    k = q.get_nowait()
                except queue.Empty:
                    return
                c = chunks[k]
                res[k] = process_cases(ctx, mode, blobs[c.start * n_in:c.stop * n_in], step_deg, range_deg, sample_size,
                                       smooth, bruteforce, postprocessing)
                st = ctx.process_stats()
                with lock:
                    for key, v in st.items():
                        stats[key] = stats.get(key, 0) + v
        except Exception as e:  # surfaced on the caller's thread
            errs.append(e)
        finally:
            ctx.close()

    threads = [threading.Thread(target=run) for _ in range(min(workers, len(chunks)))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        raise errs[0]
    outs = [b for r in res for b in r[0]]
    logs = [l for r in res for l in r[1]]
    anom = [a for r in res for a in r[2]]
    assert len(outs) == n_cases * n_out and len(logs) == n_cases * n_in
    return outs, logs, anom, stats


def make_exchange(group=None):
    """The callback mmrs_ctx_set_shard needs: an in-place all-reduce(SUM) of an int64 numpy array across the
    ranks of `group` (NCCL on GPUs, gloo on CPU). Every rank must call it the same number of times."""
    import torch
    import torch.distributed as dist

    def allreduce_sum_int64(arr: np.ndarray) -> None:
        dev = _dev()
        t = torch.from_numpy(arr).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        arr[:] = t.cpu().numpy()

    return allreduce_sum_int64


def init_comm(ctx, group=None, axis: int = 1):
    """Binds an NCCL communicator of its own to `ctx` (mmrs_ctx_comm_init): rank 0 creates the rendezvous token, the
    process group carries it to the other ranks. From then on every batched sweep of `ctx` is partitioned across the
    ranks (axis 1: whole units, 2: candidate angles) and merged by NCCL collectives on device buffers, on the
    context's own stream — no host staging (mmrs_b200.h, "multi-GPU")."""
    import torch.distributed as dist

    from . import _native as nat

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [nat.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(box[0], rank, world)
    ctx.set_partition(axis)


def enable_unit_sharding(ctx, group=None):
    """Shard the units of every batched sweep of `ctx` (one frame pair = one unit) across the ranks of `group`:
    each rank sweeps its block on its own GPU; only 32 B per unit cross NVLink (one all-reduce per search stage)."""
    import torch.distributed as dist

    ctx.set_shard(dist.get_rank(group), dist.get_world_size(group), make_exchange(group))
