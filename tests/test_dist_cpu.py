"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: unit sharding, the result all-gather,
the angle-sharded arg-min combine and the case-sharded orchestration wrapper. The per-rank GPU
compute is replaced by the CPU oracle here — this file tests the plumbing, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodars import _dist


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 398, 3996):
        for world in (1, 2, 3, 8):
            seen = [i for r in range(world) for i in _dist.shard_range(n, r, world)]
            assert seen == list(range(n))
            sizes = [len(_dist.shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    sys.path[:0] = [str(root), str(root / "multimoda-rs_b200")]
    from multimodars import _dist as D
    from oracle import oracle_py as ora
    from tests import fixtures as fx

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. unit-sharded results: 5 units dealt 3 + 2
        mine = D.shard_range(5, rank, world)
        idx = np.array([10 * u for u in mine])
        dst = np.array([0.5 + u for u in mine])
        gi, gd = D.gather_unit_results(idx, dst)
        assert gi.tolist() == [0, 10, 20, 30, 40] and gd.tolist() == [0.5, 1.5, 2.5, 3.5, 4.5]

        # 2. angle-sharded single unit vs the oracle's full sweep (ties -> lowest global index)
        rng = np.random.default_rng(0)
        phi = np.linspace(0, 2 * np.pi, 64, endpoint=False)
        ref = np.stack([2 * np.cos(phi), 1.5 * np.sin(phi)], 1) + rng.normal(0, 0.01, (64, 2))
        test = np.stack([2 * np.cos(phi + 0.3), 1.5 * np.sin(phi + 0.3)], 1)
        full = ora.sweep(test, ref, (0.0, 0.0), 0, 1.0, 180.0)
        lo, hi = D.candidate_shard(len(full["costs"]), rank, world)
        part = full["costs"][lo:hi]
        li = int(np.argmin(part))
        g_idx, g_d = D.combine_angle_sharded(li, float(part[li]), lo)
        assert g_idx == full["index"] and g_d == full["cost"]
        # exact tie across the shard boundary: the lower global index must win
        g_idx, _ = D.combine_angle_sharded(0, 1.0, lo)
        assert g_idx == 0

        # 3. case-sharded orchestration: 3 single-pullback cases over 2 ranks == serial oracle
        blobs = [ora.encode_geometry(fx.dummy_geometry()) for _ in range(3)]

        def run_local(mode, bl, step, rg, ss, sm, bf):
            outs, logs = [], []
            for b in bl:
                o, l, _ = ora.align_within(b, step, rg, sm, bf, ss)
                outs.append(o)
                logs.append(l)
            return outs, logs, [False] * len(bl)

        table, local_out = D.process_cases_sharded(1, blobs, 0.5, 30.0, 6, False, False, run_local)
        want = ora.align_within(blobs[0], 0.5, 30.0, False, False, 6)[1]
        for c in range(3):
            assert np.array_equal(table[c][0], want)
        assert sorted(local_out) == list(D.shard_range(3, rank, world))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
