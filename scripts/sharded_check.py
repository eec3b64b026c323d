"""torchrun --nproc-per-node N scripts/sharded_check.py — unit sharding across GPUs (mmrs_ctx_set_shard):
every rank must reproduce the single-GPU / oracle-golden result bit for bit, and the sweep time must drop."""
import json, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "multimoda-rs_b200")]
import torch
import torch.distributed as dist
import bench
import multimodars as mm
from multimodars import _dist, _native as nat
from tests import golden_io as gio

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = mm.get_context(local)
if os.environ.get("MMRS_SHARD_TRANSPORT", "nccl") == "callback":
    _dist.enable_unit_sharding(ctx)      # host callback transport (torch.distributed all-reduce), axis 1
else:
    _dist.init_comm(ctx)                 # the library's own NCCL communicator: collectives on device buffers

# 1. config 1 (golden): full mode on the example pullbacks, default + brute force
pack, gold = gio.inputs(), gio.oracle_outputs()
ins = [gio.py_input(mm, pack, n, d, f"{n}{d}") for n, d in (("rest", True), ("rest", False), ("stress", True), ("stress", False))]
ok = True
for tag, kw in (("default", dict(step_rotation_deg=0.5, bruteforce=False, smooth=True)),
                ("brute0p5", dict(step_rotation_deg=0.5, bruteforce=True, smooth=False))):
    ab, cd, ac, bd, logs = mm.from_array_full(*ins, range_rotation_deg=90.0, sample_size=500, write_obj=False, postprocessing=False, **kw)
    for i in range(4):
        ok &= np.array_equal(np.array(logs[i], dtype=np.float64).reshape(-1, 7), gold[f"cfg1_{tag}_logs_{i}"])
    outs = [ab.geom_a, ab.geom_b, cd.geom_a, cd.geom_b, ac.geom_a, ac.geom_b, bd.geom_a, bd.geom_b]
    for i, g in enumerate(outs):
        ok &= gio.sha(g.to_blob()) == str(gold[f"cfg1_{tag}_out_sha_{i}"])
print(f"rank {rank}/{world}: config 1 sharded == golden: {ok}", flush=True)

# 2. strong scaling of one config-2 case through the public API
def inp(seed, dia):
    fr = bench.synthetic_pullback(200, 500, seed)
    z = 0.5 * (199 - np.arange(200))
    rows = np.concatenate([np.column_stack([np.full(500, float(i)), f, np.full(500, z[i])]) for i, f in enumerate(fr)])
    last = rows[rows[:, 0] == 199][0]
    return mm.numpy_to_inputdata(rows, np.array([199, last[1] + 0.1, last[2], last[3]]), dia, label="d" if dia else "s")
a, b = inp(20261018, True), inp(20261019, False)
best = None
for _ in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    pair, logs = mm.from_array_singlepair(a, b, step_rotation_deg=0.01, range_rotation_deg=180.0, sample_size=500,
                                          write_obj=False, bruteforce=True, smooth=True, postprocessing=False)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    best = dt if best is None or dt < best else best
h = gio.sha(np.concatenate([np.array(l, dtype=np.float64).reshape(-1) for l in logs]))
hs = [None] * world
dist.all_gather_object(hs, h)
same = len(set(hs)) == 1
st = ctx.process_stats()
if rank == 0:
    out = dict(world=world, config1_matches_golden=bool(ok), config2_all_ranks_identical=bool(same), config2_logs_sha=h,
               config2_wall_s=best, config2_evals_total=14436483, config2_evals_per_s=14436483 / best, rank0_stats=st)
    print(json.dumps(out), flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"sharded_check_n{world}.json").write_text(json.dumps(out, indent=1))
dist.destroy_process_group()
sys.exit(0 if (ok and same) else 1)
