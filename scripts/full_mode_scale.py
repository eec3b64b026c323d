"""torchrun --nproc-per-node N scripts/full_mode_scale.py [cfg4_frames] [cfg5_patients]

BASELINE configs[3] ("synthetic full mode: 4 phases x 1000 frames x 2000 points, 0.005 deg sweep, sharded over 8
B200") through the public entry point from_array_full with the frame pairs of the ONE case sharded across ranks
(mmrs_ctx_set_shard), and configs[4] ("batched cohort: 256 synthetic patients in full mode, sharded by patient")
through mmrs_process_cases with whole patients dealt to ranks. Reports wall times and evaluations/s."""
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

F4 = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
P5 = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = mm.get_context(local)
out = dict(world=world)

def rows(seed, n_frames, n_points):
    fr = bench.synthetic_pullback(n_frames, n_points, seed)
    z = 0.5 * (n_frames - 1 - np.arange(n_frames))
    a = np.concatenate([np.column_stack([np.full(n_points, float(i)), f, np.full(n_points, z[i])]) for i, f in enumerate(fr)])
    last = a[a[:, 0] == n_frames - 1][0]
    return a, np.array([n_frames - 1, last[1] + 0.1, last[2], last[3]])

# warm-up
a, rp = rows(1, 6, 64)
mm.from_array_single(mm.numpy_to_inputdata(a, rp, True, label="w"), sample_size=64)

# ---- config 4: one case, units sharded across ranks ---------------------------------------------------------
if F4 > 0:
    _dist.init_comm(ctx)   # the library's NCCL communicator: every batched sweep of the call is partitioned across the ranks
    ins = []
    for k, dia in enumerate((True, False, True, False)):
        a, rp = rows(20261018 + k, F4, 2000)
        ins.append(mm.numpy_to_inputdata(a, rp, dia, label=f"phase{k}"))
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = mm.from_array_full(*ins, step_rotation_deg=0.005, range_rotation_deg=180.0, sample_size=2000, write_obj=False,
                             bruteforce=True, smooth=True, postprocessing=False)
    torch.cuda.synchronize(); dist.barrier()
    wall = time.perf_counter() - t0
    st = ctx.process_stats()   # counters are global (every rank issues every batched sweep; the library partitions it)
    tot = torch.tensor([st["evals"]], dtype=torch.float64, device="cuda")
    logs = res[4]
    h = hash(tuple(np.concatenate([np.array(l, dtype=np.float64).reshape(-1) for l in logs]).tobytes()))
    hs = [None] * world
    dist.all_gather_object(hs, h)
    out["config4"] = dict(frames_per_phase=F4, points=2000, candidates=72000, wall_s=wall, evals_total=float(tot.item()),
                          evals_per_s=float(tot.item()) / wall, all_ranks_identical=len(set(hs)) == 1, rank0_stats=st,
                          plan=ctx.plan())
    if rank == 0:
        print("config4", json.dumps(out["config4"]), flush=True)
    ctx.set_partition(0)

# ---- config 5: cohort, whole patients dealt to ranks -----------------------------------------------------------
if P5 > 0:
    mine = _dist.shard_range(P5, rank, world)
    blobs = []
    for p in mine:
        for k, dia in enumerate((True, False, True, False)):
            a, rp = rows(20261018 + 1000 * p + k, 200, 500)
            blobs.append(nat.geometry_from_arrays(a, rp, diastole=dia, label=f"pt{p}_{k}"))
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    if os.environ.get("MMRS_COHORT_PIPELINE", "1") != "0":   # host work of chunk k+1 behind the sweeps of chunk k
        outs, logs, _, st5 = _dist.process_cases_pipelined(local, 4, blobs, 0.05, 90.0, 500, False, True,
                                                           chunk_cases=int(os.environ.get("MMRS_COHORT_CHUNK", "1")),
                                                           workers=int(os.environ.get("MMRS_COHORT_WORKERS", "4")))
    else:
        outs, logs, _ = nat.process_cases(ctx, 4, blobs, 0.05, 90.0, 500, False, True)
        st5 = ctx.process_stats()
    rows_ = [np.column_stack([np.full(len(l), float(i)), l]) for i, l in enumerate(logs)]
    allrows = _dist.all_gather_rows(np.concatenate(rows_) if rows_ else np.zeros((0, 8)))
    torch.cuda.synchronize(); dist.barrier()
    wall = time.perf_counter() - t0
    st = st5
    tot = torch.tensor([st["evals"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    out["config5"] = dict(patients=P5, wall_s=wall, evals_total=float(tot.item()), evals_per_s=float(tot.item()) / wall,
                          gathered_log_rows=int(len(allrows)), rank0_stats=st,
                          pipelined=os.environ.get("MMRS_COHORT_PIPELINE", "1") != "0",
                          logs_sha=__import__("hashlib").sha256(np.ascontiguousarray(allrows).tobytes()).hexdigest())
    if rank == 0:
        print("config5", json.dumps(out["config5"]), flush=True)

if rank == 0:
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"full_mode_scale_n{world}.json").write_text(json.dumps(out, indent=1))
dist.destroy_process_group()
