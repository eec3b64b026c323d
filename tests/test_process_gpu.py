"""GPU parity of the mode orchestration (mmrs_process_cases behind from_file_* / from_array_*)
against the CPU oracle and the committed goldens: per-frame rotation / translation logs and every
output geometry must be BIT-IDENTICAL (north_star: "selected rotation angle and translation per
frame must match the reference's path exactly")."""
import numpy as np
import pytest

import multimodars as mm
from multimodars import _native as nat
from oracle import oracle_py as ora
from tests import fixtures as fx
from tests import golden_io as gio

pytestmark = pytest.mark.gpu


def oracle_blobs(pack, names_dia):
    out = []
    for name, dia in names_dia:
        a = gio.phase_arrays(pack, name, dia)
        out.append(ora.build_geometry_from_arrays(a["lumen"], a["ref_point"], a["eem"], a["calc"], a["side"],
                                                  a["records"], dia, name))
    return out


FULL = [("rest", True), ("rest", False), ("stress", True), ("stress", False)]


def logs_array(logs):
    return np.array(logs, dtype=np.float64).reshape(-1, 7)


@pytest.mark.parametrize("tag,kw", [
    ("default", dict(step_rotation_deg=0.5, range_rotation_deg=90.0, bruteforce=False, smooth=True)),
    ("brute0p5", dict(step_rotation_deg=0.5, range_rotation_deg=90.0, bruteforce=True, smooth=False)),
    ("hier0p05", dict(step_rotation_deg=0.05, range_rotation_deg=90.0, bruteforce=False, smooth=False)),
])
def test_config1_from_array_full_matches_golden_and_oracle(tag, kw):
    """BASELINE config 1: examples ivus_rest + ivus_stress, 4-phase full mode."""
    pack, gold = gio.inputs(), gio.oracle_outputs()
    ins = [gio.py_input(mm, pack, n, d, f"{n}_{'dia' if d else 'sys'}") for n, d in FULL]
    ab, cd, ac, bd, logs = mm.from_array_full(*ins, sample_size=500, write_obj=False, postprocessing=False, **kw)
    for i in range(4):
        assert np.array_equal(logs_array(logs[i]), gold[f"cfg1_{tag}_logs_{i}"]), (tag, i)
    outs = [ab.geom_a, ab.geom_b, cd.geom_a, cd.geom_b, ac.geom_a, ac.geom_b, bd.geom_a, bd.geom_b]
    for i, g in enumerate(outs):
        assert gio.sha(g.to_blob()) == str(gold[f"cfg1_{tag}_out_sha_{i}"]), (tag, i)
    assert ab.label == "rest_dia - rest_sys" and bd.label == "rest_sys - stress_sys"
    assert isinstance(ab, mm.PyGeometryPair) and isinstance(logs[0][0], tuple) and len(logs[0][0]) == 7
    st = mm.get_context().process_stats()
    assert st["units"] > 0 and st["evals"] > 0 and st["launches"] > 0


def test_config1_from_file_full(tmp_path):
    pack, gold = gio.inputs(), gio.oracle_outputs()
    rest, stress = gio.write_dir(pack, "rest", tmp_path / "ivus_rest"), gio.write_dir(pack, "stress", tmp_path / "ivus_stress")
    ab, cd, ac, bd, logs = mm.from_file_full(str(rest), str(stress), write_obj=False, postprocessing=False)
    for i in range(4):
        assert np.array_equal(logs_array(logs[i]), gold[f"cfg1_default_logs_{i}"])
    assert gio.sha(bd.geom_b.to_blob()) == str(gold["cfg1_default_out_sha_7"])
    assert ab.geom_a.label == "ivus_rest" and cd.label == "ivus_stress - ivus_stress"


@pytest.mark.parametrize("mode", [3, 2, 1])
def test_other_modes_match_oracle(mode):
    pack = gio.inputs()
    n_in = nat.N_IN[mode]
    names = FULL[:n_in]
    blobs = oracle_blobs(pack, names)
    want_out, want_logs = ora.process(mode, blobs, 0.5, 90.0, True, False, 500, threads=8)
    ins = [gio.py_input(mm, pack, n, d, n) for n, d in names]
    kw = dict(step_rotation_deg=0.5, range_rotation_deg=90.0, sample_size=500, write_obj=False)
    if mode == 3:
        ab, cd, logs = mm.from_array_doublepair(*ins, postprocessing=False, **kw)
        got = [ab.geom_a, ab.geom_b, cd.geom_a, cd.geom_b]
    elif mode == 2:
        pair, logs = mm.from_array_singlepair(*ins, postprocessing=False, **kw)
        got = [pair.geom_a, pair.geom_b]
    else:
        g, l0 = mm.from_array_single(ins[0], **kw)
        got, logs = [g], (l0,)
    for i in range(n_in):
        assert np.array_equal(logs_array(logs[i]), want_logs[i])
    for g, w in zip(got, want_out):
        assert np.array_equal(g.to_blob(), w)


def test_idealized_fixture_within_kat():
    """align_within.rs:855-887 through the product: |rot| = 15 +- 1, tx = -0.01 k, ty = +0.01 k, anomalous."""
    pack, gold = gio.inputs(), gio.oracle_outputs()
    inp = gio.py_input(mm, pack, "ideal", True, "stress")
    g, logs = mm.from_array_single(inp, step_rotation_deg=0.01, range_rotation_deg=20.0, sample_size=200, smooth=True)
    la = logs_array(logs)
    assert np.array_equal(la, gold["ideal_within_logs"])
    assert gio.sha(g.to_blob()) == str(gold["ideal_within_out_sha"])
    for k, row in enumerate(la):
        assert abs(abs(row[2]) - 15.0) <= 1.0
        assert row[3] == pytest.approx(-0.01 * (k + 1), abs=1e-3) and row[4] == pytest.approx(0.01 * (k + 1), abs=1e-3)


def test_rust_kat_simple_geometry_through_product():
    """align_within.rs:791-830 (dummy hexagon chain): rot = -15, tx = ty = -k. The 6-point polygon is
    full of exact ties, so this also exercises the chain-resolve path."""
    ctx = mm.get_context()
    blob = ora.encode_geometry(fx.dummy_geometry())
    want_out, want_logs, _ = ora.align_within(blob, 0.01, 30.0, False, False, 6)
    out, logs, _ = nat.process_cases(ctx, 1, [blob], 0.01, 30.0, 6, False, False)
    assert np.array_equal(logs[0], want_logs)
    assert np.array_equal(out[0], want_out)
    for i, row in enumerate(logs[0]):
        assert row[2] == pytest.approx(-15.0, abs=1e-6) and row[3] == pytest.approx(-(i + 1.0), abs=1e-6)


def test_between_kat_through_product():
    """align_between.rs:281-303: B = A rotated by 15 deg per frame -> found rotation -15 deg."""
    ctx = mm.get_context()
    a = ora.encode_geometry(fx.dummy_geometry_aligned_long())
    want_pairs, want_logs = ora.process(2, [a, a], 0.01, 30.0, False, False, 6)
    out, logs, _ = nat.process_cases(ctx, 2, [a, a], 0.01, 30.0, 6, False, False)
    for g, w in zip(out, want_pairs):
        assert np.array_equal(g, w)
    for l, w in zip(logs, want_logs):
        assert np.array_equal(l, w)


def test_synthetic_batch_of_cases_matches_oracle():
    """Several independent synthetic cases in ONE call (the batch axis the GPU design adds): every case
    must equal what the oracle returns for it alone."""
    ctx = mm.get_context()
    blobs = []
    for case in range(3):
        for ph in range(2):
            lumen, rp = fx.synthetic_pullback(12, 120, seed=1000 * case + ph)
            blobs.append(nat.geometry_from_arrays(lumen, rp, diastole=(ph == 0), label=f"c{case}p{ph}"))
    out, logs, _ = nat.process_cases(ctx, 2, blobs, 0.1, 45.0, 100, True, False)
    for case in range(3):
        want_out, want_logs = ora.process(2, blobs[2 * case:2 * case + 2], 0.1, 45.0, True, False, 100, threads=8)
        for k in range(2):
            assert np.array_equal(logs[2 * case + k], want_logs[k]), (case, k)
            assert np.array_equal(out[2 * case + k], want_out[k]), (case, k)


def test_reference_guards_surface_as_errors():
    ctx = mm.get_context()
    blob = ora.encode_geometry(fx.dummy_geometry())
    with pytest.raises(nat.MmrsError, match="sample_size must be > 0"):   # align_within.rs:38-40
        nat.process_cases(ctx, 1, [blob], 0.5, 30.0, 0, False, False)
    with pytest.raises(nat.MmrsError, match="Geometry contains no frames"):  # :32-34
        nat.process_cases(ctx, 1, [np.array([0.0])], 0.5, 30.0, 6, False, False)


@pytest.mark.parametrize("mode", [4, 2])
def test_postprocessing_matches_oracle(mode):
    """postprocessing=True (the reference default): postprocess_geom_pair (postprocessing.rs:12-87) on
    every pair — resample to a common z spacing, align reference frames, trim — bit-identical."""
    pack = gio.inputs()
    names = FULL[:nat.N_IN[mode]]
    blobs = oracle_blobs(pack, names)
    want_out, want_logs = ora.process(mode, blobs, 0.5, 90.0, True, False, 500, threads=8, postprocessing=True)
    ins = [gio.py_input(mm, pack, n, d, n) for n, d in names]
    kw = dict(step_rotation_deg=0.5, range_rotation_deg=90.0, sample_size=500, write_obj=False, postprocessing=True)
    if mode == 4:
        ab, cd, ac, bd, logs = mm.from_array_full(*ins, **kw)
        got = [ab.geom_a, ab.geom_b, cd.geom_a, cd.geom_b, ac.geom_a, ac.geom_b, bd.geom_a, bd.geom_b]
    else:
        pair, logs = mm.from_array_singlepair(*ins, **kw)
        got = [pair.geom_a, pair.geom_b]
    for g, w in zip(got, want_out):
        assert np.array_equal(g.to_blob(), w)
    # trimmed to the common span: both members of a pair have the same number of frames
    for k in range(0, len(got), 2):
        assert len(got[k].frames) == len(got[k + 1].frames)
    plain, _ = ora.process(mode, blobs, 0.5, 90.0, True, False, 500, threads=8, postprocessing=False)
    assert any(not np.array_equal(a, b) for a, b in zip(plain, want_out))   # the step does something


def test_unit_sharding_across_two_gpus():
    """mmrs_ctx_set_shard under torchrun with 2 ranks: every rank reproduces the golden config-1 result and
    all ranks agree on config 2 (scripts/sharded_check.py). Needs 2 GPUs; skipped on a 1-GPU box."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = gio.GOLD.parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(root / "scripts" / "sharded_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"config1_matches_golden": true' in r.stdout and '"config2_all_ranks_identical": true' in r.stdout


class _ThreadAllReduce:
    """An in-place all-reduce(SUM) over int64 arrays between THREADS of this process: the transport of
    mmrs_ctx_set_shard (any all-reduce will do: NCCL, MPI, gloo — here a barrier and a sum), so that the unit partition
    and its merge run on a 1-GPU box with one context per emulated rank."""

    def __init__(self, world):
        import threading

        self.world, self.slots = world, [None] * world
        self.barrier = threading.Barrier(world, timeout=120)
        self.calls = [0] * world

    def for_rank(self, rank):
        def allreduce(arr):
            self.calls[rank] += 1
            self.slots[rank] = arr
            self.barrier.wait()
            total = np.sum([self.slots[r] for r in range(self.world)], axis=0)
            self.barrier.wait()          # every rank has read every buffer
            arr[:] = total
            self.barrier.wait()          # every rank has written its own before the slots are reused
        return allreduce


def _run_ranks(world, body):
    """body(rank, ctx) on `world` threads, each with its own context on cuda:0 sharded through _ThreadAllReduce."""
    import threading

    ar, out, errs = _ThreadAllReduce(world), [None] * world, []

    def run(rank):
        ctx = nat.Context(0)
        try:
            ctx.set_shard(rank, world, ar.for_rank(rank))
            out[rank] = body(rank, ctx)
        except Exception as e:  # noqa: BLE001
            errs.append((rank, repr(e)))
            ar.barrier.abort()
        finally:
            ctx.close()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(300)
    assert not errs, errs
    assert not any(t.is_alive() for t in ts)
    return out, ar


@pytest.mark.parametrize("world", [2, 3])
def test_unit_partition_with_a_host_transport_on_one_gpu(world):
    """mmrs_ctx_set_shard: every emulated rank makes the same mmrs_process_cases call (config 1, brute force at 0.05 deg:
    the 80-unit intrapullback batch is dealt to the ranks in cost-balanced blocks, the 32-byte results are merged by the
    exchange; the inter-pullback batches are below the partition threshold and are swept by every rank). All ranks must
    return the unpartitioned result bit for bit."""
    pack = gio.inputs()
    blobs = oracle_blobs(pack, FULL)
    args = (4, blobs, 0.05, 90.0, 500, False, True)
    want_out, want_logs, _ = nat.process_cases(mm.get_context(), *args)
    got, ar = _run_ranks(world, lambda rank, ctx: nat.process_cases(ctx, *args))
    for rank in range(world):
        out, logs, _ = got[rank]
        assert all(np.array_equal(a, b) for a, b in zip(logs, want_logs)), rank
        assert all(np.array_equal(a, b) for a, b in zip(out, want_out)), rank
    assert ar.calls[0] >= 1 and len(set(ar.calls)) == 1    # the ranks entered the exchange equally often


def test_unit_partition_leaves_a_rank_without_units():
    """Two units on three ranks (explicit unit axis): one rank owns nothing, launches nothing, and still ends up with both
    results after the merge; a degenerate grid and an empty set keep their flags through the sum."""
    rng = np.random.default_rng(41)
    phi = np.linspace(0, 2 * np.pi, 520, endpoint=False)

    def contour(rot):
        r = 2.4 * (1 + 0.2 * np.cos(2 * phi) + 0.03 * np.cos(3 * phi + 0.4))
        return np.stack([r * np.cos(phi + rot), r * np.sin(phi + rot)], 1) + rng.normal(0, 0.004, (520, 2))

    tests = [contour(0.3), contour(-0.2), np.zeros((0, 2)), contour(0.1)]
    refs = [contour(0.0), contour(0.05), contour(0.0), contour(0.0)]
    toff = np.concatenate([[0], np.cumsum([len(t) for t in tests])])
    roff = np.concatenate([[0], np.cumsum([len(r) for r in refs])])
    grids = [nat.make_grid(0.05, 90.0), nat.make_grid(0.0, 10.0, center=0.25)]        # the second one is degenerate
    gou = np.array([0, 0, 0, 1], dtype=np.int32)
    args = (np.concatenate(tests), toff, np.concatenate(refs), roff, np.zeros((4, 2)), grids)
    kw = dict(grid_of_unit=gou, mode=0, tie_margin=1e-9)
    ctx0 = nat.Context(0)
    want = ctx0.sweep_batched(*args, partition=-1, **kw)
    ctx0.close()
    got, ar = _run_ranks(3, lambda rank, ctx: ctx.sweep_batched(*args, partition=1, **kw))
    for rank in range(3):
        for f in ("best_idx", "best_angle", "best_dist", "best_dist_f32", "n_shortlist", "n_ties", "flags"):
            assert np.array_equal(got[rank][f], want[f]), (rank, f)
    assert ar.calls == [1, 1, 1]


# ---- edge cases the reference handles on this path ---------------------------------------------------------
def _cmp_single(blob, step, rng, sample, smooth, brute):
    ctx = mm.get_context()
    want_out, want_logs, want_an = ora.align_within(blob, step, rng, smooth, brute, sample)
    out, logs, an = nat.process_cases(ctx, 1, [blob], step, rng, sample, smooth, brute)
    assert np.array_equal(logs[0], want_logs)
    assert np.array_equal(out[0], want_out)
    assert an[0] == want_an


def test_no_catheter_and_downsampling():
    """n_points = 0 (no synthetic catheter, align_within.rs:44-59 -> None) and sample_size < contour length
    (strided down-sampling, contour.rs:47-58)."""
    pack = gio.inputs()
    a = gio.phase_arrays(pack, "rest", True)
    blob = nat.geometry_from_arrays(a["lumen"], a["ref_point"], records=a["records"], diastole=True, n_points=0)
    assert np.array_equal(blob, ora.build_geometry_from_arrays(a["lumen"], a["ref_point"], records=a["records"],
                                                                diastole=True, n_points=0))
    _cmp_single(blob, 0.5, 45.0, 200, False, False)
    _cmp_single(blob, 1.0, 30.0, 37, True, True)


def test_ragged_frames_and_tiny_geometries():
    """Frames with different point counts (possible through the blob boundary: only ingest enforces equal counts,
    integrity_check.rs:121-166), a two-frame pullback and a single-frame pullback (no frame pair at all)."""
    rng = np.random.default_rng(5)
    frames = []
    for i, n in enumerate((40, 64, 33, 57, 40)):
        phi = np.linspace(0, 2 * np.pi, n, endpoint=False)
        r = 2.0 + 0.3 * np.cos(2 * phi + 0.2 * i) + rng.normal(0, 0.01, n)
        x, y = 4.5 + r * np.cos(phi + 0.1 * i), 4.5 + r * np.sin(phi + 0.1 * i)
        pts = np.stack([np.full(n, i), np.arange(n), x, y, np.full(n, float(i)), np.zeros(n)], 1)
        c = (float(x.mean()), float(y.mean()), float(i))
        frames.append(dict(id=i, centroid=c, reference_point=np.array([i, 0, 8.0, 4.5, float(i), 0]) if i == 0 else None,
                           contours={0: dict(kind=0, id=i, original_frame=10 - i, centroid=c, aortic_thickness=None,
                                             pulmonary_thickness=None, points=pts)}))
    # smoothing indexes neighbours by the current frame's point count (geometry.rs:165-239) -> keep it off for ragged input
    _cmp_single(ora.encode_geometry(frames), 0.5, 60.0, 500, False, False)
    _cmp_single(ora.encode_geometry(frames[:2]), 0.1, 20.0, 16, False, False)
    _cmp_single(ora.encode_geometry(frames[:1]), 0.5, 20.0, 16, False, False)


def test_step_regimes_of_find_best_rotation():
    """All four arms of the coarse-to-fine driver (align_within.rs:208-246): >= 1, [0.1, 1), [0.01, 0.1), < 0.01,
    plus a range smaller than the 5-degree medium window."""
    lumen, rp = fx.synthetic_pullback(6, 160, seed=42)
    blob = nat.geometry_from_arrays(lumen, rp, diastole=True)
    for step, rng_deg in ((2.0, 40.0), (0.3, 40.0), (0.03, 40.0), (0.004, 40.0), (0.03, 3.0), (0.004, 0.05)):
        _cmp_single(blob, step, rng_deg, 120, False, False)


def test_reference_defaults_write_obj(tmp_path):
    """from_file_full with the reference's defaults (write_obj=True, postprocessing=True, interpolation_steps=0):
    four output directories, per pair 2 meshes x 3 contour types x (obj, mtl, png) (to_object/process.rs:9-61)."""
    import os
    pack = gio.inputs()
    rest, stress = gio.write_dir(pack, "rest", tmp_path / "ivus_rest"), gio.write_dir(pack, "stress", tmp_path / "ivus_stress")
    outs = [str(tmp_path / "out" / n) for n in ("rest", "stress", "diastole", "systole")]
    ab, cd, ac, bd, logs = mm.from_file_full(str(rest), str(stress), output_path_ab=outs[0], output_path_cd=outs[1],
                                             output_path_ac=outs[2], output_path_bd=outs[3])
    for d, pair in zip(outs, (ab, cd, ac, bd)):
        names = sorted(os.listdir(d))
        want = sorted(f"{t}_{i:03d}_{pair.label}.{e}" for t in ("lumen", "catheter", "wall") for i in (0, 1)
                      for e in ("obj", "mtl", "png"))
        assert names == want, d
        obj = open(os.path.join(d, f"lumen_000_{pair.label}.obj")).read().splitlines()
        n_pts = sum(len(f.lumen) for f in pair.geom_a.frames)
        assert sum(l.startswith("v ") for l in obj) == n_pts + 2          # + the two cap centroids (watertight)
        assert sum(l.startswith("vt ") for l in obj) == n_pts + 2 and obj[n_pts] == f"mtllib lumen_000_{pair.label}.mtl"
        p0 = pair.geom_a.frames[0].lumen.points[0]
        assert obj[0] == f"v {p0.x!r} {p0.y!r} {p0.z!r}".replace(".0 ", " ").removesuffix(".0")
    # the single-geometry export of single_processing_rs (entry.rs:741-775)
    g, _ = mm.from_file_single(str(rest), write_obj=True, output_path=str(tmp_path / "single"))
    assert sorted(os.listdir(tmp_path / "single")) == sorted(
        f"{t}_ivus_rest.{e}" for t in ("lumen", "catheter", "wall") for e in ("obj", "mtl"))


def test_pipelined_cohort_is_identical_to_one_call():
    """_dist.process_cases_pipelined (two host threads, two contexts / streams on one GPU, chunks of cases) returns
    exactly what one mmrs_process_cases call over the whole cohort returns."""
    from multimodars import _dist
    pack = gio.inputs()
    blobs = []
    for _ in range(5):   # five identical-input cases are still five independent cases
        blobs += oracle_blobs(pack, FULL[:2])
    ctx = mm.get_context()
    want_out, want_logs, want_an = nat.process_cases(ctx, 2, blobs, 0.5, 90.0, 500, True, False, True)
    got_out, got_logs, got_an, stats = _dist.process_cases_pipelined(0, 2, blobs, 0.5, 90.0, 500, True, False, True,
                                                                   chunk_cases=2, workers=2)
    assert len(got_out) == len(want_out) == 10 and got_an == want_an
    for a, b in zip(got_out, want_out):
        assert np.array_equal(a, b)
    for a, b in zip(got_logs, want_logs):
        assert np.array_equal(a, b)
    assert stats["units"] == 5 * (ctx.process_stats()["units"] // 5)
