"""Ingest parity (host-only, no GPU): the product's mmrs_geometry_from_{dir,arrays} and the oracle's
restatement of io/build.rs:9-205 produce bit-identical geometry blobs, and both reproduce the
golden hashes made from the reference's example directories (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import multimodars as mm
from multimodars import _native as nat
from oracle import oracle_py as ora
from tests import golden_io as gio

CASES = [("rest", True), ("rest", False), ("stress", True), ("stress", False)]


@pytest.mark.parametrize("i,case", list(enumerate(CASES)))
def test_arrays_ingest_matches_oracle_and_golden(i, case):
    pack, gold = gio.inputs(), gio.oracle_outputs()
    name, dia = case
    a = gio.phase_arrays(pack, name, dia)
    want = ora.build_geometry_from_arrays(a["lumen"], a["ref_point"], a["eem"], a["calc"], a["side"], a["records"],
                                          dia, name)
    got = nat.geometry_from_arrays(a["lumen"], a["ref_point"], a["eem"], a["calc"], a["side"], a["records"], dia, name)
    assert np.array_equal(got, want)
    assert gio.sha(got) == str(gold[f"cfg1_in_sha_{i}"])   # == what the oracle built from the reference's CSV files


def test_dir_ingest_matches_oracle(tmp_path):
    pack = gio.inputs()
    for name in ("rest", "ideal"):
        d = gio.write_dir(pack, name, tmp_path / name)
        for dia in (True, False):
            assert np.array_equal(nat.geometry_from_dir(d, name, dia), ora.build_geometry_from_dir(d, name, dia))


def test_ingest_errors_match_reference_messages(tmp_path):
    with pytest.raises(nat.MmrsError, match="required contours file missing"):
        nat.geometry_from_dir(tmp_path, "x", True)
    pack = gio.inputs()
    a = gio.phase_arrays(pack, "rest", True)
    bad_ref = a["ref_point"].copy()
    bad_ref[0] = 99999  # frame absent from the lumen -> exactly-one-reference-point check fails (integrity_check.rs:107-118)
    with pytest.raises(nat.MmrsError, match="Expected exactly one reference point, found 0"):
        nat.geometry_from_arrays(a["lumen"], bad_ref, diastole=True)
    with pytest.raises(ora.OracleError, match="Expected exactly one reference point, found 0"):
        ora.build_geometry_from_arrays(a["lumen"], bad_ref, diastole=True)
    ragged = a["lumen"][:-7]  # unequal point counts per frame (integrity_check.rs:121-166)
    with pytest.raises(nat.MmrsError, match="point count mismatch"):
        nat.geometry_from_arrays(ragged, a["ref_point"], diastole=True)


def test_pyinputdata_round_trip_and_types():
    pack = gio.inputs()
    inp = gio.py_input(mm, pack, "rest", True, "rest_dia")
    assert isinstance(inp, mm.PyInputData) and len(inp.lumen) == 20 and len(inp.lumen[0]) == 501
    p = inp.lumen[0].points[3]
    assert isinstance(p, mm.PyContourPoint) and p.point_index == 3
    a = gio.phase_arrays(pack, "rest", True)
    blob = nat.geometry_from_arrays(inp._flat(inp.lumen), a["ref_point"], records=inp._records(), diastole=True,
                                    label="rest_dia")
    assert gio.sha(blob) == str(gio.oracle_outputs()["cfg1_in_sha_0"])
    g = mm.PyGeometry.from_blob(blob, "rest_dia")
    assert repr(g) == "Geometry(20 frames, label='rest_dia')"
    assert np.array_equal(g.to_blob(), blob)
    assert set(g.frames[0].extras) == {"Catheter"} and len(g.frames[0].extras["Catheter"]) == 20


def test_argument_errors_follow_the_reference():
    with pytest.raises(nat.MmrsError):      # missing input directory: anyhow error -> RuntimeError (functions.rs:228)
        mm.from_file_full("does/not/exist/a", "does/not/exist/b")
    with pytest.raises(TypeError):          # binding/align.rs:151
        mm.align_three_point(None, None, None, None, None)


def _write_case(d, rows, delim="\t", eol="\n", header=None, extra_blank=False):
    d.mkdir(parents=True, exist_ok=True)
    for phase in ("diastolic", "systolic"):
        with open(d / f"{phase}_contours.csv", "w", newline="") as f:
            if header:
                f.write(header + eol)
            for k, r in enumerate(rows):
                f.write(delim.join(r) + eol)
                if extra_blank and k % 7 == 3:
                    f.write("  " + eol)
        with open(d / f"{phase}_reference_points.csv", "w", newline="") as f:
            f.write(delim.join(["1", "3.5", "2.0", "0.5"]) + eol)
    return d


@pytest.mark.parametrize("variant", ["tab", "comma_crlf", "header_blank", "aortic_column", "number_forms", "bad_rows",
                                     "no_trailing_newline"])
def test_contour_file_dialects_match_oracle(tmp_path, variant):
    """The in-place contour-file reader (csrc/mmrs_host.cpp read_points) against the oracle's line-by-line reader on the
    file dialects io/input.rs:149-194 accepts: tab or comma, CRLF, a header row, blank lines, a fifth true/false
    column, exponent / signed / padded numbers, rows that fail to parse (dropped) or have another width (dropped)."""
    import math
    n = 12
    pts = [[(2.0 + 0.25 * f) * math.cos(2 * math.pi * k / n) + 3.0, (1.5 + 0.25 * f) * math.sin(2 * math.pi * k / n) + 3.0,
            0.5 * f] for f in range(3) for k in range(n)]
    fr = [f for f in range(3) for _ in range(n)]
    rows = [[str(f), repr(x), repr(y), repr(z)] for f, (x, y, z) in zip(fr, pts)]
    kw = {}
    if variant == "comma_crlf":
        kw = dict(delim=",", eol="\r\n")
    elif variant == "header_blank":
        kw = dict(header="frame\tx\ty\tz", extra_blank=True)
    elif variant == "aortic_column":
        rows = [r + ["true" if i % 2 else "false"] for i, r in enumerate(rows)]
    elif variant == "number_forms":
        forms = [lambda v: f"{v:.6e}", lambda v: f"+{v:.9f}", lambda v: f"  {v:.12f} ", lambda v: f"{v:.5E}",
                 lambda v: f"{v * 1000:.3f}e-3", lambda v: f"{v:.4f}".lstrip("0") if 0 < v < 1 else f"{v:.4f}"]
        rows = [[f"0{f}"] + [forms[(i + j) % len(forms)](v) for j, v in enumerate(p)] for i, (f, p) in enumerate(zip(fr, pts))]
    elif variant == "bad_rows":
        rows = rows[:5] + [["1", "abc", "2.0", "0.5"], ["-1", "1.0", "2.0", "0.5"], ["1", "1.0", "2.0"],
                           ["1", "1.0", "2.0", "0.5", "maybe"], ["1.5", "1.0", "2.0", "0.5"], ["1", "", "2.0", "0.5"],
                           ["99999999999", "1.0", "2.0", "0.5"]] + rows[5:]
    elif variant == "no_trailing_newline":
        d = _write_case(tmp_path / variant, rows)
        for phase in ("diastolic", "systolic"):
            p = d / f"{phase}_contours.csv"
            p.write_text(p.read_text().rstrip("\n"))
    d = _write_case(tmp_path / variant, rows, **kw) if variant != "no_trailing_newline" else tmp_path / variant
    got = nat.geometry_from_dir(d, "x", True)
    want = ora.build_geometry_from_dir(d, "x", True)
    assert np.array_equal(got, want)
    g = mm.PyGeometry.from_blob(got, "x")
    assert len(g.frames) == 3 and all(len(f.lumen) == n for f in g.frames)


def test_contour_reader_fuzz_matches_oracle(tmp_path):
    """Seeded fuzz of the in-place contour-file reader against the oracle's line-by-line reader: a clean 3-frame file
    with a few rows corrupted (odd number tokens, extra / missing fields, stray CRs and blanks, inserted junk rows, a
    header). Both must build the same geometry blob or fail with the same message."""
    import math
    import random
    rnd = random.Random(12345)
    tok = ["1", "0", "12", "007", "+3", "-1", "1.5", "-2.25e0", "3E0", ".5", "5.", "1e+0", "1e-0", "inf", "nan", "NaN",
           "-inf", "0x10", "1_0", "abc", "", " 2 ", "2 ", "\t2", "1,5", "true", "false", "TRUE", "4294967295",
           "4294967296", "99999999999999999999", "1e400", "1e-400", "-0", "-0.0", "١"]
    parsed = 0
    for trial in range(150):
        d = tmp_path / f"t{trial}"
        d.mkdir()
        delim, eol, n = rnd.choice(["\t", ","]), rnd.choice(["\n", "\r\n"]), 8
        rows = [[str(f), repr(3 + 2 * math.cos(2 * math.pi * k / n)), repr(3 + 1.5 * math.sin(2 * math.pi * k / n)),
                 repr(0.5 * f)] for f in range(3) for k in range(n)]
        for _ in range(rnd.randint(0, 3)):
            i = rnd.randrange(len(rows))
            r = list(rows[i])
            mode = rnd.randrange(5)
            if len(r) < 4:
                continue
            if mode == 0:
                r[rnd.randrange(4)] = rnd.choice(tok)
            elif mode == 1:
                r.append(rnd.choice(tok))
            elif mode == 2:
                r = r[:rnd.randint(0, 3)]
            elif mode == 3:
                r[rnd.randrange(4)] = " " + r[rnd.randrange(4)] + rnd.choice([" ", "\r", ""])
            else:
                rows.insert(i, [rnd.choice(tok) for _ in range(rnd.randint(1, 6))])
                continue
            rows[i] = r
        if rnd.random() < 0.3:
            rows.insert(rnd.randrange(len(rows)), [rnd.choice(["", " ", "\r"])])
        if rnd.random() < 0.2:
            rows.insert(0, ["frame", "x", "y", "z"])
        for ph in ("diastolic", "systolic"):
            with open(d / f"{ph}_contours.csv", "w", newline="") as f:
                f.write(eol.join(delim.join(r) for r in rows) + (eol if rnd.random() < 0.8 else ""))
            (d / f"{ph}_reference_points.csv").write_text(delim.join(["1", "3.5", "2.0", "0.5"]) + "\n")
        got = []
        for fn in (nat.geometry_from_dir, ora.build_geometry_from_dir):
            try:
                got.append(("ok", fn(d, "x", True)))
            except (nat.MmrsError, ora.OracleError) as e:
                got.append(("err", str(e)))
        (ka, a), (kb, b) = got
        assert ka == kb, (trial, a if ka == "err" else "blob", b if kb == "err" else "blob")
        assert np.array_equal(a, b, equal_nan=True) if ka == "ok" else a == b, trial
        parsed += ka == "ok"
    assert 30 < parsed < 140          # the fuzz exercises both outcomes


def test_array_ingest_row_orders_and_parallel_paths_match_oracle():
    """mmrs_geometry_from_arrays borrows the caller's rows, groups them by runs of equal frame ids and, for big inputs,
    fills / sorts / encodes frame-parallel. Against the oracle's plain restatement: rows in frame order, fully shuffled,
    interleaved in three passes and reversed; with and without the optional layers and records; small inputs (serial
    paths) and two that cross the thresholds of the parallel paths (>= 32 768 points, >= 131 072 blob doubles)."""
    rng = np.random.default_rng(99)

    def layers(nf, npnt):
        out = {}
        for name, scale in (("lumen", 1.0), ("eem", 1.4), ("calc", 0.5), ("side", 0.3)):
            rows = []
            for f in range(nf):
                ang = np.sort(rng.uniform(0, 2 * np.pi, npnt))
                r = scale * (2 + 0.3 * np.cos(2 * ang))
                rows.append(np.column_stack([np.full(npnt, 10 + f), 4.5 + r * np.cos(ang) + rng.normal(0, 0.01, npnt),
                                             4.5 + r * np.sin(ang) + rng.normal(0, 0.01, npnt),
                                             np.full(npnt, 0.5 * (nf - 1 - f))]))
            out[name] = np.concatenate(rows)
        return out

    for nf, npnt in ((3, 12), (7, 33), (4, 64), (80, 500), (9, 17), (300, 150)):
        base = layers(nf, npnt)
        for variant in range(4):
            arrs = {}
            for k, a in base.items():
                arrs[k] = (a, a[rng.permutation(len(a))], np.concatenate([a[i::3] for i in range(3)]), a[::-1].copy())[variant]
            ref = np.array([10 + nf - 1, 6.5, 4.5, 0.0])
            rec = np.array([[10 + f, 1.0, 1.0 + f, np.nan] for f in rng.permutation(nf)]) if variant % 2 else None
            args = (arrs["lumen"], ref, arrs["eem"] if variant != 3 else None, arrs["calc"] if variant in (0, 2) else None,
                    arrs["side"] if variant == 0 else None, rec, True, "x")
            got = nat.geometry_from_arrays(*args)
            want = ora.build_geometry_from_arrays(*args)
            assert np.array_equal(got, want), (nf, npnt, variant)
            assert np.array_equal(mm.PyGeometry.from_blob(got, "x").to_blob(), got)      # Python codec round trip


def test_host_pool_concurrent_submitters_nested_jobs_and_errors(tmp_path):
    """The library's persistent host pool (csrc/mmrs_host.cpp HostPool): several Python threads ingest big pullbacks at
    once (each call runs frame-parallel jobs, nested under the caller), the blobs equal the single-threaded ones byte
    for byte, an ingest error raised on a pool thread reaches its caller only, and the pool keeps working afterwards."""
    import threading

    rng = np.random.default_rng(5)

    def pullback(nf, npnt, seed):
        r = np.random.default_rng(seed)
        rows = []
        for f in range(nf):
            ang = np.sort(r.uniform(0, 2 * np.pi, npnt))
            rad = 2 + 0.3 * np.cos(2 * ang)
            rows.append(np.column_stack([np.full(npnt, f), 4.5 + rad * np.cos(ang), 4.5 + rad * np.sin(ang),
                                         np.full(npnt, 0.5 * (nf - 1 - f))]))
        a = np.concatenate(rows)
        return a[r.permutation(len(a))], np.array([nf - 1, 6.5, 4.5, 0.0])

    cases = [pullback(120, 400, 100 + k) for k in range(6)] + [pullback(5, 20, 200 + k) for k in range(6)]
    want = [np.array(nat.geometry_from_arrays(a, rp, diastole=True, label="p")) for a, rp in cases]
    got, errors = [None] * len(cases), []

    def worker(k):
        try:
            for _ in range(3):
                got[k] = np.array(nat.geometry_from_arrays(*cases[k], diastole=True, label="p"))
            if k % 4 == 0:   # an error inside a call: no lumen rows at all
                try:
                    nat.geometry_from_arrays(np.zeros((0, 4)), cases[k][1], diastole=True, label="bad")
                    errors.append((k, "no error raised"))
                except Exception:   # noqa: BLE001  (the message is pinned in test_ingest_errors_*)
                    pass
        except Exception as e:   # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not any(t.is_alive() for t in threads), "a pool job never completed"
    assert not errors, errors
    for w, g in zip(want, got):
        assert g is not None and np.array_equal(w, g, equal_nan=True)
    again = np.array(nat.geometry_from_arrays(*cases[0], diastole=True, label="p"))
    assert np.array_equal(want[0], again, equal_nan=True)
