"""OBJ / MTL / PNG export (mmrs_export_pair, mmrs_export_single) against the ORACLE's restatement of the reference's
writers (oracle/export_py.py: io/output.rs:10-181 write_obj_mesh, texture.rs UV coordinates, write_mtl.rs bodies) and,
restated here from their format strings, to_object/process.rs:9-121,
to_object/interpolation.rs:9-157, to_object/write_mtl.rs:15-273, to_object/texture.rs:6-95,
binding/entry.rs:741-818. Host-only: no GPU needed."""
import math
import os
import struct
import zlib

import numpy as np
import pytest

from multimodars import PyContour, PyContourPoint, PyFrame, PyGeometry
from multimodars import _native as nat
from oracle.export_py import obj_text as expected_obj
from oracle.export_py import mtl_text, rust_f64, uv_coords  # noqa: F401


def ring(fid, cz, r, n, cx=0.0, cy=0.0, kind="Lumen", phase=0.0):
    pts = [PyContourPoint(fid, i, cx + r * math.cos(phase + 2 * math.pi * i / n), cy + r * math.sin(phase + 2 * math.pi * i / n),
                          cz, False) for i in range(n)]
    c = (sum(p.x for p in pts) / n, sum(p.y for p in pts) / n, sum(p.z for p in pts) / n)
    return PyContour(fid, fid, pts, c, None, None, kind)


def geometry(label, n_frames=3, n=8, r0=2.0, dz=0.5, grow=0.0, with_catheter=True):
    frames = []
    for f in range(n_frames):
        lum = ring(f, f * dz, r0 + grow * f, n)
        extras = {"Catheter": ring(f, f * dz, 0.5, 4, kind="Catheter")} if with_catheter else {}
        ref = PyContourPoint(f, 1, 1.0, 2.0, f * dz, False) if f == n_frames - 1 else None
        frames.append(PyFrame(f, lum.centroid, lum, extras, ref))
    return PyGeometry(frames, label)


def read_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, hdr = 8, b"", None
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        crc = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0]
        assert crc == (zlib.crc32(typ + body) & 0xffffffff)
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat += body
        pos += 12 + n
    w, h, depth, ctype = hdr[:4]
    ch = {2: 3, 6: 4}[ctype]
    raw = zlib.decompress(idat)
    rows = np.frombuffer(raw, dtype=np.uint8).reshape(h, 1 + w * ch)
    assert depth == 8 and (rows[:, 0] == 0).all()
    return rows[:, 1:].reshape(h, w, ch)


def test_single_export_matches_reference_format(tmp_path):
    g = geometry("rest", n_frames=3, n=8)
    # values that exercise Rust's float formatting: integer-valued, tiny, negative zero
    g.frames[0].lumen.points[0].x = 1.0
    g.frames[0].lumen.points[1].x = 1e-7
    g.frames[0].lumen.points[2].y = -0.0
    g.frames[0].lumen.points[3].x = 1e21
    nat.export_single(g.to_blob(), "rest", str(tmp_path), True, [0, 4, 5], 0)
    assert sorted(os.listdir(tmp_path)) == ["catheter_rest.mtl", "catheter_rest.obj", "lumen_rest.mtl", "lumen_rest.obj"]
    lum = [f.lumen for f in g.frames]
    mtl = str(tmp_path / "lumen_rest.mtl")
    want = expected_obj(lum, [(0.0, 0.0)] * 24, mtl, True)
    got = open(tmp_path / "lumen_rest.obj").read()
    assert got == want
    assert "v 1 " in got and "v 0.0000001 " in got and " -0 " in got and "v 1000000000000000000000 " in got
    assert open(mtl).read() == mtl_text("lumen") == "newmtl material\nKa 1.0 1.0 1.0\nKd 1.0 1.0 1.0\nKs 0.0 0.0 0.0\n"
    assert open(tmp_path / "catheter_rest.mtl").read() == mtl_text("catheter")
    # naming 1 = to_object::write_single_geometry ("{case}_{type}")
    nat.export_single(g.to_blob(), "case", str(tmp_path / "w"), False, [0], 1)
    assert sorted(os.listdir(tmp_path / "w")) == ["case_lumen.mtl", "case_lumen.obj"]
    assert open(tmp_path / "w" / "case_lumen.obj").read() == expected_obj(lum, [(0.0, 0.0)] * 24,
                                                                        str(tmp_path / "w" / "case_lumen.mtl"), False)


def test_pair_export_files_textures_and_interpolation(tmp_path):
    a = geometry("dia", n_frames=3, n=8, r0=2.0)
    b = geometry("sys", n_frames=3, n=8, r0=2.0, grow=0.25)  # frame f is 0.25*f larger
    steps = 3
    nat.export_pair(a.to_blob(), b.to_blob(), "dia", "dia - sys", str(tmp_path), steps, True, [0, 4])
    names = sorted(os.listdir(tmp_path))
    want_names = sorted(f"{t}_{i:03d}_dia - sys.{ext}" for t in ("lumen", "catheter") for i in range(steps + 2)
                        for ext in ("obj", "mtl", "png"))
    assert names == want_names
    # interpolation t = step / (steps - 1), geometry list = [start, interp..., end] (interpolation.rs:9-94)
    ts = [None] + [s / (steps - 1) for s in range(steps)] + [None]
    for i, t in enumerate(ts):
        if t is None:
            src = a if i == 0 else b
            lum = [f.lumen for f in src.frames]
        else:
            lum = []
            for fa, fb in zip(a.frames, b.frames):
                pts = [PyContourPoint(p.frame_index, p.point_index, p.x * (1.0 - t) + q.x * t, p.y * (1.0 - t) + q.y * t,
                                      p.z * (1.0 - t) + q.z * t, p.aortic) for p, q in zip(fa.lumen.points, fb.lumen.points)]
                c = tuple(u * (1.0 - t) + v * t for u, v in zip(fa.lumen.centroid, fb.lumen.centroid))
                lum.append(PyContour(fa.lumen.id, fa.lumen.original_frame, pts, c, None, None, "Lumen"))
        uv = uv_coords(3, 8)
        got = open(tmp_path / f"lumen_{i:03d}_dia - sys.obj").read()
        assert got == expected_obj(lum, uv, f"lumen_{i:03d}_dia - sys.mtl", True)
        assert open(tmp_path / f"lumen_{i:03d}_dia - sys.mtl").read() == mtl_text("lumen", f"lumen_{i:03d}_dia - sys.png")
        assert open(tmp_path / f"catheter_{i:03d}_dia - sys.mtl").read() == mtl_text("catheter", f"catheter_{i:03d}_dia - sys.png")
        # displacement texture: texture.rs:51-74, normalised by the first-to-last max displacement
        img = read_png(tmp_path / f"lumen_{i:03d}_dia - sys.png")
        assert img.shape == (3, 8, 3)
        max_disp = 0.5  # frame 2 grows by 0.5
        for f in range(3):
            for p in range(8):
                pa, pl = a.frames[f].lumen.points[p], lum[f].points[p]
                d = math.sqrt((pl.x - pa.x) ** 2 + (pl.y - pa.y) ** 2 + (pl.z - pa.z) ** 2)
                nrm = min(max(d / max_disp, 0.0), 1.0)
                want = (int(nrm * 255.0), 0, int((1.0 - nrm) * 255.0))
                px = tuple(int(v) for v in img[(3 - 1) - f, p])
                # the displacement is recomputed in f64 from the written coordinates; allow the last-ulp flip
                assert all(abs(x - y) <= 1 for x, y in zip(px, want)), (i, f, p, px, want)
        cat = read_png(tmp_path / f"catheter_{i:03d}_dia - sys.png")
        assert cat.shape == (3, 4, 3) and not cat.any()


def test_wall_texture_is_rgba_with_reference_alpha(tmp_path):
    a = geometry("a", n_frames=2, n=6, with_catheter=False)
    for f in a.frames:
        f.extras["Wall"] = ring(f.id, f.centroid[2], 3.0, 6, kind="Wall")
    nat.export_pair(a.to_blob(), a.to_blob(), "a", "pair", str(tmp_path), 0, False, [5])
    img = read_png(tmp_path / "wall_000_pair.png")
    assert img.shape == (2, 6, 4)
    assert (img[..., :3] == 0).all() and (img[..., 3] == int(255.0 - 0.7 * 255.0)).all()  # texture.rs:84-95
    assert open(tmp_path / "wall_001_pair.mtl").read() == (
        "newmtl transparent_material\nKa 0 0 0\nKd 0 0 0\nmap_Kd wall_001_pair.png\n")
    assert "f " in open(tmp_path / "wall_000_pair.obj").read()


def test_export_errors_follow_the_reference(tmp_path):
    one = geometry("one", n_frames=1)
    with pytest.raises(nat.MmrsError, match="Need at least two contours to create a mesh"):
        nat.export_single(one.to_blob(), "one", str(tmp_path), True, [0], 0)
    g = geometry("g", n_frames=2, with_catheter=False)
    # a requested type that no frame carries: the pair export fails like write_geometry_vec_to_obj does
    with pytest.raises(nat.MmrsError, match=r"Some \.obj writes failed"):
        nat.export_pair(g.to_blob(), g.to_blob(), "g", "pair", str(tmp_path / "p"), 0, True, [0, 5])
    # ... the single export only warns and skips it (entry.rs:748-751)
    nat.export_single(g.to_blob(), "g", str(tmp_path / "s"), True, [0, 5], 0)
    assert sorted(os.listdir(tmp_path / "s")) == ["lumen_g.mtl", "lumen_g.obj"]
    uneven = geometry("u", n_frames=2, n=8)
    uneven.frames[1].lumen = ring(1, 0.5, 2.0, 6)
    with pytest.raises(nat.MmrsError, match="All contours must have the same number of points"):
        nat.export_single(uneven.to_blob(), "u", str(tmp_path / "u"), True, [0], 0)


def test_malformed_blobs_are_refused(tmp_path):
    """The blob decoder (csrc/mmrs_host.cpp decode) never trusts the counts inside a blob: a frame count the buffer cannot
    hold, a cut-off contour, an empty buffer are errors of the call."""
    import numpy as np
    from multimodars import _native as nat
    for bad in (np.array([1e15, 0.0, 0.0, 0.0]), np.array([2.0] + [0.0] * 30), np.zeros(0)):
        with pytest.raises(nat.MmrsError, match="truncated|empty"):
            nat.export_single(bad, "x", tmp_path, True, [0], 0)
    frame = np.concatenate([np.zeros(11), [1.0], np.zeros(11), [600.0], np.zeros(6 * 600)])   # one 600-point lumen
    big = np.concatenate([[40.0], np.tile(frame, 40)])
    assert len(big) > (1 << 17)                                # large enough for the frame-parallel decoder
    with pytest.raises(nat.MmrsError, match="truncated"):     # ... whose header walk notices the missing tail
        nat.export_single(big[:-100], "x", tmp_path, True, [0], 0)


def test_large_geometry_takes_the_frame_parallel_decoder(tmp_path):
    """A blob above the threshold of the frame-parallel decoder (csrc/mmrs_host.cpp decode: >= 131 072 doubles and
    >= 16 frames) exports exactly what the Python restatement of the format expects — frame order, contours and points
    intact."""
    g = geometry("big", n_frames=48, n=500, with_catheter=True)
    blob = g.to_blob()
    assert len(blob) >= (1 << 17) and len(g.frames) >= 16
    nat.export_single(blob, "big", str(tmp_path), True, [0, 4], 0)
    for kind, name in (("Lumen", "lumen"), ("Catheter", "catheter")):
        cs = [f.lumen if kind == "Lumen" else f.extras[kind] for f in g.frames]
        want = expected_obj(cs, [(0.0, 0.0)] * sum(len(c) for c in cs), str(tmp_path / f"{name}_big.mtl"), True)
        assert open(tmp_path / f"{name}_big.obj").read() == want
