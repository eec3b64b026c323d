"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
reference) may import this module. The product package never does.
See oracle/mmrs_oracle.hpp for scope, parity status and reference citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libmmrs_oracle.so"
_lib = None

c_dp = C.POINTER(C.c_double)
c_lp = C.POINTER(C.c_long)


def build(force: bool = False) -> Path:
    """Compile oracle/libmmrs_oracle.so with the committed Makefile (g++ only)."""
    srcs = [_HERE / "mmrs_oracle_capi.cpp", _HERE / "mmrs_oracle.hpp"]
    stale = (not _LIB_PATH.exists()) or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libmmrs_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.ora_last_error.restype = C.c_char_p
        _lib.ora_hausdorff.restype = C.c_double
        _lib.ora_directed_hausdorff.restype = C.c_double
        _lib.ora_search_grid.restype = C.c_long
        _lib.ora_search_range_analytic.restype = C.c_double
        _lib.ora_downsample_indices.restype = C.c_long
        _lib.ora_free.argtypes = [C.c_void_p]
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(lib().ora_last_error().decode())


def _xy(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, 2))
    return a, a.ctypes.data_as(c_dp), C.c_long(a.shape[0])


def hausdorff(a, b) -> float:
    a, pa, na = _xy(a)
    b, pb, nb = _xy(b)
    return lib().ora_hausdorff(pa, na, pb, nb)


def directed_hausdorff(a, b) -> float:
    a, pa, na = _xy(a)
    b, pb, nb = _xy(b)
    return lib().ora_directed_hausdorff(pa, na, pb, nb)


def search_grid(step_deg, range_deg, center=None, limes_deg=None, cap=1 << 20):
    """Returns (angles ndarray | None, fallback). None = search_range returns early."""
    limes_deg = range_deg if limes_deg is None else limes_deg
    out = np.empty(cap, dtype=np.float64)
    fb = C.c_double(0.0)
    n = lib().ora_search_grid(C.c_double(step_deg), C.c_double(range_deg), int(center is not None),
                              C.c_double(0.0 if center is None else center), C.c_double(limes_deg),
                              out.ctypes.data_as(c_dp), C.c_long(cap), C.byref(fb))
    if n < 0:
        return None, fb.value
    assert n <= cap
    return out[:n].copy(), fb.value


def search_range_analytic(kind, param, step_deg, range_deg, center, limes_deg, threads=1) -> float:
    return lib().ora_search_range_analytic(int(kind), C.c_double(param), C.c_double(step_deg), C.c_double(range_deg),
                                           int(center is not None), C.c_double(0.0 if center is None else center),
                                           C.c_double(limes_deg), int(threads))


def sweep(test, ref, centre, mode, step_deg, range_deg, center=None, limes_deg=None, threads=1, want_costs=True):
    """One search_range call with the Hausdorff closure. Returns dict(angle, index, cost, costs)."""
    limes_deg = range_deg if limes_deg is None else limes_deg
    t, pt, nt = _xy(test)
    r, pr, nr = _xy(ref)
    ang, idx, cost = C.c_double(), C.c_long(), C.c_double()
    cap = 0
    costs = None
    pc = None
    if want_costs:
        g, _ = search_grid(step_deg, range_deg, center, limes_deg)
        cap = 0 if g is None else len(g)
        costs = np.empty(cap, dtype=np.float64)
        pc = costs.ctypes.data_as(c_dp)
    _check(lib().ora_sweep(pt, nt, pr, nr, C.c_double(centre[0]), C.c_double(centre[1]), int(mode),
                           C.c_double(step_deg), C.c_double(range_deg), int(center is not None),
                           C.c_double(0.0 if center is None else center), C.c_double(limes_deg), int(threads),
                           C.byref(ang), C.byref(idx), C.byref(cost), pc, C.c_long(cap)))
    return dict(angle=ang.value, index=idx.value, cost=cost.value, costs=costs)


def costs(test, ref, centre, mode, angles):
    t, pt, nt = _xy(test)
    r, pr, nr = _xy(ref)
    a = np.ascontiguousarray(angles, dtype=np.float64)
    out = np.empty(len(a), dtype=np.float64)
    _check(lib().ora_costs(pt, nt, pr, nr, C.c_double(centre[0]), C.c_double(centre[1]), int(mode),
                           a.ctypes.data_as(c_dp), C.c_long(len(a)), out.ctypes.data_as(c_dp)))
    return out


def find_best_rotation(test, ref, centre, mode, step_deg, range_deg, threads=1) -> float:
    t, pt, nt = _xy(test)
    r, pr, nr = _xy(ref)
    ang = C.c_double()
    _check(lib().ora_find_best_rotation(pt, nt, pr, nr, C.c_double(centre[0]), C.c_double(centre[1]), int(mode),
                                        C.c_double(step_deg), C.c_double(range_deg), int(threads), C.byref(ang)))
    return ang.value


def downsample_indices(length, n):
    out = np.empty(max(length, n, 1), dtype=np.int64)
    k = lib().ora_downsample_indices(C.c_long(length), C.c_long(n), out.ctypes.data_as(c_lp))
    return out[:k].copy()


def sweep_batch(test_xy, test_off, ref_xy, ref_off, centre_xy, mode, step_deg, range_deg, limes_deg, threads=1):
    test_xy = np.ascontiguousarray(test_xy, dtype=np.float64)
    ref_xy = np.ascontiguousarray(ref_xy, dtype=np.float64)
    test_off = np.ascontiguousarray(test_off, dtype=np.int64)
    ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
    centre_xy = np.ascontiguousarray(centre_xy, dtype=np.float64)
    U = len(test_off) - 1
    bi = np.empty(U, dtype=np.int64)
    bc = np.empty(U, dtype=np.float64)
    _check(lib().ora_sweep_batch(test_xy.ctypes.data_as(c_dp), test_off.ctypes.data_as(c_lp),
                                 ref_xy.ctypes.data_as(c_dp), ref_off.ctypes.data_as(c_lp),
                                 centre_xy.ctypes.data_as(c_dp), C.c_long(U), int(mode), C.c_double(step_deg),
                                 C.c_double(range_deg), C.c_double(limes_deg), int(threads),
                                 bi.ctypes.data_as(c_lp), bc.ctypes.data_as(c_dp)))
    return bi, bc


# ---- geometry blobs ----------------------------------------------------------
def _take(pp, n):
    arr = np.ctypeslib.as_array(pp, shape=(max(n, 1),))[:n].copy()
    lib().ora_free(pp)
    return arr


def decode_geometry(blob):
    """f64 blob -> list of frame dicts (layout: include/mmrs_b200.h, 'geometry blob')."""
    b = np.asarray(blob, dtype=np.float64)
    pos = 0

    def nxt(k=1):
        nonlocal pos
        v = b[pos:pos + k]
        pos += k
        return v

    frames = []
    nf = int(nxt()[0])
    for _ in range(nf):
        fid, cx, cy, cz, has_ref = nxt(5)
        rp = nxt(6)
        nc = int(nxt()[0])
        contours = {}
        for k in range(nc):
            h = nxt(12)
            npts = int(h[11])
            pts = nxt(6 * npts).reshape(npts, 6)
            contours[int(h[0])] = dict(kind=int(h[0]), id=int(h[1]), original_frame=int(h[2]),
                                       centroid=tuple(h[4:7]) if h[3] else None,
                                       aortic_thickness=h[8] if h[7] else None,
                                       pulmonary_thickness=h[10] if h[9] else None, points=pts)
        frames.append(dict(id=int(fid), centroid=(cx, cy, cz), reference_point=rp.copy() if has_ref else None,
                           contours=contours))
    assert pos == len(b), "blob not fully consumed"
    return frames


def encode_geometry(frames):
    out = [float(len(frames))]
    for f in frames:
        out += [float(f["id"]), *map(float, f["centroid"]), 1.0 if f["reference_point"] is not None else 0.0]
        rp = f["reference_point"] if f["reference_point"] is not None else np.zeros(6)
        out += list(map(float, rp))
        cs = f["contours"]
        out.append(float(len(cs)))
        for kind in sorted(cs, key=lambda k: (k != 0, k)):
            c = cs[kind]
            cen = c["centroid"]
            out += [float(c["kind"]), float(c["id"]), float(c["original_frame"]), 1.0 if cen is not None else 0.0,
                    *(map(float, cen) if cen is not None else (0.0, 0.0, 0.0)),
                    1.0 if c["aortic_thickness"] is not None else 0.0, float(c["aortic_thickness"] or 0.0),
                    1.0 if c["pulmonary_thickness"] is not None else 0.0, float(c["pulmonary_thickness"] or 0.0),
                    float(len(c["points"]))]
            out += list(np.asarray(c["points"], dtype=np.float64).reshape(-1))
    return np.asarray(out, dtype=np.float64)


def build_geometry_from_dir(path, label="geom", diastole=True, image_center=(4.5, 4.5), radius=0.5, n_points=20):
    blob, ln = c_dp(), C.c_long()
    _check(lib().ora_build_geometry_from_dir(os.fsencode(str(path)), label.encode(), int(diastole),
                                             C.c_double(image_center[0]), C.c_double(image_center[1]),
                                             C.c_double(radius), C.c_uint(n_points), C.byref(blob), C.byref(ln)))
    return _take(blob, ln.value)


def build_geometry_from_arrays(lumen, ref_point, eem=None, calc=None, side=None, records=None, diastole=True,
                               label="geom", image_center=(4.5, 4.5), radius=0.5, n_points=20):
    def arr(a):
        if a is None:
            return None, None, C.c_long(0)
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, 4))
        return a, a.ctypes.data_as(c_dp), C.c_long(a.shape[0])

    l, pl, nl = arr(lumen)
    e, pe, ne = arr(eem)
    c, pc, nc = arr(calc)
    s, ps, ns = arr(side)
    r, prr, nr = arr(records)
    rp = np.ascontiguousarray(np.asarray(ref_point, dtype=np.float64).reshape(4))
    blob, ln = c_dp(), C.c_long()
    _check(lib().ora_build_geometry_from_arrays(pl, nl, pe, ne, pc, nc, ps, ns, prr, nr, rp.ctypes.data_as(c_dp),
                                                int(diastole), label.encode(), C.c_double(image_center[0]),
                                                C.c_double(image_center[1]), C.c_double(radius), C.c_uint(n_points),
                                                C.byref(blob), C.byref(ln)))
    return _take(blob, ln.value)


def align_within(blob, step_deg, range_deg, smooth, bruteforce, sample_size, threads=1, post_steps=True):
    b = np.ascontiguousarray(blob, dtype=np.float64)
    ob, ol, lg, nl, an = c_dp(), C.c_long(), c_dp(), C.c_long(), C.c_int()
    _check(lib().ora_align_within(b.ctypes.data_as(c_dp), C.c_long(len(b)), C.c_double(step_deg),
                                  C.c_double(range_deg), int(smooth), int(bruteforce), C.c_long(sample_size),
                                  int(threads), int(post_steps), C.byref(ob), C.byref(ol), C.byref(lg), C.byref(nl),
                                  C.byref(an)))
    return _take(ob, ol.value), _take(lg, nl.value * 7).reshape(-1, 7), bool(an.value)


def align_between(blob_a, blob_b, rot_deg, step_deg, sample_size, threads=1):
    a = np.ascontiguousarray(blob_a, dtype=np.float64)
    b = np.ascontiguousarray(blob_b, dtype=np.float64)
    ob, ol, ang = c_dp(), C.c_long(), C.c_double()
    _check(lib().ora_align_between(a.ctypes.data_as(c_dp), C.c_long(len(a)), b.ctypes.data_as(c_dp),
                                   C.c_long(len(b)), C.c_double(rot_deg), C.c_double(step_deg),
                                   C.c_long(sample_size), int(threads), C.byref(ob), C.byref(ol), C.byref(ang)))
    return _take(ob, ol.value), ang.value


def process(mode, blobs, step_deg, range_deg, smooth, bruteforce, sample_size, threads=1, postprocessing=False):
    """mode 4 full / 3 double pair / 2 single pair / 1 single. Returns (out_blobs, logs)."""
    n_in = 4 if mode >= 3 else mode
    n_out = {4: 8, 3: 4, 2: 2, 1: 1}[mode]
    assert len(blobs) == n_in
    arrs = [np.ascontiguousarray(b, dtype=np.float64) for b in blobs]
    ptrs = (c_dp * n_in)(*[a.ctypes.data_as(c_dp) for a in arrs])
    lens = (C.c_long * n_in)(*[len(a) for a in arrs])
    ob = (c_dp * n_out)()
    ol = (C.c_long * n_out)()
    lg = (c_dp * n_in)()
    nl = (C.c_long * n_in)()
    _check(lib().ora_process(int(mode), ptrs, lens, C.c_double(step_deg), C.c_double(range_deg), int(smooth),
                             int(bruteforce), C.c_long(sample_size), int(threads), int(postprocessing), ob, ol, lg, nl))
    outs = [_take(ob[i], ol[i]) for i in range(n_out)]
    logs = [_take(lg[i], nl[i] * 7).reshape(-1, 7) for i in range(n_in)]
    return outs, logs


def postprocess_pair(blob_a, blob_b, tol=0.03, anomalous=False):
    a = np.ascontiguousarray(blob_a, dtype=np.float64)
    b = np.ascontiguousarray(blob_b, dtype=np.float64)
    oa, la, ob, lb = c_dp(), C.c_long(), c_dp(), C.c_long()
    _check(lib().ora_postprocess_pair(a.ctypes.data_as(c_dp), C.c_long(len(a)), b.ctypes.data_as(c_dp), C.c_long(len(b)),
                                      C.c_double(tol), int(anomalous), C.byref(oa), C.byref(la), C.byref(ob),
                                      C.byref(lb)))
    return _take(oa, la.value), _take(ob, lb.value)


def predict_z_positions(ref_z, start_z, stop_z, z_diff, cap=100000):
    out = np.empty(cap, dtype=np.float64)
    lib().ora_predict_z_positions.restype = C.c_long
    n = lib().ora_predict_z_positions(C.c_double(ref_z), C.c_double(start_z), C.c_double(stop_z), C.c_double(z_diff),
                                      out.ctypes.data_as(c_dp), C.c_long(cap))
    return out[:n].copy()


def within_chain_tap(blob, step_deg, range_deg, bruteforce, sample_size, pair, threads=1):
    """Chain-state (test_xy, ref_xy, centre, best_angle) of frame pair `pair` as the reference's closure sees them."""
    b = np.ascontiguousarray(blob, dtype=np.float64)
    t, nt, r, nr, best = c_dp(), C.c_long(), c_dp(), C.c_long(), C.c_double()
    cen = (C.c_double * 2)()
    _check(lib().ora_within_chain_tap(b.ctypes.data_as(c_dp), C.c_long(len(b)), C.c_double(step_deg), C.c_double(range_deg),
                                      int(bruteforce), C.c_long(sample_size), int(threads), C.c_long(pair), C.byref(t),
                                      C.byref(nt), C.byref(r), C.byref(nr), cen, C.byref(best)))
    return (_take(t, 2 * nt.value).reshape(-1, 2), _take(r, 2 * nr.value).reshape(-1, 2), (cen[0], cen[1]), best.value)
