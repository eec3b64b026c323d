"""Pins the oracle's restatement of processing/postprocessing.rs on that file's own Rust tests
(postprocessing.rs:474-979)."""
import numpy as np
import pytest

from oracle import oracle_py as ora


def contour(cid, z, thickness, kind):  # create_test_contour, :482-513
    pts = np.array([[cid, 0, 1.0, 2.0, z, 0.0], [cid, 1, 3.0, 4.0, z, 0.0]])
    return dict(kind=kind, id=cid, original_frame=cid, centroid=(2.0, 3.0, z), aortic_thickness=thickness,
                pulmonary_thickness=None, points=pts)


def frame(fid, z, thickness, set_ref):  # create_test_frame, :516-544
    rp = np.array([fid, 0, 0.0, 0.0, z, 0.0]) if set_ref else None
    return dict(id=fid, centroid=(2.0, 3.0, z), reference_point=rp,
                contours={0: contour(fid, z, thickness, 0), 1: contour(fid, z, None, 1)})


def geometry(zs, thick=()):  # create_test_geometry, :547-579 (reference point on the middle frame)
    return [frame(i, z, thick[i] if i < len(thick) else None, i == len(zs) // 2) for i, z in enumerate(zs)]


def post(a, b, tol=0.1, anomalous=False):
    oa, ob = ora.postprocess_pair(ora.encode_geometry(a), ora.encode_geometry(b), tol, anomalous)
    return ora.decode_geometry(oa), ora.decode_geometry(ob)


def test_predict_z_positions_forward():  # :678-685
    assert ora.predict_z_positions(0.0, 0.0, 5.0, 1.0).tolist() == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]


def test_predict_z_positions_backward_and_middle():  # :688-712
    z = ora.predict_z_positions(5.0, 0.0, 5.0, 1.0)
    assert len(z) and 5.0 in z
    z = ora.predict_z_positions(2.5, 0.0, 5.0, 1.0)
    assert 2.5 in z and (z <= 1.0).any() and (z >= 4.0).any()
    assert z.tolist() == [0.5, 1.5, 2.5, 3.5, 4.5]


def test_predict_z_positions_complex_resampling():  # :930-951
    z = ora.predict_z_positions(1.5, 0.0, 2.5, 0.5)
    assert z.tolist() == [0.0, 0.5, 1.0, 1.5, 2.0, 2.5]
    assert ora.predict_z_positions(0.0, 0.0, 5.0, 0.0).tolist() == []


def test_postprocess_same_rate_resamples_and_trims():  # resample_by_diff :651-664, trim :787-814
    a = geometry([0.0, 1.0, 2.0, 3.0, 4.0])
    b = geometry([0.0, 1.0, 2.0])
    fa, fb = post(a, b)
    assert len(fa) == 3 and len(fb) == 3
    assert [f["id"] for f in fa] == [0, 1, 2] and [f["id"] for f in fb] == [0, 1, 2]
    assert [f["centroid"][2] for f in fb] == [0.0, 1.0, 2.0]


def test_postprocess_different_rates_interpolates():  # new_frames_by_sample_rate :715-737
    a = geometry([0.0, 1.0, 2.0, 3.0, 4.0], [1.0] * 5)
    b = geometry([0.0, 2.0, 4.0, 6.0, 8.0], [2.0] * 5)
    fa, fb = post(a, b)                       # signed test (da - db) = -1 < tol -> "same rate" branch (:93)
    assert len(fa) and len(fb)
    fa2, fb2 = post(b, a)                     # (2 - 1) >= tol and da > db -> A is re-sampled at B's spacing
    assert len(fa2) and len(fb2)
    z = [f["centroid"][2] for f in fa2]
    assert all(abs((z[i + 1] - z[i]) - 1.0) < 1e-12 for i in range(len(z) - 1))


def test_adjust_walls_anomalous_averages_thickness():  # :817-861
    a = geometry([0.0, 1.0], [1.0, 2.0])
    b = geometry([0.0, 1.0], [3.0, 4.0])
    # give the 2-point test contours enough points for create_aortic_wall? they have 2 -> wall creation
    # needs >= 4 points; use 8-point contours instead
    def fat(g):
        for f in g:
            for c in f["contours"].values():
                z = f["centroid"][2]
                th = np.linspace(0, 2 * np.pi, 8, endpoint=False)
                c["points"] = np.stack([np.full(8, f["id"]), np.arange(8), 2 + np.cos(th + np.pi / 2), 3 + np.sin(th + np.pi / 2),
                                        np.full(8, z), np.zeros(8)], axis=1)
        return g
    fa, fb = post(fat(a), fat(b), anomalous=True)
    assert [f["contours"][0]["aortic_thickness"] for f in fa] == [2.0, 3.0]
    assert [f["contours"][0]["aortic_thickness"] for f in fb] == [2.0, 3.0]
    assert all(5 in f["contours"] for f in fa)     # wall contour added


def test_postprocess_errors_do_not_crash():  # :901-927
    with pytest.raises(ora.OracleError):
        ora.postprocess_pair(ora.encode_geometry([]), ora.encode_geometry([]))
    one = geometry([0.0], [1.0])
    fa, fb = post(one, one)
    assert len(fa) == 1 and len(fb) == 1
