"""`read_centerline_vtp` (src/intravascular/io/input.rs:259-458): ASCII VTP polylines -> PyCenterline with branches
ordered by descending arc length, forward-difference tangents, radii from MaximumInscribedSphereRadius."""
import math

import pytest

import multimodars as mm

VTP = """<?xml version="1.0"?>
<VTKFile type="PolyData" version="0.1" byte_order="LittleEndian">
  <PolyData>
    <Piece NumberOfPoints="7" NumberOfLines="2">
      <PointData>
        <DataArray type="Float64" Name="MaximumInscribedSphereRadius" format="ascii">
          1.0 1.1 1.2 1.3 1.4 1.5 1.6
        </DataArray>
      </PointData>
      <Points>
        <DataArray type="Float32" Name="Points" NumberOfComponents="3" format="ascii">
          0 0 0  1 0 0  2 0 0
          0 0 5  0 0 3  0 4 3  0 4 0
          <InformationKey name="L2_NORM_RANGE" location="vtkDataArray" length="2"/>
        </DataArray>
      </Points>
      <Lines>
        <DataArray type="Int64" Name="connectivity" format="ascii">0 1 2 3 4 5 6</DataArray>
        <DataArray type="Int64" Name="offsets" format="ascii">3 7</DataArray>
      </Lines>
    </Piece>
  </PolyData>
</VTKFile>
"""


def test_reads_branches_longest_first(tmp_path):
    p = tmp_path / "cl.vtp"
    p.write_text(VTP)
    cl = mm.read_centerline_vtp(str(p))
    assert len(cl) == 7 and cl.branch_start_indices == [0, 4]
    # branch 0 = the 9 mm polyline (points 3..6), branch 1 = the 2 mm one
    assert [q.branch_id for q in cl.points] == [0, 0, 0, 0, 1, 1, 1]
    assert cl.points_as_tuples()[:4] == [(0.0, 0.0, 5.0), (0.0, 0.0, 3.0), (0.0, 4.0, 3.0), (0.0, 4.0, 0.0)]
    assert [q.radius for q in cl.points] == [1.3, 1.4, 1.5, 1.6, 1.0, 1.1, 1.2]
    assert cl.points[0].tangent == (0.0, 0.0, -1.0) and cl.points[1].tangent == (0.0, 1.0, 0.0)
    assert cl.points[3].tangent == cl.points[2].tangent == (0.0, 0.0, -1.0)      # the last point repeats
    assert cl.points[4].tangent == (1.0, 0.0, 0.0) and cl.points[6].tangent == (1.0, 0.0, 0.0)
    assert [q.contour_point.frame_index for q in cl.points] == list(range(7))
    assert all(math.isclose(sum(t * t for t in q.tangent), 1.0) for q in cl.points)
    # and it feeds the centerline alignment: only branch 0 is used (preprocessing.rs:14-20)
    arr = mm.to_array(cl)
    assert arr.shape == (7, 4)


def test_rejects_binary_and_malformed_files(tmp_path):
    b = tmp_path / "bin.vtp"
    b.write_bytes(b"<VTKFile>\x00\x01\x02")
    with pytest.raises(mm.MmrsError, match="appears to be a binary VTP file"):
        mm.read_centerline_vtp(str(b))
    a = tmp_path / "app.vtp"
    a.write_text(VTP.replace('Name="connectivity" format="ascii"', 'Name="connectivity" format="appended"'))
    with pytest.raises(mm.MmrsError, match="binary-encoded DataArrays detected"):
        mm.read_centerline_vtp(str(a))
    m = tmp_path / "nolines.vtp"
    m.write_text(VTP.replace("<Lines>", "<Strips>").replace("</Lines>", "</Strips>"))
    with pytest.raises(mm.MmrsError, match="<Lines> section not found"):
        mm.read_centerline_vtp(str(m))
    o = tmp_path / "off.vtp"
    o.write_text(VTP.replace(">3 7<", ">3 6<"))
    with pytest.raises(mm.MmrsError, match=r"last offset \(6\) != connectivity length \(7\)"):
        mm.read_centerline_vtp(str(o))
    r = tmp_path / "range.vtp"
    r.write_text(VTP.replace(">0 1 2 3 4 5 6<", ">0 1 2 3 4 5 9<"))
    with pytest.raises(mm.MmrsError, match="connectivity index 9 out of range"):
        mm.read_centerline_vtp(str(r))
    with pytest.raises(mm.MmrsError, match="cannot open"):
        mm.read_centerline_vtp(str(tmp_path / "missing.vtp"))
    # a radius array of the wrong length is ignored (radii default to 0.0)
    w = tmp_path / "radii.vtp"
    w.write_text(VTP.replace("1.0 1.1 1.2 1.3 1.4 1.5 1.6", "1.0 1.1"))
    assert {q.radius for q in mm.read_centerline_vtp(str(w)).points} == {0.0}
