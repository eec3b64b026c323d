"""`read_centerline_vtp` (src/intravascular/io/input.rs:259-458 behind binding/functions.rs:1542-1546): ASCII VTK
PolyData (.vtp) centerline -> PyCenterline. Each polyline of <Lines> is a branch; branches are numbered by descending
arc length (branch 0 = the longest); tangents are normalised forward differences, the last point of a branch repeats
its predecessor's; radii come from PointData/MaximumInscribedSphereRadius when its length matches."""
from __future__ import annotations

import math

from ._native import MmrsError
from ._types import PyCenterline, PyCenterlinePoint, PyContourPoint


def _section(xml: str, tag: str) -> str:
    start = xml.find(f"<{tag}")
    if start < 0:
        raise MmrsError(f"VTP: <{tag}> section not found")
    end = xml.find(f"</{tag}>", start)
    if end < 0:
        raise MmrsError(f"VTP: </{tag}> not found")
    return xml[start:end + len(tag) + 3]


def _array_text(section: str, name: str) -> str:
    pos = section.find(f'Name="{name}"')
    if pos < 0:
        raise MmrsError(f'VTP: DataArray Name="{name}" not found')
    da = section.rfind("<DataArray", 0, pos)
    if da < 0:
        raise MmrsError(f'VTP: no <DataArray before Name="{name}"')
    gt = section.find(">", da)
    if gt < 0:
        raise MmrsError(f'VTP: unclosed <DataArray Name="{name}">')
    close = section.find("</DataArray>", gt + 1)
    if close < 0:
        raise MmrsError(f'VTP: no </DataArray> for Name="{name}"')
    text = section[gt + 1:close].strip()
    lt = text.find("<")          # some exporters nest <InformationKey> nodes inside the Points array
    return (text if lt < 0 else text[:lt]).strip()


def _numbers(text: str, conv):
    out = []
    for tok in text.split():
        try:
            out.append(conv(tok))
        except ValueError:
            raise MmrsError(f"VTP: bad number '{tok}'") from None
    return out


def read_centerline_vtp(file_path: str) -> PyCenterline:
    try:
        raw = open(file_path, "rb").read()
    except OSError as e:
        raise MmrsError(f'cannot open "{file_path}": {e.strerror}') from None
    if any(b < 0x09 or 0x0d < b < 0x20 for b in raw[:512]):
        raise MmrsError(f'"{file_path}" appears to be a binary VTP file; only ASCII-format VTP is supported. '
                        "Re-export from your software with 'ASCII' data mode.")
    try:
        xml = raw.decode("utf-8")
    except UnicodeDecodeError:
        raise MmrsError(f'"{file_path}": not valid UTF-8') from None
    for fmt in ('format="binary"', 'format="appended"'):
        if fmt in xml:
            raise MmrsError(f'"{file_path}": binary-encoded DataArrays detected ({fmt}); only ASCII format is '
                            "supported. Re-export with 'ASCII' data mode.")
    flat = _numbers(_array_text(_section(xml, "Points"), "Points"), float)
    if len(flat) % 3:
        raise MmrsError(f"VTP: Points array length {len(flat)} not divisible by 3")
    coords = [(flat[i], flat[i + 1], flat[i + 2]) for i in range(0, len(flat), 3)]
    n_pts = len(coords)
    radii = [0.0] * n_pts
    try:
        r = _numbers(_array_text(_section(xml, "PointData"), "MaximumInscribedSphereRadius"), float)
        if len(r) == n_pts:
            radii = r
    except MmrsError:
        pass
    lines = _section(xml, "Lines")
    conn = _numbers(_array_text(lines, "connectivity"), int)
    offs = _numbers(_array_text(lines, "offsets"), int)
    if any(v < 0 for v in conn) or any(v < 0 for v in offs):
        raise MmrsError("VTP: bad number (negative index)")
    if not offs:
        raise MmrsError("VTP: Lines section is empty (no branches)")
    if offs[-1] != len(conn):
        raise MmrsError(f"VTP: last offset ({offs[-1]}) != connectivity length ({len(conn)})")
    branches = [conn[a:b] for a, b in zip([0] + offs[:-1], offs)]

    def arc(branch):
        total = 0.0
        for p, q in zip(branch, branch[1:]):
            (x0, y0, z0), (x1, y1, z1) = coords[p], coords[q]
            total += math.sqrt((x1 - x0) ** 2 + (y1 - y0) ** 2 + (z1 - z0) ** 2)
        return total

    for b in branches:
        for idx in b:
            if idx >= n_pts:
                raise MmrsError(f"VTP: connectivity index {idx} out of range ({n_pts} points)")
    lengths = [arc(b) for b in branches]
    order = sorted(range(len(branches)), key=lambda k: -lengths[k])
    points, starts = [], []
    for branch_id, k in enumerate(order):
        starts.append(len(points))
        branch = branches[k]
        for i, pt in enumerate(branch):
            x, y, z = coords[pt]
            idx = len(points)
            if i + 1 < len(branch):
                nx, ny, nz = coords[branch[i + 1]]
                dx, dy, dz = nx - x, ny - y, nz - z
                nrm = math.sqrt(dx * dx + dy * dy + dz * dz)
                tangent = (dx / nrm, dy / nrm, dz / nrm) if nrm > 1e-12 else (0.0, 0.0, 0.0)
            elif i > 0:
                tangent = points[-1].tangent
            else:
                tangent = (0.0, 0.0, 0.0)
            p = PyCenterlinePoint(PyContourPoint(idx, idx, x, y, z, False), tangent, branch_id)
            p.radius = radii[pt]
            points.append(p)
    cl = PyCenterline(points)
    cl.branch_start_indices = starts
    return cl
