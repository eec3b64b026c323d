"""ORACLE (test infrastructure only — nothing under multimoda-rs_b200/ imports this file).

Python restatement of the reference's OBJ / MTL text writers, from their format strings:

    src/intravascular/io/output.rs:10-155     write_obj_mesh: vertices, mtllib / usemtl, vt, vn (inward radial normals in
                                              the x-y plane, NEGATED: a zero component prints as "-0"), two triangles per
                                              quad of consecutive contours, optional end caps (:107-141, close_end :157-181)
    src/intravascular/io/output.rs:183-197    write_obj_mesh_without_uv: every vt is "0 0"
    src/intravascular/to_object/texture.rs:6-28   UV coordinates ((point + 0.5) / n_points, (contour + 0.5) / n_contours)
    src/intravascular/to_object/write_mtl.rs      the MTL bodies quoted in mtl_text()

Rust's `{}` for an f64 prints the shortest digits that round-trip and never uses scientific notation
(core::fmt::float, `float_to_decimal_display`): 1.0 -> "1", 1e-7 -> "0.0000001", 1e21 -> "1000000000000000000000",
-0.0 -> "-0". output.rs holds no unit tests of its own, so this file is pinned by reading, line by line, against the format
strings above; the product's files are compared with it byte for byte (tests/test_export_cpu.py)."""
import math


def rust_f64(v):
    """Rust's `{}` for f64: shortest round-trip digits, never scientific."""
    if v != v:
        return "NaN"
    if math.isinf(v):
        return "-inf" if v < 0 else "inf"
    r = repr(float(v))
    if "e" in r or "E" in r:
        from decimal import Decimal
        r = format(Decimal(r), "f")
    if r.endswith(".0"):
        r = r[:-2]
    return r



def uv_coords(n_contours, n_points):  # to_object/texture.rs:6-28
    return [((pi + 0.5) / n_points, (ci + 0.5) / n_contours) for ci in range(n_contours) for pi in range(n_points)]


def mtl_text(kind, png=None):
    """write_mtl.rs: the material a mesh of `kind` refers to. png = None: the single-geometry export (entry.rs:741-818)."""
    if png is None:
        k = "1.0 1.0 1.0" if kind == "lumen" else "0.0 0.0 0.0"
        return f"newmtl material\nKa {k}\nKd {k}\nKs 0.0 0.0 0.0\n"
    name, k = {"lumen": ("displacement_material", "1 1 1"), "catheter": ("black_material", "0 0 0"),
               "wall": ("transparent_material", "0 0 0")}.get(kind, ("black_material", "0 0 0"))
    return f"newmtl {name}\nKa {k}\nKd {k}\nmap_Kd {png}\n"


def obj_text(contours, uv, mtl, watertight):
    """io/output.rs:10-181 as text. contours: objects with .points (each .x .y .z) and .centroid."""
    out = []
    offs, cur = [], 1
    for c in contours:
        offs.append(cur)
        for p in c.points:
            out.append(f"v {rust_f64(p.x)} {rust_f64(p.y)} {rust_f64(p.z)}")
            cur += 1
    out.append(f"mtllib {mtl}")
    out.append("usemtl displacement_material")
    for u, v in uv:
        out.append(f"vt {rust_f64(u)} {rust_f64(v)}")
    for c in contours:
        for p in c.points:
            dx, dy = p.x - c.centroid[0], p.y - c.centroid[1]
            ln = math.sqrt(dx * dx + dy * dy)
            nx, ny = (dx / ln, dy / ln) if ln > 0.0 else (0.0, 0.0)
            out.append(f"vn {rust_f64(-nx)} {rust_f64(-ny)} {rust_f64(-0.0)}")
    ppc = len(contours[0].points)
    tri = lambda a, b, c: f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}"
    for k in range(len(contours) - 1):
        o1, o2 = offs[k], offs[k + 1]
        for j in range(ppc):
            jn = (j + 1) % ppc
            out.append(tri(o1 + j, o1 + jn, o2 + j))
            out.append(tri(o2 + j, o1 + jn, o2 + jn))
    if watertight:
        a, z = contours[0].centroid, contours[-1].centroid
        out += [f"v {rust_f64(a[0])} {rust_f64(a[1])} {rust_f64(a[2])}", "vt 0.5 0.5", "vn 0.0 0.0 -1.0"]
        out += [f"v {rust_f64(z[0])} {rust_f64(z[1])} {rust_f64(z[2])}", "vt 0.5 0.5", "vn 0.0 0.0 1.0"]
        for i in range(ppc):
            out.append(tri(offs[0] + i, offs[0] + (i + 1) % ppc, cur))
        for i in range(ppc):
            out.append(tri(cur + 1, offs[-1] + (i + 1) % ppc, offs[-1] + i))
    return "\n".join(out) + "\n"
