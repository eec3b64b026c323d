"""`multimodars.ccta` in this build holds only the centerline preparation that feeds the alignment entry points
(`centerline_prep`: load_centerline, prepare_centerline — multimodars/ccta/centerline_prep.py in the reference).
The mesh side of the reference's package (labeling, scaling, stitching, discretisation, plots; trimesh based) is a
different product area and is not part of it: asking for one of those names says so."""
from .centerline_prep import load_centerline, prepare_centerline  # noqa: F401

_MESH_SIDE = ("label", "scale", "stitch", "export_section_stl", "create_wall_mesh", "labeling", "scaling", "stitching",
              "discretization_map", "fixing_functions", "debug_plots", "boundary")


def __getattr__(name):
    if name in _MESH_SIDE:
        raise AttributeError(f"multimodars.ccta.{name} belongs to the CCTA mesh side of the reference, which is not part "
                             "of this build (only multimodars.ccta.centerline_prep is; DESIGN.md §8)")
    raise AttributeError(f"module 'multimodars.ccta' has no attribute '{name}'")
