"""multimodars/ccta/centerline_prep.py of the reference: the same two functions under the same import path."""
from .._centerline import load_centerline, prepare_centerline  # noqa: F401
