"""The claim behind the batched design: the cost curve of frame pair (i-1, i) does not depend on the
serial chain, up to f64 rounding that is orders of magnitude below the certification margin (1e-9 * max(1, Rmax)).
Measured here with the oracle alone: chain-state closure (tapped from the reference's loop) vs the "decoupled"
closure on the ORIGINAL frames, each centred on its own frame centroid — what align_within_many uploads."""
import numpy as np

from oracle import oracle_py as ora
from tests import golden_io as gio

TIE_MARGIN = 1e-9      # kTieMargin in multimoda-rs_b200/csrc/mmrs_host.cpp


def _decoupled_units(blob, sample):
    frames = ora.decode_geometry(blob)
    n_l = len(frames[0]["contours"][0]["points"])
    ratio = sample / n_l
    n_c = int(np.ceil(len(frames[0]["contours"][4]["points"]) * ratio)) if 4 in frames[0]["contours"] else None
    out = []
    for f in frames:
        lum = f["contours"][0]["points"]
        pts = lum[ora.downsample_indices(len(lum), sample), 2:4]
        if n_c is not None and 4 in f["contours"]:
            cath = f["contours"][4]["points"]
            pts = np.concatenate([pts, cath[ora.downsample_indices(len(cath), n_c), 2:4]])
        out.append(pts - np.array(f["centroid"][:2]))
    return out


def test_decoupled_costs_equal_chain_costs_to_rounding():
    pack = gio.inputs()
    a = gio.phase_arrays(pack, "stress", True)           # 25 frames x 501 points (+ 20 catheter points)
    blob = ora.build_geometry_from_arrays(a["lumen"], a["ref_point"], records=a["records"], diastole=True)
    dec = _decoupled_units(blob, 500)
    grid, _ = ora.search_grid(0.5, 90.0, None, 90.0)
    worst = 0.0
    for pair in (0, 5, 11, 17, 23):                       # late pairs carry the most accumulated chain rounding
        t, r, cen, best = ora.within_chain_tap(blob, 0.5, 90.0, True, 500, pair, threads=8)
        chain = ora.costs(t, r, cen, 0, grid)
        decoupled = ora.costs(dec[pair + 1], dec[pair], (0.0, 0.0), 0, grid)
        rmax = max(np.abs(dec[pair + 1]).max(), np.abs(dec[pair]).max())
        gap = np.abs(chain - decoupled).max()
        worst = max(worst, gap / max(1.0, rmax))
        assert gap < 1e-12, (pair, gap)
        assert int(np.argmin(chain)) == int(np.argmin(decoupled))
        assert grid[int(np.argmin(chain))] == best
    assert worst * 1e4 < TIE_MARGIN                      # >= 4 orders of magnitude of head-room
