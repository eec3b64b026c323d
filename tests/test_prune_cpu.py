"""The inequality the lower-bound pruning tier rests on (sweep_kernels.cuh K1p), checked with the CPU oracle's
Hausdorff distance (process_utils.rs:78-121 restated): for ANY row subsets A' of A and B' of B,
    H(A, B) >= max( max_{a in A'} min_{b in B} |a-b| , max_{b in B'} min_{a in A} |a-b| ),
with equality when A' = A and B' = B; and the selection rule: scoring exactly only the candidates whose bound does not
exceed the exact cost of the bound's own arg-min never loses the true arg-min (nor any candidate tied with it)."""
import numpy as np

from oracle import oracle_py as ora


def contour(rng, n, rot=0.0, e=0.2):
    phi = np.linspace(0, 2 * np.pi, n, endpoint=False)
    r = rng.uniform(1.5, 3) * (1 + e * np.cos(2 * phi) + 0.03 * np.cos(3 * phi + 0.4))
    return np.stack([r * np.cos(phi + rot), r * np.sin(phi + rot)], 1) + rng.normal(0, 0.005, (n, 2))


def directed_rows(rows, cols):
    d2 = ((rows[:, None, :] - cols[None, :, :]) ** 2).sum(-1)
    return np.sqrt(d2.min(1).max())


def rotate(p, a):
    c, s = np.cos(a), np.sin(a)
    return np.stack([p[:, 0] * c - p[:, 1] * s, p[:, 0] * s + p[:, 1] * c], 1)


def test_strided_bound_never_exceeds_hausdorff():
    rng = np.random.default_rng(0)
    for n, m, rows in ((200, 200, 32), (130, 257, 32), (64, 500, 16), (300, 300, 300)):
        a, b = contour(rng, n, 0.3), contour(rng, m)
        ia = (np.arange(min(rows, n)) * n) // min(rows, n)
        ib = (np.arange(min(rows, m)) * m) // min(rows, m)
        for ang in rng.uniform(-np.pi, np.pi, 25):
            ra = rotate(a, ang)
            h = ora.hausdorff(b, ra)
            lb = max(directed_rows(ra[ia], b), directed_rows(b[ib], ra))
            assert lb <= h * (1 + 1e-12) + 1e-15
            if rows >= max(n, m):
                assert abs(lb - h) <= 1e-12 * max(h, 1.0)
            # the second pass of k_lb rotates the reference rows by -theta instead: the same distances
            assert abs(directed_rows(rotate(b[ib], -ang), a) - directed_rows(b[ib], ra)) < 1e-12


def test_survivor_rule_keeps_the_arg_min_and_its_ties():
    rng = np.random.default_rng(1)
    for e in (0.3, 0.05, 0.0):   # 0.0: circles, a plateau where nothing can be pruned
        a, b = contour(rng, 160, 0.4, e), contour(rng, 160, 0.0, e)
        if e == 0.0:
            phi = np.linspace(0, 2 * np.pi, 160, endpoint=False)
            a = b = np.stack([2 * np.cos(phi), 2 * np.sin(phi)], 1)
        angles = np.deg2rad(np.arange(-90, 90.5, 0.5))
        ia = (np.arange(32) * 160) // 32
        h = np.array([ora.hausdorff(b, rotate(a, t)) for t in angles])
        lb = np.array([max(directed_rows(rotate(a, t)[ia], b), directed_rows(b[ia], rotate(a, t))) for t in angles])
        ub = h[int(np.argmin(lb))]                       # exact cost of the candidate with the smallest bound
        survivors = lb <= ub
        assert survivors[int(np.argmin(h))]
        assert survivors[h == h.min()].all()
        assert (h[~survivors] > h.min()).all()
        if e == 0.3:
            assert survivors.mean() < 0.5                # and it does prune when the cost depends on the angle
