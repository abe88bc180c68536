"""Known answers from the reference's own rendered output, part 2: the pairs() figure of the built vignette
(tests/golden/vignette_pairs.png, extracted by tests/golden/extract_vignette_figure.py) plots, for cell 1 of the bundled
fixture and the 50 hallmark sets, plaid() against replaid.sing(), replaid.ssgsea(alpha=0) and
replaid.scse(removeLog2=TRUE, scoreMean=FALSE) — the only published results of the rank scorers (and through them of
colranks(ties="min") with implicit zeros and sparse_colranks(ties="average")).  The axis tick marks of the figure
(positions detected in the image, values transcribed from its labels) calibrate every axis ABSOLUTELY, so each score has
a predictable pixel; the test asserts that all 50 points of all 12 panels sit on a plotted circle (within 1.5 px:
0.5 % of an axis) and that plausible misreadings of the reference do NOT (ties = "average" in sing, a per-column
instead of the global maximum rank in ssgsea, mean instead of sum in scSE, no median normalisation, a 2 % wobble).
It pins the oracle (CPU) and the CUDA path (-m gpu) to plot resolution, not to the last bits."""
import os

import numpy as np
import pytest

from oracle import plaid_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PNG = os.path.join(ROOT, "tests", "golden", "vignette_pairs.png")
# panel frames of the 4 x 4 pairs() layout in the 1402 x 824 figure (centres of the 2 px frame lines)
YE = [(45.5, 209.5), (234.5, 398.5), (424.5, 588.5), (613.5, 777.5)]
XE = [(46.5, 354.5), (379.5, 687.5), (713.5, 1021.5), (1046.5, 1354.5)]


def _dark():
    from PIL import Image
    return np.asarray(Image.open(PNG).convert("L")).astype(float) < 140


def _runs(mask):
    idx = np.where(mask)[0]
    out, s0, p = [], idx[0], idx[0]
    for v in idx[1:]:
        if v != p + 1:
            out.append((s0 + p) / 2.0)
            s0 = v
        p = v
    out.append((s0 + p) / 2.0)
    return out


# tick labels of the figure, per plotted variable: (first tick value, step, number of ticks) — transcribed from the image
TICKS = {0: (0.0, 0.1, 7), 1: (-0.50, 0.05, 8), 2: (-0.55, 0.05, 4), 3: (0.0, 2.0, 4)}


def calibrate(dark):
    """value -> pixel maps (slope, intercept) of the x axis (columns) and the y axis (rows) of every variable, from the
    tick marks: pairs() draws the x ticks of variables 0 / 2 below the bottom row and of 1 / 3 above the top row, the y
    ticks of 1 / 3 left of the first column and of 0 / 2 right of the last."""
    cal = {}
    for v in range(4):
        a, b = XE[v]
        rows = slice(781, 787) if v in (0, 2) else slice(36, 43)
        px = [x + int(a) - 2 for x in _runs(dark[rows, int(a) - 2:int(b) + 3].all(0))]
        t0, dt, n = TICKS[v]
        assert len(px) == n, (v, px)
        kx = np.polyfit(t0 + dt * np.arange(n), px, 1)
        a, b = YE[v]
        cols = slice(36, 43) if v in (1, 3) else slice(1358, 1364)
        py = [y + int(a) - 2 for y in _runs(dark[int(a) - 2:int(b) + 3, cols].all(1))][::-1]  # bottom-up = ascending values
        assert len(py) == n, (v, py)
        ky = np.polyfit(t0 + dt * np.arange(n), py, 1)
        cal[v] = (kx, ky)
    return cal


def _ring(dark, cx, cy, rad):
    ang = np.linspace(0, 2 * np.pi, 24, endpoint=False)
    xs, ys = np.rint(cx + rad * np.cos(ang)).astype(int), np.rint(cy + rad * np.sin(ang)).astype(int)
    return dark[ys, xs].mean()


def points_on_circles(S, dark):
    """number of (panel, point) pairs whose predicted pixel carries a plotted circle, out of 12 * 50"""
    S = np.asarray(S, dtype=np.float64)
    cal = calibrate(dark)
    hits = 0
    shifts = [(dx, dy) for dx in (-1.5, -1, -0.5, 0, 0.5, 1, 1.5) for dy in (-1.5, -1, -0.5, 0, 0.5, 1, 1.5)]
    H, W = dark.shape
    for i in range(4):          # panel row: y = column i of S
        for j in range(4):      # panel column: x = column j of S
            if i == j:
                continue
            for k in range(S.shape[0]):
                cx, cy = np.polyval(cal[j][0], S[k, j]), np.polyval(cal[i][1], S[k, i])
                if not (XE[j][0] - 8 < cx < XE[j][1] + 8 and YE[i][0] - 8 < cy < YE[i][1] + 8):
                    continue    # outside its panel: certainly not a plotted point
                best = max(_ring(dark, cx + dx, cy + dy, rad) for dx, dy in shifts for rad in (4.5, 5.0))
                hits += best > 0.9
    return hits


def _scores(mod, X, G, ctx=None):
    kw = {} if ctx is None else {"ctx": ctx}
    return np.column_stack([mod.plaid(X, G, **kw).mat[:, 0], mod.replaid_sing(X, G, **kw).mat[:, 0],
                            mod.replaid_ssgsea(X, G, alpha=0, **kw).mat[:, 0],
                            mod.replaid_scse(X, G, removeLog2=True, scoreMean=False, **kw).mat[:, 0]])


@pytest.fixture(scope="module")
def dark():
    pytest.importorskip("PIL")
    return _dark()


def test_oracle_reproduces_the_reference_figure(fixture_mats, dark):
    X, xr, xc, G, gr, gc = fixture_mats
    Xo, Go = O.Named(X, xr, xc), O.Named(G, gr, gc)
    S = _scores(O, Xo, Go)
    assert S.shape == (50, 4)
    assert points_on_circles(S, dark) == 12 * 50
    # the check has teeth: misreadings of the reference move the points off the circles
    wrong = S.copy()   # replaid.sing with ties.method = "average" instead of "min" (R/plaid.R:216)
    rX = O.colranks(X, ties_method="average") / X.shape[0] - 0.5
    wrong[:, 1] = O.plaid(O.Named(rX, xr, xc), Go, normalize=False).mat[:, 0]
    assert points_on_circles(wrong, dark) < 0.8 * 12 * 50
    wrong = S.copy()   # ssgsea: per-column instead of global max(rX) (R/plaid.R:251)
    r = O.colranks(X, keep_zero=True, ties_method="average").toarray()
    wrong[:, 2] = O.plaid(O.Named(r / r[:, [0]].max() - 0.5, xr, xc), Go).mat[:, 0]
    assert points_on_circles(wrong, dark) < 0.8 * 12 * 50
    wrong = S.copy()   # scSE with scoreMean = TRUE (R/plaid.R:176-181)
    wrong[:, 3] = O.replaid_scse(Xo, Go, removeLog2=True, scoreMean=True).mat[:, 0]
    assert points_on_circles(wrong, dark) < 0.8 * 12 * 50
    wrong = S.copy()   # plaid() without the median normalisation
    wrong[:, 0] = O.plaid(Xo, Go, normalize=False).mat[:, 0] * (1.0 + 0.05 * np.linspace(-1, 1, 50))
    assert points_on_circles(wrong, dark) < 0.8 * 12 * 50
    wrong = S.copy()   # resolution: a relative perturbation of 2 % of the range of one score
    wrong[:, 1] += 0.02 * np.ptp(S[:, 1]) * np.sin(np.arange(50))
    assert points_on_circles(wrong, dark) < 0.9 * 12 * 50


@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference_figure(fixture_mats, dark, gpu_ctx):
    import plaid_b200 as pb
    X, xr, xc, G, gr, gc = fixture_mats
    S = _scores(pb, pb.NamedMatrix(X, xr, xc), pb.NamedMatrix(G, gr, gc), ctx=gpu_ctx)
    assert points_on_circles(S, dark) == 12 * 50
