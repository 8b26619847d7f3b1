"""Size-independent row checks against the oracle's assembled CSR system.  TEST INFRASTRUCTURE ONLY (tests/, bench.py's
parity block): a relaxation sweep is local, so windows cut out of full-size blocks are handed to the oracle's own system
assembly (``RowCompressedMatrixSystem2d`` through ``oracle.System``: smooth.zig:923-992 interior rows, :994-1105 interface
rows) and the damped-Jacobi update computed from its rows is compared with what the CUDA path wrote for the same nodes.
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc


class _Block:
    def __init__(self, points):
        self.points = np.ascontiguousarray(points, dtype=np.float64)


class _Mesh:
    def __init__(self, blocks, connections=()):
        self.blocks = [_Block(b) for b in blocks]
        self.connections = list(connections)
        self.boundary_conditions = []


class _Range:
    def __init__(self, block, side, start, end):
        self.block, self.side, self.start, self.end = block, int(side), start, end


class _Connection:
    def __init__(self, r0, r1, periodicity=None):
        self.ranges = (r0, r1)
        self.periodicity = periodicity


def _rows_update(p, idx, v, rx, ry, flat, rows, omega):
    """x + omega (b - A x)_r / a_rr for the given rows of a CSR system"""
    out = np.empty((len(rows), 2))
    for k, r in enumerate(rows):
        a, cols = v[p[r]:p[r + 1]], idx[p[r]:p[r + 1]]
        assert len(cols) == 9
        diag = a[cols == r][0]
        out[k, 0] = flat[r, 0] + omega * (rx[r] - np.dot(a, flat[cols, 0])) / diag
        out[k, 1] = flat[r, 1] + omega * (ry[r] - np.dot(a, flat[cols, 1])) / diag
    return out


def interior_update(window: np.ndarray, omega: float) -> np.ndarray:
    """Damped-Jacobi update of the interior nodes of ``window`` (ni, nj, 2) from the oracle's interior rows; the rim is
    returned unchanged."""
    sys_ = orc.System(_Mesh([window]), orc.options(control_function="laplace"))
    sys_.fill(0)
    p, idx, v, rx, ry = sys_.csr()
    sys_.close()
    ni, nj = window.shape[:2]
    flat = np.ascontiguousarray(window, dtype=np.float64).reshape(-1, 2)
    rows = (np.arange(1, ni - 1)[:, None] * nj + np.arange(1, nj - 1)[None, :]).ravel()
    out = flat.copy()
    out[rows] = _rows_update(p, idx, v, rx, ry, flat, rows, omega)
    return out.reshape(ni, nj, 2)


def interface_update(a: np.ndarray, b: np.ndarray, side_a: int, side_b: int, omega: float, periodicity=None) -> np.ndarray:
    """Two windows joined over their whole common side (side_a of ``a`` = i_min (0) or i_max (1): its first / last j line; side
    j_min (2) / j_max (3): its first / last i line): damped-Jacobi update of the interior nodes of the interface line of
    ``a`` (the `smoothed` side-0 rows) from the oracle's interface rows.  Returns (n - 2, 2) for a line of n nodes."""
    side_a, side_b = int(side_a), int(side_b)
    n = a.shape[0] if side_a < 2 else a.shape[1]
    conn = _Connection(_Range(0, side_a, 0, n - 1), _Range(1, side_b, 0, n - 1), periodicity)
    sys_ = orc.System(_Mesh([a, b], [conn]), orc.options(control_function="laplace"))
    sys_.fill(0)
    p, idx, v, rx, ry = sys_.csr()
    sys_.close()
    flat = np.concatenate([np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 2), np.ascontiguousarray(b, dtype=np.float64).reshape(-1, 2)])
    ni, nj = a.shape[:2]
    if side_a == 0:
        rows = np.arange(1, ni - 1) * nj
    elif side_a == 1:
        rows = np.arange(1, ni - 1) * nj + nj - 1
    elif side_a == 2:
        rows = np.arange(1, nj - 1)
    else:
        rows = (ni - 1) * nj + np.arange(1, nj - 1)
    return _rows_update(p, idx, v, rx, ry, flat, rows, omega)


# ---- general form: any connection (sub-ranges, reversed traversal, periodic), windows cut around nodes k0 .. k0 + n - 1 ----
def cut_window(points: np.ndarray, rng, k0: int, n: int, depth: int):
    """Window of `depth` lines inward from the side of `rng` that holds nodes k0 .. k0 + n - 1 of the range (in traversal
    order).  Returns (window, (mini_start, mini_end), index) where `index(k)` is the (i, j) of range node k0 + k in the block."""
    ni, nj = points.shape[:2]
    side, s, e = int(rng.side), int(rng.start), int(rng.end)
    sign = 1 if e >= s else -1
    a, b = s + sign * k0, s + sign * (k0 + n - 1)
    lo, hi = min(a, b), max(a, b)
    if side == 0:
        win, index = points[lo:hi + 1, 0:depth], (lambda k: (a + sign * k, 0))
    elif side == 1:
        win, index = points[lo:hi + 1, nj - depth:nj], (lambda k: (a + sign * k, nj - 1))
    elif side == 2:
        win, index = points[0:depth, lo:hi + 1], (lambda k: (0, a + sign * k))
    else:
        win, index = points[ni - depth:ni, lo:hi + 1], (lambda k: (ni - 1, a + sign * k))
    mini = (0, n - 1) if sign > 0 else (n - 1, 0)
    return np.ascontiguousarray(win), mini, index


def connection_update(win_a, side_a, mini_a, win_b, side_b, mini_b, omega: float, periodicity=None) -> np.ndarray:
    """Damped-Jacobi update of the interior nodes (k = 1 .. n - 2, traversal order of side a) of the `smoothed` side-0 rows of
    the connection between the two windows, from the oracle's interface rows (smooth.zig:994-1105)."""
    conn = _Connection(_Range(0, side_a, mini_a[0], mini_a[1]), _Range(1, side_b, mini_b[0], mini_b[1]), periodicity)
    sys_ = orc.System(_Mesh([win_a, win_b], [conn]), orc.options(control_function="laplace"))
    sys_.fill(0)
    p, idx, v, rx, ry = sys_.csr()
    sys_.close()
    flat = np.concatenate([np.ascontiguousarray(win_a, dtype=np.float64).reshape(-1, 2), np.ascontiguousarray(win_b, dtype=np.float64).reshape(-1, 2)])
    ni, nj = win_a.shape[:2]
    n = abs(mini_a[1] - mini_a[0]) + 1
    sign = 1 if mini_a[1] >= mini_a[0] else -1
    pos = [mini_a[0] + sign * k for k in range(1, n - 1)]
    side_a = int(side_a)
    if side_a == 0:
        rows = [q * nj for q in pos]
    elif side_a == 1:
        rows = [q * nj + nj - 1 for q in pos]
    elif side_a == 2:
        rows = list(pos)
    else:
        rows = [(ni - 1) * nj + q for q in pos]
    return _rows_update(p, idx, v, rx, ry, flat, np.array(rows), omega)
