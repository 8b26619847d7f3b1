"""A second, independently written evaluation of the smoother's discrete specification.  TEST INFRASTRUCTURE ONLY.

Written from the text of SURVEY.md Appendix A (A.1 numbering, A.2 connection traversal, A.3 interface rows, A.4 node
kinds, A.5 stencil coefficients, A.6 junctions, A.7 outer loop, A.9 White) in vectorised numpy, sharing no code with
``oracle/turbomesh_oracle.c`` or with the CUDA library.  Where the C oracle transcribes the reference's procedures
(pair scans, CSR position tables, periodicity bookkeeping per appended copy), this file derives the same equations
from *what they mean*: junction groups are connected components of the end-point pair graph, the shift of every copy
relative to its primary comes from a weighted union-find over ``x1 = x0 + p``, rows are dictionaries of (column, value).

Two uses (tests/test_independent_cpu.py):
  1. the oracle's assembled CSR system must equal this evaluation row by row on hand-sized meshes and on T106;
  2. with ``dtype=np.longdouble`` (x87 80-bit, 64-bit mantissa) and a direct solve refined in that precision it is the
     extended-precision *truth* of the exact Picard sequence (smooth.zig:104-154) for the parity bound of config 1.

Reference lines each piece follows are cited at the functions.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

FIXED, SMOOTHED, CONNECTED, LAPLACIAN, SLIDING = 0, 1, 2, 3, 4   # smooth.zig:1168-1174
I_MIN, I_MAX, J_MIN, J_MAX = 0, 1, 2, 3                          # boundary.zig:8-13


class Walk:
    """A.2 (RangeFillMatrixIterator.init, smooth.zig:1556-1598): block-local ids along a range, inward and along shifts."""

    def __init__(self, rng, ni, nj):
        side, s, e = int(rng.side), int(rng.start), int(rng.end)
        if side == I_MIN:
            first, along, inward = s * nj, nj, 1
        elif side == I_MAX:
            first, along, inward = s * nj + nj - 1, nj, -1
        elif side == J_MIN:
            first, along, inward = s, 1, nj
        else:
            first, along, inward = (ni - 1) * nj + s, 1, -nj
        if s > e:
            along = -along
        self.count = abs(s - e) + 1
        self.along, self.inward = along, inward
        self.ids = first + along * np.arange(self.count, dtype=np.int64)


class IndependentSystem:
    """Global system of one Picard step, rows as COO triplets in ``dtype``."""

    def __init__(self, mesh, dtype=np.float64, white=None):
        """``white`` = None (Laplace) or (ds_target, theta_target)."""
        self.dtype = dtype
        self.sizes = [tuple(b.points.shape[:2]) for b in mesh.blocks]
        self.off = np.concatenate([[0], np.cumsum([ni * nj for ni, nj in self.sizes])]).astype(np.int64)   # A.1
        self.n = int(self.off[-1])
        self.xy = np.concatenate([np.asarray(b.points, dtype=dtype).reshape(-1, 2) for b in mesh.blocks])
        self.conns = list(mesh.connections)
        self.bcs = list(mesh.boundary_conditions)
        self.white = white
        self.cf = np.zeros((self.n, 2), dtype=dtype)           # (P, Q) per node, wall_control_function.zig:24
        self._walks = [(Walk(c.ranges[0], *self.sizes[c.ranges[0].block]), Walk(c.ranges[1], *self.sizes[c.ranges[1].block])) for c in self.conns]
        self._classify()
        self.initial = self.xy.copy()                           # rhs of fixed / sliding-x rows (smooth.zig:790-796, 853-858)
        if white is not None:
            self._white_init()
        self.outer_done = 0

    # ------------------------------------------------------------------------------------------------
    # topology: kinds (A.4), junction groups and shifts (A.6)
    # ------------------------------------------------------------------------------------------------
    def _per(self, c):
        p = self.conns[c].periodicity
        return np.zeros(2, dtype=self.dtype) if p is None else np.array([p[0], p[1]], dtype=self.dtype)

    def _boundary_ids(self, b):
        ni, nj = self.sizes[b]
        i, j = np.meshgrid(np.arange(ni), np.arange(nj), indexing="ij")
        m = (i == 0) | (i == ni - 1) | (j == 0) | (j == nj - 1)
        return self.off[b] + (i * nj + j)[m]

    def _block_of(self, g):
        return int(np.searchsorted(self.off, g, side="right") - 1)

    def _classify(self):
        kind = {}
        for b in range(len(self.sizes)):
            for g in self._boundary_ids(b):
                kind[int(g)] = FIXED
        # --- junction groups: connected components of the end-point pair graph that contain a repeated end point ---
        parent, shift = {}, {}      # weighted union-find (no compression: a handful of nodes): x[g] = x[parent[g]] + shift[g]

        def root_of(g):
            s = np.zeros(2, dtype=self.dtype)
            while parent[g] != g:
                s = s + shift[g]
                g = parent[g]
            return g, s             # x[g_in] = x[root] + s

        occurrences = {}
        pairs = []
        for c, (w0, w1) in enumerate(self._walks):
            o0, o1 = self.off[self.conns[c].ranges[0].block], self.off[self.conns[c].ranges[1].block]
            for k in (0, w0.count - 1):
                g0, g1 = int(o0 + w0.ids[k]), int(o1 + w1.ids[k])
                pairs.append((g0, g1, self._per(c)))
                for g in (g0, g1):
                    occurrences[g] = occurrences.get(g, 0) + 1
                    if g not in parent:
                        parent[g] = g
                        shift[g] = np.zeros(2, dtype=self.dtype)
        for g0, g1, p in pairs:     # x[g1] = x[g0] + p
            r0, s0 = root_of(g0)
            r1, s1 = root_of(g1)
            if r0 != r1:            # x[r1] = x[g1] - s1 = x[r0] + s0 + p - s1
                parent[r1] = r0
                shift[r1] = s0 + p - s1
        comps = {}
        for g in parent:
            comps.setdefault(root_of(g)[0], []).append(g)
        self.junctions = []
        for root, members in comps.items():
            if not any(occurrences[g] >= 2 for g in members):
                continue
            members = sorted(members)
            primary = members[0]
            base = root_of(primary)[1]
            copies = [(g, root_of(g)[1] - base) for g in members]   # x[g] = x[primary] + s
            self.junctions.append({"primary": primary, "copies": copies})
        self.junctions.sort(key=lambda j: j["primary"])
        self.master = {}                    # connected node -> (master node, shift): x_self = x_master + shift
        for jn in self.junctions:
            for g, s in jn["copies"]:
                if g == jn["primary"]:
                    kind[g] = LAPLACIAN
                else:
                    kind[g] = CONNECTED
                    self.master[g] = (jn["primary"], s)
        # --- inlet / outlet ranges slide (smooth.zig:1265-1277); walls stay fixed ---
        self.sliding_inner = {}
        for bc in self.bcs:
            if int(bc.kind) == 0:
                continue
            b = bc.range.block
            w = Walk(bc.range, *self.sizes[b])
            for l in w.ids:
                g = int(self.off[b] + l)
                kind[g] = SLIDING
                self.master.pop(g, None)
                self.sliding_inner[g] = g + w.inward
        # --- per connection, in declaration order (smooth.zig:1279-1329) ---
        self.smoothed_conn = {}
        for c, (w0, w1) in enumerate(self._walks):
            o0, o1 = self.off[self.conns[c].ranges[0].block], self.off[self.conns[c].ranges[1].block]
            p = self._per(c)
            for k in range(w0.count):
                g0, g1 = int(o0 + w0.ids[k]), int(o1 + w1.ids[k])
                if k == 0 or k == w0.count - 1:
                    if kind[g0] in (FIXED, SLIDING):
                        kind[g1] = CONNECTED
                        self.master[g1] = (g0, p)
                        self.sliding_inner.pop(g1, None)
                else:
                    kind[g0] = SMOOTHED
                    kind[g1] = CONNECTED
                    self.smoothed_conn[g0] = (c, k)
                    self.master[g1] = (g0, p)
        self.kind = kind

    def kinds_flat(self):
        """Node kinds in the reference's flat boundary numbering (A.1, boundary.zig:248-285)."""
        out = []
        for b, (ni, nj) in enumerate(self.sizes):
            k = np.empty(2 * (ni + nj - 2), dtype=np.uint8)
            o = int(self.off[b])
            for j in range(nj):
                k[j] = self.kind[o + j]
                k[nj + 2 * (ni - 2) + j] = self.kind[o + (ni - 1) * nj + j]
            for i in range(1, ni - 1):
                k[nj + 2 * (i - 1)] = self.kind[o + i * nj]
                k[nj - 1 + 2 * i] = self.kind[o + i * nj + nj - 1]
            out.append(k)
        return np.concatenate(out)

    # ------------------------------------------------------------------------------------------------
    # A.5: the nine coefficients from W, E, S, N and (P, Q)   (StencilData.init, smooth.zig:192-215)
    # ------------------------------------------------------------------------------------------------
    def _stencil(self, W, E, S, N, P, Q):
        half = self.dtype(0.5)
        x_xi, y_xi = half * (E[..., 0] - W[..., 0]), half * (E[..., 1] - W[..., 1])
        x_eta, y_eta = half * (N[..., 0] - S[..., 0]), half * (N[..., 1] - S[..., 1])
        g22 = x_eta * x_eta + y_eta * y_eta
        g12 = x_xi * x_eta + y_xi * y_eta
        g11 = x_xi * x_xi + y_xi * y_xi
        return {(0, 0): -2 * g22 - 2 * g11,
                (1, 0): g22 * (1 + half * P), (-1, 0): g22 * (1 - half * P),
                (0, 1): g11 * (1 + half * Q), (0, -1): g11 * (1 - half * Q),
                (1, 1): -half * g12, (1, -1): half * g12, (-1, 1): half * g12, (-1, -1): -half * g12}

    # ------------------------------------------------------------------------------------------------
    # assembly of one Picard step; y_mode only changes the sliding rows (A.4, smooth.zig:1115-1165)
    # ------------------------------------------------------------------------------------------------
    def assemble(self, y_mode):
        rows, cols, vals = [], [], []
        rhs = np.zeros(self.n, dtype=self.dtype)
        comp = 1 if y_mode else 0
        X = self.xy

        def emit(r, c, v):
            rows.append(np.asarray(r, dtype=np.int64).ravel()); cols.append(np.asarray(c, dtype=np.int64).ravel()); vals.append(np.asarray(v, dtype=self.dtype).ravel())

        # interior rows (A.4 / A.5; fillBlockInternalPointData, smooth.zig:923-992)
        for b, (ni, nj) in enumerate(self.sizes):
            o = int(self.off[b])
            P3 = X[o:o + ni * nj].reshape(ni, nj, 2)
            cf = self.cf[o:o + ni * nj].reshape(ni, nj, 2)
            a = self._stencil(P3[:-2, 1:-1], P3[2:, 1:-1], P3[1:-1, :-2], P3[1:-1, 2:], cf[1:-1, 1:-1, 0], cf[1:-1, 1:-1, 1])
            i, j = np.meshgrid(np.arange(1, ni - 1), np.arange(1, nj - 1), indexing="ij")
            g = o + i * nj + j
            for (di, dj), v in a.items():
                emit(g, g + di * nj + dj, v)
        # block-boundary rows
        for g, k in self.kind.items():
            if k == FIXED:
                emit(g, g, 1.0)
                rhs[g] = self.initial[g, comp]
            elif k == CONNECTED:            # x_master - x_self = -shift
                m, s = self.master[g]
                emit([g, g], [m, g], [1.0, -1.0])
                rhs[g] = -s[comp]
            elif k == SLIDING:
                if not y_mode:
                    emit(g, g, 1.0)
                    rhs[g] = self.initial[g, 0]
                else:                        # [1, -1] on the two columns in ascending order (positional, smooth.zig:837-859)
                    lo, hi = sorted((g, self.sliding_inner[g]))
                    emit([g, g], [lo, hi], [1.0, -1.0])
        # interface rows (A.3; fillBlockConnectionData, smooth.zig:994-1105)
        by_conn = {}
        for g0, (c, k) in self.smoothed_conn.items():
            by_conn.setdefault(c, []).append((g0, k))
        for c, lst in by_conn.items():
            lst.sort(key=lambda t: t[1])
            w0, w1 = self._walks[c]
            o0, o1 = int(self.off[self.conns[c].ranges[0].block]), int(self.off[self.conns[c].ranges[1].block])
            ks = np.array([k for _, k in lst], dtype=np.int64)
            g0 = o0 + w0.ids[ks]
            g1 = o1 + w1.ids[ks]
            d0, n0, d1, n1 = w0.along, w0.inward, w1.along, w1.inward
            p = self._per(c)
            periodic = self.conns[c].periodicity is not None
            N = X[g1 + n1] - p                                                        # smooth.zig:1032
            P, Q = self.cf[g0, 0], self.cf[g0, 1]
            if not periodic:
                P, Q = Q, P                                                           # smooth.zig:1082-1083 vs 1040-1041
            a = self._stencil(X[g0 - d0], X[g0 + d0], X[g0 + n0], N, P, Q)
            for (da, dn), v in a.items():
                col = (g0 + da * d0 + n0) if dn == -1 else (g0 + da * d0) if dn == 0 else (g1 + da * d1 + n1)
                emit(g0, col, v)
            rhs[g0] = p[comp] * (a[(-1, 1)] + a[(0, 1)] + a[(1, 1)])                  # smooth.zig:1060-1061
        # junction rows (A.6 step 4; smooth.zig:813-836, 917-920, 1457-1511)
        for jn in self.junctions:
            g = jn["primary"]
            nbrs, total = [], np.zeros(2, dtype=self.dtype)
            for cg, s in jn["copies"]:
                b = self._block_of(cg)
                ni, nj = self.sizes[b]
                l = cg - int(self.off[b]); i, j = divmod(l, nj)
                ii = [1] if i == 0 else [ni - 2] if i == ni - 1 else [i - 1, i + 1]
                jj = [1] if j == 0 else [nj - 2] if j == nj - 1 else [j - 1, j + 1]
                for a_ in ii:
                    for b_ in jj:
                        nbrs.append(int(self.off[b]) + a_ * nj + b_)
                        total = total + s
            assert self.kind[g] == LAPLACIAN or self.kind[g] in (SLIDING, CONNECTED)
            if self.kind[g] != LAPLACIAN:
                continue                     # re-classified by a later rule (A.4 precedence): that rule's row stands
            emit([g] * len(nbrs), nbrs, [1.0] * len(nbrs))
            emit(g, g, -float(len(nbrs)))
            rhs[g] = total[comp]
        r, c, v = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
        return r, c, v, rhs

    def csr(self, y_mode, dtype=np.float64):
        r, c, v, rhs = self.assemble(y_mode)
        A = sp.coo_matrix((v.astype(dtype), (r, c)), shape=(self.n, self.n)).tocsr()
        A.sum_duplicates()
        return A, rhs

    # ------------------------------------------------------------------------------------------------
    # White control function (A.9)
    # ------------------------------------------------------------------------------------------------
    def _blend(self, o, ni, nj, pq_wall):
        """every i line: wall value times 1 - j/(nj-1)   (wall_control_function.zig:104-111)"""
        j = np.arange(nj, dtype=self.dtype)
        factor = 1 - j / (self.dtype(nj) - 1)
        factor[0] = 1
        self.cf[o:o + ni * nj] = (pq_wall[:, None, :] * factor[None, :, None]).reshape(-1, 2)

    @staticmethod
    def _eq610(xi, xi2, eta, eta2):
        g11 = xi[..., 0] ** 2 + xi[..., 1] ** 2
        g22 = eta[..., 0] ** 2 + eta[..., 1] ** 2
        dot = lambda a, b: a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]
        p = -dot(xi, xi2) / g11 - dot(xi, eta2) / g22
        q = -dot(eta, eta2) / g22 - dot(eta, xi2) / g11
        return np.stack([p, q], axis=-1)

    def _wall_frames(self, b, second):
        """xi derivative along the wall (central; one-sided at the two ends), eta one-sided into the block; optionally the second differences"""
        ni, nj = self.sizes[b]
        o = int(self.off[b])
        P3 = self.xy[o:o + ni * nj].reshape(ni, nj, 2)
        w = P3[:, 0]
        half = self.dtype(0.5)
        xi = np.empty_like(w); xi2 = np.empty_like(w)
        xi[1:-1] = half * (w[2:] - w[:-2]); xi2[1:-1] = w[2:] - 2 * w[1:-1] + w[:-2]
        xi[0] = -w[0] + w[1]; xi2[0] = w[0] - 2 * w[1] + w[2]
        xi[-1] = w[-1] - w[-2]; xi2[-1] = w[-1] - 2 * w[-2] + w[-3]
        eta = -P3[:, 0] + P3[:, 1]
        eta2 = P3[:, 0] - 2 * P3[:, 1] + P3[:, 2]
        return (xi, xi2, eta, eta2) if second else (xi, eta)

    def _le_frame(self):
        """the shared wall node of connection 0: xi across the two O-grid halves, eta along the connection (:203-279, 394-431)"""
        w0, w1 = self._walks[0]
        c = self.conns[0]
        assert c.ranges[0].block == 0 and c.ranges[1].block == 1 and int(c.ranges[0].side) == J_MIN and int(c.ranges[1].side) == J_MIN
        assert c.ranges[0].start == 0 and c.ranges[1].start == 0 and c.periodicity is None
        X = self.xy
        o1 = int(self.off[1])
        return X[0], X[w0.inward], X[o1 + w1.inward], X[w0.along], X[2 * w0.along]

    def _white_init(self):
        for b in (0, 1):
            ni, nj = self.sizes[b]
            xi, xi2, eta, eta2 = self._wall_frames(b, True)
            self._blend(int(self.off[b]), ni, nj, self._eq610(xi, xi2, eta, eta2))
        c, ip1, im1, jp1, jp2 = self._le_frame()
        half = self.dtype(0.5)
        pq = self._eq610(half * (ip1 - im1), ip1 - 2 * c + im1, -c + jp1, c - 2 * jp1 + jp2)
        nj = self.sizes[0][1]
        self._blend(0, 1, nj, pq[None, :])

    def _delta(self, xi, eta):
        ds_t, th_t = self.dtype(self.white[0]), self.dtype(self.white[1])
        g11 = xi[..., 0] ** 2 + xi[..., 1] ** 2
        g12 = xi[..., 0] * eta[..., 0] + xi[..., 1] * eta[..., 1]
        g22 = eta[..., 0] ** 2 + eta[..., 1] ** 2
        ds = np.sqrt(g22)
        theta = np.arccos(g12 / np.sqrt(g11 * g22))
        tenth = self.dtype(0.1)
        return np.stack([tenth * -np.arctan2(th_t - theta, th_t), tenth * np.arctan2(ds_t - ds, ds_t)], axis=-1)

    def white_update(self):
        for b in (0, 1):
            ni, nj = self.sizes[b]
            o = int(self.off[b])
            xi, eta = self._wall_frames(b, False)
            wall = self.cf[o:o + ni * nj].reshape(ni, nj, 2)[:, 0].copy()
            self._blend(o, ni, nj, wall + self._delta(xi, eta))
        c, ip1, im1, jp1, _ = self._le_frame()
        half = self.dtype(0.5)
        pq = self.cf[0] + self._delta(-half * (ip1 - im1), -c + jp1)      # sign flip: wall_control_function.zig:429-431
        self._blend(0, 1, self.sizes[0][1], pq[None, :])

    # ------------------------------------------------------------------------------------------------
    # A.7: one exact Picard step (direct solve; refined in `dtype` when that is wider than fp64)
    # ------------------------------------------------------------------------------------------------
    def _solve(self, r, c, v, rhs, x0):
        A64 = sp.coo_matrix((v.astype(np.float64), (r, c)), shape=(self.n, self.n)).tocsc()
        lu = spla.splu(A64)
        if self.dtype == np.float64:
            return lu.solve(rhs.astype(np.float64)), 0.0
        order = np.argsort(r, kind="stable")
        rs, cs, vs = r[order], c[order], v[order]
        starts = np.flatnonzero(np.concatenate([[True], rs[1:] != rs[:-1]]))
        assert len(starts) == self.n
        x = x0.astype(self.dtype).copy()
        last = None
        for _ in range(12):
            res = rhs - np.add.reduceat(vs * x[cs], starts)
            dx = lu.solve(res.astype(np.float64))
            x = x + dx.astype(self.dtype)
            last = float(np.abs(dx).max())
            if last < 1e-19:
                break
        return x, last

    def step(self):
        if self.white is not None and self.outer_done > 0:
            self.white_update()
        new = np.empty_like(self.xy)
        self.last_refinement = []
        for comp in (0, 1):
            r, c, v, rhs = self.assemble(bool(comp))
            new[:, comp], last = self._solve(r, c, v, rhs, self.xy[:, comp])
            self.last_refinement.append(last)
        self.xy = new
        self.outer_done += 1

    def blocks(self):
        return [self.xy[self.off[b]:self.off[b + 1]].reshape(ni, nj, 2) for b, (ni, nj) in enumerate(self.sizes)]
