"""Preconditioner study on the reference's own configurations (host, scipy): BiCGStab iteration counts on the assembled
Picard system of the oracle with point-Jacobi (what the CUDA path uses), line solves along j / along i / the better of the
two per block / the approximate factorisation T_j D^-1 T_i, and ILU(0)-like (the reference's GMRES.zig:199-298 class).
Usage: python scripts/precond_probe.py [t106_white|ls89x4_white] [picard_iterations_before]"""
import os, sys, json
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from util import load_fixture
from turbomesh_b200 import synthetic
from oracle import oracle as orc

name = sys.argv[1] if len(sys.argv) > 1 else "t106_white"
n_before = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec, z, meta = load_fixture(name)
mesh = synthetic.materialize(spec, orc.tfi)
O = orc.System(mesh, orc.tight_options(control_function="white", ds_target=meta["ds_target"], theta_target=meta["theta_target"]))
for n in range(n_before):
    O.iterate(n)
O.fill(n_before)
p, idx, v, rx, ry = O.csr()
A = sp.csr_matrix((v, idx, p))
dof = A.shape[0]
d = A.diagonal()
A = sp.diags(1.0 / d) @ A          # row scaling (what the CUDA path solves)
A = A.tocsr()
bx, by = rx / d, ry / d
shapes = [b.points.shape[:2] for b in mesh.blocks]
offs = np.concatenate([[0], np.cumsum([a * b for a, b in shapes])])
x0 = np.concatenate([b.points.reshape(-1, 2) for b in mesh.blocks])
print(name, "dof", dof, "blocks", shapes)

blk = np.zeros(dof, dtype=np.int64); ii = np.zeros(dof, dtype=np.int64); jj = np.zeros(dof, dtype=np.int64)
for k, (ni, nj) in enumerate(shapes):
    r = np.arange(offs[k], offs[k + 1])
    blk[r] = k; ii[r] = (r - offs[k]) // nj; jj[r] = (r - offs[k]) % nj
C = A.tocoo()
same = blk[C.row] == blk[C.col]
di, dj = ii[C.col] - ii[C.row], jj[C.col] - jj[C.row]
def sub(mask):
    return sp.csr_matrix((C.data[mask], (C.row[mask], C.col[mask])), shape=A.shape)
Tj = sub(same & (di == 0) & (np.abs(dj) <= 1))
Ti = sub(same & (dj == 0) & (np.abs(di) <= 1))
# strength per block: mean |off-diagonal| along each direction
sj = np.zeros(len(shapes)); si = np.zeros(len(shapes))
mj = same & (di == 0) & (np.abs(dj) == 1); mi = same & (dj == 0) & (np.abs(di) == 1)
np.add.at(sj, blk[C.row[mj]], np.abs(C.data[mj])); np.add.at(si, blk[C.row[mi]], np.abs(C.data[mi]))
print("coupling j / i per block:", [f"{a / max(b, 1e-300):.2f}" for a, b in zip(sj, si)])
pick_j = sj >= si
Tbest = sub(same & (((di == 0) & (np.abs(dj) <= 1) & pick_j[blk[C.row]]) | ((dj == 0) & (np.abs(di) <= 1) & ~pick_j[blk[C.row]])))
lu_j, lu_i, lu_b = spl.splu(Tj.tocsc()), spl.splu(Ti.tocsc()), spl.splu(Tbest.tocsc())
ilu = spl.spilu(A.tocsc(), fill_factor=1.0, drop_tol=0.0)
I = sp.identity(dof, format="csr")

def bicgstab(M, b, x, rtol=1e-6, atol=1e-8, cap=2000):
    """right-preconditioned BiCGStab; returns iterations (2 operator applications each)"""
    r = b - A @ x
    tol = max(atol, rtol * np.linalg.norm(b))
    rh = r.copy(); rho = alpha = om = 1.0; vv = np.zeros_like(b); pp = np.zeros_like(b)
    for it in range(1, cap + 1):
        rho1 = rh @ r
        beta = (rho1 / rho) * (alpha / om); rho = rho1
        pp = r + beta * (pp - om * vv)
        ph = M(pp); vv = A @ ph
        alpha = rho / (rh @ vv)
        s = r - alpha * vv
        if np.linalg.norm(s) <= tol:
            return it - 0.5
        sh = M(s); t = A @ sh
        om = (t @ s) / (t @ t)
        x = x + alpha * ph + om * sh
        r = s - om * t
        if np.linalg.norm(r) <= tol:
            return it
    return cap

precs = {
    "jacobi": lambda r: r,
    "line j": lu_j.solve,
    "line i": lu_i.solve,
    "line, stronger direction per block": lu_b.solve,
    "T_j D^-1 T_i": lambda r: lu_i.solve(lu_j.solve(r)),          # D = I after the row scaling
    "T_i D^-1 T_j": lambda r: lu_j.solve(lu_i.solve(r)),
    "ilu(0)-class (spilu, fill 1)": ilu.solve,
}
for tolname, (rt, at) in {"reference defaults 1e-6/1e-8": (1e-6, 1e-8), "tight 0/1e-13": (0.0, 1e-13)}.items():
    print(tolname)
    for nm, M in precs.items():
        kx = bicgstab(M, bx, x0[:, 0].copy(), rt, at); ky = bicgstab(M, by, x0[:, 1].copy(), rt, at)
        print(f"  {nm:38s} x {kx:7.1f}  y {ky:7.1f} iterations")

# ---- two-level: piecewise-constant aggregates a x a inside each block, Galerkin coarse operator, direct coarse solve ----
kinds_free = np.abs(A - I).sum(axis=1).A1 > 0          # rows that are not plain identity rows (fixed nodes)
for a in (4, 8, 16):
    agg = np.full(dof, -1, dtype=np.int64); nagg = 0
    for k, (ni, nj) in enumerate(shapes):
        gi, gj = (ni + a - 1) // a, (nj + a - 1) // a
        r = np.arange(offs[k], offs[k + 1])
        agg[r] = nagg + (ii[r] // a) * gj + jj[r] // a
        nagg += gi * gj
    rows = np.nonzero(kinds_free)[0]
    P = sp.csr_matrix((np.ones(len(rows)), (rows, agg[rows])), shape=(dof, nagg))
    keep = np.asarray(P.sum(axis=0)).ravel() > 0
    P = P[:, keep]
    Ac = (P.T @ A @ P).tocsc()
    lu_c = spl.splu(Ac)
    def add(r, P=P, lu_c=lu_c):
        return r + P @ lu_c.solve(P.T @ r)
    def mult(r, P=P, lu_c=lu_c):
        zc = P @ lu_c.solve(P.T @ r)
        return zc + (r - A @ zc)
    def mult_sym(r, P=P, lu_c=lu_c):
        z = r.copy()
        z = z + P @ lu_c.solve(P.T @ (r - A @ z))
        return z + (r - A @ z)
    print(f"aggregates {a}x{a}: {P.shape[1]} coarse unknowns")
    for tolname, (rt, at) in {"reference defaults": (1e-6, 1e-8), "tight": (0.0, 1e-13)}.items():
        for nm, M in {"additive": add, "coarse then jacobi": mult, "jacobi, coarse, jacobi": mult_sym}.items():
            kx = bicgstab(M, bx, x0[:, 0].copy(), rt, at); ky = bicgstab(M, by, x0[:, 1].copy(), rt, at)
            print(f"  {tolname:20s} {nm:24s} x {kx:7.1f}  y {ky:7.1f} iterations")

# ---- two-level with bilinear interpolation from every a-th node of each block (+ the last line), Galerkin, direct coarse solve ----
def hat(n, a):
    """(n x nc) linear interpolation from coarse lines 0, a, 2a, ..., n-1"""
    cl = sorted(set(list(range(0, n, a)) + [n - 1]))
    if len(cl) > 2 and cl[-1] - cl[-2] < a // 2:
        cl.pop(-2)
    W = np.zeros((n, len(cl)))
    for q in range(len(cl) - 1):
        lo, hi = cl[q], cl[q + 1]
        for t in range(lo, hi + 1):
            w = (t - lo) / (hi - lo); W[t, q] = max(W[t, q], 1 - w) if t == lo else 1 - w; W[t, q + 1] = w
    return sp.csr_matrix(W)
for a in (4, 8, 16):
    Ps = [sp.kron(hat(ni, a), hat(nj, a), format="csr") for ni, nj in shapes]
    P = sp.block_diag(Ps, format="csr")
    P = sp.diags(kinds_free.astype(float)) @ P
    keep = np.asarray(abs(P).sum(axis=0)).ravel() > 0
    P = P[:, keep].tocsr()
    Ac = (P.T @ A @ P).tocsc()
    lu_c = spl.splu(Ac)
    def add(r, P=P, lu_c=lu_c):
        return r + P @ lu_c.solve(P.T @ r)
    def mult(r, P=P, lu_c=lu_c):
        zc = P @ lu_c.solve(P.T @ r)
        return zc + (r - A @ zc)
    print(f"bilinear {a}: {P.shape[1]} coarse unknowns")
    for tolname, (rt, at) in {"reference defaults": (1e-6, 1e-8), "tight": (0.0, 1e-13)}.items():
        for nm, M in {"additive": add, "coarse then jacobi": mult}.items():
            kx = bicgstab(M, bx, x0[:, 0].copy(), rt, at); ky = bicgstab(M, by, x0[:, 1].copy(), rt, at)
            print(f"  {tolname:20s} {nm:24s} x {kx:7.1f}  y {ky:7.1f} iterations")
