// mg_kernels.cuh -- device code of the two FAS multigrid hierarchies: transfers of the non-nested single-block hierarchy,
// nested multi-block transfers (block kernels + row tables), Anderson acceleration on level-1 samples, the single-CTA
// multi-sweep kernel of tiny levels and the side-length reduction that decides the semi-coarsening.
#pragma once
#include "kernels.cuh"

namespace tmesh {

// ---------------------------------------------------------------------------------------------------
// Geometric FAS multigrid for a block whose boundary nodes are all fixed (the single-block configuration).
// Levels are NOT nested (8192 nodes = 8191 intervals is prime): a coarse level has about half the nodes per direction
// and all transfers are bilinear interpolations in index space -- the coordinates are smooth functions of (xi, eta), and
// the Winslow row (undivided differences) of a smooth field scales by (r_xi r_eta)^2 between levels.
// ---------------------------------------------------------------------------------------------------
struct MgLevelDims {
    int ni_f, nj_f, ni_c, nj_c;
    double r_i, r_j;  // (ni_f-1)/(ni_c-1), (nj_f-1)/(nj_c-1)
};

__device__ __forceinline__ double2 bilerp(const double2* __restrict__ f, int nj, int i0, int j0, double ti, double tj) {
    const double2 a = f[(size_t)i0 * nj + j0], b = f[(size_t)i0 * nj + j0 + 1], c = f[(size_t)(i0 + 1) * nj + j0], d = f[(size_t)(i0 + 1) * nj + j0 + 1];
    const double w00 = (1.0 - ti) * (1.0 - tj), w01 = (1.0 - ti) * tj, w10 = ti * (1.0 - tj), w11 = ti * tj;
    return make_double2(w00 * a.x + w01 * b.x + w10 * c.x + w11 * d.x, w00 * a.y + w01 * b.y + w10 * c.y + w11 * d.y);
}

// coarse <- fine: the iterate by interpolation (boundary included); the residual by interpolation of its
// [1 2 1]x[1 2 1]/16 average (full-weighting-like), scaled to coarse row units.  One thread per coarse node.
__global__ void mg_restrict_kernel(MgLevelDims d, const double2* __restrict__ u_f, const double2* __restrict__ res_f, double2* __restrict__ u_c,
                                   double2* __restrict__ e_c, double2* __restrict__ res_c, double scale) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x, I = blockIdx.y;
    if (J >= d.nj_c || I >= d.ni_c) return;
    const double xi = fmin(I * d.r_i, (double)(d.ni_f - 1)), eta = fmin(J * d.r_j, (double)(d.nj_f - 1));
    const int i0 = min((int)xi, d.ni_f - 2), j0 = min((int)eta, d.nj_f - 2);
    const double ti = xi - i0, tj = eta - j0;
    const size_t k = (size_t)I * d.nj_c + J;
    const double2 uc = bilerp(u_f, d.nj_f, i0, j0, ti, tj);
    u_c[k] = uc;
    e_c[k] = uc;
    double2 r = make_double2(0.0, 0.0);
    if (I > 0 && I < d.ni_c - 1 && J > 0 && J < d.nj_c - 1) {
        // averaged residual at the four surrounding fine nodes (the fine residual is zero on the block boundary)
        double2 acc[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int ic = i0 + a, jc = j0 + b;
                double sx = 0.0, sy = 0.0;
#pragma unroll
                for (int p = -1; p <= 1; ++p)
#pragma unroll
                    for (int q = -1; q <= 1; ++q) {
                        const int ii = ic + p, jj = jc + q;
                        if (ii < 0 || ii >= d.ni_f || jj < 0 || jj >= d.nj_f) continue;
                        const double w = (p == 0 ? 2.0 : 1.0) * (q == 0 ? 2.0 : 1.0) * (1.0 / 16.0);
                        const double2 v = res_f[(size_t)ii * d.nj_f + jj];
                        sx += w * v.x; sy += w * v.y;
                    }
                acc[a][b] = make_double2(sx, sy);
            }
        const double w00 = (1.0 - ti) * (1.0 - tj), w01 = (1.0 - ti) * tj, w10 = ti * (1.0 - tj), w11 = ti * tj;
        r.x = scale * (w00 * acc[0][0].x + w01 * acc[0][1].x + w10 * acc[1][0].x + w11 * acc[1][1].x);
        r.y = scale * (w00 * acc[0][0].y + w01 * acc[0][1].y + w10 * acc[1][0].y + w11 * acc[1][1].y);
    }
    res_c[k] = r;
}

// Anderson acceleration of the single-block cycle (see aa_* kernels below): the fine iterate sampled by interpolation on
// the nodes of level 1; g_new = sample, f_new = sample - x_prev.  One thread per coarse node.
__global__ void mg_sample_kernel(MgLevelDims d, const double2* __restrict__ u_f, const double2* __restrict__ x_prev, double2* __restrict__ g_new,
                                 double2* __restrict__ f_new) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x, I = blockIdx.y;
    if (J >= d.nj_c || I >= d.ni_c) return;
    const double xi = fmin(I * d.r_i, (double)(d.ni_f - 1)), eta = fmin(J * d.r_j, (double)(d.nj_f - 1));
    const int i0 = min((int)xi, d.ni_f - 2), j0 = min((int)eta, d.nj_f - 2);
    const size_t k = (size_t)I * d.nj_c + J;
    const double2 g = bilerp(u_f, d.nj_f, i0, j0, xi - i0, eta - j0);
    const double2 x = x_prev[k];
    g_new[k] = g;
    f_new[k] = make_double2(g.x - x.x, g.y - x.y);
}

// coarse right-hand side of FAS: tau_c = row_c(I u_f) + restricted residual.  `rel_c` holds row_c(I u_f) (MODE_REL output)
// on interior nodes; `res_c` the scaled restricted residual; the result overwrites res_c.
__global__ void mg_coarse_rhs_kernel(int ni, int nj, const double2* __restrict__ rel_c, double2* __restrict__ res_c) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= nj || i >= ni) return;
    const size_t k = (size_t)i * nj + j;
    if (i == 0 || i == ni - 1 || j == 0 || j == nj - 1) { res_c[k] = make_double2(0.0, 0.0); return; }
    const double2 a = rel_c[k], r = res_c[k];
    res_c[k] = make_double2(a.x + r.x, a.y + r.y);
}

// fine += interpolated coarse correction (u_c - e_c); interior fine nodes only.  One thread per fine node.
__global__ void mg_prolong_kernel(MgLevelDims d, const double2* __restrict__ u_c, const double2* __restrict__ e_c, double2* __restrict__ u_f) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j <= 0 || j >= d.nj_f - 1 || i <= 0 || i >= d.ni_f - 1) return;
    const double xi = i / d.r_i, eta = j / d.r_j;
    const int i0 = min((int)xi, d.ni_c - 2), j0 = min((int)eta, d.nj_c - 2);
    const double ti = xi - i0, tj = eta - j0;
    const double2 a = bilerp(u_c, d.nj_c, i0, j0, ti, tj), b = bilerp(e_c, d.nj_c, i0, j0, ti, tj);
    const size_t k = (size_t)i * d.nj_f + j;
    double2 v = u_f[k];
    v.x += a.x - b.x; v.y += a.y - b.y;
    u_f[k] = v;
}

// ---------------------------------------------------------------------------------------------------
// Multi-block / multi-GPU FAS multigrid: NESTED coarsening (every coarse node is a fine node), per block by a factor
// f_i, f_j in {1, 2} per direction.  Block-local transfers; rows that straddle blocks (interface, junction, sliding
// rows) are restricted by a flat table (RestrictRow), and the copies of a node are re-derived from their root after
// every transfer, so all copies stay bit-consistent on every level.
// ---------------------------------------------------------------------------------------------------
struct BlockXfer {
    int64_t off_f, off_c;          // local offsets of the block on the fine / coarse level
    int32_t ni_f, nj_f, ni_c, nj_c;
    int32_t fi, fj;                // fine index = f * coarse index
    int32_t slide, _pad;           // bit 0..3: the whole side j_min (i = 0) / j_max (i = ni-1) / i_min (j = 0) / i_max (j = nj-1) slides
};
struct RestrictRow {               // rhs_c[dst] = sum_k w[k] * res_f[src[k]]  (scale and sign folded into w)
    int64_t dst;
    int64_t src[9];
    double w[9];
    int32_t n, _pad;
};

// coarse <- fine for one block: iterate by injection (all nodes, so that copies stay exact copies); residual of the
// interior rows by full weighting in the coarsened directions, times (f_i f_j)^2 (the undivided Winslow row of a smooth
// field scales like h_xi^2 h_eta^2).  The result R goes to the coarse level's scratch field; the FAS right-hand side
// tau_c = row_c(I u_f) - R is then produced by ONE launch of the coarse rows in MODE_REL with R as "rhs".
constexpr int MGB_ROWS = 1;  // rows per CTA in the block transfer kernels (marching several rows per thread measured slower: less memory-level parallelism)
__global__ void __launch_bounds__(128) mgb_restrict_kernel(const BlockXfer* __restrict__ blocks /* one per blockIdx.z */, const double2* __restrict__ u_f, const double2* __restrict__ res_f, double2* __restrict__ u_c,
                                                           double2* e_c, double2* __restrict__ rhs_c,
                                                           unsigned long long* __restrict__ change /* may be NULL */,
                                                           const double2* e_prev /* the previous cycle's restricted iterate (may alias e_c) */) {
    const BlockXfer b = blocks[blockIdx.z];
    const double scale = (double)(b.fi * b.fj) * (double)(b.fi * b.fj);  // restricted residual, in coarse row units
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    const int I_end = min((int)(blockIdx.y + 1) * MGB_ROWS, b.ni_c);
    double dmax = 0.0;
    if (J < b.nj_c) {
        const int j = J * b.fj;
        const int pi = b.fi == 2 ? 1 : 0, pj = b.fj == 2 ? 1 : 0;
        // Next to a sliding (Neumann-type) side the boundary unknown follows its inner neighbour, so the coarse boundary
        // node carries no row of its own: the quarter of the first interior row's residual that full weighting would
        // send there belongs to this row instead (the Galerkin restriction after eliminating y_0 = y_1).  That residual
        // is what drives the sliding modes; with the plain weights the coarse correction is half of what is needed.
        const double wjm = (J == 1 && (b.slide & 4)) ? 0.5 : 0.25, wjp = (J == b.nj_c - 2 && (b.slide & 8)) ? 0.5 : 0.25;
        for (int I = blockIdx.y * MGB_ROWS; I < I_end; ++I) {
            const int i = I * b.fi;
            const size_t kc = (size_t)b.off_c + (size_t)I * b.nj_c + J;
            const size_t kf = (size_t)b.off_f + (size_t)i * b.nj_f + j;
            const double2 uc = u_f[kf];
            if (change) {  // how far this node moved since the previous cycle's restriction (the cycle's convergence measure)
                const double2 old = e_prev[kc];
                dmax = fmax(dmax, fmax(fabs(uc.x - old.x), fabs(uc.y - old.y)));
            }
            u_c[kc] = uc;
            e_c[kc] = uc;
            double2 r = make_double2(0.0, 0.0);
            if (I > 0 && I < b.ni_c - 1 && J > 0 && J < b.nj_c - 1) {
                const double wim = (I == 1 && (b.slide & 1)) ? 0.5 : 0.25, wip = (I == b.ni_c - 2 && (b.slide & 2)) ? 0.5 : 0.25;
                for (int p = -pi; p <= pi; ++p)
                    for (int q = -pj; q <= pj; ++q) {
                        const double w = (pi ? (p == 0 ? 0.5 : (p < 0 ? wim : wip)) : 1.0) * (pj ? (q == 0 ? 0.5 : (q < 0 ? wjm : wjp)) : 1.0);
                        const double2 v = res_f[kf + (long long)p * b.nj_f + q];
                        r.x += w * v.x; r.y += w * v.y;
                    }
                r.x *= scale; r.y *= scale;
            }
            rhs_c[kc] = r;
        }
    }
    if (change) {
        dmax = warp_max(dmax);
        // non-negative doubles order like integers; the plain (possibly stale) read filters out nearly every atomic
        const unsigned long long bits = (unsigned long long)__double_as_longlong(dmax);
        if ((threadIdx.x & 31) == 0 && bits > *(volatile unsigned long long*)change) atomicMax(change, bits);
    }
}

// the cycle's convergence measure as one record of per-CTA partials (reduce_kernel / all-reduce take it from there)
__global__ void mgb_change_kernel(unsigned long long* __restrict__ change, double* __restrict__ partials, int n_records) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_records) return;
    double* p = partials + (size_t)k * 5;
    p[0] = p[1] = p[2] = p[3] = 0.0;
    p[4] = k == 0 ? __longlong_as_double((long long)*change) : 0.0;
}

__global__ void mgb_restrict_rows_kernel(const RestrictRow* __restrict__ rows, int n, const double2* __restrict__ res_f, double2* __restrict__ rhs_c) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const RestrictRow row = rows[k];
    double2 r = make_double2(0.0, 0.0);
    for (int q = 0; q < row.n; ++q) {
        const double2 v = res_f[row.src[q]];
        r.x += row.w[q] * v.x; r.y += row.w[q] * v.y;
    }
    rhs_c[row.dst] = r;
}

// fine += bilinear interpolation of the coarse correction (u_c - e_c).  Interior nodes per block; the free rows on
// block boundaries (interface, junction, sliding rows) by table, interpolating along their boundary line.  Fixed nodes
// are never touched (a fixed wall node next to a moving junction must not pick up half of its correction), and copies
// are re-derived from their roots afterwards.
__device__ __forceinline__ double2 mgb_correction(const BlockXfer& b, int i, int j, const double2* __restrict__ u_c, const double2* __restrict__ e_c) {
    const int I0 = i / b.fi, J0 = j / b.fj;
    const bool hi = (i % b.fi) != 0, hj = (j % b.fj) != 0;  // halfway between two coarse nodes
    const double2* uc = u_c + b.off_c;
    const double2* ec = e_c + b.off_c;
    auto corr = [&](int I, int J) {
        const size_t k = (size_t)I * b.nj_c + J;
        const double2 a = uc[k], e = ec[k];
        return make_double2(a.x - e.x, a.y - e.y);
    };
    double2 c = corr(I0, J0);
    if (hi && hj) {
        const double2 c1 = corr(I0 + 1, J0), c2 = corr(I0, J0 + 1), c3 = corr(I0 + 1, J0 + 1);
        c = make_double2(0.25 * ((c.x + c3.x) + (c1.x + c2.x)), 0.25 * ((c.y + c3.y) + (c1.y + c2.y)));
    } else if (hi) {
        const double2 c1 = corr(I0 + 1, J0);
        c = make_double2(0.5 * (c.x + c1.x), 0.5 * (c.y + c1.y));
    } else if (hj) {
        const double2 c2 = corr(I0, J0 + 1);
        c = make_double2(0.5 * (c.x + c2.x), 0.5 * (c.y + c2.y));
    }
    return c;
}
__global__ void __launch_bounds__(128) mgb_prolong_kernel(const BlockXfer* __restrict__ blocks /* one per blockIdx.z */, const double2* __restrict__ u_c,
                                                          const double2* __restrict__ e_c, double2* __restrict__ u_f) {
    const BlockXfer b = blocks[blockIdx.z];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j <= 0 || j >= b.nj_f - 1) return;
    const int i_begin = max(1, (int)blockIdx.y * MGB_ROWS), i_end = min((int)(blockIdx.y + 1) * MGB_ROWS, b.ni_f - 1);
    for (int i = i_begin; i < i_end; ++i) {
        const double2 c = mgb_correction(b, i, j, u_c, e_c);
        const size_t k = (size_t)b.off_f + (size_t)i * b.nj_f + j;
        double2 v = u_f[k];
        v.x += c.x; v.y += c.y;
        u_f[k] = v;
    }
}
// The common case f_i = f_j = 2: one thread per COARSE cell (I, J) updates the 2 x 2 fine nodes (2I..2I+1, 2J..2J+1) from
// the four corner corrections -- a quarter of the threads, no redundant coarse loads, 32 contiguous bytes per fine row.
__global__ void __launch_bounds__(128) mgb_prolong_2x2_kernel(const BlockXfer* __restrict__ blocks /* one per blockIdx.z */, const double2* __restrict__ u_c,
                                                              const double2* __restrict__ e_c, double2* __restrict__ u_f) {
    const BlockXfer b = blocks[blockIdx.z];
    const int J = blockIdx.x * blockDim.x + threadIdx.x, I = blockIdx.y;
    if (J >= b.nj_c - 1 || I >= b.ni_c - 1) return;
    const double2* uc = u_c + b.off_c;
    const double2* ec = e_c + b.off_c;
    auto corr = [&](int II, int JJ) {
        const size_t k = (size_t)II * b.nj_c + JJ;
        const double2 a = uc[k], e = ec[k];
        return make_double2(a.x - e.x, a.y - e.y);
    };
    const double2 c = corr(I, J), c1 = corr(I + 1, J), c2 = corr(I, J + 1), c3 = corr(I + 1, J + 1);
    double2* f = u_f + b.off_f + (size_t)(2 * I) * b.nj_f + 2 * J;
    auto add = [](double2* p, double dx, double dy) { double2 v = *p; v.x += dx; v.y += dy; *p = v; };
    if (I > 0 && J > 0) add(f, c.x, c.y);
    if (I > 0) add(f + 1, 0.5 * (c.x + c2.x), 0.5 * (c.y + c2.y));
    if (J > 0) add(f + b.nj_f, 0.5 * (c.x + c1.x), 0.5 * (c.y + c1.y));
    add(f + b.nj_f + 1, 0.25 * ((c.x + c3.x) + (c1.x + c2.x)), 0.25 * ((c.y + c3.y) + (c1.y + c2.y)));
}
__global__ void mgb_prolong_rows_kernel(const BlockXfer* __restrict__ blocks, int n_blocks, const SmoothedRow* __restrict__ srows, int n_s,
                                        const JunctionRow* __restrict__ jrows, int n_j, const SlidingRow* __restrict__ lrows, int n_l,
                                        const double2* __restrict__ u_c, const double2* __restrict__ e_c, double2* __restrict__ u_f) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_s + n_j + n_l) return;
    const int64_t self = r < n_s ? srows[r].g0 : (r < n_s + n_j ? jrows[r - n_s].self : lrows[r - n_s - n_j].self);
    int lo = 0, hi = n_blocks - 1;  // own blocks are stored in ascending offset order
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (blocks[mid].off_f <= self) lo = mid; else hi = mid - 1;
    }
    const BlockXfer b = blocks[lo];
    const int64_t local = self - b.off_f;
    const int i = (int)(local / b.nj_f), j = (int)(local - (int64_t)i * b.nj_f);
    const double2 c = mgb_correction(b, i, j, u_c, e_c);
    double2 v = u_f[self];
    v.x += c.x; v.y += c.y;
    u_f[self] = v;
}

// ---------------------------------------------------------------------------------------------------
// Anderson acceleration of the multigrid cycle on level-1 samples.  The slow modes of the cycle are smooth (on the
// cascade: the global shift anchored only by the plate, whose tip singularity every level resolves differently), so their
// history is kept where it is cheap -- on the nodes of level 1, a quarter of the mesh.  One "iteration" runs from the
// restriction point of a cycle to that of the next: X_j = the (accelerated) iterate sampled there, G_j = what the cycle
// made of it one cycle later, F_j = G_j - X_j.  alpha minimises |sum alpha_j F_j| subject to sum alpha_j = 1 over the
// last q <= 3 iterations and the new iterate is sum alpha_j G_j: the difference to the current one lives on level 1 and is
// interpolated to the fine mesh like a coarse-grid correction.
// ---------------------------------------------------------------------------------------------------
constexpr int AA_MAX = 5;                            // residuals in the window
constexpr int AA_GRAM = AA_MAX * (AA_MAX + 1) / 2;   // upper triangle of the Gram matrix, row-major
struct AaFields { double2* G[AA_MAX]; double2* F[AA_MAX]; int q; };  // chronological, index q-1 = newest
// samples the fine iterate on the level-1 nodes of one block: G_new = sample, F_new = sample - X_prev
__global__ void __launch_bounds__(128) aa_sample_kernel(const BlockXfer* __restrict__ blocks /* one per blockIdx.z */, const double2* __restrict__ u_f,
                                                        const double2* __restrict__ x_prev, double2* __restrict__ g_new, double2* __restrict__ f_new) {
    const BlockXfer b = blocks[blockIdx.z];
    const int J = blockIdx.x * blockDim.x + threadIdx.x, I = blockIdx.y;
    if (J >= b.nj_c || I >= b.ni_c) return;
    const size_t kc = (size_t)b.off_c + (size_t)I * b.nj_c + J;
    const double2 g = u_f[(size_t)b.off_f + (size_t)(I * b.fi) * b.nj_f + (size_t)J * b.fj];
    const double2 x = x_prev[kc];
    g_new[kc] = g;
    f_new[kc] = make_double2(g.x - x.x, g.y - x.y);
}
__global__ void __launch_bounds__(256) aa_gram_kernel(int64_t n, AaFields h, double* __restrict__ partials /* grid x AA_GRAM */) {
    double g[AA_GRAM];
#pragma unroll
    for (int e = 0; e < AA_GRAM; ++e) g[e] = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (int64_t)gridDim.x * 256) {
        double2 f[AA_MAX];
#pragma unroll
        for (int i = 0; i < AA_MAX; ++i) f[i] = i < h.q ? h.F[i][k] : make_double2(0.0, 0.0);
        int e = 0;
#pragma unroll
        for (int a = 0; a < AA_MAX; ++a)
#pragma unroll
            for (int c = a; c < AA_MAX; ++c) g[e++] += f[a].x * f[c].x + f[a].y * f[c].y;
    }
    __shared__ double sh[AA_GRAM][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < AA_GRAM; ++e) { g[e] = warp_sum(g[e]); if (lane == 0) sh[e][w] = g[e]; }
    __syncthreads();
    if (threadIdx.x < AA_GRAM) {
        double s = 0.0;
        for (int q = 0; q < 8; ++q) s += sh[threadIdx.x][q];
        partials[(size_t)blockIdx.x * AA_GRAM + threadIdx.x] = s;
    }
}
__global__ void aa_reduce_kernel(const double* __restrict__ partials, int n_part, double* __restrict__ gram) {  // one CTA of 32 * AA_GRAM threads, fixed order
    __shared__ double sh[AA_GRAM][32];
    const int e = threadIdx.x / 32, l = threadIdx.x & 31;
    double s = 0.0;
    for (int k = l; k < n_part; k += 32) s += partials[(size_t)k * AA_GRAM + e];
    sh[e][l] = s;
    __syncthreads();
    if (l == 0) { double t = 0.0; for (int q = 0; q < 32; ++q) t += sh[e][q]; gram[e] = t; }
}
struct SumPtrs { double* p[16]; };
__global__ void combine_sum_kernel(SumPtrs v, int n_ranks, int count) {  // in-process emulation of the all-reduce
    const int k = threadIdx.x;
    if (k >= count) return;
    double s = 0.0;
    for (int r = 0; r < n_ranks; ++r) s += v.p[r][k];
    __syncthreads();
    for (int r = 0; r < n_ranks; ++r) v.p[r][k] = s;
}
__global__ void __launch_bounds__(256) combine_sum_fields_kernel(SumPtrs v, int n_ranks, int64_t count /* doubles */) {
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < count; k += (int64_t)gridDim.x * 256) {
        double s = 0.0;
        for (int r = 0; r < n_ranks; ++r) s += v.p[r][k];
        for (int r = 0; r < n_ranks; ++r) v.p[r][k] = s;
    }
}
// alpha_0..alpha_{q-1} (sum 1); falls back to "newest only" (no extrapolation) when the window is short, the
// least-squares problem is degenerate, the weights are wild or the newest residual grew
__global__ void aa_solve_kernel(const double* __restrict__ gram, int q, double* __restrict__ alpha) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int i = 0; i < AA_MAX; ++i) alpha[i] = 0.0;
    alpha[q - 1] = 1.0;
    if (q < 2) return;
    double G[AA_MAX][AA_MAX];
    {
        int e = 0;
        for (int a = 0; a < AA_MAX; ++a)
            for (int c = a; c < AA_MAX; ++c) { G[a][c] = gram[e]; G[c][a] = gram[e]; ++e; }
    }
    if (G[q - 1][q - 1] > 4.0 * G[q - 2][q - 2]) return;  // the residual doubled: the history is not trustworthy
    double z[AA_MAX];
    double tr = 0.0;
    for (int i = 0; i < q; ++i) { z[i] = 1.0; tr += G[i][i]; }
    if (!(tr > 0.0)) return;
    for (int i = 0; i < q; ++i) G[i][i] += 1e-12 * tr;   // Tikhonov guard against a degenerate window
    for (int c = 0; c < q; ++c) {                        // Gaussian elimination with partial pivoting, G z = 1
        int piv = c;
        for (int r = c + 1; r < q; ++r) if (fabs(G[r][c]) > fabs(G[piv][c])) piv = r;
        if (fabs(G[piv][c]) < 1e-300) return;
        if (piv != c) { for (int k = 0; k < q; ++k) { const double t = G[c][k]; G[c][k] = G[piv][k]; G[piv][k] = t; } const double t = z[c]; z[c] = z[piv]; z[piv] = t; }
        for (int r = c + 1; r < q; ++r) {
            const double f = G[r][c] / G[c][c];
            for (int k = c; k < q; ++k) G[r][k] -= f * G[c][k];
            z[r] -= f * z[c];
        }
    }
    for (int r = q - 1; r >= 0; --r) {
        double t = z[r];
        for (int k = r + 1; k < q; ++k) t -= G[r][k] * z[k];
        z[r] = t / G[r][r];
    }
    double sum = 0.0;
    for (int i = 0; i < q; ++i) sum += z[i];
    if (!(fabs(sum) > 1e-300)) return;
    double a[AA_MAX];
    for (int i = 0; i < q; ++i) { a[i] = z[i] / sum; if (!(fabs(a[i]) < 20.0)) return; }
    for (int i = 0; i < q; ++i) alpha[i] = a[i];
}
// x_new = sum alpha_j G_j ; d = x_new - G_newest (the extrapolation, interpolated to the fine mesh next) ; x_store = x_new
__global__ void __launch_bounds__(256) aa_combine_kernel(int64_t n, AaFields h, const double* __restrict__ alpha, double2* __restrict__ d, double2* __restrict__ x_store) {
    double a[AA_MAX];
#pragma unroll
    for (int i = 0; i < AA_MAX; ++i) a[i] = alpha[i];
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (int64_t)gridDim.x * 256) {
        const double2 gn = h.G[h.q - 1][k];
        // written as newest + sum alpha_j (G_j - newest): exactly zero wherever all samples agree (fixed nodes)
        double dx = 0.0, dy = 0.0;
#pragma unroll
        for (int i = 0; i < AA_MAX; ++i)
            if (i < h.q - 1) { const double2 t = h.G[i][k]; dx += a[i] * (t.x - gn.x); dy += a[i] * (t.y - gn.y); }
        d[k] = make_double2(dx, dy);
        x_store[k] = make_double2(gn.x + dx, gn.y + dy);
    }
}

// ---------------------------------------------------------------------------------------------------
// Tiny multigrid levels (a few thousand nodes: the coarsest levels of a block-structured hierarchy cannot get smaller
// than 3x3 nodes per block): ALL sweeps of a visit in ONE launch of ONE CTA.  Launch latency, not bandwidth, is what
// such levels cost; the level's data lives in L2 and __syncthreads() separates the sweeps.  Interior rows come from a flat
// node list, boundary rows reuse boundary_rows(); ping-pong between xa and xb, the result is in xa after an even and
// in xb after an odd number of sweeps.
// ---------------------------------------------------------------------------------------------------
struct SmallNode { int64_t idx; int32_t block, i, j, _pad; };  // an interior node: local index, owning block, (i, j)
__global__ void __launch_bounds__(1024) winslow_small_level_kernel(const SmallNode* nodes, int n_nodes, const DevBlock* blocks, BndArgs bnd, double2* xa, double2* xb,
                                                                   const double2* rhs /* may be NULL */, double omega, int sweeps) {
    const int n_rows = bnd.n_s + bnd.n_j + bnd.n_l;
    for (int sw = 0; sw < sweeps; ++sw) {
        const double2* u = (sw & 1) ? xb : xa;
        double2* out = (sw & 1) ? xa : xb;
        for (int k = threadIdx.x; k < n_nodes; k += blockDim.x) {
            const SmallNode nd = nodes[k];
            const DevBlock b = blocks[nd.block];
            const double2* c = u + nd.idx;
            const int nj = b.nj;
            const double2 C = c[0], W = c[-nj], E = c[nj], S = c[-1], N = c[1];
            const double2 SW = c[-nj - 1], NW = c[-nj + 1], SE = c[nj - 1], NE = c[nj + 1];
            Metric m = metric_terms(W, E, N - S);
            if (rhs && b.slide) {
                if ((nd.i == 1 && (b.slide & 1)) || (nd.i == b.ni - 2 && (b.slide & 2))) m.g11 *= b.tan_i;
                if ((nd.j == 1 && (b.slide & 4)) || (nd.j == nj - 2 && (b.slide & 8))) m.g22 *= b.tan_j;
            }
            double2 rel = row_rel<false>(m, 0.0, 0.0, C, W, E, (N - C) + (S - C), N - S, NE - SE, NW - SW);
            if (rhs) { const double2 f = rhs[nd.idx]; rel.x -= f.x; rel.y -= f.y; }
            out[nd.idx] = row_result<MODE_RELAX>(m, rel, C, omega);
        }
        for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
            if (rhs) boundary_rows<MODE_RELAX, false, false, 0, true>(0, bnd.srows, bnd.n_s, bnd.jrows, bnd.n_j, bnd.lrows, bnd.n_l, bnd.slaves, u, u, nullptr, out, omega, nullptr, nullptr, rhs, r);
            else boundary_rows<MODE_RELAX, false, false, 0, false>(0, bnd.srows, bnd.n_s, bnd.jrows, bnd.n_j, bnd.lrows, bnd.n_l, bnd.slaves, u, u, nullptr, out, omega, nullptr, nullptr, nullptr, r);
        }
        __syncthreads();
    }
}

// polyline length of one side of a block (multigrid: mean cell size per direction decides the semi-coarsening);
// one CTA per (own block, side); out[4 * global block + side]
struct SideLenJob { int64_t off; int32_t ni, nj, block; };
__global__ void __launch_bounds__(256) side_length_kernel(const SideLenJob* __restrict__ jobs, const double2* __restrict__ x, double* __restrict__ out) {
    const SideLenJob jb = jobs[blockIdx.x >> 2];
    const int side = blockIdx.x & 3;  // tm_side order: i_min (j = 0), i_max (j = nj-1), j_min (i = 0), j_max (i = ni-1)
    const int n = side < 2 ? jb.ni : jb.nj;
    const long long stride = side < 2 ? jb.nj : 1;
    const long long base = side == 0 ? 0 : side == 1 ? jb.nj - 1 : side == 2 ? 0 : (long long)(jb.ni - 1) * jb.nj;
    const double2* p = x + jb.off + base;
    double s = 0.0;
    for (int k = threadIdx.x; k + 1 < n; k += 256) {
        const double2 a = p[(long long)k * stride], b = p[(long long)(k + 1) * stride];
        s += sqrt((b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y));
    }
    __shared__ double red[5];
    double sums[4] = {s, 0.0, 0.0, 0.0};
    block_reduce_store<4, 256>(sums, 0.0, red);
    __syncthreads();
    if (threadIdx.x == 0) out[4 * (size_t)jb.block + side] = red[0];
}

}  // namespace tmesh
