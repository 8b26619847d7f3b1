// krylov_kernels.cuh -- the inner linear solves of one outer (Picard) iteration as ONE persistent cooperative kernel.
//
// What it replaces: solver.Solver.solve for x and y (src/core/smoothing/solver.zig:29-94, BiCGStab.zig:279-370 with the
// `diagonal` preconditioner), i.e. > 95 % of the reference's run time.  Round 1 ran a BiCGStab iteration
// as ~12 dependent launches (~50 us per iteration on the reference's mesh sizes, 430 B/node of vector traffic on a batch)
// and solved a batch of independent cuts as ONE system with ONE stopping test.  Here:
//
//   * The mesh is split into its connected COMPONENTS (blocks joined by connections): every 2D cut of a batch is a system
//     of its own with its own Krylov scalars, its own ||b||, tolerance, iteration count and stopping test -- exactly what
//     the reference does when it meshes the cuts one after the other.
//   * CTAs are partitioned into GROUPS; a group works through its components one after the other, each to convergence, so the
//     working set of a solve (10 fields x 16 B x 25 k nodes = 4 MB for a T106 cut) stays in the 126 MB L2 while ~30 cuts
//     are in flight: the iteration is bound by L2 latency / bandwidth, not by HBM and not by launch cadence.
//   * All phases of an iteration (p update, v = A p, s update, t = A s, r update) run inside the kernel, separated by a
//     group barrier (sense-reversing counter in global memory, release / acquire at gpu scope; co-residency comes from the
//     cooperative launch); the partial dot products of a phase are combined redundantly by every CTA in a fixed order, so
//     all CTAs of a group hold bit-identical scalars without a second barrier.
//   * Rows are evaluated by their owner: interior nodes by warp tiles (32 columns x a few rows, 3-row register window on
//     the vector and on the lagged coordinates), interface / junction / sliding rows one per thread; `connected` copies
//     are mirrored by the thread of their root, so no separate copy pass exists.
//
// The arithmetic of a row is the one of kernels.cuh (difference form, row-scaled system D^-1 A x = D^-1 b).
#pragma once
#include "kernels.cuh"

namespace tmesh {

constexpr int K_THREADS = 256;
constexpr int K_WARPS = K_THREADS / 32;

struct WTile { int32_t block, i0, j0, rows; };   // 32 columns starting at interior column j0, `rows` rows starting at interior row i0

struct KComp {   // one independent system: ranges into the component-sorted tables
    int32_t wt_begin, wt_end;
    int32_t s_begin, s_end, j_begin, j_end, l_begin, l_end;
    int32_t rt_begin, rt_end;
    int32_t nodes, _pad;
};
struct KCtl {    // per component; index 0 = x solve, 1 = y solve
    double tol[2], norm_b[2], norm_r[2];
    int32_t done[2];     // 1 converged, 2 breakdown, 3 iteration cap
    int32_t iters[2];
    int32_t applications, cycles;
};
struct KGroup { int32_t comp_begin, comp_end, cta_begin, n_ctas; };   // components [comp_begin, comp_end) of group_comps
struct alignas(128) KBarrier { unsigned int count, gen; unsigned int _pad[30]; };

struct KArgs {
    const WTile* wtiles;
    const DevBlock* blocks;
    const SmoothedRow* srows;
    const JunctionRow* jrows;
    const SlidingRow* lrows;
    const SlaveRow* slaves;
    const RhsTerm* rterms;
    const KComp* comps;
    const int32_t* group_comps;
    const KGroup* groups;
    const int32_t* cta_group;
    KCtl* ctl;
    KBarrier* bars;
    double* partials;            // 2 x n_ctas x 4
    const double2* xc;           // lagged coordinates (the mesh before this outer iteration)
    const double2* pq;           // control function (HAS_PQ)
    double2* xnew;               // iterate, warm-started from xc by the caller
    double2 *r, *rhat, *p, *v, *s, *t, *d;
    double rtol, atol;
    int32_t max_iters, max_restarts, n_ctas_total, _pad;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// all CTAs of a group; `gen` is the CTA's private copy of the barrier generation
__device__ __forceinline__ void group_barrier(KBarrier* b, unsigned int n, unsigned int& gen) {
    __syncthreads();
    gen += 1;
    if (n > 1 && threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&b->count, 1u) == n - 1) {
            atomicExch(&b->count, 0u);
            __threadfence();
            st_release_gpu(&b->gen, gen);
        } else {
            while ((int)(ld_acquire_gpu(&b->gen) - gen) < 0) {}
        }
        __threadfence();
    }
    __syncthreads();
}

// ---- row evaluation (lagged coefficients) ------------------------------------------------------------
// MODE_APPLY: (A u)_i / a_ii, homogeneous.  MODE_RESID: (b - A u)_i / a_ii with the periodic shifts folded in.
template <int MODE, bool HAS_PQ>
__device__ __forceinline__ double2 k_smoothed(const SmoothedRow& row, const double2* u, const double2* __restrict__ xc, const double2* __restrict__ pq, double& b2x,
                                              double& b2y) {
    const double2 per = make_double2(row.px, row.py);
    const double2 sh = MODE == MODE_RESID ? per : make_double2(0.0, 0.0);
    const double2 C = u[row.g0];
    const double2 W = u[row.g0 - row.d0], E = u[row.g0 + row.d0];
    const double2 S = u[row.g0 + row.n0], SW = u[row.g0 - row.d0 + row.n0], SE = u[row.g0 + row.d0 + row.n0];
    const double2 N = u[row.iN] - sh, NW = u[row.iNW] - sh, NE = u[row.iNE] - sh;
    const double2 cW = ldg2(xc + row.g0 - row.d0), cE = ldg2(xc + row.g0 + row.d0), cS = ldg2(xc + row.g0 + row.n0);
    const double2 cN = ldg2(xc + row.iN) - per;  // smooth.zig:1032
    const Metric m = metric_terms(cW, cE, cN - cS);
    double P = 0.0, Q = 0.0;
    if (HAS_PQ) {
        const double2 f = ldg2(pq + row.g0);
        if (row.periodic) { P = f.x; Q = f.y; } else { P = f.y; Q = f.x; }  // smooth.zig:1040-1041 vs 1082-1083
    }
    const double2 rel = row_rel<HAS_PQ>(m, P, Q, C, W, E, (N - C) + (S - C), N - S, NE - SE, NW - SW);
    if (MODE == MODE_RESID && row.periodic) {  // rhs of the reference's row: p (a(i-1,j+1) + a(i,j+1) + a(i+1,j+1)) = p g11 (1 + Q/2)
        const double a = 0.25 * m.g11 * (1.0 + 0.5 * Q);
        b2x = (row.px * a) * (row.px * a); b2y = (row.py * a) * (row.py * a);
    }
    return row_result<MODE>(m, rel, C, 1.0);
}
template <int MODE>
__device__ __forceinline__ double2 k_junction(const JunctionRow& row, const double2* u) {
    const double2 C = u[row.self];
    double2 sum = make_double2(0.0, 0.0);
    for (int k = 0; k < row.n; ++k) sum = sum + (u[row.nbr[k]] - C);
    const double n = (double)row.n;
    if (MODE == MODE_APPLY) return make_double2(-(sum.x / n), -(sum.y / n));
    return make_double2((sum.x - row.rhs_x) / n, (sum.y - row.rhs_y) / n);
}
template <int MODE>
__device__ __forceinline__ double2 k_sliding(const SlidingRow& row, const double2* u) {
    const double2 C = u[row.self], I = u[row.inner];
    if (MODE == MODE_APPLY) return make_double2(C.x, C.y - I.y);
    return make_double2(row.rhs_x - C.x, (double)row.ysign * row.rhs_y - (C.y - I.y));
}

// interior rows of one warp tile: epi(index into the fields, row result) for every active node
template <int MODE, bool HAS_PQ, class Epi>
__device__ __forceinline__ void k_interior(const WTile& t, const DevBlock& b, const double2* u, const double2* __restrict__ xc, const double2* __restrict__ pq, Epi&& epi) {
    const int lane = threadIdx.x & 31;
    const int nj = b.nj;
    const int j = t.j0 + lane;
    const bool active = j <= nj - 2;
    const int jc = active ? j : nj - 2;
    const int i_end = min(t.i0 + t.rows, b.ni - 1);
    const double2* ub = u + b.off;
    const double2* cb = xc + b.off;
    size_t idx = (size_t)(t.i0 - 1) * nj + jc;
    double2 Cm = ub[idx], Dm = ub[idx + 1] - ub[idx - 1];
    double2 cCm = ldg2(cb + idx);
    idx += nj;
    double2 l = ub[idx - 1], r = ub[idx + 1];
    double2 C0 = ub[idx], D0 = r - l, R0 = (r - C0) + (l - C0);
    double2 cC0 = ldg2(cb + idx), cD0 = ldg2(cb + idx + 1) - ldg2(cb + idx - 1);
    for (int i = t.i0; i < i_end; ++i) {
        const size_t ip = idx + nj;
        const double2 lp = ub[ip - 1], rp = ub[ip + 1];
        const double2 Cp = ub[ip], Dp = rp - lp, Rp = (rp - Cp) + (lp - Cp);
        const double2 cCp = ldg2(cb + ip), cDp = ldg2(cb + ip + 1) - ldg2(cb + ip - 1);
        const Metric m = metric_terms(cCm, cCp, cD0);
        double P = 0.0, Q = 0.0;
        if (HAS_PQ) {
            const double2 f = ldg2(pq + b.off + idx);
            P = f.x; Q = f.y;
        }
        const double2 rel = row_rel<HAS_PQ>(m, P, Q, C0, Cm, Cp, R0, D0, Dp, Dm);
        if (active) epi((size_t)b.off + idx, row_result<MODE>(m, rel, C0, 1.0));
        Cm = C0; Dm = D0;
        C0 = Cp; D0 = Dp; R0 = Rp;
        cCm = cC0; cC0 = cCp; cD0 = cDp;
        idx = ip;
    }
}
// the same nodes without a stencil (vector updates)
template <class Body>
__device__ __forceinline__ void k_interior_nodes(const WTile& t, const DevBlock& b, Body&& body) {
    const int j = t.j0 + (threadIdx.x & 31);
    if (j > b.nj - 2) return;
    const int i_end = min(t.i0 + t.rows, b.ni - 1);
    for (int i = t.i0; i < i_end; ++i) body((size_t)b.off + (size_t)i * b.nj + j);
}

struct KScal {   // solver scalars of one component, identical in every thread of the group
    double rho_old[2], rho_new[2], alpha[2], omega[2], beta[2], tol[2], norm_b[2], norm_r[2];
    int done[2], iters[2];
};

template <bool HAS_PQ>
__global__ void __launch_bounds__(K_THREADS) bicgstab_persistent_kernel(const KArgs a) {
    __shared__ double sh_part[K_WARPS][4];
    __shared__ double sh_red[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const KGroup G = a.groups[a.cta_group[blockIdx.x]];
    KBarrier* bar = a.bars + a.cta_group[blockIdx.x];
    const int crank = (int)blockIdx.x - G.cta_begin;                // CTA within the group
    const int gwarp = crank * K_WARPS + warp, gwarps = G.n_ctas * K_WARPS;
    const int gthread = crank * K_THREADS + tid, gthreads = G.n_ctas * K_THREADS;
    unsigned int gen = ld_acquire_gpu(&bar->gen);                    // nobody can have passed a barrier yet
    int parity = 0;
    const double eps = 1e-30;                                        // breakdown_eps, BiCGStab.zig:280

    // barrier + combination of the group's partial sums (fixed order: bit-identical in every CTA)
    auto reduce = [&](double (&acc)[4]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) sh_part[warp][k] = acc[k];
        }
        __syncthreads();
        double* mine = a.partials + ((size_t)parity * a.n_ctas_total + blockIdx.x) * 4;
        if (tid < 4) {
            double s = 0.0;
            for (int w = 0; w < K_WARPS; ++w) s += sh_part[w][tid];
            __stcg(mine + tid, s);
        }
        group_barrier(bar, (unsigned)G.n_ctas, gen);
        if (warp == 0) {
            const double* base = a.partials + ((size_t)parity * a.n_ctas_total + G.cta_begin) * 4;
            double s[4] = {0.0, 0.0, 0.0, 0.0};
            for (int c = lane; c < G.n_ctas; c += 32) {
#pragma unroll
                for (int k = 0; k < 4; ++k) s[k] += __ldcg(base + (size_t)c * 4 + k);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) s[k] = warp_sum(s[k]);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) sh_red[k] = s[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = sh_red[k];
        parity ^= 1;
    };

    for (int gc = G.comp_begin; gc < G.comp_end; ++gc) {
        const int comp = a.group_comps[gc];
        const KComp K = a.comps[comp];
        KScal S;
        int applications = 0, cycle = 0;
#pragma unroll
        for (int c = 0; c < 2; ++c) { S.tol[c] = 0.0; S.norm_b[c] = 0.0; S.norm_r[c] = 0.0; S.done[c] = 0; S.iters[c] = 0; }

        // row-owner loops: interior warp tiles, then boundary rows (one thread each); `mirror` hands a value to the copies
        auto mirror = [&](double2* f, int sb, int se, double2 val) {
            for (int k = sb; k < se; ++k) f[a.slaves[k].self] = val;
        };
        for (;; ++cycle) {
            // ---- R0: r = D^-1 (b - A x); rhat = r; p = v = d = 0; ||r||^2 (and ||b||^2 in the first cycle) ----
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int w = K.wt_begin + gwarp; w < K.wt_end; w += gwarps) {
                const WTile t = a.wtiles[w];
                const DevBlock b = a.blocks[t.block];
                k_interior<MODE_RESID, HAS_PQ>(t, b, a.xnew, a.xc, a.pq, [&](size_t k, double2 res) {
                    a.r[k] = res; a.rhat[k] = res;
                    a.p[k] = make_double2(0.0, 0.0); a.v[k] = make_double2(0.0, 0.0); a.d[k] = make_double2(0.0, 0.0);
                    acc[0] += res.x * res.x; acc[1] += res.y * res.y;
                });
            }
            auto init_row = [&](int64_t self, int sb, int se, double2 res) {
                const double2 z = make_double2(0.0, 0.0);
                a.r[self] = res; a.rhat[self] = res; a.p[self] = z; a.v[self] = z; a.d[self] = z;
                mirror(a.p, sb, se, z); mirror(a.s, sb, se, z);
                acc[0] += res.x * res.x; acc[1] += res.y * res.y;
            };
            for (int q = K.s_begin + gthread; q < K.s_end; q += gthreads) {
                const SmoothedRow row = a.srows[q];
                double b2x = 0.0, b2y = 0.0;
                const double2 res = k_smoothed<MODE_RESID, HAS_PQ>(row, a.xnew, a.xc, a.pq, b2x, b2y);
                init_row(row.g0, row.slave_begin, row.slave_end, res);
                if (cycle == 0) { acc[2] += b2x; acc[3] += b2y; }
            }
            for (int q = K.j_begin + gthread; q < K.j_end; q += gthreads) {
                const JunctionRow row = a.jrows[q];
                init_row(row.self, row.slave_begin, row.slave_end, k_junction<MODE_RESID>(row, a.xnew));
            }
            for (int q = K.l_begin + gthread; q < K.l_end; q += gthreads) {
                const SlidingRow row = a.lrows[q];
                init_row(row.self, row.slave_begin, row.slave_end, k_sliding<MODE_RESID>(row, a.xnew));
            }
            if (cycle == 0) {  // constant part of the reference's ||b||^2 (BiCGStab.zig:289-291)
                for (int q = K.rt_begin + gthread; q < K.rt_end; q += gthreads) {
                    const RhsTerm t = a.rterms[q];
                    double bx = t.cx, by = t.cy;
                    if (t.from_x | t.from_y) { const double2 x0 = ldg2(a.xc + t.g); if (t.from_x) bx = x0.x; if (t.from_y) by = x0.y; }
                    acc[2] += bx * bx; acc[3] += by * by;
                }
            }
            reduce(acc);
            applications += 1;
            bool stop = cycle > a.max_restarts;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                S.norm_r[c] = sqrt(acc[c]);
                if (cycle == 0) { S.norm_b[c] = sqrt(acc[2 + c]); S.tol[c] = fmax(a.atol, a.rtol * S.norm_b[c]); }   // GMRES.zig:305-306 / BiCGStab.zig:291
                S.rho_old[c] = 1.0; S.alpha[c] = 1.0; S.omega[c] = 1.0;
                S.rho_new[c] = acc[c];                                  // rhat = r
                S.done[c] = S.norm_r[c] <= S.tol[c] ? 1 : (S.iters[c] >= a.max_iters ? 3 : 0);
                if (!S.done[c] && fabs(S.rho_new[c]) < eps) S.done[c] = 2;
                S.beta[c] = S.rho_new[c];
            }
            if ((S.done[0] == 1 && S.done[1] == 1) || S.done[0] == 3 || S.done[1] == 3 || stop) break;

            // ---- BiCGStab iterations (BiCGStab.zig:303-366), x and y in lock-step ----
            while (!(S.done[0] && S.done[1])) {
                const bool dx = S.done[0] != 0, dy = S.done[1] != 0;
                // P1: p = r + beta (p - omega v)
                {
                    const double bx = S.beta[0], by = S.beta[1], ox = S.omega[0], oy = S.omega[1];
                    auto upd = [&](size_t k) {
                        const double2 rr = a.r[k], vv = a.v[k];
                        double2 pp = a.p[k];
                        pp.x = dx ? 0.0 : rr.x + bx * (pp.x - ox * vv.x);
                        pp.y = dy ? 0.0 : rr.y + by * (pp.y - oy * vv.y);
                        a.p[k] = pp;
                        return pp;
                    };
                    for (int w = K.wt_begin + gwarp; w < K.wt_end; w += gwarps) {
                        const WTile t = a.wtiles[w];
                        k_interior_nodes(t, a.blocks[t.block], [&](size_t k) { upd(k); });
                    }
                    for (int q = K.s_begin + gthread; q < K.s_end; q += gthreads) { const SmoothedRow& row = a.srows[q]; mirror(a.p, row.slave_begin, row.slave_end, upd((size_t)row.g0)); }
                    for (int q = K.j_begin + gthread; q < K.j_end; q += gthreads) { const JunctionRow& row = a.jrows[q]; mirror(a.p, row.slave_begin, row.slave_end, upd((size_t)row.self)); }
                    for (int q = K.l_begin + gthread; q < K.l_end; q += gthreads) { const SlidingRow& row = a.lrows[q]; mirror(a.p, row.slave_begin, row.slave_end, upd((size_t)row.self)); }
                    group_barrier(bar, (unsigned)G.n_ctas, gen);
                }
                // P2: v = A p, rhat . v
                double acc2[4] = {0.0, 0.0, 0.0, 0.0};
                {
                    auto put = [&](size_t k, double2 res) {
                        a.v[k] = res;
                        const double2 h = a.rhat[k];
                        acc2[0] += h.x * res.x; acc2[1] += h.y * res.y;
                    };
                    for (int w = K.wt_begin + gwarp; w < K.wt_end; w += gwarps) {
                        const WTile t = a.wtiles[w];
                        k_interior<MODE_APPLY, HAS_PQ>(t, a.blocks[t.block], a.p, a.xc, a.pq, put);
                    }
                    double u0, u1;
                    for (int q = K.s_begin + gthread; q < K.s_end; q += gthreads) { const SmoothedRow row = a.srows[q]; put((size_t)row.g0, k_smoothed<MODE_APPLY, HAS_PQ>(row, a.p, a.xc, a.pq, u0, u1)); }
                    for (int q = K.j_begin + gthread; q < K.j_end; q += gthreads) { const JunctionRow row = a.jrows[q]; put((size_t)row.self, k_junction<MODE_APPLY>(row, a.p)); }
                    for (int q = K.l_begin + gthread; q < K.l_end; q += gthreads) { const SlidingRow row = a.lrows[q]; put((size_t)row.self, k_sliding<MODE_APPLY>(row, a.p)); }
                    reduce(acc2);
                    applications += 1;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (S.done[c]) continue;
                        if (fabs(acc2[c]) < eps) { S.done[c] = 2; S.alpha[c] = 0.0; }
                        else S.alpha[c] = S.rho_new[c] / acc2[c];
                    }
                }
                // P3: s = r - alpha v; d += alpha p; ||s||^2   (a component that just broke down is masked from here on)
                double acc3[4] = {0.0, 0.0, 0.0, 0.0};
                {
                    const bool mx = S.done[0] != 0, my = S.done[1] != 0;
                    const double ax = S.alpha[0], ay = S.alpha[1];
                    auto upd = [&](size_t k) {
                        const double2 rr = a.r[k], vv = a.v[k], pp = a.p[k];
                        double2 ss = make_double2(0.0, 0.0), dd = a.d[k];
                        if (!mx) { ss.x = rr.x - ax * vv.x; dd.x += ax * pp.x; acc3[0] += ss.x * ss.x; }
                        if (!my) { ss.y = rr.y - ay * vv.y; dd.y += ay * pp.y; acc3[1] += ss.y * ss.y; }
                        a.s[k] = ss; a.d[k] = dd;
                        return ss;
                    };
                    for (int w = K.wt_begin + gwarp; w < K.wt_end; w += gwarps) {
                        const WTile t = a.wtiles[w];
                        k_interior_nodes(t, a.blocks[t.block], [&](size_t k) { upd(k); });
                    }
                    for (int q = K.s_begin + gthread; q < K.s_end; q += gthreads) { const SmoothedRow& row = a.srows[q]; mirror(a.s, row.slave_begin, row.slave_end, upd((size_t)row.g0)); }
                    for (int q = K.j_begin + gthread; q < K.j_end; q += gthreads) { const JunctionRow& row = a.jrows[q]; mirror(a.s, row.slave_begin, row.slave_end, upd((size_t)row.self)); }
                    for (int q = K.l_begin + gthread; q < K.l_end; q += gthreads) { const SlidingRow& row = a.lrows[q]; mirror(a.s, row.slave_begin, row.slave_end, upd((size_t)row.self)); }
                    reduce(acc3);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (S.done[c]) continue;
                        S.iters[c] += 1;
                        S.norm_r[c] = sqrt(acc3[c]);
                        if (S.norm_r[c] <= S.tol[c]) S.done[c] = 1;
                    }
                }
                if (S.done[0] && S.done[1]) break;
                // P4: t = A s, t . s, t . t
                double acc4[4] = {0.0, 0.0, 0.0, 0.0};
                {
                    auto put = [&](size_t k, double2 res) {
                        a.t[k] = res;
                        const double2 ss = a.s[k];
                        acc4[0] += ss.x * res.x; acc4[1] += ss.y * res.y;
                        acc4[2] += res.x * res.x; acc4[3] += res.y * res.y;
                    };
                    for (int w = K.wt_begin + gwarp; w < K.wt_end; w += gwarps) {
                        const WTile t = a.wtiles[w];
                        k_interior<MODE_APPLY, HAS_PQ>(t, a.blocks[t.block], a.s, a.xc, a.pq, put);
                    }
                    double u0, u1;
                    for (int q = K.s_begin + gthread; q < K.s_end; q += gthreads) { const SmoothedRow row = a.srows[q]; put((size_t)row.g0, k_smoothed<MODE_APPLY, HAS_PQ>(row, a.s, a.xc, a.pq, u0, u1)); }
                    for (int q = K.j_begin + gthread; q < K.j_end; q += gthreads) { const JunctionRow row = a.jrows[q]; put((size_t)row.self, k_junction<MODE_APPLY>(row, a.s)); }
                    for (int q = K.l_begin + gthread; q < K.l_end; q += gthreads) { const SlidingRow row = a.lrows[q]; put((size_t)row.self, k_sliding<MODE_APPLY>(row, a.s)); }
                    reduce(acc4);
                    applications += 1;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (S.done[c]) continue;
                        if (fabs(acc4[2 + c]) < eps) { S.done[c] = 2; S.omega[c] = 0.0; }
                        else {
                            S.omega[c] = acc4[c] / acc4[2 + c];
                            if (fabs(S.omega[c]) < eps) { S.done[c] = 2; S.omega[c] = 0.0; }
                        }
                    }
                }
                // P5: d += omega s; r = s - omega t; ||r||^2, rhat . r
                double acc5[4] = {0.0, 0.0, 0.0, 0.0};
                {
                    const bool ex = S.done[0] != 0, ey = S.done[1] != 0;
                    const double ox = S.omega[0], oy = S.omega[1];
                    auto upd = [&](size_t k) {
                        const double2 ss = a.s[k], tt = a.t[k], rh = a.rhat[k];
                        double2 rr = a.r[k], dd = a.d[k];
                        if (!ex) { dd.x += ox * ss.x; rr.x = ss.x - ox * tt.x; acc5[0] += rr.x * rr.x; acc5[2] += rh.x * rr.x; }
                        if (!ey) { dd.y += oy * ss.y; rr.y = ss.y - oy * tt.y; acc5[1] += rr.y * rr.y; acc5[3] += rh.y * rr.y; }
                        a.r[k] = rr; a.d[k] = dd;
                    };
                    for (int w = K.wt_begin + gwarp; w < K.wt_end; w += gwarps) {
                        const WTile t = a.wtiles[w];
                        k_interior_nodes(t, a.blocks[t.block], upd);
                    }
                    for (int q = K.s_begin + gthread; q < K.s_end; q += gthreads) upd((size_t)a.srows[q].g0);
                    for (int q = K.j_begin + gthread; q < K.j_end; q += gthreads) upd((size_t)a.jrows[q].self);
                    for (int q = K.l_begin + gthread; q < K.l_end; q += gthreads) upd((size_t)a.lrows[q].self);
                    reduce(acc5);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (S.done[c]) continue;
                        S.norm_r[c] = sqrt(acc5[c]);
                        if (S.norm_r[c] <= S.tol[c]) { S.done[c] = 1; continue; }
                        S.rho_old[c] = S.rho_new[c];
                        S.rho_new[c] = acc5[2 + c];
                        if (S.iters[c] >= a.max_iters) { S.done[c] = 3; continue; }
                        if (fabs(S.rho_new[c]) < eps) { S.done[c] = 2; continue; }
                        S.beta[c] = (S.rho_new[c] / S.rho_old[c]) * (S.alpha[c] / S.omega[c]);
                    }
                }
            }
            // ---- x += d on the rows of this component; copies follow their root (x_copy = x_root + shift) ----
            {
                auto add = [&](size_t k) {
                    const double2 dd = a.d[k];
                    double2 xx = a.xnew[k];
                    xx.x += dd.x; xx.y += dd.y;
                    a.xnew[k] = xx;
                    return xx;
                };
                auto copies = [&](int sb, int se, double2 xx) {
                    for (int k = sb; k < se; ++k) { const SlaveRow sl = a.slaves[k]; a.xnew[sl.self] = make_double2(xx.x + sl.sx, xx.y + sl.sy); }
                };
                for (int w = K.wt_begin + gwarp; w < K.wt_end; w += gwarps) {
                    const WTile t = a.wtiles[w];
                    k_interior_nodes(t, a.blocks[t.block], [&](size_t k) { add(k); });
                }
                for (int q = K.s_begin + gthread; q < K.s_end; q += gthreads) { const SmoothedRow& row = a.srows[q]; copies(row.slave_begin, row.slave_end, add((size_t)row.g0)); }
                for (int q = K.j_begin + gthread; q < K.j_end; q += gthreads) { const JunctionRow& row = a.jrows[q]; copies(row.slave_begin, row.slave_end, add((size_t)row.self)); }
                for (int q = K.l_begin + gthread; q < K.l_end; q += gthreads) { const SlidingRow& row = a.lrows[q]; copies(row.slave_begin, row.slave_end, add((size_t)row.self)); }
                group_barrier(bar, (unsigned)G.n_ctas, gen);
            }
        }
        if (crank == 0 && tid == 0) {
            KCtl out;
#pragma unroll
            for (int c = 0; c < 2; ++c) { out.tol[c] = S.tol[c]; out.norm_b[c] = S.norm_b[c]; out.norm_r[c] = S.norm_r[c]; out.done[c] = S.done[c]; out.iters[c] = S.iters[c]; }
            out.applications = applications; out.cycles = cycle + 1;
            a.ctl[comp] = out;
        }
    }
}

}  // namespace tmesh
