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
//   * An iteration is THREE phases inside the kernel, separated by a group barrier (sense-reversing counter in global memory,
//     release / acquire at gpu scope; co-residency comes from the cooperative launch):
//         A  v = A p  with  p = r + beta (p - omega v)  formed on the fly at every stencil node (p and v ping-pong)
//         B  t = A s  with  s = r - alpha v             formed on the fly likewise; d += alpha p; ||s||, t.s, t.t
//         C  d += omega s;  r = s - omega t;  ||r||, rhat.r
//     The partial dot products of a phase are combined redundantly by every CTA in a fixed order, so all CTAs of a group hold
//     bit-identical scalars without a second barrier.
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
constexpr int K_NACC = 6;        // sums a phase reduces
constexpr int K_TILE_ROWS = 2;   // rows of a warp tile

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
struct alignas(128) KBarrier { unsigned int count; unsigned int _pad0[31]; unsigned int gen; unsigned int _pad1[31]; };   // arrivals and the polled generation on separate lines

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
    double* partials;            // 2 x n_ctas x K_NACC
    const double2* xc;           // lagged coordinates (the mesh before this outer iteration)
    const double2* pq;           // control function (HAS_PQ)
    double2* xnew;               // iterate, warm-started from xc by the caller
    double2 *r, *rhat, *p[2], *v[2], *s, *t, *d;   // p and v ping-pong: a phase reads the old field at its neighbours while it writes the new one
    double rtol, atol;
    int32_t max_iters, max_restarts, n_ctas_total, polish;   // polish: refinement cycles after convergence (each asks for 10x less residual)
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
// `u(k)` yields the value the row is applied to at node k (an array, or a vector update formed on the fly).
// MODE_APPLY: (A u)_i / a_ii, homogeneous.  MODE_RESID: (b - A u)_i / a_ii with the periodic shifts folded in.
struct KRow { double2 res, centre; };
template <int MODE, bool HAS_PQ, class U>
__device__ __forceinline__ KRow k_smoothed(const SmoothedRow& row, U&& u, const double2* __restrict__ xc, const double2* __restrict__ pq, double& b2x, double& b2y) {
    const double2 per = make_double2(row.px, row.py);
    const double2 sh = MODE == MODE_RESID ? per : make_double2(0.0, 0.0);
    const double2 C = u(row.g0);
    const double2 W = u(row.g0 - row.d0), E = u(row.g0 + row.d0);
    const double2 S = u(row.g0 + row.n0), SW = u(row.g0 - row.d0 + row.n0), SE = u(row.g0 + row.d0 + row.n0);
    const double2 N = u(row.iN) - sh, NW = u(row.iNW) - sh, NE = u(row.iNE) - sh;
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
    return KRow{row_result<MODE>(m, rel, C, 1.0), C};
}
template <int MODE, class U>
__device__ __forceinline__ KRow k_junction(const JunctionRow& row, U&& u) {
    const double2 C = u(row.self);
    double2 sum = make_double2(0.0, 0.0);
    for (int k = 0; k < row.n; ++k) sum = sum + (u(row.nbr[k]) - C);
    const double n = (double)row.n;
    if (MODE == MODE_APPLY) return KRow{make_double2(-(sum.x / n), -(sum.y / n)), C};
    return KRow{make_double2((sum.x - row.rhs_x) / n, (sum.y - row.rhs_y) / n), C};
}
template <int MODE, class U>
__device__ __forceinline__ KRow k_sliding(const SlidingRow& row, U&& u) {
    const double2 C = u(row.self), I = u(row.inner);
    if (MODE == MODE_APPLY) return KRow{make_double2(C.x, C.y - I.y), C};
    return KRow{make_double2(row.rhs_x - C.x, (double)row.ysign * row.rhs_y - (C.y - I.y)), C};
}

// Interior rows of one warp tile (32 columns x at most K_TILE_ROWS rows): epi(node, row result, centre value) for every
// active node.  The solves live in L2, so what counts is memory-level parallelism, not re-use: every row of the tile
// gathers its nine values independently (the neighbours are L1 hits of the same lines) and nothing is stored before all
// rows of the tile have been loaded and evaluated.
template <int MODE, bool HAS_PQ, class U, class Epi>
__device__ __forceinline__ void k_interior(const WTile& t, const DevBlock& b, U&& u, const double2* __restrict__ xc, const double2* __restrict__ pq, Epi&& epi) {
    const int lane = threadIdx.x & 31;
    const int nj = b.nj;
    const int j = t.j0 + lane;
    const bool active = j <= nj - 2;
    const int jc = active ? j : nj - 2;
    const double2* cb = xc + b.off;
    KRow out[K_TILE_ROWS];
#pragma unroll
    for (int q = 0; q < K_TILE_ROWS; ++q) {
        if (q < t.rows) {
            const int64_t l = (int64_t)(t.i0 + q) * nj + jc, g = b.off + l;
            const double2 C = u(g), W = u(g - nj), E = u(g + nj), S = u(g - 1), N = u(g + 1);
            const double2 SW = u(g - nj - 1), NW = u(g - nj + 1), SE = u(g + nj - 1), NE = u(g + nj + 1);
            const double2 cW = ldg2(cb + l - nj), cE = ldg2(cb + l + nj), cS = ldg2(cb + l - 1), cN = ldg2(cb + l + 1);
            const Metric m = metric_terms(cW, cE, cN - cS);
            double P = 0.0, Q = 0.0;
            if (HAS_PQ) {
                const double2 f = ldg2(pq + g);
                P = f.x; Q = f.y;
            }
            const double2 rel = row_rel<HAS_PQ>(m, P, Q, C, W, E, (N - C) + (S - C), N - S, NE - SE, NW - SW);
            out[q] = KRow{row_result<MODE>(m, rel, C, 1.0), C};
        }
    }
    if (!active) return;
#pragma unroll
    for (int q = 0; q < K_TILE_ROWS; ++q)
        if (q < t.rows) epi(b.off + (int64_t)(t.i0 + q) * nj + j, out[q].res, out[q].centre);
}
// Marching form for tall tiles (the HBM-bound phased path): a thread owns one column and walks down the rows with a 3-row
// register window on the values and on the lagged coordinates, so a row is fetched once per tile (its j-1 / j+1 neighbours
// are L1 hits of the same lines) instead of three times.  (Tried and dropped: neighbours by warp shuffle -- fewer L1
// wavefronts, but the kernel is latency-bound at 4 CTAs/SM and forcing 6 CTAs/SM only added spills: 740 -> 820 ms on 128 cuts.)
template <int MODE, bool HAS_PQ, class U, class Epi>
__device__ __forceinline__ void k_interior_march(const WTile& t, const DevBlock& b, U&& u, const double2* __restrict__ xc, const double2* __restrict__ pq, Epi&& epi) {
    const int lane = threadIdx.x & 31;
    const int nj = b.nj;
    const int j = t.j0 + lane;
    const bool active = j <= nj - 2;
    const int jc = active ? j : nj - 2;
    const int i_end = min(t.i0 + t.rows, b.ni - 1);
    const double2* cb = xc + b.off;
    int64_t l = (int64_t)(t.i0 - 1) * nj + jc;
    double2 Cm = u(b.off + l), Dm = u(b.off + l + 1) - u(b.off + l - 1);
    double2 cCm = ldg2(cb + l);
    l += nj;
    double2 lf = u(b.off + l - 1), rt = u(b.off + l + 1);
    double2 C0 = u(b.off + l), D0 = rt - lf, R0 = (rt - C0) + (lf - C0);
    double2 cC0 = ldg2(cb + l), cD0 = ldg2(cb + l + 1) - ldg2(cb + l - 1);
#pragma unroll 1
    for (int i = t.i0; i < i_end; ++i) {
        const int64_t lp = l + nj;
        const double2 lfp = u(b.off + lp - 1), rtp = u(b.off + lp + 1);
        const double2 Cp = u(b.off + lp), Dp = rtp - lfp, Rp = (rtp - Cp) + (lfp - Cp);
        const double2 cCp = ldg2(cb + lp), cDp = ldg2(cb + lp + 1) - ldg2(cb + lp - 1);
        const Metric m = metric_terms(cCm, cCp, cD0);
        double P = 0.0, Q = 0.0;
        if (HAS_PQ) {
            const double2 f = ldg2(pq + b.off + l);
            P = f.x; Q = f.y;
        }
        const double2 rel = row_rel<HAS_PQ>(m, P, Q, C0, Cm, Cp, R0, D0, Dp, Dm);
        if (active) epi(b.off + l, row_result<MODE>(m, rel, C0, 1.0), C0);
        Cm = C0; Dm = D0;
        C0 = Cp; D0 = Dp; R0 = Rp;
        cCm = cC0; cC0 = cCp; cD0 = cDp;
        l = lp;
    }
}
// vector updates over a tall tile
template <class Load, class Store>
__device__ __forceinline__ void k_interior_nodes_march(const WTile& t, const DevBlock& b, Load&& load, Store&& store) {
    const int j = t.j0 + (threadIdx.x & 31);
    if (j > b.nj - 2) return;
    const int i_end = min(t.i0 + t.rows, b.ni - 1);
#pragma unroll 4
    for (int i = t.i0; i < i_end; ++i) {
        const int64_t k = b.off + (int64_t)i * b.nj + j;
        store(k, load(k));
    }
}

// the same nodes without a stencil (vector updates): load(k) for all rows first, then store(k, loaded)
template <class Load, class Store>
__device__ __forceinline__ void k_interior_nodes(const WTile& t, const DevBlock& b, Load&& load, Store&& store) {
    const int j = t.j0 + (threadIdx.x & 31);
    if (j > b.nj - 2) return;
    decltype(load((int64_t)0)) tmp[K_TILE_ROWS];
#pragma unroll
    for (int q = 0; q < K_TILE_ROWS; ++q)
        if (q < t.rows) tmp[q] = load(b.off + (int64_t)(t.i0 + q) * b.nj + j);
#pragma unroll
    for (int q = 0; q < K_TILE_ROWS; ++q)
        if (q < t.rows) store(b.off + (int64_t)(t.i0 + q) * b.nj + j, tmp[q]);
}

struct KScal {   // solver scalars of one component, identical in every thread of the group
    double rho_old[2], rho_new[2], alpha[2], omega[2], beta[2], tol[2], tol_eff[2], norm_b[2], norm_r[2];
    int done[2], iters[2];
};

template <bool HAS_PQ>
__global__ void __launch_bounds__(K_THREADS) bicgstab_persistent_kernel(const KArgs a) {
    __shared__ double sh_part[K_WARPS][K_NACC];
    __shared__ double sh_red[K_NACC];
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
    auto reduce = [&](double (&acc)[K_NACC]) {
#pragma unroll
        for (int k = 0; k < K_NACC; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K_NACC; ++k) sh_part[warp][k] = acc[k];
        }
        __syncthreads();
        if (tid < K_NACC) {
            double s = 0.0;
            for (int w = 0; w < K_WARPS; ++w) s += sh_part[w][tid];
            __stcg(a.partials + ((size_t)parity * a.n_ctas_total + blockIdx.x) * K_NACC + tid, s);
        }
        group_barrier(bar, (unsigned)G.n_ctas, gen);
        if (warp == 0) {
            const double* base = a.partials + ((size_t)parity * a.n_ctas_total + G.cta_begin) * K_NACC;
            double s[K_NACC];
#pragma unroll
            for (int k = 0; k < K_NACC; ++k) s[k] = 0.0;
            for (int c = lane; c < G.n_ctas; c += 32) {
#pragma unroll
                for (int k = 0; k < K_NACC; ++k) s[k] += __ldcg(base + (size_t)c * K_NACC + k);
            }
#pragma unroll
            for (int k = 0; k < K_NACC; ++k) s[k] = warp_sum(s[k]);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < K_NACC; ++k) sh_red[k] = s[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K_NACC; ++k) acc[k] = sh_red[k];
        parity ^= 1;
    };

    for (int gc = G.comp_begin; gc < G.comp_end; ++gc) {
        const int comp = a.group_comps[gc];
        const KComp K = a.comps[comp];
        const int n_s = K.s_end - K.s_begin, n_j = K.j_end - K.j_begin, n_l = K.l_end - K.l_begin, n_bnd = n_s + n_j + n_l;
        KScal S;
        int applications = 0, cycle = 0, pb = 0, vb = 0, polish_left = a.polish;
#pragma unroll
        for (int c = 0; c < 2; ++c) { S.tol[c] = 0.0; S.norm_b[c] = 0.0; S.norm_r[c] = 0.0; S.done[c] = 0; S.iters[c] = 0; }

        // Row-owner loops.  Interior nodes: warp tiles.  Boundary rows: one thread per row, taken from the END of the group's
        // threads (the warps without a tile).  `mirror` hands a value to the `connected` copies of a row's node.
        auto mirror = [&](double2* f, int sb, int se, double2 val) {
            for (int k = sb; k < se; ++k) f[a.slaves[k].self] = val;
        };
        auto for_tiles = [&](auto&& fn) {
            for (int w = K.wt_begin + gwarp; w < K.wt_end; w += gwarps) {
                const WTile t = a.wtiles[w];
                fn(t, a.blocks[t.block]);
            }
        };
        auto for_bnd = [&](auto&& fs, auto&& fj, auto&& fl) {
            for (int q = gthreads - 1 - gthread; q < n_bnd; q += gthreads) {
                if (q < n_s) fs(a.srows[K.s_begin + q]);
                else if (q < n_s + n_j) fj(a.jrows[K.j_begin + q - n_s]);
                else fl(a.lrows[K.l_begin + q - n_s - n_j]);
            }
        };
        for (;; ++cycle) {
            // ---- R0: r = D^-1 (b - A x); rhat = r; p = v = d = 0; ||r||^2 (and ||b||^2 in the first cycle) ----
            double acc[K_NACC] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            {
                double2* const P0 = a.p[pb];
                double2* const V0 = a.v[vb];
                const double2 z = make_double2(0.0, 0.0);
                auto xval = [&](int64_t k) { return a.xnew[k]; };
                auto init = [&](int64_t k, double2 res, int sb, int se) {
                    a.r[k] = res; a.rhat[k] = res; P0[k] = z; V0[k] = z; a.d[k] = z;
                    mirror(a.r, sb, se, res); mirror(P0, sb, se, z); mirror(V0, sb, se, z);
                    acc[0] += res.x * res.x; acc[1] += res.y * res.y;
                };
                for_tiles([&](const WTile& t, const DevBlock& b) {
                    k_interior<MODE_RESID, HAS_PQ>(t, b, xval, a.xc, a.pq, [&](int64_t k, double2 res, double2) { init(k, res, 0, 0); });
                });
                for_bnd([&](const SmoothedRow& row) {
                            double b2x = 0.0, b2y = 0.0;
                            const KRow o = k_smoothed<MODE_RESID, HAS_PQ>(row, xval, a.xc, a.pq, b2x, b2y);
                            init(row.g0, o.res, row.slave_begin, row.slave_end);
                            if (cycle == 0) { acc[2] += b2x; acc[3] += b2y; }
                        },
                        [&](const JunctionRow& row) { init(row.self, k_junction<MODE_RESID>(row, xval).res, row.slave_begin, row.slave_end); },
                        [&](const SlidingRow& row) { init(row.self, k_sliding<MODE_RESID>(row, xval).res, row.slave_begin, row.slave_end); });
                if (cycle == 0) {  // constant part of the reference's ||b||^2 (BiCGStab.zig:289-291)
                    for (int q = K.rt_begin + gthread; q < K.rt_end; q += gthreads) {
                        const RhsTerm t = a.rterms[q];
                        double bx = t.cx, by = t.cy;
                        if (t.from_x | t.from_y) { const double2 x0 = ldg2(a.xc + t.g); if (t.from_x) bx = x0.x; if (t.from_y) by = x0.y; }
                        acc[2] += bx * bx; acc[3] += by * by;
                    }
                }
            }
            reduce(acc);
            applications += 1;
            const bool stop = cycle > a.max_restarts;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                S.norm_r[c] = sqrt(acc[c]);
                if (cycle == 0) { S.norm_b[c] = sqrt(acc[2 + c]); S.tol[c] = fmax(a.atol, a.rtol * S.norm_b[c]); }   // GMRES.zig:305-306 / BiCGStab.zig:291
                S.tol_eff[c] = S.tol[c];
                S.rho_old[c] = 1.0; S.alpha[c] = 1.0; S.omega[c] = 1.0;
                S.rho_new[c] = acc[c];                                  // rhat = r
                S.done[c] = S.norm_r[c] <= S.tol[c] ? 1 : (S.iters[c] >= a.max_iters ? 3 : 0);
                if (!S.done[c] && fabs(S.rho_new[c]) < eps) S.done[c] = 2;
                S.beta[c] = S.rho_new[c];
            }
            if (S.done[0] == 1 && S.done[1] == 1 && polish_left > 0 && cycle > 0 && !stop) {
                // converged; a refinement cycle asks for a 10x smaller (true) residual -- the 2-norm of the row-scaled residual
                // says little about the smooth part of the error (tight, "exact Picard step" settings only)
                polish_left -= 1;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    S.tol_eff[c] = 0.1 * S.norm_r[c];
                    S.done[c] = fabs(S.rho_new[c]) < eps ? 2 : 0;
                }
            }
            if ((S.done[0] == 1 && S.done[1] == 1) || S.done[0] == 3 || S.done[1] == 3 || stop) break;

            // ---- BiCGStab iterations (BiCGStab.zig:303-366), x and y in lock-step ----
            while (!(S.done[0] && S.done[1])) {
                // A: p = r + beta (p - omega v) formed at every stencil node; v = A p; rhat . v
                double accA[K_NACC] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                {
                    const bool dx = S.done[0] != 0, dy = S.done[1] != 0;
                    const double bx = S.beta[0], by = S.beta[1], ox = S.omega[0], oy = S.omega[1];
                    const double2* const Pold = a.p[pb];
                    const double2* const Vold = a.v[vb];
                    double2* const Pnew = a.p[pb ^ 1];
                    double2* const Vnew = a.v[vb ^ 1];
                    auto pval = [&](int64_t k) {
                        const double2 rr = a.r[k], vv = Vold[k], pp = Pold[k];
                        return make_double2(dx ? 0.0 : rr.x + bx * (pp.x - ox * vv.x), dy ? 0.0 : rr.y + by * (pp.y - oy * vv.y));
                    };
                    auto put = [&](int64_t k, const KRow& o, int sb, int se) {
                        Pnew[k] = o.centre; Vnew[k] = o.res;
                        mirror(Pnew, sb, se, o.centre); mirror(Vnew, sb, se, o.res);
                        const double2 h = a.rhat[k];
                        accA[0] += h.x * o.res.x; accA[1] += h.y * o.res.y;
                    };
                    for_tiles([&](const WTile& t, const DevBlock& b) {
                        k_interior<MODE_APPLY, HAS_PQ>(t, b, pval, a.xc, a.pq, [&](int64_t k, double2 res, double2 c) { put(k, KRow{res, c}, 0, 0); });
                    });
                    double u0, u1;
                    for_bnd([&](const SmoothedRow& row) { put(row.g0, k_smoothed<MODE_APPLY, HAS_PQ>(row, pval, a.xc, a.pq, u0, u1), row.slave_begin, row.slave_end); },
                            [&](const JunctionRow& row) { put(row.self, k_junction<MODE_APPLY>(row, pval), row.slave_begin, row.slave_end); },
                            [&](const SlidingRow& row) { put(row.self, k_sliding<MODE_APPLY>(row, pval), row.slave_begin, row.slave_end); });
                }
                reduce(accA);
                pb ^= 1; vb ^= 1;
                applications += 1;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (S.done[c]) continue;
                    if (fabs(accA[c]) < eps) { S.done[c] = 2; S.alpha[c] = 0.0; }
                    else S.alpha[c] = S.rho_new[c] / accA[c];
                }
                // B: s = r - alpha v formed at every stencil node; d += alpha p; t = A s; ||s||^2, t . s, t . t
                double accB[K_NACC] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                {
                    const bool mx = S.done[0] != 0, my = S.done[1] != 0;   // a component that just broke down is masked from here on
                    const double ax = S.alpha[0], ay = S.alpha[1];
                    const double2* const Pcur = a.p[pb];
                    const double2* const Vcur = a.v[vb];
                    auto sval = [&](int64_t k) {
                        const double2 rr = a.r[k], vv = Vcur[k];
                        return make_double2(mx ? 0.0 : rr.x - ax * vv.x, my ? 0.0 : rr.y - ay * vv.y);
                    };
                    auto put = [&](int64_t k, const KRow& o) {
                        const double2 pp = Pcur[k];
                        double2 dd = a.d[k];
                        if (!mx) dd.x += ax * pp.x;
                        if (!my) dd.y += ay * pp.y;
                        a.s[k] = o.centre; a.t[k] = o.res; a.d[k] = dd;
                        accB[0] += o.centre.x * o.centre.x; accB[1] += o.centre.y * o.centre.y;
                        accB[2] += o.centre.x * o.res.x; accB[3] += o.centre.y * o.res.y;
                        accB[4] += o.res.x * o.res.x; accB[5] += o.res.y * o.res.y;
                    };
                    for_tiles([&](const WTile& t, const DevBlock& b) {
                        k_interior<MODE_APPLY, HAS_PQ>(t, b, sval, a.xc, a.pq, [&](int64_t k, double2 res, double2 c) { put(k, KRow{res, c}); });
                    });
                    double u0, u1;
                    for_bnd([&](const SmoothedRow& row) { put(row.g0, k_smoothed<MODE_APPLY, HAS_PQ>(row, sval, a.xc, a.pq, u0, u1)); },
                            [&](const JunctionRow& row) { put(row.self, k_junction<MODE_APPLY>(row, sval)); },
                            [&](const SlidingRow& row) { put(row.self, k_sliding<MODE_APPLY>(row, sval)); });
                }
                reduce(accB);
                applications += 1;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (S.done[c]) continue;
                    S.iters[c] += 1;
                    S.norm_r[c] = sqrt(accB[c]);
                    if (S.norm_r[c] <= S.tol_eff[c]) { S.done[c] = 1; continue; }   // x += alpha p has been taken; no omega part
                    if (fabs(accB[4 + c]) < eps) { S.done[c] = 2; S.omega[c] = 0.0; }
                    else {
                        S.omega[c] = accB[2 + c] / accB[4 + c];
                        if (fabs(S.omega[c]) < eps) { S.done[c] = 2; S.omega[c] = 0.0; }
                    }
                }
                if (S.done[0] && S.done[1]) break;
                // C: d += omega s; r = s - omega t; ||r||^2, rhat . r
                double accC[K_NACC] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                {
                    const bool ex = S.done[0] != 0, ey = S.done[1] != 0;
                    const double ox = S.omega[0], oy = S.omega[1];
                    struct RD { double2 r, d, h; };
                    auto load = [&](int64_t k) {
                        const double2 ss = a.s[k], tt = a.t[k];
                        RD o{a.r[k], a.d[k], a.rhat[k]};
                        if (!ex) { o.d.x += ox * ss.x; o.r.x = ss.x - ox * tt.x; }
                        if (!ey) { o.d.y += oy * ss.y; o.r.y = ss.y - oy * tt.y; }
                        return o;
                    };
                    auto store = [&](int64_t k, const RD& o) {
                        a.r[k] = o.r; a.d[k] = o.d;
                        if (!ex) { accC[0] += o.r.x * o.r.x; accC[2] += o.h.x * o.r.x; }
                        if (!ey) { accC[1] += o.r.y * o.r.y; accC[3] += o.h.y * o.r.y; }
                    };
                    auto upd = [&](int64_t k, int sb, int se) { const RD o = load(k); store(k, o); mirror(a.r, sb, se, o.r); };
                    for_tiles([&](const WTile& t, const DevBlock& b) { k_interior_nodes(t, b, load, store); });
                    for_bnd([&](const SmoothedRow& row) { upd(row.g0, row.slave_begin, row.slave_end); },
                            [&](const JunctionRow& row) { upd(row.self, row.slave_begin, row.slave_end); },
                            [&](const SlidingRow& row) { upd(row.self, row.slave_begin, row.slave_end); });
                }
                reduce(accC);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (S.done[c]) continue;
                    S.norm_r[c] = sqrt(accC[c]);
                    if (S.norm_r[c] <= S.tol_eff[c]) { S.done[c] = 1; continue; }
                    S.rho_old[c] = S.rho_new[c];
                    S.rho_new[c] = accC[2 + c];
                    if (S.iters[c] >= a.max_iters) { S.done[c] = 3; continue; }
                    if (fabs(S.rho_new[c]) < eps) { S.done[c] = 2; continue; }
                    S.beta[c] = (S.rho_new[c] / S.rho_old[c]) * (S.alpha[c] / S.omega[c]);
                }
            }
            // ---- x += d on the rows of this component; copies follow their root (x_copy = x_root + shift) ----
            {
                auto load = [&](int64_t k) { const double2 dd = a.d[k]; double2 xx = a.xnew[k]; xx.x += dd.x; xx.y += dd.y; return xx; };
                auto upd = [&](int64_t k, int sb, int se) {
                    const double2 xx = load(k);
                    a.xnew[k] = xx;
                    for (int q = sb; q < se; ++q) { const SlaveRow sl = a.slaves[q]; a.xnew[sl.self] = make_double2(xx.x + sl.sx, xx.y + sl.sy); }
                };
                for_tiles([&](const WTile& t, const DevBlock& b) { k_interior_nodes(t, b, load, [&](int64_t k, double2 xx) { a.xnew[k] = xx; }); });
                for_bnd([&](const SmoothedRow& row) { upd(row.g0, row.slave_begin, row.slave_end); },
                        [&](const JunctionRow& row) { upd(row.self, row.slave_begin, row.slave_end); },
                        [&](const SlidingRow& row) { upd(row.self, row.slave_begin, row.slave_end); });
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
