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
//   * An iteration is TWO phases inside the kernel, each ended by one exchange between the CTAs of the group that is barrier
//     and reduction at once (a stamped record per CTA, see reduce(); co-residency comes from the cooperative launch):
//         A  v = A p  with  p = r + beta (p - omega v),  r = s - omega t  formed on the fly at every stencil node (p and v
//            ping-pong); the owner also stores r and takes d += omega s of the previous iteration;  rhat.v
//         B  t = A s  with  s = r - alpha v  formed on the fly likewise; d += alpha p;  s.s, t.s, t.t, rhat.s, rhat.t
//     The third synchronisation of the textbook iteration (||r||, rhat.r after r = s - omega t; BiCGStab.zig:349-360) is
//     replaced by algebra on B's sums: ||r||^2 = s.s - 2 omega t.s + omega^2 t.t,  rhat.r = rhat.s - omega rhat.t.  These
//     recurrences only steer the iteration; convergence is always confirmed on the true residual (R0 of the next cycle).
//     The partial dot products of a phase are combined redundantly by every CTA in a fixed order, so all CTAs of a group hold
//     bit-identical scalars.
//   * Rows are evaluated by their owner: interior nodes by warp tiles (32 columns x a few rows, 3-row register window on
//     the vector and on the lagged coordinates), interface / junction / sliding rows one per thread; `connected` copies
//     are mirrored by the thread of their root, so no separate copy pass exists.
//
//   * COARSE: two-level right preconditioner M^-1 = I + P G P^T over aggregates (krylov_coarse.cuh), without an extra
//     barrier.  Everything coarse is linear, so the kernel keeps e_x = G P^T x for x = r, v, t (and by recurrence for p, s)
//     and the phases work with the `hatted` directions p^ = p + P e_p, s^ = s + P e_s formed on the fly:
//         A  p^ = (r + P e_r) + beta (p^ - omega (v + P e_v));  v = A p^          (p^ is what is stored)
//         B  s^ = (r + P e_r) - alpha (v + P e_v);  t = A s^;  d += alpha p^      (s is stored un-hatted: it is the residual)
//         then  e_r = e_r - alpha e_v - omega e_t  (local), and the next A takes d += omega (s + P e_s), e_s = e_r + omega e_t
//     A phase that produces a vector (r, v, t) also leaves its restriction: every warp tile sums its nodes per 8-lane
//     segment into a CONTRIBUTION slot, every boundary row has a slot of its own; after the phase's barrier a CTA adds
//     the slots of each aggregate in the fixed order of a host-built list (c = P^T x, all of it: a few KB), and forms
//     e = G c only for the aggregates its own rows touch (host-built list per CTA).  No atomics: bit-reproducible.
//
// The arithmetic of a row is the one of kernels.cuh (difference form, row-scaled system D^-1 A x = D^-1 b).
#pragma once
#include <type_traits>
#include "kernels.cuh"

namespace tmesh {

constexpr int K_THREADS = 256;
constexpr int K_WARPS = K_THREADS / 32;
constexpr int K_NACC = 10;       // sums a phase reduces
constexpr int K_TILE_ROWS = 2;   // rows of a warp tile (at most; 4 was tried: spills and 1.5 x slower on LS89 x4)
constexpr int K_MAX_GROUP = 160; // CTAs of a group (one per SM)
constexpr int K_REC = 16;        // doubles of a CTA's record in the exchange (128 bytes): K_NACC sums, the stamp in the last one

struct WTile { int32_t block, i0, j0, rows; };   // 32 columns starting at interior column j0, `rows` rows starting at interior row i0
// STRIP tiles (phased path only): the last <= 16 columns of a block, whose own tile would fill at most half of the warp, packed
// four (<= 8 columns) or two (<= 16) row groups to a warp -- with W = 8 or 16 lanes per group, lanes W g .. W g + W - 1 march rows
// i0 + g * rows .. of columns j0 .. j0 + W - 1.  `rows` then carries the number of row groups in bits 16 .. 23 and W in bits 24 .. 31.
__host__ __device__ inline int wtile_rows(const WTile& t) { return t.rows & 0xffff; }
__host__ __device__ inline int wtile_groups(const WTile& t) { return (t.rows >> 16) & 0xff; }   // 0: a plain tile
__host__ __device__ inline int wtile_width(const WTile& t) { return (t.rows >> 24) & 0xff; }
struct WLane { int i0, j; bool active; };
__device__ __forceinline__ WLane wtile_lane(const WTile& t, const DevBlock& b) {
    const int lane = threadIdx.x & 31, groups = wtile_groups(t);
    WLane l;
    if (groups == 0) {
        l.i0 = t.i0; l.j = t.j0 + lane; l.active = l.j <= b.nj - 2;
    } else {
        const int W = wtile_width(t), g = lane / W;
        l.i0 = t.i0 + g * wtile_rows(t); l.j = t.j0 + (lane - g * W);
        l.active = g < groups && l.j <= b.nj - 2 && l.i0 <= b.ni - 2;
        if (!l.active) l.i0 = t.i0;
    }
    return l;
}

struct KComp {   // one independent system: ranges into the component-sorted tables
    int32_t wt_begin, wt_end;
    int32_t s_begin, s_end, j_begin, j_end, l_begin, l_end;
    int32_t rt_begin, rt_end;
    int32_t nodes;
    int32_t bnd_warps;   // warps of the group (from its end) that take the boundary rows and no tiles
    int32_t bw_s, bw_j;  // of these, the first bw_s take the interface rows, the next bw_j the junction rows, the rest the sliding rows
                         // (a warp with rows of one kind does not serialise three row evaluators); 0, 0: all kinds on all of them
};
// which warp (counted from the END of the group) and which rows of its kind a boundary-row warp takes; shared with the host (krylov.inl)
struct KBndPart { int kind, first, stride, count, base; };   // rows first + lane * stride, + 32 * stride, ... < count of `kind`; q = base + row
__host__ __device__ inline KBndPart k_bnd_part(const KComp& K, int back) {
    const int n_s = K.s_end - K.s_begin, n_j = K.j_end - K.j_begin, n_l = K.l_end - K.l_begin;
    if (K.bw_s == 0 && K.bw_j == 0) return KBndPart{-1, back, K.bnd_warps, n_s + n_j + n_l, 0};
    if (back < K.bw_s) return KBndPart{0, back, K.bw_s, n_s, 0};
    if (back < K.bw_s + K.bw_j) return KBndPart{1, back - K.bw_s, K.bw_j, n_j, n_s};
    return KBndPart{2, back - K.bw_s - K.bw_j, K.bnd_warps - K.bw_s - K.bw_j, n_l, n_s + n_j};
}
struct KCtl {    // per component; index 0 = x solve, 1 = y solve
    double tol[2], norm_b[2], norm_r[2];
    int32_t done[2];     // 1 converged, 2 breakdown, 3 iteration cap
    int32_t iters[2];
    int32_t applications, cycles;
};
struct KCoarse {         // coarse space of one component (krylov_coarse.cuh)
    int32_t nc;          // aggregates (0: none, plain Jacobi)
    int32_t agg_base;    // first aggregate in the mesh-wide numbering of the contribution lists
    int32_t slot_base;   // first contribution slot: 4 per warp tile (8-lane segments) in tile order, then one per boundary row
    int32_t need_base;   // lists of the aggregates a CTA needs e on: need_ptr[need_base + CTA within the group]
    int64_t g_off;       // G = (P^T A P)^-1, row-major nc x nc
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
    double* partials;            // 2 x n_ctas records of K_REC doubles {sums ..., stamp}: see reduce()
    unsigned long long epoch;    // launch counter (> 0): the high half of the stamps
    const double2* xc;           // lagged coordinates (the mesh before this outer iteration)
    const double2* pq;           // control function (HAS_PQ)
    double2* xnew;               // iterate, warm-started from xc by the caller
    double2 *r, *rhat, *p[2], *v[2], *s, *t, *d;   // p and v ping-pong: a phase reads the old field at its neighbours while it writes the new one
    double rtol, atol;
    int32_t max_iters, max_restarts, n_ctas_total, polish;   // polish: refinement cycles after convergence (each asks for 10x less residual)
    // COARSE
    const KCoarse* coarse;
    const int32_t* coarse_ok;    // per component: G is usable
    const int32_t* agg;          // per node: aggregate within its component, -1 outside the coarse space
    const int32_t *contrib_ptr, *contrib_src;   // per aggregate (mesh-wide numbering): the slots to add, in order
    const int32_t *need_ptr, *need;
    const double* G;
    double2* contrib;            // 2 x n_slots (ping-pong like `partials`)
    int64_t n_slots;
    int32_t nc_max;
    int32_t need_max, src_max;   // shared-memory caches of a CTA's static tables: its needed aggregates + their rows of G (need_max > 0), the slot lists (src_max > 0)
    int32_t _pad2;
    long long* timing;           // optional (TM_KRYLOV_TIMING): 8 cycle counters per CTA -- R0, A, B, C compute / barrier + sums / coarse products / rest
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
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
    const int nj = b.nj;
    const WLane wl = wtile_lane(t, b);
    const bool active = wl.active;
    const int jc = active ? wl.j : nj - 2;
    const int i_end = min(wl.i0 + wtile_rows(t), b.ni - 1);
    const double2* cb = xc + b.off;
    int64_t l = (int64_t)(wl.i0 - 1) * nj + jc;
    double2 Cm = u(b.off + l), Dm = u(b.off + l + 1) - u(b.off + l - 1);
    double2 cCm = ldg2(cb + l);
    l += nj;
    double2 lf = u(b.off + l - 1), rt = u(b.off + l + 1);
    double2 C0 = u(b.off + l), D0 = rt - lf, R0 = (rt - C0) + (lf - C0);
    double2 cC0 = ldg2(cb + l), cD0 = ldg2(cb + l + 1) - ldg2(cb + l - 1);
#pragma unroll 1
    for (int i = wl.i0; i < i_end; ++i) {
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
    const WLane wl = wtile_lane(t, b);
    if (!wl.active) return;
    const int j = wl.j;
    const int i_end = min(wl.i0 + wtile_rows(t), b.ni - 1);
#pragma unroll 4
    for (int i = wl.i0; i < i_end; ++i) {
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
    double pend[2];   // omega of the last iteration whose d += omega s^ is still to be taken (by the next phase A or the final update)
    int done[2], iters[2];
};

template <bool HAS_PQ, bool COARSE>
__global__ void __launch_bounds__(K_THREADS) bicgstab_persistent_kernel(const KArgs a) {
    __shared__ double sh_part[K_WARPS][K_NACC];
    __shared__ double sh_red[K_NACC];
    __shared__ double sh_all[K_MAX_GROUP][K_NACC];
    extern __shared__ double2 sh_coarse[];                          // COARSE: c (restriction), e_r, e_v, e_t: nc_max each
    double2* const c_full = sh_coarse;
    double2* const e_r = sh_coarse + (COARSE ? a.nc_max : 0);
    double2* const e_v = sh_coarse + (COARSE ? 2 * a.nc_max : 0);
    double2* const e_t = sh_coarse + (COARSE ? 3 * a.nc_max : 0);
    double2* const sh_val = sh_coarse + (COARSE ? 4 * a.nc_max : 0);                                       // src_max: the slots of a phase, staged
    double* const sh_G = reinterpret_cast<double*>(sh_val + (COARSE ? a.src_max : 0));                    // need_max x nc_max
    int32_t* const sh_need = reinterpret_cast<int32_t*>(sh_G + (COARSE ? (size_t)a.need_max * a.nc_max : 0));   // need_max
    int32_t* const sh_cptr = sh_need + (COARSE ? a.need_max : 0);                                         // nc_max + 1
    int32_t* const sh_csrc = sh_cptr + (COARSE ? a.nc_max + 1 : 0);                                       // src_max
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const KGroup G = a.groups[a.cta_group[blockIdx.x]];
    KBarrier* bar = a.bars + a.cta_group[blockIdx.x];
    const int crank = (int)blockIdx.x - G.cta_begin;                // CTA within the group
    const int gwarp = crank * K_WARPS + warp, gwarps = G.n_ctas * K_WARPS;
    const int gthread = crank * K_THREADS + tid, gthreads = G.n_ctas * K_THREADS;
    unsigned int gen = ld_acquire_gpu(&bar->gen);                    // nobody can have passed a barrier yet
    int parity = 0;
    const double eps = 1e-30;                                        // breakdown_eps, BiCGStab.zig:280
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
    auto tick = [&](int slot) { if (a.timing) { const long long now = clock64(); tacc[slot] += now - tlast; tlast = now; } };

    // Barrier + combination of the group's partial sums in one exchange (fixed order: bit-identical in every CTA).  A CTA
    // writes its sums into a 128-byte record and publishes the record's STAMP with a release store (cumulative over what
    // its threads wrote before the __syncthreads); thread c of every CTA polls the stamp of CTA c with acquire loads and,
    // once it carries this exchange's number, reads that record: when all are in, every CTA has arrived AND the sums are
    // here -- no arrival counter, and only one polled word per pair of CTAs (polling stamped VALUES costs n^2 N loads per round:
    // 12.7k cycles per exchange for 99 CTAs and 10 sums against 6.9k this way, scripts/micro/exchange_bench.cu).  Two buffers:
    // a CTA can be at most one exchange ahead of the slowest.
    unsigned long long stamp = a.epoch << 32;
    auto reduce = [&](double (&acc)[K_NACC], auto n_tag) {
        constexpr int N = decltype(n_tag)::value;                    // sums in use: acc[0 .. N)
#pragma unroll
        for (int k = 0; k < N; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < N; ++k) sh_part[warp][k] = acc[k];
        }
        __syncthreads();
        stamp += 1;
        double* const mine = a.partials + ((size_t)parity * a.n_ctas_total + blockIdx.x) * K_REC;
        if (tid < N) {
            double s = 0.0;
            for (int w = 0; w < K_WARPS; ++w) s += sh_part[w][tid];
            __stcg(mine + tid, s);
        }
        __syncthreads();
        if (tid == 0) st_release_u64(reinterpret_cast<unsigned long long*>(mine + K_REC - 1), stamp);
        {
            const double* base = a.partials + ((size_t)parity * a.n_ctas_total + G.cta_begin) * K_REC;
            for (int c = tid; c < G.n_ctas; c += K_THREADS) {
                const double* rec = base + (size_t)c * K_REC;
                while (ld_acquire_u64(reinterpret_cast<const unsigned long long*>(rec + K_REC - 1)) != stamp) {}
#pragma unroll
                for (int k = 0; k < N; k += 2) {
                    const double2 v = __ldcg(reinterpret_cast<const double2*>(rec + k));
                    sh_all[c][k] = v.x;
                    if (k + 1 < N) sh_all[c][k + 1] = v.y;
                }
            }
        }
        __syncthreads();
        if (warp == 0) {
            double s[N];
#pragma unroll
            for (int k = 0; k < N; ++k) s[k] = 0.0;
            for (int c = lane; c < G.n_ctas; c += 32) {
#pragma unroll
                for (int k = 0; k < N; ++k) s[k] += sh_all[c][k];
            }
#pragma unroll
            for (int k = 0; k < N; ++k) s[k] = warp_sum(s[k]);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < N; ++k) sh_red[k] = s[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < N; ++k) acc[k] = sh_red[k];
        parity ^= 1;
    };
    using N2 = std::integral_constant<int, 2>;
    using N4 = std::integral_constant<int, 4>;
    using N10 = std::integral_constant<int, 10>;

    for (int gc = G.comp_begin; gc < G.comp_end; ++gc) {
        const int comp = a.group_comps[gc];
        const KComp K = a.comps[comp];
        const int n_s = K.s_end - K.s_begin, n_j = K.j_end - K.j_begin, n_l = K.l_end - K.l_begin, n_bnd = n_s + n_j + n_l;
        KScal S;
        int applications = 0, cycle = 0, pb = 0, vb = 0, polish_left = a.polish;
        KCoarse CS{};
        if (COARSE) {
            CS = a.coarse[comp];
            if (!a.coarse_ok[comp]) CS.nc = 0;
        }
        const int nc = CS.nc;                                        // uniform over the group
        const int n_tiles = K.wt_end - K.wt_begin;
        const int tile_warps = gwarps - K.bnd_warps;
        const double2 z2 = make_double2(0.0, 0.0);
        int n_need = 0, n_src = 0;
        const int32_t* need = nullptr;
        const int32_t *cptr = nullptr, *csrc = nullptr;              // slot lists, indexed from the component's first aggregate / first list entry
        if (COARSE && nc > 0) {
            // static tables of this CTA for this component -> shared memory (once per solve): the iteration then only fetches slots
            const int nb = a.need_ptr[CS.need_base + crank];
            n_need = a.need_ptr[CS.need_base + crank + 1] - nb;
            need = a.need + nb;
            const int src0 = a.contrib_ptr[CS.agg_base];
            cptr = a.contrib_ptr + CS.agg_base; csrc = a.contrib_src + src0;
            __syncthreads();                                         // the previous component's tables are no longer read
            for (int J = tid; J < nc; J += K_THREADS) { e_r[J] = z2; e_v[J] = z2; e_t[J] = z2; }
            if (a.need_max > 0) {
                for (int q = tid; q < n_need; q += K_THREADS) sh_need[q] = need[q];
                const double* Gc = a.G + CS.g_off;
                for (int q = warp; q < n_need; q += K_WARPS) {
                    const double* row = Gc + (size_t)need[q] * nc;
                    for (int J = lane; J < nc; J += 32) sh_G[(size_t)q * nc + J] = __ldg(row + J);
                }
                need = sh_need;
            }
            if (a.src_max > 0) {
                n_src = a.contrib_ptr[CS.agg_base + nc] - src0;
                for (int J = tid; J <= nc; J += K_THREADS) sh_cptr[J] = cptr[J] - src0;
                for (int q = tid; q < n_src; q += K_THREADS) sh_csrc[q] = csrc[q];
                cptr = sh_cptr; csrc = sh_csrc;
            }
            __syncthreads();
        }
        const int src_shift = (COARSE && nc > 0 && a.src_max == 0) ? a.contrib_ptr[CS.agg_base] : 0;
        // restriction: a warp tile leaves the sums of its four 8-lane segments, a boundary row its value
        auto tile_contrib = [&](double2* cw, int w, double2 sum) {
            if (COARSE && nc > 0) {
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) { sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o); sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o); }
                if ((lane & 7) == 0) __stcg(cw + (size_t)(w - K.wt_begin) * 4 + (lane >> 3), sum);
            }
        };
        auto row_contrib = [&](double2* cw, int q, double2 val) {
            if (COARSE && nc > 0) __stcg(cw + (size_t)n_tiles * 4 + q, val);
        };
        // after the phase's barrier: c = P^T x from the slots (fixed order), e = G c on the aggregates this CTA's rows touch
        auto coarse_post = [&](int cpar, double2* e_dst) {
            if (!COARSE || nc == 0) return;
            const double2* cs = a.contrib + (size_t)cpar * a.n_slots;
            if (a.src_max > 0) {
                // all slots of the component in sweeps of independent loads (the lists have very different lengths: an aggregate
                // between two interfaces has 50 entries, most have 8), then the sums from shared memory
                for (int q0 = tid; q0 < n_src; q0 += 8 * K_THREADS) {
                    double2 tmp[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int q = q0 + u * K_THREADS;
                        if (q < n_src) tmp[u] = __ldcg(cs + sh_csrc[q]);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int q = q0 + u * K_THREADS;
                        if (q < n_src) sh_val[q] = tmp[u];
                    }
                }
                __syncthreads();
                for (int J = tid; J < nc; J += K_THREADS) {
                    double2 sum = z2;
                    const int qe = sh_cptr[J + 1];
                    for (int q = sh_cptr[J]; q < qe; ++q) sum = sum + sh_val[q];
                    c_full[J] = sum;
                }
            } else {
                for (int J = tid; J < nc; J += K_THREADS) {
                    double2 sum = z2;
                    const int qe = cptr[J + 1] - src_shift;
#pragma unroll 4
                    for (int q = cptr[J] - src_shift; q < qe; ++q) sum = sum + __ldcg(cs + csrc[q]);
                    c_full[J] = sum;
                }
            }
            __syncthreads();
            if (a.need_max > 0) {
                // a quarter warp per needed aggregate: its row of G from shared memory
                const int sub = lane >> 3, l8 = lane & 7;
                for (int q = warp * 4 + sub; q < ((n_need + 3) & ~3); q += K_WARPS * 4) {
                    double sx = 0.0, sy = 0.0;
                    if (q < n_need) {
                        const double* row = sh_G + (size_t)q * nc;
                        for (int J = l8; J < nc; J += 8) {
                            const double g = row[J];
                            const double2 c = c_full[J];
                            sx += g * c.x; sy += g * c.y;
                        }
                    }
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
                    if (q < n_need && l8 == 0) e_dst[sh_need[q]] = make_double2(sx, sy);
                }
            } else {
                const double* Gc = a.G + CS.g_off;
                for (int q = warp; q < n_need; q += K_WARPS) {
                    const int Jp = need[q];
                    const double* row = Gc + (size_t)Jp * nc;
                    double sx = 0.0, sy = 0.0;
                    for (int J = lane; J < nc; J += 32) {
                        const double g = __ldg(row + J);
                        const double2 c = c_full[J];
                        sx += g * c.x; sy += g * c.y;
                    }
                    sx = warp_sum(sx); sy = warp_sum(sy);
                    if (lane == 0) e_dst[Jp] = make_double2(sx, sy);
                }
            }
            __syncthreads();
        };
        auto e_at = [&](const double2* e, int64_t k) {
            const int J = __ldg(a.agg + k);
            return J >= 0 ? e[J] : z2;
        };
#pragma unroll
        for (int c = 0; c < 2; ++c) { S.tol[c] = 0.0; S.norm_b[c] = 0.0; S.norm_r[c] = 0.0; S.done[c] = 0; S.iters[c] = 0; }

        // Row-owner loops.  Interior nodes: warp tiles.  Boundary rows: one thread per row, taken from the END of the group's
        // threads (the warps without a tile).  `mirror` hands a value to the `connected` copies of a row's node.
auto mirror = [&](int sb, int se, auto&& put_copy) {   // put_copy(node of the copy) stores every field of the phase
            // the indices of up to four copies are fetched before anything is stored: a store may alias the table as far as the
            // compiler knows, so index loads behind stores would cost an L2 round trip each (a junction node has up to four copies)
            for (int k0 = sb; k0 < se; k0 += 4) {
                int64_t id[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) id[u] = k0 + u < se ? a.slaves[k0 + u].self : (int64_t)-1;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (id[u] >= 0) put_copy(id[u]);
            }
        };
        auto for_tiles = [&](auto&& fn) {
            if (gwarp >= tile_warps) return;
            for (int w = K.wt_begin + gwarp; w < K.wt_end; w += tile_warps) {
                const WTile t = a.wtiles[w];
                fn(w, t, a.blocks[t.block]);
            }
        };
        // Boundary rows: spread over the K.bnd_warps last warps of the group, which have no tiles (krylov.inl sizes them so
        // that every warp has about the same work) -- consecutive rows on consecutive warps, so a warp holds few rows of
        // each kind and no CTA is a straggler.  Mirrored by the host (coarse_plan_build).
        const int w_bnd = K.bnd_warps;
        auto for_bnd = [&](auto&& fs, auto&& fj, auto&& fl) {
            const int back = gwarps - 1 - gwarp;
            if (back >= w_bnd) return;
            const KBndPart part = k_bnd_part(K, back);
            for (int row = part.first + lane * part.stride; row < part.count; row += 32 * part.stride) {
                const int q = part.base + row;
                if (part.kind == 0 || (part.kind < 0 && q < n_s)) fs(q, a.srows[K.s_begin + q]);
                else if (part.kind == 1 || (part.kind < 0 && q < n_s + n_j)) fj(q, a.jrows[K.j_begin + q - n_s]);
                else fl(q, a.lrows[K.l_begin + q - n_s - n_j]);
            }
        };
        for (;; ++cycle) {
            // ---- R0: r = D^-1 (b - A x); rhat = r; p = v = d = 0; ||r||^2 (and ||b||^2 in the first cycle) ----
            double acc[K_NACC] = {};
            const int cpar0 = parity;
            {
                double2* const cw = a.contrib + (size_t)parity * a.n_slots + CS.slot_base;
                double2* const P0 = a.p[pb];
                double2* const V0 = a.v[vb];
                const double2 z = make_double2(0.0, 0.0);
                auto xval = [&](int64_t k) { return a.xnew[k]; };
                auto init = [&](int64_t k, double2 res, int sb, int se) {   // phase A forms r = s - omega t: s = r, t = 0
                    a.s[k] = res; a.t[k] = z; a.rhat[k] = res; P0[k] = z; V0[k] = z; a.d[k] = z;
                    mirror(sb, se, [&](int64_t c) { a.s[c] = res; a.t[c] = z; P0[c] = z; V0[c] = z; });
                    acc[0] += res.x * res.x; acc[1] += res.y * res.y;
                };
                for_tiles([&](int w, const WTile& t, const DevBlock& b) {
                    double2 tsum = z2;
                    k_interior<MODE_RESID, HAS_PQ>(t, b, xval, a.xc, a.pq, [&](int64_t k, double2 res, double2) { init(k, res, 0, 0); tsum = tsum + res; });
                    tile_contrib(cw, w, tsum);
                });
                for_bnd([&](int q, const SmoothedRow& row) {
                            double b2x = 0.0, b2y = 0.0;
                            const KRow o = k_smoothed<MODE_RESID, HAS_PQ>(row, xval, a.xc, a.pq, b2x, b2y);
                            init(row.g0, o.res, row.slave_begin, row.slave_end);
                            row_contrib(cw, q, o.res);
                            if (cycle == 0) { acc[2] += b2x; acc[3] += b2y; }
                        },
                        [&](int q, const JunctionRow& row) {
                            const double2 res = k_junction<MODE_RESID>(row, xval).res;
                            init(row.self, res, row.slave_begin, row.slave_end);
                            row_contrib(cw, q, res);
                        },
                        [&](int, const SlidingRow& row) { init(row.self, k_sliding<MODE_RESID>(row, xval).res, row.slave_begin, row.slave_end); });
                if (cycle == 0) {  // constant part of the reference's ||b||^2 (BiCGStab.zig:289-291)
                    for (int q = K.rt_begin + gthread; q < K.rt_end; q += gthreads) {
                        const RhsTerm t = a.rterms[q];
                        double bx = t.cx, by = t.cy;
                        if (t.from_x | t.from_y) { const double2 x0 = ldg2(a.xc + t.g); if (t.from_x) bx = x0.x; if (t.from_y) by = x0.y; }
                        acc[2] += bx * bx; acc[3] += by * by;
                    }
                }
            }
            tick(0);
            reduce(acc, N4{});
            tick(4);
            applications += 1;
            coarse_post(cpar0, e_r);
            tick(5);
            if (COARSE && nc > 0) {                                  // v = 0
                for (int J = tid; J < nc; J += K_THREADS) e_v[J] = z2;
                __syncthreads();
            }
            const bool stop = cycle > a.max_restarts;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                S.norm_r[c] = sqrt(acc[c]);
                if (cycle == 0) { S.norm_b[c] = sqrt(acc[2 + c]); S.tol[c] = fmax(a.atol, a.rtol * S.norm_b[c]); }   // GMRES.zig:305-306 / BiCGStab.zig:291
                S.tol_eff[c] = S.tol[c];
                S.rho_old[c] = 1.0; S.alpha[c] = 1.0; S.omega[c] = 1.0; S.pend[c] = 0.0;
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
                // A: r = s - omega t and p = r + beta (p - omega v) formed at every stencil node; v = A p; rhat . v
                double accA[K_NACC] = {};
                const int cparA = parity;
                {
                    double2* const cw = a.contrib + (size_t)parity * a.n_slots + CS.slot_base;
                    const bool dx = S.done[0] != 0, dy = S.done[1] != 0;
                    const double bx = S.beta[0], by = S.beta[1], ox = S.omega[0], oy = S.omega[1];
                    const double qx = S.pend[0], qy = S.pend[1];
                    const double2* const Pold = a.p[pb];
                    const double2* const Vold = a.v[vb];
                    double2* const Pnew = a.p[pb ^ 1];
                    double2* const Vnew = a.v[vb ^ 1];
                    auto pval = [&](int64_t k) {
                        const double2 ss = a.s[k], tt = a.t[k], pp = Pold[k];
                        double2 vv = Vold[k];
                        double2 rr = make_double2(ss.x - ox * tt.x, ss.y - oy * tt.y);
                        if (COARSE && nc > 0) { rr = rr + e_at(e_r, k); vv = vv + e_at(e_v, k); }
                        return make_double2(dx ? 0.0 : rr.x + bx * (pp.x - ox * vv.x), dy ? 0.0 : rr.y + by * (pp.y - oy * vv.y));
                    };
                    auto put = [&](int64_t k, const KRow& o, int sb, int se) {
                        const double2 ss = a.s[k], tt = a.t[k];
                        const double2 rr = make_double2(ss.x - ox * tt.x, ss.y - oy * tt.y);
                        if (qx != 0.0 || qy != 0.0) {                       // the omega part of the previous iteration
                            double2 sh = ss;                                // COARSE: s^ = s + P e_s, e_s = e_r + omega e_t (e_r is the new one)
                            if (COARSE && nc > 0) {
                                const int J = __ldg(a.agg + k);
                                if (J >= 0) { const double2 er = e_r[J], et = e_t[J]; sh.x += er.x + qx * et.x; sh.y += er.y + qy * et.y; }
                            }
                            double2 dd = a.d[k];
                            dd.x += qx * sh.x; dd.y += qy * sh.y;
                            a.d[k] = dd;
                        }
                        a.r[k] = rr; Pnew[k] = o.centre; Vnew[k] = o.res;
                        mirror(sb, se, [&](int64_t c) { a.r[c] = rr; Pnew[c] = o.centre; Vnew[c] = o.res; });
                        const double2 h = a.rhat[k];
                        accA[0] += h.x * o.res.x; accA[1] += h.y * o.res.y;
                    };
                    for_tiles([&](int w, const WTile& t, const DevBlock& b) {
                        double2 tsum = z2;
                        k_interior<MODE_APPLY, HAS_PQ>(t, b, pval, a.xc, a.pq, [&](int64_t k, double2 res, double2 c) { put(k, KRow{res, c}, 0, 0); tsum = tsum + res; });
                        tile_contrib(cw, w, tsum);
                    });
                    double u0, u1;
                    for_bnd([&](int q, const SmoothedRow& row) {
                                const KRow o = k_smoothed<MODE_APPLY, HAS_PQ>(row, pval, a.xc, a.pq, u0, u1);
                                put(row.g0, o, row.slave_begin, row.slave_end);
                                row_contrib(cw, q, o.res);
                            },
                            [&](int q, const JunctionRow& row) {
                                const KRow o = k_junction<MODE_APPLY>(row, pval);
                                put(row.self, o, row.slave_begin, row.slave_end);
                                row_contrib(cw, q, o.res);
                            },
                            [&](int, const SlidingRow& row) { put(row.self, k_sliding<MODE_APPLY>(row, pval), row.slave_begin, row.slave_end); });
                }
                tick(1);
                reduce(accA, N2{});
                tick(4);
                pb ^= 1; vb ^= 1;
                applications += 1;
                S.pend[0] = 0.0; S.pend[1] = 0.0;
                coarse_post(cparA, e_v);
                tick(5);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (S.done[c]) continue;
                    if (fabs(accA[c]) < eps) { S.done[c] = 2; S.alpha[c] = 0.0; }
                    else S.alpha[c] = S.rho_new[c] / accA[c];
                }
                // B: s = r - alpha v formed at every stencil node; d += alpha p; t = A s; s.s, t.s, t.t, rhat.s, rhat.t
                double accB[K_NACC] = {};
                const int cparB = parity;
                {
                    double2* const cw = a.contrib + (size_t)parity * a.n_slots + CS.slot_base;
                    const bool mx = S.done[0] != 0, my = S.done[1] != 0;   // a component that just broke down is masked from here on
                    const double ax = S.alpha[0], ay = S.alpha[1];
                    const double2* const Pcur = a.p[pb];
                    const double2* const Vcur = a.v[vb];
                    auto sval = [&](int64_t k) {
                        double2 rr = a.r[k], vv = Vcur[k];
                        if (COARSE && nc > 0) { rr = rr + e_at(e_r, k); vv = vv + e_at(e_v, k); }
                        return make_double2(mx ? 0.0 : rr.x - ax * vv.x, my ? 0.0 : rr.y - ay * vv.y);
                    };
                    auto put = [&](int64_t k, const KRow& o, int sb, int se) {
                        const double2 pp = Pcur[k];
                        double2 dd = a.d[k];
                        if (!mx) dd.x += ax * pp.x;
                        if (!my) dd.y += ay * pp.y;
                        double2 ss = o.centre;                              // COARSE: the centre is s^; the residual s is kept
                        if (COARSE && nc > 0) {
                            const double2 rr = a.r[k], vv = Vcur[k];
                            ss = make_double2(mx ? 0.0 : rr.x - ax * vv.x, my ? 0.0 : rr.y - ay * vv.y);
                        }
                        a.s[k] = ss; a.t[k] = o.res; a.d[k] = dd;
                        mirror(sb, se, [&](int64_t c) { a.s[c] = ss; a.t[c] = o.res; });
                        const double2 h = a.rhat[k];
                        accB[0] += ss.x * ss.x; accB[1] += ss.y * ss.y;
                        accB[2] += ss.x * o.res.x; accB[3] += ss.y * o.res.y;
                        accB[4] += o.res.x * o.res.x; accB[5] += o.res.y * o.res.y;
                        accB[6] += h.x * ss.x; accB[7] += h.y * ss.y;
                        accB[8] += h.x * o.res.x; accB[9] += h.y * o.res.y;
                    };
                    for_tiles([&](int w, const WTile& t, const DevBlock& b) {
                        double2 tsum = z2;
                        k_interior<MODE_APPLY, HAS_PQ>(t, b, sval, a.xc, a.pq, [&](int64_t k, double2 res, double2 c) { put(k, KRow{res, c}, 0, 0); tsum = tsum + res; });
                        tile_contrib(cw, w, tsum);
                    });
                    double u0, u1;
                    for_bnd([&](int q, const SmoothedRow& row) {
                                const KRow o = k_smoothed<MODE_APPLY, HAS_PQ>(row, sval, a.xc, a.pq, u0, u1);
                                put(row.g0, o, row.slave_begin, row.slave_end);
                                row_contrib(cw, q, o.res);
                            },
                            [&](int q, const JunctionRow& row) {
                                const KRow o = k_junction<MODE_APPLY>(row, sval);
                                put(row.self, o, row.slave_begin, row.slave_end);
                                row_contrib(cw, q, o.res);
                            },
                            [&](int, const SlidingRow& row) { put(row.self, k_sliding<MODE_APPLY>(row, sval), row.slave_begin, row.slave_end); });
                }
                tick(2);
                reduce(accB, N10{});
                tick(4);
                applications += 1;
                coarse_post(cparB, e_t);
                tick(5);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (S.done[c]) continue;
                    S.iters[c] += 1;
                    S.norm_r[c] = sqrt(accB[c]);
                    if (S.norm_r[c] <= S.tol_eff[c]) { S.done[c] = 1; S.omega[c] = 0.0; continue; }   // x += alpha p has been taken; no omega part
                    if (fabs(accB[4 + c]) < eps) { S.done[c] = 2; S.omega[c] = 0.0; continue; }
                    S.omega[c] = accB[2 + c] / accB[4 + c];
                    if (fabs(S.omega[c]) < eps) { S.done[c] = 2; S.omega[c] = 0.0; continue; }
                    // r = s - omega t: its norm and rhat . r from the sums of this phase (BiCGStab.zig:349-360 without the third reduction)
                    const double om = S.omega[c];
                    S.pend[c] = om;
                    S.norm_r[c] = sqrt(fmax(0.0, accB[c] - om * (2.0 * accB[2 + c] - om * accB[4 + c])));
                    if (S.norm_r[c] <= S.tol_eff[c]) { S.done[c] = 1; continue; }
                    S.rho_old[c] = S.rho_new[c];
                    S.rho_new[c] = accB[6 + c] - om * accB[8 + c];
                    if (S.iters[c] >= a.max_iters) { S.done[c] = 3; continue; }
                    if (fabs(S.rho_new[c]) < eps) { S.done[c] = 2; continue; }
                    S.beta[c] = (S.rho_new[c] / S.rho_old[c]) * (S.alpha[c] / om);
                }
                if (COARSE && nc > 0) {                              // e_r follows r = s - omega t = r - alpha v - omega t
                    const double ax = S.alpha[0], ay = S.alpha[1], ox = S.omega[0], oy = S.omega[1];
                    for (int J = tid; J < nc; J += K_THREADS) {
                        const double2 er = e_r[J], ev = e_v[J], et = e_t[J];
                        e_r[J] = make_double2((er.x - ax * ev.x) - ox * et.x, (er.y - ay * ev.y) - oy * et.y);
                    }
                    __syncthreads();
                }
            }
            // ---- x += d on the rows of this component; copies follow their root (x_copy = x_root + shift) ----
            {
                const double qx = S.pend[0], qy = S.pend[1];              // a component that converged on r = s - omega t still owes d += omega s^
                auto load = [&](int64_t k) {
                    double2 dd = a.d[k];
                    if (qx != 0.0 || qy != 0.0) {
                        double2 sh = a.s[k];
                        if (COARSE && nc > 0) {
                            const int J = __ldg(a.agg + k);
                            if (J >= 0) { const double2 er = e_r[J], et = e_t[J]; sh.x += er.x + qx * et.x; sh.y += er.y + qy * et.y; }
                        }
                        dd.x += qx * sh.x; dd.y += qy * sh.y;
                    }
                    double2 xx = a.xnew[k];
                    xx.x += dd.x; xx.y += dd.y;
                    return xx;
                };
                auto upd = [&](int64_t k, int sb, int se) {
                    const double2 xx = load(k);
                    a.xnew[k] = xx;
                    for (int q = sb; q < se; ++q) { const SlaveRow sl = a.slaves[q]; a.xnew[sl.self] = make_double2(xx.x + sl.sx, xx.y + sl.sy); }
                };
                for_tiles([&](int, const WTile& t, const DevBlock& b) { k_interior_nodes(t, b, load, [&](int64_t k, double2 xx) { a.xnew[k] = xx; }); });
                for_bnd([&](int, const SmoothedRow& row) { upd(row.g0, row.slave_begin, row.slave_end); },
                        [&](int, const JunctionRow& row) { upd(row.self, row.slave_begin, row.slave_end); },
                        [&](int, const SlidingRow& row) { upd(row.self, row.slave_begin, row.slave_end); });
                group_barrier(bar, (unsigned)G.n_ctas, gen);
            }
        }
        tick(6);
        if (crank == 0 && tid == 0) {
            KCtl out;
#pragma unroll
            for (int c = 0; c < 2; ++c) { out.tol[c] = S.tol[c]; out.norm_b[c] = S.norm_b[c]; out.norm_r[c] = S.norm_r[c]; out.done[c] = S.done[c]; out.iters[c] = S.iters[c]; }
            out.applications = applications; out.cycles = cycle + 1;
            a.ctl[comp] = out;
        }
    }
    if (a.timing && tid == 0) {
        for (int k = 0; k < 8; ++k) a.timing[(size_t)blockIdx.x * 8 + k] += tacc[k];
    }
}

}  // namespace tmesh
