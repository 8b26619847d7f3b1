// krylov_phased.cuh -- the same BiCGStab iteration as krylov_kernels.cuh (three fused phases A, B, C; one linear system per
// connected component with its own scalars and stopping test), but every phase is ONE LAUNCH over ALL components that are
// still iterating.  This is the throughput form for batches of cuts: 128 T106 cuts are 3.2 M nodes, far beyond L2, so the
// phases are HBM-bound and want many small, high-occupancy CTAs rather than a few fat persistent ones; the components advance
// in lock-step launches but every one keeps its own alpha / omega / rho, tolerance and iteration count, and drops out of the
// launches (its tiles return at once) as soon as it has converged.  Per phase the partial dot products are written per warp
// tile / boundary chunk and combined per component in a fixed order by a small finalisation launch (deterministic).
//
// Bytes per node and iteration: A reads r, p, v, rhat + lagged x (+ P,Q), writes p', v'; B reads r, v, p, d + x (+ P,Q), writes
// s, t, d; C reads s, t, r, d, rhat, writes r, d  =  23 x 16 B = 368 B (+ 32 B of P,Q), against 432 B and 12 launches in round 1.
#pragma once
#include "krylov_kernels.cuh"

namespace tmesh {

constexpr int KP_THREADS = 128;              // 4 warp tiles per CTA
constexpr int KP_WARPS = KP_THREADS / 32;
#ifndef KP_MIN_CTAS
#define KP_MIN_CTAS 4
#endif

struct BChunk { int32_t comp, s_begin, s_end, j_begin, j_end, l_begin, l_end, _pad; };   // <= KP_THREADS boundary rows of ONE component

struct KState {      // per component, lives in global memory; index 0 = x solve, 1 = y solve
    double rho_old[2], rho_new[2], alpha[2], omega[2], beta[2], tol[2], tol_eff[2], norm_b[2], norm_r[2];
    int32_t done[2];         // 0 iterating, 1 converged, 2 breakdown, 3 iteration cap
    int32_t iters[2];
    int32_t final_, cycles, applications, polish_left;
};
struct KPComp { int32_t wt_begin, wt_end, ch_begin, ch_end, rt_begin, rt_end, nodes, _pad; };

struct KPArgs {
    const WTile* wtiles;
    const int32_t* wt_comp;
    const BChunk* chunks;
    const DevBlock* blocks;
    const SmoothedRow* srows;
    const JunctionRow* jrows;
    const SlidingRow* lrows;
    const SlaveRow* slaves;
    const RhsTerm* rterms;
    const KPComp* comps;
    KState* state;
    double* partials;            // (n_wtiles + n_chunks) x K_NACC
    const double2* xc;
    const double2* pq;
    double2* xnew;
    double2 *r, *rhat, *s, *t, *d;
    const double2 *p_old, *v_old;
    double2 *p_new, *v_new;
    int32_t n_wtiles, n_chunks, n_comp, cycle;
    double rtol, atol;
    int32_t max_iters, max_restarts, polish, _pad;
    // COARSE (krylov_coarse.cuh): the two-level preconditioner, where an iteration is bandwidth-bound and halving their number pays
    const KCoarse* coarse;       // per component; need_base unused, slot numbering below
    const int32_t* coarse_ok;
    const int32_t* agg;          // per node: aggregate within its component, -1 outside the coarse space
    const int32_t *contrib_ptr, *contrib_src;
    const double* G;
    double2* contrib;            // slots: 4 per warp tile (8-lane segments) | one per smoothed row | one per junction row
    double2 *e_r, *e_v, *e_t;    // G P^T x per aggregate (mesh-wide numbering)
    int32_t sslot0, jslot0, nc_max, _pad2;
};

enum KPhase : int { KP_R0 = 0, KP_A = 1, KP_B = 2, KP_C = 3, KP_ADD = 4 };

// TILES = true: the interior warp tiles (grid = ceil(n_wtiles / 4)); false: the boundary chunks (grid = n_chunks).  Two kernels
// rather than one so that the register count of the tile path (the bandwidth path) is not set by the gather-heavy row path.
template <int PHASE, bool HAS_PQ, bool TILES, bool COARSE>
__global__ void __launch_bounds__(KP_THREADS, KP_MIN_CTAS) krylov_phase_kernel(const KPArgs a) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_tile_ctas = TILES ? 0x7fffffff : 0;
    const bool is_tile = TILES;
    int comp, slot;
    WTile t{};
    BChunk ch{};
    if (is_tile) {
        const int w = (int)blockIdx.x * KP_WARPS + warp;
        if (w >= a.n_wtiles) return;
        t = a.wtiles[w];
        comp = a.wt_comp[w];
        slot = w;
    } else {
        ch = a.chunks[(int)blockIdx.x - n_tile_ctas];
        comp = ch.comp;
        slot = a.n_wtiles + ((int)blockIdx.x - n_tile_ctas);
    }
    const KState& S = a.state[comp];
    if (S.final_) return;
    const int d0 = S.done[0], d1 = S.done[1];
    if (PHASE != KP_R0 && PHASE != KP_ADD && d0 && d1) return;
    double acc[K_NACC] = {};
    auto mirror = [&](double2* f, int sb, int se, double2 val) {
        for (int k = sb; k < se; ++k) f[a.slaves[k].self] = val;
    };
    // COARSE: the hatted directions of krylov_kernels.cuh; e_x live in global memory (written by krylov_coarse_kernel)
    const double2 z2 = make_double2(0.0, 0.0);
    bool co = false;
    const double2 *er = nullptr, *ev = nullptr;
    if (COARSE) {
        const KCoarse C = a.coarse[comp];
        co = C.nc > 0 && a.coarse_ok[comp] != 0;
        er = a.e_r + C.agg_base; ev = a.e_v + C.agg_base;
    }
    auto e_at = [&](const double2* e, int64_t k) {
        const int J = __ldg(a.agg + k);
        return J >= 0 ? ldg2(e + J) : z2;                          // written by an earlier launch: read-only here
    };
    double2 tsum = z2;                                             // restriction of what the phase produces: per lane, then per 8-lane segment
    auto tile_contrib = [&]() {
        if (COARSE && co) {
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) { tsum.x += __shfl_xor_sync(0xffffffffu, tsum.x, o); tsum.y += __shfl_xor_sync(0xffffffffu, tsum.y, o); }
            if ((lane & 7) == 0) a.contrib[(size_t)slot * 4 + (lane >> 3)] = tsum;
        }
    };
    // boundary rows of a chunk: thread q of the CTA takes row q of the concatenation smoothed | junction | sliding
    auto for_bnd = [&](auto&& fs, auto&& fj, auto&& fl) {
        const int n_s = ch.s_end - ch.s_begin, n_j = ch.j_end - ch.j_begin, n_l = ch.l_end - ch.l_begin;
        const int q = tid;
        if (q < n_s) fs(a.sslot0 + ch.s_begin + q, a.srows[ch.s_begin + q]);
        else if (q < n_s + n_j) fj(a.jslot0 + ch.j_begin + q - n_s, a.jrows[ch.j_begin + q - n_s]);
        else if (q < n_s + n_j + n_l) fl(a.lrows[ch.l_begin + q - n_s - n_j]);
    };
    if (PHASE == KP_R0) {
        // r = D^-1 (b - A x); rhat = r; p = v = d = 0; ||r||^2 (and ||b||^2 in the first cycle)
        double2* const P0 = a.p_new;
        double2* const V0 = a.v_new;
        const double2 z = make_double2(0.0, 0.0);
        auto xval = [&](int64_t k) { return a.xnew[k]; };
        auto init = [&](int64_t k, double2 res, int sb, int se) {
            a.r[k] = res; a.rhat[k] = res; P0[k] = z; V0[k] = z; a.d[k] = z;
            mirror(a.r, sb, se, res); mirror(P0, sb, se, z); mirror(V0, sb, se, z);
            acc[0] += res.x * res.x; acc[1] += res.y * res.y;
        };
        if (is_tile) {
            k_interior_march<MODE_RESID, HAS_PQ>(t, a.blocks[t.block], xval, a.xc, a.pq, [&](int64_t k, double2 res, double2) { init(k, res, 0, 0); tsum = tsum + res; });
            tile_contrib();
        } else {
            for_bnd([&](int sl, const SmoothedRow& row) {
                        double b2x = 0.0, b2y = 0.0;
                        const KRow o = k_smoothed<MODE_RESID, HAS_PQ>(row, xval, a.xc, a.pq, b2x, b2y);
                        init(row.g0, o.res, row.slave_begin, row.slave_end);
                        if (COARSE && co) a.contrib[sl] = o.res;
                        if (a.cycle == 0) { acc[2] += b2x; acc[3] += b2y; }
                    },
                    [&](int sl, const JunctionRow& row) {
                        const double2 res = k_junction<MODE_RESID>(row, xval).res;
                        init(row.self, res, row.slave_begin, row.slave_end);
                        if (COARSE && co) a.contrib[sl] = res;
                    },
                    [&](const SlidingRow& row) { init(row.self, k_sliding<MODE_RESID>(row, xval).res, row.slave_begin, row.slave_end); });
            if (a.cycle == 0 && (int)blockIdx.x == a.comps[comp].ch_begin) {  // the component's first chunk also sums the constant part of ||b||^2
                const KPComp K = a.comps[comp];
                for (int q = K.rt_begin + tid; q < K.rt_end; q += KP_THREADS) {
                    const RhsTerm rt = a.rterms[q];
                    double bx = rt.cx, by = rt.cy;
                    if (rt.from_x | rt.from_y) { const double2 x0 = ldg2(a.xc + rt.g); if (rt.from_x) bx = x0.x; if (rt.from_y) by = x0.y; }
                    acc[2] += bx * bx; acc[3] += by * by;
                }
            }
        }
    } else if (PHASE == KP_A) {
        const bool dx = d0 != 0, dy = d1 != 0;
        const double bx = S.beta[0], by = S.beta[1], ox = S.omega[0], oy = S.omega[1];
        auto pval = [&](int64_t k) {
            double2 rr = a.r[k], vv = a.v_old[k];
            const double2 pp = a.p_old[k];
            if (COARSE && co) { rr = rr + e_at(er, k); vv = vv + e_at(ev, k); }
            return make_double2(dx ? 0.0 : rr.x + bx * (pp.x - ox * vv.x), dy ? 0.0 : rr.y + by * (pp.y - oy * vv.y));
        };
        auto put = [&](int64_t k, const KRow& o, int sb, int se) {
            a.p_new[k] = o.centre; a.v_new[k] = o.res;
            mirror(a.p_new, sb, se, o.centre); mirror(a.v_new, sb, se, o.res);
            const double2 h = a.rhat[k];
            acc[0] += h.x * o.res.x; acc[1] += h.y * o.res.y;
        };
        if (is_tile) {
            k_interior_march<MODE_APPLY, HAS_PQ>(t, a.blocks[t.block], pval, a.xc, a.pq, [&](int64_t k, double2 res, double2 c) { put(k, KRow{res, c}, 0, 0); tsum = tsum + res; });
            tile_contrib();
        } else {
            double u0, u1;
            for_bnd([&](int sl, const SmoothedRow& row) {
                        const KRow o = k_smoothed<MODE_APPLY, HAS_PQ>(row, pval, a.xc, a.pq, u0, u1);
                        put(row.g0, o, row.slave_begin, row.slave_end);
                        if (COARSE && co) a.contrib[sl] = o.res;
                    },
                    [&](int sl, const JunctionRow& row) {
                        const KRow o = k_junction<MODE_APPLY>(row, pval);
                        put(row.self, o, row.slave_begin, row.slave_end);
                        if (COARSE && co) a.contrib[sl] = o.res;
                    },
                    [&](const SlidingRow& row) { put(row.self, k_sliding<MODE_APPLY>(row, pval), row.slave_begin, row.slave_end); });
        }
    } else if (PHASE == KP_B) {
        // p_old / v_old are the CURRENT p and v here (the host hands the buffers phase A wrote)
        const bool mx = d0 != 0, my = d1 != 0;
        const double ax = S.alpha[0], ay = S.alpha[1];
        auto sval = [&](int64_t k) {
            double2 rr = a.r[k], vv = a.v_old[k];
            if (COARSE && co) { rr = rr + e_at(er, k); vv = vv + e_at(ev, k); }
            return make_double2(mx ? 0.0 : rr.x - ax * vv.x, my ? 0.0 : rr.y - ay * vv.y);
        };
        auto put = [&](int64_t k, const KRow& o) {
            const double2 pp = a.p_old[k];
            double2 dd = a.d[k];
            if (!mx) dd.x += ax * pp.x;
            if (!my) dd.y += ay * pp.y;
            double2 ss = o.centre;                                  // COARSE: the centre is s^; the residual s is kept
            if (COARSE && co) {
                const double2 rr = a.r[k], vv = a.v_old[k];
                ss = make_double2(mx ? 0.0 : rr.x - ax * vv.x, my ? 0.0 : rr.y - ay * vv.y);
            }
            a.s[k] = ss; a.t[k] = o.res; a.d[k] = dd;
            acc[0] += ss.x * ss.x; acc[1] += ss.y * ss.y;
            acc[2] += ss.x * o.res.x; acc[3] += ss.y * o.res.y;
            acc[4] += o.res.x * o.res.x; acc[5] += o.res.y * o.res.y;
        };
        if (is_tile) {
            k_interior_march<MODE_APPLY, HAS_PQ>(t, a.blocks[t.block], sval, a.xc, a.pq, [&](int64_t k, double2 res, double2 c) { put(k, KRow{res, c}); tsum = tsum + res; });
            tile_contrib();
        } else {
            double u0, u1;
            for_bnd([&](int sl, const SmoothedRow& row) {
                        const KRow o = k_smoothed<MODE_APPLY, HAS_PQ>(row, sval, a.xc, a.pq, u0, u1);
                        put(row.g0, o);
                        if (COARSE && co) a.contrib[sl] = o.res;
                    },
                    [&](int sl, const JunctionRow& row) {
                        const KRow o = k_junction<MODE_APPLY>(row, sval);
                        put(row.self, o);
                        if (COARSE && co) a.contrib[sl] = o.res;
                    },
                    [&](const SlidingRow& row) { put(row.self, k_sliding<MODE_APPLY>(row, sval)); });
        }
    } else if (PHASE == KP_C) {
        const bool ex = d0 != 0, ey = d1 != 0;
        const double ox = S.omega[0], oy = S.omega[1];
        const double ax = S.alpha[0], ay = S.alpha[1];
        struct RD { double2 r, d, h; };
        auto load = [&](int64_t k) {
            const double2 ss = a.s[k], tt = a.t[k];
            RD o{a.r[k], a.d[k], a.rhat[k]};
            double2 sh = ss;                                        // COARSE: d takes s^ = s + P e_s, e_s = e_r - alpha e_v
            if (COARSE && co) {
                const int J = __ldg(a.agg + k);
                if (J >= 0) { const double2 e1 = ldg2(er + J), e2 = ldg2(ev + J); sh.x += e1.x - ax * e2.x; sh.y += e1.y - ay * e2.y; }
            }
            if (!ex) { o.d.x += ox * sh.x; o.r.x = ss.x - ox * tt.x; }
            if (!ey) { o.d.y += oy * sh.y; o.r.y = ss.y - oy * tt.y; }
            return o;
        };
        auto store = [&](int64_t k, const RD& o) {
            a.r[k] = o.r; a.d[k] = o.d;
            if (!ex) { acc[0] += o.r.x * o.r.x; acc[2] += o.h.x * o.r.x; }
            if (!ey) { acc[1] += o.r.y * o.r.y; acc[3] += o.h.y * o.r.y; }
        };
        auto upd = [&](int64_t k, int sb, int se) { const RD o = load(k); store(k, o); mirror(a.r, sb, se, o.r); };
        if (is_tile) {
            k_interior_nodes_march(t, a.blocks[t.block], load, store);
        } else {
            for_bnd([&](int, const SmoothedRow& row) { upd(row.g0, row.slave_begin, row.slave_end); },
                    [&](int, const JunctionRow& row) { upd(row.self, row.slave_begin, row.slave_end); },
                    [&](const SlidingRow& row) { upd(row.self, row.slave_begin, row.slave_end); });
        }
    } else {  // KP_ADD: x += d on the rows of the component; copies follow their root (x_copy = x_root + shift)
        auto load = [&](int64_t k) { const double2 dd = a.d[k]; double2 xx = a.xnew[k]; xx.x += dd.x; xx.y += dd.y; return xx; };
        auto upd = [&](int64_t k, int sb, int se) {
            const double2 xx = load(k);
            a.xnew[k] = xx;
            for (int q = sb; q < se; ++q) { const SlaveRow sl = a.slaves[q]; a.xnew[sl.self] = make_double2(xx.x + sl.sx, xx.y + sl.sy); }
        };
        if (is_tile) {
            k_interior_nodes_march(t, a.blocks[t.block], load, [&](int64_t k, double2 xx) { a.xnew[k] = xx; });
        } else {
            for_bnd([&](int, const SmoothedRow& row) { upd(row.g0, row.slave_begin, row.slave_end); },
                    [&](int, const JunctionRow& row) { upd(row.self, row.slave_begin, row.slave_end); },
                    [&](const SlidingRow& row) { upd(row.self, row.slave_begin, row.slave_end); });
        }
        return;
    }
    // partial sums: one record per warp tile, one per boundary chunk
    if (is_tile) {
#pragma unroll
        for (int k = 0; k < K_NACC; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K_NACC; ++k) a.partials[(size_t)slot * K_NACC + k] = acc[k];
        }
    } else {
        __shared__ double sh[KP_WARPS][K_NACC];
#pragma unroll
        for (int k = 0; k < K_NACC; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K_NACC; ++k) sh[warp][k] = acc[k];
        }
        __syncthreads();
        if (tid < K_NACC) {
            double s = 0.0;
            for (int w = 0; w < KP_WARPS; ++w) s += sh[w][tid];
            a.partials[(size_t)slot * K_NACC + tid] = s;
        }
    }
}

// One warp per component: sums the component's partial records in a fixed order and advances its scalars exactly as the
// persistent kernel does after the corresponding phase (BiCGStab.zig:303-366).
template <int PHASE>
__global__ void __launch_bounds__(128) krylov_finalize_kernel(const KPArgs a) {
    const int comp = (int)blockIdx.x * 4 + ((int)threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (comp >= a.n_comp) return;
    KState& S = a.state[comp];
    if (S.final_) return;
    if (PHASE != KP_R0 && S.done[0] && S.done[1]) return;
    const KPComp K = a.comps[comp];
    double acc[K_NACC] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int n_t = K.wt_end - K.wt_begin, n_c = K.ch_end - K.ch_begin;
    for (int q = lane; q < n_t + n_c; q += 32) {
        const double* p = a.partials + (size_t)(q < n_t ? K.wt_begin + q : a.n_wtiles + K.ch_begin + (q - n_t)) * K_NACC;
#pragma unroll
        for (int k = 0; k < K_NACC; ++k) acc[k] += p[k];
    }
#pragma unroll
    for (int k = 0; k < K_NACC; ++k) acc[k] = warp_sum(acc[k]);
    if (lane != 0) return;
    const double eps = 1e-30;  // breakdown_eps, BiCGStab.zig:280
    if (PHASE == KP_R0) {
        S.applications += 1;
        S.cycles = a.cycle + 1;
        const bool stop = a.cycle > a.max_restarts;
        for (int c = 0; c < 2; ++c) {
            S.norm_r[c] = sqrt(acc[c]);
            if (a.cycle == 0) { S.norm_b[c] = sqrt(acc[2 + c]); S.tol[c] = fmax(a.atol, a.rtol * S.norm_b[c]); S.iters[c] = 0; S.polish_left = a.polish; }
            S.tol_eff[c] = S.tol[c];
            S.rho_old[c] = 1.0; S.alpha[c] = 1.0; S.omega[c] = 1.0;
            S.rho_new[c] = acc[c];
            S.done[c] = S.norm_r[c] <= S.tol[c] ? 1 : (S.iters[c] >= a.max_iters ? 3 : 0);
            if (!S.done[c] && fabs(S.rho_new[c]) < eps) S.done[c] = 2;
            S.beta[c] = S.rho_new[c];
        }
        if (S.done[0] == 1 && S.done[1] == 1 && S.polish_left > 0 && a.cycle > 0 && !stop) {   // refinement cycle, see krylov_kernels.cuh
            S.polish_left -= 1;
            for (int c = 0; c < 2; ++c) {
                S.tol_eff[c] = 0.1 * S.norm_r[c];
                S.done[c] = fabs(S.rho_new[c]) < eps ? 2 : 0;
            }
        }
        if ((S.done[0] == 1 && S.done[1] == 1) || S.done[0] == 3 || S.done[1] == 3 || stop) S.final_ = 1;
    } else if (PHASE == KP_A) {
        S.applications += 1;
        for (int c = 0; c < 2; ++c) {
            if (S.done[c]) continue;
            if (fabs(acc[c]) < eps) { S.done[c] = 2; S.alpha[c] = 0.0; }
            else S.alpha[c] = S.rho_new[c] / acc[c];
        }
    } else if (PHASE == KP_B) {
        S.applications += 1;
        for (int c = 0; c < 2; ++c) {
            if (S.done[c]) continue;
            S.iters[c] += 1;
            S.norm_r[c] = sqrt(acc[c]);
            if (S.norm_r[c] <= S.tol_eff[c]) { S.done[c] = 1; continue; }
            if (fabs(acc[4 + c]) < eps) { S.done[c] = 2; S.omega[c] = 0.0; }
            else {
                S.omega[c] = acc[2 + c] / acc[4 + c];
                if (fabs(S.omega[c]) < eps) { S.done[c] = 2; S.omega[c] = 0.0; }
            }
        }
    } else if (PHASE == KP_C) {
        for (int c = 0; c < 2; ++c) {
            if (S.done[c]) continue;
            S.norm_r[c] = sqrt(acc[c]);
            if (S.norm_r[c] <= S.tol_eff[c]) { S.done[c] = 1; continue; }
            S.rho_old[c] = S.rho_new[c];
            S.rho_new[c] = acc[2 + c];
            if (S.iters[c] >= a.max_iters) { S.done[c] = 3; continue; }
            if (fabs(S.rho_new[c]) < eps) { S.done[c] = 2; continue; }
            S.beta[c] = (S.rho_new[c] / S.rho_old[c]) * (S.alpha[c] / S.omega[c]);
        }
    }
}

// COARSE, after the finalisation of a phase: a CTA per component adds the contribution slots of every aggregate in the fixed order
// of the host's lists (c = P^T x) and forms e = G c -- R0: e_r (and e_v = 0), A: e_v, B: e_t; after C: e_r -= alpha e_v + omega e_t.
constexpr int KC_THREADS = 1024, KC_SPLIT = 4;   // the sum over the aggregates is split in KC_SPLIT parts of KC_THREADS / KC_SPLIT output rows each
template <int PHASE>
__global__ void __launch_bounds__(KC_THREADS) krylov_coarse_kernel(const KPArgs a) {
    extern __shared__ double2 sh_c[];                               // c (nc_max), then the partial products (KC_SPLIT x nc_max)
    const int comp = (int)blockIdx.x, tid = threadIdx.x;
    const KState& S = a.state[comp];
    if (S.final_) return;
    const KCoarse C = a.coarse[comp];
    const int nc = C.nc;
    if (nc == 0 || !a.coarse_ok[comp]) return;
    double2* const er = a.e_r + C.agg_base;
    double2* const ev = a.e_v + C.agg_base;
    double2* const et = a.e_t + C.agg_base;
    if (PHASE == KP_C) {
        const double ax = S.alpha[0], ay = S.alpha[1], ox = S.omega[0], oy = S.omega[1];
        for (int J = tid; J < nc; J += KC_THREADS) {
            const double2 r0 = er[J], v0 = ev[J], t0 = et[J];
            er[J] = make_double2((r0.x - ax * v0.x) - ox * t0.x, (r0.y - ay * v0.y) - oy * t0.y);
        }
        return;
    }
    if (PHASE != KP_R0 && S.done[0] && S.done[1]) return;
    for (int J = tid; J < nc; J += KC_THREADS) {
        double2 sum = make_double2(0.0, 0.0);
        const int qe = a.contrib_ptr[C.agg_base + J + 1];
#pragma unroll 4
        for (int q = a.contrib_ptr[C.agg_base + J]; q < qe; ++q) sum = sum + __ldcg(a.contrib + a.contrib_src[q]);
        sh_c[J] = sum;
    }
    __syncthreads();
    // e = G c with G stored transposed (coarse_assemble_kernel): thread (part, J') adds its quarter of the aggregates, coalesced over J'
    double2* const part = sh_c + a.nc_max;
    const double* GT = a.G + C.g_off;
    constexpr int ROWS = KC_THREADS / KC_SPLIT;
    const int g = tid / ROWS, l = tid - g * ROWS;
    const int per = (nc + KC_SPLIT - 1) / KC_SPLIT, j0 = g * per, j1 = min(nc, j0 + per);
    for (int base = 0; base < nc; base += ROWS) {
        const int Jp = base + l;
        if (Jp < nc) {
            double sx = 0.0, sy = 0.0;
#pragma unroll 8
            for (int J = j0; J < j1; ++J) {
                const double gv = __ldg(GT + (size_t)J * nc + Jp);
                const double2 c = sh_c[J];
                sx += gv * c.x; sy += gv * c.y;
            }
            part[(size_t)g * a.nc_max + Jp] = make_double2(sx, sy);
        }
    }
    __syncthreads();
    double2* const dst = PHASE == KP_R0 ? er : (PHASE == KP_A ? ev : et);
    for (int Jp = tid; Jp < nc; Jp += KC_THREADS) {
        double2 sum = part[Jp];
#pragma unroll
        for (int q = 1; q < KC_SPLIT; ++q) sum = sum + part[(size_t)q * a.nc_max + Jp];
        dst[Jp] = sum;
        if (PHASE == KP_R0) ev[Jp] = make_double2(0.0, 0.0);
    }
}

// how many components are still iterating / not final (one int each), for the host's poll
__global__ void krylov_count_kernel(const KState* __restrict__ state, int n_comp, int* __restrict__ out /* [0] iterating, [1] not final */) {
    int it = 0, nf = 0;
    for (int c = threadIdx.x; c < n_comp; c += blockDim.x) {
        const KState& S = state[c];
        if (!S.final_) { nf += 1; if (!(S.done[0] && S.done[1])) it += 1; }
    }
    it = __reduce_add_sync(0xffffffffu, it);
    nf = __reduce_add_sync(0xffffffffu, nf);
    __shared__ int sh[2][32];
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = it; sh[1][threadIdx.x >> 5] = nf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a0 = 0, a1 = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a0 += sh[0][w]; a1 += sh[1][w]; }
        out[0] = a0; out[1] = a1;
    }
}
__global__ void krylov_reset_kernel(KState* __restrict__ state, int n_comp) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_comp) return;
    KState z{};
    state[c] = z;
}

}  // namespace tmesh
