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
};

enum KPhase : int { KP_R0 = 0, KP_A = 1, KP_B = 2, KP_C = 3, KP_ADD = 4 };

// TILES = true: the interior warp tiles (grid = ceil(n_wtiles / 4)); false: the boundary chunks (grid = n_chunks).  Two kernels
// rather than one so that the register count of the tile path (the bandwidth path) is not set by the gather-heavy row path.
template <int PHASE, bool HAS_PQ, bool TILES>
__global__ void __launch_bounds__(KP_THREADS) krylov_phase_kernel(const KPArgs a) {
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
    double acc[K_NACC] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    auto mirror = [&](double2* f, int sb, int se, double2 val) {
        for (int k = sb; k < se; ++k) f[a.slaves[k].self] = val;
    };
    // boundary rows of a chunk: thread q of the CTA takes row q of the concatenation smoothed | junction | sliding
    auto for_bnd = [&](auto&& fs, auto&& fj, auto&& fl) {
        const int n_s = ch.s_end - ch.s_begin, n_j = ch.j_end - ch.j_begin, n_l = ch.l_end - ch.l_begin;
        const int q = tid;
        if (q < n_s) fs(a.srows[ch.s_begin + q]);
        else if (q < n_s + n_j) fj(a.jrows[ch.j_begin + q - n_s]);
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
            k_interior_march<MODE_RESID, HAS_PQ>(t, a.blocks[t.block], xval, a.xc, a.pq, [&](int64_t k, double2 res, double2) { init(k, res, 0, 0); });
        } else {
            for_bnd([&](const SmoothedRow& row) {
                        double b2x = 0.0, b2y = 0.0;
                        const KRow o = k_smoothed<MODE_RESID, HAS_PQ>(row, xval, a.xc, a.pq, b2x, b2y);
                        init(row.g0, o.res, row.slave_begin, row.slave_end);
                        if (a.cycle == 0) { acc[2] += b2x; acc[3] += b2y; }
                    },
                    [&](const JunctionRow& row) { init(row.self, k_junction<MODE_RESID>(row, xval).res, row.slave_begin, row.slave_end); },
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
            const double2 rr = a.r[k], vv = a.v_old[k], pp = a.p_old[k];
            return make_double2(dx ? 0.0 : rr.x + bx * (pp.x - ox * vv.x), dy ? 0.0 : rr.y + by * (pp.y - oy * vv.y));
        };
        auto put = [&](int64_t k, const KRow& o, int sb, int se) {
            a.p_new[k] = o.centre; a.v_new[k] = o.res;
            mirror(a.p_new, sb, se, o.centre); mirror(a.v_new, sb, se, o.res);
            const double2 h = a.rhat[k];
            acc[0] += h.x * o.res.x; acc[1] += h.y * o.res.y;
        };
        if (is_tile) {
            k_interior_march<MODE_APPLY, HAS_PQ>(t, a.blocks[t.block], pval, a.xc, a.pq, [&](int64_t k, double2 res, double2 c) { put(k, KRow{res, c}, 0, 0); });
        } else {
            double u0, u1;
            for_bnd([&](const SmoothedRow& row) { put(row.g0, k_smoothed<MODE_APPLY, HAS_PQ>(row, pval, a.xc, a.pq, u0, u1), row.slave_begin, row.slave_end); },
                    [&](const JunctionRow& row) { put(row.self, k_junction<MODE_APPLY>(row, pval), row.slave_begin, row.slave_end); },
                    [&](const SlidingRow& row) { put(row.self, k_sliding<MODE_APPLY>(row, pval), row.slave_begin, row.slave_end); });
        }
    } else if (PHASE == KP_B) {
        // p_old / v_old are the CURRENT p and v here (the host hands the buffers phase A wrote)
        const bool mx = d0 != 0, my = d1 != 0;
        const double ax = S.alpha[0], ay = S.alpha[1];
        auto sval = [&](int64_t k) {
            const double2 rr = a.r[k], vv = a.v_old[k];
            return make_double2(mx ? 0.0 : rr.x - ax * vv.x, my ? 0.0 : rr.y - ay * vv.y);
        };
        auto put = [&](int64_t k, const KRow& o) {
            const double2 pp = a.p_old[k];
            double2 dd = a.d[k];
            if (!mx) dd.x += ax * pp.x;
            if (!my) dd.y += ay * pp.y;
            a.s[k] = o.centre; a.t[k] = o.res; a.d[k] = dd;
            acc[0] += o.centre.x * o.centre.x; acc[1] += o.centre.y * o.centre.y;
            acc[2] += o.centre.x * o.res.x; acc[3] += o.centre.y * o.res.y;
            acc[4] += o.res.x * o.res.x; acc[5] += o.res.y * o.res.y;
        };
        if (is_tile) {
            k_interior_march<MODE_APPLY, HAS_PQ>(t, a.blocks[t.block], sval, a.xc, a.pq, [&](int64_t k, double2 res, double2 c) { put(k, KRow{res, c}); });
        } else {
            double u0, u1;
            for_bnd([&](const SmoothedRow& row) { put(row.g0, k_smoothed<MODE_APPLY, HAS_PQ>(row, sval, a.xc, a.pq, u0, u1)); },
                    [&](const JunctionRow& row) { put(row.self, k_junction<MODE_APPLY>(row, sval)); },
                    [&](const SlidingRow& row) { put(row.self, k_sliding<MODE_APPLY>(row, sval)); });
        }
    } else if (PHASE == KP_C) {
        const bool ex = d0 != 0, ey = d1 != 0;
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
            if (!ex) { acc[0] += o.r.x * o.r.x; acc[2] += o.h.x * o.r.x; }
            if (!ey) { acc[1] += o.r.y * o.r.y; acc[3] += o.h.y * o.r.y; }
        };
        auto upd = [&](int64_t k, int sb, int se) { const RD o = load(k); store(k, o); mirror(a.r, sb, se, o.r); };
        if (is_tile) {
            k_interior_nodes_march(t, a.blocks[t.block], load, store);
        } else {
            for_bnd([&](const SmoothedRow& row) { upd(row.g0, row.slave_begin, row.slave_end); },
                    [&](const JunctionRow& row) { upd(row.self, row.slave_begin, row.slave_end); },
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
            for_bnd([&](const SmoothedRow& row) { upd(row.g0, row.slave_begin, row.slave_end); },
                    [&](const JunctionRow& row) { upd(row.self, row.slave_begin, row.slave_end); },
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
