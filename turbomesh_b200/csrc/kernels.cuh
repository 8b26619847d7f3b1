// kernels.cuh -- sm_100a device code for turbomesh's hot path (fp64, HBM-bound, no tensor cores).
//
//   tfi_kernel              tfi.linear2dBoundaryBlendedControlFunction          src/core/tfi.zig:112-208
//   winslow_interior_kernel StencilData.init + fillBlockInternalPointData       src/core/smoothing/smooth.zig:192-215, 923-992
//                           fused with the solver's use of the row (relaxation sweep / operator apply / residual),
//                           so the 9 coefficients live only in registers (matrix-free)
//   boundary_rows           fillBlockConnectionData (interface rows), junction rows, sliding rows, connected copies
//                           (device function: its CTAs ride in the launch of the interior kernels)
//                                                                               smooth.zig:994-1105, 813-859, 1115-1165
//   white_*                 White.initControlFunction / White.update            wall_control_function.zig:70-473
//   vector kernels          BiCGStab.zig:279-370 with x and y advanced in lock-step (double2 per node)
//
// Data layout: all blocks concatenated in the reference's global row order, AoS double2 (x,y) per node,
// j fastest inside a block (types.zig:78-101).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "topology.hpp"

namespace tmesh {

struct DevBlock {
    int64_t off;
    int32_t ni, nj;
    // coarse multigrid levels only (HAS_RHS kernels): sides that slide as a whole (bit 0..3: i = 0, i = ni-1, j = 0, j = nj-1)
    // and the weight of the tangential second difference in the rows next to them (see mgb_restrict_kernel)
    int32_t slide = 0, _pad = 0;
    double tan_i = 1.0, tan_j = 1.0;
};
struct Tile {
    int32_t block, i0, j0, rows;  // rows marched by the CTA starting at interior row i0
};

constexpr int TILE_J = 128;  // threads per CTA = nodes along j per tile
constexpr int TILE_I = 32;   // default rows marched per CTA (the host balances waves, see build_tiles)

enum Mode : int { MODE_RELAX = 0, MODE_APPLY = 1, MODE_RESID = 2, MODE_REL = 3 };

// ---------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ld2(const double2* p) { return *p; }
__device__ __forceinline__ double2 ldg2(const double2* p) { return __ldg(p); }
__device__ __forceinline__ double2 operator+(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 operator-(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-level reduction of K sums + 1 max; thread 0 writes them to out[0..K] (K sums then the max).
template <int K, int NT>
__device__ __forceinline__ void block_reduce_store(double (&sums)[K], double mx, double* out) {
    __shared__ double sh[(K + 1) * (NT / 32)];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) sums[k] = warp_sum(sums[k]);
    mx = warp_max(mx);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) sh[k * (NT / 32) + w] = sums[k];
        sh[K * (NT / 32) + w] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s = 0.0;
            for (int q = 0; q < NT / 32; ++q) s += sh[k * (NT / 32) + q];
            out[k] = s;
        }
        double m = 0.0;
        for (int q = 0; q < NT / 32; ++q) m = fmax(m, sh[K * (NT / 32) + q]);
        out[K] = m;
    }
}

// ---------------------------------------------------------------------------------------------------
// TFI.  Bit-exact with the reference: explicit round-to-nearest intrinsics are never contracted into
// FMAs, and the association follows tfi.zig:185-197 (scale, add, addAll left-to-right from (0,0)).
// One thread owns one j column and marches ROWS rows; edge data of the column stays in registers,
// edge data of the row is a warp-uniform (broadcast) read-only load.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double tfi_component(double u, double v, double omu, double omv, double uv, double u_omv, double omu_v, double omu_omv,
                                                double x0j, double xnj, double xi0, double xim, double x00, double xn0, double x0m, double xnm) {
    const double u_ij = __dadd_rn(__dmul_rn(omu, x0j), __dmul_rn(u, xnj));
    const double v_ij = __dadd_rn(__dmul_rn(omv, xi0), __dmul_rn(v, xim));
    double acc = __dadd_rn(0.0, __dmul_rn(uv, xnm));
    acc = __dadd_rn(acc, __dmul_rn(u_omv, xn0));
    acc = __dadd_rn(acc, __dmul_rn(omu_v, x0m));
    acc = __dadd_rn(acc, __dmul_rn(omu_omv, x00));
    return __dsub_rn(__dadd_rn(u_ij, v_ij), acc);
}

constexpr int TFI_ROWS = 16;
__global__ void __launch_bounds__(TILE_J) tfi_kernel(int ni, int nj, const double2* __restrict__ x_i_min, const double2* __restrict__ x_i_max,
                                                     const double2* __restrict__ x_j_min, const double2* __restrict__ x_j_max,
                                                     const double* __restrict__ s1, const double* __restrict__ s2, const double* __restrict__ t1,
                                                     const double* __restrict__ t2, double2* __restrict__ out) {
    const int j = blockIdx.x * TILE_J + threadIdx.x;
    const int i_begin = blockIdx.y * TFI_ROWS;
    if (j >= nj) return;
    const double t1_j = __ldg(t1 + j), t2_j = __ldg(t2 + j);
    const double2 x0j = ldg2(x_j_min + j), xnj = ldg2(x_j_max + j);
    const double2 x00 = ldg2(x_i_min), xn0 = ldg2(x_i_min + (ni - 1)), x0m = ldg2(x_j_min + (nj - 1)), xnm = ldg2(x_i_max + (ni - 1));
    const double omt1 = __dsub_rn(1.0, t1_j), dt = __dsub_rn(t2_j, t1_j);
    const int i_end = min(i_begin + TFI_ROWS, ni);
    for (int i = i_begin; i < i_end; ++i) {
        const double s1_i = __ldg(s1 + i), s2_i = __ldg(s2 + i);
        const double2 xi0 = ldg2(x_i_min + i), xim = ldg2(x_i_max + i);
        const double ds = __dsub_rn(s2_i, s1_i);
        // tfi.zig:185-186 (the two denominators are the same product with the factors swapped; IEEE multiplication commutes)
        // When the two clusterings of a direction coincide (s1 == s2 or t1 == t2 -- every synthetic workload, most O4H
        // blocks) the product is exactly 0, the denominator exactly 1 and x / 1 == x bit for bit: the two IEEE divisions,
        // which otherwise make this kernel fp64-bound instead of store-bound, are skipped.  ds is warp-uniform.
        const double prod = __dmul_rn(ds, dt);
        double u = __dadd_rn(__dmul_rn(omt1, s1_i), __dmul_rn(t1_j, s2_i));
        double v = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s1_i), t1_j), __dmul_rn(s1_i, t2_j));
        if (prod != 0.0) {
            const double den = __dsub_rn(1.0, prod);
            u = __ddiv_rn(u, den);
            v = __ddiv_rn(v, den);
        }
        const double omu = __dsub_rn(1.0, u), omv = __dsub_rn(1.0, v);
        const double uv = __dmul_rn(u, v), u_omv = __dmul_rn(u, omv), omu_v = __dmul_rn(omu, v), omu_omv = __dmul_rn(omu, omv);
        double2 r;
        r.x = tfi_component(u, v, omu, omv, uv, u_omv, omu_v, omu_omv, x0j.x, xnj.x, xi0.x, xim.x, x00.x, xn0.x, x0m.x, xnm.x);
        r.y = tfi_component(u, v, omu, omv, uv, u_omv, omu_v, omu_omv, x0j.y, xnj.y, xi0.y, xim.y, x00.y, xn0.y, x0m.y, xnm.y);
        out[(size_t)i * nj + j] = r;
    }
}

// ---------------------------------------------------------------------------------------------------
// The Winslow row (StencilData.init, smooth.zig:192-215) in "difference form".
//   W,E = nodes (i-1,j),(i+1,j); metric terms use central differences of the LAGGED coordinates:
//     x_xi = (E-W)/2, x_eta = (N-S)/2; g11 = |x_xi|^2, g22 = |x_eta|^2, g12 = x_xi.x_eta
//   row:  g22[(1+P/2)E + (1-P/2)W] + g11[(1+Q/2)N + (1-Q/2)S] - (g12/2)[(NE-SE)-(NW-SW)] - 2(g11+g22) C
//   The off-diagonal coefficients sum to the diagonal 2(g11+g22), so the row only sees differences to C:
//     rel = g22[(E-C)+(W-C) + P/2 (E-W)] + g11[R_i + Q/2 D_i] - (g12/2)(D_{i+1} - D_{i-1})
//   with D_r = u[r][j+1]-u[r][j-1] and R_r = (u[r][j+1]-u[r][j]) + (u[r][j-1]-u[r][j]).  Evaluating the row this
//   way is translation invariant: its rounding error scales with the cell size, not with |x|.
// ---------------------------------------------------------------------------------------------------
struct Metric {  // 4x the reference's g11, g22, g12: built from undivided central differences (the common factor
    double g11, g22, g12;  // cancels in every row result, and scaling by 4 is exact in binary floating point)
};
__device__ __forceinline__ Metric metric_terms(double2 W, double2 E, double2 Deta /* N - S */) {
    const double ax = E.x - W.x, ay = E.y - W.y;  // 2 x_xi, 2 y_xi
    Metric m;
    m.g22 = Deta.x * Deta.x + Deta.y * Deta.y;
    m.g12 = ax * Deta.x + ay * Deta.y;
    m.g11 = ax * ax + ay * ay;
    return m;
}
// rel = sum_k a_k (u_k - C) over the 8 neighbours (times the common factor 4)
template <bool HAS_PQ>
__device__ __forceinline__ double2 row_rel(const Metric& m, double P, double Q, double2 C, double2 W, double2 E, double2 Rj, double2 Deta, double2 Dp, double2 Dm) {
    double ex = (E.x - C.x) + (W.x - C.x), ey = (E.y - C.y) + (W.y - C.y);
    double nx = Rj.x, ny = Rj.y;
    if (HAS_PQ) {
        ex += 0.5 * P * (E.x - W.x); ey += 0.5 * P * (E.y - W.y);
        nx += 0.5 * Q * Deta.x; ny += 0.5 * Q * Deta.y;
    }
    const double h = 0.5 * m.g12;
    double2 r;
    r.x = m.g22 * ex + m.g11 * nx - h * (Dp.x - Dm.x);
    r.y = m.g22 * ey + m.g11 * ny - h * (Dp.y - Dm.y);
    return r;
}

// What a row produces, shared by interior and interface rows.  The Krylov modes work on the row-scaled system
// D^-1 A x = D^-1 b (D = diagonal, the reference's `diagonal` preconditioner, GMRES.zig:176-196, applied from the
// left as GMRES.zig:300-423 does): residuals are then "Jacobi updates", i.e. lengths, and the stopping test is
// meaningful at any geometric scale.  With a_ii = -2(g11+g22):
//   RELAX: C + w * rel / (2(g11+g22))                 (damped Jacobi update of the row)
//   APPLY: (A v)_i / a_ii     = -rel / (2(g11+g22))   (homogeneous)
//   RESID: (b - A x)_i / a_ii = +rel / (2(g11+g22))   (rhs of these rows is 0 once periodic shifts are folded in)
template <int MODE>
__device__ __forceinline__ double2 row_result(const Metric& m, double2 rel, double2 C, double omega) {
    const double diag = 2.0 * (m.g11 + m.g22);
    const double inv = diag == 0.0 ? 1.0 : 1.0 / diag;  // zero diagonal -> 1.0 as in GMRES.zig:190-194
    if (MODE == MODE_RELAX) {
        return make_double2(C.x + omega * (rel.x * inv), C.y + omega * (rel.y * inv));
    } else if (MODE == MODE_APPLY) {
        return make_double2(-(rel.x * inv), -(rel.y * inv));
    } else if (MODE == MODE_RESID) {
        return make_double2(rel.x * inv, rel.y * inv);
    } else {
        return rel;  // MODE_REL: the unscaled row (minus the multigrid right-hand side), see mg_* kernels
    }
}

// ---------------------------------------------------------------------------------------------------
// Interior rows of all blocks.  One CTA = one tile (TILE_J columns x TILE_I rows) of one block; a thread owns
// one j column and marches along i with a 3-row register window, so every node is loaded once per sweep
// from HBM/L2 (its j-1/j+1 neighbours are L1 hits of the same 128-byte lines).  Stores are coalesced 16 B.
//   u      field the row is applied to            xc   lagged coordinates (LAGGED; otherwise xc == u)
//   pq     control function P,Q (HAS_PQ)          out  result field
//   stats  per-CTA partials: sum dx^2, sum dy^2, (dot slots), max|d|   (only when STATS)
// STATS for RELAX: d = out - u.  For APPLY: partial dots with `dotv` (rhat.v, or t.s and t.t) -- see K.
// ---------------------------------------------------------------------------------------------------
constexpr int BND_THREADS = 128;
template <int MODE, bool LAGGED, bool HAS_PQ, int STATS, bool HAS_RHS = false>
__device__ __forceinline__ void boundary_rows(int cta, const SmoothedRow* __restrict__ srows, int n_s, const JunctionRow* __restrict__ jrows, int n_j,
                                              const SlidingRow* __restrict__ lrows, int n_l, const SlaveRow* __restrict__ slaves,
                                              const double2* __restrict__ u, const double2* __restrict__ xc, const double2* __restrict__ pq,
                                              double2* __restrict__ out, double omega, const double2* __restrict__ dot_a, double* __restrict__ partials,
                                              const double2* __restrict__ rhs = nullptr, int row_override = -1);
// The boundary rows ride in the same launch as the interior tiles (the first n_ctas CTAs of the grid): they are few
// and latency-bound, so they hide behind the interior work instead of costing a launch of their own.
struct BndArgs {
    const SmoothedRow* srows;
    const JunctionRow* jrows;
    const SlidingRow* lrows;
    const SlaveRow* slaves;
    double* partials;
    int n_s, n_j, n_l, n_ctas;
};

template <int MODE, bool LAGGED, bool HAS_PQ, int STATS>
__global__ void __launch_bounds__(TILE_J) winslow_interior_kernel(const Tile* __restrict__ tiles, const DevBlock* __restrict__ blocks,
                                                                   const double2* __restrict__ u, const double2* __restrict__ xc,
                                                                   const double2* __restrict__ pq, double2* __restrict__ out, double omega,
                                                                   const double2* __restrict__ dot_a, double* __restrict__ partials, const BndArgs bnd) {
    if ((int)blockIdx.x < bnd.n_ctas) {  // the boundary rows ride in the same launch (first n_ctas CTAs), as in the bulk kernel
        boundary_rows<MODE, LAGGED, HAS_PQ, STATS>(blockIdx.x, bnd.srows, bnd.n_s, bnd.jrows, bnd.n_j, bnd.lrows, bnd.n_l, bnd.slaves, u, xc, pq, out, omega, dot_a,
                                                   bnd.partials);
        return;
    }
    const int tile_id = (int)blockIdx.x - bnd.n_ctas;
    const Tile t = tiles[tile_id];
    const DevBlock b = blocks[t.block];
    const int nj = b.nj;
    const int j = t.j0 + threadIdx.x;
    const bool active = j <= nj - 2;
    const int jc = active ? j : nj - 2;  // clamp: inactive lanes recompute the last column, never store
    const int i_begin = t.i0, i_end = min(t.i0 + t.rows, b.ni - 1);
    const double2* ub = u + b.off;
    const double2* cb = LAGGED ? xc + b.off : ub;
    double2* ob = out + b.off;

    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, mx = 0.0;

    // register window: rows i-1 (m), i (0), i+1 (p)
    size_t idx = (size_t)(i_begin - 1) * nj + jc;
    double2 Cm = ld2(ub + idx), Dm = ld2(ub + idx + 1) - ld2(ub + idx - 1);
    idx += nj;
    double2 l = ld2(ub + idx - 1), r = ld2(ub + idx + 1);
    double2 C0 = ld2(ub + idx), D0 = r - l, R0 = (r - C0) + (l - C0);
    double2 cCm, cC0, cD0;
    if (LAGGED) {
        cCm = ld2(cb + idx - nj);
        cC0 = ld2(cb + idx);
        cD0 = ld2(cb + idx + 1) - ld2(cb + idx - 1);
    }
#pragma unroll 4
    for (int i = i_begin; i < i_end; ++i) {
        const size_t ip = idx + nj;  // row i+1
        const double2 lp = ld2(ub + ip - 1), rp = ld2(ub + ip + 1);
        const double2 Cp = ld2(ub + ip), Dp = rp - lp, Rp = (rp - Cp) + (lp - Cp);
        double2 cCp, cDp;
        Metric m;
        if (LAGGED) {
            cCp = ld2(cb + ip);
            cDp = ld2(cb + ip + 1) - ld2(cb + ip - 1);
            m = metric_terms(cCm, cCp, cD0);
        } else {
            m = metric_terms(Cm, Cp, D0);
        }
        double P = 0.0, Q = 0.0;
        if (HAS_PQ) {
            const double2 f = ldg2(pq + b.off + idx);
            P = f.x; Q = f.y;
        }
        const double2 rel = row_rel<HAS_PQ>(m, P, Q, C0, Cm, Cp, R0, D0, Dp, Dm);
        const double2 res = row_result<MODE>(m, rel, C0, omega);
        if (active) {
            ob[idx] = res;
            if (STATS == 1) {  // update norms (smooth.zig:112-134) of a relaxation sweep
                const double dx = res.x - C0.x, dy = res.y - C0.y;
                s0 += dx * dx; s1 += dy * dy;
                mx = fmax(mx, fmax(fabs(dx), fabs(dy)));
            } else if (STATS == 2) {  // dot(a, res) per component
                const double2 a = ld2(dot_a + b.off + idx);
                s0 += a.x * res.x; s1 += a.y * res.y;
            } else if (STATS == 3) {  // dot(res, a) and dot(res, res) per component  (t.s, t.t)
                const double2 a = ld2(dot_a + b.off + idx);
                s0 += a.x * res.x; s1 += a.y * res.y;
                s2 += res.x * res.x; s3 += res.y * res.y;
            } else if (STATS == 4) {  // sum of squares of res (||r||^2)
                s0 += res.x * res.x; s1 += res.y * res.y;
            }
        }
        Cm = C0; Dm = D0;
        C0 = Cp; D0 = Dp; R0 = Rp;
        if (LAGGED) { cCm = cC0; cC0 = cCp; cD0 = cDp; }
        idx = ip;
    }
    if (STATS != 0) {
        double sums[4] = {s0, s1, s2, s3};
        block_reduce_store<4, TILE_J>(sums, mx, partials + (size_t)tile_id * 5);
    }
}

// ---------------------------------------------------------------------------------------------------
// Same rows, fed by the async copy engine (the throughput path: coefficients from the field itself).
// Thread 0 streams row segments of TILE_J+2 nodes into a shared-memory ring with 1-D bulk async copies
// (cp.async.bulk, SASS UBLKCP) that complete on an mbarrier; the ring keeps BULK_NS*BULK_R rows in flight per CTA
// without holding registers, which is what a latency-bound fp64 stencil needs to approach the HBM roofline.
// Consumers read a node and its j-1/j+1 neighbours from shared memory (3 x LDS.128), keep the 3-row window in
// registers and store results straight from registers (coalesced 16 B per thread).
// ---------------------------------------------------------------------------------------------------
constexpr int BULK_R = 4;    // rows per pipeline stage
constexpr int BULK_NS = 3;   // stages
constexpr int BULK_ROW = TILE_J + 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int MODE, bool HAS_PQ, int STATS, bool HAS_RHS = false>
__global__ void __launch_bounds__(TILE_J) winslow_interior_bulk_kernel(const Tile* __restrict__ tiles, const DevBlock* __restrict__ blocks,
                                                                        const double2* __restrict__ u, const double2* __restrict__ pq,
                                                                        double2* __restrict__ out, double omega, const double2* __restrict__ dot_a,
                                                                        double* __restrict__ partials, const BndArgs bnd,
                                                                        const double2* __restrict__ rhs /* HAS_RHS: multigrid tau term, in row units */) {
    __shared__ __align__(128) double2 ring[BULK_NS][BULK_R][BULK_ROW];
    __shared__ __align__(8) uint64_t full[BULK_NS];
    if ((int)blockIdx.x < bnd.n_ctas) {
        boundary_rows<MODE, false, HAS_PQ, STATS, HAS_RHS>(blockIdx.x, bnd.srows, bnd.n_s, bnd.jrows, bnd.n_j, bnd.lrows, bnd.n_l, bnd.slaves, u, u, pq, out, omega,
                                                           dot_a, bnd.partials, rhs);
        return;
    }
    const int tile_id = (int)blockIdx.x - bnd.n_ctas;
    const Tile t = tiles[tile_id];
    const DevBlock b = blocks[t.block];
    const int nj = b.nj;
    const int tid = threadIdx.x;
    const int j = t.j0 + tid;
    const bool active = j <= nj - 2;
    const int i_begin = t.i0, i_end = min(t.i0 + t.rows, b.ni - 1);
    const int width = min(TILE_J, nj - 1 - t.j0) + 2;          // nodes per row segment incl. the two halo columns
    const uint32_t row_bytes = (uint32_t)width * (uint32_t)sizeof(double2);
    const int n_rows = (i_end - i_begin) + 2;                   // rows i_begin-1 .. i_end
    const int n_chunks = (n_rows + BULK_R - 1) / BULK_R;
    const double2* ub = u + b.off;
    double2* ob = out + b.off;
    const double2* src0 = ub + (size_t)(i_begin - 1) * nj + (t.j0 - 1);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < BULK_NS; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue_chunk = [&](int c) {  // thread 0 only
        const int s = c % BULK_NS;
        const int r0 = c * BULK_R, r1 = min(r0 + BULK_R, n_rows);
        mbar_expect_tx(&full[s], (uint32_t)(r1 - r0) * row_bytes);
        for (int r = r0; r < r1; ++r) bulk_load(&ring[s][r - r0][0], src0 + (size_t)r * nj, row_bytes, &full[s]);
    };
    if (tid == 0) {
        for (int c = 0; c < BULK_NS && c < n_chunks; ++c) issue_chunk(c);
    }

    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, mx = 0.0;
    double2 Cm = make_double2(0, 0), Dm = Cm, C0 = Cm, D0 = Cm, R0 = Cm;
    const int tl = active ? tid : 0;  // inactive lanes read a valid column, never store
    long long idx = (long long)(i_begin - 2) * nj + (active ? j : t.j0);  // block-local index of the row being finished (k-2)
    int stage = 0;
    uint32_t parity = 0;
    for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(&full[stage], parity);
#pragma unroll
        for (int slot = 0; slot < BULK_R; ++slot) {
            const int k = c * BULK_R + slot;
            if (k < n_rows) {  // uniform over the CTA
                const double2 lp = ring[stage][slot][tl], Cp = ring[stage][slot][tl + 1], rp = ring[stage][slot][tl + 2];
                const double2 Dp = rp - lp, Rp = (rp - Cp) + (lp - Cp);
                if (k >= 2) {
                    Metric m = metric_terms(Cm, Cp, D0);
                    if (HAS_RHS && b.slide) {  // coarse level, row next to a sliding side: Galerkin weight of the tangential term
                        const int irow = i_begin - 2 + k;
                        if ((irow == 1 && (b.slide & 1)) || (irow == b.ni - 2 && (b.slide & 2))) m.g11 *= b.tan_i;
                        if ((j == 1 && (b.slide & 4)) || (j == nj - 2 && (b.slide & 8))) m.g22 *= b.tan_j;
                    }
                    double P = 0.0, Q = 0.0;
                    if (HAS_PQ) {
                        const double2 f = ldg2(pq + b.off + idx);
                        P = f.x; Q = f.y;
                    }
                    double2 rel = row_rel<HAS_PQ>(m, P, Q, C0, Cm, Cp, R0, D0, Dp, Dm);
                    if (HAS_RHS) {
                        const double2 f = ld2(rhs + b.off + idx);
                        rel.x -= f.x; rel.y -= f.y;
                    }
                    const double2 res = row_result<MODE>(m, rel, C0, omega);
                    if (active) {
                        ob[idx] = res;
                        if (STATS == 1) {
                            const double dx = res.x - C0.x, dy = res.y - C0.y;
                            s0 += dx * dx; s1 += dy * dy;
                            mx = fmax(mx, fmax(fabs(dx), fabs(dy)));
                        } else if (STATS == 2) {
                            const double2 a = ld2(dot_a + b.off + idx);
                            s0 += a.x * res.x; s1 += a.y * res.y;
                        } else if (STATS == 3) {
                            const double2 a = ld2(dot_a + b.off + idx);
                            s0 += a.x * res.x; s1 += a.y * res.y;
                            s2 += res.x * res.x; s3 += res.y * res.y;
                        } else if (STATS == 4) {
                            s0 += res.x * res.x; s1 += res.y * res.y;
                        }
                    }
                }
                Cm = C0; Dm = D0;
                C0 = Cp; D0 = Dp; R0 = Rp;
                idx += nj;
            }
        }
        __syncthreads();  // every thread has copied this stage into registers: the stage may be refilled
        if (tid == 0 && c + BULK_NS < n_chunks) issue_chunk(c + BULK_NS);
        if (++stage == BULK_NS) { stage = 0; parity ^= 1u; }
    }
    if (STATS != 0) {
        double sums[4] = {s0, s1, s2, s3};
        block_reduce_store<4, TILE_J>(sums, mx, partials + (size_t)tile_id * 5);
    }
}

// ---------------------------------------------------------------------------------------------------
// Boundary rows: one thread per free boundary row (smoothed interface rows, then junction rows, then sliding
// rows).  RELAX additionally writes the row's `connected` copies (x_slave = x_root + shift).
// ---------------------------------------------------------------------------------------------------
// HAS_RHS (coarse multigrid levels): rhs[self] is the FAS tau term of the row -- in row units for the Winslow
// interface rows, in update units (lengths) for junction and sliding rows.
template <int MODE, bool LAGGED, bool HAS_PQ, int STATS, bool HAS_RHS>
__device__ __forceinline__ void boundary_rows(int cta, const SmoothedRow* __restrict__ srows, int n_s, const JunctionRow* __restrict__ jrows, int n_j,
                                              const SlidingRow* __restrict__ lrows, int n_l, const SlaveRow* __restrict__ slaves,
                                              const double2* __restrict__ u, const double2* __restrict__ xc, const double2* __restrict__ pq,
                                              double2* __restrict__ out, double omega, const double2* __restrict__ dot_a, double* __restrict__ partials,
                                              const double2* __restrict__ rhs, int row_override /* >= 0: this row (STATS == 0 callers only) */) {
    const int r = row_override >= 0 ? row_override : cta * BND_THREADS + threadIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, mx = 0.0;
    double2 res = make_double2(0.0, 0.0), old = make_double2(0.0, 0.0);
    int64_t self = -1;
    int sb = 0, se = 0, n_copies = 0;
    if (r < n_s) {
        const SmoothedRow row = srows[r];
        self = row.g0;
        sb = row.slave_begin; se = row.slave_end; n_copies = row.n_copies;
        const double2 per = make_double2(row.px, row.py);
        // values the row is applied to; block-1 columns are shifted by -periodicity in the affine modes
        // (equivalent to the reference's rhs = p * (a(i-1,j+1)+a(i,j+1)+a(i+1,j+1)), smooth.zig:1060-1061)
        const bool affine = (MODE == MODE_RELAX || MODE == MODE_RESID || MODE == MODE_REL);
        const double2 sh = affine ? per : make_double2(0.0, 0.0);
        const double2 C = ld2(u + row.g0);
        const double2 W = ld2(u + row.g0 - row.d0), E = ld2(u + row.g0 + row.d0);
        const double2 S = ld2(u + row.g0 + row.n0), SW = ld2(u + row.g0 - row.d0 + row.n0), SE = ld2(u + row.g0 + row.d0 + row.n0);
        const double2 N = ld2(u + row.iN) - sh, NW = ld2(u + row.iNW) - sh, NE = ld2(u + row.iNE) - sh;
        Metric m;
        if (LAGGED) {
            const double2 cW = ld2(xc + row.g0 - row.d0), cE = ld2(xc + row.g0 + row.d0), cS = ld2(xc + row.g0 + row.n0);
            const double2 cN = ld2(xc + row.iN) - per;  // smooth.zig:1032
            m = metric_terms(cW, cE, cN - cS);
        } else {
            m = metric_terms(W, E, (ld2(u + row.iN) - per) - S);
        }
        double P = 0.0, Q = 0.0;
        if (HAS_PQ) {
            const double2 f = ldg2(pq + row.g0);
            if (row.periodic) { P = f.x; Q = f.y; } else { P = f.y; Q = f.x; }  // smooth.zig:1040-1041 vs 1082-1083
        }
        double2 rel = row_rel<HAS_PQ>(m, P, Q, C, W, E, (N - C) + (S - C), N - S, NE - SE, NW - SW);
        if (HAS_RHS) { const double2 f = ld2(rhs + row.g0); rel.x -= f.x; rel.y -= f.y; }
        res = row_result<MODE>(m, rel, C, omega);
        old = C;
        if (STATS == 4 && row.periodic) {  // rhs of the reference's row: p * (a(i-1,j+1)+a(i,j+1)+a(i+1,j+1)) = p * g11 (1 + Q/2)
            const double a = 0.25 * m.g11 * (1.0 + 0.5 * Q);  // Metric holds 4 g11
            s2 = (row.px * a) * (row.px * a); s3 = (row.py * a) * (row.py * a);
        }
    } else if (r < n_s + n_j) {
        const JunctionRow row = jrows[r - n_s];
        self = row.self;
        sb = row.slave_begin; se = row.slave_end; n_copies = row.n_copies;
        const double2 C = ld2(u + row.self);
        double2 sum = make_double2(0.0, 0.0);  // sum_k (x_k - C): translation invariant like the Winslow rows
        for (int k = 0; k < row.n; ++k) sum = sum + (ld2(u + row.nbr[k]) - C);
        const double n = (double)row.n;
        double2 tau = make_double2(0.0, 0.0);
        if (HAS_RHS) tau = ld2(rhs + row.self);
        if (MODE == MODE_RELAX) {
            res = make_double2(C.x + omega * ((sum.x - row.rhs_x) / n - tau.x), C.y + omega * ((sum.y - row.rhs_y) / n - tau.y));
        } else if (MODE == MODE_APPLY) {  // row / a_ii with a_ii = -n
            res = make_double2(-(sum.x / n), -(sum.y / n));
        } else {
            res = make_double2((sum.x - row.rhs_x) / n - tau.x, (sum.y - row.rhs_y) / n - tau.y);
        }
        old = C;
    } else if (r < n_s + n_j + n_l) {
        const SlidingRow row = lrows[r - n_s - n_j];
        self = row.self;
        sb = row.slave_begin; se = row.slave_end; n_copies = row.n_copies;
        const double2 C = ld2(u + row.self), I = ld2(u + row.inner);
        const double ys = (double)row.ysign;
        double tau_y = 0.0;
        if (HAS_RHS) tau_y = ld2(rhs + row.self).y;
        if (MODE == MODE_RELAX) {
            res = make_double2(row.rhs_x, I.y + ys * row.rhs_y - tau_y);
        } else if (MODE == MODE_APPLY) {  // a_ii = 1 (x) and ysign (y)
            res = make_double2(C.x, C.y - I.y);
        } else if (MODE == MODE_REL) {    // multigrid residual in update units; x is a Dirichlet value (no residual)
            res = make_double2(0.0, ys * row.rhs_y - (C.y - I.y) - tau_y);
        } else {
            res = make_double2(row.rhs_x - C.x, ys * row.rhs_y - (C.y - I.y));
        }
        old = C;
    }
    if (self >= 0) {
        out[self] = res;
        if (MODE == MODE_RELAX) {
            for (int k = sb; k < se; ++k) {
                const SlaveRow sl = slaves[k];
                out[sl.self] = make_double2(res.x + sl.sx, res.y + sl.sy);
            }
        }
        if (STATS == 1) {
            // the reference sums over ALL nodes (smooth.zig:117-133): count the row and its copies
            const double dx = res.x - old.x, dy = res.y - old.y;
            const double w = 1.0 + (double)n_copies;
            s0 = w * dx * dx; s1 = w * dy * dy;
            mx = fmax(fabs(dx), fabs(dy));
        } else if (STATS == 2) {
            const double2 a = ld2(dot_a + self);
            s0 = a.x * res.x; s1 = a.y * res.y;
        } else if (STATS == 3) {
            const double2 a = ld2(dot_a + self);
            s0 = a.x * res.x; s1 = a.y * res.y; s2 = res.x * res.x; s3 = res.y * res.y;
        } else if (STATS == 4) {
            s0 = res.x * res.x; s1 = res.y * res.y;  // s2, s3 may already hold the periodic rhs term
        }
    }
    if (STATS != 0) {
        double sums[4] = {s0, s1, s2, s3};
        block_reduce_store<4, BND_THREADS>(sums, mx, partials + (size_t)cta * 5);
    }
}

// mode 0: v[slave] = v[root]; mode 1: v[slave] = v[root] + shift (keeps `connected` copies consistent,
// smooth.zig:804-812); mode 2: v[slave] = 0 (Krylov vectors carry zeros in eliminated rows)
__global__ void sync_slaves_kernel(const SlaveRow* __restrict__ slaves, int n, double2* __restrict__ v, int mode) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const SlaveRow s = slaves[k];
    if (mode == 2) { v[s.self] = make_double2(0.0, 0.0); return; }
    const double2 r = v[s.root];
    v[s.self] = mode == 1 ? make_double2(r.x + s.sx, r.y + s.sy) : r;
}

// halo exchange, send side: gathers the owned nodes that peers ghost into one contiguous buffer (per-peer segments)
__global__ void pack_kernel(const int64_t* __restrict__ idx, int64_t n, const double2* __restrict__ v, double2* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = v[idx[k]];
}

// ---------------------------------------------------------------------------------------------------
// Halo exchange over NVLink peer memory (one process per GPU, buffers mapped with CUDA IPC): the owner PUSHES the nodes
// its neighbours ghost straight into the tail of their fields -- gather and remote store in one kernel, no staging buffer,
// no NCCL launch -- and the last CTA to finish raises the rank's slot of every neighbour's flag array to the exchange
// counter (release at system scope, after every writer fenced its stores).  The consumer side is p2p_wait_kernel.
// ---------------------------------------------------------------------------------------------------
constexpr int P2P_MAX_RANKS = 16;
struct PushArgs {
    double2* dst[P2P_MAX_RANKS];                 // peer field + offset of this rank's ghost segment there
    int64_t base[P2P_MAX_RANKS + 1];             // send list offsets per peer
    unsigned long long* flag[P2P_MAX_RANKS];     // this rank's slot in the peer's flag array (NULL: not a neighbour)
    int n_ranks;
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// One launch per exchange, at most one CTA per SM (all CTAs are co-resident, so spinning cannot starve a CTA that still
// has to push): every CTA pushes its share (grid-stride) and fences; the last CTA to finish signals the neighbours; then
// ALL CTAs wait until the neighbours' pushes number `epoch` have landed here (bounded spin: a rank that never arrives
// makes the wait give up and raise *err instead of hanging the GPU) and re-derive their share of the copies whose root is
// a ghost (mode as in sync_slaves_kernel; n_slaves = 0 for fields without copies).
__global__ void __launch_bounds__(256) p2p_exchange_kernel(const int64_t* __restrict__ idx, int64_t n, double2* __restrict__ v, PushArgs a,
                                                           unsigned long long epoch, unsigned int* __restrict__ counter,
                                                           const unsigned long long* __restrict__ flags, unsigned int nb_mask, int* __restrict__ err,
                                                           const SlaveRow* __restrict__ slaves, int n_slaves, int slave_mode) {
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (int64_t)gridDim.x * 256) {
        int p = 0;
        while (k >= a.base[p + 1]) ++p;
        a.dst[p][k - a.base[p]] = v[idx[k]];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(counter, 1u);
        if (ticket == gridDim.x - 1) {
            *counter = 0u;
            __threadfence_system();
            for (int p = 0; p < a.n_ranks; ++p)
                if (a.flag[p]) st_release_sys(a.flag[p], epoch);
        }
    }
    const int p = threadIdx.x;
    if (p < P2P_MAX_RANKS && ((nb_mask >> p) & 1u)) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + p) < epoch) {
            if (clock64() - t0 > 60000000000ll) { *err = 1; break; }  // ~30 s: rank skew (a peer still building a hierarchy) is legitimate
            __nanosleep(64);
        }
    }
    __syncthreads();
    for (int q = blockIdx.x * 256 + threadIdx.x; q < n_slaves; q += gridDim.x * 256) {
        const SlaveRow s = slaves[q];
        if (slave_mode == 2) { v[s.self] = make_double2(0.0, 0.0); continue; }
        // the root is a ghost written by a peer GPU: read it past the (non-coherent) L1
        const double2 r = __ldcg(v + s.root);
        v[s.self] = slave_mode == 1 ? make_double2(r.x + s.sx, r.y + s.sy) : r;
    }
}

// begin_smoothing: capture rhs_x of sliding rows from the initial mesh (smooth.zig:853-857), apply the
// (normally empty) fixed overrides.
__global__ void capture_boundary_kernel(SlidingRow* __restrict__ lrows, int n_l, const FixedOverride* __restrict__ fo, int n_fo, double2* __restrict__ x) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_l) {
        if (lrows[k].rhs_x_from_initial) lrows[k].rhs_x = x[lrows[k].self].x;
    } else if (k < n_l + n_fo) {
        const FixedOverride f = fo[k - n_l];
        x[f.self] = make_double2(f.x, f.y);
    }
}

// connectionDataCheck (smooth.zig:220-275): max over all interface node pairs of |x0 + p - x1|_inf; also the
// index of the worst pair so the error message can name it.
__global__ void pair_check_kernel(const PairCheck* __restrict__ pairs, int n, const double2* __restrict__ x, double tol, unsigned long long* __restrict__ worst) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const PairCheck p = pairs[k];
    const double2 a = x[p.g0], b = x[p.g1];
    const double ex = fabs((a.x + p.px) - b.x), ey = fabs((a.y + p.py) - b.y);
    const double e = fmax(ex, ey);
    if (!(e <= tol)) {
        // pack (error as ordered bits of a non-negative double, pair index) -> atomicMax keeps the worst
        const unsigned long long bits = isnan(e) ? 0x7ff8000000000000ull : (unsigned long long)__double_as_longlong(e);
        atomicMax(worst, (bits & 0xffffffff00000000ull) | (unsigned long long)(unsigned)k);
    }
}

// ---------------------------------------------------------------------------------------------------
// White wall control function (wall_control_function.zig:70-473).  The reference hard-codes it to blocks 0 and 1
// (the two O-grid halves, wall = line j = 0) joined at the leading edge by connection 0; here every such pair of
// blocks is a *group* (default: the reference's single group), so that a batch of independent cuts is handled in the
// same launches.  wall_pq holds the accumulated (P,Q) of every wall node of a group:
// [block A: niA entries][block B: niB entries] starting at wall_base.
// ---------------------------------------------------------------------------------------------------
struct WhiteParams {
    int64_t off0, off1;           // local offsets of the group's two blocks
    int32_t ni0, nj0, ni1, nj1;
    int32_t c_in0, c_in1, c_al0;  // leading-edge connection: inward shifts on both sides, along shift on side 0
    int32_t wall_base;            // first entry of the group in wall_pq
};
struct WhiteNode {
    int32_t group, t;             // t in [0, ni0 + ni1): wall node of block A, then of block B
};

__device__ __forceinline__ void white_eq610(double x_xi, double y_xi, double x_xi2, double y_xi2, double x_eta, double y_eta, double x_eta2, double y_eta2,
                                            double& p, double& q) {
    const double g11 = x_xi * x_xi + y_xi * y_xi;
    const double g22 = x_eta * x_eta + y_eta * y_eta;
    p = -(x_xi * x_xi2 + y_xi * y_xi2) / g11 - (x_xi * x_eta2 + y_xi * y_eta2) / g22;
    q = -(x_eta * x_eta2 + y_eta * y_eta2) / g22 - (x_eta * x_xi2 + y_eta * y_xi2) / g11;
}
__device__ __forceinline__ void white_delta(double ds_target, double theta_target, double x_xi, double y_xi, double x_eta, double y_eta, double& p, double& q) {
    // White.computeUpdate, wall_control_function.zig:293-309
    const double g11 = x_xi * x_xi + y_xi * y_xi;
    const double g12 = x_xi * x_eta + y_xi * y_eta;
    const double g22 = x_eta * x_eta + y_eta * y_eta;
    const double ds = sqrt(g22);
    const double theta = acos(g12 / sqrt(g11 * g22));
    const double delta_p = -atan2(theta_target - theta, theta_target);
    const double delta_q = atan2(ds_target - ds, ds_target);
    p += 0.1 * delta_p;
    q += 0.1 * delta_q;
}

// one thread per wall node; `update` = 0: initControlFunction, 1: update
__global__ void white_wall_kernel(const WhiteParams* __restrict__ groups, const WhiteNode* __restrict__ nodes, int n_nodes, double ds_target,
                                  double theta_target, const double2* __restrict__ x, double2* __restrict__ wall_pq, int update) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes) return;
    const WhiteNode node = nodes[k];
    const WhiteParams w = groups[node.group];
    const int t = node.t;
    const int blk = t < w.ni0 ? 0 : 1;
    const int i = blk ? t - w.ni0 : t;
    const int ni = blk ? w.ni1 : w.ni0, nj = blk ? w.nj1 : w.nj0;
    const double2* d = x + (blk ? w.off1 : w.off0);
    const size_t l = (size_t)i * nj;
    const double2 c = d[l], e1 = d[l + 1];
    double x_xi, y_xi, x_xi2 = 0, y_xi2 = 0;
    if (i == 0) {  // forward
        const double2 a = d[l + nj];
        x_xi = -c.x + a.x; y_xi = -c.y + a.y;
        if (!update) { const double2 a2 = d[l + 2 * (size_t)nj]; x_xi2 = c.x - 2 * a.x + a2.x; y_xi2 = c.y - 2 * a.y + a2.y; }
    } else if (i == ni - 1) {  // backward
        const double2 a = d[l - nj];
        x_xi = c.x - a.x; y_xi = c.y - a.y;
        if (!update) { const double2 a2 = d[l - 2 * (size_t)nj]; x_xi2 = c.x - 2 * a.x + a2.x; y_xi2 = c.y - 2 * a.y + a2.y; }
    } else {  // central
        const double2 a = d[l + nj], b = d[l - nj];
        x_xi = 0.5 * (a.x - b.x); y_xi = 0.5 * (a.y - b.y);
        x_xi2 = a.x - 2 * c.x + b.x; y_xi2 = a.y - 2 * c.y + b.y;
    }
    const double x_eta = -c.x + e1.x, y_eta = -c.y + e1.y;
    double p, q;
    if (!update) {
        const double2 e2 = d[l + 2];
        white_eq610(x_xi, y_xi, x_xi2, y_xi2, x_eta, y_eta, c.x - 2 * e1.x + e2.x, c.y - 2 * e1.y + e2.y, p, q);
    } else {
        const double2 acc = wall_pq[w.wall_base + t];
        p = acc.x; q = acc.y;
        white_delta(ds_target, theta_target, x_xi, y_xi, x_eta, y_eta, p, q);
    }
    if (t == 0) {
        // the leading-edge connection (blockA:j_min <-> blockB:j_min) shares wall node 0; xi runs across the two blocks,
        // eta along the connection (wall_control_function.zig:203-279, 394-472)
        const double2* d1 = x + w.off1;
        const double2 ip1 = d[w.c_in0], im1 = d1[w.c_in1], jp1 = d[w.c_al0];
        if (!update) {
            const double2 jp2 = d[2 * w.c_al0];
            white_eq610(0.5 * (ip1.x - im1.x), 0.5 * (ip1.y - im1.y), ip1.x - 2 * c.x + im1.x, ip1.y - 2 * c.y + im1.y, -c.x + jp1.x, -c.y + jp1.y,
                        c.x - 2 * jp1.x + jp2.x, c.y - 2 * jp1.y + jp2.y, p, q);  // overwrites the corner value
        } else {
            // applied on top of the corner update above; note the sign flip of the xi derivative (:429-431)
            white_delta(ds_target, theta_target, -0.5 * (ip1.x - im1.x), -0.5 * (ip1.y - im1.y), -c.x + jp1.x, -c.y + jp1.y, p, q);
        }
    }
    wall_pq[w.wall_base + t] = make_double2(p, q);
}

// blends the wall values linearly along j: factor = 1 - j/(nj-1)   (wall_control_function.zig:104-111); one CTA per wall node
__global__ void white_blend_kernel(const WhiteParams* __restrict__ groups, const WhiteNode* __restrict__ nodes, int n_nodes,
                                   const double2* __restrict__ wall_pq, double2* __restrict__ pq) {
    const int k = blockIdx.x;
    if (k >= n_nodes) return;
    const WhiteNode node = nodes[k];
    const WhiteParams w = groups[node.group];
    const int blk = node.t < w.ni0 ? 0 : 1;
    const int i = blk ? node.t - w.ni0 : node.t;
    const int nj = blk ? w.nj1 : w.nj0;
    const double2 a = wall_pq[w.wall_base + node.t];
    double2* line = pq + (blk ? w.off1 : w.off0) + (size_t)i * nj;
    for (int j = threadIdx.x; j < nj; j += blockDim.x) {
        double2 r = a;
        if (j > 0) {
            const double factor = 1 - (double)j / ((double)nj - 1);
            r = make_double2(factor * a.x, factor * a.y);
        }
        line[j] = r;
    }
}

// ---------------------------------------------------------------------------------------------------
// Reductions and the BiCGStab control block.  All scalars stay on the device; the host only polls `done`.
// ---------------------------------------------------------------------------------------------------
struct SolveCtl {
    double rho_old[2], rho_new[2], alpha[2], omega[2], beta[2];
    double tol[2], norm_b[2], norm_r[2];
    double sumsq[2], max_update;   // outer-iteration statistics (smooth.zig:112-137)
    int32_t done[2];               // 1 converged, 2 breakdown, 3 iteration cap
    int32_t iters[2];
    int32_t stage_dummy, _pad;
};

enum ReduceOp : int {
    RED_UPDATE_STATS = 0,  // sums[0,1] -> sumsq, max -> max_update
    RED_INIT = 1,          // sums[0,1] = ||r||^2 ; sums[2,3] = ||b||^2 -> tol, rho_new = ||r||^2 (rhat = r), done if ||r|| <= tol
    RED_ALPHA = 2,         // sums[0,1] = rhat.v -> alpha = rho_new / that
    RED_NORM_S = 3,        // sums[0,1] = ||s||^2 -> done if <= tol
    RED_OMEGA = 4,         // sums[0,1] = t.s ; sums[2,3] = t.t -> omega
    RED_NORM_R = 5,        // sums[0,1] = ||r||^2 ; sums[2,3] = rhat.r -> done?, rho_old = rho_new, rho_new = rhat.r, beta
    RED_RESTART = 6,       // like RED_INIT from the true residual, but keeps tol / iteration counts (restart after breakdown)
};

// Turns the reduced sums (red[0..3]) and max (red[4]) into solver scalars; thread 0 of one CTA.
__device__ __forceinline__ void finalize_reduction(const double* out, int op, SolveCtl* __restrict__ ctl, double rtol, double atol, int max_iters) {
    const double eps = 1e-30;  // breakdown_eps, BiCGStab.zig:280
    if (op == RED_UPDATE_STATS) {
        ctl->sumsq[0] = out[0]; ctl->sumsq[1] = out[1]; ctl->max_update = out[4];
        return;
    }
    for (int c = 0; c < 2; ++c) {
        if (op == RED_INIT || op == RED_RESTART) {
            const double nr = sqrt(out[c]);
            ctl->norm_r[c] = nr;
            if (op == RED_INIT) {
                const double nb = sqrt(out[2 + c]);
                ctl->norm_b[c] = nb;
                ctl->tol[c] = fmax(atol, rtol * nb);             // GMRES.zig:305-306 / BiCGStab.zig:291
                ctl->iters[c] = 0;
            }
            ctl->rho_old[c] = 1.0; ctl->alpha[c] = 1.0; ctl->omega[c] = 1.0;
            ctl->rho_new[c] = out[c];                             // rhat = r
            ctl->done[c] = nr <= ctl->tol[c] ? 1 : (ctl->iters[c] >= max_iters ? 3 : 0);
            if (!ctl->done[c] && fabs(ctl->rho_new[c]) < eps) ctl->done[c] = 2;
            ctl->beta[c] = (ctl->rho_new[c] / ctl->rho_old[c]) * (ctl->alpha[c] / ctl->omega[c]);
            continue;
        }
        if (ctl->done[c]) continue;
        if (op == RED_ALPHA) {
            if (fabs(out[c]) < eps) { ctl->done[c] = 2; ctl->alpha[c] = 0.0; }
            else ctl->alpha[c] = ctl->rho_new[c] / out[c];
        } else if (op == RED_NORM_S) {
            ctl->iters[c] += 1;
            ctl->norm_r[c] = sqrt(out[c]);
            if (ctl->norm_r[c] <= ctl->tol[c]) ctl->done[c] = 1;
        } else if (op == RED_OMEGA) {
            if (fabs(out[2 + c]) < eps) { ctl->done[c] = 2; ctl->omega[c] = 0.0; }
            else {
                ctl->omega[c] = out[c] / out[2 + c];
                if (fabs(ctl->omega[c]) < eps) { ctl->done[c] = 2; ctl->omega[c] = 0.0; }
            }
        } else if (op == RED_NORM_R) {
            ctl->norm_r[c] = sqrt(out[c]);
            if (ctl->norm_r[c] <= ctl->tol[c]) { ctl->done[c] = 1; continue; }
            ctl->rho_old[c] = ctl->rho_new[c];
            ctl->rho_new[c] = out[2 + c];
            if (ctl->iters[c] >= max_iters) { ctl->done[c] = 3; continue; }
            if (fabs(ctl->rho_new[c]) < eps) { ctl->done[c] = 2; continue; }
            ctl->beta[c] = (ctl->rho_new[c] / ctl->rho_old[c]) * (ctl->alpha[c] / ctl->omega[c]);
        }
    }
}

// Sums `n_part` (+ `n_extra`) per-CTA partial records (5 doubles each: 4 sums + 1 max) in a fixed order ->
// deterministic.  The rank-local result goes to red[0..4]; with one rank the solver scalars are finalised in the same
// launch, with several ranks red is all-reduced first (sum / max) and finalize_kernel follows.
template <int NT>
__global__ void __launch_bounds__(NT) reduce_kernel(const double* __restrict__ partials, int n_part, const double* __restrict__ extra, int n_extra,
                                                    double* __restrict__ red, const double* __restrict__ bconst /* RED_INIT: this rank's constant part of ||b||^2 */,
                                                    int finalize, int op, SolveCtl* __restrict__ ctl, double rtol, double atol, int max_iters) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    double mx = 0.0;
    for (int k = threadIdx.x; k < n_part + n_extra; k += NT) {
        const double* p = k < n_part ? partials + (size_t)k * 5 : extra + (size_t)(k - n_part) * 5;
        s[0] += p[0]; s[1] += p[1]; s[2] += p[2]; s[3] += p[3];
        mx = fmax(mx, p[4]);
    }
    __shared__ double out[5];
    block_reduce_store<4, NT>(s, mx, out);
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (op == RED_INIT) { out[2] += bconst[0]; out[3] += bconst[1]; }
    for (int k = 0; k < 5; ++k) red[k] = out[k];
    if (finalize) finalize_reduction(out, op, ctl, rtol, atol, max_iters);
}
__global__ void finalize_kernel(const double* __restrict__ red, int op, SolveCtl* __restrict__ ctl, double rtol, double atol, int max_iters) {
    if (threadIdx.x == 0 && blockIdx.x == 0) finalize_reduction(red, op, ctl, rtol, atol, max_iters);
}
// In-process emulation of several ranks on one GPU (tests): the "all-reduce" of their red[] records.
struct RedPtrs { double* p[16]; };
__global__ void combine_red_kernel(RedPtrs reds, int n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double out[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < n; ++r) {
        for (int k = 0; k < 4; ++k) out[k] += reds.p[r][k];
        out[4] = fmax(out[4], reds.p[r][4]);
    }
    for (int r = 0; r < n; ++r)
        for (int k = 0; k < 5; ++k) reds.p[r][k] = out[k];
}

constexpr int VEC_THREADS = 256;

// p = r + beta (p - omega v)                                        (BiCGStab.zig:310-312)
__global__ void __launch_bounds__(VEC_THREADS) bicg_p_kernel(int64_t n, const SolveCtl* __restrict__ ctl, const double2* __restrict__ r, double2* __restrict__ p,
                                                             const double2* __restrict__ v) {
    const double bx = ctl->beta[0], by = ctl->beta[1], ox = ctl->omega[0], oy = ctl->omega[1];
    const bool dx = ctl->done[0] != 0, dy = ctl->done[1] != 0;
    for (int64_t k = (int64_t)blockIdx.x * VEC_THREADS + threadIdx.x; k < n; k += (int64_t)gridDim.x * VEC_THREADS) {
        const double2 rr = r[k], vv = v[k];
        double2 pp = p[k];
        pp.x = dx ? 0.0 : rr.x + bx * (pp.x - ox * vv.x);
        pp.y = dy ? 0.0 : rr.y + by * (pp.y - oy * vv.y);
        p[k] = pp;
    }
}

// s = r - alpha v; x += alpha p; partial ||s||^2                    (BiCGStab.zig:324-334)
// x is advanced on free rows only (p is read before its connected copies are meaningful for x); the copies of x are
// restored by one affine sync after the solve.
__global__ void __launch_bounds__(VEC_THREADS) bicg_s_kernel(int64_t n, const SolveCtl* __restrict__ ctl, const double2* __restrict__ r, const double2* __restrict__ v,
                                                             double2* __restrict__ s, double2* __restrict__ x, const double2* __restrict__ p,
                                                             double* __restrict__ partials) {
    const double ax = ctl->alpha[0], ay = ctl->alpha[1];
    const bool dx = ctl->done[0] != 0, dy = ctl->done[1] != 0;
    double s0 = 0.0, s1 = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * VEC_THREADS + threadIdx.x; k < n; k += (int64_t)gridDim.x * VEC_THREADS) {
        const double2 rr = r[k], vv = v[k], pp = p[k];
        double2 ss = make_double2(0.0, 0.0), xx = x[k];
        if (!dx) { ss.x = rr.x - ax * vv.x; xx.x += ax * pp.x; s0 += ss.x * ss.x; }
        if (!dy) { ss.y = rr.y - ay * vv.y; xx.y += ay * pp.y; s1 += ss.y * ss.y; }
        s[k] = ss; x[k] = xx;
    }
    double sums[4] = {s0, s1, 0.0, 0.0};
    block_reduce_store<4, VEC_THREADS>(sums, 0.0, partials + (size_t)blockIdx.x * 5);
}

// x += omega s; r = s - omega t; partials ||r||^2 and rhat.r       (BiCGStab.zig:352-366)
__global__ void __launch_bounds__(VEC_THREADS) bicg_r_kernel(int64_t n, const SolveCtl* __restrict__ ctl, const double2* __restrict__ s, const double2* __restrict__ t,
                                                             double2* __restrict__ r, double2* __restrict__ x, const double2* __restrict__ rhat,
                                                             double* __restrict__ partials) {
    const double ox = ctl->omega[0], oy = ctl->omega[1];
    const bool dx = ctl->done[0] != 0, dy = ctl->done[1] != 0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * VEC_THREADS + threadIdx.x; k < n; k += (int64_t)gridDim.x * VEC_THREADS) {
        const double2 ss = s[k], tt = t[k], rh = rhat[k];
        double2 rr = r[k], xx = x[k];
        if (!dx) { xx.x += ox * ss.x; rr.x = ss.x - ox * tt.x; s0 += rr.x * rr.x; s2 += rh.x * rr.x; }
        if (!dy) { xx.y += oy * ss.y; rr.y = ss.y - oy * tt.y; s1 += rr.y * rr.y; s3 += rh.y * rr.y; }
        r[k] = rr; x[k] = xx;
    }
    double sums[4] = {s0, s1, s2, s3};
    block_reduce_store<4, VEC_THREADS>(sums, 0.0, partials + (size_t)blockIdx.x * 5);
}

// x += d  (end of a BiCGStab cycle: the correction accumulated from zero is added to the iterate once, so its
// rounding errors scale with |d|, not |x| -- iterative refinement)
__global__ void __launch_bounds__(VEC_THREADS) add_correction_kernel(int64_t n, double2* __restrict__ x, double2* __restrict__ d) {
    for (int64_t k = (int64_t)blockIdx.x * VEC_THREADS + threadIdx.x; k < n; k += (int64_t)gridDim.x * VEC_THREADS) {
        const double2 dd = d[k];
        double2 xx = x[k];
        xx.x += dd.x; xx.y += dd.y;
        x[k] = xx;
        d[k] = make_double2(0.0, 0.0);
    }
}

// Constant part of ||b||^2 of the reference's full right-hand side (BiCGStab.zig:289-291): fixed rows carry their
// coordinate, connected rows their (periodic) rhs, sliding rows (rhs_x, rhs_y), junction rows their periodic rhs.
// Interior rows are 0 and periodic interface rows are added per solve (they depend on the lagged coordinates).
__global__ void __launch_bounds__(VEC_THREADS) rhs_const_kernel(const RhsTerm* __restrict__ terms, int n, const double2* __restrict__ x, double* __restrict__ out2) {
    double s0 = 0.0, s1 = 0.0;
    for (int k = threadIdx.x; k < n; k += VEC_THREADS) {
        const RhsTerm t = terms[k];
        double bx = t.cx, by = t.cy;
        if (t.from_x | t.from_y) { const double2 v = x[t.g]; if (t.from_x) bx = v.x; if (t.from_y) by = v.y; }
        s0 += bx * bx; s1 += by * by;
    }
    __shared__ double red[5];
    double sums[4] = {s0, s1, 0.0, 0.0};
    block_reduce_store<4, VEC_THREADS>(sums, 0.0, red);
    __syncthreads();
    if (threadIdx.x == 0) { out2[0] = red[0]; out2[1] = red[1]; }
}

// update statistics between two full coordinate fields (Picard modes): sum dx^2, sum dy^2, max |d|
__global__ void __launch_bounds__(VEC_THREADS) diff_stats_kernel(int64_t n, const double2* __restrict__ a, const double2* __restrict__ b, double* __restrict__ partials) {
    double s0 = 0.0, s1 = 0.0, mx = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * VEC_THREADS + threadIdx.x; k < n; k += (int64_t)gridDim.x * VEC_THREADS) {
        const double2 p = a[k], q = b[k];
        const double dx = p.x - q.x, dy = p.y - q.y;
        s0 += dx * dx; s1 += dy * dy;
        mx = fmax(mx, fmax(fabs(dx), fabs(dy)));
    }
    double sums[4] = {s0, s1, 0.0, 0.0};
    block_reduce_store<4, VEC_THREADS>(sums, mx, partials + (size_t)blockIdx.x * 5);
}

}  // namespace tmesh
