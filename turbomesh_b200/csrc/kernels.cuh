// kernels.cuh -- sm_100a device code for turbomesh's hot path (fp64, HBM-bound, no tensor cores).
//
//   tfi_kernel              tfi.linear2dBoundaryBlendedControlFunction          src/core/tfi.zig:112-208
//   winslow_interior_kernel StencilData.init + fillBlockInternalPointData       src/core/smoothing/smooth.zig:192-215, 923-992
//                           fused with the solver's use of the row (relaxation sweep / operator apply / residual),
//                           so the 9 coefficients live only in registers (matrix-free)
//   winslow_boundary_kernel fillBlockConnectionData (interface rows), junction rows, sliding rows, connected copies
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
template <int MODE, bool LAGGED, bool HAS_PQ, int STATS>
__global__ void __launch_bounds__(TILE_J) winslow_interior_kernel(const Tile* __restrict__ tiles, const DevBlock* __restrict__ blocks,
                                                                   const double2* __restrict__ u, const double2* __restrict__ xc,
                                                                   const double2* __restrict__ pq, double2* __restrict__ out, double omega,
                                                                   const double2* __restrict__ dot_a, double* __restrict__ partials) {
    const Tile t = tiles[blockIdx.x];
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
        block_reduce_store<4, TILE_J>(sums, mx, partials + (size_t)blockIdx.x * 5);
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

template <int MODE, bool LAGGED, bool HAS_PQ, int STATS>
__global__ void __launch_bounds__(BND_THREADS) winslow_boundary_kernel(const SmoothedRow* __restrict__ srows, int n_s, const JunctionRow* __restrict__ jrows,
                                                                       int n_j, const SlidingRow* __restrict__ lrows, int n_l,
                                                                       const SlaveRow* __restrict__ slaves, const double2* __restrict__ u,
                                                                       const double2* __restrict__ xc, const double2* __restrict__ pq,
                                                                       double2* __restrict__ out, double omega, const double2* __restrict__ dot_a,
                                                                       double* __restrict__ partials) {
    boundary_rows<MODE, LAGGED, HAS_PQ, STATS>(blockIdx.x, srows, n_s, jrows, n_j, lrows, n_l, slaves, u, xc, pq, out, omega, dot_a, partials);
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
            if (clock64() - t0 > 8000000000ll) { *err = 1; break; }
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

// ---------------------------------------------------------------------------------------------------
// Structured output (the step right after the path, src/core/cgns.zig:69-101, 110-161): the AoS block (x,y interleaved,
// j fastest) as two SoA arrays with i fastest -- what cg_coord_write / cg_field_write take.  A 32x32 tile transpose
// through shared memory: coalesced 16 B loads along j, coalesced 8 B stores along i.  HBM-bound, 32 B per node.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) aos_to_soa_kernel(int ni, int nj, const double2* __restrict__ in, double* __restrict__ x, double* __restrict__ y) {
    __shared__ double2 tile[32][33];
    const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + tx;
        if (i < ni && j < nj) tile[r][tx] = in[(size_t)i * nj + j];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r, i = i0 + tx;
        if (i < ni && j < nj) {
            const double2 v = tile[tx][r];
            x[(size_t)j * ni + i] = v.x;
            y[(size_t)j * ni + i] = v.y;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Edge discretisation, the step right before the path: discrete.Edge.init =
// clustering.create + Curve.interpolate (src/core/discrete.zig:17-31), batched -- one CTA per edge, one thread per point.
//   clustering   clustering.zig:9-17 (uniform), :24-42 (Roberts), :56-95 (Vinokur tanh; delta comes from the host)
//   line         geometry.zig:26-40
//   spline       FittingSpline.interpolate = eval(paramAtArcFraction(u)), spline.zig:74-81, 112-139, 202-222, on the tables of
//                an already fitted spline (params, points, second derivatives, arc-length table)
// The arithmetic of the curves uses round-to-nearest intrinsics in the reference's operation order (no FMA contraction):
// with a uniform clustering the points are bit-exact; pow / tanh of the other clusterings are CUDA's libm (<= 2 ulp).
// ---------------------------------------------------------------------------------------------------
struct EdgeJob {
    int64_t out_off;        // first point of the edge in the output arrays
    int64_t spline_off;     // spline tables in `tables`: params[m], points[2m], zx[m], zy[m], arc[n_samples]
    int32_t n, curve, clustering, spline_m, n_samples, _pad;
    double line[4];         // start x,y ; end x,y
    double alpha, beta, delta, total_length;
};
__device__ __forceinline__ double edge_clustering(const EdgeJob& e, int i) {
    const double n_1 = (double)(e.n - 1);
    const double u = __ddiv_rn((double)i, n_1);
    if (e.clustering == 1) {  // Roberts
        const double tmp = pow(__ddiv_rn(__dadd_rn(e.beta, 1.0), __dsub_rn(e.beta, 1.0)), __ddiv_rn(__dsub_rn(u, e.alpha), __dsub_rn(1.0, e.alpha)));
        const double tbar = __dadd_rn(__dsub_rn(__dmul_rn(__dadd_rn(e.beta, __dmul_rn(2.0, e.alpha)), tmp), e.beta), __dmul_rn(2.0, e.alpha));
        return __ddiv_rn(tbar, __dmul_rn(__dadd_rn(__dmul_rn(2.0, e.alpha), 1.0), __dadd_rn(1.0, tmp)));
    }
    if (e.clustering == 2 && i > 0)  // Vinokur: 1 + tanh(delta/2 (u - 1)) / tanh(delta/2)
        return __dadd_rn(1.0, __ddiv_rn(tanh(__dmul_rn(__dmul_rn(0.5, e.delta), __dsub_rn(u, 1.0))), tanh(__dmul_rn(0.5, e.delta))));
    return u;
}
__device__ __forceinline__ double2 edge_spline_point(const EdgeJob& e, const double* __restrict__ tables, double u) {
    const int m = e.spline_m;
    const double* params = tables + e.spline_off;
    const double* points = params + m;
    const double* zx = points + 2 * m;
    const double* zy = zx + m;
    const double* arc = zy + m;
    double param = 0.0;
    if (e.total_length != 0.0) {  // paramAtArcFraction
        const double target = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        int lo = 0, hi = e.n_samples - 1;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if (arc[mid] < target) lo = mid + 1; else hi = mid;
        }
        if (lo > 0) {
            const double a0 = arc[lo - 1], a1 = arc[lo];
            const double ns = (double)(e.n_samples - 1);
            const double p0 = __ddiv_rn((double)(lo - 1), ns), p1 = __ddiv_rn((double)lo, ns);
            const double t = a1 > a0 ? __ddiv_rn(__dsub_rn(target, a0), __dsub_rn(a1, a0)) : 0.0;
            param = __dadd_rn(p0, __dmul_rn(t, __dsub_rn(p1, p0)));
        }
    }
    const double uu = param < 0.0 ? 0.0 : (param > 1.0 ? 1.0 : param);
    // idx = number of knots params[1..] below uu (the reference scans linearly), at most m-2
    int lo = 0, hi = m - 1;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (params[mid + 1] < uu) lo = mid + 1; else hi = mid;
    }
    const int idx = lo >= m - 1 ? m - 2 : lo;
    const double h = __dsub_rn(params[idx + 1], params[idx]);
    const double a = __ddiv_rn(__dsub_rn(params[idx + 1], uu), h), b = __ddiv_rn(__dsub_rn(uu, params[idx]), h);
    const double a3 = __dsub_rn(__dmul_rn(__dmul_rn(a, a), a), a), b3 = __dsub_rn(__dmul_rn(__dmul_rn(b, b), b), b);
    const double hh = __dmul_rn(h, h);
    auto comp = [&](double y0, double y1, double z0, double z1) {
        const double lin = __dadd_rn(__dmul_rn(a, y0), __dmul_rn(b, y1));
        const double cub = __ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(a3, z0), __dmul_rn(b3, z1)), hh), 6.0);
        return __dadd_rn(lin, cub);
    };
    return make_double2(comp(points[2 * idx], points[2 * idx + 2], zx[idx], zx[idx + 1]), comp(points[2 * idx + 1], points[2 * idx + 3], zy[idx], zy[idx + 1]));
}
__global__ void __launch_bounds__(128) edge_discretize_kernel(const EdgeJob* __restrict__ jobs, const double* __restrict__ tables, double2* __restrict__ points,
                                                              double* __restrict__ clustering) {
    const EdgeJob e = jobs[blockIdx.x];
    for (int i = threadIdx.x; i < e.n; i += blockDim.x) {
        const double u = edge_clustering(e, i);
        double2 p;
        if (e.curve == 0) {
            p.x = __dadd_rn(e.line[0], __dmul_rn(u, __dsub_rn(e.line[2], e.line[0])));
            p.y = __dadd_rn(e.line[1], __dmul_rn(u, __dsub_rn(e.line[3], e.line[1])));
        } else {
            p = edge_spline_point(e, tables, u);
        }
        clustering[e.out_off + i] = u;
        points[e.out_off + i] = p;
    }
}

// ---------------------------------------------------------------------------------------------------
// Viewer buffers (src/gui/lib.zig:227-318): f32 copies of all points with their x / y ranges
// (createPointBuffer) and the wireframe line indices (createWireframeElementBuffer: per block first the segments along j,
// then those along i), produced on the device -- into a mapped GL buffer if the caller hands one in.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) viewer_points_kernel(int64_t n, const double2* __restrict__ x, float2* __restrict__ out, float* __restrict__ partials /* grid x 4 */) {
    float xmin = 3.402823466e+38f, xmax = 1.175494351e-38f, ymin = 3.402823466e+38f, ymax = 1.175494351e-38f;  // floatMax / floatMin as in the reference
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (int64_t)gridDim.x * 256) {
        const double2 p = x[k];
        const float2 f = make_float2(__double2float_rn(p.x), __double2float_rn(p.y));
        out[k] = f;
        xmin = fminf(xmin, f.x); xmax = fmaxf(xmax, f.x); ymin = fminf(ymin, f.y); ymax = fmaxf(ymax, f.y);
    }
    __shared__ float sh[4][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[0][w] = xmin; sh[1][w] = xmax; sh[2][w] = ymin; sh[3][w] = ymax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) { xmin = fminf(xmin, sh[0][q]); xmax = fmaxf(xmax, sh[1][q]); ymin = fminf(ymin, sh[2][q]); ymax = fmaxf(ymax, sh[3][q]); }
        float* p = partials + (size_t)blockIdx.x * 4;
        p[0] = xmin; p[1] = xmax; p[2] = ymin; p[3] = ymax;
    }
}
__global__ void viewer_ranges_kernel(const float* __restrict__ partials, int n_part, float* __restrict__ ranges /* xmin xmax ymin ymax */) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float r[4] = {3.402823466e+38f, 1.175494351e-38f, 3.402823466e+38f, 1.175494351e-38f};
    for (int k = 0; k < n_part; ++k) {
        r[0] = fminf(r[0], partials[4 * k]); r[1] = fmaxf(r[1], partials[4 * k + 1]);
        r[2] = fminf(r[2], partials[4 * k + 2]); r[3] = fmaxf(r[3], partials[4 * k + 3]);
    }
    for (int k = 0; k < 4; ++k) ranges[k] = r[k];
}
struct ViewerBlock { int64_t point_off, index_off; int32_t ni, nj; };  // offsets of the block in the point / index buffers
__global__ void __launch_bounds__(256) viewer_wireframe_kernel(const ViewerBlock* __restrict__ blocks, uint2* __restrict__ lines) {
    const ViewerBlock b = blocks[blockIdx.y];
    const int64_t n_j = (int64_t)b.ni * (b.nj - 1), n_i = (int64_t)b.nj * (b.ni - 1);   // segments along j, along i
    uint2* out = lines + b.index_off / 2;
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n_j + n_i; k += (int64_t)gridDim.x * 256) {
        unsigned p, q;
        if (k < n_j) {              // i-th line, j-th segment: (i*nj + j, i*nj + j + 1)
            const int64_t i = k / (b.nj - 1), j = k - i * (b.nj - 1);
            p = (unsigned)(b.point_off + i * b.nj + j); q = p + 1u;
        } else {                    // j-th column, i-th segment: (i*nj + j, (i+1)*nj + j)
            const int64_t kk = k - n_j, j = kk / (b.ni - 1), i = kk - j * (b.ni - 1);
            p = (unsigned)(b.point_off + i * b.nj + j); q = p + (unsigned)b.nj;
        }
        out[k] = make_uint2(p, q);
    }
}

}  // namespace tmesh
