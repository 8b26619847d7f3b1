// krylov_coarse.cuh -- coarse space of the two-level preconditioner of the BiCGStab kernels: on by default in the phased launches
// of a batch (krylov_phased.cuh), opt-in in the persistent kernel (krylov_kernels.cuh); DESIGN.md 4 has the measurements.
//
// What it replaces: the reference's strong preconditioner (ILU(0), src/core/smoothing/GMRES.zig:199-298).  The inner systems
// are elliptic: with point-Jacobi (BiCGStab.zig's `diagonal`) the iteration count grows with the node count along the
// longest block chain, and line solves or an ILU(0)-class factorisation do not change that (scripts/precond_probe.py: 195
// iterations for the y system of the T106 mesh with Jacobi, 159-243 with line solves).  What does is a coarse space: the
// nodes of a component are grouped into AGGREGATES (ai x aj patches of interior nodes of a block; an interface or junction
// node joins the patch of the interior node next to it; `connected` copies follow their root; fixed and sliding nodes
// stay outside), P = piecewise constant prolongation, and
//
//         M^-1 = I + P (P^T A P)^-1 P^T        on the row-scaled system (A = D^-1 A_ref),
//
// applied as a right preconditioner.  Aggregates are plain node sets, so sub-range, reversed and periodic connections,
// 3- and 5-block junctions need nothing special (the geometric multigrid of multigrid.inl does not converge on the O4H
// topology).  Both solves share the coarse operator: the x and y systems differ only in the sliding rows,
// which are outside the coarse space.
//
//   coarse_assemble_kernel   A_c[I][J] = sum over the rows r of aggregate I of (A 1_J)_r: a warp per non-zero (I, J), the
//                            rows evaluated matrix-free with the lagged coefficients by the row evaluators of
//                            krylov_kernels.cuh applied to the indicator of J.  Fixed summation order.
//   coarse_invert_kernel     G = A_c^-1 in place, Gauss-Jordan without pivoting (A_c is the aggregated form of a row-scaled
//                            elliptic operator: weakly diagonally dominant); a CTA per component.  A vanishing or non-finite
//                            pivot switches the coarse space of that component off (the solve is then plain Jacobi-BiCGStab).
//
// The application (restriction inside the phases, the product with G after the phase's exchange / finalisation) is in
// krylov_kernels.cuh and krylov_phased.cuh (krylov_coarse_kernel).
#pragma once
#include "krylov_kernels.cuh"

namespace tmesh {

struct CoarseItem { int32_t comp, I, J; };

constexpr int COARSE_INV_THREADS = 1024;

// one interior row (node l of block b, not on the rim) applied to u; homogeneous, row-scaled
template <bool HAS_PQ, class U>
__device__ __forceinline__ double k_interior_row_x(const DevBlock& b, int64_t l, U&& u, const double2* __restrict__ xc, const double2* __restrict__ pq) {
    const int nj = b.nj;
    const int64_t g = b.off + l;
    const double2* cb = xc + b.off;
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
    return row_result<MODE_APPLY>(m, rel, C, 1.0).x;
}

template <bool HAS_PQ>
__global__ void __launch_bounds__(256) coarse_assemble_kernel(const CoarseItem* __restrict__ items, int n_items, const KCoarse* __restrict__ coarse,
                                                              const int32_t* __restrict__ mem_ptr, const int32_t* __restrict__ mem_code,
                                                              const int32_t* __restrict__ agg_block, const int32_t* __restrict__ agg, const DevBlock* __restrict__ blocks,
                                                              const SmoothedRow* __restrict__ srows, const JunctionRow* __restrict__ jrows,
                                                              const double2* __restrict__ xc, const double2* __restrict__ pq, double* __restrict__ G, int transposed) {
    const int lane = threadIdx.x & 31;
    const int w = int((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (w >= n_items) return;
    const CoarseItem it = items[w];
    const KCoarse C = coarse[it.comp];
    const int gI = C.agg_base + it.I;
    const int32_t J = it.J;
    auto u = [&](int64_t k) { const double v = __ldg(agg + k) == J ? 1.0 : 0.0; return make_double2(v, v); };
    const DevBlock b = blocks[agg_block[gI]];
    double sum = 0.0;
    for (int q = mem_ptr[gI] + lane; q < mem_ptr[gI + 1]; q += 32) {
        const uint32_t code = (uint32_t)mem_code[q];
        const uint32_t kind = code >> 30, idx = code & 0x3fffffffu;
        double u0, u1;
        if (kind == 0) sum += k_interior_row_x<HAS_PQ>(b, (int64_t)idx - b.off, u, xc, pq);
        else if (kind == 1) sum += k_smoothed<MODE_APPLY, HAS_PQ>(srows[idx], u, xc, pq, u0, u1).res.x;
        else sum += k_junction<MODE_APPLY>(jrows[idx], u).res.x;
    }
    sum = warp_sum(sum);
    // transposed: the inverse of A_c^T is G^T, which the phased path reads column-wise (coalesced over the output index)
    if (lane == 0) G[C.g_off + (transposed ? (int64_t)J * C.nc + it.I : (int64_t)it.I * C.nc + J)] = sum;
}

__global__ void __launch_bounds__(COARSE_INV_THREADS) coarse_invert_kernel(const KCoarse* __restrict__ coarse, double* G, int32_t* __restrict__ ok) {
    extern __shared__ double sh_ci[];
    __shared__ int bad;
    const KCoarse C = coarse[blockIdx.x];
    const int n = C.nc;
    if (n == 0) {
        if (threadIdx.x == 0) ok[blockIdx.x] = 0;
        return;
    }
    double* const A = G + C.g_off;
    double* const col = sh_ci;
    double* const row = sh_ci + n;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        for (int i = threadIdx.x; i < n; i += COARSE_INV_THREADS) {
            col[i] = __ldcg(A + (int64_t)i * n + k);
            row[i] = __ldcg(A + (int64_t)k * n + i);
        }
        __syncthreads();
        const double p = col[k];
        if (!(fabs(p) > 1e-10) || !(fabs(p) < 1e300)) {   // uniform: every thread reads the same value
            if (threadIdx.x == 0) bad = 1;
            break;
        }
        const double ip = 1.0 / p;
        for (int i = ty; i < n; i += COARSE_INV_THREADS / 32) {
            const double f = col[i];
            if (i != k && f == 0.0) continue;              // the row does not change (A_c is sparse at the start; the inverse fills in)
            double* const Ai = A + (int64_t)i * n;
            for (int j = tx; j < n; j += 32) {
                double v;
                if (i == k) v = j == k ? ip : row[j] * ip;
                else v = j == k ? -f * ip : __ldcg(Ai + j) - f * (row[j] * ip);
                __stcg(Ai + j, v);
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) ok[blockIdx.x] = bad ? 0 : 1;
}

}  // namespace tmesh
