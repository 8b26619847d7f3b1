/*
 * turbomesh_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the reference's (pascalPost/turbomesh) hot path:
 * boundary-blended linear TFI + the multi-block elliptic smoother with its built-in Krylov solvers,
 * and of the data-parallel steps either side of it (edge discretisation: clustering + line / spline
 * interpolation; structured SoA output; viewer buffers) at the end of this file.
 * Each function cites the reference file:line it follows (paths relative to the turbomesh repository).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load
 * this library -- as the checker or as the timed CPU baseline, never as part of the product path.
 * The product (turbomesh_b200/, include/turbomesh_gpu.h) must never link, import or call it.
 *
 * PARITY UNPINNED: the reference is Zig 0.15.2 and cannot be built in this image (no zig, no network),
 * and its own tests hold no golden vector for TFI, the stencil assembly, the Krylov solvers, the White
 * control function or the interface coupling (tfi.zig:230-260 is commented out, O4H.zig:576-612 asserts
 * nothing).  Trust is earned instead by (tests/test_oracle_*.py): analytic identities, an independent
 * scipy sparse-LU solve of the assembled system, the adjacent known-answer vectors the reference does
 * hold (umfpack.zig:71-97 5x5 system, discrete.zig:219-290 Edge.combine, csv.zig:59-67, spline.zig:235+)
 * and the reference's runtime invariants (ascending CSR columns, coincident interface nodes, ...).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (Zig never contracts a*b+c into an FMA and never
 * re-associates; sums are plain left-to-right loops).
 *
 * Documented deviations from the reference (all deliberate, none numerical):
 *  D1  zero connections: `for (0..endpoint_ids.len - 1)` underflows (smooth.zig:1364); here a mesh
 *      without connections simply has no junction points (needed for the single-block config).
 *  D2  the CSR capacity over-allocation `9 * sum(cumulative dof)` (smooth.zig:330-343) is replaced by
 *      exactly-sized arrays; ILU(0) copies the used entries only (GMRES.zig:200,228 copy the capacity).
 *  D3  debug assertions / panics become error codes.
 *  D4  Krylov tolerances, restart length and iteration cap are parameters (defaults = GMRES.zig:21-24,
 *      BiCGStab.zig:19-21) so that a tight-tolerance "exact Picard step" mode exists.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_ARG -1
#define ORC_ERR_TOPOLOGY -4
#define ORC_ERR_UNSUPPORTED -5
#define ORC_ERR_NOMEM -7

enum { SIDE_I_MIN = 0, SIDE_I_MAX = 1, SIDE_J_MIN = 2, SIDE_J_MAX = 3 };                       /* boundary.zig:8-13 */
enum { BC_WALL = 0, BC_INLET = 1, BC_OUTLET = 2 };                                             /* boundary.zig:172-176 */
enum { KIND_FIXED = 0, KIND_SMOOTHED = 1, KIND_CONNECTED = 2, KIND_LAPLACIAN = 3, KIND_SLIDING = 4 }; /* smooth.zig:1168-1174 */
enum { SOLVER_GMRES = 0, SOLVER_BICGSTAB = 1 };                                                /* solver.zig:10-15 */
enum { PRECOND_DIAGONAL = 0, PRECOND_ILU0 = 1 };                                               /* preconditioner.zig:1-4 */
enum { CF_LAPLACE = 0, CF_WHITE = 1 };                                                         /* wall_control_function.zig:10-14 */

typedef struct { uint64_t ni, nj; double *xy; } orc_block;                                     /* types.zig:78-101 */
typedef struct { uint64_t block; uint32_t side; uint32_t _pad; uint64_t start, end; } orc_range; /* boundary.zig:15-19 */
typedef struct { orc_range ranges[2]; int32_t has_periodicity; int32_t _pad; double periodicity[2]; } orc_connection; /* boundary.zig:119-123 */
typedef struct { orc_range range; uint32_t kind; uint32_t _pad; } orc_condition;               /* boundary.zig:178-181 */

typedef struct {
    int32_t solver;          /* SOLVER_*                                          */
    int32_t preconditioner;  /* PRECOND_*                                         */
    int32_t control_function;/* CF_*                                              */
    int32_t restart;         /* GMRES.zig:21  (30)                                */
    uint64_t max_iters;      /* GMRES.zig:22  (1000)                              */
    double rtol, atol;       /* GMRES.zig:23-24 (1e-6, 1e-8)                      */
    double ds_target, theta_target; /* wall_control_function.zig:56-61            */
} orc_options;

typedef struct {
    uint64_t outer_iterations;
    uint64_t krylov_iterations;   /* summed over all x and y solves                 */
    uint64_t matvecs;             /* CSR matVec calls                               */
    uint64_t precond_applies;
    uint64_t not_converged;       /* solves that hit the `did not converge` warning */
    double last_sumsq_x, last_sumsq_y, last_residual; /* smooth.zig:112-137        */
    double last_max_update;
    double seconds_fill, seconds_solve, seconds_total;
} orc_stats;

static char g_err[512];
const char *orc_last_error(void) { return g_err; }
#define FAIL(code, ...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return (code); } while (0)

#include <time.h>
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

/* ================================================================================================
 * TFI -- tfi.linear2dBoundaryBlendedControlFunction, src/core/tfi.zig:112-208
 * (Thompson et al., Handbook of Grid Generation, 3.5.1 / 3.6.5).  Operation order exactly as written
 * there: scale() then add() per component, addAll() left to right starting from (0,0).
 * ================================================================================================ */
int orc_tfi(uint64_t n, uint64_t m,
            const double *x_i_min, const double *x_i_max, const double *x_j_min, const double *x_j_max,
            const double *s1, const double *s2, const double *t1, const double *t2, double *out)
{
    if (n < 2 || m < 2) FAIL(ORC_ERR_ARG, "tfi: block needs at least 2x2 nodes");
    /* tfi.zig:135-145 */
    if (s1[0] != 0 || s1[n - 1] != 1.0 || s2[0] != 0 || s2[n - 1] != 1.0 ||
        t1[0] != 0 || t1[m - 1] != 1.0 || t2[0] != 0 || t2[m - 1] != 1.0)
        FAIL(ORC_ERR_ARG, "tfi: clustering must start at 0 and end at 1");
    /* corners, tfi.zig:152-162 */
    const double x00[2] = { x_i_min[0], x_i_min[1] };
    const double xn0[2] = { x_i_min[2 * (n - 1)], x_i_min[2 * (n - 1) + 1] };
    const double x0m[2] = { x_j_min[2 * (m - 1)], x_j_min[2 * (m - 1) + 1] };
    const double xnm[2] = { x_i_max[2 * (n - 1)], x_i_max[2 * (n - 1) + 1] };
    const double tol = 1e-10;
#define APPROX(a, b0, b1) (fabs((a)[0] - (b0)) <= tol && fabs((a)[1] - (b1)) <= tol)
    if (!APPROX(x00, x_j_min[0], x_j_min[1]) || !APPROX(xn0, x_j_max[0], x_j_max[1]) ||
        !APPROX(x0m, x_i_max[0], x_i_max[1]) || !APPROX(xnm, x_j_max[2 * (m - 1)], x_j_max[2 * (m - 1) + 1]))
        FAIL(ORC_ERR_ARG, "tfi: edge corner points are not consistent (tfi.zig:150-162)");
#undef APPROX
    uint64_t idx = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const double s1_i = s1[i], s2_i = s2[i];
        const double xi0[2] = { x_i_min[2 * i], x_i_min[2 * i + 1] };
        const double xim[2] = { x_i_max[2 * i], x_i_max[2 * i + 1] };
        for (uint64_t j = 0; j < m; ++j) {
            const double t1_j = t1[j], t2_j = t2[j];
            const double x0j[2] = { x_j_min[2 * j], x_j_min[2 * j + 1] };
            const double xnj[2] = { x_j_max[2 * j], x_j_max[2 * j + 1] };
            /* tfi.zig:185-186 */
            const double u = ((1.0 - t1_j) * s1_i + t1_j * s2_i) / (1.0 - (s2_i - s1_i) * (t2_j - t1_j));
            const double v = ((1.0 - s1_i) * t1_j + s1_i * t2_j) / (1.0 - (t2_j - t1_j) * (s2_i - s1_i));
            for (int c = 0; c < 2; ++c) {
                /* tfi.zig:188-197 */
                const double u_ij = (1.0 - u) * x0j[c] + u * xnj[c];
                const double v_ij = (1.0 - v) * xi0[c] + v * xim[c];
                double uv = 0.0;                       /* addAll starts from Vec2d.init(0,0), types.zig:51-55 */
                uv = uv + (u * v) * xnm[c];
                uv = uv + (u * (1.0 - v)) * xn0[c];
                uv = uv + ((1.0 - u) * v) * x0m[c];
                uv = uv + ((1.0 - u) * (1.0 - v)) * x00[c];
                out[2 * idx + c] = (u_ij + v_ij) - uv;
            }
            ++idx;
        }
    }
    return ORC_OK;
}

/* ================================================================================================
 * Mesh model helpers -- boundary.zig, smooth.zig:1618-1668
 * ================================================================================================ */
typedef struct {
    const orc_block *blocks; size_t n_blocks;
    const orc_connection *conns; size_t n_conns;
    const orc_condition *bcs; size_t n_bcs;
    uint64_t *block_start;      /* IndexConverter.global_point_index_range_start, smooth.zig:1623-1637 */
    uint64_t *bbuf_start;       /* PointDataBufferIndexConverter.block_range_start_table, boundary.zig:218-239 */
    uint64_t dof, n_boundary;
} mesh_t;

static uint64_t range_len(const orc_range *r) { return r->start > r->end ? r->start - r->end + 1 : r->end - r->start + 1; } /* boundary.zig:21-26 */

/* base local id + signed increment of a range; boundary.zig:28-62 / smooth.zig:1556-1598 */
static void range_walk(const mesh_t *M, const orc_range *r, int64_t *base, int64_t *inc, int64_t *inward)
{
    const int64_t ni = (int64_t)M->blocks[r->block].ni, nj = (int64_t)M->blocks[r->block].nj;
    switch (r->side) {
    case SIDE_I_MIN: *base = (int64_t)r->start * nj;            *inc = nj; *inward = 1;   break;
    case SIDE_I_MAX: *base = (int64_t)r->start * nj + nj - 1;   *inc = nj; *inward = -1;  break;
    case SIDE_J_MIN: *base = (int64_t)r->start;                 *inc = 1;  *inward = nj;  break;
    default:         *base = (ni - 1) * nj + (int64_t)r->start; *inc = 1;  *inward = -nj; break;
    }
    if (r->start > r->end) *inc = -*inc;
}

/* Range.endpoints, boundary.zig:65-77 */
static void range_endpoints(const mesh_t *M, const orc_range *r, uint64_t out[2])
{
    const uint64_t ni = M->blocks[r->block].ni, nj = M->blocks[r->block].nj;
    switch (r->side) {
    case SIDE_I_MIN: out[0] = r->start * nj; out[1] = r->end * nj; break;
    case SIDE_J_MAX: out[0] = (ni - 1) * nj + r->start; out[1] = (ni - 1) * nj + r->end; break;
    case SIDE_I_MAX: out[0] = r->start * nj + nj - 1; out[1] = r->end * nj + nj - 1; break;
    default:         out[0] = r->start; out[1] = r->end; break;
    }
}

/* IndexConverter.localIndex, smooth.zig:1647-1652 */
static size_t block_of_global(const mesh_t *M, uint64_t g)
{
    size_t b = M->n_blocks - 1;
    while (g < M->block_start[b]) --b;
    return b;
}

/* PointDataBufferIndexConverter.bufferIndex, boundary.zig:248-285; -1 = not a boundary node */
static int64_t buffer_index(const mesh_t *M, size_t block, uint64_t local)
{
    const uint64_t ni = M->blocks[block].ni, nj = M->blocks[block].nj;
    const uint64_t i = local / nj, j = local - i * nj;   /* index2d, smooth.zig:1654-1662 */
    uint64_t k;
    if (i == 0) k = j;
    else if (i == ni - 1) k = nj + 2 * (ni - 2) + j;
    else if (j == 0) k = nj + (i - 1) * 2;
    else if (j == nj - 1) k = nj - 1 + i * 2;
    else return -1;
    return (int64_t)(M->bbuf_start[block] + k);
}

static int range_valid(const mesh_t *M, const orc_range *r)
{
    if (r->block >= M->n_blocks || r->side > 3) return 0;
    const uint64_t ext = (r->side == SIDE_I_MIN || r->side == SIDE_I_MAX) ? M->blocks[r->block].ni : M->blocks[r->block].nj;
    return r->start < ext && r->end < ext;
}

static int mesh_init(mesh_t *M, const orc_block *blocks, size_t nb, const orc_connection *conns, size_t nc,
                     const orc_condition *bcs, size_t nbc)
{
    memset(M, 0, sizeof *M);
    if (nb == 0) FAIL(ORC_ERR_ARG, "mesh without blocks");
    M->blocks = blocks; M->n_blocks = nb; M->conns = conns; M->n_conns = nc; M->bcs = bcs; M->n_bcs = nbc;
    M->block_start = (uint64_t *)malloc(nb * sizeof(uint64_t));
    M->bbuf_start = (uint64_t *)malloc(nb * sizeof(uint64_t));
    if (!M->block_start || !M->bbuf_start) FAIL(ORC_ERR_NOMEM, "out of memory");
    for (size_t b = 0; b < nb; ++b) {
        if (blocks[b].ni < 3 || blocks[b].nj < 3) FAIL(ORC_ERR_ARG, "block %zu smaller than 3x3", b);
        M->block_start[b] = M->dof; M->dof += blocks[b].ni * blocks[b].nj;
        M->bbuf_start[b] = M->n_boundary; M->n_boundary += 2 * (blocks[b].ni + blocks[b].nj - 2);
    }
    for (size_t c = 0; c < nc; ++c) {
        if (!range_valid(M, &conns[c].ranges[0]) || !range_valid(M, &conns[c].ranges[1])) FAIL(ORC_ERR_TOPOLOGY, "connection %zu: range out of bounds", c);
        if (range_len(&conns[c].ranges[0]) != range_len(&conns[c].ranges[1])) FAIL(ORC_ERR_TOPOLOGY, "connection %zu: range lengths differ", c);
    }
    for (size_t c = 0; c < nbc; ++c)
        if (!range_valid(M, &bcs[c].range)) FAIL(ORC_ERR_TOPOLOGY, "condition %zu: range out of bounds", c);
    return ORC_OK;
}
static void mesh_free(mesh_t *M) { free(M->block_start); free(M->bbuf_start); }

/* connectionDataCheck, smooth.zig:220-275 */
static int connection_data_check(const mesh_t *M)
{
    const double abs_tol = 1e-15;
    for (size_t c = 0; c < M->n_conns; ++c) {
        const orc_connection *cn = &M->conns[c];
        int64_t b0, d0, n0, b1, d1, n1;
        range_walk(M, &cn->ranges[0], &b0, &d0, &n0);
        range_walk(M, &cn->ranges[1], &b1, &d1, &n1);
        const uint64_t len = range_len(&cn->ranges[0]);
        const double *p0 = M->blocks[cn->ranges[0].block].xy, *p1 = M->blocks[cn->ranges[1].block].xy;
        for (uint64_t k = 0; k < len; ++k) {
            const int64_t l0 = b0 + (int64_t)k * d0, l1 = b1 + (int64_t)k * d1;
            double x0 = p0[2 * l0], y0 = p0[2 * l0 + 1];
            if (cn->has_periodicity) { x0 = x0 + cn->periodicity[0]; y0 = y0 + cn->periodicity[1]; }
            if (!(fabs(x0 - p1[2 * l1]) <= abs_tol && fabs(y0 - p1[2 * l1 + 1]) <= abs_tol))
                FAIL(ORC_ERR_TOPOLOGY, "non matching points for connection %zu point %llu: (%.17g,%.17g) vs (%.17g,%.17g)",
                     c, (unsigned long long)k, x0, y0, p1[2 * l1], p1[2 * l1 + 1]);
        }
    }
    return ORC_OK;
}

/* ================================================================================================
 * Junction ("laplacian") points -- BlockBoundaryPoints.initLaplacianPoints, smooth.zig:1340-1514
 * ================================================================================================ */
typedef struct { uint64_t global_id; double periodicity[2]; } overlap_t;       /* smooth.zig:1334-1337 */
typedef struct {
    overlap_t overlapping[4]; size_t n_overlapping;                           /* BoundedArray(...,4), smooth.zig:1221 */
    int32_t stencil_ids[6]; size_t n_stencil;                                 /* BoundedArray(c_int,6), smooth.zig:1224 */
    double rhs[2];
} laplacian_t;

static int append_if_unique(laplacian_t *lp, uint64_t id, const double p[2])    /* smooth.zig:1516-1522 */
{
    for (size_t k = 0; k < lp->n_overlapping; ++k) if (lp->overlapping[k].global_id == id) return ORC_OK;
    if (lp->n_overlapping >= 4) FAIL(ORC_ERR_UNSUPPORTED, "junction with more than 4 overlapping points (smooth.zig:1221)");
    lp->overlapping[lp->n_overlapping].global_id = id;
    lp->overlapping[lp->n_overlapping].periodicity[0] = p[0];
    lp->overlapping[lp->n_overlapping].periodicity[1] = p[1];
    lp->n_overlapping++;
    return ORC_OK;
}
static int cmp_overlap(const void *a, const void *b) { uint64_t x = ((const overlap_t *)a)->global_id, y = ((const overlap_t *)b)->global_id; return x < y ? -1 : x > y; }
static int cmp_laplacian(const void *a, const void *b) { uint64_t x = ((const laplacian_t *)a)->overlapping[0].global_id, y = ((const laplacian_t *)b)->overlapping[0].global_id; return x < y ? -1 : x > y; }
static int cmp_i32(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return x < y ? -1 : x > y; }

static void conn_periodicity(const mesh_t *M, size_t c, double p[2])
{
    if (M->conns[c].has_periodicity) { p[0] = M->conns[c].periodicity[0]; p[1] = M->conns[c].periodicity[1]; }
    else { p[0] = 0; p[1] = 0; }
}

static int init_laplacian_points(const mesh_t *M, laplacian_t **out, size_t *n_out)
{
    *out = NULL; *n_out = 0;
    if (M->n_conns == 0) return ORC_OK;                                       /* deviation D1 */
    const size_t ne = M->n_conns * 4;
    uint64_t *ids = (uint64_t *)malloc(ne * sizeof(uint64_t));
    laplacian_t *lps = (laplacian_t *)calloc(M->n_conns * 2 + 1, sizeof(laplacian_t)); /* capacity hint smooth.zig:1360 (grows there; here 2/connection suffices: each new point consumes one pair) */
    size_t cap = M->n_conns * 2 + 1, n = 0;
    if (!ids || !lps) { free(ids); free(lps); FAIL(ORC_ERR_NOMEM, "out of memory"); }
    /* smooth.zig:1349-1356 */
    for (size_t c = 0; c < M->n_conns; ++c) {
        uint64_t e0[2], e1[2];
        range_endpoints(M, &M->conns[c].ranges[0], e0);
        range_endpoints(M, &M->conns[c].ranges[1], e1);
        ids[4 * c + 0] = M->block_start[M->conns[c].ranges[0].block] + e0[0];
        ids[4 * c + 1] = M->block_start[M->conns[c].ranges[1].block] + e1[0];
        ids[4 * c + 2] = M->block_start[M->conns[c].ranges[0].block] + e0[1];
        ids[4 * c + 3] = M->block_start[M->conns[c].ranges[1].block] + e1[1];
    }
    int rc = ORC_OK;
    /* smooth.zig:1364-1439 */
    for (size_t a = 0; a + 1 < ne && rc == ORC_OK; ++a) {
        const uint64_t endpoint = ids[a];
        for (size_t b = a + 1; b < ne && rc == ORC_OK; ++b) {
            if (endpoint != ids[b]) continue;
            int found = 0;
            for (size_t l = 0; l < n && rc == ORC_OK; ++l) {
                /* NOTE: the reference iterates over a slice taken before the append; appended entries are
                 * therefore not revisited in this inner loop (smooth.zig:1373). */
                const size_t n_before = lps[l].n_overlapping;
                for (size_t k = 0; k < n_before && rc == ORC_OK; ++k) {
                    if (lps[l].overlapping[k].global_id == endpoint) {
                        found = 1;
                        const size_t to_add = (b % 2 == 0) ? b + 1 : b - 1;   /* smooth.zig:1378 */
                        double p[2]; conn_periodicity(M, to_add / 4, p);
                        rc = append_if_unique(&lps[l], ids[to_add], p);
                    }
                }
            }
            if (!found && rc == ORC_OK) {
                const size_t pt0 = a / 2, pt1 = b / 2;                         /* smooth.zig:1390 */
                if (pt0 == pt1) { rc = ORC_ERR_TOPOLOGY; snprintf(g_err, sizeof g_err, "degenerate connection: both sides share an endpoint (smooth.zig:1391)"); break; }
                if (n == cap) { rc = ORC_ERR_NOMEM; snprintf(g_err, sizeof g_err, "junction table overflow"); break; }
                laplacian_t *lp = &lps[n];
                memset(lp, 0, sizeof *lp);
                double p[2], zero[2] = { 0, 0 };
                conn_periodicity(M, pt0 / 2, p);
                lp->overlapping[0].global_id = ids[pt0 * 2]; lp->overlapping[0].periodicity[0] = 0; lp->overlapping[0].periodicity[1] = 0;
                lp->overlapping[1].global_id = ids[pt0 * 2 + 1]; lp->overlapping[1].periodicity[0] = p[0]; lp->overlapping[1].periodicity[1] = p[1];
                lp->n_overlapping = 2; (void)zero;
                if (lp->overlapping[0].global_id == lp->overlapping[1].global_id) { rc = ORC_ERR_TOPOLOGY; snprintf(g_err, sizeof g_err, "connection joins a node with itself (smooth.zig:1404)"); break; }
                conn_periodicity(M, pt1 / 2, p);
                rc = append_if_unique(lp, ids[pt1 * 2], p);
                if (rc == ORC_OK) rc = append_if_unique(lp, ids[pt1 * 2 + 1], p);
                ++n;
            }
        }
    }
    if (rc != ORC_OK) { free(ids); free(lps); return rc; }
    /* smooth.zig:1441-1455 */
    for (size_t l = 0; l < n; ++l) qsort(lps[l].overlapping, lps[l].n_overlapping, sizeof(overlap_t), cmp_overlap);
    qsort(lps, n, sizeof(laplacian_t), cmp_laplacian);
    /* smooth.zig:1457-1511 */
    for (size_t l = 0; l < n && rc == ORC_OK; ++l) {
        laplacian_t *lp = &lps[l];
        lp->n_stencil = 0; lp->rhs[0] = lp->rhs[1] = 0;
        lp->stencil_ids[lp->n_stencil++] = (int32_t)lp->overlapping[0].global_id;
        for (size_t k = 0; k < lp->n_overlapping && rc == ORC_OK; ++k) {
            const uint64_t g = lp->overlapping[k].global_id;
            const size_t blk = block_of_global(M, g);
            const uint64_t local = g - M->block_start[blk];
            const uint64_t ni = M->blocks[blk].ni, nj = M->blocks[blk].nj;
            const uint64_t i = local / nj, j = local - i * nj;
            uint64_t pi[2], pj[2]; int np = 0;
            if (i == 0) {
                if (j == 0) { pi[0] = 1; pj[0] = 1; np = 1; }
                else if (j == nj - 1) { pi[0] = 1; pj[0] = nj - 2; np = 1; }
                else { pi[0] = 1; pj[0] = j - 1; pi[1] = 1; pj[1] = j + 1; np = 2; }
            } else if (i == ni - 1) {
                if (j == 0) { pi[0] = ni - 2; pj[0] = 1; np = 1; }
                else if (j == nj - 1) { pi[0] = ni - 2; pj[0] = nj - 2; np = 1; }
                else { pi[0] = ni - 2; pj[0] = j - 1; pi[1] = ni - 2; pj[1] = j + 1; np = 2; }
            } else if (j == 0) { pi[0] = i - 1; pj[0] = 1; pi[1] = i + 1; pj[1] = 1; np = 2; }
            else if (j == nj - 1) { pi[0] = i - 1; pj[0] = j - 1; pi[1] = i + 1; pj[1] = j - 1; np = 2; }
            else { rc = ORC_ERR_TOPOLOGY; snprintf(g_err, sizeof g_err, "junction copy is not a block boundary node (smooth.zig:1489)"); break; }
            for (int q = 0; q < np; ++q) {
                if (lp->n_stencil >= 6) { rc = ORC_ERR_UNSUPPORTED; snprintf(g_err, sizeof g_err, "junction stencil with more than 6 ids (smooth.zig:1224)"); break; }
                lp->stencil_ids[lp->n_stencil++] = (int32_t)(M->block_start[blk] + pi[q] * nj + pj[q]);
                lp->rhs[0] += lp->overlapping[k].periodicity[0];
                lp->rhs[1] += lp->overlapping[k].periodicity[1];
            }
        }
        qsort(lp->stencil_ids, lp->n_stencil, sizeof(int32_t), cmp_i32);
    }
    free(ids);
    if (rc != ORC_OK) { free(lps); return rc; }
    *out = lps; *n_out = n;
    return ORC_OK;
}

/* ================================================================================================
 * RowCompressedMatrixSystem2d -- smooth.zig:277-1166
 * ================================================================================================ */
typedef struct {
    int64_t count;
    int64_t inward[2];  /* first_internal_point_shift      */
    int64_t along[2];   /* in_connection_direction_shift   */
    int64_t pos[2];     /* position (local ids)            */
} fill_it;              /* RangeFillMatrixIterator, smooth.zig:1531-1599 */

static void fill_it_init(fill_it *it, const mesh_t *M, const orc_connection *cn)
{
    for (int s = 0; s < 2; ++s) range_walk(M, &cn->ranges[s], &it->pos[s], &it->along[s], &it->inward[s]);
    it->count = (int64_t)range_len(&cn->ranges[1]);
}
static int fill_it_next(fill_it *it, int64_t out[2])
{
    if (it->count == 0) return 0;
    out[0] = it->pos[0]; out[1] = it->pos[1];
    it->count -= 1; it->pos[0] += it->along[0]; it->pos[1] += it->along[1];
    return 1;
}

typedef struct orc_system {
    mesh_t M;
    orc_options opt;
    uint8_t *kind;                    /* BlockBoundaryPoints.kind.buffer                  */
    laplacian_t *lps; size_t n_lps;   /* BlockBoundaryPoints.laplacian_points              */
    int32_t *lhs_p, *lhs_i; double *lhs_values; size_t nnz, nnz_cap;
    double *rhs_x, *rhs_y, *x_new, *y_new;
    double *cf;                       /* ControlFunction.data (P,Q interleaved)            */
    /* solver state */
    int seeded;
    double *ilu; int32_t *diag_pos, *marker;
    double *work; size_t work_len;
    orc_stats stats;
} orc_system;

static int nz_push(orc_system *S, int32_t v)
{
    if (S->nnz == S->nnz_cap) {
        size_t cap = S->nnz_cap ? S->nnz_cap * 2 : 1024;
        int32_t *p = (int32_t *)realloc(S->lhs_i, cap * sizeof(int32_t));
        if (!p) FAIL(ORC_ERR_NOMEM, "out of memory");
        S->lhs_i = p; S->nnz_cap = cap;
    }
    S->lhs_i[S->nnz++] = v;
    return ORC_OK;
}

/* BlockBoundaryPoints.init, smooth.zig:1234-1332 */
static int init_kinds(orc_system *S)
{
    const mesh_t *M = &S->M;
    S->kind = (uint8_t *)malloc(M->n_boundary);
    if (!S->kind) FAIL(ORC_ERR_NOMEM, "out of memory");
    memset(S->kind, KIND_FIXED, M->n_boundary);
    for (size_t l = 0; l < S->n_lps; ++l) {
        for (size_t k = 0; k < S->lps[l].n_overlapping; ++k) {
            const uint64_t g = S->lps[l].overlapping[k].global_id;
            const size_t b = block_of_global(M, g);
            const int64_t bi = buffer_index(M, b, g - M->block_start[b]);
            if (bi < 0) FAIL(ORC_ERR_TOPOLOGY, "junction point is not a boundary node");
            S->kind[bi] = (k == 0) ? KIND_LAPLACIAN : KIND_CONNECTED;
        }
    }
    for (size_t c = 0; c < M->n_bcs; ++c) {
        if (M->bcs[c].kind == BC_WALL) continue;
        int64_t base, inc, inward; range_walk(M, &M->bcs[c].range, &base, &inc, &inward);
        const uint64_t len = range_len(&M->bcs[c].range);
        for (uint64_t k = 0; k < len; ++k) {
            const int64_t bi = buffer_index(M, M->bcs[c].range.block, (uint64_t)(base + (int64_t)k * inc));
            S->kind[bi] = KIND_SLIDING;
        }
    }
    for (size_t c = 0; c < M->n_conns; ++c) {
        const orc_connection *cn = &M->conns[c];
        int64_t b0, d0, n0, b1, d1, n1;
        range_walk(M, &cn->ranges[0], &b0, &d0, &n0);
        range_walk(M, &cn->ranges[1], &b1, &d1, &n1);
        const uint64_t len = range_len(&cn->ranges[0]);
        for (uint64_t k = 0; k < len; ++k) {
            const int64_t i0 = buffer_index(M, cn->ranges[0].block, (uint64_t)(b0 + (int64_t)k * d0));
            const int64_t i1 = buffer_index(M, cn->ranges[1].block, (uint64_t)(b1 + (int64_t)k * d1));
            if (k == 0 || k == len - 1) {                        /* smooth.zig:1286-1300, 1314-1328 */
                if (S->kind[i0] == KIND_FIXED || S->kind[i0] == KIND_SLIDING) S->kind[i1] = KIND_CONNECTED;
            } else {                                             /* smooth.zig:1303-1311 */
                S->kind[i0] = KIND_SMOOTHED; S->kind[i1] = KIND_CONNECTED;
            }
        }
    }
    return ORC_OK;
}

/* computeConnectionStencilPositions, smooth.zig:518-616 */
static int connection_stencil_positions(const orc_connection *cn, const fill_it *it, size_t pos[9])
{
    if (cn->ranges[0].block == cn->ranges[1].block) {
        if (cn->ranges[0].side == SIDE_I_MIN && cn->ranges[1].side == SIDE_I_MAX) {
            if (it->along[0] > 0) { const size_t p[9] = { 1, 4, 7, 0, 3, 6, 2, 5, 8 }; memcpy(pos, p, sizeof p); }
            else                  { const size_t p[9] = { 7, 4, 1, 6, 3, 0, 8, 5, 2 }; memcpy(pos, p, sizeof p); }
            return ORC_OK;
        }
        FAIL(ORC_ERR_UNSUPPORTED, "same-block connections are only supported as i_min -> i_max (smooth.zig:522-559)");
    }
    if (!(cn->ranges[0].block < cn->ranges[1].block)) FAIL(ORC_ERR_TOPOLOGY, "connection ranges[0].block must be <= ranges[1].block (smooth.zig:562,627)");
    const int dm0 = it->along[0] > 0 ? 1 : -1, dm1 = it->along[1] > 0 ? 1 : -1;
    switch (cn->ranges[0].side) {
    case SIDE_I_MIN: pos[0] = (size_t)(3 - 2 * dm0); pos[1] = 3; pos[2] = (size_t)(3 + 2 * dm0); pos[3] = (size_t)(2 - 2 * dm0); pos[4] = 2; pos[5] = (size_t)(2 + 2 * dm0); break;
    case SIDE_I_MAX: pos[0] = (size_t)(2 - 2 * dm0); pos[1] = 2; pos[2] = (size_t)(2 + 2 * dm0); pos[3] = (size_t)(3 - 2 * dm0); pos[4] = 3; pos[5] = (size_t)(3 + 2 * dm0); break;
    case SIDE_J_MIN: pos[0] = (size_t)(4 - dm0); pos[1] = 4; pos[2] = (size_t)(4 + dm0); pos[3] = (size_t)(1 - dm0); pos[4] = 1; pos[5] = (size_t)(1 + dm0); break;
    default:         pos[0] = (size_t)(1 - dm0); pos[1] = 1; pos[2] = (size_t)(1 + dm0); pos[3] = (size_t)(4 - dm0); pos[4] = 4; pos[5] = (size_t)(4 + dm0); break;
    }
    pos[6] = (size_t)(7 - dm1); pos[7] = 7; pos[8] = (size_t)(7 + dm1);
    return ORC_OK;
}

/* initNonZeroMatrixEntriesForBoundaryPoint, smooth.zig:421-458 */
static int nz_boundary_point(orc_system *S, uint64_t *bpid, int32_t *row, size_t *lap_count)
{
    int rc = ORC_OK;
    switch (S->kind[*bpid]) {
    case KIND_FIXED: rc = nz_push(S, *row); break;
    case KIND_SMOOTHED: for (int k = 0; k < 9 && rc == ORC_OK; ++k) rc = nz_push(S, -1); break;
    case KIND_CONNECTED: rc = nz_push(S, -1); if (rc == ORC_OK) rc = nz_push(S, -1); break;
    case KIND_LAPLACIAN: {
        if (*lap_count >= S->n_lps) FAIL(ORC_ERR_TOPOLOGY, "junction bookkeeping out of sync");
        const laplacian_t *lp = &S->lps[*lap_count];
        for (size_t k = 0; k < lp->n_stencil && rc == ORC_OK; ++k) rc = nz_push(S, lp->stencil_ids[k]);
        *lap_count += 1;
    } break;
    default: rc = nz_push(S, -1); if (rc == ORC_OK) rc = nz_push(S, -1); break; /* sliding_circ */
    }
    *bpid += 1;
    S->lhs_p[*row + 1] = (int32_t)S->nnz;
    *row += 1;
    return rc;
}

/* initNonZeroMatrixForConnectionEndpoint, smooth.zig:695-721 */
static int nz_connection_endpoint(orc_system *S, const orc_connection *cn, const int64_t local[2])
{
    const mesh_t *M = &S->M;
    const int64_t bi = buffer_index(M, cn->ranges[0].block, (uint64_t)local[0]);
    if (bi < 0) FAIL(ORC_ERR_TOPOLOGY, "connection endpoint is not a boundary node");
    switch (S->kind[bi]) {
    case KIND_FIXED: case KIND_SLIDING: {
        const int32_t g0 = (int32_t)(M->block_start[cn->ranges[0].block] + (uint64_t)local[0]);
        const int32_t g1 = (int32_t)(M->block_start[cn->ranges[1].block] + (uint64_t)local[1]);
        if (!(g0 < g1)) FAIL(ORC_ERR_TOPOLOGY, "connection endpoint: side-0 id must be below side-1 id (smooth.zig:712)");
        const size_t s = (size_t)S->lhs_p[g1];
        if (S->lhs_p[g1 + 1] - S->lhs_p[g1] != 2) FAIL(ORC_ERR_TOPOLOGY, "connected endpoint row does not have 2 entries");
        S->lhs_i[s] = g0; S->lhs_i[s + 1] = g1;
    } break;
    case KIND_LAPLACIAN: case KIND_CONNECTED: break;
    default: FAIL(ORC_ERR_UNSUPPORTED, "connection endpoint lies inside another connection (smooth.zig:719 unreachable)");
    }
    return ORC_OK;
}

/* initNonZeroMatrixEntries, smooth.zig:723-778 (+ :460-516, :618-693) */
static int init_nonzero_entries(orc_system *S)
{
    const mesh_t *M = &S->M;
    int rc;
    uint64_t bpid = 0; int32_t row = 0; size_t lap_count = 0;
    S->lhs_p[0] = 0;
    for (size_t b = 0; b < M->n_blocks; ++b) {
        const int32_t col = (int32_t)M->blocks[b].nj;
        for (uint64_t j = 0; j < M->blocks[b].nj; ++j) if ((rc = nz_boundary_point(S, &bpid, &row, &lap_count))) return rc;
        for (uint64_t i = 1; i + 1 < M->blocks[b].ni; ++i) {
            if ((rc = nz_boundary_point(S, &bpid, &row, &lap_count))) return rc;
            for (uint64_t j = 1; j + 1 < M->blocks[b].nj; ++j) {               /* smooth.zig:493-505 */
                const int32_t e[9] = { row - col - 1, row - col, row - col + 1, row - 1, row, row + 1, row + col - 1, row + col, row + col + 1 };
                for (int k = 0; k < 9; ++k) if ((rc = nz_push(S, e[k]))) return rc;
                S->lhs_p[row + 1] = (int32_t)S->nnz; row += 1;
            }
            if ((rc = nz_boundary_point(S, &bpid, &row, &lap_count))) return rc;
        }
        for (uint64_t j = 0; j < M->blocks[b].nj; ++j) if ((rc = nz_boundary_point(S, &bpid, &row, &lap_count))) return rc;
    }
    /* smooth.zig:738-747 */
    for (size_t l = 0; l < S->n_lps; ++l) {
        const int32_t smoothed = (int32_t)S->lps[l].overlapping[0].global_id;
        for (size_t k = 1; k < S->lps[l].n_overlapping; ++k) {
            const int32_t g = (int32_t)S->lps[l].overlapping[k].global_id;
            const size_t s = (size_t)S->lhs_p[g];
            if (S->lhs_p[g + 1] - S->lhs_p[g] != 2) FAIL(ORC_ERR_TOPOLOGY, "junction copy row does not have 2 entries");
            S->lhs_i[s] = smoothed; S->lhs_i[s + 1] = g;
        }
    }
    /* initNonZeroMatrixEntriesConnectionBased, smooth.zig:618-693 */
    for (size_t c = 0; c < M->n_conns; ++c) {
        const orc_connection *cn = &M->conns[c];
        if (!(cn->ranges[0].block <= cn->ranges[1].block)) FAIL(ORC_ERR_TOPOLOGY, "connection %zu: ranges[0].block must be <= ranges[1].block (smooth.zig:627)", c);
        if (!(range_len(&cn->ranges[0]) > 5)) FAIL(ORC_ERR_UNSUPPORTED, "connection %zu: needs at least 6 nodes (lenInternal() > 3, smooth.zig:631)", c);
        fill_it it; fill_it_init(&it, M, cn);
        size_t pos[9];
        if ((rc = connection_stencil_positions(cn, &it, pos))) return rc;
        int64_t loc[2];
        fill_it_next(&it, loc);
        if ((rc = nz_connection_endpoint(S, cn, loc))) return rc;
        const int64_t middle = it.count - 1;
        for (int64_t k = 0; k < middle; ++k) {
            fill_it_next(&it, loc);
            const int32_t g0 = (int32_t)(M->block_start[cn->ranges[0].block] + (uint64_t)loc[0]);
            const int32_t g1 = (int32_t)(M->block_start[cn->ranges[1].block] + (uint64_t)loc[1]);
            if (!(g0 < g1)) FAIL(ORC_ERR_TOPOLOGY, "connection %zu: side-0 id must be below side-1 id (smooth.zig:651)", c);
            { const size_t s = (size_t)S->lhs_p[g1];
              if (S->lhs_p[g1 + 1] - S->lhs_p[g1] != 2) FAIL(ORC_ERR_TOPOLOGY, "connection %zu: partner row is not a `connected` row", c);
              S->lhs_i[s] = g0; S->lhs_i[s + 1] = g1; }
            { const size_t s = (size_t)S->lhs_p[g0];
              if (S->lhs_p[g0 + 1] - S->lhs_p[g0] != 9) FAIL(ORC_ERR_TOPOLOGY, "connection %zu: side-0 row is not a `smoothed` row", c);
              const int32_t d0 = (int32_t)it.along[0], n0 = (int32_t)it.inward[0], d1 = (int32_t)it.along[1], n1 = (int32_t)it.inward[1];
              S->lhs_i[s + pos[0]] = g0 - d0 + n0; S->lhs_i[s + pos[1]] = g0 + n0; S->lhs_i[s + pos[2]] = g0 + d0 + n0;
              S->lhs_i[s + pos[3]] = g0 - d0;      S->lhs_i[s + pos[4]] = g0;      S->lhs_i[s + pos[5]] = g0 + d0;
              S->lhs_i[s + pos[6]] = g1 - d1 + n1; S->lhs_i[s + pos[7]] = g1 + n1; S->lhs_i[s + pos[8]] = g1 + d1 + n1;
              for (int q = 0; q < 8; ++q)                                     /* smooth.zig:679-687 */
                  if (S->lhs_i[s + q] >= S->lhs_i[s + q + 1]) FAIL(ORC_ERR_TOPOLOGY, "connection %zu: interface row columns are not ascending (smooth.zig:679-687)", c);
            }
        }
        loc[0] = it.pos[0]; loc[1] = it.pos[1];
        if ((rc = nz_connection_endpoint(S, cn, loc))) return rc;
    }
    /* smooth.zig:751-777 */
    for (size_t c = 0; c < M->n_bcs; ++c) {
        const orc_condition *bc = &M->bcs[c];
        if (bc->kind == BC_WALL) FAIL(ORC_ERR_UNSUPPORTED, "wall boundary conditions are `unreachable` in the reference (smooth.zig:775)");
        int64_t base, inc, inward; range_walk(M, &bc->range, &base, &inc, &inward);
        const uint64_t len = range_len(&bc->range);
        for (uint64_t k = 0; k < len; ++k) {
            const uint64_t local = (uint64_t)(base + (int64_t)k * inc);
            if (S->kind[buffer_index(M, bc->range.block, local)] != KIND_SLIDING) continue;
            const int32_t g = (int32_t)(M->block_start[bc->range.block] + local);
            const size_t s = (size_t)S->lhs_p[g];
            if (inward > 0) { S->lhs_i[s] = g; S->lhs_i[s + 1] = g + (int32_t)inward; }
            else            { S->lhs_i[s] = g + (int32_t)inward; S->lhs_i[s + 1] = g; }
        }
    }
    for (size_t k = 0; k < S->nnz; ++k) if (S->lhs_i[k] < 0 || (uint64_t)S->lhs_i[k] >= M->dof) FAIL(ORC_ERR_TOPOLOGY, "matrix entry %zu was never initialised (inconsistent topology)", k);
    return ORC_OK;
}

/* initBoundaryPointData + initBoundaryData, smooth.zig:780-921 */
static void init_boundary_point_data(orc_system *S, const orc_block *blk, uint64_t *bpid, uint64_t *row, uint64_t *pid, size_t *nz, size_t *lap_count)
{
    switch (S->kind[*bpid]) {
    case KIND_FIXED:
        S->lhs_values[*nz] = 1; *nz += 1;
        S->rhs_x[*row] = blk->xy[2 * *pid]; S->rhs_y[*row] = blk->xy[2 * *pid + 1];
        break;
    case KIND_SMOOTHED: *nz += 9; S->rhs_x[*row] = 0; S->rhs_y[*row] = 0; break;
    case KIND_CONNECTED:
        S->lhs_values[*nz] = 1; S->lhs_values[*nz + 1] = -1; *nz += 2;
        S->rhs_x[*row] = 0; S->rhs_y[*row] = 0;
        break;
    case KIND_LAPLACIAN: {
        const laplacian_t *lp = &S->lps[*lap_count];
        const int32_t self = (int32_t)lp->overlapping[0].global_id;
        size_t where = 0;
        for (size_t k = 0; k < lp->n_stencil; ++k) { if (lp->stencil_ids[k] == self) break; where += 1; }
        for (size_t k = 0; k < lp->n_stencil; ++k) S->lhs_values[*nz + k] = 1;
        S->lhs_values[*nz + where] = -(double)lp->n_stencil + 1;
        *nz += lp->n_stencil;
        S->rhs_x[*row] = 0; S->rhs_y[*row] = 0;
        *lap_count += 1;
    } break;
    default: /* sliding_circ: values are set per solve (fillX/YSpecific) */
        *nz += 2;
        S->rhs_x[*row] = blk->xy[2 * *pid]; S->rhs_y[*row] = 0.0;
        break;
    }
    *row += 1; *bpid += 1; *pid += 1;
}

static void init_boundary_data(orc_system *S)
{
    const mesh_t *M = &S->M;
    size_t nz = 0, lap_count = 0; uint64_t row = 0, bpid = 0;
    for (size_t b = 0; b < M->n_blocks; ++b) {
        const orc_block *blk = &M->blocks[b];
        uint64_t pid = 0;
        for (uint64_t j = 0; j < blk->nj; ++j) init_boundary_point_data(S, blk, &bpid, &row, &pid, &nz, &lap_count);
        for (uint64_t i = 1; i + 1 < blk->ni; ++i) {
            init_boundary_point_data(S, blk, &bpid, &row, &pid, &nz, &lap_count);
            for (uint64_t j = 1; j + 1 < blk->nj; ++j) { nz += 9; pid += 1; row += 1; }
            init_boundary_point_data(S, blk, &bpid, &row, &pid, &nz, &lap_count);
        }
        for (uint64_t j = 0; j < blk->nj; ++j) init_boundary_point_data(S, blk, &bpid, &row, &pid, &nz, &lap_count);
    }
    /* periodicity, smooth.zig:903-915: every side-1 node of a periodic connection, end points included */
    for (size_t c = 0; c < M->n_conns; ++c) {
        const orc_connection *cn = &M->conns[c];
        if (!cn->has_periodicity) continue;
        fill_it it; fill_it_init(&it, M, cn);
        int64_t loc[2];
        while (fill_it_next(&it, loc)) {
            const uint64_t g1 = M->block_start[cn->ranges[1].block] + (uint64_t)loc[1];
            S->rhs_x[g1] = -cn->periodicity[0]; S->rhs_y[g1] = -cn->periodicity[1];
        }
    }
    for (size_t l = 0; l < S->n_lps; ++l) {                                   /* smooth.zig:917-920 */
        const uint64_t g = S->lps[l].overlapping[0].global_id;
        S->rhs_x[g] = S->lps[l].rhs[0]; S->rhs_y[g] = S->lps[l].rhs[1];
    }
}

/* StencilData.init, smooth.zig:192-215; out order = enum index (i_j, ip1_j, im1_j, i_jp1, i_jm1, ip1_jp1, ip1_jm1, im1_jp1, im1_jm1) */
enum { ST_I_J, ST_IP1_J, ST_IM1_J, ST_I_JP1, ST_I_JM1, ST_IP1_JP1, ST_IP1_JM1, ST_IM1_JP1, ST_IM1_JM1 };
static void stencil_init(const double im1_j[2], const double ip1_j[2], const double i_jm1[2], const double i_jp1[2], double P, double Q, double d[9])
{
    const double x_xi = 0.5 * (ip1_j[0] - im1_j[0]);
    const double x_eta = 0.5 * (i_jp1[0] - i_jm1[0]);
    const double y_xi = 0.5 * (ip1_j[1] - im1_j[1]);
    const double y_eta = 0.5 * (i_jp1[1] - i_jm1[1]);
    const double g22 = x_eta * x_eta + y_eta * y_eta;
    const double g12 = x_xi * x_eta + y_xi * y_eta;
    const double g11 = x_xi * x_xi + y_xi * y_xi;
    d[ST_I_J] = -2.0 * g22 - 2.0 * g11;
    d[ST_IP1_J] = g22 * (1 + 0.5 * P);
    d[ST_IM1_J] = g22 * (1 - 0.5 * P);
    d[ST_I_JP1] = g11 * (1 + 0.5 * Q);
    d[ST_I_JM1] = g11 * (1 - 0.5 * Q);
    d[ST_IP1_JP1] = -0.5 * g12;
    d[ST_IP1_JM1] = 0.5 * g12;
    d[ST_IM1_JP1] = 0.5 * g12;
    d[ST_IM1_JM1] = -0.5 * g12;
}

/* fillBlockInternalPointData, smooth.zig:923-992 */
static void fill_internal(orc_system *S)
{
    const mesh_t *M = &S->M;
    uint64_t row = 0;
    for (size_t b = 0; b < M->n_blocks; ++b) {
        const orc_block *blk = &M->blocks[b];
        const uint64_t nj = blk->nj;
        uint64_t pid = 0;
        pid += nj; row += nj;
        for (uint64_t i = 1; i + 1 < blk->ni; ++i) {
            row += 1; pid += 1;
            for (uint64_t j = 1; j + 1 < nj; ++j) {
                double st[9];
                stencil_init(&blk->xy[2 * (pid - nj)], &blk->xy[2 * (pid + nj)], &blk->xy[2 * (pid - 1)], &blk->xy[2 * (pid + 1)],
                             S->cf[2 * row], S->cf[2 * row + 1], st);
                double *v = &S->lhs_values[S->lhs_p[row]];
                v[0] = st[ST_IM1_JM1]; v[1] = st[ST_IM1_J]; v[2] = st[ST_IM1_JP1];
                v[3] = st[ST_I_JM1];   v[4] = st[ST_I_J];   v[5] = st[ST_I_JP1];
                v[6] = st[ST_IP1_JM1]; v[7] = st[ST_IP1_J]; v[8] = st[ST_IP1_JP1];
                S->rhs_x[row] = 0; S->rhs_y[row] = 0;
                pid += 1; row += 1;
            }
            row += 1; pid += 1;
        }
        row += nj;
    }
}

/* fillBlockConnectionData, smooth.zig:994-1105 */
static int fill_connections(orc_system *S)
{
    const mesh_t *M = &S->M;
    for (size_t c = 0; c < M->n_conns; ++c) {
        const orc_connection *cn = &M->conns[c];
        const double *p0 = M->blocks[cn->ranges[0].block].xy, *p1 = M->blocks[cn->ranges[1].block].xy;
        fill_it it; fill_it_init(&it, M, cn);
        int64_t loc[2];
        fill_it_next(&it, loc); it.count -= 1;                               /* limitToRangeInternalPoints, smooth.zig:1551-1554 */
        size_t pos[9]; int rc;
        if ((rc = connection_stencil_positions(cn, &it, pos))) return rc;
        while (fill_it_next(&it, loc)) {
            const uint64_t g0 = M->block_start[cn->ranges[0].block] + (uint64_t)loc[0];
            const size_t s = (size_t)S->lhs_p[g0];
            const double *im1_j = &p0[2 * (loc[0] - it.along[0])];
            const double *i_jm1 = &p0[2 * (loc[0] + it.inward[0])];
            const double *ip1_j = &p0[2 * (loc[0] + it.along[0])];
            double i_jp1[2] = { p1[2 * (loc[1] + it.inward[1])], p1[2 * (loc[1] + it.inward[1]) + 1] };
            double st[9];
            if (cn->has_periodicity) {
                i_jp1[0] = i_jp1[0] + (-cn->periodicity[0]); i_jp1[1] = i_jp1[1] + (-cn->periodicity[1]);   /* smooth.zig:1032 */
                stencil_init(im1_j, ip1_j, i_jm1, i_jp1, S->cf[2 * g0], S->cf[2 * g0 + 1], st);           /* :1040-1041 */
            } else {
                stencil_init(im1_j, ip1_j, i_jm1, i_jp1, S->cf[2 * g0 + 1], S->cf[2 * g0], st);           /* :1082-1083 (P,Q swapped) */
            }
            double *v = &S->lhs_values[s];
            v[pos[0]] = st[ST_IM1_JM1]; v[pos[1]] = st[ST_I_JM1]; v[pos[2]] = st[ST_IP1_JM1];
            v[pos[3]] = st[ST_IM1_J];   v[pos[4]] = st[ST_I_J];   v[pos[5]] = st[ST_IP1_J];
            v[pos[6]] = st[ST_IM1_JP1]; v[pos[7]] = st[ST_I_JP1]; v[pos[8]] = st[ST_IP1_JP1];
            if (cn->has_periodicity) {                                        /* smooth.zig:1060-1061 */
                S->rhs_x[g0] = cn->periodicity[0] * (st[ST_IM1_JP1] + st[ST_I_JP1] + st[ST_IP1_JP1]);
                S->rhs_y[g0] = cn->periodicity[1] * (st[ST_IM1_JP1] + st[ST_I_JP1] + st[ST_IP1_JP1]);
            }
        }
    }
    return ORC_OK;
}

/* fillXSpecific / fillYSpecific, smooth.zig:1115-1165 */
static void fill_specific(orc_system *S, int y_mode)
{
    const mesh_t *M = &S->M;
    for (size_t c = 0; c < M->n_bcs; ++c) {
        const orc_condition *bc = &M->bcs[c];
        int64_t base, inc, inward; range_walk(M, &bc->range, &base, &inc, &inward);
        const uint64_t len = range_len(&bc->range);
        for (uint64_t k = 0; k < len; ++k) {
            const uint64_t local = (uint64_t)(base + (int64_t)k * inc);
            if (S->kind[buffer_index(M, bc->range.block, local)] != KIND_SLIDING) continue;
            const size_t s = (size_t)S->lhs_p[M->block_start[bc->range.block] + local];
            if (y_mode) { S->lhs_values[s] = 1.0; S->lhs_values[s + 1] = -1.0; }
            else if (inward > 0) { S->lhs_values[s] = 1.0; S->lhs_values[s + 1] = 0.0; }
            else { S->lhs_values[s] = 0.0; S->lhs_values[s + 1] = 1.0; }
        }
    }
}

/* ================================================================================================
 * White wall control function -- wall_control_function.zig:56-474
 * ================================================================================================ */
static void white_blend_line(double *cf, uint64_t start, uint64_t nj, double p, double q)
{   /* wall_control_function.zig:104-111 (and the identical loops at :144-151, :186-193, :268-275, :309-319) */
    cf[2 * start] = p; cf[2 * start + 1] = q;
    for (uint64_t j = 1; j < nj; ++j) {
        const double factor = 1 - (double)j / ((double)nj - 1);
        cf[2 * (start + j)] = factor * p; cf[2 * (start + j) + 1] = factor * q;
    }
}
static void white_pq(double x_xi, double y_xi, double x_xi2, double y_xi2, double x_eta, double y_eta, double x_eta2, double y_eta2, double *p, double *q)
{   /* eq. 6.10, wall_control_function.zig:97-102 */
    const double g11 = x_xi * x_xi + y_xi * y_xi;
    const double g22 = x_eta * x_eta + y_eta * y_eta;
    *p = -(x_xi * x_xi2 + y_xi * y_xi2) / g11 - (x_xi * x_eta2 + y_xi * y_eta2) / g22;
    *q = -(x_eta * x_eta2 + y_eta * y_eta2) / g22 - (x_eta * x_xi2 + y_eta * y_xi2) / g11;
}
static int white_check_connection0(const mesh_t *M)
{   /* wall_control_function.zig:212-217, 403-408 */
    if (M->n_blocks < 2 || M->n_conns < 1) FAIL(ORC_ERR_UNSUPPORTED, "white control function needs blocks 0,1 and connection 0 (wall_control_function.zig:72,204)");
    const orc_connection *cn = &M->conns[0];
    if (!(cn->ranges[0].block == 0 && cn->ranges[0].start == 0 && cn->ranges[0].side == SIDE_J_MIN &&
          cn->ranges[1].block == 1 && cn->ranges[1].start == 0 && cn->ranges[1].side == SIDE_J_MIN && !cn->has_periodicity))
        FAIL(ORC_ERR_UNSUPPORTED, "white control function: connection 0 must be block0:j_min[0..] <-> block1:j_min[0..], non periodic");
    if (M->blocks[0].nj < 3 || M->blocks[1].nj < 3 || M->blocks[0].ni < 3 || M->blocks[1].ni < 3) FAIL(ORC_ERR_UNSUPPORTED, "white: blocks too small");
    return ORC_OK;
}

/* White.initControlFunction, wall_control_function.zig:70-280 */
static int white_init(orc_system *S)
{
    const mesh_t *M = &S->M;
    int rc = white_check_connection0(M); if (rc) return rc;
    double *cf = S->cf;
    uint64_t start = 0;
    for (size_t b = 0; b < 2; ++b) {
        const double *d = M->blocks[b].xy; const uint64_t ni = M->blocks[b].ni, nj = M->blocks[b].nj;
        uint64_t l = 0; double p, q;
#define X(k) d[2 * (k)]
#define Y(k) d[2 * (k) + 1]
        {   /* corner 0,0: forward differences, :78-112 */
            const double x_xi = -X(l) + X(l + nj), y_xi = -Y(l) + Y(l + nj);
            const double x_xi2 = X(l) - 2 * X(l + nj) + X(l + 2 * nj), y_xi2 = Y(l) - 2 * Y(l + nj) + Y(l + 2 * nj);
            const double x_eta = -X(l) + X(l + 1), y_eta = -Y(l) + Y(l + 1);
            const double x_eta2 = X(l) - 2 * X(l + 1) + X(l + 2), y_eta2 = Y(l) - 2 * Y(l + 1) + Y(l + 2);
            white_pq(x_xi, y_xi, x_xi2, y_xi2, x_eta, y_eta, x_eta2, y_eta2, &p, &q);
            white_blend_line(cf, start + l, nj, p, q); l += nj;
        }
        for (uint64_t i = 1; i + 1 < ni; ++i) {   /* :115-153 */
            const double x_xi = 0.5 * (X(l + nj) - X(l - nj)), y_xi = 0.5 * (Y(l + nj) - Y(l - nj));
            const double x_xi2 = X(l + nj) - 2 * X(l) + X(l - nj), y_xi2 = Y(l + nj) - 2 * Y(l) + Y(l - nj);
            const double x_eta = -X(l) + X(l + 1), y_eta = -Y(l) + Y(l + 1);
            const double x_eta2 = X(l) - 2 * X(l + 1) + X(l + 2), y_eta2 = Y(l) - 2 * Y(l + 1) + Y(l + 2);
            white_pq(x_xi, y_xi, x_xi2, y_xi2, x_eta, y_eta, x_eta2, y_eta2, &p, &q);
            white_blend_line(cf, start + l, nj, p, q); l += nj;
        }
        {   /* corner n,0: backward differences, :155-194 */
            const double x_xi = X(l) - X(l - nj), y_xi = Y(l) - Y(l - nj);
            const double x_xi2 = X(l) - 2 * X(l - nj) + X(l - 2 * nj), y_xi2 = Y(l) - 2 * Y(l - nj) + Y(l - 2 * nj);
            const double x_eta = -X(l) + X(l + 1), y_eta = -Y(l) + Y(l + 1);
            const double x_eta2 = X(l) - 2 * X(l + 1) + X(l + 2), y_eta2 = Y(l) - 2 * Y(l + 1) + Y(l + 2);
            white_pq(x_xi, y_xi, x_xi2, y_xi2, x_eta, y_eta, x_eta2, y_eta2, &p, &q);
            white_blend_line(cf, start + l, nj, p, q);
        }
#undef X
#undef Y
        start += ni * nj;
    }
    {   /* connection 0 override, :203-279 */
        const orc_connection *cn = &M->conns[0];
        fill_it it; fill_it_init(&it, M, cn);
        const double *d0 = M->blocks[0].xy, *d1 = M->blocks[1].xy;
        const int64_t p0 = it.pos[0], p1 = it.pos[1];
        const double x_i_j = d0[2 * p0], y_i_j = d0[2 * p0 + 1];
        const double x_ip1_j = d0[2 * (p0 + it.inward[0])], y_ip1_j = d0[2 * (p0 + it.inward[0]) + 1];
        const double x_im1_j = d1[2 * (p1 + it.inward[1])], y_im1_j = d1[2 * (p1 + it.inward[1]) + 1];
        const double x_i_jp1 = d0[2 * (p0 + it.along[0])], y_i_jp1 = d0[2 * (p0 + it.along[0]) + 1];
        const double x_i_jp2 = d0[2 * (p0 + 2 * it.along[0])], y_i_jp2 = d0[2 * (p0 + 2 * it.along[0]) + 1];
        const double x_xi = 0.5 * (x_ip1_j - x_im1_j), y_xi = 0.5 * (y_ip1_j - y_im1_j);
        const double x_xi2 = x_ip1_j - 2 * x_i_j + x_im1_j, y_xi2 = y_ip1_j - 2 * y_i_j + y_im1_j;
        const double x_eta = -x_i_j + x_i_jp1, y_eta = -y_i_j + y_i_jp1;
        const double x_eta2 = x_i_j - 2 * x_i_jp1 + x_i_jp2, y_eta2 = y_i_j - 2 * y_i_jp1 + y_i_jp2;
        double p, q; white_pq(x_xi, y_xi, x_xi2, y_xi2, x_eta, y_eta, x_eta2, y_eta2, &p, &q);
        white_blend_line(cf, 0, M->blocks[0].nj, p, q);
    }
    return ORC_OK;
}

/* White.computeUpdate, wall_control_function.zig:282-320 */
static void white_compute_update(const orc_options *o, double *cf, uint64_t id, uint64_t nj, double x_xi, double y_xi, double x_eta, double y_eta)
{
    const double g11 = x_xi * x_xi + y_xi * y_xi;
    const double g12 = x_xi * x_eta + y_xi * y_eta;
    const double g22 = x_eta * x_eta + y_eta * y_eta;
    const double ds = sqrt(g22);
    const double theta = acos(g12 / sqrt(g11 * g22));
    const double delta_ds = o->ds_target - ds;
    const double delta_theta = o->theta_target - theta;
    const double delta_p = -atan2(delta_theta, o->theta_target);
    const double delta_q = atan2(delta_ds, o->ds_target);
    double p = cf[2 * id], q = cf[2 * id + 1];
    p += 0.1 * delta_p; q += 0.1 * delta_q;
    white_blend_line(cf, id, nj, p, q);
}

/* White.update, wall_control_function.zig:322-473 */
static void white_update(orc_system *S)
{
    const mesh_t *M = &S->M; const orc_options *o = &S->opt; double *cf = S->cf;
    uint64_t start = 0;
    for (size_t b = 0; b < 2; ++b) {
        const double *d = M->blocks[b].xy; const uint64_t ni = M->blocks[b].ni, nj = M->blocks[b].nj;
        uint64_t l = 0;
#define X(k) d[2 * (k)]
#define Y(k) d[2 * (k) + 1]
        white_compute_update(o, cf, start + l, nj, -X(l) + X(l + nj), -Y(l) + Y(l + nj), -X(l) + X(l + 1), -Y(l) + Y(l + 1)); l += nj;
        for (uint64_t i = 1; i + 1 < ni; ++i) {
            white_compute_update(o, cf, start + l, nj, 0.5 * (X(l + nj) - X(l - nj)), 0.5 * (Y(l + nj) - Y(l - nj)), -X(l) + X(l + 1), -Y(l) + Y(l + 1)); l += nj;
        }
        white_compute_update(o, cf, start + l, nj, X(l) - X(l - nj), Y(l) - Y(l - nj), -X(l) + X(l + 1), -Y(l) + Y(l + 1));
#undef X
#undef Y
        start += ni * nj;
    }
    {   /* connection 0, :394-472 */
        const orc_connection *cn = &M->conns[0];
        fill_it it; fill_it_init(&it, M, cn);
        const double *d0 = M->blocks[0].xy, *d1 = M->blocks[1].xy;
        const int64_t p0 = it.pos[0], p1 = it.pos[1];
        const double x_i_j = d0[2 * p0], y_i_j = d0[2 * p0 + 1];
        const double x_ip1_j = d0[2 * (p0 + it.inward[0])], y_ip1_j = d0[2 * (p0 + it.inward[0]) + 1];
        const double x_im1_j = d1[2 * (p1 + it.inward[1])], y_im1_j = d1[2 * (p1 + it.inward[1]) + 1];
        const double x_i_jp1 = d0[2 * (p0 + it.along[0])], y_i_jp1 = d0[2 * (p0 + it.along[0]) + 1];
        const double x_xi = -0.5 * (x_ip1_j - x_im1_j), y_xi = -0.5 * (y_ip1_j - y_im1_j);   /* sign flip, :429-431 */
        const double x_eta = -x_i_j + x_i_jp1, y_eta = -y_i_j + y_i_jp1;
        /* same arithmetic as computeUpdate, applied to control_function[0] which the block loop above has
         * already advanced once (:450) */
        white_compute_update(o, cf, 0, M->blocks[0].nj, x_xi, y_xi, x_eta, y_eta);
    }
}

/* ================================================================================================
 * Krylov solvers -- GMRES.zig, BiCGStab.zig (shared helpers are duplicated verbatim there)
 * ================================================================================================ */
static double dotp(const double *a, const double *b, uint64_t n) { double s = 0.0; for (uint64_t i = 0; i < n; ++i) s += a[i] * b[i]; return s; } /* GMRES.zig:526-532 */
static double norm2(const double *a, uint64_t n) { return sqrt(dotp(a, a, n)); }                                                                /* GMRES.zig:534-536 */

static void mat_vec(orc_system *S, const double *x, double *out)   /* GMRES.zig:477-488 */
{
    const uint64_t dof = S->M.dof;
    for (uint64_t row = 0; row < dof; ++row) {
        double sum = 0.0;
        for (int32_t k = S->lhs_p[row]; k < S->lhs_p[row + 1]; ++k) sum += S->lhs_values[k] * x[S->lhs_i[k]];
        out[row] = sum;
    }
    S->stats.matvecs += 1;
}

static void update_diag_inv(orc_system *S, double *diag_inv)       /* GMRES.zig:176-196 */
{
    const uint64_t dof = S->M.dof;
    for (uint64_t row = 0; row < dof; ++row) {
        double diag = 0.0;
        for (int32_t k = S->lhs_p[row]; k < S->lhs_p[row + 1]; ++k) if ((uint64_t)S->lhs_i[k] == row) { diag = S->lhs_values[k]; break; }
        diag_inv[row] = (diag == 0.0) ? 1.0 : 1.0 / diag;
    }
}

static int update_ilu0(orc_system *S)                              /* GMRES.zig:199-298 */
{
    const uint64_t dof = S->M.dof;
    if (!S->ilu) {
        S->ilu = (double *)malloc(S->nnz * sizeof(double));
        S->diag_pos = (int32_t *)malloc(dof * sizeof(int32_t));
        S->marker = (int32_t *)malloc(dof * sizeof(int32_t));
        if (!S->ilu || !S->diag_pos || !S->marker) FAIL(ORC_ERR_NOMEM, "out of memory");
    }
    double *lu = S->ilu; int32_t *diag_pos = S->diag_pos, *marker = S->marker;
    memcpy(lu, S->lhs_values, S->nnz * sizeof(double));
    for (uint64_t r = 0; r < dof; ++r) { diag_pos[r] = -1; marker[r] = -1; }
    for (uint64_t row = 0; row < dof; ++row)
        for (int32_t k = S->lhs_p[row]; k < S->lhs_p[row + 1]; ++k) if ((uint64_t)S->lhs_i[k] == row) { diag_pos[row] = k; break; }
    for (uint64_t row = 0; row < dof; ++row) {
        const int32_t start = S->lhs_p[row], end = S->lhs_p[row + 1];
        for (int32_t k = start; k < end; ++k) marker[S->lhs_i[k]] = k;
        for (int32_t k = start; k < end; ++k) {
            const uint64_t col = (uint64_t)S->lhs_i[k];
            if (col >= row) continue;
            double diag = 1.0;
            if (diag_pos[col] >= 0) { diag = lu[diag_pos[col]]; if (diag == 0.0) diag = 1.0; }
            const double lij = lu[k] / diag;
            lu[k] = lij;
            for (int32_t q = S->lhs_p[col]; q < S->lhs_p[col + 1]; ++q) {
                const uint64_t col_k = (uint64_t)S->lhs_i[q];
                if (col_k <= col) continue;
                const int32_t pos = marker[col_k];
                if (pos >= 0) lu[pos] -= lij * lu[q];
            }
        }
        for (int32_t k = start; k < end; ++k) marker[S->lhs_i[k]] = -1;
    }
    return ORC_OK;
}

static void apply_ilu0(orc_system *S, const double *rhs, double *out)   /* GMRES.zig:437-475 */
{
    const uint64_t dof = S->M.dof; const double *lu = S->ilu;
    for (uint64_t row = 0; row < dof; ++row) {
        double sum = rhs[row];
        for (int32_t k = S->lhs_p[row]; k < S->lhs_p[row + 1]; ++k) { const uint64_t col = (uint64_t)S->lhs_i[k]; if (col < row) sum -= lu[k] * out[col]; }
        out[row] = sum;
    }
    for (uint64_t row = dof; row-- > 0;) {
        double sum = out[row];
        for (int32_t k = S->lhs_p[row]; k < S->lhs_p[row + 1]; ++k) { const uint64_t col = (uint64_t)S->lhs_i[k]; if (col > row) sum -= lu[k] * out[col]; }
        double diag = 1.0;
        if (S->diag_pos[row] >= 0) { diag = lu[S->diag_pos[row]]; if (diag == 0.0) diag = 1.0; }
        out[row] = sum / diag;
    }
}

static void apply_precond(orc_system *S, const double *rhs, double *out, const double *diag_inv)
{
    if (S->opt.preconditioner == PRECOND_DIAGONAL) { for (uint64_t i = 0; i < S->M.dof; ++i) out[i] = rhs[i] * diag_inv[i]; }
    else apply_ilu0(S, rhs, out);
    S->stats.precond_applies += 1;
}
static int precondition(orc_system *S, double *diag_inv)
{
    if (S->opt.preconditioner == PRECOND_DIAGONAL) { update_diag_inv(S, diag_inv); return ORC_OK; }
    return update_ilu0(S);
}

static void compute_givens(double a, double b, double *c, double *s, double *r)   /* GMRES.zig:510-524 */
{
    if (b == 0.0) { *c = 1.0; *s = 0.0; *r = a; return; }
    if (fabs(b) > fabs(a)) { const double t = a / b; const double sn = 1.0 / sqrt(1.0 + t * t); *c = sn * t; *s = sn; *r = b / sn; return; }
    const double t = b / a; const double cs = 1.0 / sqrt(1.0 + t * t); *c = cs; *s = cs * t; *r = a / cs;
}

/* GMRESSolver.solveSystem, GMRES.zig:300-423 */
static void gmres_solve(orc_system *S, const double *rhs, double *x, size_t restart, double *v, double *h, double *cs, double *sn, double *g,
                        double *r, double *w, double *z, const double *diag_inv)
{
    const uint64_t dof = S->M.dof;
    const double breakdown_eps = 1e-30;
    if (restart == 0) return;
    const double tol = fmax(S->opt.atol, S->opt.rtol * norm2(rhs, dof));
    uint64_t iter_total = 0;
#define H(row, col) h[(row) + (restart + 1) * (col)]
    while (iter_total < S->opt.max_iters) {
        mat_vec(S, x, w);
        for (uint64_t i = 0; i < dof; ++i) r[i] = rhs[i] - w[i];
        apply_precond(S, r, z, diag_inv);
        const double beta = norm2(z, dof);
        if (beta <= tol) return;
        for (uint64_t i = 0; i < dof; ++i) v[i] = z[i] / beta;
        memset(h, 0, (restart + 1) * restart * sizeof(double));
        memset(cs, 0, restart * sizeof(double)); memset(sn, 0, restart * sizeof(double)); memset(g, 0, (restart + 1) * sizeof(double));
        g[0] = beta;
        size_t cols_used = 0; int converged = 0; double resid = beta;
        for (size_t j = 0; j < restart && iter_total < S->opt.max_iters; ++j) {
            mat_vec(S, v + j * dof, w);
            apply_precond(S, w, z, diag_inv);
            for (size_t i = 0; i < j + 1; ++i) {
                const double *vi = v + i * dof;
                const double h_ij = dotp(z, vi, dof);
                H(i, j) = h_ij;
                for (uint64_t k = 0; k < dof; ++k) z[k] -= h_ij * vi[k];
            }
            const double h_next = norm2(z, dof);
            H(j + 1, j) = h_next;
            if (h_next > breakdown_eps) { double *vn = v + (j + 1) * dof; for (uint64_t k = 0; k < dof; ++k) vn[k] = z[k] / h_next; }
            for (size_t i = 0; i < j; ++i) {
                const double h_i = H(i, j), h_ip1 = H(i + 1, j);
                const double temp = cs[i] * h_i + sn[i] * h_ip1;
                H(i + 1, j) = -sn[i] * h_i + cs[i] * h_ip1;
                H(i, j) = temp;
            }
            double c, s, rr; compute_givens(H(j, j), H(j + 1, j), &c, &s, &rr);
            cs[j] = c; sn[j] = s; H(j, j) = rr; H(j + 1, j) = 0.0;
            const double g_j = g[j], g_jp1 = g[j + 1];
            g[j] = c * g_j + s * g_jp1;
            g[j + 1] = -s * g_j + c * g_jp1;
            resid = fabs(g[j + 1]);
            iter_total += 1; cols_used = j + 1;
            if (resid <= tol) { converged = 1; break; }
        }
        S->stats.krylov_iterations += cols_used;
        if (cols_used == 0) break;
        double *y = w;
        for (size_t idx = cols_used; idx-- > 0;) {
            double sum = g[idx];
            for (size_t k = idx + 1; k < cols_used; ++k) sum -= H(idx, k) * y[k];
            const double h_ii = H(idx, idx);
            if (h_ii == 0.0) break;
            y[idx] = sum / h_ii;
        }
        for (size_t i = 0; i < cols_used; ++i) { const double *vi = v + i * dof; const double yi = y[i]; for (uint64_t k = 0; k < dof; ++k) x[k] += yi * vi[k]; }
        if (converged) return;
        if (resid <= tol) return;
    }
#undef H
    S->stats.not_converged += 1;   /* log.warn, GMRES.zig:422 */
}

/* BiCGStabSolver.solveSystem, BiCGStab.zig:279-370 */
static void bicgstab_solve(orc_system *S, const double *rhs, double *x, double *r, double *r_hat, double *p, double *v, double *s, double *t, double *precond, const double *diag_inv)
{
    const uint64_t dof = S->M.dof;
    const double breakdown_eps = 1e-30;
    mat_vec(S, x, v);
    for (uint64_t i = 0; i < dof; ++i) r[i] = rhs[i] - v[i];
    memcpy(r_hat, r, dof * sizeof(double));
    const double norm_b = norm2(rhs, dof);
    double norm_r = norm2(r, dof);
    const double tol = fmax(S->opt.atol, S->opt.rtol * norm_b);
    if (norm_r <= tol) return;
    memset(p, 0, dof * sizeof(double)); memset(v, 0, dof * sizeof(double));
    double rho_old = 1.0, alpha = 1.0, omega = 1.0;
    for (uint64_t iter = 0; iter < S->opt.max_iters; ++iter) {
        const double rho_new = dotp(r_hat, r, dof);
        if (fabs(rho_new) < breakdown_eps) break;
        const double beta = (rho_new / rho_old) * (alpha / omega);
        for (uint64_t i = 0; i < dof; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
        apply_precond(S, p, precond, diag_inv);
        mat_vec(S, precond, v);
        const double denom = dotp(r_hat, v, dof);
        if (fabs(denom) < breakdown_eps) break;
        alpha = rho_new / denom;
        for (uint64_t i = 0; i < dof; ++i) s[i] = r[i] - alpha * v[i];
        for (uint64_t i = 0; i < dof; ++i) x[i] += alpha * precond[i];
        S->stats.krylov_iterations += 1;
        const double norm_s = norm2(s, dof);
        if (norm_s <= tol) return;
        apply_precond(S, s, precond, diag_inv);
        mat_vec(S, precond, t);
        const double t_dot_t = dotp(t, t, dof);
        if (fabs(t_dot_t) < breakdown_eps) break;
        omega = dotp(t, s, dof) / t_dot_t;
        if (fabs(omega) < breakdown_eps) break;
        for (uint64_t i = 0; i < dof; ++i) x[i] += omega * precond[i];
        for (uint64_t i = 0; i < dof; ++i) r[i] = s[i] - omega * t[i];
        norm_r = norm2(r, dof);
        if (norm_r <= tol) return;
        rho_old = rho_new;
    }
    S->stats.not_converged += 1;   /* log.warn, BiCGStab.zig:368-369 */
}

/* GMRESSolver.solve / BiCGStabSolver.solve, GMRES.zig:76-90 / BiCGStab.zig:71-85 */
static int solver_solve(orc_system *S)
{
    const mesh_t *M = &S->M; const uint64_t dof = M->dof;
    if (!S->seeded) {   /* seedInitialGuess, GMRES.zig:157-174 */
        for (size_t b = 0; b < M->n_blocks; ++b) {
            const uint64_t n = M->blocks[b].ni * M->blocks[b].nj, off = M->block_start[b];
            for (uint64_t k = 0; k < n; ++k) { S->x_new[off + k] = M->blocks[b].xy[2 * k]; S->y_new[off + k] = M->blocks[b].xy[2 * k + 1]; }
        }
        S->seeded = 1;
    }
    int rc;
    if (S->opt.solver == SOLVER_GMRES) {
        const size_t restart = (size_t)((uint64_t)S->opt.restart < dof ? (uint64_t)S->opt.restart : dof);   /* GMRES.zig:93 */
        const size_t total = (restart + 1) * dof + (restart + 1) * restart + restart + restart + (restart + 1) + 4 * dof;
        if (S->work_len != total) { free(S->work); S->work = (double *)malloc(total * sizeof(double)); S->work_len = total; if (!S->work) FAIL(ORC_ERR_NOMEM, "out of memory (GMRES workspace)"); }
        double *v = S->work, *h = v + (restart + 1) * dof, *cs = h + (restart + 1) * restart, *sn = cs + restart, *g = sn + restart;
        double *r = g + restart + 1, *w = r + dof, *z = w + dof, *diag_inv = z + dof;
        fill_specific(S, 0); if ((rc = precondition(S, diag_inv))) return rc;
        gmres_solve(S, S->rhs_x, S->x_new, restart, v, h, cs, sn, g, r, w, z, diag_inv);
        fill_specific(S, 1); if ((rc = precondition(S, diag_inv))) return rc;
        gmres_solve(S, S->rhs_y, S->y_new, restart, v, h, cs, sn, g, r, w, z, diag_inv);
    } else {
        const size_t total = 8 * dof;                                                                       /* BiCGStab.zig:87-89 */
        if (S->work_len != total) { free(S->work); S->work = (double *)malloc(total * sizeof(double)); S->work_len = total; if (!S->work) FAIL(ORC_ERR_NOMEM, "out of memory (BiCGStab workspace)"); }
        double *diag_inv = S->work, *r = diag_inv + dof, *r_hat = r + dof, *p = r_hat + dof, *v = p + dof, *s = v + dof, *t = s + dof, *pc = t + dof;
        fill_specific(S, 0); if ((rc = precondition(S, diag_inv))) return rc;
        bicgstab_solve(S, S->rhs_x, S->x_new, r, r_hat, p, v, s, t, pc, diag_inv);
        fill_specific(S, 1); if ((rc = precondition(S, diag_inv))) return rc;
        bicgstab_solve(S, S->rhs_y, S->y_new, r, r_hat, p, v, s, t, pc, diag_inv);
    }
    return ORC_OK;
}

/* ================================================================================================
 * Public system API (lets the tests inspect kinds / CSR / control function step by step)
 * ================================================================================================ */
void orc_options_default(orc_options *o)
{
    memset(o, 0, sizeof *o);
    o->solver = SOLVER_GMRES; o->preconditioner = PRECOND_ILU0; o->control_function = CF_LAPLACE;
    o->restart = 30; o->max_iters = 1000; o->rtol = 1e-6; o->atol = 1e-8;
    o->ds_target = 1e-6; o->theta_target = 0.5 * 3.14159265358979323846;
}

void orc_system_destroy(orc_system *S)
{
    if (!S) return;
    mesh_free(&S->M);
    free(S->kind); free(S->lps); free(S->lhs_p); free(S->lhs_i); free(S->lhs_values);
    free(S->rhs_x); free(S->rhs_y); free(S->x_new); free(S->y_new); free(S->cf);
    free(S->ilu); free(S->diag_pos); free(S->marker); free(S->work);
    free(S);
}

/* RowCompressedMatrixSystem2d.init, smooth.zig:309-385.  The block/connection/condition arrays are borrowed
 * and must outlive the system; block coordinates are updated in place by orc_system_iterate. */
int orc_system_create(const orc_block *blocks, size_t nb, const orc_connection *conns, size_t nc, const orc_condition *bcs, size_t nbc,
                      const orc_options *opt, orc_system **out)
{
    *out = NULL;
    orc_system *S = (orc_system *)calloc(1, sizeof *S);
    if (!S) FAIL(ORC_ERR_NOMEM, "out of memory");
    S->opt = *opt;
    int rc = mesh_init(&S->M, blocks, nb, conns, nc, bcs, nbc);
    if (rc == ORC_OK) rc = connection_data_check(&S->M);
    if (rc == ORC_OK && S->M.dof >= (uint64_t)INT32_MAX / 9) { rc = ORC_ERR_UNSUPPORTED; snprintf(g_err, sizeof g_err, "mesh too large for the reference's 32-bit CSR indices (smooth.zig:280-287)"); }
    if (rc == ORC_OK) rc = init_laplacian_points(&S->M, &S->lps, &S->n_lps);
    if (rc == ORC_OK) rc = init_kinds(S);
    if (rc != ORC_OK) { orc_system_destroy(S); return rc; }
    const uint64_t dof = S->M.dof;
    S->lhs_p = (int32_t *)calloc(dof + 1, sizeof(int32_t));
    S->rhs_x = (double *)calloc(dof, sizeof(double)); S->rhs_y = (double *)calloc(dof, sizeof(double));
    S->x_new = (double *)calloc(dof, sizeof(double)); S->y_new = (double *)calloc(dof, sizeof(double));
    S->cf = (double *)calloc(2 * dof, sizeof(double));
    if (!S->lhs_p || !S->rhs_x || !S->rhs_y || !S->x_new || !S->y_new || !S->cf) { orc_system_destroy(S); FAIL(ORC_ERR_NOMEM, "out of memory"); }
    if (S->opt.control_function == CF_WHITE) rc = white_init(S);               /* ControlFunction.init, wall_control_function.zig:27-42 */
    if (rc == ORC_OK) rc = init_nonzero_entries(S);
    if (rc == ORC_OK) {
        S->lhs_values = (double *)calloc(S->nnz ? S->nnz : 1, sizeof(double));
        if (!S->lhs_values) { rc = ORC_ERR_NOMEM; snprintf(g_err, sizeof g_err, "out of memory"); }
    }
    if (rc != ORC_OK) { orc_system_destroy(S); return rc; }
    init_boundary_data(S);
    *out = S;
    return ORC_OK;
}

/* system.fill(n), smooth.zig:1107-1113 */
int orc_system_fill(orc_system *S, uint64_t iteration)
{
    if (iteration > 0 && S->opt.control_function == CF_WHITE) white_update(S);
    fill_internal(S);
    return fill_connections(S);
}
void orc_system_fill_specific(orc_system *S, int y_mode) { fill_specific(S, y_mode); }

/* one pass of the loop body of smooth.mesh, smooth.zig:104-154 */
int orc_system_iterate(orc_system *S, uint64_t n)
{
    const mesh_t *M = &S->M;
    double t0 = now_s();
    int rc = orc_system_fill(S, n); if (rc) return rc;
    double t1 = now_s();
    rc = solver_solve(S); if (rc) return rc;
    double t2 = now_s();
    double sx = 0.0, sy = 0.0, mx = 0.0;                                         /* smooth.zig:112-134 */
    for (size_t b = 0; b < M->n_blocks; ++b) {
        const uint64_t cnt = M->blocks[b].ni * M->blocks[b].nj, off = M->block_start[b];
        for (uint64_t k = 0; k < cnt; ++k) {
            const double dx = M->blocks[b].xy[2 * k] - S->x_new[off + k];
            const double dy = M->blocks[b].xy[2 * k + 1] - S->y_new[off + k];
            sx += dx * dx; sy += dy * dy;
            if (fabs(dx) > mx) mx = fabs(dx);
            if (fabs(dy) > mx) mx = fabs(dy);
        }
    }
    S->stats.last_sumsq_x = sx; S->stats.last_sumsq_y = sy; S->stats.last_residual = (sx + sy) * (sx + sy); S->stats.last_max_update = mx;
    for (size_t b = 0; b < M->n_blocks; ++b) {                                   /* smooth.zig:139-153 */
        const uint64_t cnt = M->blocks[b].ni * M->blocks[b].nj, off = M->block_start[b];
        for (uint64_t k = 0; k < cnt; ++k) { M->blocks[b].xy[2 * k] = S->x_new[off + k]; M->blocks[b].xy[2 * k + 1] = S->y_new[off + k]; }
    }
    S->stats.outer_iterations += 1;
    S->stats.seconds_fill += t1 - t0; S->stats.seconds_solve += t2 - t1; S->stats.seconds_total += now_s() - t0;
    return ORC_OK;
}

/* smoothing.smooth.mesh, smooth.zig:74-166 */
int orc_smooth_mesh(orc_block *blocks, size_t nb, const orc_connection *conns, size_t nc, const orc_condition *bcs, size_t nbc,
                    uint64_t iterations, const orc_options *opt, orc_stats *stats)
{
    orc_system *S = NULL;
    const double t0 = now_s();
    int rc = orc_system_create(blocks, nb, conns, nc, bcs, nbc, opt, &S);
    if (rc) return rc;
    for (uint64_t n = 0; n < iterations && rc == ORC_OK; ++n) rc = orc_system_iterate(S, n);
    S->stats.seconds_total = now_s() - t0;
    if (stats) *stats = S->stats;
    orc_system_destroy(S);
    return rc;
}

/* ---- inspection accessors ---- */
uint64_t orc_system_dof(const orc_system *S) { return S->M.dof; }
uint64_t orc_system_nnz(const orc_system *S) { return S->nnz; }
uint64_t orc_system_n_boundary(const orc_system *S) { return S->M.n_boundary; }
const int32_t *orc_system_lhs_p(const orc_system *S) { return S->lhs_p; }
const int32_t *orc_system_lhs_i(const orc_system *S) { return S->lhs_i; }
const double *orc_system_lhs_values(const orc_system *S) { return S->lhs_values; }
const double *orc_system_rhs_x(const orc_system *S) { return S->rhs_x; }
const double *orc_system_rhs_y(const orc_system *S) { return S->rhs_y; }
const double *orc_system_control_function(const orc_system *S) { return S->cf; }
const uint8_t *orc_system_kinds(const orc_system *S) { return S->kind; }
void orc_system_stats(const orc_system *S, orc_stats *out) { *out = S->stats; }
uint64_t orc_system_n_junctions(const orc_system *S) { return S->n_lps; }
/* junction l: ids[4] (unused = UINT64_MAX), periodicities[8], stencil[6] (unused = -1), rhs[2] */
void orc_system_junction(const orc_system *S, uint64_t l, uint64_t ids[4], double per[8], int32_t stencil[6], double rhs[2])
{
    const laplacian_t *lp = &S->lps[l];
    for (size_t k = 0; k < 4; ++k) { ids[k] = k < lp->n_overlapping ? lp->overlapping[k].global_id : UINT64_MAX; per[2 * k] = k < lp->n_overlapping ? lp->overlapping[k].periodicity[0] : 0; per[2 * k + 1] = k < lp->n_overlapping ? lp->overlapping[k].periodicity[1] : 0; }
    for (size_t k = 0; k < 6; ++k) stencil[k] = k < lp->n_stencil ? lp->stencil_ids[k] : -1;
    rhs[0] = lp->rhs[0]; rhs[1] = lp->rhs[1];
}

/* Generic CSR solve with the restated Krylov solvers: lets the tests run the reference's own 5x5
 * known-answer system (umfpack.zig:71-97) and compare against scipy on assembled systems. */
int orc_csr_solve(uint64_t dof, const int32_t *lhs_p, const int32_t *lhs_i, const double *lhs_values, const double *rhs, double *x,
                  const orc_options *opt, orc_stats *stats)
{
    orc_system S; memset(&S, 0, sizeof S);
    S.opt = *opt; S.M.dof = dof; S.nnz = (size_t)lhs_p[dof];
    S.lhs_p = (int32_t *)lhs_p; S.lhs_i = (int32_t *)lhs_i; S.lhs_values = (double *)lhs_values;
    int rc = ORC_OK;
    if (opt->solver == SOLVER_GMRES) {
        const size_t restart = (size_t)((uint64_t)opt->restart < dof ? (uint64_t)opt->restart : dof);
        const size_t total = (restart + 1) * dof + (restart + 1) * restart + restart + restart + (restart + 1) + 4 * dof;
        double *wk = (double *)malloc(total * sizeof(double)); if (!wk) FAIL(ORC_ERR_NOMEM, "out of memory");
        double *v = wk, *h = v + (restart + 1) * dof, *cs = h + (restart + 1) * restart, *sn = cs + restart, *g = sn + restart;
        double *r = g + restart + 1, *w = r + dof, *z = w + dof, *diag_inv = z + dof;
        rc = precondition(&S, diag_inv);
        if (rc == ORC_OK) gmres_solve(&S, rhs, x, restart, v, h, cs, sn, g, r, w, z, diag_inv);
        free(wk);
    } else {
        double *wk = (double *)malloc(8 * dof * sizeof(double)); if (!wk) FAIL(ORC_ERR_NOMEM, "out of memory");
        double *diag_inv = wk, *r = diag_inv + dof, *r_hat = r + dof, *p = r_hat + dof, *v = p + dof, *s = v + dof, *t = s + dof, *pc = t + dof;
        rc = precondition(&S, diag_inv);
        if (rc == ORC_OK) bicgstab_solve(&S, rhs, x, r, r_hat, p, v, s, t, pc, diag_inv);
        free(wk);
    }
    free(S.ilu); free(S.diag_pos); free(S.marker);
    if (stats) *stats = S.stats;
    return rc;
}

/* Structured output: the AoS block (x,y interleaved, j fastest) as two SoA arrays with i fastest, the layout the
 * reference hands to cg_coord_write for CoordinateX / CoordinateY (src/core/cgns.zig:69-101) and, for the control
 * function, to cg_field_write (cgns.zig:110-161). */
void orc_block_to_soa(uint64_t ni, uint64_t nj, const double *xy, double *x, double *y)
{
    uint64_t idx = 0;
    for (uint64_t j = 0; j < nj; ++j)          /* cgns.zig:74-83 */
        for (uint64_t i = 0; i < ni; ++i) {
            x[idx] = xy[2 * (i * nj + j)];
            y[idx] = xy[2 * (i * nj + j) + 1];
            ++idx;
        }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Edge discretisation, the step right before the path: discrete.Edge.init = clustering.create + Curve.interpolate
 * (src/core/discrete.zig:17-31).  The spline itself (chord parameters, second derivatives, arc-length table) is an input
 * here: its construction (spline.zig:141-200, 87-110) is covered by the host-side mirror and the reference's known answers.
 * ------------------------------------------------------------------------------------------------------------------ */
/* clustering.zig:9-17 (uniform), :24-42 (roberts), :56-95 (single hyperbolic); kind 0 / 1 / 2 */
int orc_clustering(int kind, double alpha, double beta, double delta_s, uint64_t n, double *out)
{
    if (n < 2) FAIL(-1, "clustering needs at least 2 points");
    if (kind == 0) {
        for (uint64_t i = 0; i < n; ++i) out[i] = (double)i / (double)(n - 1);
    } else if (kind == 1) {
        for (uint64_t i = 0; i < n; ++i) {
            const double u = (double)i / (double)(n - 1);
            const double tmp = pow((beta + 1.0) / (beta - 1.0), (u - alpha) / (1.0 - alpha));
            const double tbar = (beta + 2.0 * alpha) * tmp - beta + 2.0 * alpha;
            out[i] = tbar / ((2.0 * alpha + 1.0) * (1.0 + tmp));
        }
    } else if (kind == 2) {
        const double n_1 = (double)(n - 1);
        const double b = n_1 * delta_s;
        const double y = 1.0 / b;
        double delta;
        if (y < 1.0) FAIL(-1, "single hyperbolic clustering needs (n-1)*delta_s <= 1 (clustering.zig:68-76)");
        if (y < 2.7829681) {
            const double y_bar = y - 1.0;
            delta = sqrt(6.0 * y_bar) * (1.0 + y_bar * (-0.15 + y_bar * (0.057321429 + y_bar * (-0.024907295 + y_bar * (0.0077424461 - 0.0010794123 * y_bar)))));
        } else {
            const double w = 1.0 / y - 0.028527431;
            const double v = log(y);
            delta = v + (1.0 + 1.0 / v) * log(2.0 * v) - 0.02041793 + w * (0.24902722 + w * (1.9496443 + w * (-2.6294547 + 8.56795911 * w)));
        }
        for (uint64_t i = 0; i < n; ++i) out[i] = (double)i / n_1;
        for (uint64_t i = 1; i < n; ++i) out[i] = 1.0 + tanh(0.5 * delta * (out[i] - 1.0)) / tanh(0.5 * delta);
    } else {
        FAIL(-1, "unknown clustering kind");
    }
    return 0;
}

/* geometry.zig:26-40 */
void orc_line_interpolate(const double start[2], const double end[2], const double *u, uint64_t n, double *out)
{
    const double dx = end[0] - start[0], dy = end[1] - start[1];
    for (uint64_t k = 0; k < n; ++k) {
        out[2 * k] = start[0] + u[k] * dx;
        out[2 * k + 1] = start[1] + u[k] * dy;
    }
}

/* FittingSpline.interpolate (spline.zig:74-81) = eval(paramAtArcFraction(u)) (spline.zig:112-139, 202-222); the arc table
 * has n_samples entries at the uniform parameters i / (n_samples - 1) (spline.zig:87-110, 201 in the reference) */
void orc_spline_interpolate(uint64_t m, const double *params, const double *points, const double *zx, const double *zy,
                            uint64_t n_samples, const double *sample_arc, double total_length, const double *u, uint64_t n, double *out)
{
    for (uint64_t k = 0; k < n; ++k) {
        double param = 0.0;
        if (total_length != 0.0) {
            const double target = u[k] < 0.0 ? 0.0 : (u[k] > 1.0 ? 1.0 : u[k]);
            uint64_t lo = 0, hi = n_samples - 1;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) / 2;
                if (sample_arc[mid] < target) lo = mid + 1; else hi = mid;
            }
            if (lo == 0) {
                param = 0.0;
            } else {
                const double a0 = sample_arc[lo - 1], a1 = sample_arc[lo];
                const double p0 = (double)(lo - 1) / (double)(n_samples - 1), p1 = (double)lo / (double)(n_samples - 1);
                const double t = a1 > a0 ? (target - a0) / (a1 - a0) : 0.0;
                param = p0 + t * (p1 - p0);
            }
        }
        const double uu = param < 0.0 ? 0.0 : (param > 1.0 ? 1.0 : param);
        uint64_t idx = 0;
        while (idx + 1 < m && params[idx + 1] < uu) ++idx;
        if (idx >= m - 1) idx = m - 2;
        const double h = params[idx + 1] - params[idx];
        const double a = (params[idx + 1] - uu) / h, b = (uu - params[idx]) / h;
        const double *z[2] = {zx, zy};
        for (int d = 0; d < 2; ++d) {
            const double y0 = points[2 * idx + d], y1 = points[2 * (idx + 1) + d];
            const double z0 = z[d][idx], z1 = z[d][idx + 1];
            out[2 * k + d] = a * y0 + b * y1 + ((a * a * a - a) * z0 + (b * b * b - b) * z1) * (h * h) / 6.0;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Viewer buffers (src/gui/lib.zig:227-318): all points of all blocks as f32 pairs with their x / y ranges
 * (createPointBuffer, :227-265; the maxima start at floatMin(f32), the smallest positive normal, as in the reference) and
 * the wireframe line indices (createWireframeElementBuffer, :267-318: per block first the segments along j, then along i).
 * The reference allocates 8x more indices than it writes; n_indices here is what it writes.
 * ------------------------------------------------------------------------------------------------------------------ */
#include <float.h>
void orc_viewer_points(const orc_block *blocks, size_t nb, float *points, float range_x[2], float range_y[2])
{
    range_x[0] = FLT_MAX; range_x[1] = FLT_MIN; range_y[0] = FLT_MAX; range_y[1] = FLT_MIN;
    size_t id = 0;
    for (size_t b = 0; b < nb; ++b)
        for (uint64_t k = 0; k < blocks[b].ni * blocks[b].nj; ++k) {
            points[id] = (float)blocks[b].xy[2 * k];
            if (points[id] < range_x[0]) range_x[0] = points[id];
            if (points[id] > range_x[1]) range_x[1] = points[id];
            points[id + 1] = (float)blocks[b].xy[2 * k + 1];
            if (points[id + 1] < range_y[0]) range_y[0] = points[id + 1];
            if (points[id + 1] > range_y[1]) range_y[1] = points[id + 1];
            id += 2;
        }
}
uint64_t orc_viewer_index_count(const orc_block *blocks, size_t nb)
{
    uint64_t total = 0;
    for (size_t b = 0; b < nb; ++b) total += blocks[b].ni * (blocks[b].nj - 1) * 2 + blocks[b].nj * (blocks[b].ni - 1) * 2;
    return total;
}
void orc_viewer_wireframe(const orc_block *blocks, size_t nb, uint32_t *idx)
{
    size_t at = 0;
    uint32_t point_offset = 0;
    for (size_t b = 0; b < nb; ++b) {
        const uint64_t ni = blocks[b].ni, nj = blocks[b].nj;
        uint32_t point_id = 0;
        for (uint64_t i = 0; i < ni; ++i) {            /* segments along j */
            for (uint64_t j = 1; j < nj; ++j) { idx[at] = point_offset + point_id; idx[at + 1] = point_offset + point_id + 1; point_id += 1; at += 2; }
            point_id += 1;
        }
        for (uint64_t j = 0; j < nj; ++j) {            /* segments along i */
            point_id = (uint32_t)j;
            for (uint64_t i = 1; i < ni; ++i) { idx[at] = point_offset + point_id; idx[at + 1] = point_offset + point_id + (uint32_t)nj; point_id += (uint32_t)nj; at += 2; }
        }
        point_offset += (uint32_t)(ni * nj);
    }
}
