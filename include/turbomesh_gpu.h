/*
 * turbomesh_gpu.h -- C ABI of the B200 (sm_100a) back-end for turbomesh's hot path:
 * 2D boundary-blended linear TFI + multi-block elliptic (Winslow) smoothing.
 *
 * This is the drop-in boundary: plain pointers and sizes only, no C++/torch types.
 * Every entry point names the reference interface it replaces (paths relative to the
 * turbomesh repository).  The Zig-side binding is in zig/cuda.zig, the integration
 * recipe in INTEGRATION.md.
 *
 * Conventions
 *  - all coordinates are fp64, interleaved x0,y0,x1,y1,... exactly like
 *    `Mat2d.data` / `Vec2d` (src/core/types.zig:16-27, 78-101): node (i,j) of a block of
 *    size (ni,nj) lives at doubles [2*(i*nj+j), 2*(i*nj+j)+1]  (j is the fastest index).
 *  - every function returns 0 (TM_OK) on success or a negative tm_status; the message of
 *    the last failure on the calling thread is available from tm_last_error().
 *  - the library never frees or keeps host pointers handed to it; results are written in
 *    place into the caller's block arrays (as smooth.zig:139-153 does).
 *  - there is NO CPU fallback: without a usable CUDA device every compute entry point
 *    returns TM_ERR_NO_DEVICE.
 */
#ifndef TURBOMESH_GPU_H
#define TURBOMESH_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TM_ABI_VERSION 1

typedef enum tm_status {
    TM_OK = 0,
    TM_ERR_INVALID_ARGUMENT = -1,
    TM_ERR_NO_DEVICE = -2,        /* no CUDA device / driver: there is no CPU fallback        */
    TM_ERR_CUDA = -3,             /* a CUDA runtime call failed                                */
    TM_ERR_TOPOLOGY = -4,         /* inconsistent mesh topology (what smooth.zig:220-275,
                                     627-651 asserts or panics on)                             */
    TM_ERR_UNSUPPORTED = -5,      /* topology the reference cannot express either
                                     (e.g. > 4 overlapping junction copies, smooth.zig:1221)   */
    TM_ERR_NOT_CONVERGED = -6,    /* only returned when opts.fail_on_no_convergence != 0       */
    TM_ERR_OUT_OF_MEMORY = -7
} tm_status;

/* boundary.zig:8-13 (enum order is part of the ABI) */
typedef enum tm_side { TM_SIDE_I_MIN = 0, TM_SIDE_I_MAX = 1, TM_SIDE_J_MIN = 2, TM_SIDE_J_MAX = 3 } tm_side;

/* boundary.zig:172-176 */
typedef enum tm_condition_kind { TM_BC_WALL = 0, TM_BC_INLET = 1, TM_BC_OUTLET = 2 } tm_condition_kind;

/* view of discrete.Block2d.points (discrete.zig:138-140, types.zig:78-81) */
typedef struct tm_block {
    uint64_t ni, nj; /* Mat2d.size                                                     */
    double *xy;      /* Mat2d.data.ptr viewed as [*]f64, 2*ni*nj doubles, host memory  */
} tm_block;

/* boundary.Range (boundary.zig:15-19); start > end means reversed traversal */
typedef struct tm_range {
    uint64_t block;
    uint32_t side; /* tm_side */
    uint32_t _pad;
    uint64_t start, end;
} tm_range;

/* boundary.Connection (boundary.zig:119-123); periodicity maps ranges[0] onto ranges[1]: x0 + p == x1 */
typedef struct tm_connection {
    tm_range ranges[2];
    int32_t has_periodicity;
    int32_t _pad;
    double periodicity[2];
} tm_connection;

/* boundary.Condition (boundary.zig:178-181) */
typedef struct tm_condition {
    tm_range range;
    uint32_t kind; /* tm_condition_kind */
    uint32_t _pad;
} tm_condition;

/* wall_control_function.Algorithm (wall_control_function.zig:10-20, 56-61) */
typedef enum tm_control_function { TM_CF_LAPLACE = 0, TM_CF_WHITE = 1 } tm_control_function;

/* How one outer iteration is advanced.
 *  TM_SOLVER_PICARD_BICGSTAB  reference semantics (smooth.zig:104-154): coefficients lagged at the
 *      previous outer iterate, the linear system solved for x and y by a matrix-free right-preconditioned
 *      BiCGStab (BiCGStab.zig:279-370, preconditioner `diagonal`, solver.zig:22-24).
 *  TM_SOLVER_RELAX  throughput path: `sweeps_per_iteration` damped-Jacobi sweeps of the same 9-point
 *      operator with the coefficients recomputed from the current iterate (32 B/node-update); converges
 *      to the same fixed point as the Picard iteration.
 *  TM_SOLVER_FAS_MULTIGRID  time-to-converged path: `iterations` V(nu,nu) cycles of a geometric full-approximation-scheme
 *      multigrid (nu = sweeps_per_iteration) with the same damped-Jacobi sweep as smoother on every level, stopped early
 *      by stop_max_update (max-norm movement of the mesh over one cycle).  A single block with fixed boundary nodes gets a
 *      non-nested hierarchy (config 3); any multi-block topology -- interfaces, periodic pairs, junctions, sliding inlet /
 *      outlet, one or several GPUs -- gets nested coarse multi-block meshes with Anderson acceleration (config 4, see
 *      tm_mg_plan).  Same fixed point as the two other solvers.  White control function: TM_ERR_UNSUPPORTED.
 */
typedef enum tm_solver { TM_SOLVER_PICARD_BICGSTAB = 0, TM_SOLVER_RELAX = 1, TM_SOLVER_FAS_MULTIGRID = 2 } tm_solver;

typedef struct tm_smooth_options {
    uint32_t struct_size;          /* sizeof(tm_smooth_options), for ABI evolution               */
    uint32_t solver;               /* tm_solver                                                  */
    uint64_t iterations;           /* outer iterations (`iterations` of smooth.mesh)             */
    uint32_t control_function;     /* tm_control_function                                        */
    uint32_t fail_on_no_convergence;
    double white_ds_target;        /* White.ds_target                                            */
    double white_theta_target;     /* White.theta_target (default pi/2)                          */
    /* inner linear solve (Picard modes); defaults of the reference: rtol 1e-6, atol 1e-8, 1000 */
    double rtol, atol;
    uint64_t max_inner_iterations;
    /* relaxation */
    double omega;                  /* Jacobi damping, 0 < omega <= 1                             */
    uint64_t sweeps_per_iteration; /* TM_SOLVER_RELAX: sweeps per outer iteration                */
    double stop_max_update;        /* > 0: stop the outer loop early once max|x_new-x_old| <= this */
    int32_t device;                /* CUDA device ordinal, -1 = current                          */
    int32_t inner_refinement_cycles; /* TM_SOLVER_PICARD_BICGSTAB: after an inner solve has met its tolerance, this many further
                                      cycles each ask for a 10x smaller TRUE residual (iterative refinement).  0 = the reference's
                                      behaviour.  For "exact Picard step" parity runs: the 2-norm of the row-scaled residual bounds the
                                      smooth part of the error only loosely. */
} tm_smooth_options;

typedef struct tm_smooth_stats {
    uint64_t outer_iterations;     /* outer iterations actually run                              */
    uint64_t inner_iterations;     /* total Krylov iterations / Jacobi sweeps                    */
    uint64_t operator_applications;/* applications of the 9-point operator to the whole mesh     */
    uint64_t nodes;                /* total node count                                           */
    double last_sumsq_x, last_sumsq_y; /* sums of smooth.zig:112-134 for the last outer iteration */
    double last_residual;          /* (sumsq_x+sumsq_y)^2, the number smooth.zig:136-137 logs    */
    double last_max_update;        /* max-norm of the last outer update                          */
    double last_inner_residual;    /* max over x,y of the final inner residual norm              */
    double gpu_seconds;            /* device time of the smoothing loop (CUDA events)            */
    int32_t converged;             /* all inner solves reached their tolerance                   */
    int32_t streamed_chunks;       /* tm_smooth_mesh: row chunks the block was streamed in (0 = resident, see tm_smooth_stream_plan) */
} tm_smooth_stats;

/* -------------------------------------------------------------------------------------------------
 * One-shot entry points with HOST buffers (what the Zig call sites bind)
 * ------------------------------------------------------------------------------------------------- */

/* Replaces tfi.linear2dBoundaryBlendedControlFunction (src/core/tfi.zig:112-208) as called from
 * discrete.Block2d.init (src/core/discrete.zig:142-159).  Edges are interleaved x,y host arrays:
 * x_i_min,x_i_max,s1,s2 have ni entries; x_j_min,x_j_max,t1,t2 have nj entries.  out_xy receives
 * 2*ni*nj doubles.  Bit-exact with the reference's operation order (no FMA contraction). */
int tm_tfi_block(uint64_t ni, uint64_t nj,
                 const double *x_i_min, const double *x_i_max,
                 const double *x_j_min, const double *x_j_max,
                 const double *s1, const double *s2, const double *t1, const double *t2,
                 double *out_xy);

/* Replaces smoothing.smooth.mesh (src/core/smoothing/smooth.zig:74-166): smooths all blocks in place.
 * A large single block whose boundary nodes are all fixed, smoothed by a fixed number of TM_SOLVER_RELAX sweeps, is
 * streamed through the device in row chunks so that the host<->device copies overlap the sweeps (bit-identical result,
 * see tm_smooth_stream_plan; stats->streamed_chunks tells); everything else is uploaded, smoothed and downloaded whole.
 * If a streamed call fails half way the block may be partly smoothed. */
int tm_smooth_mesh(tm_block *blocks, size_t n_blocks,
                   const tm_connection *connections, size_t n_connections,
                   const tm_condition *conditions, size_t n_conditions,
                   const tm_smooth_options *opts, tm_smooth_stats *stats);

/* Fills *opts with the defaults (reference tolerances, Laplace control function, Picard/BiCGStab). */
void tm_smooth_options_default(tm_smooth_options *opts);

/* -------------------------------------------------------------------------------------------------
 * Device-resident mesh handle (benchmarks, batches, multi-GPU): H2D/D2H only when asked
 * ------------------------------------------------------------------------------------------------- */
typedef struct tm_mesh tm_mesh; /* opaque */

/* Builds the device mesh: topology analysis (node kinds, junctions, interface tables; smooth.zig:1234-1529),
 * device allocation.  `stream` is a cudaStream_t (NULL = a stream owned by the handle). Block coordinate
 * pointers may be NULL at creation (fill them with tm_mesh_tfi_block or tm_mesh_upload_block). */
int tm_mesh_create(const tm_block *blocks, size_t n_blocks,
                   const tm_connection *connections, size_t n_connections,
                   const tm_condition *conditions, size_t n_conditions,
                   int device, void *stream, tm_mesh **out);
void tm_mesh_destroy(tm_mesh *mesh);

int tm_mesh_upload_block(tm_mesh *mesh, size_t block, const double *xy);     /* host -> device  */
int tm_mesh_download_block(tm_mesh *mesh, size_t block, double *xy);         /* device -> host  */
/* The same copy-back (smooth.zig:139-153) without holding the mesh: the block is snapshot on the device and the snapshot goes
 * to `xy` (pinned host memory, for the copy to overlap) on a stream of its own, so the next tm_mesh_tfi_block /
 * tm_mesh_smooth may start while the result of this one is still on its way.  `xy` is valid after tm_mesh_download_wait. */
int tm_mesh_download_block_async(tm_mesh *mesh, size_t block, double *xy);
int tm_mesh_download_wait(tm_mesh *mesh);
/* TFI of one block directly into the device mesh; edge arrays are HOST pointers (O(ni+nj) data) and stay
 * cached on the device, so tm_mesh_tfi_block_resident can re-run the TFI without any host traffic. */
int tm_mesh_tfi_block(tm_mesh *mesh, size_t block,
                      const double *x_i_min, const double *x_i_max,
                      const double *x_j_min, const double *x_j_max,
                      const double *s1, const double *s2, const double *t1, const double *t2);
int tm_mesh_tfi_block_resident(tm_mesh *mesh, size_t block);
/* White control function groups (extension for batches of independent cuts in one mesh).  The reference applies White
 * to blocks 0 and 1 and connection 0 only (wall_control_function.zig:72, 204-217) -- that is the default.  A batch names
 * one pair (A, B) of O-grid half blocks per cut; each pair needs a connection A:j_min[0..] <-> B:j_min[0..].
 * block_pairs holds 2*n_groups block indices.  Call before tm_mesh_begin_smoothing. */
int tm_mesh_set_white_groups(tm_mesh *mesh, const uint64_t *block_pairs, size_t n_groups);
/* Freezes the current coordinates as the initial mesh: checks interface coincidence
 * (connectionDataCheck, smooth.zig:220-275), captures fixed/sliding boundary values
 * (smooth.zig:790-796, 853-858) and initialises the control function (wall_control_function.zig:27-42). */
int tm_mesh_begin_smoothing(tm_mesh *mesh, const tm_smooth_options *opts);
/* Runs opts->iterations outer iterations on the device-resident mesh (may be called repeatedly). */
int tm_mesh_smooth(tm_mesh *mesh, const tm_smooth_options *opts, tm_smooth_stats *stats);
/* Blocks until all work queued on the mesh's stream has finished. */
int tm_mesh_synchronize(tm_mesh *mesh);

/* Independent systems.  Blocks that no connection joins never exchange a value: every connected component of the block
 * graph -- every 2D cut of a batch -- is a linear system of its own.  TM_SOLVER_PICARD_BICGSTAB solves each component for
 * x and y with its own Krylov scalars, its own ||b||, tolerance max(atol, rtol ||b||) (GMRES.zig:305-306 /
 * BiCGStab.zig:291), iteration count and stopping test, exactly as the reference does when it meshes the cuts one after the
 * other (smooth.zig:104-154 per mesh).  Components are numbered in the order of their lowest block.  The per-component
 * record describes the LAST outer iteration of the last tm_mesh_smooth call (single-process meshes; index 0 = x solve,
 * 1 = y solve; status 1 converged, 2 breakdown, 3 iteration cap). */
typedef struct tm_component_stats {
    uint64_t nodes;
    uint64_t iterations[2];
    double tolerance[2], norm_b[2], norm_r[2];
    int32_t status[2];
    uint64_t operator_applications;   /* applications of the 9-point operator to this component (both solves advance together) */
    uint64_t restarts;                /* true-residual restarts (1 = none) */
} tm_component_stats;
uint64_t tm_mesh_component_count(const tm_mesh *mesh);
int tm_mesh_component_of_block(const tm_mesh *mesh, size_t block, uint64_t *component);
int tm_mesh_component_stats(const tm_mesh *mesh, size_t component, tm_component_stats *out);

/* Accessors (mirror the WASM surface, src/wasm/lib.zig:97-124) */
uint64_t tm_mesh_block_count(const tm_mesh *mesh);
uint64_t tm_mesh_node_count(const tm_mesh *mesh);
int tm_mesh_block_size(const tm_mesh *mesh, size_t block, uint64_t *ni, uint64_t *nj);
/* device pointer to the current coordinates of a block (interleaved x,y) */
double *tm_mesh_block_device_ptr(tm_mesh *mesh, size_t block);
/* device pointer / host copy of the control function (P,Q per node, wall_control_function.zig:24) */
int tm_mesh_download_control_function(tm_mesh *mesh, size_t block, double *pq);
/* Structured output, the step right after the path: the block as two struct-of-arrays fields with i fastest
 * (x[j*ni + i], y[j*ni + i]) -- exactly the buffers the reference's CGNS writer passes to cg_coord_write for
 * CoordinateX / CoordinateY (src/core/cgns.zig:69-101) and, with TM_FIELD_CONTROL_FUNCTION, to cg_field_write for the
 * P / Q solution fields (cgns.zig:110-161).  The transposition runs on the device; x and y receive ni*nj doubles each
 * (host memory, ideally pinned). */
typedef enum tm_field { TM_FIELD_COORDINATES = 0, TM_FIELD_CONTROL_FUNCTION = 1 } tm_field;
int tm_mesh_download_block_soa(tm_mesh *mesh, size_t block, int field /* tm_field */, double *x, double *y);
/* Structured writer fed by the same device-side transposition: all blocks held by this process as a multi-block 2D PLOT3D
 * grid file (binary, C stream layout, fp64, no IBLANK: int32 nblocks; nblocks x (int32 ni, int32 nj); per block all x with i
 * fastest, then all y) -- the wire format of structured multi-block grids, and block for block the arrays the reference hands
 * to cg_coord_write (src/core/cgns.zig:26-168; CGNS itself needs libcgns / HDF5, which the reference links as system
 * libraries).  function_path (may be NULL) receives the control function as a PLOT3D function file with two variables
 * (P, Q), the fields cgns.zig:110-161 writes as a FlowSolution.  Blocks of other ranks are not written. */
int tm_mesh_write_plot3d(tm_mesh *mesh, const char *grid_path, const char *function_path);

/* Viewer buffers of a (single-GPU) device mesh, built on the device: what createPointBuffer and
 * createWireframeElementBuffer build on the host (src/gui/lib.zig:227-318).  points receives 2*n_points floats (x,y of
 * all blocks in block order, f64 -> f32 round to nearest), ranges receives x_min, x_max, y_min, y_max (the maxima start
 * at the smallest positive normal float, as in the reference), indices receives n_indices 32-bit point indices: per
 * block first the line segments along j, then those along i.  Any of the three may be NULL; each may be host memory or
 * device memory (e.g. the mapped pointer of a CUDA-registered OpenGL buffer). */
int tm_mesh_viewer_sizes(const tm_mesh *mesh, uint64_t *n_points, uint64_t *n_indices);
int tm_mesh_viewer_buffers(tm_mesh *mesh, float *points, float *ranges /* 4 */, uint32_t *indices);
/* node kind per block-boundary node in the reference's flat boundary numbering (boundary.zig:248-285);
 * values: 0 fixed, 1 smoothed, 2 connected, 3 laplacian_smoothed, 4 sliding_circ (smooth.zig:1168-1174).
 * `kinds` receives 2*(ni+nj-2) bytes. */
int tm_mesh_download_boundary_kinds(tm_mesh *mesh, size_t block, uint8_t *kinds);

/* -------------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, whole blocks assigned to ranks (DESIGN.md "Multi-GPU")
 * ------------------------------------------------------------------------------------------------- */
#define TM_UNIQUE_ID_BYTES 128

/* Rank 0 obtains an id (ncclGetUniqueId) and hands the bytes to the other ranks with its own plumbing
 * (MPI, torch.distributed, a file ...) before every rank calls tm_mesh_create_distributed. */
int tm_dist_get_unique_id(uint8_t *id /* TM_UNIQUE_ID_BYTES */);

/* Like tm_mesh_create, for rank `rank` of `n_ranks`: every rank passes the SAME global block / connection / condition
 * arrays (coordinates are only read for the blocks it owns and may be NULL elsewhere) and block_owner[n_blocks].
 * Rows are computed by the owner of their node; per sweep / operator application the ranks exchange the few nodes
 * their neighbours read (NCCL send/recv) and all-reduce the residual / Krylov scalars.  All other tm_mesh_* calls
 * keep global block indices and are valid for owned blocks; tm_mesh_begin_smoothing and tm_mesh_smooth are collective.
 * rank = -1 builds ALL ranks inside this process on one GPU (no NCCL): an emulation for testing the multi-rank
 * logic, every block is then addressable through the one handle. */
int tm_mesh_create_distributed(const tm_block *blocks, size_t n_blocks,
                               const tm_connection *connections, size_t n_connections,
                               const tm_condition *conditions, size_t n_conditions,
                               const int32_t *block_owner, int rank, int n_ranks,
                               const uint8_t *unique_id /* TM_UNIQUE_ID_BYTES; may be NULL when n_ranks == 1 or rank == -1 */,
                               int device, void *stream, tm_mesh **out);
uint64_t tm_mesh_local_node_count(const tm_mesh *mesh); /* nodes of the blocks held by this process */
/* How the per-sweep halo exchange of this mesh travels.  TM_HALO_PEER_MEMORY: the owner pushes the nodes its neighbours
 * ghost straight into their fields over NVLink (buffers mapped with CUDA IPC, one gather+store+signal kernel per
 * exchange); chosen when every rank could map its neighbours' buffers (TM_P2P=0 in the environment forces NCCL).
 * TM_HALO_NCCL: pack kernel + ncclSend/ncclRecv group. */
typedef enum tm_halo_path { TM_HALO_NONE = 0, TM_HALO_EMULATED = 1, TM_HALO_NCCL = 2, TM_HALO_PEER_MEMORY = 3 } tm_halo_path;
int tm_mesh_halo_path(const tm_mesh *mesh);

/* Host-only view of the partition (needs no GPU): sizes of rank `rank`'s local field and its exchange lists.
 * ghost_ids / send_ids (may be NULL) receive the global node ids grouped by peer rank in ascending rank order;
 * counts (may be NULL) receives 2*n_ranks entries: [ghosts from peer p, sends to peer p]. */
typedef struct tm_dist_plan_info {
    uint64_t n_own, n_ghost, n_synth, n_send;
    uint64_t n_smoothed, n_junction, n_sliding, n_slaves;
} tm_dist_plan_info;
int tm_dist_plan(const tm_block *blocks, size_t n_blocks,
                 const tm_connection *connections, size_t n_connections,
                 const tm_condition *conditions, size_t n_conditions,
                 const int32_t *block_owner, int rank, int n_ranks,
                 tm_dist_plan_info *info, int64_t *ghost_ids, int64_t *send_ids, int64_t *counts);

/* -------------------------------------------------------------------------------------------------
 * Edge discretisation, the step right before the path (batched): discrete.Edge.init = clustering.create +
 * Curve.interpolate (src/core/discrete.zig:17-31; clustering.zig:9-116, geometry.zig:26-40, spline.zig:74-139, 202-222)
 * ------------------------------------------------------------------------------------------------- */
typedef enum tm_clustering_kind { TM_CLUSTERING_UNIFORM = 0, TM_CLUSTERING_ROBERTS = 1, TM_CLUSTERING_SINGLE_HYPERBOLIC = 2 } tm_clustering_kind;
typedef enum tm_curve_kind { TM_CURVE_LINE = 0, TM_CURVE_SPLINE = 1 } tm_curve_kind;
/* views of the fields of an already fitted spline.FittingSpline(2) (spline.zig:24-40): the library borrows them */
typedef struct tm_spline {
    uint64_t n_points;            /* knots                                                            */
    const double *params;         /* chord-length parameters, n_points                                */
    const double *points;         /* interleaved x,y, 2*n_points                                      */
    const double *second_derivs_x, *second_derivs_y; /* n_points each                                 */
    uint64_t n_samples;           /* arc-length table entries (201 in the reference, spline.zig:22)   */
    const double *sample_arc;     /* normalised arc length at the uniform parameters i/(n_samples-1)  */
    double total_length;
} tm_spline;
typedef struct tm_edge_job {
    uint64_t n;                   /* points of the edge                                               */
    uint32_t curve_kind;          /* tm_curve_kind                                                    */
    uint32_t clustering_kind;     /* tm_clustering_kind                                               */
    double line_start[2], line_end[2]; /* TM_CURVE_LINE                                               */
    const tm_spline *spline;      /* TM_CURVE_SPLINE                                                  */
    double alpha, beta;           /* Roberts (clustering.zig:24-27)                                   */
    double delta_s;               /* single hyperbolic (clustering.zig:56-59)                         */
    double *points;               /* out: 2*n doubles, interleaved x,y (Edge.points), host memory     */
    double *clustering;           /* out: n doubles (Edge.clustering), host memory                    */
} tm_edge_job;
/* All edges in one launch (one CTA per edge).  Curve arithmetic is bit-exact with the reference's operation order; the
 * pow / tanh inside the Roberts and hyperbolic clusterings are CUDA's (a few ulp from any host libm). */
int tm_edges_discretize(const tm_edge_job *jobs, size_t n_jobs, int device);

/* Spline fit, batched on the device (one CTA per spline): spline.FittingSpline.init (src/core/spline.zig:24-110, 141-200) --
 * chord-length parameters, natural-cubic second derivatives, the arc-length table of n_samples entries (201 in the
 * reference) and the total length.  The outputs are exactly the slices a tm_spline views; bit-exact with the reference's
 * operation order.  Coincident consecutive points (CoincidentParameters, spline.zig:176-178): TM_ERR_INVALID_ARGUMENT. */
typedef struct tm_spline_fit_job {
    uint64_t n_points;            /* >= 2                                                              */
    const double *points;         /* interleaved x,y, 2*n_points                                       */
    uint64_t n_samples;           /* >= 2                                                              */
    double *params, *second_derivs_x, *second_derivs_y;   /* out: n_points each, host memory           */
    double *sample_arc;           /* out: n_samples                                                    */
    double *total_length;         /* out: 1                                                            */
} tm_spline_fit_job;
int tm_splines_fit(const tm_spline_fit_job *jobs, size_t n_jobs, int device);

/* The two other edge operations of the automated blocking, batched on the device (one CTA per job, bit-exact with the
 * reference's operation order):
 *   Edge.combine (src/core/discrete.zig:38-91 over the views of :94-136): the views' points back to back without the
 *   duplicated joints (which must agree within 1e-10, else TM_ERR_INVALID_ARGUMENT); clustering re-accumulated and normalised.
 *   projectNormal (src/core/templates/O4H.zig:531-574): every point moved by `distance` along the normal of the edge. */
typedef struct tm_edge_view {
    const double *points;         /* source Edge.points, interleaved x,y                               */
    const double *clustering;     /* source Edge.clustering                                            */
    uint64_t n;                   /* points of the source edge                                         */
    uint64_t start, end;          /* EdgeView.start / .end (start > end: traversed backwards)          */
} tm_edge_view;
typedef struct tm_combine_job {
    const tm_edge_view *views;    /* at least 2, at most 16                                            */
    uint64_t n_views;
    double *points;               /* out: 2 * n doubles with n = sum(len) - (n_views - 1), host memory */
    double *clustering;           /* out: n doubles                                                    */
} tm_combine_job;
int tm_edges_combine(const tm_combine_job *jobs, size_t n_jobs, int device);
typedef struct tm_project_job {
    const double *points;         /* Edge.points, 2 * n doubles                                        */
    uint64_t n;                   /* >= 2                                                              */
    double distance;
    double *out;                  /* 2 * n doubles, host memory                                        */
} tm_project_job;
int tm_edges_project_normal(const tm_project_job *jobs, size_t n_jobs, int device);

/* Host-only view of the multigrid hierarchy TM_SOLVER_FAS_MULTIGRID builds for a multi-block mesh (needs no GPU): nested
 * coarsening of block sizes and connection / condition ranges, directions tied into classes by the connections.
 * cell_size (may be NULL = all equal) holds the mean cell size per (block, direction), 2*n_blocks entries -- the solver
 * measures it on the device.  *n_levels receives the number of levels (>= 1); sizes (may be NULL) receives (ni, nj) of
 * every block on every level: sizes[(level*n_blocks + block)*2 + {0,1}] for level < max_levels. */
int tm_mg_plan(const tm_block *blocks, size_t n_blocks,
               const tm_connection *connections, size_t n_connections,
               const tm_condition *conditions, size_t n_conditions,
               const double *cell_size, size_t max_levels, uint64_t *n_levels, uint64_t *sizes);

/* Host-only view of how tm_smooth_mesh streams a large single block through the device (needs no GPU).  A block of
 * ni x nj nodes whose boundary nodes are all fixed (no connections, no inlet / outlet), smoothed by TM_SOLVER_RELAX with
 * the Laplace control function for a fixed number of `sweeps`, is cut into row chunks: chunk k owns the rows
 * [owned_first[k], owned_first[k+1]) and travels as the window of `*window_rows` rows starting at window_first[k], which
 * holds at least `sweeps` extra rows on every side that is not a block boundary -- enough for `sweeps` Jacobi sweeps of
 * the window to leave the owned rows exactly as sweeps of the whole block would; upload, sweeps and download of
 * successive chunks overlap.  *n_chunks = 0: the block is smoothed resident (too small, too many sweeps for its
 * extent, or TM_STREAM=0).  window_first needs 8 entries, owned_first 9. */
int tm_smooth_stream_plan(uint64_t ni, uint64_t nj, uint64_t sweeps, uint64_t *n_chunks, uint64_t *window_rows,
                          uint64_t *window_first, uint64_t *owned_first);

/* -------------------------------------------------------------------------------------------------
 * Misc
 * ------------------------------------------------------------------------------------------------- */
const char *tm_last_error(void);
/* Device buffers (and the pinned scalar block of a mesh) are recycled through a process-wide cache when a mesh is
 * destroyed, so that the one-shot entry points do not pay cudaMalloc / cudaFree per call; at most TM_CACHE_GB GiB
 * (environment, default 8) are kept.  The three window meshes of the streamed tm_smooth_mesh are parked between calls
 * likewise.  This returns everything to the driver (call it before unloading the library). */
void tm_release_cached_memory(void);
int tm_abi_version(void);
/* number of CUDA kernel launches issued by this library since load (bench.py's gpu_launches) */
uint64_t tm_kernel_launch_count(void);
/* name, SM count, memory of the device in use; returns TM_ERR_NO_DEVICE when there is none */
int tm_device_info(int device, char *name, size_t name_len, int *sm_count, uint64_t *global_mem_bytes);

#ifdef __cplusplus
}
#endif
#endif /* TURBOMESH_GPU_H */
