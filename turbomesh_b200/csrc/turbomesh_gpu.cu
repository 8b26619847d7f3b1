// turbomesh_gpu.cu -- implementation of the C ABI declared in include/turbomesh_gpu.h.
//
// Host orchestration of the sm_100a kernels in kernels.cuh: device mesh handle, topology upload, the outer
// (Picard) loop of smoothing.smooth.mesh (src/core/smoothing/smooth.zig:74-166), the matrix-free BiCGStab
// (src/core/smoothing/BiCGStab.zig:279-370) and the relaxation sweeps.  No CPU compute path exists here:
// without a CUDA device every entry point fails with TM_ERR_NO_DEVICE.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/turbomesh_gpu.h"
#include "kernels.cuh"

using namespace tmesh;

namespace {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int set_error(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                           \
    do {                                                                                                         \
        cudaError_t _e = (expr);                                                                                 \
        if (_e != cudaSuccess) {                                                                                 \
            const int _code = (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? TM_ERR_NO_DEVICE  \
                              : (_e == cudaErrorMemoryAllocation ? TM_ERR_OUT_OF_MEMORY : TM_ERR_CUDA);          \
            TM_THROW(_code, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);         \
        }                                                                                                        \
    } while (0)

#define LAUNCH(kernel, grid, block, stream, ...)                     \
    do {                                                             \
        auto _kfn = kernel;                                          \
        _kfn<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);         \
        g_launches.fetch_add(1, std::memory_order_relaxed);          \
        CUDA_TRY(cudaGetLastError());                                \
    } while (0)

void require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        TM_THROW(TM_ERR_NO_DEVICE, "no usable CUDA device (%s); turbomesh_gpu has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device >= n) TM_THROW(TM_ERR_INVALID_ARGUMENT, "device ordinal %d out of range (%d devices)", device, n);
    if (device >= 0) CUDA_TRY(cudaSetDevice(device));
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        release();
        if (count == 0) return;
        CUDA_TRY(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void upload(const std::vector<T>& h, cudaStream_t s) {
        alloc(h.size());
        if (!h.empty()) {
            CUDA_TRY(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaStreamSynchronize(s));  // h may be a temporary
        }
    }
    void zero(cudaStream_t s) {
        if (p) CUDA_TRY(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

struct EdgeCache {  // device copies of the four edges + clusterings of one block (TFI inputs)
    DevBuf<double> buf;
    bool valid = false;
};

}  // namespace

struct tm_mesh {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    Topology topo;
    int64_t N = 0;

    DevBuf<double2> X[2];
    int cur = 0;
    DevBuf<double2> pq, wall_pq;
    DevBuf<DevBlock> d_blocks;
    DevBuf<Tile> d_tiles;
    DevBuf<SmoothedRow> d_srows;
    DevBuf<JunctionRow> d_jrows;
    DevBuf<SlidingRow> d_lrows;
    DevBuf<SlaveRow> d_slaves, d_cslaves;
    DevBuf<FixedOverride> d_fo;
    DevBuf<PairCheck> d_pairs;
    DevBuf<RhsTerm> d_rhs_terms;
    DevBuf<double> part_int, part_bnd, part_vec, bconst;
    DevBuf<unsigned long long> d_worst;
    DevBuf<SolveCtl> d_ctl;
    SolveCtl* h_ctl = nullptr;  // pinned
    DevBuf<double2> kr, krhat, kp, kv, ks, kt;
    bool krylov_ready = false;
    std::vector<EdgeCache> edges;
    std::vector<uint8_t> have_coords;

    int n_tiles = 0, n_bnd_rows = 0, n_bnd_ctas = 0, vec_grid = 1;
    int tile_rows = TILE_I;   // TM_TILE_ROWS overrides (tuning aid)
    bool use_bulk = true;     // TM_INTERIOR=regs selects the register-only interior kernel (tuning aid)
    bool begun = false;
    int cf = TM_CF_LAPLACE;
    WhiteParams wp{};
    uint64_t outer_done = 0;  // outer iterations since begin_smoothing (the `n` of system.fill(n), smooth.zig:1107-1110)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    ~tm_mesh() {
        if (h_ctl) cudaFreeHost(h_ctl);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

namespace {

void build_tiles(tm_mesh* m) {
    std::vector<Tile> tiles;
    std::vector<DevBlock> blocks;
    for (size_t b = 0; b < m->topo.blocks.size(); ++b) {
        const auto& B = m->topo.blocks[b];
        blocks.push_back(DevBlock{B.off, int32_t(B.ni), int32_t(B.nj)});
        // rows per CTA: about TILE_I, evened out over the block so no CTA gets a short remainder
        const int64_t interior_i = B.ni - 2;
        const int64_t n_i = std::max<int64_t>(1, (interior_i + m->tile_rows - 1) / m->tile_rows);
        const int64_t rows = (interior_i + n_i - 1) / n_i;
        for (int64_t i0 = 1; i0 <= B.ni - 2; i0 += rows)
            for (int64_t j0 = 1; j0 <= B.nj - 2; j0 += TILE_J) tiles.push_back(Tile{int32_t(b), int32_t(i0), int32_t(j0), int32_t(rows)});
    }
    m->n_tiles = int(tiles.size());
    m->d_tiles.upload(tiles, m->stream);
    m->d_blocks.upload(blocks, m->stream);
}

void build_rhs_terms(tm_mesh* m, std::vector<RhsTerm>& terms) {
    // the rows of the reference system whose rhs is not zero by construction (smooth.zig:780-921)
    const Topology& T = m->topo;
    std::vector<uint8_t> over(size_t(T.n_boundary), 0);
    for (const auto& f : T.fixed_overrides) {
        terms.push_back(RhsTerm{f.self, f.x, f.y, 0, 0});
        over[size_t(T.bid_of_global(f.self))] = 1;
    }
    for (size_t b = 0; b < T.blocks.size(); ++b) {
        const auto& B = T.blocks[b];
        auto visit = [&](int64_t i, int64_t j) {
            const int64_t local = i * B.nj + j;
            const size_t id = size_t(T.bid(b, local));
            if (T.kind[id] == K_FIXED && !over[id]) terms.push_back(RhsTerm{B.off + local, 0.0, 0.0, 1, 1});
        };
        for (int64_t j = 0; j < B.nj; ++j) { visit(0, j); visit(B.ni - 1, j); }
        for (int64_t i = 1; i + 1 < B.ni; ++i) { visit(i, 0); visit(i, B.nj - 1); }
    }
    for (const auto& s : T.sliding) terms.push_back(RhsTerm{s.self, s.rhs_x, s.rhs_y, s.rhs_x_from_initial, 0});
    for (const auto& j : T.junction_rows) terms.push_back(RhsTerm{j.self, j.rhs_x, j.rhs_y, 0, 0});
}

void ensure_krylov(tm_mesh* m) {
    if (m->krylov_ready) return;
    for (DevBuf<double2>* v : {&m->kr, &m->krhat, &m->kp, &m->kv, &m->ks, &m->kt}) {
        v->alloc(size_t(m->N));
        v->zero(m->stream);
    }
    m->krylov_ready = true;
}

int bnd_ctas(int rows) { return (rows + BND_THREADS - 1) / BND_THREADS; }

// ---- kernel dispatch over the (LAGGED, HAS_PQ) template space -------------------------------------
template <int MODE, int STATS>
void launch_rows(tm_mesh* m, bool lagged, const double2* u, const double2* xc, double2* out, double omega, const double2* dot_a) {
    const bool has_pq = m->cf == TM_CF_WHITE;
    const double2* pq = m->pq.p;
    cudaStream_t s = m->stream;
#define TM_ROWS(LAG, PQ)                                                                                                                              \
    do {                                                                                                                                              \
        if (m->n_tiles > 0)                                                                                                                           \
            LAUNCH((winslow_interior_kernel<MODE, LAG, PQ, STATS>), m->n_tiles, TILE_J, s, m->d_tiles.p, m->d_blocks.p, u, xc, pq, out, omega, dot_a, \
                   m->part_int.p);                                                                                                                    \
        if (m->n_bnd_rows > 0)                                                                                                                        \
            LAUNCH((winslow_boundary_kernel<MODE, LAG, PQ, STATS>), m->n_bnd_ctas, BND_THREADS, s, m->d_srows.p, int(m->topo.smoothed.size()),        \
                   m->d_jrows.p, int(m->topo.junction_rows.size()), m->d_lrows.p, int(m->topo.sliding.size()), m->d_slaves.p, u, xc, pq, out, omega,  \
                   dot_a, m->part_bnd.p);                                                                                                             \
    } while (0)
#define TM_ROWS_BULK(PQ)                                                                                                                          \
    do {                                                                                                                                          \
        if (m->n_tiles > 0)                                                                                                                       \
            LAUNCH((winslow_interior_bulk_kernel<MODE, PQ, STATS>), m->n_tiles, TILE_J, s, m->d_tiles.p, m->d_blocks.p, u, pq, out, omega, dot_a, \
                   m->part_int.p);                                                                                                                \
        if (m->n_bnd_rows > 0)                                                                                                                    \
            LAUNCH((winslow_boundary_kernel<MODE, false, PQ, STATS>), m->n_bnd_ctas, BND_THREADS, s, m->d_srows.p, int(m->topo.smoothed.size()),  \
                   m->d_jrows.p, int(m->topo.junction_rows.size()), m->d_lrows.p, int(m->topo.sliding.size()), m->d_slaves.p, u, xc, pq, out,     \
                   omega, dot_a, m->part_bnd.p);                                                                                                  \
    } while (0)
    if (lagged) { if (has_pq) TM_ROWS(true, true); else TM_ROWS(true, false); }
    else if (m->use_bulk) { if (has_pq) TM_ROWS_BULK(true); else TM_ROWS_BULK(false); }
    else        { if (has_pq) TM_ROWS(false, true); else TM_ROWS(false, false); }
#undef TM_ROWS
#undef TM_ROWS_BULK
}

void launch_reduce(tm_mesh* m, int op, const tm_smooth_options* o, bool from_rows) {
    const int max_it = o->max_inner_iterations > 0x7fffffffull ? 0x7fffffff : int(o->max_inner_iterations);
    if (from_rows)
        LAUNCH((reduce_kernel<256>), 1, 256, m->stream, m->part_int.p, m->n_tiles, op, m->d_ctl.p, o->rtol, o->atol, max_it, m->part_bnd.p, m->n_bnd_ctas, m->bconst.p);
    else
        LAUNCH((reduce_kernel<256>), 1, 256, m->stream, m->part_vec.p, m->vec_grid, op, m->d_ctl.p, o->rtol, o->atol, max_it, (const double*)nullptr, 0, m->bconst.p);
}

void sync_slaves(tm_mesh* m, double2* v, int mode) {
    const int n = int(m->topo.slaves.size());
    if (n > 0) LAUNCH(sync_slaves_kernel, (n + 127) / 128, 128, m->stream, m->d_slaves.p, n, v, mode);
}

void fetch_ctl(tm_mesh* m) {
    CUDA_TRY(cudaMemcpyAsync(m->h_ctl, m->d_ctl.p, sizeof(SolveCtl), cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
}

void white_step(tm_mesh* m, bool update) {
    const int nw = m->wp.ni0 + m->wp.ni1;
    LAUNCH(white_wall_kernel, (nw + 127) / 128, 128, m->stream, m->wp, m->X[m->cur].p, m->wall_pq.p, update ? 1 : 0);
    const int64_t nn = int64_t(m->wp.ni0) * m->wp.nj0 + int64_t(m->wp.ni1) * m->wp.nj1;
    LAUNCH(white_blend_kernel, unsigned((nn + 255) / 256), 256, m->stream, m->wp, m->wall_pq.p, m->pq.p);
}

void validate_options(const tm_smooth_options* o) {
    if (!o) TM_THROW(TM_ERR_INVALID_ARGUMENT, "options are NULL");
    if (o->struct_size != sizeof(tm_smooth_options)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "tm_smooth_options.struct_size mismatch (ABI version?)");
    if (o->solver > TM_SOLVER_RELAX) TM_THROW(TM_ERR_INVALID_ARGUMENT, "unknown solver %u", o->solver);
    if (o->control_function > TM_CF_WHITE) TM_THROW(TM_ERR_INVALID_ARGUMENT, "unknown control function %u", o->control_function);
    if (!(o->omega > 0.0 && o->omega <= 1.0)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "omega must be in (0, 1]");
    if (o->solver == TM_SOLVER_RELAX && o->sweeps_per_iteration == 0) TM_THROW(TM_ERR_INVALID_ARGUMENT, "sweeps_per_iteration must be > 0");
    if (!(o->rtol >= 0.0) || !(o->atol >= 0.0)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "tolerances must be non-negative");
}

// ---- the two ways of advancing one outer iteration -------------------------------------------------
void run_relax(tm_mesh* m, const tm_smooth_options* o, tm_smooth_stats* st) {
    for (uint64_t it = 0; it < o->iterations; ++it) {
        if (m->cf == TM_CF_WHITE && m->outer_done > 0) white_step(m, true);
        for (uint64_t sw = 0; sw < o->sweeps_per_iteration; ++sw) {
            const bool last = sw + 1 == o->sweeps_per_iteration;
            const double2* u = m->X[m->cur].p;
            double2* out = m->X[1 - m->cur].p;
            if (last) launch_rows<MODE_RELAX, 1>(m, false, u, u, out, o->omega, nullptr);
            else launch_rows<MODE_RELAX, 0>(m, false, u, u, out, o->omega, nullptr);
            m->cur = 1 - m->cur;
            st->inner_iterations += 1;
            st->operator_applications += 1;
        }
        launch_reduce(m, RED_UPDATE_STATS, o, true);
        m->outer_done += 1;
        st->outer_iterations += 1;
        if (o->stop_max_update > 0.0) {
            fetch_ctl(m);
            if (m->h_ctl->max_update <= o->stop_max_update) break;
        }
    }
}

// BiCGStab iterations (BiCGStab.zig:303-366) on the row-scaled system until both components report done.
void bicgstab_cycle(tm_mesh* m, const tm_smooth_options* o, tm_smooth_stats* st, double2* x, const double2* xc) {
    cudaStream_t s = m->stream;
    const int64_t N = m->N;
    const int check_every = 8;
    for (uint64_t k = 0;; ++k) {
        if (k % check_every == 0) {
            fetch_ctl(m);
            if (m->h_ctl->done[0] && m->h_ctl->done[1]) break;
        }
        LAUNCH(bicg_p_kernel, m->vec_grid, VEC_THREADS, s, N, m->d_ctl.p, m->kr.p, m->kp.p, m->kv.p);
        sync_slaves(m, m->kp.p, 0);
        launch_rows<MODE_APPLY, 2>(m, true, m->kp.p, xc, m->kv.p, 1.0, m->krhat.p);   // v = A p, partial rhat.v
        launch_reduce(m, RED_ALPHA, o, true);
        LAUNCH(bicg_s_kernel, m->vec_grid, VEC_THREADS, s, N, m->d_ctl.p, m->kr.p, m->kv.p, m->ks.p, x, m->kp.p, m->part_vec.p);
        launch_reduce(m, RED_NORM_S, o, false);
        sync_slaves(m, m->ks.p, 0);
        launch_rows<MODE_APPLY, 3>(m, true, m->ks.p, xc, m->kt.p, 1.0, m->ks.p);      // t = A s, partials t.s and t.t
        sync_slaves(m, m->ks.p, 2);
        launch_reduce(m, RED_OMEGA, o, true);
        LAUNCH(bicg_r_kernel, m->vec_grid, VEC_THREADS, s, N, m->d_ctl.p, m->ks.p, m->kt.p, m->kr.p, x, m->krhat.p, m->part_vec.p);
        launch_reduce(m, RED_NORM_R, o, false);
        st->operator_applications += 2;
    }
}

// One outer (Picard) iteration = the reference's fill + solve(x) + solve(y) (smooth.zig:104-154), with the two
// solves advanced in lock-step by a matrix-free BiCGStab on the row-scaled system.  One extension over BiCGStab.zig
// that only matters when the tolerance is tighter than the reference's: whenever the solver reports convergence or
// breaks down, the recursive residual is replaced by the true one and, if that is still above the tolerance, the
// iteration restarts from there (at most `max_restarts` times).
void run_picard_bicgstab(tm_mesh* m, const tm_smooth_options* o, tm_smooth_stats* st) {
    ensure_krylov(m);
    cudaStream_t s = m->stream;
    const int64_t N = m->N;
    const size_t bytes = size_t(N) * sizeof(double2);
    const int max_restarts = 40;
    st->converged = 1;
    for (uint64_t it = 0; it < o->iterations; ++it) {
        if (m->cf == TM_CF_WHITE && m->outer_done > 0) white_step(m, true);
        const double2* xc = m->X[m->cur].p;      // lagged coordinates: the mesh before this iteration
        double2* x = m->X[1 - m->cur].p;         // x_new / y_new, warm-started from the mesh (GMRES.zig:157-174)
        CUDA_TRY(cudaMemcpyAsync(x, xc, bytes, cudaMemcpyDeviceToDevice, s));
        for (int cycle = 0;; ++cycle) {
            if (cycle > 0) sync_slaves(m, x, 1);
            launch_rows<MODE_RESID, 4>(m, true, x, xc, m->kr.p, 1.0, nullptr);        // r = D^-1 (b - A x)
            st->operator_applications += 1;
            launch_reduce(m, cycle == 0 ? RED_INIT : RED_RESTART, o, true);
            fetch_ctl(m);
            const int d0 = m->h_ctl->done[0], d1 = m->h_ctl->done[1];
            if ((d0 == 1 && d1 == 1) || d0 == 3 || d1 == 3 || cycle > max_restarts) break;
            CUDA_TRY(cudaMemcpyAsync(m->krhat.p, m->kr.p, bytes, cudaMemcpyDeviceToDevice, s));
            m->kp.zero(s);
            m->kv.zero(s);
            bicgstab_cycle(m, o, st, x, xc);
        }
        sync_slaves(m, x, 1);
        st->inner_iterations += uint64_t(m->h_ctl->iters[0]) + uint64_t(m->h_ctl->iters[1]);
        st->last_inner_residual = std::fmax(m->h_ctl->norm_r[0], m->h_ctl->norm_r[1]);
        if (m->h_ctl->done[0] != 1 || m->h_ctl->done[1] != 1) st->converged = 0;  // log.warn "did not converge", BiCGStab.zig:368-369
        LAUNCH(diff_stats_kernel, m->vec_grid, VEC_THREADS, s, N, xc, (const double2*)x, m->part_vec.p);
        launch_reduce(m, RED_UPDATE_STATS, o, false);
        m->cur = 1 - m->cur;  // copy-back (smooth.zig:139-153) is a buffer swap
        m->outer_done += 1;
        st->outer_iterations += 1;
        if (o->stop_max_update > 0.0) {
            fetch_ctl(m);
            if (m->h_ctl->max_update <= o->stop_max_update) break;
        }
    }
}

template <class F>
int guarded(F&& f) {
    try {
        f();
        return TM_OK;
    } catch (const Error& e) {
        return set_error(e.code, e.msg);
    } catch (const std::bad_alloc&) {
        return set_error(TM_ERR_OUT_OF_MEMORY, "host allocation failed");
    } catch (const std::exception& e) {
        return set_error(TM_ERR_CUDA, e.what());
    }
}

void check_mesh(const tm_mesh* m) {
    if (!m) TM_THROW(TM_ERR_INVALID_ARGUMENT, "mesh handle is NULL");
}
void check_block(const tm_mesh* m, size_t block) {
    check_mesh(m);
    if (block >= m->topo.blocks.size()) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block index %zu out of range", block);
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
#pragma GCC visibility push(default)
extern "C" {

const char* tm_last_error(void) { return g_last_error.c_str(); }
int tm_abi_version(void) { return TM_ABI_VERSION; }
uint64_t tm_kernel_launch_count(void) { return g_launches.load(); }

int tm_device_info(int device, char* name, size_t name_len, int* sm_count, uint64_t* global_mem_bytes) {
    return guarded([&] {
        require_device(device);
        int dev = device;
        if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
        cudaDeviceProp p;
        CUDA_TRY(cudaGetDeviceProperties(&p, dev));
        if (name && name_len) { std::strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
        if (sm_count) *sm_count = p.multiProcessorCount;
        if (global_mem_bytes) *global_mem_bytes = p.totalGlobalMem;
    });
}

void tm_smooth_options_default(tm_smooth_options* o) {
    if (!o) return;
    std::memset(o, 0, sizeof *o);
    o->struct_size = sizeof *o;
    o->solver = TM_SOLVER_PICARD_BICGSTAB;
    o->iterations = 0;                       // input.zig:28
    o->control_function = TM_CF_LAPLACE;     // input.zig:30
    o->white_ds_target = 1e-6;
    o->white_theta_target = 0.5 * 3.14159265358979323846;  // wall_control_function.zig:61
    o->rtol = 1e-6; o->atol = 1e-8; o->max_inner_iterations = 1000;  // BiCGStab.zig:19-21
    o->omega = 1.0;
    o->sweeps_per_iteration = 1;
    o->stop_max_update = 0.0;
    o->device = -1;
}

int tm_mesh_create(const tm_block* blocks, size_t n_blocks, const tm_connection* connections, size_t n_connections,
                   const tm_condition* conditions, size_t n_conditions, int device, void* stream, tm_mesh** out) {
    if (out) *out = nullptr;
    tm_mesh* m = nullptr;
    int rc = guarded([&] {
        if (!out) TM_THROW(TM_ERR_INVALID_ARGUMENT, "out is NULL");
        if ((n_connections && !connections) || (n_conditions && !conditions)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "NULL connection / condition array");
        require_device(device);
        m = new tm_mesh();
        if (device < 0) CUDA_TRY(cudaGetDevice(&m->device)); else m->device = device;
        m->topo.build(blocks, n_blocks, connections, n_connections, conditions, n_conditions);
        if (const char* e = std::getenv("TM_TILE_ROWS")) m->tile_rows = std::max(4, std::atoi(e));
        if (const char* e = std::getenv("TM_INTERIOR")) m->use_bulk = std::strcmp(e, "regs") != 0;
        if (stream) m->stream = (cudaStream_t)stream;
        else { CUDA_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking)); m->own_stream = true; }
        CUDA_TRY(cudaEventCreate(&m->ev0));
        CUDA_TRY(cudaEventCreate(&m->ev1));
        m->N = m->topo.n_nodes;
        m->X[0].alloc(size_t(m->N));
        m->X[1].alloc(size_t(m->N));
        build_tiles(m);
        m->d_srows.upload(m->topo.smoothed, m->stream);
        m->d_jrows.upload(m->topo.junction_rows, m->stream);
        m->d_lrows.upload(m->topo.sliding, m->stream);
        m->d_slaves.upload(m->topo.slaves, m->stream);
        m->d_cslaves.upload(m->topo.const_slaves, m->stream);
        m->d_fo.upload(m->topo.fixed_overrides, m->stream);
        m->d_pairs.upload(m->topo.pairs, m->stream);
        {
            std::vector<RhsTerm> terms;
            build_rhs_terms(m, terms);
            for (const auto& c : m->topo.connected_rhs) terms.push_back(RhsTerm{c.self, c.x, c.y, 0, 0});  // smooth.zig:904-915
            m->d_rhs_terms.upload(terms, m->stream);
        }
        m->n_bnd_rows = int(m->topo.smoothed.size() + m->topo.junction_rows.size() + m->topo.sliding.size());
        m->n_bnd_ctas = bnd_ctas(m->n_bnd_rows);
        int sms = 148;
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device));
        const int64_t want = (m->N + VEC_THREADS - 1) / VEC_THREADS;
        m->vec_grid = int(std::max<int64_t>(1, std::min<int64_t>(want, int64_t(sms) * 8)));
        m->part_int.alloc(size_t(std::max(m->n_tiles, 1)) * 5);
        m->part_bnd.alloc(size_t(std::max(m->n_bnd_ctas, 1)) * 5);
        m->part_vec.alloc(size_t(m->vec_grid) * 5);
        m->part_int.zero(m->stream); m->part_bnd.zero(m->stream); m->part_vec.zero(m->stream);
        m->bconst.alloc(2); m->bconst.zero(m->stream);
        m->d_worst.alloc(1);
        m->d_ctl.alloc(1); m->d_ctl.zero(m->stream);
        CUDA_TRY(cudaMallocHost(&m->h_ctl, sizeof(SolveCtl)));
        std::memset(m->h_ctl, 0, sizeof(SolveCtl));
        m->edges.resize(n_blocks);
        m->have_coords.assign(n_blocks, 0);
        for (size_t b = 0; b < n_blocks; ++b) {
            if (blocks[b].xy) {
                CUDA_TRY(cudaMemcpyAsync(m->X[0].p + m->topo.blocks[b].off, blocks[b].xy, size_t(blocks[b].ni * blocks[b].nj) * sizeof(double2), cudaMemcpyHostToDevice, m->stream));
                m->have_coords[b] = 1;
            }
        }
        CUDA_TRY(cudaStreamSynchronize(m->stream));
        *out = m;
    });
    if (rc != TM_OK) delete m;
    return rc;
}

void tm_mesh_destroy(tm_mesh* mesh) {
    if (!mesh) return;
    cudaSetDevice(mesh->device);
    if (mesh->stream) cudaStreamSynchronize(mesh->stream);
    delete mesh;
}

int tm_mesh_upload_block(tm_mesh* m, size_t block, const double* xy) {
    return guarded([&] {
        check_block(m, block);
        if (!xy) TM_THROW(TM_ERR_INVALID_ARGUMENT, "xy is NULL");
        CUDA_TRY(cudaSetDevice(m->device));
        const auto& B = m->topo.blocks[block];
        CUDA_TRY(cudaMemcpyAsync(m->X[m->cur].p + B.off, xy, size_t(B.ni * B.nj) * sizeof(double2), cudaMemcpyHostToDevice, m->stream));
        CUDA_TRY(cudaStreamSynchronize(m->stream));
        m->have_coords[block] = 1;
        m->begun = false;
    });
}

int tm_mesh_download_block(tm_mesh* m, size_t block, double* xy) {
    return guarded([&] {
        check_block(m, block);
        if (!xy) TM_THROW(TM_ERR_INVALID_ARGUMENT, "xy is NULL");
        CUDA_TRY(cudaSetDevice(m->device));
        const auto& B = m->topo.blocks[block];
        CUDA_TRY(cudaMemcpyAsync(xy, m->X[m->cur].p + B.off, size_t(B.ni * B.nj) * sizeof(double2), cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(cudaStreamSynchronize(m->stream));
    });
}

static void tfi_launch(tm_mesh* m, size_t block) {
    const auto& B = m->topo.blocks[block];
    const int ni = int(B.ni), nj = int(B.nj);
    const double* e = m->edges[block].buf.p;
    // layout of the cache: x_i_min[2ni] x_i_max[2ni] x_j_min[2nj] x_j_max[2nj] s1[ni] s2[ni] t1[nj] t2[nj]
    const double2* x_i_min = (const double2*)e;
    const double2* x_i_max = x_i_min + ni;
    const double2* x_j_min = x_i_max + ni;
    const double2* x_j_max = x_j_min + nj;
    const double* s1 = (const double*)(x_j_max + nj);
    const double *s2 = s1 + ni, *t1 = s2 + ni, *t2 = t1 + nj;
    dim3 grid((nj + TILE_J - 1) / TILE_J, (ni + TFI_ROWS - 1) / TFI_ROWS);
    LAUNCH(tfi_kernel, grid, TILE_J, m->stream, ni, nj, x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2, m->X[m->cur].p + B.off);
    m->have_coords[block] = 1;
    m->begun = false;
}

static void tfi_validate_host(uint64_t ni, uint64_t nj, const double* x_i_min, const double* x_i_max, const double* x_j_min, const double* x_j_max,
                              const double* s1, const double* s2, const double* t1, const double* t2) {
    if (!x_i_min || !x_i_max || !x_j_min || !x_j_max || !s1 || !s2 || !t1 || !t2) TM_THROW(TM_ERR_INVALID_ARGUMENT, "tfi: NULL edge array");
    if (ni < 2 || nj < 2) TM_THROW(TM_ERR_INVALID_ARGUMENT, "tfi: a block needs at least 2x2 nodes");
    // what tfi.zig:135-162 asserts: clustering runs from exactly 0 to exactly 1, corners agree within 1e-10
    if (s1[0] != 0 || s1[ni - 1] != 1.0 || s2[0] != 0 || s2[ni - 1] != 1.0 || t1[0] != 0 || t1[nj - 1] != 1.0 || t2[0] != 0 || t2[nj - 1] != 1.0)
        TM_THROW(TM_ERR_INVALID_ARGUMENT, "tfi: clustering must start at 0 and end at 1 (tfi.zig:135-145)");
    auto near = [](const double* a, const double* b) { return std::fabs(a[0] - b[0]) <= 1e-10 && std::fabs(a[1] - b[1]) <= 1e-10; };
    if (!near(x_i_min, x_j_min) || !near(x_i_min + 2 * (ni - 1), x_j_max) || !near(x_j_min + 2 * (nj - 1), x_i_max) ||
        !near(x_i_max + 2 * (ni - 1), x_j_max + 2 * (nj - 1)))
        TM_THROW(TM_ERR_INVALID_ARGUMENT, "tfi: edge corner points are not consistent (tfi.zig:150-162)");
}

int tm_mesh_tfi_block(tm_mesh* m, size_t block, const double* x_i_min, const double* x_i_max, const double* x_j_min, const double* x_j_max,
                      const double* s1, const double* s2, const double* t1, const double* t2) {
    return guarded([&] {
        check_block(m, block);
        CUDA_TRY(cudaSetDevice(m->device));
        const auto& B = m->topo.blocks[block];
        const size_t ni = size_t(B.ni), nj = size_t(B.nj);
        tfi_validate_host(ni, nj, x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2);
        EdgeCache& ec = m->edges[block];
        const size_t total = 6 * (ni + nj);
        if (ec.buf.n != total) ec.buf.alloc(total);
        double* d = ec.buf.p;
        const double* src[8] = {x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2};
        const size_t cnt[8] = {2 * ni, 2 * ni, 2 * nj, 2 * nj, ni, ni, nj, nj};
        for (int k = 0; k < 8; ++k) {
            CUDA_TRY(cudaMemcpyAsync(d, src[k], cnt[k] * sizeof(double), cudaMemcpyHostToDevice, m->stream));
            d += cnt[k];
        }
        ec.valid = true;
        tfi_launch(m, block);
        CUDA_TRY(cudaStreamSynchronize(m->stream));  // the host edge arrays may be freed after return
    });
}

int tm_mesh_tfi_block_resident(tm_mesh* m, size_t block) {
    return guarded([&] {
        check_block(m, block);
        CUDA_TRY(cudaSetDevice(m->device));
        if (!m->edges[block].valid) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block %zu has no cached edges; call tm_mesh_tfi_block first", block);
        tfi_launch(m, block);
    });
}

int tm_mesh_begin_smoothing(tm_mesh* m, const tm_smooth_options* o) {
    return guarded([&] {
        check_mesh(m);
        validate_options(o);
        CUDA_TRY(cudaSetDevice(m->device));
        for (size_t b = 0; b < m->have_coords.size(); ++b)
            if (!m->have_coords[b]) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block %zu has no coordinates yet", b);
        cudaStream_t s = m->stream;
        double2* x = m->X[m->cur].p;
        // connectionDataCheck (smooth.zig:220-275)
        const int np = int(m->topo.pairs.size());
        if (np > 0) {
            CUDA_TRY(cudaMemsetAsync(m->d_worst.p, 0, sizeof(unsigned long long), s));
            LAUNCH(pair_check_kernel, (np + 255) / 256, 256, s, m->d_pairs.p, np, (const double2*)x, 1e-15, m->d_worst.p);
            unsigned long long worst = 0;
            CUDA_TRY(cudaMemcpyAsync(&worst, m->d_worst.p, sizeof worst, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            if (worst != 0) {
                const PairCheck& p = m->topo.pairs[size_t(worst & 0xffffffffull)];
                TM_THROW(TM_ERR_TOPOLOGY, "non matching points for connection %d point %d (tolerance 1e-15 abs, smooth.zig:220-275)", p.conn, p.point);
            }
        }
        // rhs of fixed / sliding rows is captured from the initial mesh (smooth.zig:790-796, 853-858)
        const int n_l = int(m->topo.sliding.size()), n_fo = int(m->topo.fixed_overrides.size());
        if (n_l + n_fo > 0) LAUNCH(capture_boundary_kernel, (n_l + n_fo + 127) / 128, 128, s, m->d_lrows.p, n_l, m->d_fo.p, n_fo, x);
        const int n_cs = int(m->topo.const_slaves.size());
        if (n_cs > 0) LAUNCH(sync_slaves_kernel, (n_cs + 127) / 128, 128, s, m->d_cslaves.p, n_cs, x, 1);
        CUDA_TRY(cudaMemcpyAsync(m->X[1 - m->cur].p, x, size_t(m->N) * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        // constant part of ||b||^2 (the term table was uploaded by tm_mesh_create)
        if (m->d_rhs_terms.n > 0) LAUNCH(rhs_const_kernel, 1, VEC_THREADS, s, m->d_rhs_terms.p, int(m->d_rhs_terms.n), (const double2*)x, m->bconst.p);
        else m->bconst.zero(s);
        // control function (ControlFunction.init, wall_control_function.zig:27-42)
        m->cf = int(o->control_function);
        if (m->cf == TM_CF_WHITE) {
            if (!m->topo.white_ok) TM_THROW(TM_ERR_UNSUPPORTED, "%s", m->topo.white_why.c_str());
            const auto& B0 = m->topo.blocks[0];
            const auto& B1 = m->topo.blocks[1];
            m->wp.off0 = B0.off; m->wp.off1 = B1.off;
            m->wp.ni0 = int32_t(B0.ni); m->wp.nj0 = int32_t(B0.nj); m->wp.ni1 = int32_t(B1.ni); m->wp.nj1 = int32_t(B1.nj);
            m->wp.c_in0 = int32_t(B0.nj); m->wp.c_in1 = int32_t(B1.nj); m->wp.c_al0 = 1;  // j_min sides starting at node 0, running towards +j
            m->wp.ds_target = o->white_ds_target; m->wp.theta_target = o->white_theta_target;
            if (m->pq.n != size_t(m->N)) m->pq.alloc(size_t(m->N));
            m->pq.zero(s);
            m->wall_pq.alloc(size_t(B0.ni + B1.ni));
            white_step(m, false);
        }
        m->outer_done = 0;
        m->begun = true;
        CUDA_TRY(cudaStreamSynchronize(s));
    });
}

int tm_mesh_smooth(tm_mesh* m, const tm_smooth_options* o, tm_smooth_stats* stats) {
    tm_smooth_stats st;
    std::memset(&st, 0, sizeof st);
    int rc = guarded([&] {
        check_mesh(m);
        validate_options(o);
        if (!m->begun) TM_THROW(TM_ERR_INVALID_ARGUMENT, "tm_mesh_begin_smoothing has not been called for the current coordinates");
        if (int(o->control_function) != m->cf) TM_THROW(TM_ERR_INVALID_ARGUMENT, "control function differs from the one given to tm_mesh_begin_smoothing");
        CUDA_TRY(cudaSetDevice(m->device));
        if (m->cf == TM_CF_WHITE) { m->wp.ds_target = o->white_ds_target; m->wp.theta_target = o->white_theta_target; }
        st.nodes = uint64_t(m->N);
        st.converged = 1;
        CUDA_TRY(cudaEventRecord(m->ev0, m->stream));
        if (o->solver == TM_SOLVER_RELAX) run_relax(m, o, &st);
        else run_picard_bicgstab(m, o, &st);
        CUDA_TRY(cudaEventRecord(m->ev1, m->stream));
        fetch_ctl(m);
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
        st.gpu_seconds = 1e-3 * double(ms);
        st.last_sumsq_x = m->h_ctl->sumsq[0];
        st.last_sumsq_y = m->h_ctl->sumsq[1];
        const double ss = st.last_sumsq_x + st.last_sumsq_y;
        st.last_residual = ss * ss;  // smooth.zig:136
        st.last_max_update = m->h_ctl->max_update;
        if (!st.converged && o->fail_on_no_convergence) TM_THROW(TM_ERR_NOT_CONVERGED, "inner solve did not converge (residual %.3e)", st.last_inner_residual);
    });
    if (stats) *stats = st;
    return rc;
}

int tm_mesh_synchronize(tm_mesh* m) {
    return guarded([&] {
        check_mesh(m);
        CUDA_TRY(cudaSetDevice(m->device));
        CUDA_TRY(cudaStreamSynchronize(m->stream));
    });
}

uint64_t tm_mesh_block_count(const tm_mesh* m) { return m ? m->topo.blocks.size() : 0; }
uint64_t tm_mesh_node_count(const tm_mesh* m) { return m ? uint64_t(m->N) : 0; }
int tm_mesh_block_size(const tm_mesh* m, size_t block, uint64_t* ni, uint64_t* nj) {
    return guarded([&] {
        check_block(m, block);
        if (ni) *ni = uint64_t(m->topo.blocks[block].ni);
        if (nj) *nj = uint64_t(m->topo.blocks[block].nj);
    });
}
double* tm_mesh_block_device_ptr(tm_mesh* m, size_t block) {
    if (!m || block >= m->topo.blocks.size()) return nullptr;
    return reinterpret_cast<double*>(m->X[m->cur].p + m->topo.blocks[block].off);
}
int tm_mesh_download_control_function(tm_mesh* m, size_t block, double* pqv) {
    return guarded([&] {
        check_block(m, block);
        if (!pqv) TM_THROW(TM_ERR_INVALID_ARGUMENT, "pq is NULL");
        CUDA_TRY(cudaSetDevice(m->device));
        const auto& B = m->topo.blocks[block];
        const size_t bytes = size_t(B.ni * B.nj) * sizeof(double2);
        if (m->cf != TM_CF_WHITE || !m->pq.p) { std::memset(pqv, 0, bytes); return; }  // laplace: all zero (wall_control_function.zig:29-33)
        CUDA_TRY(cudaMemcpyAsync(pqv, m->pq.p + B.off, bytes, cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(cudaStreamSynchronize(m->stream));
    });
}
int tm_mesh_download_boundary_kinds(tm_mesh* m, size_t block, uint8_t* kinds) {
    return guarded([&] {
        check_block(m, block);
        if (!kinds) TM_THROW(TM_ERR_INVALID_ARGUMENT, "kinds is NULL");
        const auto& B = m->topo.blocks[block];
        std::memcpy(kinds, m->topo.kind.data() + B.bbuf, size_t(2 * (B.ni + B.nj - 2)));
    });
}

int tm_tfi_block(uint64_t ni, uint64_t nj, const double* x_i_min, const double* x_i_max, const double* x_j_min, const double* x_j_max, const double* s1,
                 const double* s2, const double* t1, const double* t2, double* out_xy) {
    tm_mesh* m = nullptr;
    int rc = guarded([&] {
        if (!out_xy) TM_THROW(TM_ERR_INVALID_ARGUMENT, "out_xy is NULL");
        if (ni < 3 || nj < 3) TM_THROW(TM_ERR_INVALID_ARGUMENT, "tfi: block smaller than 3x3 nodes");
    });
    if (rc != TM_OK) return rc;
    tm_block blk{ni, nj, nullptr};
    rc = tm_mesh_create(&blk, 1, nullptr, 0, nullptr, 0, -1, nullptr, &m);
    if (rc == TM_OK) rc = tm_mesh_tfi_block(m, 0, x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2);
    if (rc == TM_OK) rc = tm_mesh_download_block(m, 0, out_xy);
    std::string keep = g_last_error;
    tm_mesh_destroy(m);
    g_last_error = keep;
    return rc;
}

int tm_smooth_mesh(tm_block* blocks, size_t n_blocks, const tm_connection* connections, size_t n_connections, const tm_condition* conditions,
                   size_t n_conditions, const tm_smooth_options* opts, tm_smooth_stats* stats) {
    tm_mesh* m = nullptr;
    int rc = guarded([&] {
        validate_options(opts);
        if (!blocks) TM_THROW(TM_ERR_INVALID_ARGUMENT, "blocks is NULL");
        for (size_t b = 0; b < n_blocks; ++b)
            if (!blocks[b].xy) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block %zu has no coordinates", b);
    });
    if (rc != TM_OK) return rc;
    rc = tm_mesh_create(blocks, n_blocks, connections, n_connections, conditions, n_conditions, opts->device, nullptr, &m);
    if (rc == TM_OK) rc = tm_mesh_begin_smoothing(m, opts);
    if (rc == TM_OK) rc = tm_mesh_smooth(m, opts, stats);
    if (rc == TM_OK || rc == TM_ERR_NOT_CONVERGED) {
        for (size_t b = 0; b < n_blocks; ++b) {
            const int rc2 = tm_mesh_download_block(m, b, blocks[b].xy);
            if (rc2 != TM_OK) { rc = rc2; break; }
        }
    }
    std::string keep = g_last_error;
    tm_mesh_destroy(m);
    g_last_error = keep;
    return rc;
}

}  // extern "C"
#pragma GCC visibility pop
