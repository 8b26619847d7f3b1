// streamed.inl -- tm_smooth_mesh on a large single block with HOST buffers: PCIe copies overlapped with the sweeps.
//
// The one-shot entry point (what smooth.mesh, smooth.zig:74-166, binds to) is bound by the host link for a large block:
// 1 GB up, 100 sweeps, 1 GB down are 19 + 35 + 19 ms when they run one after the other.  A damped-Jacobi sweep is
// local -- after T sweeps a node depends on the initial values within T rows of it only -- so a block whose boundary
// nodes are all fixed is cut into K row chunks; each chunk travels with T extra rows on either side (a window), gets T
// sweeps as a single-block mesh of its own (the window's artificial first / last row is "fixed" and goes stale, the
// error moves inwards by exactly one row per sweep and never reaches the rows the chunk owns), and only the owned rows
// travel back.  Windows rotate through three device meshes on three streams, so the upload of chunk k+1, the sweeps of
// chunk k and the download of chunk k-1 overlap; every node gets exactly the arithmetic of the resident path, the result
// is bit-identical (tests/test_gpu_first.py::test_streamed_host_smoothing_is_bit_identical).  Price: redundant rows --
// about T/C of the block, because sweep t skips the tile rows within t rows of an artificial edge: they are stale
// already and nothing the chunk owns can see them any more.
// The whole pipeline is queued without a host-side wait (events order the copies that share host rows, see below); the
// three window meshes are parked between calls like the large device buffers (tm_release_cached_memory frees them).
//
// Taken when: one block, no connections, no inlet / outlet condition, TM_SOLVER_RELAX with the Laplace control function
// and a fixed sweep count (stop_max_update == 0), at least TM_STREAM_MIN_NODES nodes (default 8 Mi) and chunks of at
// least 4T rows.  TM_STREAM=0 switches it off, TM_STREAM_CHUNKS overrides K (default: up to 8).

struct StreamPlan {
    int64_t T = 0, W = 0;                 // sweeps in total; rows per window
    std::vector<int64_t> w0, o0;          // first row of window k; first row owned by chunk k (o0[K] = ni)
    int K() const { return int(w0.size()); }
};

// equally sized windows, evenly spaced over the block; the owned ranges meet in the middle of the overlaps
inline bool plan_streaming(int64_t ni, int64_t nj, int64_t sweeps, StreamPlan& P) {
    int64_t min_nodes = int64_t(8) << 20;
    if (const char* e = std::getenv("TM_STREAM")) if (std::atoi(e) == 0) return false;
    if (const char* e = std::getenv("TM_STREAM_MIN_NODES")) min_nodes = std::atoll(e);
    if (sweeps < 1 || ni * nj < min_nodes) return false;
    int64_t K = std::min<int64_t>(8, ni / (4 * sweeps));
    if (const char* e = std::getenv("TM_STREAM_CHUNKS")) K = std::min<int64_t>(K, std::atoll(e));
    if (K < 2) return false;
    const int64_t T = sweeps;
    const int64_t W = (ni + (2 * T + 2) * (K - 1) + K - 1) / K;
    if (W >= ni || W < 3 || W * nj >= (int64_t(1) << 31)) return false;  // a window is a block of its own: < 2^31 nodes
    P.T = T; P.W = W;
    P.w0.resize(size_t(K)); P.o0.assign(size_t(K) + 1, 0);
    for (int64_t k = 0; k < K; ++k) P.w0[size_t(k)] = (k * (ni - W)) / (K - 1);
    P.o0[size_t(K)] = ni;
    for (int64_t k = 1; k < K; ++k) {
        const int64_t lo = P.w0[size_t(k)], hi = P.w0[size_t(k - 1)] + W;  // rows both windows hold
        if (hi - lo < 2 * T) return false;
        P.o0[size_t(k)] = lo + (hi - lo) / 2;
    }
    for (int64_t k = 0; k < K; ++k) {
        const int64_t a = P.o0[size_t(k)], b = P.o0[size_t(k + 1)], lo = P.w0[size_t(k)], hi = lo + W;
        if (!(a < b)) return false;
        if (lo > 0 && a - lo < T) return false;        // T rows between an artificial window edge and the owned rows
        if (hi < ni && hi - b < T) return false;
        if (k >= 2 && lo < P.o0[size_t(k - 1)]) return false;  // a window reads owned rows of its direct neighbours only
    }
    return true;
}

inline bool can_stream(const tm_block* blocks, size_t n_blocks, size_t n_connections, const tm_condition* conditions, size_t n_conditions,
                       const tm_smooth_options* o, StreamPlan& P) {
    if (n_blocks != 1 || n_connections != 0) return false;
    for (size_t c = 0; c < n_conditions; ++c)
        if (conditions[c].kind != TM_BC_WALL) return false;
    if (o->solver != TM_SOLVER_RELAX || o->control_function != TM_CF_LAPLACE || o->stop_max_update > 0.0) return false;
    if (o->iterations == 0 || o->sweeps_per_iteration > (uint64_t(1) << 20) || o->iterations > (uint64_t(1) << 20)) return false;
    return plan_streaming(int64_t(blocks[0].ni), int64_t(blocks[0].nj), int64_t(o->iterations * o->sweeps_per_iteration), P);
}

// the three window meshes, their events and the pinned statistics of a call; parked between calls
struct StreamSlots {
    static constexpr int SLOTS = 3, MAX_CHUNKS = 8;
    int device = -1;
    int64_t W = 0, nj = 0;
    tm_mesh* slot[SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_up[SLOTS] = {nullptr, nullptr, nullptr}, ev_done[SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    SolveCtl* h_stats = nullptr;  // pinned, one entry per chunk
    ~StreamSlots() {
        for (tm_mesh* m : slot) tm_mesh_destroy(m);
        for (cudaEvent_t e : ev_up) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : ev_done) if (e) cudaEventDestroy(e);
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_end) cudaEventDestroy(ev_end);
        if (h_stats) cudaFreeHost(h_stats);
    }
};
std::mutex g_slots_mu;
StreamSlots* g_slots_parked = nullptr;  // most recent set; never destroyed at process exit (the CUDA context may be gone)

inline std::unique_ptr<StreamSlots> acquire_slots(int device, int64_t W, int64_t nj) {
    {
        std::lock_guard<std::mutex> lock(g_slots_mu);
        if (g_slots_parked && g_slots_parked->device == device && g_slots_parked->W == W && g_slots_parked->nj == nj) {
            std::unique_ptr<StreamSlots> q(g_slots_parked);
            g_slots_parked = nullptr;
            return q;
        }
    }
    std::unique_ptr<StreamSlots> S(new StreamSlots());
    S->device = device; S->W = W; S->nj = nj;
    const tm_block window{uint64_t(W), uint64_t(nj), nullptr};
    for (int s = 0; s < StreamSlots::SLOTS; ++s) {
        const int rc = tm_mesh_create(&window, 1, nullptr, 0, nullptr, 0, device, nullptr, &S->slot[s]);
        if (rc != TM_OK) throw Error{rc, g_last_error};
        S->slot[s]->cf = TM_CF_LAPLACE;
        S->slot[s]->ranks[0]->has_pq = false;
        S->slot[s]->ranks[0]->have_coords[0] = 1;
        CUDA_TRY(cudaEventCreateWithFlags(&S->ev_up[s], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&S->ev_done[s], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreate(&S->ev_begin));
    CUDA_TRY(cudaEventCreate(&S->ev_end));
    CUDA_TRY(cudaMallocHost(&S->h_stats, sizeof(SolveCtl) * StreamSlots::MAX_CHUNKS));
    return S;
}
inline void park_slots(std::unique_ptr<StreamSlots> S) {
    StreamSlots* old = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_slots_mu);
        old = g_slots_parked;
        g_slots_parked = S.release();
    }
    delete old;
}
inline void release_parked_slots() {
    StreamSlots* old = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_slots_mu);
        old = g_slots_parked;
        g_slots_parked = nullptr;
    }
    delete old;
}

inline void smooth_streamed(tm_block* blocks, const tm_smooth_options* o, const StreamPlan& P, tm_smooth_stats* stats) {
    constexpr int SLOTS = StreamSlots::SLOTS;
    const bool trace = std::getenv("TM_STREAM_TRACE") != nullptr;  // host-side phase times on stderr (tuning aid)
    const bool serial = std::getenv("TM_STREAM_SERIAL") != nullptr && std::atoi(std::getenv("TM_STREAM_SERIAL")) != 0;
    const auto t_start = std::chrono::steady_clock::now();
    auto ms_since_start = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    const int64_t ni = int64_t(blocks[0].ni), nj = int64_t(blocks[0].nj), W = P.W, T = P.T;
    const int K = P.K();
    if (K > StreamSlots::MAX_CHUNKS) TM_THROW(TM_ERR_INVALID_ARGUMENT, "internal: %d chunks", K);
    double2* host = reinterpret_cast<double2*>(blocks[0].xy);
    int device = o->device;
    require_device(device);
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    std::unique_ptr<StreamSlots> S = acquire_slots(device, W, nj);
    if (trace) std::fprintf(stderr, "[stream] %d chunks, windows of %lld rows; slots ready at %.2f ms\n", K, (long long)W, ms_since_start());
    auto drain = [&] { for (tm_mesh* m : S->slot) if (m && m->stream) cudaStreamSynchronize(m->stream); };
    try {
        // tile rows of a window mesh (build_rank: i0-major tile list, `rows` rows per tile row)
        RankMesh& r0 = *S->slot[0]->ranks[0];
        const int64_t interior = W - 2, n_i = std::max<int64_t>(1, (interior + S->slot[0]->tile_rows - 1) / S->slot[0]->tile_rows);
        const int64_t rows = (interior + n_i - 1) / n_i, n_tile_rows = (interior + rows - 1) / rows, n_cols = (nj - 2 + TILE_J - 1) / TILE_J;
        const bool trapezoid = S->slot[0]->use_bulk && r0.n_bnd_ctas == 0 && int64_t(r0.n_tiles) == n_tile_rows * n_cols && std::getenv("TM_STREAM_FULL_WINDOWS") == nullptr;
        CUDA_TRY(cudaEventRecord(S->ev_begin, S->slot[0]->stream));
        for (int k = 0; k < K; ++k) {
            const int s = k % SLOTS;
            tm_mesh* m = S->slot[s];
            RankMesh& r = *m->ranks[0];
            const int64_t w0 = P.w0[size_t(k)], a = P.o0[size_t(k)], b = P.o0[size_t(k + 1)];
            // upload of the window; on this stream it follows the download of chunk k - SLOTS, which read the same buffers
            CUDA_TRY(cudaMemcpyAsync(r.X[r.cur].p, host + w0 * nj, size_t(W * nj) * sizeof(double2), cudaMemcpyHostToDevice, m->stream));
            CUDA_TRY(cudaEventRecord(S->ev_up[s], m->stream));
            // the sweeps write interior nodes only: the other buffer needs the fixed boundary nodes as well
            CUDA_TRY(cudaMemcpyAsync(r.X[1 - r.cur].p, r.X[r.cur].p, size_t(W * nj) * sizeof(double2), cudaMemcpyDeviceToDevice, m->stream));
            if (k > 0) {
                // the window of chunk k holds rows chunk k-1 owns: only once it has been read may the results of chunk k-1
                // overwrite them on the host (windows read owned rows of their direct neighbours only, plan_streaming)
                tm_mesh* mp = S->slot[(k - 1) % SLOTS];
                RankMesh& rp = *mp->ranks[0];
                const int64_t pa = P.o0[size_t(k - 1)], pb = P.o0[size_t(k)];
                CUDA_TRY(cudaStreamWaitEvent(mp->stream, S->ev_up[s], 0));
                CUDA_TRY(cudaMemcpyAsync(host + pa * nj, rp.X[rp.cur].p + (pa - P.w0[size_t(k - 1)]) * nj, size_t((pb - pa) * nj) * sizeof(double2), cudaMemcpyDeviceToHost, mp->stream));
                // TM_STREAM_SERIAL=1: one chunk is swept at a time.  Default: the sweeps of two uploaded chunks may share the
                // device -- a window is only ~2 waves of CTAs, and the tail of one chunk's sweep is filled by the other's
                if (serial) CUDA_TRY(cudaStreamWaitEvent(m->stream, S->ev_done[(k - 1) % SLOTS], 0));
            }
            for (int64_t t = 1; t <= T; ++t) {
                RowPart part;
                if (trapezoid) {  // rows within t of an artificial edge are stale and out of reach of the owned rows
                    const int64_t lo = w0 > 0 ? (t - 1) / rows : 0;
                    const int64_t hi = w0 + W < ni ? std::min<int64_t>(n_tile_rows, std::max<int64_t>(0, W - 2 - t) / rows + 1) : n_tile_rows;
                    part.first = int(lo * n_cols); part.count = int(std::max<int64_t>(0, hi - lo) * n_cols);
                }
                launch_rows<MODE_RELAX, 0>(m, r, false, r.X[r.cur].p, r.X[r.cur].p, r.X[1 - r.cur].p, o->omega, nullptr, part);
                r.cur = 1 - r.cur;
            }
            // statistics of the last sweep over the owned rows (X[cur] = iterate T, X[1-cur] = iterate T-1)
            LAUNCH(diff_stats_kernel, r.vec_grid, VEC_THREADS, m->stream, (b - a) * nj, (const double2*)r.X[r.cur].p + (a - w0) * nj,
                   (const double2*)r.X[1 - r.cur].p + (a - w0) * nj, r.part_vec.p);
            launch_reduce(m, RED_UPDATE_STATS, o, false);
            CUDA_TRY(cudaEventRecord(S->ev_done[s], m->stream));
            CUDA_TRY(cudaMemcpyAsync(S->h_stats + k, r.d_ctl.p, sizeof(SolveCtl), cudaMemcpyDeviceToHost, m->stream));
        }
        {
            tm_mesh* mp = S->slot[(K - 1) % SLOTS];
            RankMesh& rp = *mp->ranks[0];
            const int64_t pa = P.o0[size_t(K - 1)], pb = P.o0[size_t(K)];
            CUDA_TRY(cudaMemcpyAsync(host + pa * nj, rp.X[rp.cur].p + (pa - P.w0[size_t(K - 1)]) * nj, size_t((pb - pa) * nj) * sizeof(double2), cudaMemcpyDeviceToHost, mp->stream));
            for (int s = 0; s < SLOTS; ++s) {  // the last stream joins the others, then carries the end mark
                if (S->slot[s] == mp) continue;
                CUDA_TRY(cudaEventRecord(S->ev_up[s], S->slot[s]->stream));
                CUDA_TRY(cudaStreamWaitEvent(mp->stream, S->ev_up[s], 0));
            }
            CUDA_TRY(cudaEventRecord(S->ev_end, mp->stream));
            if (trace) std::fprintf(stderr, "[stream] everything queued at %.2f ms\n", ms_since_start());
            CUDA_TRY(cudaStreamSynchronize(mp->stream));
        }
        if (trace) std::fprintf(stderr, "[stream] drained at %.2f ms\n", ms_since_start());
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, S->ev_begin, S->ev_end));
        tm_smooth_stats st;
        std::memset(&st, 0, sizeof st);
        for (int k = 0; k < K; ++k) {
            st.last_sumsq_x += S->h_stats[k].sumsq[0]; st.last_sumsq_y += S->h_stats[k].sumsq[1];
            st.last_max_update = std::fmax(st.last_max_update, S->h_stats[k].max_update);
        }
        st.outer_iterations = o->iterations;
        st.inner_iterations = st.operator_applications = uint64_t(T);
        st.nodes = uint64_t(ni * nj);
        st.last_residual = (st.last_sumsq_x + st.last_sumsq_y) * (st.last_sumsq_x + st.last_sumsq_y);  // smooth.zig:136
        st.gpu_seconds = 1e-3 * double(ms);
        st.converged = 1;
        st.streamed_chunks = K;
        if (stats) *stats = st;
    } catch (...) {
        drain();  // nothing may still write the host block; the slots are not parked again
        throw;
    }
    park_slots(std::move(S));
    if (trace) std::fprintf(stderr, "[stream] slots parked at %.2f ms\n", ms_since_start());
}
