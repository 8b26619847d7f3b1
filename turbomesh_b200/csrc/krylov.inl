// krylov.inl -- host side of the persistent BiCGStab kernel (krylov_kernels.cuh): the plan (components, warp tiles, CTA
// groups) and one launch per outer iteration.  Textually included into the anonymous namespace of turbomesh_gpu.cu.
// Single-rank meshes only (one process, one GPU -- the reference's configurations and the batches of cuts, which need no
// communication); meshes spread over several ranks keep the launch-per-phase path of run_picard_bicgstab.

bool krylov_persistent_possible(const tm_mesh* m) {
    if (m->n_ranks != 1 || m->emulated) return false;
    if (const char* e = std::getenv("TM_KRYLOV")) if (std::strcmp(e, "launches") == 0) return false;   // round 1's launch-per-vector-operation path (kept for ranks > 1)
    int coop = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, m->device) != cudaSuccess || !coop) return false;
    return true;
}

void krylov_plan_build(tm_mesh* m, RankMesh& r) {
    if (r.kplan && r.kplan->built_pq == r.has_pq) return;
    r.kplan.reset(new KrylovPlan());
    KrylovPlan& P = *r.kplan;
    P.built_pq = r.has_pq;
    cudaStream_t s = m->stream;
    const Topology& T = m->topo;
    const int n_comp = T.n_comp;
    // how many CTAs can be co-resident (cooperative launch)
    int per_sm = 0;
    if (r.has_pq) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bicgstab_persistent_kernel<true>, K_THREADS, 0));
    else CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bicgstab_persistent_kernel<false>, K_THREADS, 0));
    if (per_sm < 1) TM_THROW(TM_ERR_CUDA, "the persistent Krylov kernel does not fit on an SM");
    if (const char* e = std::getenv("TM_KRYLOV_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, std::atoi(e)));
    const int max_ctas = per_sm * m->sm_count;
    std::vector<int64_t> nodes(size_t(n_comp), 0);
    for (size_t b = 0; b < T.blocks.size(); ++b) nodes[size_t(T.comp_of_block[b])] += T.blocks[b].ni * T.blocks[b].nj;
    const int64_t largest = *std::max_element(nodes.begin(), nodes.end());
    // Groups.  A solve touches ~10 fields of 16 B per node; the components in flight should fit in L2 together, and a CTA
    // wants about a thousand nodes (a few rows per warp and phase) -- more CTAs per component only buy barrier latency.
    int l2_bytes = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, m->device));
    double l2_fraction = 0.75;
    if (const char* e = std::getenv("TM_KRYLOV_L2_FRACTION")) l2_fraction = std::atof(e);
    int nodes_per_cta = 256;
    if (const char* e = std::getenv("TM_KRYLOV_NODES_PER_CTA")) nodes_per_cta = std::max(64, std::atoi(e));
    const int want_ctas = int(std::max<int64_t>(1, std::min<int64_t>(max_ctas, (largest + nodes_per_cta - 1) / nodes_per_cta)));
    const int fit_l2 = int(std::max(1.0, l2_fraction * double(l2_bytes) / (160.0 * double(largest))));
    int n_groups = std::max(1, std::min(std::min(n_comp, fit_l2), max_ctas / std::min(want_ctas, 4)));
    if (const char* e = std::getenv("TM_KRYLOV_GROUPS")) n_groups = std::max(1, std::min(std::min(n_comp, max_ctas), std::atoi(e)));
    {   // fewer groups that finish in the same number of rounds leave more CTAs to each
        const int rounds = (n_comp + n_groups - 1) / n_groups;
        n_groups = (n_comp + rounds - 1) / rounds;
    }
    int group_ctas = std::max(1, std::min(want_ctas, max_ctas / n_groups));
    if (const char* e = std::getenv("TM_KRYLOV_GROUP_CTAS")) group_ctas = std::max(1, std::min(max_ctas / n_groups, std::atoi(e)));
    P.n_groups = n_groups; P.group_ctas = group_ctas; P.n_ctas = n_groups * group_ctas;
    // components to groups: largest first onto the least loaded group
    std::vector<int32_t> order((size_t)n_comp), group_of((size_t)n_comp, 0);
    for (int c = 0; c < n_comp; ++c) order[size_t(c)] = c;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return nodes[size_t(x)] > nodes[size_t(y)]; });
    std::vector<int64_t> load(size_t(n_groups), 0);
    for (int32_t c : order) {
        const size_t g = size_t(std::min_element(load.begin(), load.end()) - load.begin());
        group_of[size_t(c)] = int32_t(g);
        load[g] += nodes[size_t(c)];
    }
    std::vector<KGroup> groups((size_t)n_groups);
    std::vector<int32_t> group_comps, cta_group((size_t)P.n_ctas);
    for (int g = 0; g < n_groups; ++g) {
        groups[size_t(g)].comp_begin = int32_t(group_comps.size());
        for (int c = 0; c < n_comp; ++c) if (group_of[size_t(c)] == g) group_comps.push_back(c);
        groups[size_t(g)].comp_end = int32_t(group_comps.size());
        groups[size_t(g)].cta_begin = g * group_ctas;
        groups[size_t(g)].n_ctas = group_ctas;
        for (int k = 0; k < group_ctas; ++k) cta_group[size_t(g * group_ctas + k)] = g;
    }
    // warp tiles per component: rows chosen so that every warp of the group gets about two tiles per phase
    std::vector<KComp> comps((size_t)n_comp);
    std::vector<WTile> wtiles;
    const int group_warps = group_ctas * K_WARPS;
    for (int c = 0; c < n_comp; ++c) {
        KComp& K = comps[size_t(c)];
        K = KComp{};
        K.nodes = int32_t(std::min<int64_t>(nodes[size_t(c)], 0x7fffffff));
        int rows = int(std::max<int64_t>(1, std::min<int64_t>(K_TILE_ROWS, nodes[size_t(c)] / (32 * int64_t(group_warps)))));
        if (const char* e = std::getenv("TM_KRYLOV_TILE_ROWS")) rows = std::max(1, std::min(K_TILE_ROWS, std::atoi(e)));
        K.wt_begin = int32_t(wtiles.size());
        for (size_t b = 0; b < T.blocks.size(); ++b) {
            if (T.comp_of_block[b] != c) continue;
            const auto& B = T.blocks[b];
            const int64_t interior_i = B.ni - 2;
            const int64_t n_i = std::max<int64_t>(1, (interior_i + rows - 1) / rows);
            const int64_t rr = (interior_i + n_i - 1) / n_i;
            for (int64_t i0 = 1; i0 <= B.ni - 2; i0 += rr)
                for (int64_t j0 = 1; j0 <= B.nj - 2; j0 += 32) wtiles.push_back(WTile{int32_t(b), int32_t(i0), int32_t(j0), int32_t(std::min<int64_t>(rr, B.ni - 1 - i0))});
        }
        K.wt_end = int32_t(wtiles.size());
    }
    // boundary rows / rhs terms are grouped by component already (build_rank): find the ranges
    auto ranges = [&](auto& rows, auto node_of, auto set) {
        size_t k = 0;
        for (int c = 0; c < n_comp; ++c) {
            const size_t b = k;
            while (k < rows.size() && T.comp_of_block[T.block_of(node_of(rows[k]))] == c) ++k;
            set(comps[size_t(c)], int32_t(b), int32_t(k));
        }
        if (k != rows.size()) TM_THROW(TM_ERR_TOPOLOGY, "internal: boundary rows are not grouped by component");
    };
    ranges(r.L.smoothed, [](const SmoothedRow& x) { return x.g0; }, [](KComp& K, int32_t b, int32_t e) { K.s_begin = b; K.s_end = e; });
    ranges(r.L.junction_rows, [](const JunctionRow& x) { return x.self; }, [](KComp& K, int32_t b, int32_t e) { K.j_begin = b; K.j_end = e; });
    ranges(r.L.sliding, [](const SlidingRow& x) { return x.self; }, [](KComp& K, int32_t b, int32_t e) { K.l_begin = b; K.l_end = e; });
    ranges(r.L.rhs_terms, [](const RhsTerm& x) { return x.g; }, [](KComp& K, int32_t b, int32_t e) { K.rt_begin = b; K.rt_end = e; });
    P.h_comps = comps;
    P.h_ctl.assign(size_t(n_comp), KCtl{});
    P.wtiles.upload(wtiles, s);
    P.comps.upload(comps, s);
    P.groups.upload(groups, s);
    P.group_comps.upload(group_comps, s);
    P.cta_group.upload(cta_group, s);
    P.ctl.alloc(size_t(n_comp)); P.ctl.zero(s);
    P.bars.alloc(size_t(n_groups)); P.bars.zero(s);
    P.partials.alloc(size_t(2) * size_t(P.n_ctas) * K_NACC); P.partials.zero(s);
    CUDA_TRY(cudaStreamSynchronize(s));
}

// One outer (Picard) iteration: every component solved for x and y by the persistent kernel.  X[cur] = lagged mesh,
// X[1 - cur] = the new iterate (warm start = the mesh, GMRES.zig:157-174).
void krylov_solve_persistent(tm_mesh* m, RankMesh& r, const tm_smooth_options* o, tm_smooth_stats* st) {
    krylov_plan_build(m, r);
    KrylovPlan& P = *r.kplan;
    cudaStream_t s = m->stream;
    KArgs a{};
    a.wtiles = P.wtiles.p; a.blocks = r.d_blocks.p;
    a.srows = r.d_srows.p; a.jrows = r.d_jrows.p; a.lrows = r.d_lrows.p; a.slaves = r.d_slaves.p; a.rterms = r.d_rhs_terms.p;
    a.comps = P.comps.p; a.group_comps = P.group_comps.p; a.groups = P.groups.p; a.cta_group = P.cta_group.p;
    a.ctl = P.ctl.p; a.bars = P.bars.p; a.partials = P.partials.p;
    a.xc = r.X[r.cur].p; a.pq = r.pq.p; a.xnew = r.X[1 - r.cur].p;
    a.r = r.kr.p; a.rhat = r.krhat.p; a.p[0] = r.kp.p; a.p[1] = r.kp2.p; a.v[0] = r.kv.p; a.v[1] = r.kv2.p; a.s = r.ks.p; a.t = r.kt.p; a.d = r.kd.p;
    a.rtol = o->rtol; a.atol = o->atol;
    a.max_iters = o->max_inner_iterations > 0x7fffffffull ? 0x7fffffff : int32_t(o->max_inner_iterations);
    a.max_restarts = 60;
    a.n_ctas_total = P.n_ctas;
    a.polish = std::max(0, int(o->inner_refinement_cycles));
    void* params[] = {&a};
    const void* fn = r.has_pq ? (const void*)bicgstab_persistent_kernel<true> : (const void*)bicgstab_persistent_kernel<false>;
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(unsigned(P.n_ctas)), dim3(K_THREADS), params, 0, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaMemcpyAsync(P.h_ctl.data(), P.ctl.p, P.h_ctl.size() * sizeof(KCtl), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    double weighted_apps = 0.0, worst = 0.0;
    for (size_t c = 0; c < P.h_ctl.size(); ++c) {
        const KCtl& k = P.h_ctl[c];
        st->inner_iterations += uint64_t(k.iters[0]) + uint64_t(k.iters[1]);
        weighted_apps += double(k.applications) * double(P.h_comps[c].nodes);
        worst = std::fmax(worst, std::fmax(k.norm_r[0] / std::fmax(k.tol[0], 1e-300), k.norm_r[1] / std::fmax(k.tol[1], 1e-300)));
        st->last_inner_residual = std::fmax(c == 0 ? 0.0 : st->last_inner_residual, std::fmax(k.norm_r[0], k.norm_r[1]));
        if (k.done[0] != 1 || k.done[1] != 1) st->converged = 0;  // log.warn "did not converge", BiCGStab.zig:368-369
    }
    st->operator_applications += uint64_t(weighted_apps / double(std::max<int64_t>(m->topo.n_nodes, 1)) + 0.5);
    (void)worst;
}

// ---------------------------------------------------------------------------------------------------------------
// Phased form (krylov_phased.cuh): one launch per phase over all components, scalars per component; the default.
// ---------------------------------------------------------------------------------------------------------------
void phased_plan_build(tm_mesh* m, RankMesh& r) {
    if (r.pplan) return;
    r.pplan.reset(new PhasedPlan());
    PhasedPlan& P = *r.pplan;
    cudaStream_t s = m->stream;
    const Topology& T = m->topo;
    const int n_comp = T.n_comp;
    std::vector<KPComp> comps((size_t)n_comp);
    std::vector<WTile> wtiles;
    std::vector<int32_t> wt_comp;
    std::vector<BChunk> chunks;
    int tile_rows = T.n_nodes >= 2000000 ? 16 : (T.n_nodes >= 200000 ? 8 : 2);
    if (const char* e = std::getenv("TM_KRYLOV_TILE_ROWS")) tile_rows = std::max(1, std::atoi(e));
    std::vector<std::vector<size_t>> blocks_of((size_t)n_comp);
    for (size_t b = 0; b < T.blocks.size(); ++b) blocks_of[size_t(T.comp_of_block[b])].push_back(b);
    // row ranges per component (the tables are grouped by component, build_rank)
    std::vector<std::array<int32_t, 8>> rng((size_t)n_comp);
    auto ranges = [&](auto& rows, auto node_of, int slot) {
        size_t k = 0;
        for (int c = 0; c < n_comp; ++c) {
            const size_t b = k;
            while (k < rows.size() && T.comp_of_block[T.block_of(node_of(rows[k]))] == c) ++k;
            rng[size_t(c)][size_t(slot)] = int32_t(b); rng[size_t(c)][size_t(slot) + 1] = int32_t(k);
        }
        if (k != rows.size()) TM_THROW(TM_ERR_TOPOLOGY, "internal: boundary rows are not grouped by component");
    };
    ranges(r.L.smoothed, [](const SmoothedRow& x) { return x.g0; }, 0);
    ranges(r.L.junction_rows, [](const JunctionRow& x) { return x.self; }, 2);
    ranges(r.L.sliding, [](const SlidingRow& x) { return x.self; }, 4);
    ranges(r.L.rhs_terms, [](const RhsTerm& x) { return x.g; }, 6);
    for (int c = 0; c < n_comp; ++c) {
        KPComp& K = comps[size_t(c)];
        K = KPComp{};
        K.wt_begin = int32_t(wtiles.size());
        int64_t nodes = 0;
        for (size_t b : blocks_of[size_t(c)]) {
            const auto& B = T.blocks[b];
            nodes += B.ni * B.nj;
            // tall tiles for bandwidth (marching window), short ones when the whole mesh is small and launch-bound anyway
            const int64_t interior_i = B.ni - 2;
            const int64_t n_i = std::max<int64_t>(1, (interior_i + tile_rows - 1) / tile_rows);
            const int64_t rr = (interior_i + n_i - 1) / n_i;
            for (int64_t i0 = 1; i0 <= B.ni - 2; i0 += rr)
                for (int64_t j0 = 1; j0 <= B.nj - 2; j0 += 32) {
                    wtiles.push_back(WTile{int32_t(b), int32_t(i0), int32_t(j0), int32_t(std::min<int64_t>(rr, B.ni - 1 - i0))});
                    wt_comp.push_back(c);
                }
        }
        K.wt_end = int32_t(wtiles.size());
        K.nodes = int32_t(std::min<int64_t>(nodes, 0x7fffffff));
        K.rt_begin = rng[size_t(c)][6]; K.rt_end = rng[size_t(c)][7];
        // boundary rows in chunks of <= KP_THREADS rows (smoothed | junction | sliding); at least one chunk per component
        K.ch_begin = int32_t(chunks.size());
        int32_t cur[3] = {rng[size_t(c)][0], rng[size_t(c)][2], rng[size_t(c)][4]};
        const int32_t end[3] = {rng[size_t(c)][1], rng[size_t(c)][3], rng[size_t(c)][5]};
        do {
            BChunk ch{};
            ch.comp = c;
            int room = KP_THREADS;
            int32_t* b[3] = {&ch.s_begin, &ch.j_begin, &ch.l_begin};
            int32_t* e[3] = {&ch.s_end, &ch.j_end, &ch.l_end};
            for (int k = 0; k < 3; ++k) {
                const int take = std::min<int>(room, end[k] - cur[k]);
                *b[k] = cur[k]; *e[k] = cur[k] + take;
                cur[k] += take; room -= take;
            }
            chunks.push_back(ch);
        } while (cur[0] < end[0] || cur[1] < end[1] || cur[2] < end[2]);
        K.ch_end = int32_t(chunks.size());
    }
    P.n_comp = n_comp; P.n_wtiles = int(wtiles.size()); P.n_chunks = int(chunks.size());
    P.h_comps = comps;
    P.h_state.assign((size_t)n_comp, KState{});
    P.wtiles.upload(wtiles, s);
    P.wt_comp.upload(wt_comp, s);
    P.chunks.upload(chunks, s);
    P.comps.upload(comps, s);
    P.state.alloc((size_t)n_comp); P.state.zero(s);
    P.partials.alloc(size_t(P.n_wtiles + P.n_chunks) * K_NACC); P.partials.zero(s);
    P.count.alloc(2); P.count.zero(s);
    CUDA_TRY(cudaMallocHost(&P.h_count, 2 * sizeof(int)));
    CUDA_TRY(cudaStreamSynchronize(s));
}

template <int PHASE>
void phased_launch(tm_mesh* m, RankMesh& r, const KPArgs& a) {
    const PhasedPlan& P = *r.pplan;
    const unsigned g_tiles = unsigned((P.n_wtiles + KP_WARPS - 1) / KP_WARPS), g_chunks = unsigned(P.n_chunks);
    if (r.has_pq) {
        if (g_tiles) LAUNCH((krylov_phase_kernel<PHASE, true, true>), g_tiles, KP_THREADS, m->stream, a);
        if (g_chunks) LAUNCH((krylov_phase_kernel<PHASE, true, false>), g_chunks, KP_THREADS, m->stream, a);
    } else {
        if (g_tiles) LAUNCH((krylov_phase_kernel<PHASE, false, true>), g_tiles, KP_THREADS, m->stream, a);
        if (g_chunks) LAUNCH((krylov_phase_kernel<PHASE, false, false>), g_chunks, KP_THREADS, m->stream, a);
    }
    if (PHASE != KP_ADD) LAUNCH((krylov_finalize_kernel<PHASE>), unsigned((P.n_comp + 3) / 4), 128, m->stream, a);
}

void krylov_solve_phased(tm_mesh* m, RankMesh& r, const tm_smooth_options* o, tm_smooth_stats* st) {
    phased_plan_build(m, r);
    PhasedPlan& P = *r.pplan;
    cudaStream_t s = m->stream;
    KPArgs a{};
    a.wtiles = P.wtiles.p; a.wt_comp = P.wt_comp.p; a.chunks = P.chunks.p; a.blocks = r.d_blocks.p;
    a.srows = r.d_srows.p; a.jrows = r.d_jrows.p; a.lrows = r.d_lrows.p; a.slaves = r.d_slaves.p; a.rterms = r.d_rhs_terms.p;
    a.comps = P.comps.p; a.state = P.state.p; a.partials = P.partials.p;
    a.xc = r.X[r.cur].p; a.pq = r.pq.p; a.xnew = r.X[1 - r.cur].p;
    a.r = r.kr.p; a.rhat = r.krhat.p; a.s = r.ks.p; a.t = r.kt.p; a.d = r.kd.p;
    a.n_wtiles = P.n_wtiles; a.n_chunks = P.n_chunks; a.n_comp = P.n_comp;
    a.rtol = o->rtol; a.atol = o->atol;
    a.max_iters = o->max_inner_iterations > 0x7fffffffull ? 0x7fffffff : int32_t(o->max_inner_iterations);
    a.max_restarts = 60;
    a.polish = std::max(0, int(o->inner_refinement_cycles));
    double2* const Pb[2] = {r.kp.p, r.kp2.p};
    double2* const Vb[2] = {r.kv.p, r.kv2.p};
    LAUNCH(krylov_reset_kernel, unsigned((P.n_comp + 127) / 128), 128, s, P.state.p, P.n_comp);
    auto poll = [&]() {
        LAUNCH(krylov_count_kernel, 1, 256, s, (const KState*)P.state.p, P.n_comp, P.count.p);
        CUDA_TRY(cudaMemcpyAsync(P.h_count, P.count.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    };
    // two iterations (the ping-pong of p and v returns to its start) as one graph, replayed
    auto two_iterations = [&]() {
        for (int half = 0; half < 2; ++half) {
            KPArgs x = a;
            x.p_old = Pb[half]; x.v_old = Vb[half]; x.p_new = Pb[half ^ 1]; x.v_new = Vb[half ^ 1];
            phased_launch<KP_A>(m, r, x);
            x.p_old = Pb[half ^ 1]; x.v_old = Vb[half ^ 1];   // the current p and v from here on
            phased_launch<KP_B>(m, r, x);
            phased_launch<KP_C>(m, r, x);
        }
    };
    cudaGraphExec_t exec = nullptr;
    uint64_t launches_per_graph = 0;
    if (m->use_graph) {
        const uint64_t before = g_launches.load();
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        try {
            two_iterations();
        } catch (...) {
            cudaStreamEndCapture(s, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        CUDA_TRY(cudaStreamEndCapture(s, &graph));
        launches_per_graph = g_launches.load() - before;
        g_launches.store(before);  // captured, not run
        const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        CUDA_TRY(e);
    }
    try {
        for (int cycle = 0;; ++cycle) {
            KPArgs x = a;
            x.cycle = cycle; x.p_new = Pb[0]; x.v_new = Vb[0];   // every cycle starts on buffer 0 (R0 zeroes it)
            phased_launch<KP_R0>(m, r, x);
            poll();
            if (P.h_count[1] == 0) break;                         // every component is final
            while (P.h_count[0] > 0) {
                for (int k = 0; k < 4; ++k) {
                    if (exec) { CUDA_TRY(cudaGraphLaunch(exec, s)); g_launches.fetch_add(launches_per_graph, std::memory_order_relaxed); }
                    else two_iterations();
                }
                poll();
            }
            phased_launch<KP_ADD>(m, r, a);
        }
    } catch (...) {
        if (exec) cudaGraphExecDestroy(exec);
        throw;
    }
    if (exec) cudaGraphExecDestroy(exec);
    CUDA_TRY(cudaMemcpyAsync(P.h_state.data(), P.state.p, P.h_state.size() * sizeof(KState), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    double weighted_apps = 0.0;
    for (size_t c = 0; c < P.h_state.size(); ++c) {
        const KState& k = P.h_state[c];
        st->inner_iterations += uint64_t(k.iters[0]) + uint64_t(k.iters[1]);
        weighted_apps += double(k.applications) * double(P.h_comps[c].nodes);
        st->last_inner_residual = std::fmax(c == 0 ? 0.0 : st->last_inner_residual, std::fmax(k.norm_r[0], k.norm_r[1]));
        if (k.done[0] != 1 || k.done[1] != 1) st->converged = 0;  // log.warn "did not converge", BiCGStab.zig:368-369
    }
    st->operator_applications += uint64_t(weighted_apps / double(std::max<int64_t>(m->topo.n_nodes, 1)) + 0.5);
}
