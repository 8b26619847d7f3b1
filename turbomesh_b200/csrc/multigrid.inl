// multigrid.inl -- host drivers of the two geometric FAS multigrid hierarchies (TM_SOLVER_FAS_MULTIGRID): the non-nested
// single-block one (config 3) and the nested multi-block / multi-GPU one (config 4) with its Anderson acceleration,
// replicated small levels and transfer tables.  Textually included into the anonymous namespace of turbomesh_gpu.cu
// (after the sweep / exchange helpers it builds on); kernels are in kernels.cuh, the hierarchy planning in mg_plan.hpp.

// ---- geometric FAS multigrid (single block, all boundary nodes fixed) ------------------------------------------
void mg_build(tm_mesh* m, RankMesh& r) {
    if (!r.mg.empty()) return;
    cudaStream_t s = m->stream;
    const auto& B = m->topo.blocks[0];
    int ni = int(B.ni), nj = int(B.nj);
    // Mean physical extents of the block along i and j (from its corner nodes) give the mean cell aspect ratio per level:
    // with point relaxation, a direction whose spacing is much finer than the other's is strongly coupled and is the only
    // one coarsened until the cells are roughly square (semi-coarsening).
    double len_i = 1.0, len_j = 1.0;
    {
        double2 c[4];
        const int64_t idx[4] = {0, int64_t(ni - 1) * nj, int64_t(nj - 1), int64_t(ni - 1) * nj + nj - 1};
        for (int k = 0; k < 4; ++k) CUDA_TRY(cudaMemcpyAsync(&c[k], r.X[r.cur].p + r.L.loff[0] + idx[k], sizeof(double2), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        auto dist = [](double2 a, double2 b) { return std::hypot(a.x - b.x, a.y - b.y); };
        len_i = 0.5 * (dist(c[0], c[1]) + dist(c[2], c[3]));
        len_j = 0.5 * (dist(c[0], c[2]) + dist(c[1], c[3]));
        if (!(len_i > 0.0) || !(len_j > 0.0)) len_i = len_j = 1.0;
    }
    for (int l = 0;; ++l) {
        std::unique_ptr<MgLevel> L(new MgLevel());
        L->ni = ni; L->nj = nj;
        const size_t n = size_t(ni) * size_t(nj);
        std::vector<Tile> tiles;
        const int64_t interior_i = ni - 2;
        const int64_t n_i = std::max<int64_t>(1, (interior_i + m->tile_rows - 1) / m->tile_rows);
        const int64_t rows = (interior_i + n_i - 1) / n_i;
        for (int64_t i0 = 1; i0 <= ni - 2; i0 += rows)
            for (int64_t j0 = 1; j0 <= nj - 2; j0 += TILE_J) tiles.push_back(Tile{0, int32_t(i0), int32_t(j0), int32_t(rows)});
        L->n_tiles = int(tiles.size());
        L->tiles.upload(tiles, s);
        L->blk.upload(std::vector<DevBlock>{DevBlock{0, ni, nj}}, s);
        L->tmp.alloc(n); L->tmp.zero(s);
        if (l == 0) L->E.alloc(n);
        if (l > 0) {
            L->U.alloc(n); L->V.alloc(n); L->rhs.alloc(n); L->E.alloc(n);
            L->U.zero(s); L->V.zero(s); L->rhs.zero(s); L->E.zero(s);
        }
        if (l > 0 && int64_t(n) <= m->mg_small_nodes) {  // all sweeps of a visit in one single-CTA launch
            std::vector<SmallNode> nodes;
            for (int i = 1; i + 1 < ni; ++i)
                for (int j = 1; j + 1 < nj; ++j) nodes.push_back(SmallNode{int64_t(i) * nj + j, 0, i, j, 0});
            L->n_small = int(nodes.size());
            L->small.upload(nodes, s);
        }
        if (l == 1 && m->mg_aa) {
            for (int k = 0; k < m->mg_aa_window; ++k)
                for (auto* ring : {&L->aa_G, &L->aa_F}) {
                    ring->emplace_back(new DevBuf<double2>());
                    ring->back()->alloc(n);
                    ring->back()->zero(s);
                }
            for (DevBuf<double2>* v : {&L->aa_X, &L->aa_D, &L->aa_zero}) { v->alloc(n); v->zero(s); }
            L->aa_part.alloc(size_t(r.vec_grid) * AA_GRAM);
            L->aa_gram.alloc(AA_GRAM); L->aa_gram.zero(s);
            L->aa_coef.alloc(AA_MAX); L->aa_coef.zero(s);
        }
        const double hi = len_i / double(ni - 1), hj = len_j / double(nj - 1);
        bool do_i = ni >= 9, do_j = nj >= 9;
        if (do_i && do_j) {
            if (hi < 0.6 * hj) do_j = false;       // i is the strongly coupled direction
            else if (hj < 0.6 * hi) do_i = false;  // j is
        }
        const int ci = do_i ? (ni + 1) / 2 : ni, cj = do_j ? (nj + 1) / 2 : nj;
        const bool last = (ci == ni && cj == nj);
        if (!last) {
            L->to_coarse = MgLevelDims{ni, nj, ci, cj, double(ni - 1) / double(ci - 1), double(nj - 1) / double(cj - 1)};
            L->scale = (L->to_coarse.r_i * L->to_coarse.r_j) * (L->to_coarse.r_i * L->to_coarse.r_j);
        }
        r.mg.push_back(std::move(L));
        if (last) break;
        ni = ci; nj = cj;
    }
}

// `sweeps` damped-Jacobi sweeps on one level; the fields ping-pong between *pu and *pv
void mg_smooth(tm_mesh* m, RankMesh& r, MgLevel& L, double2** pu, double2** pv, int level, uint64_t sweeps, double omega, bool stats_on_last) {
    const BndArgs none{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0};
    if (level > 0 && L.n_small > 0 && !stats_on_last) {
        LAUNCH(winslow_small_level_kernel, 1, 1024, m->stream, (const SmallNode*)L.small.p, L.n_small, (const DevBlock*)L.blk.p, none, *pu, *pv,
               (const double2*)L.rhs.p, omega, int(sweeps));
        if (sweeps & 1) std::swap(*pu, *pv);
        return;
    }
    for (uint64_t k = 0; k < sweeps; ++k) {
        const bool st = stats_on_last && k + 1 == sweeps;
        const double2* u = *pu;
        double2* out = *pv;
        if (level == 0) {
            if (st) LAUNCH((winslow_interior_bulk_kernel<MODE_RELAX, false, 1, false>), L.n_tiles, TILE_J, m->stream, (const Tile*)L.tiles.p, (const DevBlock*)L.blk.p, u,
                           (const double2*)nullptr, out, omega, (const double2*)nullptr, r.part_int.p, none, (const double2*)nullptr);
            else LAUNCH((winslow_interior_bulk_kernel<MODE_RELAX, false, 0, false>), L.n_tiles, TILE_J, m->stream, (const Tile*)L.tiles.p, (const DevBlock*)L.blk.p, u,
                        (const double2*)nullptr, out, omega, (const double2*)nullptr, r.part_int.p, none, (const double2*)nullptr);
        } else {
            LAUNCH((winslow_interior_bulk_kernel<MODE_RELAX, false, 0, true>), L.n_tiles, TILE_J, m->stream, (const Tile*)L.tiles.p, (const DevBlock*)L.blk.p, u,
                   (const double2*)nullptr, out, omega, (const double2*)nullptr, r.part_int.p, none, (const double2*)L.rhs.p);
        }
        std::swap(*pu, *pv);
    }
}

// tmp = row(u) - rhs on interior nodes (the rim of tmp is never written and stays 0); the restriction flips the sign
void mg_residual(tm_mesh* m, RankMesh& r, MgLevel& L, const double2* u, int level) {
    const BndArgs none{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0};
    if (level == 0)
        LAUNCH((winslow_interior_bulk_kernel<MODE_REL, false, 0, false>), L.n_tiles, TILE_J, m->stream, (const Tile*)L.tiles.p, (const DevBlock*)L.blk.p, u,
               (const double2*)nullptr, L.tmp.p, 1.0, (const double2*)nullptr, r.part_int.p, none, (const double2*)nullptr);
    else
        LAUNCH((winslow_interior_bulk_kernel<MODE_REL, false, 0, true>), L.n_tiles, TILE_J, m->stream, (const Tile*)L.tiles.p, (const DevBlock*)L.blk.p, u,
               (const double2*)nullptr, L.tmp.p, 1.0, (const double2*)nullptr, r.part_int.p, none, (const double2*)L.rhs.p);
}

// V(nu,nu) cycles of the full approximation scheme until the fine-level Jacobi update drops below stop_max_update
// Anderson acceleration of the single-block cycle on the nodes of level 1 (same scheme as anderson_step for multi-block
// meshes; the levels are not nested here, so the samples and the extrapolation travel by interpolation)
void anderson_step_single(tm_mesh* m, RankMesh& r, MgLevel& F, MgLevel& C, double2* u_f) {
    cudaStream_t s = m->stream;
    const int64_t n = int64_t(C.ni) * C.nj;
    const bool have_x = C.aa_have_x;
    if (have_x) { C.aa_head = (C.aa_head + 1) % m->mg_aa_window; C.aa_count = std::min(C.aa_count + 1, m->mg_aa_window); }
    double2* g_new = have_x ? C.aa_G[size_t(C.aa_head)]->p : C.aa_X.p;
    double2* f_new = have_x ? C.aa_F[size_t(C.aa_head)]->p : C.aa_D.p;
    dim3 gc((C.nj + 127) / 128, C.ni);
    LAUNCH(mg_sample_kernel, gc, 128, s, F.to_coarse, (const double2*)u_f, (const double2*)C.aa_X.p, g_new, f_new);
    C.aa_have_x = true;
    if (!have_x) return;
    AaFields h{};
    h.q = C.aa_count;
    for (int i = 0; i < h.q; ++i) {
        const size_t slot = size_t((C.aa_head + m->mg_aa_window - (h.q - 1) + i) % m->mg_aa_window);
        h.G[i] = C.aa_G[slot]->p; h.F[i] = C.aa_F[slot]->p;
    }
    LAUNCH(aa_gram_kernel, r.vec_grid, 256, s, n, h, C.aa_part.p);
    LAUNCH(aa_reduce_kernel, 1, 32 * AA_GRAM, s, (const double*)C.aa_part.p, r.vec_grid, C.aa_gram.p);
    LAUNCH(aa_solve_kernel, 1, 32, s, (const double*)C.aa_gram.p, h.q, C.aa_coef.p);
    LAUNCH(aa_combine_kernel, r.vec_grid, 256, s, n, h, (const double*)C.aa_coef.p, C.aa_D.p, C.aa_X.p);
    if (h.q < 2) return;
    dim3 gf((F.nj + 127) / 128, F.ni);
    LAUNCH(mg_prolong_kernel, gf, 128, s, F.to_coarse, (const double2*)C.aa_D.p, (const double2*)C.aa_zero.p, u_f);
}

void run_fas_multigrid_blocks(tm_mesh* m, const tm_smooth_options* o, tm_smooth_stats* st);
bool mg_single_block_case(const tm_mesh* m) {
    return m->n_ranks == 1 && m->topo.blocks.size() == 1 && m->topo.smoothed.empty() && m->topo.junction_rows.empty() && m->topo.sliding.empty() &&
           m->topo.slaves.empty() && m->cf == TM_CF_LAPLACE && !std::getenv("TM_MG_NESTED");
}
void run_fas_multigrid(tm_mesh* m, const tm_smooth_options* o, tm_smooth_stats* st) {
    if (!mg_single_block_case(m)) { run_fas_multigrid_blocks(m, o, st); return; }
    RankMesh& r = *m->ranks[0];
    mg_build(m, r);
    cudaStream_t s = m->stream;
    const int n_levels = int(r.mg.size());
    const uint64_t nu = o->sweeps_per_iteration;
    const BndArgs none{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0};
    std::vector<double2*> U, V;
    U.resize(size_t(n_levels));
    V.resize(size_t(n_levels));
    double fine_work = 0.0;  // operator applications in units of one fine-grid application
    for (uint64_t cyc = 0; cyc < o->iterations; ++cyc) {
        U[0] = r.X[r.cur].p; V[0] = r.X[1 - r.cur].p;
        for (int l = 1; l < n_levels; ++l) { U[size_t(l)] = r.mg[size_t(l)]->U.p; V[size_t(l)] = r.mg[size_t(l)]->V.p; }
        const bool want_change = o->stop_max_update > 0.0 || cyc + 1 == o->iterations;  // mesh change over the whole cycle
        if (want_change) CUDA_TRY(cudaMemcpyAsync(r.mg[0]->E.p, U[0], size_t(r.N) * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        for (int l = 0; l + 1 < n_levels; ++l) {  // down: smooth, restrict iterate and residual, build the coarse tau term
            MgLevel& F = *r.mg[size_t(l)];
            MgLevel& C = *r.mg[size_t(l) + 1];
            const double w = double(F.ni) * F.nj / (double(r.mg[0]->ni) * r.mg[0]->nj);
            mg_smooth(m, r, F, &U[size_t(l)], &V[size_t(l)], l, nu, o->omega, false);
            if (l == 0 && m->mg_aa && !C.aa_G.empty()) anderson_step_single(m, r, F, C, U[0]);
            mg_residual(m, r, F, U[size_t(l)], l);
            fine_work += w * double(nu + 1);
            dim3 gc((C.nj + 127) / 128, C.ni);
            LAUNCH(mg_restrict_kernel, gc, 128, s, F.to_coarse, (const double2*)U[size_t(l)], (const double2*)F.tmp.p, U[size_t(l) + 1], C.E.p, C.rhs.p, -F.scale);
            CUDA_TRY(cudaMemcpyAsync(V[size_t(l) + 1], U[size_t(l) + 1], size_t(C.ni) * C.nj * sizeof(double2), cudaMemcpyDeviceToDevice, s));
            LAUNCH((winslow_interior_bulk_kernel<MODE_REL, false, 0, false>), C.n_tiles, TILE_J, s, (const Tile*)C.tiles.p, (const DevBlock*)C.blk.p,
                   (const double2*)U[size_t(l) + 1], (const double2*)nullptr, C.tmp.p, 1.0, (const double2*)nullptr, r.part_int.p, none, (const double2*)nullptr);
            LAUNCH(mg_coarse_rhs_kernel, gc, 128, s, C.ni, C.nj, (const double2*)C.tmp.p, C.rhs.p);
        }
        {   // coarsest level: relax (almost) to convergence
            const int l = n_levels - 1;
            MgLevel& C = *r.mg[size_t(l)];
            const uint64_t nc = n_levels > 1 ? 4 * uint64_t(std::max(C.ni, C.nj)) : nu;
            mg_smooth(m, r, C, &U[size_t(l)], &V[size_t(l)], l, nc, o->omega, false);
            fine_work += double(nc) * double(C.ni) * C.nj / (double(r.mg[0]->ni) * r.mg[0]->nj);
        }
        for (int l = n_levels - 2; l >= 0; --l) {  // up: interpolate the coarse correction, smooth
            MgLevel& F = *r.mg[size_t(l)];
            MgLevel& C = *r.mg[size_t(l) + 1];
            dim3 gf((F.nj + 127) / 128, F.ni);
            LAUNCH(mg_prolong_kernel, gf, 128, s, F.to_coarse, (const double2*)U[size_t(l) + 1], (const double2*)C.E.p, U[size_t(l)]);
            mg_smooth(m, r, F, &U[size_t(l)], &V[size_t(l)], l, nu, o->omega, false);
            fine_work += double(nu) * double(F.ni) * F.nj / (double(r.mg[0]->ni) * r.mg[0]->nj);
        }
        if (U[0] != r.X[r.cur].p) r.cur = 1 - r.cur;  // the finest iterate lives in the mesh's own ping-pong pair
        for (int l = 1; l < n_levels; ++l)
            if (U[size_t(l)] != r.mg[size_t(l)]->U.p) std::swap(r.mg[size_t(l)]->U.p, r.mg[size_t(l)]->V.p);
        if (want_change) {
            LAUNCH(diff_stats_kernel, r.vec_grid, VEC_THREADS, s, r.L.n_own, (const double2*)r.mg[0]->E.p, (const double2*)U[0], r.part_vec.p);
            launch_reduce(m, RED_UPDATE_STATS, o, false);
        }
        m->outer_done += 1;
        st->outer_iterations += 1;
        st->inner_iterations += 1;
        if (o->stop_max_update > 0.0) {
            fetch_ctl(m);
            if (m->h_ctl->max_update <= o->stop_max_update) break;
        }
    }
    st->operator_applications += uint64_t(fine_work + 0.5);
}

// =====================================================================================================
// Multi-block / multi-GPU geometric FAS multigrid (config 4: time to converged mesh on the cascade).
//
// Coarse levels are complete multi-block meshes of their own: the block sizes and all connection / condition ranges are
// halved (NESTED coarsening: every coarse node is a fine node), the topology analysis and the rank-local tables are
// rebuilt per level with the very same code as the fine level (Topology::build + localize), and the smoother of a level
// is the same sweep launch (interior tiles + interface / junction / sliding rows) with the FAS tau term as right-hand
// side.  Directions are tied into classes by the connections (the along and the normal direction of the two sides of a
// connection must coarsen together); a class coarsens when every extent and every range end point in it is even, and
// only while its mean cell size is not much larger than the smallest one (semi-coarsening).
// =====================================================================================================

// mean cell size per (block, direction), all blocks of the mesh (all-reduced over the ranks)
std::vector<double> block_cell_sizes(tm_mesh* m) {
    cudaStream_t s = m->stream;
    const size_t nb = m->h_blocks.size();
    DevBuf<double> d_len;
    d_len.alloc(4 * nb);
    d_len.zero(s);
    for (auto& rp : m->ranks) {
        RankMesh& r = *rp;
        std::vector<SideLenJob> jobs;
        for (int32_t b : r.L.own_blocks) jobs.push_back(SideLenJob{r.L.loff[size_t(b)], int32_t(m->topo.blocks[size_t(b)].ni), int32_t(m->topo.blocks[size_t(b)].nj), b});
        if (jobs.empty()) continue;
        DevBuf<SideLenJob> d_jobs;
        d_jobs.upload(jobs, s);
        LAUNCH(side_length_kernel, unsigned(4 * jobs.size()), 256, s, (const SideLenJob*)d_jobs.p, (const double2*)r.X[r.cur].p, d_len.p);
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    if (m->n_ranks > 1 && !m->emulated) NCCL_TRY(g_nccl.AllReduce(d_len.p, d_len.p, 4 * nb, ncclDouble, ncclSum, m->comm, s));
    std::vector<double> len(4 * nb, 0.0), h(2 * nb, 0.0);
    CUDA_TRY(cudaMemcpyAsync(len.data(), d_len.p, 4 * nb * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (size_t b = 0; b < nb; ++b) {
        h[2 * b] = 0.5 * (len[4 * b] + len[4 * b + 1]) / double(m->h_blocks[b].ni - 1);      // sides i_min / i_max run along i
        h[2 * b + 1] = 0.5 * (len[4 * b + 2] + len[4 * b + 3]) / double(m->h_blocks[b].nj - 1);
    }
    return h;
}

// bit 0..3: every node strictly inside the side i = 0 / i = ni-1 / j = 0 / j = nj-1 of the block is a sliding node
int32_t side_slide_mask(const Topology& T, size_t k) {
    const int64_t ni = T.blocks[k].ni, nj = T.blocks[k].nj;
    auto slides = [&](int64_t base, int64_t stride, int64_t n) {
        for (int64_t q = 1; q + 1 < n; ++q)
            if (T.kind[size_t(T.bid(k, base + q * stride))] != K_SLIDING) return false;
        return n > 2;
    };
    return (slides(0, 1, nj) ? 1 : 0) | (slides((ni - 1) * nj, 1, nj) ? 2 : 0) | (slides(0, nj, ni) ? 4 : 0) | (slides(nj - 1, nj, ni) ? 8 : 0);
}

// residual restriction of the rows that straddle blocks: coarse row <- full weighting of the fine residuals around it
void mgb_build_transfer(tm_mesh* m, const Topology& TF, const MgbLevel& F, RankMesh& rf, const Topology& TC, const MgbLevel& C, RankMesh& rc) {
    const int rank = rf.L.rank;
    // who computes what: the owner on the fine level (a replicated fine level: this rank, for every block)
    const std::vector<int32_t> all_mine(m->owner.size(), int32_t(rank));
    const std::vector<int32_t>& owner = rf.replicated ? all_mine : m->owner;
    const std::vector<int32_t>& owner_c = rc.replicated ? all_mine : m->owner;
    rf.xfer_blocks.clear();
    for (int32_t b : rf.L.own_blocks) {
        const size_t k = size_t(b);
        if (TF.blocks[k].ni > 65535 || TF.blocks[k].nj > 0x7fffffff) TM_THROW(TM_ERR_UNSUPPORTED, "multigrid: block %d too large for the transfer kernels", b);
        const int64_t ni = TF.blocks[k].ni, nj = TF.blocks[k].nj;
        const int32_t slide = side_slide_mask(TF, k);
        rf.xfer_blocks.push_back(BlockXfer{rf.L.loff[k], rc.L.loff[k], int32_t(ni), int32_t(nj), int32_t(TC.blocks[k].ni), int32_t(TC.blocks[k].nj), F.fi[k],
                                           F.fj[k], slide, 0});
    }
    std::vector<RestrictRow> rows;
    auto lf = [&](int64_t g) { return local_index(TF, owner, rf.L, g); };
    auto lc = [&](int64_t g) { return local_index(TC, owner_c, rc.L, g); };
    auto add = [](RestrictRow& row, int64_t src, double w) { row.src[row.n] = src; row.w[row.n] = w; ++row.n; };
    for (size_t c = 0; c < C.conns.size(); ++c) {
        const tm_connection& cc = C.conns[c];
        const tm_connection& cf = F.conns[c];
        const size_t b0 = size_t(cc.ranges[0].block), b1 = size_t(cc.ranges[1].block);
        if (owner[b0] != rank) continue;
        int64_t B0, A0, N0, f0, a0, n0, f1, a1, n1;
        TC.walk(cc.ranges[0], B0, A0, N0);
        TF.walk(cf.ranges[0], f0, a0, n0);
        TF.walk(cf.ranges[1], f1, a1, n1);
        const int al = along_dir(cc.ranges[0].side);
        const int fa = al == 0 ? F.fi[b0] : F.fj[b0], fn = al == 0 ? F.fj[b0] : F.fi[b0];
        const double scale = double(fa * fn) * double(fa * fn);
        const int64_t nC = Topology::range_len(cc.ranges[0]);
        for (int64_t K = 1; K + 1 < nC; ++K) {
            const int64_t g0c = TC.blocks[b0].off + B0 + K * A0;
            if (TC.kind[size_t(TC.bid_of_global(g0c))] != K_SMOOTHED) continue;
            const int64_t k = K * fa;
            const int64_t g0f = TF.blocks[b0].off + f0 + k * a0, g1f = TF.blocks[b1].off + f1 + k * a1;
            RestrictRow row{};
            row.dst = lc(g0c);
            const int pa = fa == 2 ? 1 : 0;
            // an end point of the connection that slides carries no Winslow row: its share of the neighbouring residual
            // stays with this row (see mgb_restrict_kernel)
            auto end_slides = [&](int64_t Kend) { return TC.kind[size_t(TC.bid_of_global(TC.blocks[b0].off + B0 + Kend * A0))] == K_SLIDING; };
            const double w_lo = (K == 1 && end_slides(0)) ? 0.5 : 0.25, w_hi = (K == nC - 2 && end_slides(nC - 1)) ? 0.5 : 0.25;
            for (int da = -pa; da <= pa; ++da) {
                const double wa = pa ? (da == 0 ? 0.5 : (da < 0 ? w_lo : w_hi)) : 1.0;
                if (fn == 2) {
                    add(row, lf(g0f + da * a0), scale * wa * 0.5);
                    add(row, lf(g0f + da * a0 + n0), scale * wa * 0.25);
                    add(row, lf(g1f + da * a1 + n1), scale * wa * 0.25);
                } else {
                    add(row, lf(g0f + da * a0), scale * wa);
                }
            }
            rows.push_back(row);
        }
    }
    auto fine_of = [&](int64_t gc, size_t& b) {  // the fine node a coarse node coincides with
        b = TC.block_of(gc);
        const int64_t local = gc - TC.blocks[b].off, I = local / TC.blocks[b].nj, J = local - I * TC.blocks[b].nj;
        return TF.blocks[b].off + (I * F.fi[b]) * TF.blocks[b].nj + J * F.fj[b];
    };
    for (const auto& jr : TC.junction_rows) {  // update units (lengths): second differences over the diagonal neighbours
        size_t b;
        const int64_t gf = fine_of(jr.self, b);
        if (owner[b] != rank) continue;
        RestrictRow row{};
        row.dst = lc(jr.self);
        add(row, lf(gf), double(F.fi[b] * F.fj[b]));
        rows.push_back(row);
    }
    for (const auto& sr : TC.sliding) {       // update units: a first difference along the inward normal
        size_t b;
        const int64_t gf = fine_of(sr.self, b);
        if (owner[b] != rank) continue;
        const int64_t inward = sr.inner - sr.self;
        const int fn = (inward == 1 || inward == -1) ? F.fj[b] : F.fi[b];
        RestrictRow row{};
        row.dst = lc(sr.self);
        add(row, lf(gf), double(fn));
        rows.push_back(row);
    }
    rf.n_rrows = int(rows.size());
    rf.d_rrows.upload(rows, m->stream);
    rf.d_xfer_blocks.upload(rf.xfer_blocks, m->stream);
    rf.xf_ni_f = rf.xf_nj_f = rf.xf_ni_c = rf.xf_nj_c = 0;
    rf.xf_all_2x2 = !rf.xfer_blocks.empty();
    for (const BlockXfer& b : rf.xfer_blocks) {
        rf.xf_all_2x2 = rf.xf_all_2x2 && b.fi == 2 && b.fj == 2;
        rf.xf_ni_f = std::max(rf.xf_ni_f, b.ni_f); rf.xf_nj_f = std::max(rf.xf_nj_f, b.nj_f);
        rf.xf_ni_c = std::max(rf.xf_ni_c, b.ni_c); rf.xf_nj_c = std::max(rf.xf_nj_c, b.nj_c);
    }
    if (rf.xfer_blocks.size() > 65535) TM_THROW(TM_ERR_UNSUPPORTED, "multigrid: more than 65535 blocks per rank");
}

void mgb_build(tm_mesh* m) {
    if (!m->mgb.empty()) return;
    cudaStream_t s = m->stream;
    const size_t nb = m->h_blocks.size();
    // the hierarchy is planned on the host (mg_plan.hpp); here its levels get their device meshes and transfer tables
    int max_levels = 20;
    if (const char* e = std::getenv("TM_MG_MAX_LEVELS")) max_levels = std::max(1, std::atoi(e));   // tuning / diagnostics
    std::vector<MgPlanLevel> plan = plan_multigrid(m->h_blocks, m->h_conns, m->h_bcs, block_cell_sizes(m), max_levels);
    {
        std::unique_ptr<MgbLevel> L0(new MgbLevel());
        L0->blocks = plan[0].blocks; L0->conns = plan[0].conns; L0->bcs = plan[0].bcs;
        L0->tan_i.assign(nb, 1.0); L0->tan_j.assign(nb, 1.0);
        m->mgb.push_back(std::move(L0));
    }
    for (auto& rp : m->ranks) {
        rp->mg_tmp.alloc(size_t(std::max<int64_t>(rp->N, 1))); rp->mg_tmp.zero(s);
        rp->mg_E.alloc(size_t(std::max<int64_t>(rp->N, 1))); rp->mg_E.zero(s);
        rp->d_change.alloc(1); rp->d_change.zero(s);
        CUDA_TRY(cudaStreamSynchronize(s));
        p2p_add_tmp(m, *rp);
    }
    const double n0 = double(m->topo.n_nodes);
    for (size_t level = 1; level < plan.size(); ++level) {
        MgbLevel& F = *m->mgb.back();
        F.fi = plan[level - 1].fi; F.fj = plan[level - 1].fj;
        std::unique_ptr<MgbLevel> C(new MgbLevel());
        C->blocks = std::move(plan[level].blocks);
        C->conns = std::move(plan[level].conns);
        C->bcs = std::move(plan[level].bcs);
        C->topo = std::move(plan[level].topo);
        C->work = double(C->topo.n_nodes) / n0;
        // Rows next to a sliding side: the Galerkin coarse operator (restriction weights 1/2, 1/2, 1/4 over the first three
        // rows, boundary unknown eliminated) carries 5/4 of the tangential term of the level below; f' = f/2 + 3/4.
        C->tan_i.resize(nb); C->tan_j.resize(nb);
        for (size_t b = 0; b < nb; ++b) {
            C->tan_i[b] = F.fi[b] == 2 ? 0.5 * F.tan_i[b] + 0.75 : F.tan_i[b];
            C->tan_j[b] = F.fj[b] == 2 ? 0.5 * F.tan_j[b] + 0.75 : F.tan_j[b];
        }
        // Small levels are REPLICATED on every rank of a multi-GPU run (every rank holds all blocks and does the same work):
        // their sweeps then need no halo exchange at all, which is what such levels cost.  Levels of a few thousand nodes
        // additionally run all sweeps of a visit in one single-CTA launch (winslow_small_level_kernel).
        const bool replicate = m->n_ranks > 1 && C->topo.n_nodes <= m->mg_replicate_nodes;
        for (auto& rp : m->ranks) {
            C->ranks.emplace_back(new RankMesh());
            RankMesh& rc = *C->ranks.back();
            rc.replicated = replicate;
            const std::vector<int32_t> all_mine(nb, int32_t(rp->L.rank));
            build_rank(m, C->topo, rc, rp->L.rank, replicate ? &all_mine : nullptr);
            if (C->topo.n_nodes <= m->mg_small_nodes && (replicate || m->n_ranks == 1)) {
                std::vector<SmallNode> nodes;
                for (int32_t b : rc.L.own_blocks) {
                    const auto& B = C->topo.blocks[size_t(b)];
                    for (int64_t i = 1; i + 1 < B.ni; ++i)
                        for (int64_t j = 1; j + 1 < B.nj; ++j) nodes.push_back(SmallNode{rc.L.loff[size_t(b)] + i * B.nj + j, b, int32_t(i), int32_t(j), 0});
                }
                rc.n_small = int(nodes.size());
                rc.d_small.upload(nodes, s);
            }
            for (DevBuf<double2>* v : {&rc.mg_rhs, &rc.mg_tmp, &rc.mg_E, &rc.mg_zero}) { v->alloc(size_t(std::max<int64_t>(rc.N, 1))); v->zero(s); }
            std::fill(rc.have_coords.begin(), rc.have_coords.end(), uint8_t(1));
            std::vector<DevBlock> blocks(nb, DevBlock{0, 0, 0});
            for (int32_t b : rc.L.own_blocks) {
                const size_t k = size_t(b);
                blocks[k] = DevBlock{rc.L.loff[k], int32_t(C->topo.blocks[k].ni), int32_t(C->topo.blocks[k].nj), side_slide_mask(C->topo, k), 0, C->tan_i[k], C->tan_j[k]};
            }
            rc.d_blocks.upload(blocks, s);
            if (m->mgb.size() == 1 && m->mg_aa) {  // this is level 1
                for (int k = 0; k < m->mg_aa_window; ++k)
                    for (auto* ring : {&rc.aa_G, &rc.aa_F}) {
                        ring->emplace_back(new DevBuf<double2>());
                        ring->back()->alloc(size_t(std::max<int64_t>(rc.N, 1)));
                        ring->back()->zero(s);
                    }
                rc.aa_X.alloc(size_t(std::max<int64_t>(rc.N, 1))); rc.aa_X.zero(s);
                rc.aa_D.alloc(size_t(std::max<int64_t>(rc.N, 1))); rc.aa_D.zero(s);
                rc.aa_part.alloc(size_t(rc.vec_grid) * AA_GRAM);
                rc.aa_gram.alloc(AA_GRAM); rc.aa_gram.zero(s);
                rc.aa_coef.alloc(AA_MAX); rc.aa_coef.zero(s);
            }
            CUDA_TRY(cudaStreamSynchronize(s));
            if (!replicate) {
                p2p_setup(m, rc);
                p2p_add_tmp(m, rc);
            }
        }
        const Topology& TF = m->mgb.size() == 1 ? m->topo : F.topo;
        RankList& RF = m->mgb.size() == 1 ? m->ranks : F.ranks;
        for (size_t q = 0; q < RF.size(); ++q) mgb_build_transfer(m, TF, F, *RF[q], C->topo, *C, *C->ranks[q]);
        m->mgb.push_back(std::move(C));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
}

// fine += interpolated (u_c - e_c) on the interior nodes of every own block
void launch_prolong_blocks(tm_mesh* m, RankMesh& rf, const double2* u_c, const double2* e_c, double2* u_f) {
    if (rf.xfer_blocks.empty()) return;
    if (rf.xf_all_2x2) {
        dim3 g((rf.xf_nj_c - 1 + 127) / 128, rf.xf_ni_c - 1, unsigned(rf.xfer_blocks.size()));
        LAUNCH(mgb_prolong_2x2_kernel, g, 128, m->stream, (const BlockXfer*)rf.d_xfer_blocks.p, u_c, e_c, u_f);
    } else {
        dim3 g((rf.xf_nj_f + 127) / 128, (rf.xf_ni_f + MGB_ROWS - 1) / MGB_ROWS, unsigned(rf.xfer_blocks.size()));
        LAUNCH(mgb_prolong_kernel, g, 128, m->stream, (const BlockXfer*)rf.d_xfer_blocks.p, u_c, e_c, u_f);
    }
}

// Anderson acceleration at the restriction point of a cycle (see kernels.cuh): sample, least squares over the last <= 3
// iterations, interpolate the extrapolation to the fine mesh.
void anderson_step(tm_mesh* m, RankList& RF, RankList& RC) {
    cudaStream_t s = m->stream;
    auto xcur = [](RankMesh& r) { return r.X[r.cur].p; };
    const bool have_x = RC[0]->aa_have_x;
    for (size_t q = 0; q < RF.size(); ++q) {
        RankMesh& rf = *RF[q];
        RankMesh& rc = *RC[q];
        if (have_x) { rc.aa_head = (rc.aa_head + 1) % m->mg_aa_window; rc.aa_count = std::min(rc.aa_count + 1, m->mg_aa_window); }
        // without a previous state there is no residual yet: the sample only becomes the state X
        double2* g_new = have_x ? rc.aa_G[size_t(rc.aa_head)]->p : rc.aa_X.p;
        double2* f_new = have_x ? rc.aa_F[size_t(rc.aa_head)]->p : rc.aa_D.p;
        if (!rf.xfer_blocks.empty()) {
            dim3 g((rf.xf_nj_c + 127) / 128, rf.xf_ni_c, unsigned(rf.xfer_blocks.size()));
            LAUNCH(aa_sample_kernel, g, 128, s, (const BlockXfer*)rf.d_xfer_blocks.p, (const double2*)xcur(rf), (const double2*)rc.aa_X.p, g_new, f_new);
        }
        rc.aa_have_x = true;
    }
    if (!have_x) return;
    const int q_res = RC[0]->aa_count;
    auto fields = [&](RankMesh& rc) {
        AaFields h{};
        h.q = q_res;
        for (int i = 0; i < q_res; ++i) {
            const size_t slot = size_t((rc.aa_head + m->mg_aa_window - (q_res - 1) + i) % m->mg_aa_window);
            h.G[i] = rc.aa_G[slot]->p; h.F[i] = rc.aa_F[slot]->p;
        }
        return h;
    };
    for (auto& rp : RC) {
        RankMesh& rc = *rp;
        LAUNCH(aa_gram_kernel, rc.vec_grid, 256, s, rc.L.n_own, fields(rc), rc.aa_part.p);
        LAUNCH(aa_reduce_kernel, 1, 32 * AA_GRAM, s, (const double*)rc.aa_part.p, rc.vec_grid, rc.aa_gram.p);
    }
    if (m->n_ranks > 1) {
        if (m->emulated) {
            SumPtrs ptrs{};
            for (size_t k = 0; k < RC.size(); ++k) ptrs.p[k] = RC[k]->aa_gram.p;
            LAUNCH(combine_sum_kernel, 1, 32, s, ptrs, int(RC.size()), AA_GRAM);
        } else {
            NCCL_TRY(g_nccl.AllReduce(RC[0]->aa_gram.p, RC[0]->aa_gram.p, AA_GRAM, ncclDouble, ncclSum, m->comm, s));
        }
    }
    for (size_t q = 0; q < RF.size(); ++q) {
        RankMesh& rf = *RF[q];
        RankMesh& rc = *RC[q];
        LAUNCH(aa_solve_kernel, 1, 32, s, (const double*)rc.aa_gram.p, q_res, rc.aa_coef.p);
        LAUNCH(aa_combine_kernel, rc.vec_grid, 256, s, rc.L.n_own, fields(rc), (const double*)rc.aa_coef.p, rc.aa_D.p, rc.aa_X.p);
        if (q_res < 2) continue;  // nothing to extrapolate from yet (d = 0)
        launch_prolong_blocks(m, rf, (const double2*)rc.aa_D.p, (const double2*)rc.mg_zero.p, xcur(rf));
        if (rf.n_bnd_rows > 0)
            LAUNCH(mgb_prolong_rows_kernel, (rf.n_bnd_rows + 127) / 128, 128, s, (const BlockXfer*)rf.d_xfer_blocks.p, int(rf.xfer_blocks.size()),
                   (const SmoothedRow*)rf.d_srows.p, int(rf.L.smoothed.size()), (const JunctionRow*)rf.d_jrows.p, int(rf.L.junction_rows.size()),
                   (const SlidingRow*)rf.d_lrows.p, int(rf.L.sliding.size()), (const double2*)rc.aa_D.p, (const double2*)rc.mg_zero.p, xcur(rf));
    }
    if (q_res >= 2) {
        exchange_on(m, RF, xcur);
        for (auto& rp : RF) sync_slaves(m, *rp, xcur(*rp), 1);
    }
}

void run_fas_multigrid_blocks(tm_mesh* m, const tm_smooth_options* o, tm_smooth_stats* st) {
    if (m->cf != TM_CF_LAPLACE) TM_THROW(TM_ERR_UNSUPPORTED, "the multigrid solver supports the Laplace control function only");
    mgb_build(m);
    cudaStream_t s = m->stream;
    const int nl = int(m->mgb.size());
    const uint64_t nu = o->sweeps_per_iteration;
    auto ranks_of = [&](int l) -> RankList& { return l == 0 ? m->ranks : m->mgb[size_t(l)]->ranks; };
    auto xcur = [](RankMesh& r) { return r.X[r.cur].p; };
    auto tmp_of = [](RankMesh& r) { return r.mg_tmp.p; };
    double fine_work = 0.0;
    auto smooth = [&](int l, uint64_t sweeps) {
        RankList& R = ranks_of(l);
        if (l > 0 && !R.empty() && R[0]->n_small > 0) {  // a tiny level: all sweeps in one single-CTA launch per rank held here
            for (auto& rp : R) {
                RankMesh& r = *rp;
                const BndArgs bnd{r.d_srows.p, r.d_jrows.p, r.d_lrows.p, r.d_slaves.p, r.part_bnd.p, int(r.L.smoothed.size()),
                                  int(r.L.junction_rows.size()), int(r.L.sliding.size()), r.n_bnd_ctas};
                LAUNCH(winslow_small_level_kernel, 1, 1024, s, (const SmallNode*)r.d_small.p, r.n_small, (const DevBlock*)r.d_blocks.p, bnd, r.X[r.cur].p, r.X[1 - r.cur].p,
                       (const double2*)r.mg_rhs.p, o->omega, int(sweeps));
                if (sweeps & 1) r.cur = 1 - r.cur;
            }
            fine_work += double(sweeps) * m->mgb[size_t(l)]->work;
            return;
        }
        for (uint64_t k = 0; k < sweeps; ++k) relax_sweep<0>(m, R, o->omega, true, l > 0);
        relax_join(m);
        fine_work += double(sweeps) * m->mgb[size_t(l)]->work;
    };
    auto refresh = [&](int l) {  // ghosts, then every copy re-derived from its root
        RankList& R = ranks_of(l);
        exchange_on(m, R, xcur);
        for (auto& rp : R) sync_slaves(m, *rp, xcur(*rp), 1);
    };
    uint64_t n_coarsest = nu;
    if (nl > 1) {
        const MgbLevel& C = *m->mgb.back();
        uint64_t ext = 0;
        for (const auto& b : C.blocks) ext = std::max<uint64_t>(ext, std::max(b.ni, b.nj));
        n_coarsest = std::min<uint64_t>(400, 4 * ext * uint64_t(std::ceil(std::sqrt(double(C.blocks.size())))));
        if (const char* e = std::getenv("TM_MG_COARSEST_SWEEPS")) n_coarsest = uint64_t(std::max(1, std::atoi(e)));
    }
    for (uint64_t cyc = 0; cyc < o->iterations; ++cyc) {
        // Convergence measure of a cycle.  With coarse levels: how far the level-1 nodes moved between this cycle's and the
        // previous cycle's restriction (taken inside the restriction kernel, no extra pass over the mesh; the very first
        // cycle after begin_smoothing compares against the initial mesh).  Without: the change of the whole mesh.
        const bool want_change = o->stop_max_update > 0.0 || cyc + 1 == o->iterations;
        const bool track = nl > 1;
        if (want_change && !track)
            for (auto& rp : m->ranks) CUDA_TRY(cudaMemcpyAsync(rp->mg_E.p, xcur(*rp), size_t(rp->N) * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        if (track)
            for (auto& rp : m->ranks) CUDA_TRY(cudaMemsetAsync(rp->d_change.p, 0, sizeof(unsigned long long), s));
        for (int l = 0; l + 1 < nl; ++l) {
            RankList& RF = ranks_of(l);
            RankList& RC = ranks_of(l + 1);
            smooth(l, nu);
            if (l == 0 && m->mg_aa && !RC.empty() && !RC[0]->aa_G.empty()) anderson_step(m, RF, RC);
            for (auto& rp : RF) launch_rows_mg<MODE_REL, 0>(m, *rp, xcur(*rp), rp->mg_tmp.p, 1.0, l > 0 ? (const double2*)rp->mg_rhs.p : nullptr);
            exchange_on(m, RF, tmp_of);
            fine_work += m->mgb[size_t(l)]->work;
            const bool gather = !RC.empty() && RC[0]->replicated && !RF[0]->replicated;  // every rank restricts its blocks, all need all
            if (gather)
                for (auto& rp : RC) {
                    CUDA_TRY(cudaMemsetAsync(xcur(*rp), 0, size_t(rp->L.n_own) * sizeof(double2), s));
                    CUDA_TRY(cudaMemsetAsync(rp->mg_tmp.p, 0, size_t(rp->L.n_own) * sizeof(double2), s));
                }
            for (size_t q = 0; q < RF.size(); ++q) {
                RankMesh& rf = *RF[q];
                RankMesh& rc = *RC[q];
                const double2* e_prev = rc.E();
                if (!rf.xfer_blocks.empty()) {  // all own blocks in one launch (blockIdx.z = block)
                    dim3 g((rf.xf_nj_c + 127) / 128, (rf.xf_ni_c + MGB_ROWS - 1) / MGB_ROWS, unsigned(rf.xfer_blocks.size()));
                    LAUNCH(mgb_restrict_kernel, g, 128, s, (const BlockXfer*)rf.d_xfer_blocks.p, (const double2*)xcur(rf), (const double2*)rf.mg_tmp.p, xcur(rc), rc.E(),
                           rc.mg_tmp.p, l == 0 ? rf.d_change.p : (unsigned long long*)nullptr, e_prev);
                }
                if (rf.n_rrows > 0)
                    LAUNCH(mgb_restrict_rows_kernel, (rf.n_rrows + 127) / 128, 128, s, (const RestrictRow*)rf.d_rrows.p, rf.n_rrows, (const double2*)rf.mg_tmp.p, rc.mg_tmp.p);
            }
            if (gather) {  // sum of "mine, zero elsewhere" = everybody's blocks (exact: one non-zero contribution per node)
                for (double2* (*field)(RankMesh&) : {+[](RankMesh& r) { return r.X[r.cur].p; }, +[](RankMesh& r) { return r.mg_tmp.p; }}) {
                    if (m->emulated) {
                        SumPtrs ptrs{};
                        for (size_t k = 0; k < RC.size(); ++k) ptrs.p[k] = reinterpret_cast<double*>(field(*RC[k]));
                        LAUNCH(combine_sum_fields_kernel, 64, 256, s, ptrs, int(RC.size()), 2 * RC[0]->L.n_own);
                    } else {
                        double* f = reinterpret_cast<double*>(field(*RC[0]));
                        NCCL_TRY(g_nccl.AllReduce(f, f, size_t(2 * RC[0]->L.n_own), ncclDouble, ncclSum, m->comm, s));
                    }
                }
            }
            refresh(l + 1);
            for (auto& rp : RC) {
                RankMesh& rc = *rp;
                const int n_l = int(rc.L.sliding.size());
                if (n_l > 0) LAUNCH(capture_boundary_kernel, (n_l + 127) / 128, 128, s, rc.d_lrows.p, n_l, (const FixedOverride*)nullptr, 0, xcur(rc));
                if (!rc.mg_primed) {  // nodes no row writes (fixed ones) never change after the first restriction
                    CUDA_TRY(cudaMemcpyAsync(rc.X[1 - rc.cur].p, xcur(rc), size_t(rc.N) * sizeof(double2), cudaMemcpyDeviceToDevice, s));
                    rc.mg_primed = true;
                }
                // tau_c = row_c(I u_f) - R with the level's (HAS_RHS) operator; R is the restricted residual sitting in mg_tmp
                launch_rows_mg<MODE_REL, 0>(m, rc, xcur(rc), rc.mg_rhs.p, 1.0, (const double2*)rc.mg_tmp.p);
            }
            fine_work += m->mgb[size_t(l) + 1]->work;
        }
        smooth(nl - 1, nl > 1 ? n_coarsest : nu);
        for (int l = nl - 2; l >= 0; --l) {
            RankList& RF = ranks_of(l);
            RankList& RC = ranks_of(l + 1);
            for (size_t q = 0; q < RF.size(); ++q) {
                RankMesh& rf = *RF[q];
                RankMesh& rc = *RC[q];
                launch_prolong_blocks(m, rf, (const double2*)xcur(rc), (const double2*)rc.E(), xcur(rf));
                if (rf.n_bnd_rows > 0)
                    LAUNCH(mgb_prolong_rows_kernel, (rf.n_bnd_rows + 127) / 128, 128, s, (const BlockXfer*)rf.d_xfer_blocks.p, int(rf.xfer_blocks.size()),
                           (const SmoothedRow*)rf.d_srows.p, int(rf.L.smoothed.size()), (const JunctionRow*)rf.d_jrows.p, int(rf.L.junction_rows.size()),
                           (const SlidingRow*)rf.d_lrows.p, int(rf.L.sliding.size()), (const double2*)xcur(rc), (const double2*)rc.E(), xcur(rf));
            }
            refresh(l);  // nodes no row writes (fixed ones) got a zero correction: both ping-pong buffers still agree there
            smooth(l, nu);
        }
        if (want_change) {
            for (auto& rp : m->ranks) {
                if (track) LAUNCH(mgb_change_kernel, (rp->vec_grid + 127) / 128, 128, s, rp->d_change.p, rp->part_vec.p, rp->vec_grid);
                else LAUNCH(diff_stats_kernel, rp->vec_grid, VEC_THREADS, s, rp->L.n_own, (const double2*)rp->mg_E.p, (const double2*)xcur(*rp), rp->part_vec.p);
            }
            launch_reduce(m, RED_UPDATE_STATS, o, false);
        }
        m->outer_done += 1;
        st->outer_iterations += 1;
        st->inner_iterations += 1;
        if (o->stop_max_update > 0.0) {
            fetch_ctl(m);
            if (m->h_ctl->max_update <= o->stop_max_update) break;
        }
    }
    st->operator_applications += uint64_t(fine_work + 0.5);
}
