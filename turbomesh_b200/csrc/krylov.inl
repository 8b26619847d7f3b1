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

// ---- two-level preconditioner (krylov_coarse.cuh): aggregates, contribution lists, needed lists, assembly work items ----
int coarse_max_aggregates() {
    int cap = 256;
    if (const char* e = std::getenv("TM_KRYLOV_COARSE_MAX")) cap = std::max(8, std::min(1024, std::atoi(e)));
    return cap;
}
size_t coarse_smem_bytes(int nc_max, int need_max = 0, int src_max = 0) {
    return size_t(4) * size_t(nc_max) * sizeof(double2) + size_t(need_max) * size_t(nc_max) * sizeof(double) + (size_t(need_max) + size_t(nc_max) + 1 + size_t(src_max)) * sizeof(int32_t) + size_t(src_max) * sizeof(double2);
}
bool coarse_enabled() {
    // opt-in: halves the iteration count on the reference's meshes but costs more per iteration than it saves at these sizes, and
    // on meshes with a collapsed cell (LS89 + White after a few outer iterations) its corrections along the nearly singular
    // direction are invisible to the residual test
    if (const char* e = std::getenv("TM_KRYLOV_COARSE")) return std::atoi(e) != 0;
    return false;
}

void coarse_plan_build(tm_mesh* m, RankMesh& r, KrylovPlan& P, const std::vector<KComp>& comps, const std::vector<WTile>& wtiles) {
    P.n_items = 0;
    const Topology& T = m->topo;
    if (!coarse_enabled() || T.n_nodes >= (int64_t(1) << 30)) return;
    cudaStream_t s = m->stream;
    const int n_comp = T.n_comp;
    const int cap = coarse_max_aggregates();
    std::vector<std::vector<size_t>> blocks_of((size_t)n_comp);
    for (size_t b = 0; b < T.blocks.size(); ++b) blocks_of[size_t(T.comp_of_block[b])].push_back(b);
    // patch sizes: the finest of the list that keeps a component within `cap` aggregates (rows even, columns a multiple of 8:
    // a warp tile's two rows and each of its 8-lane segments then lie inside one aggregate)
    static const int cand[][2] = {{8, 8}, {16, 8}, {16, 16}, {32, 16}, {32, 32}, {64, 32}, {64, 64}, {128, 64}, {128, 128}, {256, 128}, {256, 256}};
    int fixed_ai = 0, fixed_aj = 0;
    if (const char* e = std::getenv("TM_KRYLOV_COARSE_PATCH")) {
        if (std::sscanf(e, "%dx%d", &fixed_ai, &fixed_aj) != 2 || fixed_ai < 2 || fixed_ai % 2 || fixed_aj < 8 || fixed_aj % 8) fixed_ai = fixed_aj = 0;
    }
    std::vector<int32_t> agg(size_t(T.n_nodes), -1);
    std::vector<KCoarse> coarse((size_t)n_comp, KCoarse{});
    std::vector<int32_t> agg_block;                       // per aggregate (mesh-wide numbering)
    struct Patch { int ai, aj, gi, gj, base; };
    std::vector<Patch> patch(T.blocks.size(), Patch{0, 0, 0, 0, 0});
    auto counts = [&](const BlockInfo& B, int ai, int aj, int& gi, int& gj) {
        gi = int(std::max<int64_t>(1, (B.ni - 2 + ai / 2) / ai));
        gj = int(std::max<int64_t>(1, (B.nj - 2 + aj / 2) / aj));
    };
    int64_t g_size = 0;
    int nc_max = 0;
    for (int c = 0; c < n_comp; ++c) {
        bool ok = true;
        for (size_t b : blocks_of[size_t(c)]) if (T.blocks[b].ni < 3 || T.blocks[b].nj < 3) ok = false;
        if (!ok) continue;
        int ai = 0, aj = 0;
        for (const auto& cd : cand) {
            int64_t n = 0;
            for (size_t b : blocks_of[size_t(c)]) { int gi, gj; counts(T.blocks[b], cd[0], cd[1], gi, gj); n += int64_t(gi) * gj; }
            if (n <= cap) { ai = cd[0]; aj = cd[1]; break; }
        }
        if (fixed_ai) { ai = fixed_ai; aj = fixed_aj; }
        if (!ai) continue;
        int nc = 0;
        for (size_t b : blocks_of[size_t(c)]) {
            int gi, gj; counts(T.blocks[b], ai, aj, gi, gj);
            patch[b] = Patch{ai, aj, gi, gj, nc};
            nc += gi * gj;
        }
        if (nc > 1024 || nc < 2) { for (size_t b : blocks_of[size_t(c)]) patch[b] = Patch{0, 0, 0, 0, 0}; continue; }
        KCoarse& C = coarse[size_t(c)];
        C.nc = nc;
        C.agg_base = int32_t(agg_block.size());
        C.g_off = g_size;
        g_size += int64_t(nc) * nc;
        nc_max = std::max(nc_max, nc);
        for (size_t b : blocks_of[size_t(c)]) for (int k = 0; k < patch[b].gi * patch[b].gj; ++k) agg_block.push_back(int32_t(b));
    }
    if (nc_max == 0) return;
    // aggregate of the interior node nearest to (i, j) of block b
    auto patch_of = [&](size_t b, int64_t i, int64_t j) -> int32_t {
        const Patch& p = patch[b];
        if (!p.ai) return -1;
        const auto& B = T.blocks[b];
        const int64_t ic = std::min(std::max<int64_t>(i, 1), B.ni - 2) - 1, jc = std::min(std::max<int64_t>(j, 1), B.nj - 2) - 1;
        return int32_t(p.base + std::min<int64_t>(ic / p.ai, p.gi - 1) * p.gj + std::min<int64_t>(jc / p.aj, p.gj - 1));
    };
    auto patch_of_node = [&](int64_t g) -> int32_t {
        const size_t b = T.block_of(g);
        const int64_t l = g - T.blocks[b].off;
        return patch_of(b, l / T.blocks[b].nj, l % T.blocks[b].nj);
    };
    for (size_t b = 0; b < T.blocks.size(); ++b) {
        if (!patch[b].ai) continue;
        const auto& B = T.blocks[b];
        for (int64_t i = 1; i <= B.ni - 2; ++i)
            for (int64_t j = 1; j <= B.nj - 2; ++j) agg[size_t(B.off + i * B.nj + j)] = patch_of(b, i, j);
    }
    for (const SmoothedRow& row : r.L.smoothed) agg[size_t(row.g0)] = patch_of_node(row.g0);
    for (const JunctionRow& row : r.L.junction_rows) agg[size_t(row.self)] = patch_of_node(row.self);
    for (const SlaveRow& sl : r.L.slaves) agg[size_t(sl.self)] = agg[size_t(sl.root)];
    // members (the rows of an aggregate), contribution slots, neighbour aggregates
    const int32_t n_agg = int32_t(agg_block.size());
    std::vector<std::vector<int32_t>> members((size_t)n_agg), slots((size_t)n_agg), nbrs((size_t)n_agg);
    auto add_unique = [](std::vector<int32_t>& v, int32_t x) { if (x >= 0 && std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
    auto srow_nodes = [](const SmoothedRow& row, int64_t (&out)[9]) {
        out[0] = row.g0; out[1] = row.g0 - row.d0; out[2] = row.g0 + row.d0; out[3] = row.g0 + row.n0; out[4] = row.g0 - row.d0 + row.n0;
        out[5] = row.g0 + row.d0 + row.n0; out[6] = row.iN; out[7] = row.iNW; out[8] = row.iNE;
    };
    int64_t n_slots = 0;
    std::vector<int32_t> need_ptr(1, 0), need;
    const int gwarps = P.group_ctas * K_WARPS;
    for (int c = 0; c < n_comp; ++c) {
        KCoarse& C = coarse[size_t(c)];
        const KComp& K = comps[size_t(c)];
        C.slot_base = int32_t(n_slots);
        C.need_base = int32_t(need_ptr.size() - 1);
        const int n_tiles = K.wt_end - K.wt_begin;
        const int n_s = K.s_end - K.s_begin, n_j = K.j_end - K.j_begin, n_l = K.l_end - K.l_begin;
        n_slots += int64_t(4) * n_tiles + n_s + n_j + n_l;
        if (n_slots >= 0x7fffffff) TM_THROW(TM_ERR_UNSUPPORTED, "internal: too many contribution slots");
        if (C.nc == 0) { for (int k = 0; k < P.group_ctas; ++k) need_ptr.push_back(int32_t(need.size())); continue; }
        std::vector<std::vector<int32_t>> need_of((size_t)P.group_ctas);
        const int32_t ab = C.agg_base;
        for (int w = K.wt_begin; w < K.wt_end; ++w) {
            const WTile& t = wtiles[size_t(w)];
            const auto& B = T.blocks[size_t(t.block)];
            const int crank = ((w - K.wt_begin) % (gwarps - K.bnd_warps)) / K_WARPS;
            for (int seg = 0; seg < 4; ++seg) {
                const int64_t j = t.j0 + 8 * seg;
                if (j > B.nj - 2) break;
                slots[size_t(ab + patch_of(size_t(t.block), t.i0, j))].push_back(int32_t(C.slot_base + 4 * (w - K.wt_begin) + seg));
            }
            for (int64_t i = t.i0 - 1; i <= std::min<int64_t>(t.i0 + t.rows, B.ni - 1); ++i)
                for (int64_t j = t.j0 - 1; j <= std::min<int64_t>(t.j0 + 32, B.nj - 1); ++j) {
                    const int32_t a0 = agg[size_t(B.off + i * B.nj + j)];
                    add_unique(need_of[size_t(crank)], a0);
                    if (i >= t.i0 && i < t.i0 + t.rows && j >= t.j0 && j < t.j0 + 32 && j <= B.nj - 2) {   // a row of the tile: its 3 x 3 neighbourhood
                        const int32_t I = agg[size_t(B.off + i * B.nj + j)];
                        members[size_t(ab + I)].push_back(int32_t(B.off + i * B.nj + j));
                        for (int di = -1; di <= 1; ++di)
                            for (int dj = -1; dj <= 1; ++dj) add_unique(nbrs[size_t(ab + I)], agg[size_t(B.off + (i + di) * B.nj + j + dj)]);
                    }
                }
        }
        auto bnd_crank = [&](int q) {   // the kernel's for_bnd: which CTA of the group evaluates boundary row q of the component
            int back = 0;
            if (K.bw_s == 0 && K.bw_j == 0) back = q % std::max(K.bnd_warps, 1);
            else if (q < n_s) back = q % K.bw_s;
            else if (q < n_s + n_j) back = K.bw_s + (q - n_s) % K.bw_j;
            else back = K.bw_s + K.bw_j + (q - n_s - n_j) % (K.bnd_warps - K.bw_s - K.bw_j);
            return (gwarps - 1 - back) / K_WARPS;
        };
        for (int q = 0; q < n_s; ++q) {
            const SmoothedRow& row = r.L.smoothed[size_t(K.s_begin + q)];
            const int32_t I = agg[size_t(row.g0)];
            slots[size_t(ab + I)].push_back(int32_t(C.slot_base + 4 * n_tiles + q));
            members[size_t(ab + I)].push_back(int32_t((1u << 30) | uint32_t(K.s_begin + q)));
            int64_t nodes[9];
            srow_nodes(row, nodes);
            for (int64_t g : nodes) { add_unique(need_of[size_t(bnd_crank(q))], agg[size_t(g)]); add_unique(nbrs[size_t(ab + I)], agg[size_t(g)]); }
        }
        for (int q = 0; q < n_j; ++q) {
            const JunctionRow& row = r.L.junction_rows[size_t(K.j_begin + q)];
            const int32_t I = agg[size_t(row.self)];
            slots[size_t(ab + I)].push_back(int32_t(C.slot_base + 4 * n_tiles + n_s + q));
            members[size_t(ab + I)].push_back(int32_t((2u << 30) | uint32_t(K.j_begin + q)));
            const int cr = bnd_crank(n_s + q);
            add_unique(need_of[size_t(cr)], I); add_unique(nbrs[size_t(ab + I)], I);
            for (int k = 0; k < row.n; ++k) { add_unique(need_of[size_t(cr)], agg[size_t(row.nbr[k])]); add_unique(nbrs[size_t(ab + I)], agg[size_t(row.nbr[k])]); }
        }
        for (int q = 0; q < n_l; ++q) {
            const SlidingRow& row = r.L.sliding[size_t(K.l_begin + q)];
            add_unique(need_of[size_t(bnd_crank(n_s + n_j + q))], agg[size_t(row.inner)]);
        }
        for (int k = 0; k < P.group_ctas; ++k) {
            std::sort(need_of[size_t(k)].begin(), need_of[size_t(k)].end());
            need.insert(need.end(), need_of[size_t(k)].begin(), need_of[size_t(k)].end());
            need_ptr.push_back(int32_t(need.size()));
        }
    }
    std::vector<int32_t> contrib_ptr(1, 0), contrib_src, mem_ptr(1, 0), mem_code;
    std::vector<CoarseItem> items;
    for (int c = 0; c < n_comp; ++c) {
        const KCoarse& C = coarse[size_t(c)];
        for (int I = 0; I < C.nc; ++I) {
            const size_t g = size_t(C.agg_base + I);
            contrib_src.insert(contrib_src.end(), slots[g].begin(), slots[g].end());
            contrib_ptr.push_back(int32_t(contrib_src.size()));
            mem_code.insert(mem_code.end(), members[g].begin(), members[g].end());
            mem_ptr.push_back(int32_t(mem_code.size()));
            std::sort(nbrs[g].begin(), nbrs[g].end());
            for (int32_t J : nbrs[g]) items.push_back(CoarseItem{c, I, J});
        }
    }
    P.n_items = int(items.size());
    P.nc_max = nc_max;
    P.n_slots = n_slots;
    P.g_size = g_size;
    // shared-memory caches of the static tables, as far as they fit (the rows of G first)
    int need_max = 0, src_max = 0;
    for (size_t k = 0; k + 1 < need_ptr.size(); ++k) need_max = std::max(need_max, need_ptr[k + 1] - need_ptr[k]);
    for (int c = 0; c < n_comp; ++c)
        if (coarse[size_t(c)].nc) src_max = std::max(src_max, contrib_ptr[size_t(coarse[size_t(c)].agg_base + coarse[size_t(c)].nc)] - contrib_ptr[size_t(coarse[size_t(c)].agg_base)]);
    const size_t smem_cap = 190 * 1024;
    if (std::getenv("TM_KRYLOV_COARSE_NOCACHE")) need_max = src_max = 0;
    if (coarse_smem_bytes(nc_max, need_max, 0) > smem_cap) need_max = 0;
    if (coarse_smem_bytes(nc_max, need_max, src_max) > smem_cap) src_max = 0;
    P.need_max = need_max; P.src_max = src_max;
    P.smem = coarse_smem_bytes(nc_max, need_max, src_max);
    P.coarse_every = 1;
    if (const char* e = std::getenv("TM_KRYLOV_COARSE_EVERY")) P.coarse_every = std::max(1, std::atoi(e));
    P.coarse.upload(coarse, s);
    P.agg.upload(agg, s);
    P.agg_block.upload(agg_block, s);
    P.contrib_ptr.upload(contrib_ptr, s); P.contrib_src.upload(contrib_src, s);
    P.need_ptr.upload(need_ptr, s); P.need.upload(need, s);
    P.mem_ptr.upload(mem_ptr, s); P.mem_code.upload(mem_code, s);
    P.items.upload(items, s);
    P.G.alloc(size_t(g_size)); P.G.zero(s);
    P.contrib.alloc(size_t(2) * size_t(n_slots)); P.contrib.zero(s);
    P.coarse_ok.alloc(size_t(n_comp)); P.coarse_ok.zero(s);
    P.h_coarse_ok.assign(size_t(n_comp), 0);
    P.coarse_age = -1;
    CUDA_TRY(cudaFuncSetAttribute(coarse_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * 1024 * sizeof(double))));
}

// G = (P^T A P)^-1 with the lagged coefficients of this outer iteration (X[cur], pq)
template <class Plan>
void coarse_refresh(tm_mesh* m, RankMesh& r, Plan& P) {
    cudaStream_t s = m->stream;
    CUDA_TRY(cudaMemsetAsync(P.G.p, 0, size_t(P.g_size) * sizeof(double), s));
    const unsigned grid = unsigned((int64_t(P.n_items) * 32 + 255) / 256);
    if (r.has_pq)
        LAUNCH(coarse_assemble_kernel<true>, grid, 256, s, (const CoarseItem*)P.items.p, P.n_items, (const KCoarse*)P.coarse.p, (const int32_t*)P.mem_ptr.p, (const int32_t*)P.mem_code.p,
               (const int32_t*)P.agg_block.p, (const int32_t*)P.agg.p, (const DevBlock*)r.d_blocks.p, (const SmoothedRow*)r.d_srows.p, (const JunctionRow*)r.d_jrows.p,
               (const double2*)r.X[r.cur].p, (const double2*)r.pq.p, P.G.p, int(std::is_same<Plan, PhasedPlan>::value));
    else
        LAUNCH(coarse_assemble_kernel<false>, grid, 256, s, (const CoarseItem*)P.items.p, P.n_items, (const KCoarse*)P.coarse.p, (const int32_t*)P.mem_ptr.p, (const int32_t*)P.mem_code.p,
               (const int32_t*)P.agg_block.p, (const int32_t*)P.agg.p, (const DevBlock*)r.d_blocks.p, (const SmoothedRow*)r.d_srows.p, (const JunctionRow*)r.d_jrows.p,
               (const double2*)r.X[r.cur].p, (const double2*)r.pq.p, P.G.p, int(std::is_same<Plan, PhasedPlan>::value));
    coarse_invert_kernel<<<unsigned(m->topo.n_comp), COARSE_INV_THREADS, 2 * size_t(P.nc_max) * sizeof(double), s>>>((const KCoarse*)P.coarse.p, P.G.p, P.coarse_ok.p);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
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
    if (coarse_enabled()) {
        if (r.has_pq) { CUDA_TRY(cudaFuncSetAttribute(bicgstab_persistent_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(200 * 1024))); }
        else { CUDA_TRY(cudaFuncSetAttribute(bicgstab_persistent_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(200 * 1024))); }
    }
    if (r.has_pq) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bicgstab_persistent_kernel<true, true>, K_THREADS, coarse_enabled() ? size_t(190 * 1024) : size_t(0)));
    else CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bicgstab_persistent_kernel<false, true>, K_THREADS, coarse_enabled() ? size_t(190 * 1024) : size_t(0)));
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
    int group_ctas = std::max(1, std::min(std::min(want_ctas, K_MAX_GROUP), max_ctas / n_groups));
    if (const char* e = std::getenv("TM_KRYLOV_GROUP_CTAS")) group_ctas = std::max(1, std::min(std::min(max_ctas / n_groups, K_MAX_GROUP), std::atoi(e)));
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
    // boundary rows / rhs terms are grouped by component already (build_rank): find the ranges
    std::vector<KComp> comps((size_t)n_comp, KComp{});
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
    // Warp tiles per component.  One row per tile gives the most parallelism, two rows halve the number of tiles: two when the
    // tiles of a phase and the boundary rows (a warp per 32 of them) then fit the group's warps in ONE round -- every warp has
    // one task, the warps left over share the boundary rows (k_bnd_warps) -- or when there are several rounds anyway.
    std::vector<WTile> wtiles;
    const int group_warps = group_ctas * K_WARPS;
    for (int c = 0; c < n_comp; ++c) {
        KComp& K = comps[size_t(c)];
        K.nodes = int32_t(std::min<int64_t>(nodes[size_t(c)], 0x7fffffff));
        auto tiles_with = [&](int64_t rows) {
            int64_t n = 0;
            for (size_t b = 0; b < T.blocks.size(); ++b) {
                if (T.comp_of_block[b] != c) continue;
                const auto& B = T.blocks[b];
                n += ((B.ni - 2 + rows - 1) / rows) * ((B.nj - 2 + 31) / 32);
            }
            return n;
        };
        const int64_t bnd_chunks = (int64_t(K.s_end - K.s_begin) + (K.j_end - K.j_begin) + (K.l_end - K.l_begin) + 31) / 32;
        int rows = 1;   // the fewest rows per tile with which the phase is a single round, else the most
        while (rows < K_TILE_ROWS && tiles_with(rows) + bnd_chunks > group_warps) rows *= 2;
        if (const char* e = std::getenv("TM_KRYLOV_TILE_ROWS")) rows = std::max(1, std::min(K_TILE_ROWS, std::atoi(e)));
        // warps set aside for the boundary rows: all that are left in a single round, else a share by work (32 boundary rows
        // cost about two tiles: three kinds of rows, gathers across blocks)
        const int64_t nt = tiles_with(rows);
        if (bnd_chunks == 0) K.bnd_warps = 0;
        else if (nt + bnd_chunks <= group_warps) K.bnd_warps = int32_t(group_warps - nt);
        else K.bnd_warps = int32_t(std::max<int64_t>(1, std::min<int64_t>(group_warps / 2, (2 * bnd_chunks * group_warps + (nt + 2 * bnd_chunks) / 2) / (nt + 2 * bnd_chunks))));
        if (const char* e = std::getenv("TM_KRYLOV_BND_WARPS")) K.bnd_warps = std::max(bnd_chunks ? 1 : 0, std::min(group_warps - 1, std::atoi(e)));
        // the boundary warps by kind of row, shared out by work (an interface row 1, a junction row 1.5, a sliding row 0.2); every kind
        // present needs a warp, else all kinds share all warps
        K.bw_s = K.bw_j = 0;
        {
            const int n_s = K.s_end - K.s_begin, n_j = K.j_end - K.j_begin, n_l = K.l_end - K.l_begin;
            const int kinds = (n_s > 0) + (n_j > 0) + (n_l > 0);
            if (kinds >= 2 && K.bnd_warps >= 2 * kinds && !std::getenv("TM_KRYLOV_BND_MIXED")) {
                const double ws = n_s, wj = 1.5 * n_j, wl = 0.2 * n_l, total = ws + wj + wl;
                int bj = n_j ? std::max(1, int(K.bnd_warps * wj / total + 0.5)) : 0;
                int bl = n_l ? std::max(1, int(K.bnd_warps * wl / total + 0.5)) : 0;
                int bs = K.bnd_warps - bj - bl;
                if (n_s == 0) { bl += bs; bs = 0; }
                if (bs >= (n_s ? 1 : 0) && (bs > 0 || bj > 0)) { K.bw_s = bs; K.bw_j = bj; }
                if (K.bw_s == 0 && K.bw_j == 0) { K.bw_s = 0; K.bw_j = 0; }
            }
        }
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
    P.h_comps = comps;
    P.h_ctl.assign(size_t(n_comp), KCtl{});
    P.wtiles.upload(wtiles, s);
    P.comps.upload(comps, s);
    P.groups.upload(groups, s);
    P.group_comps.upload(group_comps, s);
    P.cta_group.upload(cta_group, s);
    P.ctl.alloc(size_t(n_comp)); P.ctl.zero(s);
    P.bars.alloc(size_t(n_groups)); P.bars.zero(s);
    P.partials.alloc(size_t(2) * size_t(P.n_ctas) * K_REC); P.partials.zero(s);
    coarse_plan_build(m, r, P, comps, wtiles);
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
    a.ctl = P.ctl.p; a.bars = P.bars.p; a.partials = P.partials.p; a.epoch = ++P.epoch;
    a.xc = r.X[r.cur].p; a.pq = r.pq.p; a.xnew = r.X[1 - r.cur].p;
    a.r = r.kr.p; a.rhat = r.krhat.p; a.p[0] = r.kp.p; a.p[1] = r.kp2.p; a.v[0] = r.kv.p; a.v[1] = r.kv2.p; a.s = r.ks.p; a.t = r.kt.p; a.d = r.kd.p;
    a.rtol = o->rtol; a.atol = o->atol;
    a.max_iters = o->max_inner_iterations > 0x7fffffffull ? 0x7fffffff : int32_t(o->max_inner_iterations);
    a.max_restarts = 60;
    a.n_ctas_total = P.n_ctas;
    a.polish = std::max(0, int(o->inner_refinement_cycles));
    const bool coarse = P.n_items > 0;
    if (coarse) {
        if (P.coarse_age < 0 || P.coarse_age >= P.coarse_every) { coarse_refresh(m, r, P); P.coarse_age = 0; }
        P.coarse_age += 1;
        a.coarse = P.coarse.p; a.coarse_ok = P.coarse_ok.p; a.agg = P.agg.p; a.contrib_ptr = P.contrib_ptr.p; a.contrib_src = P.contrib_src.p;
        a.need_ptr = P.need_ptr.p; a.need = P.need.p; a.G = P.G.p; a.contrib = P.contrib.p; a.n_slots = P.n_slots; a.nc_max = P.nc_max; a.need_max = P.need_max; a.src_max = P.src_max;
    }
    const bool timing = std::getenv("TM_KRYLOV_TIMING") != nullptr;
    if (timing) {
        if (!P.timing.p) P.timing.alloc(size_t(P.n_ctas) * 8);
        P.timing.zero(s);
        a.timing = P.timing.p;
    }
    void* params[] = {&a};
    const void* fn = coarse ? (r.has_pq ? (const void*)bicgstab_persistent_kernel<true, true> : (const void*)bicgstab_persistent_kernel<false, true>)
                            : (r.has_pq ? (const void*)bicgstab_persistent_kernel<true, false> : (const void*)bicgstab_persistent_kernel<false, false>);
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(unsigned(P.n_ctas)), dim3(K_THREADS), params, coarse ? P.smem : 0, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaMemcpyAsync(P.h_ctl.data(), P.ctl.p, P.h_ctl.size() * sizeof(KCtl), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (timing) {
        std::vector<long long> t(size_t(P.n_ctas) * 8);
        CUDA_TRY(cudaMemcpy(t.data(), P.timing.p, t.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        double sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < P.n_ctas; ++c) for (int k = 0; k < 8; ++k) sum[k] += double(t[size_t(c) * 8 + size_t(k)]) / P.n_ctas;
        for (int k : {1, 4}) {
            long long mn = t[size_t(k)], mx = mn; int imx = 0, imn = 0;
            for (int c = 0; c < P.n_ctas; ++c) { const long long v = t[size_t(c) * 8 + size_t(k)]; if (v > mx) { mx = v; imx = c; } if (v < mn) { mn = v; imn = c; } }
            std::fprintf(stderr, "  slot %d: min %lld (CTA %d) max %lld (CTA %d);", k, mn, imn, mx, imx);
            for (int c = 0; c < P.n_ctas; c += (std::getenv("TM_KRYLOV_TIMING_ALL") ? 1 : std::max(1, P.n_ctas / 12))) std::fprintf(stderr, " %lld", t[size_t(c) * 8 + size_t(k)] / 1000);
            std::fprintf(stderr, "\n");
        }
        std::fprintf(stderr, "krylov timing (mean cycles per CTA): R0 %.0f  A %.0f  B %.0f  C %.0f  barrier+sums %.0f  coarse %.0f  rest %.0f  (%d CTAs; coarse: %d aggregates, needed lists <= %d, slot lists <= %d, %zu B shared)\n", sum[0], sum[1], sum[2], sum[3], sum[4],
                     sum[5], sum[6], P.n_ctas, P.nc_max, P.need_max, P.src_max, P.smem);
    }
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
// Coarse space of the phased path: an aggregate is a segment group of a warp tile (the tile's rows x 8, 16 or 32 columns), so
// that a tile's restriction is its segment sums; boundary rows as in coarse_plan_build.  Slots: 4 per warp tile, then one
// per smoothed row, then one per junction row (table order).
bool coarse_enabled_batch() {
    if (const char* e = std::getenv("TM_KRYLOV_COARSE")) return std::atoi(e) != 0;
    return true;
}
void phased_coarse_build(tm_mesh* m, RankMesh& r, PhasedPlan& P, const std::vector<KPComp>& comps, const std::vector<WTile>& wtiles, int tile_rows) {
    P.n_items = 0;
    const Topology& T = m->topo;
    if (!coarse_enabled_batch() || tile_rows < 8 || tile_rows % 2 || T.n_nodes >= (int64_t(1) << 30)) return;
    cudaStream_t s = m->stream;
    const int n_comp = T.n_comp;
    int cap = 512;
    if (const char* e = std::getenv("TM_KRYLOV_COARSE_MAX")) cap = std::max(8, std::min(1024, std::atoi(e)));
    std::vector<std::vector<size_t>> blocks_of((size_t)n_comp);
    for (size_t b = 0; b < T.blocks.size(); ++b) blocks_of[size_t(T.comp_of_block[b])].push_back(b);
    struct Patch { int ai, aj, gi, gj, base; };
    std::vector<Patch> patch(T.blocks.size(), Patch{0, 0, 0, 0, 0});
    auto rows_of = [&](const BlockInfo& B) {   // the tile height phased_plan_build gives the block
        const int64_t interior_i = B.ni - 2;
        const int64_t n_i = std::max<int64_t>(1, (interior_i + tile_rows - 1) / tile_rows);
        return int((interior_i + n_i - 1) / n_i);
    };
    std::vector<KCoarse> coarse((size_t)n_comp, KCoarse{});
    std::vector<int32_t> agg_block;
    int64_t g_size = 0;
    int nc_max = 0;
    for (int c = 0; c < n_comp; ++c) {
        bool ok = true;
        for (size_t b : blocks_of[size_t(c)]) if (T.blocks[b].ni < 3 || T.blocks[b].nj < 3) ok = false;
        if (!ok) continue;
        int aj = 0, nc = 0;
        for (int cand : {8, 16, 32}) {
            int64_t n = 0;
            for (size_t b : blocks_of[size_t(c)]) {
                const auto& B = T.blocks[b];
                const int ai = rows_of(B);
                n += ((B.ni - 2 + ai - 1) / ai) * std::max<int64_t>(1, (B.nj - 2 + cand / 2) / cand);
            }
            if (n <= cap) { aj = cand; nc = int(n); break; }
        }
        if (!aj || nc < 2) continue;
        int base = 0;
        for (size_t b : blocks_of[size_t(c)]) {
            const auto& B = T.blocks[b];
            const int ai = rows_of(B);
            const int gi = int((B.ni - 2 + ai - 1) / ai), gj = int(std::max<int64_t>(1, (B.nj - 2 + aj / 2) / aj));
            patch[b] = Patch{ai, aj, gi, gj, base};
            base += gi * gj;
        }
        KCoarse& C = coarse[size_t(c)];
        C.nc = nc; C.agg_base = int32_t(agg_block.size()); C.g_off = g_size;
        g_size += int64_t(nc) * nc;
        nc_max = std::max(nc_max, nc);
        for (size_t b : blocks_of[size_t(c)]) for (int k = 0; k < patch[b].gi * patch[b].gj; ++k) agg_block.push_back(int32_t(b));
    }
    if (nc_max == 0) return;
    auto patch_of = [&](size_t b, int64_t i, int64_t j) -> int32_t {
        const Patch& p = patch[b];
        if (!p.ai) return -1;
        const auto& B = T.blocks[b];
        const int64_t ic = std::min(std::max<int64_t>(i, 1), B.ni - 2) - 1, jc = std::min(std::max<int64_t>(j, 1), B.nj - 2) - 1;
        return int32_t(p.base + std::min<int64_t>(ic / p.ai, p.gi - 1) * p.gj + std::min<int64_t>(jc / p.aj, p.gj - 1));
    };
    auto patch_of_node = [&](int64_t g) -> int32_t {
        const size_t b = T.block_of(g);
        const int64_t l = g - T.blocks[b].off;
        return patch_of(b, l / T.blocks[b].nj, l % T.blocks[b].nj);
    };
    std::vector<int32_t> agg(size_t(T.n_nodes), -1);
    for (size_t b = 0; b < T.blocks.size(); ++b) {
        if (!patch[b].ai) continue;
        const auto& B = T.blocks[b];
        for (int64_t i = 1; i <= B.ni - 2; ++i)
            for (int64_t j = 1; j <= B.nj - 2; ++j) agg[size_t(B.off + i * B.nj + j)] = patch_of(b, i, j);
    }
    for (const SmoothedRow& row : r.L.smoothed) agg[size_t(row.g0)] = patch_of_node(row.g0);
    for (const JunctionRow& row : r.L.junction_rows) agg[size_t(row.self)] = patch_of_node(row.self);
    for (const SlaveRow& sl : r.L.slaves) agg[size_t(sl.self)] = agg[size_t(sl.root)];
    const int32_t n_agg = int32_t(agg_block.size());
    std::vector<std::vector<int32_t>> members((size_t)n_agg), slots((size_t)n_agg), nbrs((size_t)n_agg);
    auto add_unique = [](std::vector<int32_t>& v, int32_t x) { if (x >= 0 && std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
    const int64_t sslot0 = int64_t(4) * int64_t(wtiles.size()), jslot0 = sslot0 + int64_t(r.L.smoothed.size());
    const int64_t n_slots = jslot0 + int64_t(r.L.junction_rows.size());
    if (n_slots >= 0x7fffffff) return;
    for (size_t w = 0; w < wtiles.size(); ++w) {
        const WTile& t = wtiles[w];
        if (!patch[size_t(t.block)].ai) continue;
        const auto& B = T.blocks[size_t(t.block)];
        const int32_t ab = coarse[size_t(T.comp_of_block[size_t(t.block)])].agg_base;
        // the four 8-lane segments of the warp: columns j0 + 8 s of one row group (plain tile) or row group s of a strip tile
        const int groups = wtile_groups(t), rr = wtile_rows(t), segs_per_group = groups ? wtile_width(t) / 8 : 4;
        for (int seg = 0; seg < 4; ++seg) {
            const int g = groups ? seg / segs_per_group : 0;
            const int64_t i_first = t.i0 + int64_t(g) * rr, j_first = t.j0 + 8 * (seg - g * segs_per_group);
            if (j_first > B.nj - 2 || i_first > B.ni - 2 || (groups && g >= groups)) continue;
            slots[size_t(ab + patch_of(size_t(t.block), i_first, j_first))].push_back(int32_t(4 * w + size_t(seg)));
            const int64_t i_last = std::min<int64_t>(i_first + rr, B.ni - 1), j_last = std::min<int64_t>(j_first + 8, B.nj - 1);
            for (int64_t i = i_first; i < i_last; ++i)
                for (int64_t j = j_first; j < j_last; ++j) {
                    const int32_t I = agg[size_t(B.off + i * B.nj + j)];
                    members[size_t(ab + I)].push_back(int32_t(B.off + i * B.nj + j));
                    for (int di = -1; di <= 1; ++di)
                        for (int dj = -1; dj <= 1; ++dj) add_unique(nbrs[size_t(ab + I)], agg[size_t(B.off + (i + di) * B.nj + j + dj)]);
                }
        }
    }
    for (size_t q = 0; q < r.L.smoothed.size(); ++q) {
        const SmoothedRow& row = r.L.smoothed[q];
        const int32_t I = agg[size_t(row.g0)];
        if (I < 0) continue;
        const int32_t ab = coarse[size_t(T.comp_of_block[T.block_of(row.g0)])].agg_base;
        slots[size_t(ab + I)].push_back(int32_t(sslot0 + int64_t(q)));
        members[size_t(ab + I)].push_back(int32_t((1u << 30) | uint32_t(q)));
        const int64_t nodes[9] = {row.g0, row.g0 - row.d0, row.g0 + row.d0, row.g0 + row.n0, row.g0 - row.d0 + row.n0, row.g0 + row.d0 + row.n0, row.iN, row.iNW, row.iNE};
        for (int64_t g : nodes) add_unique(nbrs[size_t(ab + I)], agg[size_t(g)]);
    }
    for (size_t q = 0; q < r.L.junction_rows.size(); ++q) {
        const JunctionRow& row = r.L.junction_rows[q];
        const int32_t I = agg[size_t(row.self)];
        if (I < 0) continue;
        const int32_t ab = coarse[size_t(T.comp_of_block[T.block_of(row.self)])].agg_base;
        slots[size_t(ab + I)].push_back(int32_t(jslot0 + int64_t(q)));
        members[size_t(ab + I)].push_back(int32_t((2u << 30) | uint32_t(q)));
        add_unique(nbrs[size_t(ab + I)], I);
        for (int k = 0; k < row.n; ++k) add_unique(nbrs[size_t(ab + I)], agg[size_t(row.nbr[k])]);
    }
    std::vector<int32_t> contrib_ptr(1, 0), contrib_src, mem_ptr(1, 0), mem_code;
    std::vector<CoarseItem> items;
    for (int c = 0; c < n_comp; ++c) {
        const KCoarse& C = coarse[size_t(c)];
        for (int I = 0; I < C.nc; ++I) {
            const size_t g = size_t(C.agg_base + I);
            contrib_src.insert(contrib_src.end(), slots[g].begin(), slots[g].end());
            contrib_ptr.push_back(int32_t(contrib_src.size()));
            mem_code.insert(mem_code.end(), members[g].begin(), members[g].end());
            mem_ptr.push_back(int32_t(mem_code.size()));
            std::sort(nbrs[g].begin(), nbrs[g].end());
            for (int32_t J : nbrs[g]) items.push_back(CoarseItem{c, I, J});
        }
    }
    (void)comps;
    P.n_items = int(items.size());
    P.nc_max = nc_max; P.g_size = g_size; P.sslot0 = int(sslot0); P.jslot0 = int(jslot0);
    P.coarse_every = 1;
    if (const char* e = std::getenv("TM_KRYLOV_COARSE_EVERY")) P.coarse_every = std::max(1, std::atoi(e));
    P.coarse.upload(coarse, s); P.agg.upload(agg, s); P.agg_block.upload(agg_block, s);
    P.contrib_ptr.upload(contrib_ptr, s); P.contrib_src.upload(contrib_src, s);
    P.mem_ptr.upload(mem_ptr, s); P.mem_code.upload(mem_code, s);
    P.items.upload(items, s);
    P.G.alloc(size_t(g_size)); P.G.zero(s);
    P.contrib.alloc(size_t(n_slots)); P.contrib.zero(s);
    P.e_r.alloc(size_t(n_agg)); P.e_r.zero(s); P.e_v.alloc(size_t(n_agg)); P.e_v.zero(s); P.e_t.alloc(size_t(n_agg)); P.e_t.zero(s);
    P.coarse_ok.alloc(size_t(n_comp)); P.coarse_ok.zero(s);
    P.coarse_age = -1;
    CUDA_TRY(cudaFuncSetAttribute(coarse_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * 1024 * sizeof(double))));
}

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
            // the last <= 16 columns of a block would fill at most half of a warp: four or two row groups of them share one (strip tiles)
            const int64_t rem = (B.nj - 2) % 32, n_full = (B.nj - 2) / 32;
            const bool strips = rem > 0 && rem <= 16 && n_i >= 2 && !std::getenv("TM_KRYLOV_NO_STRIPS");
            const int64_t strip_w = rem <= 8 ? 8 : 16, per_warp = 32 / strip_w;
            for (int64_t i0 = 1; i0 <= B.ni - 2; i0 += rr)
                for (int64_t j0 = 1; j0 <= B.nj - 2; j0 += 32) {
                    if (strips && j0 == 1 + 32 * n_full) continue;
                    wtiles.push_back(WTile{int32_t(b), int32_t(i0), int32_t(j0), int32_t(std::min<int64_t>(rr, B.ni - 1 - i0))});
                    wt_comp.push_back(c);
                }
            if (strips)
                for (int64_t g0 = 0; g0 < n_i; g0 += per_warp) {
                    wtiles.push_back(WTile{int32_t(b), int32_t(1 + g0 * rr), int32_t(1 + 32 * n_full), int32_t(rr | (std::min<int64_t>(per_warp, n_i - g0) << 16) | (strip_w << 24))});
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
    phased_coarse_build(m, r, P, comps, wtiles, tile_rows);
    CUDA_TRY(cudaMallocHost(&P.h_count, 2 * sizeof(int)));
    CUDA_TRY(cudaStreamSynchronize(s));
}

template <int PHASE>
void phased_launch(tm_mesh* m, RankMesh& r, const KPArgs& a) {
    const PhasedPlan& P = *r.pplan;
    const unsigned g_tiles = unsigned((P.n_wtiles + KP_WARPS - 1) / KP_WARPS), g_chunks = unsigned(P.n_chunks);
    const bool co = a.coarse != nullptr && PHASE != KP_ADD;
    // the boundary chunks neither read what the tiles of the same phase write nor write the same nodes: they run beside them
    // (the mesh's second, high-priority stream; inside the captured graph this is a fork and a join)
    cudaStream_t st = m->stream, sc = m->stream;
    if (g_tiles && g_chunks && !std::getenv("TM_KRYLOV_SERIAL_CHUNKS")) {
        sc = m->comm_stream;
        CUDA_TRY(cudaEventRecord(m->ev_rim, st));
        CUDA_TRY(cudaStreamWaitEvent(sc, m->ev_rim, 0));
    }
    if (r.has_pq) {
        if (co) {
            if (g_tiles) LAUNCH((krylov_phase_kernel<PHASE, true, true, true>), g_tiles, KP_THREADS, st, a);
            if (g_chunks) LAUNCH((krylov_phase_kernel<PHASE, true, false, true>), g_chunks, KP_THREADS, sc, a);
        } else {
            if (g_tiles) LAUNCH((krylov_phase_kernel<PHASE, true, true, false>), g_tiles, KP_THREADS, st, a);
            if (g_chunks) LAUNCH((krylov_phase_kernel<PHASE, true, false, false>), g_chunks, KP_THREADS, sc, a);
        }
    } else {
        if (co) {
            if (g_tiles) LAUNCH((krylov_phase_kernel<PHASE, false, true, true>), g_tiles, KP_THREADS, st, a);
            if (g_chunks) LAUNCH((krylov_phase_kernel<PHASE, false, false, true>), g_chunks, KP_THREADS, sc, a);
        } else {
            if (g_tiles) LAUNCH((krylov_phase_kernel<PHASE, false, true, false>), g_tiles, KP_THREADS, st, a);
            if (g_chunks) LAUNCH((krylov_phase_kernel<PHASE, false, false, false>), g_chunks, KP_THREADS, sc, a);
        }
    }
    if (sc != st) {
        CUDA_TRY(cudaEventRecord(m->ev_x, sc));
        CUDA_TRY(cudaStreamWaitEvent(st, m->ev_x, 0));
    }
    if (PHASE != KP_ADD) LAUNCH((krylov_finalize_kernel<PHASE>), unsigned((P.n_comp + 3) / 4), 128, m->stream, a);
    if (co) {
        krylov_coarse_kernel<PHASE><<<unsigned(P.n_comp), KC_THREADS, size_t(P.nc_max) * (1 + KC_SPLIT) * sizeof(double2), m->stream>>>(a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CUDA_TRY(cudaGetLastError());
    }
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
    // Without TM_KRYLOV_COARSE in the environment the coarse space is used with the reference's kind of settings only: "exact
    // Picard step" runs (inner_refinement_cycles > 0, parity against the oracle) stay with point-Jacobi, whose iterates never
    // move along the nearly singular direction of a mesh with a collapsed cell.
    const bool use_coarse = P.n_items > 0 && (std::getenv("TM_KRYLOV_COARSE") != nullptr || o->inner_refinement_cycles == 0);
    if (use_coarse) {
        if (P.coarse_age < 0 || P.coarse_age >= P.coarse_every) { coarse_refresh(m, r, P); P.coarse_age = 0; }
        P.coarse_age += 1;
        a.coarse = P.coarse.p; a.coarse_ok = P.coarse_ok.p; a.agg = P.agg.p; a.contrib_ptr = P.contrib_ptr.p; a.contrib_src = P.contrib_src.p;
        a.G = P.G.p; a.contrib = P.contrib.p; a.e_r = P.e_r.p; a.e_v = P.e_v.p; a.e_t = P.e_t.p; a.sslot0 = P.sslot0; a.jslot0 = P.jslot0; a.nc_max = P.nc_max;
    }
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
    // graphs (of two iterations) between two looks at the systems' states: a look costs a stream synchronisation, not looking
    // costs iterations nobody needs; with the coarse space a solve is ~60 iterations, without ~150
    int polls_every = use_coarse ? 2 : 4;
    if (const char* e = std::getenv("TM_KRYLOV_POLL")) polls_every = std::max(1, std::atoi(e));
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
                for (int k = 0; k < polls_every; ++k) {
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
