// exchange.inl -- CUDA IPC plumbing of the halo exchange over NVLink peer memory (one process per GPU): exporting a
// rank's fields and flag array, mapping the neighbours' copies, agreeing on the transport.  Textually included into the
// anonymous namespace of turbomesh_gpu.cu; the exchange itself (exchange_on) and its kernel (p2p_exchange_kernel) are in
// turbomesh_gpu.cu / kernels.cuh.

// ---- CUDA IPC plumbing of the peer-memory exchange ---------------------------------------------------------------
// All ranks call these in lock-step (they contain NCCL collectives).  Handles travel through an all-reduce of bytes in
// which every rank fills only its own slot; a rank on which anything fails votes the feature off for everybody, so the
// ranks can never disagree about the exchange path (the NCCL send/recv path stays as the alternative).
bool p2p_vote(tm_mesh* m, bool ok) {
    DevBuf<double> d;
    d.alloc(1);
    const double mine = ok ? 1.0 : 0.0;
    CUDA_TRY(cudaMemcpyAsync(d.p, &mine, sizeof mine, cudaMemcpyHostToDevice, m->stream));
    NCCL_TRY(g_nccl.AllReduce(d.p, d.p, 1, ncclDouble, ncclMin, m->comm, m->stream));
    double all = 0.0;
    CUDA_TRY(cudaMemcpyAsync(&all, d.p, sizeof all, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    return all > 0.5;
}

// exports `ptr` of every rank and maps the neighbours' copies into out[p]; returns false (on every rank) if any rank failed
bool p2p_share(tm_mesh* m, RankMesh& r, void* ptr, void** out) {
    const int n = m->n_ranks, me = r.L.rank;
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    std::vector<unsigned char> all(size_t(n) * hs, 0);
    bool ok = ptr != nullptr;
    if (ok) {
        cudaIpcMemHandle_t h;
        if (cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) { (void)cudaGetLastError(); ok = false; }
        else std::memcpy(all.data() + size_t(me) * hs, &h, hs);
    }
    DevBuf<unsigned char> d;
    d.alloc(all.size());
    CUDA_TRY(cudaMemcpyAsync(d.p, all.data(), all.size(), cudaMemcpyHostToDevice, m->stream));
    NCCL_TRY(g_nccl.AllReduce(d.p, d.p, all.size(), ncclUint8, ncclSum, m->comm, m->stream));
    CUDA_TRY(cudaMemcpyAsync(all.data(), d.p, all.size(), cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    ok = p2p_vote(m, ok);
    if (ok) {
        for (int p = 0; p < n; ++p) {
            out[p] = nullptr;
            if (p == me || !((r.p2p.nb_mask >> p) & 1u)) continue;
            cudaIpcMemHandle_t h;
            std::memcpy(&h, all.data() + size_t(p) * hs, hs);
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); ok = false; break; }
            r.p2p.opened.push_back(q);
            out[p] = q;
        }
    }
    return p2p_vote(m, ok);
}

void p2p_setup(tm_mesh* m, RankMesh& r) {
    if (m->n_ranks < 2 || m->emulated || m->n_ranks > P2P_MAX_RANKS) return;
    if (const char* e = std::getenv("TM_P2P")) if (std::atoi(e) == 0) return;
    const int n = m->n_ranks, me = r.L.rank;
    r.p2p.nb_mask = 0;
    for (int p = 0; p < n; ++p)
        if (p != me && (r.L.send_base[size_t(p) + 1] > r.L.send_base[size_t(p)] || r.L.ghost_base[size_t(p) + 1] > r.L.ghost_base[size_t(p)])) r.p2p.nb_mask |= 1u << p;
    r.p2p.flags.alloc(size_t(n)); r.p2p.flags.zero(m->stream);
    r.p2p.counter.alloc(1); r.p2p.counter.zero(m->stream);
    r.p2p.err.alloc(1); r.p2p.err.zero(m->stream);
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    void* tmp[P2P_MAX_RANKS];
    bool ok = p2p_share(m, r, r.p2p.flags.p, tmp);
    if (ok) for (int p = 0; p < n; ++p) r.p2p.peer_flags[p] = static_cast<unsigned long long*>(tmp[p]);
    for (int w = 0; w < 2 && ok; ++w) {
        ok = p2p_share(m, r, r.X[w].p, tmp);
        if (ok) { for (int p = 0; p < n; ++p) r.p2p.peer[w][p] = static_cast<double2*>(tmp[p]); r.p2p.have[w] = true; }
    }
    r.p2p.ready = ok;
}
void p2p_add_tmp(tm_mesh* m, RankMesh& r) {  // the residual scratch field of a multigrid level
    if (!r.p2p.ready || r.p2p.have[2]) return;
    void* tmp[P2P_MAX_RANKS];
    if (p2p_share(m, r, r.mg_tmp.p, tmp)) { for (int p = 0; p < m->n_ranks; ++p) r.p2p.peer[2][p] = static_cast<double2*>(tmp[p]); r.p2p.have[2] = true; }
}
void p2p_check(tm_mesh* m, RankMesh& r) {  // after a synchronisation point: did a wait give up?
    if (!r.p2p.ready) return;
    int e = 0;
    CUDA_TRY(cudaMemcpyAsync(&e, r.p2p.err.p, sizeof e, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (e) TM_THROW(TM_ERR_CUDA, "peer-memory halo exchange timed out waiting for a neighbour rank");
}
