// turbomesh_gpu.cu -- implementation of the C ABI declared in include/turbomesh_gpu.h.
//
// Host orchestration of the sm_100a kernels in kernels.cuh: device mesh handle, topology upload, the outer
// (Picard) loop of smoothing.smooth.mesh (src/core/smoothing/smooth.zig:74-166), the matrix-free BiCGStab
// (src/core/smoothing/BiCGStab.zig:279-370), the relaxation sweeps and the multi-GPU halo exchange (one process per
// GPU, NCCL send/recv over NVLink once per sweep / operator application).  No CPU compute path exists here: without a
// CUDA device every entry point fails with TM_ERR_NO_DEVICE.
#include <dlfcn.h>
#include <nccl.h>  // types only: NCCL is resolved at run time with dlopen, single-GPU use needs no libnccl

#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/turbomesh_gpu.h"
#include "kernels.cuh"
#include "krylov_coarse.cuh"
#include "krylov_phased.cuh"
#include "io_kernels.cuh"
#include "mg_kernels.cuh"
#include "mg_plan.hpp"
#include "partition.hpp"

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

// Device -> host of per-job slices: consecutive jobs whose host slices are adjacent (a caller that keeps the edges of a batch in
// one buffer) travel in ONE copy instead of one each -- thousands of small copies into pageable memory cost ~15 us apiece.
template <class T, class Host, class Off, class Cnt>
void copy_out_runs(size_t n_jobs, const T* dev, Host host, Off off, Cnt cnt, cudaStream_t s) {
    size_t k = 0;
    while (k < n_jobs) {
        T* const h0 = host(k);
        const int64_t o0 = off(k);
        size_t n = cnt(k), e = k + 1;
        while (e < n_jobs && host(e) == h0 + n && off(e) == o0 + int64_t(n)) { n += cnt(e); ++e; }
        CUDA_TRY(cudaMemcpyAsync(h0, dev + o0, n * sizeof(T), cudaMemcpyDeviceToHost, s));
        k = e;
    }
}

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

// Device allocations are recycled through a small process-wide cache: the one-shot entry points (tm_tfi_block,
// tm_smooth_mesh) create and destroy a device mesh per call, and cudaMalloc / cudaFree of GB-sized fields would otherwise
// cost more than the kernels they bracket -- and the ~25 small tables of a mesh matter as well: every cudaFree is a
// device-wide synchronisation that was measured to take anything from 0.1 to 400 ms on a busy box.  Exact-size reuse
// only; at most TM_CACHE_GB (default 8) GiB and kMaxEntries buffers are kept; tm_release_cached_memory() returns
// everything to the driver.
struct DeviceCache {
    struct Entry { void* p; size_t bytes; int device; };
    std::mutex mu;
    std::vector<Entry> free_list;
    size_t held = 0;
    static constexpr size_t kMinBytes = 1;
    static constexpr size_t kMaxEntries = 1024;
    size_t cap() const {
        static const size_t c = [] { const char* e = std::getenv("TM_CACHE_GB"); return size_t((e ? std::atof(e) : 8.0) * double(size_t(1) << 30)); }();
        return c;
    }
    void* take(size_t bytes, int device) {
        std::lock_guard<std::mutex> lock(mu);
        for (size_t k = 0; k < free_list.size(); ++k)
            if (free_list[k].bytes == bytes && free_list[k].device == device) {
                void* q = free_list[k].p;
                held -= bytes;
                free_list.erase(free_list.begin() + long(k));
                return q;
            }
        return nullptr;
    }
    bool give(void* q, size_t bytes, int device) {
        std::lock_guard<std::mutex> lock(mu);
        if (bytes < kMinBytes || bytes > cap()) return false;
        // full: the oldest entries make room (sizes nobody asks for any more must not block the ones in use)
        while (!free_list.empty() && (held + bytes > cap() || free_list.size() >= kMaxEntries)) {
            cudaFree(free_list.front().p);
            held -= free_list.front().bytes;
            free_list.erase(free_list.begin());
        }
        free_list.push_back(Entry{q, bytes, device});
        held += bytes;
        return true;
    }
    void clear() {
        std::lock_guard<std::mutex> lock(mu);
        for (auto& e : free_list) cudaFree(e.p);
        free_list.clear();
        held = 0;
    }
};
DeviceCache g_cache;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int device = 0;
    void alloc(size_t count) {
        release();
        if (count == 0) return;
        const size_t bytes = count * sizeof(T);
        CUDA_TRY(cudaGetDevice(&device));
        if (bytes >= DeviceCache::kMinBytes) p = static_cast<T*>(g_cache.take(bytes, device));
        if (!p) {
            cudaError_t e = cudaMalloc(&p, bytes);
            if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and retry once
                (void)cudaGetLastError();
                g_cache.clear();
                e = cudaMalloc(&p, bytes);
            }
            CUDA_TRY(e);
        }
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
        if (p && !g_cache.give(p, n * sizeof(T), device)) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), device(o.device) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; device = o.device; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
};

// ---- NCCL, resolved lazily so that single-GPU users need no libnccl ------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    void load() {
        if (lib) return;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) TM_THROW(TM_ERR_UNSUPPORTED, "multi-GPU needs NCCL but libnccl.so.2 could not be loaded: %s", dlerror());
#define TM_SYM(field, sym)                                                                         \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, sym));                                    \
    if (!field) TM_THROW(TM_ERR_UNSUPPORTED, "libnccl lacks symbol %s", sym)
        TM_SYM(GetUniqueId, "ncclGetUniqueId");
        TM_SYM(CommInitRank, "ncclCommInitRank");
        TM_SYM(CommDestroy, "ncclCommDestroy");
        TM_SYM(Send, "ncclSend");
        TM_SYM(Recv, "ncclRecv");
        TM_SYM(AllReduce, "ncclAllReduce");
        TM_SYM(GroupStart, "ncclGroupStart");
        TM_SYM(GroupEnd, "ncclGroupEnd");
        TM_SYM(GetErrorString, "ncclGetErrorString");
#undef TM_SYM
    }
};
NcclApi g_nccl;
#define NCCL_TRY(expr)                                                                                              \
    do {                                                                                                            \
        ncclResult_t _r = (expr);                                                                                   \
        if (_r != ncclSuccess) TM_THROW(TM_ERR_CUDA, "%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
    } while (0)

struct EdgeCache {  // device copies of the four edges + clusterings of one block (TFI inputs)
    DevBuf<double> buf;
    bool valid = false;
};

// One level of the geometric multigrid hierarchy of a single block (level 0 aliases the mesh's own fields).
struct MgLevel {
    int ni = 0, nj = 0, n_tiles = 0;
    MgLevelDims to_coarse{};                 // transfer to the next coarser level
    double scale = 1.0;                      // (r_i r_j)^2: row units of the next coarser level per row unit here
    DevBuf<double2> U, V, rhs, tmp, E;       // iterate ping-pong, tau term, scratch (row / residual), restricted iterate
    DevBuf<Tile> tiles;
    DevBuf<DevBlock> blk;
    DevBuf<SmallNode> small;                 // tiny levels: interior node list of winslow_small_level_kernel
    int n_small = 0;
    // level 1 only: Anderson acceleration history on this level's nodes (see anderson_step)
    std::vector<std::unique_ptr<DevBuf<double2>>> aa_G, aa_F;
    DevBuf<double2> aa_X, aa_D, aa_zero;
    DevBuf<double> aa_part, aa_gram, aa_coef;
    int aa_head = -1, aa_count = 0;
    bool aa_have_x = false;
};

// plan of the persistent Krylov kernel (krylov.inl): components, warp tiles, CTA groups
struct KrylovPlan {
    DevBuf<WTile> wtiles;
    DevBuf<KComp> comps;
    DevBuf<int32_t> group_comps, cta_group;
    DevBuf<KGroup> groups;
    DevBuf<KCtl> ctl;
    DevBuf<KBarrier> bars;
    DevBuf<double> partials;
    unsigned long long epoch = 0;
    std::vector<KComp> h_comps;
    std::vector<KCtl> h_ctl;
    int n_groups = 0, n_ctas = 0, group_ctas = 0;
    bool built_pq = false;
    // two-level preconditioner (krylov_coarse.cuh); n_items == 0: not in use
    DevBuf<KCoarse> coarse;
    DevBuf<int32_t> coarse_ok, agg, contrib_ptr, contrib_src, need_ptr, need, mem_ptr, mem_code, agg_block;
    DevBuf<CoarseItem> items;
    DevBuf<double> G;
    DevBuf<double2> contrib;
    int64_t n_slots = 0, g_size = 0;
    int n_items = 0, nc_max = 0, need_max = 0, src_max = 0, coarse_age = -1, coarse_every = 1;
    size_t smem = 0;
    std::vector<int32_t> h_coarse_ok;
    DevBuf<long long> timing;   // TM_KRYLOV_TIMING
};


// plan of the phased Krylov path (krylov.inl / krylov_phased.cuh): warp tiles and boundary chunks per component
struct PhasedPlan {
    DevBuf<WTile> wtiles;
    DevBuf<int32_t> wt_comp;
    DevBuf<BChunk> chunks;
    DevBuf<KPComp> comps;
    DevBuf<KState> state;
    DevBuf<double> partials;
    DevBuf<int> count;
    int* h_count = nullptr;   // pinned
    std::vector<KPComp> h_comps;
    std::vector<KState> h_state;
    int n_comp = 0, n_wtiles = 0, n_chunks = 0;
    // two-level preconditioner (krylov_coarse.cuh); n_items == 0: not in use
    DevBuf<KCoarse> coarse;
    DevBuf<int32_t> coarse_ok, agg, contrib_ptr, contrib_src, mem_ptr, mem_code, agg_block;
    DevBuf<CoarseItem> items;
    DevBuf<double> G;
    DevBuf<double2> contrib, e_r, e_v, e_t;
    int64_t g_size = 0;
    int n_items = 0, nc_max = 0, coarse_age = -1, coarse_every = 1, sslot0 = 0, jslot0 = 0;
    ~PhasedPlan() { if (h_count) cudaFreeHost(h_count); }
};

// Everything one rank keeps on its GPU.  A distributed mesh holds exactly one; the in-process emulation of several
// ranks on one GPU (tests of the multi-rank logic) holds all of them.
struct RankMesh {
    LocalTables L;
    int64_t N = 0;  // local field length: own + ghosts + synthesised copies
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
    DevBuf<int64_t> d_send_idx, d_check_send_idx;
    DevBuf<double2> sendbuf;
    DevBuf<double> part_int, part_bnd, part_vec, bconst, red;
    DevBuf<unsigned long long> d_worst;
    DevBuf<SolveCtl> d_ctl;
    DevBuf<double2> kr, krhat, kp, kv, ks, kt, kd, kp2, kv2;
    bool krylov_ready = false;
    DevBuf<double2> snapshot;                // tm_mesh_download_block_async
    std::vector<cudaEvent_t> snap_copied;    // per block (global index): its snapshot has reached the host
    std::unique_ptr<KrylovPlan> kplan;       // single-rank meshes: built on the first Picard solve
    std::unique_ptr<PhasedPlan> pplan;
    std::vector<EdgeCache> edges;        // indexed by position in L.own_blocks
    std::vector<uint8_t> have_coords;
    int n_tiles = 0, n_rim_tiles = 0, n_bnd_rows = 0, n_bnd_ctas = 0, vec_grid = 1;
    bool has_pq = false;
    DevBuf<WhiteParams> d_wgroups;           // White groups handled by this rank
    DevBuf<WhiteNode> d_wnodes;
    int n_wnodes = 0;
    std::vector<std::unique_ptr<MgLevel>> mg;  // built on first use of TM_SOLVER_FAS_MULTIGRID (single fixed-boundary block)
    // multi-block multigrid (one RankMesh per rank per level): FAS tau term, residual scratch, restricted iterate
    DevBuf<double2> mg_rhs, mg_tmp, mg_E, mg_zero;
    std::vector<BlockXfer> xfer_blocks;      // own blocks: this level -> next coarser level
    DevBuf<BlockXfer> d_xfer_blocks;
    int xf_ni_f = 0, xf_nj_f = 0, xf_ni_c = 0, xf_nj_c = 0;  // largest extents over the own blocks (grid of the batched transfer kernels)
    bool xf_all_2x2 = false;                                 // every own block halves both directions
    DevBuf<double> soa_stage;                // device staging of the structured (SoA) output
    bool replicated = false;                 // multigrid level held in full by every rank (no halo exchange, redundant work)
    DevBuf<SmallNode> d_small;               // tiny levels: flat interior node list of winslow_small_level_kernel
    int n_small = 0;
    DevBuf<unsigned long long> d_change;     // level 0: max-norm movement of the level-1 nodes between two restrictions
    bool mg_primed = false;                  // coarse levels: both ping-pong buffers hold the (constant) fixed-node values
    // level 1 only: Anderson acceleration history (rings of mg_aa_window samples G_j and residuals F_j, the accelerated state X)
    std::vector<std::unique_ptr<DevBuf<double2>>> aa_G, aa_F;
    DevBuf<double2> aa_X, aa_D;
    int aa_head = -1, aa_count = 0;          // newest slot, entries in the rings
    bool aa_have_x = false;
    DevBuf<double> aa_part, aa_gram, aa_coef;
    double2* E() { return mg_E.p; }
    // NVLink peer-memory halo exchange (real multi-rank meshes): peers' fields mapped with CUDA IPC
    struct P2P {
        bool ready = false;                          // flags + X[0] + X[1] mapped on every rank
        bool have[3] = {false, false, false};        // X[0], X[1], mg_tmp
        double2* peer[3][P2P_MAX_RANKS] = {};        // mapped base pointers of the peers' buffers
        unsigned long long* peer_flags[P2P_MAX_RANKS] = {};
        std::vector<void*> opened;                   // for cudaIpcCloseMemHandle
        DevBuf<unsigned long long> flags;            // slot p: number of pushes rank p has completed into this rank's buffers
        DevBuf<unsigned int> counter;
        DevBuf<int> err;
        unsigned int nb_mask = 0;
        unsigned long long epoch = 0;
    } p2p;
    ~RankMesh() {
        for (void* q : p2p.opened) cudaIpcCloseMemHandle(q);
        for (cudaEvent_t e : snap_copied) if (e) cudaEventDestroy(e);
    }
    RankMesh() = default;
    DevBuf<RestrictRow> d_rrows;             // boundary rows of the next coarser level <- residuals of this level
    int n_rrows = 0;
};

// One level of the multi-block multigrid hierarchy; level 0 aliases the mesh's own topology and rank data.
struct MgbLevel {
    Topology topo;                                   // levels >= 1
    std::vector<tm_block> blocks;
    std::vector<tm_connection> conns;
    std::vector<tm_condition> bcs;
    std::vector<std::unique_ptr<RankMesh>> ranks;    // levels >= 1
    std::vector<int> fi, fj;                         // per block: coarsening factors towards the next level (empty on the coarsest)
    std::vector<double> tan_i, tan_j;                // per block: weight of the tangential term next to sliding sides (1 on level 0)
    double work = 1.0;                               // nodes relative to level 0
};

// the pinned scalar block a mesh polls (cudaMallocHost / cudaFreeHost cost milliseconds: recycled like the device fields)
std::mutex g_pinned_mu;
std::vector<SolveCtl*> g_pinned_free;
SolveCtl* acquire_pinned_ctl() {
    {
        std::lock_guard<std::mutex> lock(g_pinned_mu);
        if (!g_pinned_free.empty()) { SolveCtl* q = g_pinned_free.back(); g_pinned_free.pop_back(); return q; }
    }
    SolveCtl* q = nullptr;
    CUDA_TRY(cudaMallocHost(&q, sizeof(SolveCtl)));
    return q;
}
void release_pinned_ctl(SolveCtl* q) {
    std::lock_guard<std::mutex> lock(g_pinned_mu);
    if (g_pinned_free.size() < 8) g_pinned_free.push_back(q); else cudaFreeHost(q);
}

}  // namespace

struct tm_mesh {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    Topology topo;
    std::vector<int32_t> owner;
    int n_ranks = 1;
    bool emulated = false;                       // all ranks live in this process on one GPU (tests)
    std::vector<std::unique_ptr<RankMesh>> ranks;  // the ranks held by this process
    ncclComm_t comm = nullptr;
    SolveCtl* h_ctl = nullptr;                   // pinned
    bool begun = false;
    int cf = TM_CF_LAPLACE;
    uint64_t outer_done = 0;  // outer iterations since begin_smoothing (the `n` of system.fill(n), smooth.zig:1107-1110)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t comm_stream = nullptr;          // the halo exchange of a sweep runs here, next to the bulk of the interior
    cudaStream_t copy_stream = nullptr;          // tm_mesh_download_block_async: snapshots go to the host here while the mesh is used again
    cudaEvent_t ev_snap = nullptr, ev_copied = nullptr;
    bool copies_pending = false;
    cudaEvent_t ev_rim = nullptr, ev_x = nullptr;
    bool use_graph = true;                       // TM_GRAPH=0: launch the BiCGStab iteration kernel by kernel
    bool overlap = true;                         // TM_OVERLAP=0: exchange on the main stream after the whole sweep
    bool ev_x_valid = false, overlap_pending = false;
    std::vector<tm_block> h_blocks;              // host copies of the topology description (xy = NULL): multigrid coarsening
    std::vector<tm_connection> h_conns;
    std::vector<tm_condition> h_bcs;
    std::vector<std::unique_ptr<MgbLevel>> mgb;  // multi-block multigrid hierarchy, built on first use
    int sm_count = 148;
    int64_t mg_replicate_nodes = 32768;  // multigrid levels up to this size are replicated on every rank (TM_MG_REPLICATE_NODES)
    int64_t mg_small_nodes = 4096;       // ... and up to this size swept by the single-CTA kernel (TM_MG_SMALL_NODES)
    int mg_aa_window = 3;       // residuals kept by the Anderson acceleration (TM_MG_AA_WINDOW, 1..AA_MAX)
    bool mg_aa = true;        // Anderson acceleration of the multi-block multigrid cycle (TM_MG_AA=0 switches it off)
    int tile_rows = TILE_I;   // TM_TILE_ROWS overrides (tuning aid)
    bool use_bulk = true;     // TM_INTERIOR=regs selects the register-only interior kernel (tuning aid)

    ~tm_mesh() {
        mgb.clear();
        ranks.clear();
        if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
        if (h_ctl) release_pinned_ctl(h_ctl);
        if (ev_rim) cudaEventDestroy(ev_rim);
        if (ev_x) cudaEventDestroy(ev_x);
        if (comm_stream) cudaStreamDestroy(comm_stream);
        if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
        if (ev_snap) cudaEventDestroy(ev_snap);
        if (ev_copied) cudaEventDestroy(ev_copied);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

namespace {

int bnd_ctas(int rows) { return (rows + BND_THREADS - 1) / BND_THREADS; }

void build_rank(tm_mesh* m, const Topology& topo, RankMesh& r, int rank, const std::vector<int32_t>* owner_override = nullptr) {
    cudaStream_t s = m->stream;
    r.L = localize(topo, owner_override ? *owner_override : m->owner, rank, m->n_ranks);
    r.N = r.L.n_local;
    {   // boundary rows and rhs terms grouped by component (independent systems): the persistent Krylov kernel walks them per component
        auto comp_of_local = [&](int64_t l) {
            size_t lo = 0, hi = r.L.own_blocks.size() - 1;  // own blocks ascend in the local field
            while (lo < hi) {
                const size_t mid = (lo + hi + 1) / 2;
                if (r.L.loff[size_t(r.L.own_blocks[mid])] <= l) lo = mid; else hi = mid - 1;
            }
            return topo.comp_of_block[size_t(r.L.own_blocks[lo])];
        };
        auto group_by_component = [&](auto& rows, auto node_of) {
            std::vector<int32_t> key(rows.size());
            bool sorted = true;
            for (size_t k = 0; k < rows.size(); ++k) {
                key[k] = comp_of_local(node_of(rows[k]));
                sorted = sorted && (k == 0 || key[k - 1] <= key[k]);
            }
            if (sorted) return;  // the usual case: a batch is concatenated cut by cut
            std::vector<size_t> perm(rows.size());
            for (size_t k = 0; k < perm.size(); ++k) perm[k] = k;
            std::stable_sort(perm.begin(), perm.end(), [&](size_t x, size_t y) { return key[x] < key[y]; });
            auto copy = rows;
            for (size_t k = 0; k < perm.size(); ++k) rows[k] = copy[perm[k]];
        };
        if (!r.L.own_blocks.empty() && topo.n_comp > 1) {
            group_by_component(r.L.smoothed, [](const SmoothedRow& x) { return x.g0; });
            group_by_component(r.L.junction_rows, [](const JunctionRow& x) { return x.self; });
            group_by_component(r.L.sliding, [](const SlidingRow& x) { return x.self; });
            group_by_component(r.L.rhs_terms, [](const RhsTerm& x) { return x.g; });
        }
    }
    r.X[0].alloc(size_t(std::max<int64_t>(r.N, 1)));
    r.X[1].alloc(size_t(std::max<int64_t>(r.N, 1)));
    r.X[0].zero(s);
    r.X[1].zero(s);
    std::vector<Tile> tiles;
    std::vector<DevBlock> blocks(topo.blocks.size(), DevBlock{0, 0, 0});
    for (size_t k = 0; k < r.L.own_blocks.size(); ++k) {
        const size_t b = size_t(r.L.own_blocks[k]);
        const auto& B = topo.blocks[b];
        blocks[b] = DevBlock{r.L.loff[b], int32_t(B.ni), int32_t(B.nj)};
        // rows per CTA: about tile_rows, evened out over the block so that no CTA gets a short remainder
        const int64_t interior_i = B.ni - 2;
        const int64_t n_i = std::max<int64_t>(1, (interior_i + m->tile_rows - 1) / m->tile_rows);
        const int64_t rows = (interior_i + n_i - 1) / n_i;
        for (int64_t i0 = 1; i0 <= B.ni - 2; i0 += rows)
            for (int64_t j0 = 1; j0 <= B.nj - 2; j0 += TILE_J) tiles.push_back(Tile{int32_t(b), int32_t(i0), int32_t(j0), int32_t(rows)});
    }
    // Rim tiles first: the tiles that write a node some peer ghosts, or read a copy whose root is a ghost.  Everything the
    // halo exchange of a sweep needs is produced by the rim tiles (and the boundary CTAs), so the exchange can travel while
    // the bulk of the interior is still being swept (relax_sweep).
    r.n_rim_tiles = 0;
    if (m->n_ranks > 1 && !tiles.empty()) {
        std::vector<uint8_t> rim(tiles.size(), 0);
        std::vector<size_t> first_tile(topo.blocks.size(), 0);  // index of a block's first tile
        {
            size_t t = 0;
            for (int32_t b : r.L.own_blocks) {
                first_tile[size_t(b)] = t;
                while (t < tiles.size() && tiles[t].block == b) ++t;
            }
        }
        auto mark_near = [&](int64_t lidx) {  // tiles holding an interior node within one node of local node lidx
            if (lidx < 0 || lidx >= r.L.n_own) return;
            size_t b = size_t(r.L.own_blocks.back());
            for (int32_t cand : r.L.own_blocks)
                if (lidx >= r.L.loff[size_t(cand)] && lidx < r.L.loff[size_t(cand)] + topo.blocks[size_t(cand)].ni * topo.blocks[size_t(cand)].nj) { b = size_t(cand); break; }
            const int64_t ni = topo.blocks[b].ni, nj = topo.blocks[b].nj, local = lidx - r.L.loff[b], i = local / nj, j = local - i * nj;
            const int64_t rows = tiles[first_tile[b]].rows, ncols = (nj - 2 + TILE_J - 1) / TILE_J;
            for (int64_t di = -1; di <= 1; ++di)
                for (int64_t dj = -1; dj <= 1; ++dj) {
                    const int64_t ii = std::min<int64_t>(std::max<int64_t>(i + di, 1), ni - 2), jj = std::min<int64_t>(std::max<int64_t>(j + dj, 1), nj - 2);
                    rim[first_tile[b] + size_t(((ii - 1) / rows) * ncols + (jj - 1) / TILE_J)] = 1;
                }
        };
        for (int64_t l : r.L.send_lidx) mark_near(l);
        for (size_t k = size_t(r.L.n_slaves_local_root); k < r.L.slaves.size(); ++k) mark_near(r.L.slaves[k].self);
        std::vector<Tile> ordered;
        for (size_t t = 0; t < tiles.size(); ++t) if (rim[t]) ordered.push_back(tiles[t]);
        r.n_rim_tiles = int(ordered.size());
        for (size_t t = 0; t < tiles.size(); ++t) if (!rim[t]) ordered.push_back(tiles[t]);
        tiles.swap(ordered);
    }
    r.n_tiles = int(tiles.size());
    r.d_tiles.upload(tiles, s);
    r.d_blocks.upload(blocks, s);
    r.d_srows.upload(r.L.smoothed, s);
    r.d_jrows.upload(r.L.junction_rows, s);
    r.d_lrows.upload(r.L.sliding, s);
    r.d_slaves.upload(r.L.slaves, s);
    r.d_cslaves.upload(r.L.const_slaves, s);
    r.d_fo.upload(r.L.fixed_overrides, s);
    r.d_pairs.upload(r.L.pairs, s);
    r.d_rhs_terms.upload(r.L.rhs_terms, s);
    r.d_send_idx.upload(r.L.send_lidx, s);
    r.d_check_send_idx.upload(r.L.check_send_lidx, s);
    r.sendbuf.alloc(std::max(r.L.send_lidx.size(), r.L.check_send_lidx.size()));
    r.n_bnd_rows = int(r.L.smoothed.size() + r.L.junction_rows.size() + r.L.sliding.size());
    r.n_bnd_ctas = bnd_ctas(r.n_bnd_rows);
    int sms = 148;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device));
    const int64_t want = (r.N + VEC_THREADS - 1) / VEC_THREADS;
    r.vec_grid = int(std::max<int64_t>(1, std::min<int64_t>(want, int64_t(sms) * 8)));
    r.part_int.alloc(size_t(std::max(r.n_tiles, 1)) * 5);
    r.part_bnd.alloc(size_t(std::max(r.n_bnd_ctas, 1)) * 5);
    r.part_vec.alloc(size_t(r.vec_grid) * 5);
    r.part_int.zero(s); r.part_bnd.zero(s); r.part_vec.zero(s);
    r.bconst.alloc(2); r.bconst.zero(s);
    r.red.alloc(5); r.red.zero(s);
    r.d_worst.alloc(1);
    r.d_ctl.alloc(1); r.d_ctl.zero(s);
    r.edges.resize(r.L.own_blocks.size());
    r.have_coords.assign(r.L.own_blocks.size(), 0);
}

#include "exchange.inl"  // p2p_setup / p2p_add_tmp / p2p_check: CUDA IPC plumbing of the peer-memory halo exchange

bool krylov_persistent_possible(const tm_mesh* m);
void ensure_krylov(tm_mesh* m) {
    for (auto& rp : m->ranks) {
        RankMesh& r = *rp;
        if (r.krylov_ready) continue;
        for (DevBuf<double2>* v : {&r.kr, &r.krhat, &r.kp, &r.kv, &r.ks, &r.kt, &r.kd, &r.kp2, &r.kv2}) {
            if ((v == &r.kp2 || v == &r.kv2) && !krylov_persistent_possible(m)) continue;  // ping-pong partners of the persistent kernel
            v->alloc(size_t(std::max<int64_t>(r.N, 1)));
            v->zero(m->stream);
        }
        r.krylov_ready = true;
    }
}

// ---- halo exchange: every rank's ghost slots of `field` are refreshed from their owners ------------------------
// check = true: the one-time exchange of raw side-0 coordinates for connectionDataCheck (separate slots).
using RankList = std::vector<std::unique_ptr<RankMesh>>;
void sync_slaves(tm_mesh* m, RankMesh& r, double2* v, int mode, bool only_remote_root = false, cudaStream_t xs = nullptr);
// slave_mode >= 0: afterwards the copies whose root is a ghost are re-derived from it (mode as in sync_slaves_kernel)
template <class Get>
void exchange_on(tm_mesh* m, RankList& ranks, Get get, bool check = false, int slave_mode = -1, cudaStream_t xs = nullptr) {
    if (m->n_ranks == 1 || ranks.empty() || ranks[0]->replicated) return;  // (a replicated level has no remote roots either)
    cudaStream_t s = xs ? xs : m->stream;
    auto send_base = [&](RankMesh& r) -> const std::vector<int64_t>& { return check ? r.L.check_send_base : r.L.send_base; };
    auto ghost_base = [&](RankMesh& r) -> const std::vector<int64_t>& { return check ? r.L.check_ghost_base : r.L.ghost_base; };
    auto region = [&](RankMesh& r) { return check ? r.L.n_own + r.L.n_ghost + r.L.n_synth : r.L.n_own; };
    bool p2p = false;
    int which = -1;
    if (!m->emulated && !check) {
        RankMesh& r = *ranks[0];
        double2* f = get(r);
        which = f == r.X[0].p ? 0 : f == r.X[1].p ? 1 : (r.mg_tmp.p && f == r.mg_tmp.p) ? 2 : -1;
        p2p = r.p2p.ready && which >= 0 && r.p2p.have[which];
    }
    for (auto& rp : ranks) {
        if (p2p) break;
        RankMesh& r = *rp;
        const int64_t n = send_base(r).back();
        const int64_t* idx = check ? r.d_check_send_idx.p : r.d_send_idx.p;
        if (n > 0) LAUNCH(pack_kernel, unsigned((n + 255) / 256), 256, s, idx, n, (const double2*)get(r), r.sendbuf.p);
    }
    if (m->emulated) {
        for (auto& rp : ranks) {
            RankMesh& r = *rp;
            for (int p = 0; p < m->n_ranks; ++p) {
                const int64_t cnt = send_base(r)[size_t(p) + 1] - send_base(r)[size_t(p)];
                if (cnt == 0) continue;
                RankMesh& d = *ranks[size_t(p)];
                CUDA_TRY(cudaMemcpyAsync(get(d) + region(d) + ghost_base(d)[size_t(r.L.rank)], r.sendbuf.p + send_base(r)[size_t(p)],
                                         size_t(cnt) * sizeof(double2), cudaMemcpyDeviceToDevice, s));
            }
        }
    } else {
        RankMesh& r = *ranks[0];
        if (p2p) {
            // peer-memory push: gather + remote store + signal in one launch, then wait for the neighbours' pushes
            PushArgs a{};
            a.n_ranks = m->n_ranks;
            for (int p = 0; p <= m->n_ranks; ++p) a.base[p] = r.L.send_base[size_t(p)];
            for (int p = 0; p < m->n_ranks; ++p) {
                a.dst[p] = r.p2p.peer[which][p] ? r.p2p.peer[which][p] + r.L.peer_ghost_offset[size_t(p)] : nullptr;
                a.flag[p] = ((r.p2p.nb_mask >> p) & 1u) ? r.p2p.peer_flags[p] + r.L.rank : nullptr;
            }
            r.p2p.epoch += 1;
            const int64_t n = r.L.send_base.back();
            const int64_t first = r.L.n_slaves_local_root;
            const int n_sl = slave_mode >= 0 ? int(int64_t(r.L.slaves.size()) - first) : 0;
            const int64_t want = (std::max<int64_t>(n, n_sl) + 255) / 256;
            const int64_t most = xs ? 32 : m->sm_count;  // next to a running sweep: few CTAs, they only move O(interface) nodes
            LAUNCH(p2p_exchange_kernel, unsigned(std::max<int64_t>(1, std::min<int64_t>(want, most))), 256, s, (const int64_t*)r.d_send_idx.p, n, get(r), a, r.p2p.epoch,
                   r.p2p.counter.p, (const unsigned long long*)r.p2p.flags.p, r.p2p.nb_mask, r.p2p.err.p, (const SlaveRow*)(r.d_slaves.p + first), n_sl,
                   std::max(slave_mode, 0));
            return;
        }
        NCCL_TRY(g_nccl.GroupStart());
        for (int p = 0; p < m->n_ranks; ++p) {
            const int64_t ns = send_base(r)[size_t(p) + 1] - send_base(r)[size_t(p)];
            const int64_t ng = ghost_base(r)[size_t(p) + 1] - ghost_base(r)[size_t(p)];
            if (ns > 0) NCCL_TRY(g_nccl.Send(r.sendbuf.p + send_base(r)[size_t(p)], size_t(ns) * 2, ncclDouble, p, m->comm, s));
            if (ng > 0) NCCL_TRY(g_nccl.Recv(get(r) + region(r) + ghost_base(r)[size_t(p)], size_t(ng) * 2, ncclDouble, p, m->comm, s));
        }
        NCCL_TRY(g_nccl.GroupEnd());
    }
    if (slave_mode >= 0)
        for (auto& rp : ranks) sync_slaves(m, *rp, get(*rp), slave_mode, true, s);
}

template <class Get>
void exchange(tm_mesh* m, Get get, bool check = false, int slave_mode = -1) { exchange_on(m, m->ranks, get, check, slave_mode); }

// copies of nodes: mode 0 homogeneous / 1 affine / 2 zero; `only_remote_root` restricts to copies whose root is a ghost
void sync_slaves(tm_mesh* m, RankMesh& r, double2* v, int mode, bool only_remote_root, cudaStream_t xs) {
    const int64_t first = only_remote_root ? r.L.n_slaves_local_root : 0;
    const int n = int(int64_t(r.L.slaves.size()) - first);
    if (n > 0) LAUNCH(sync_slaves_kernel, (n + 127) / 128, 128, xs ? xs : m->stream, (const SlaveRow*)(r.d_slaves.p + first), n, v, mode);
}

// which rows of a rank one launch covers (bulk kernel only): tiles [first, first+count) and, optionally, the boundary CTAs
struct RowPart {
    int first = 0, count = -1;  // count < 0: all tiles
    bool bnd = true;
    cudaStream_t stream = nullptr;  // NULL: the mesh's stream
};

// ---- kernel dispatch over the (LAGGED, HAS_PQ) template space -------------------------------------
template <int MODE, int STATS>
void launch_rows(tm_mesh* m, RankMesh& r, bool lagged, const double2* u, const double2* xc, double2* out, double omega, const double2* dot_a,
                 RowPart part = RowPart()) {
    const bool has_pq = r.has_pq;
    const double2* pq = r.pq.p;
    cudaStream_t s = part.stream ? part.stream : m->stream;
    const int t_first = part.first, t_count = part.count < 0 ? r.n_tiles : part.count, b_ctas = part.bnd ? r.n_bnd_ctas : 0;
#define TM_ROWS(LAG, PQ)                                                                                                                          \
    do {                                                                                                                                          \
        const BndArgs bnd{r.d_srows.p, r.d_jrows.p, r.d_lrows.p, r.d_slaves.p, r.part_bnd.p, int(r.L.smoothed.size()),                            \
                          int(r.L.junction_rows.size()), int(r.L.sliding.size()), r.n_bnd_ctas};                                                  \
        if (r.n_tiles + r.n_bnd_ctas > 0)                                                                                                         \
            LAUNCH((winslow_interior_kernel<MODE, LAG, PQ, STATS>), r.n_tiles + r.n_bnd_ctas, TILE_J, s, (const Tile*)r.d_tiles.p,                \
                   (const DevBlock*)r.d_blocks.p, u, xc, pq, out, omega, dot_a, r.part_int.p, bnd);                                               \
    } while (0)
#define TM_ROWS_BULK(PQ)                                                                                                                          \
    do {                                                                                                                                          \
        const BndArgs bnd{r.d_srows.p, r.d_jrows.p, r.d_lrows.p, r.d_slaves.p, r.part_bnd.p, int(r.L.smoothed.size()),                            \
                          int(r.L.junction_rows.size()), int(r.L.sliding.size()), b_ctas};                                                        \
        if (t_count + b_ctas > 0)                                                                                                                 \
            LAUNCH((winslow_interior_bulk_kernel<MODE, PQ, STATS>), t_count + b_ctas, TILE_J, s, (const Tile*)r.d_tiles.p + t_first,              \
                   (const DevBlock*)r.d_blocks.p, u, pq, out, omega, dot_a, r.part_int.p + size_t(t_first) * 5, bnd, (const double2*)nullptr);    \
    } while (0)
    if (lagged) { if (has_pq) TM_ROWS(true, true); else TM_ROWS(true, false); }
    else if (m->use_bulk) { if (has_pq) TM_ROWS_BULK(true); else TM_ROWS_BULK(false); }
    else { if (has_pq) TM_ROWS(false, true); else TM_ROWS(false, false); }
#undef TM_ROWS
#undef TM_ROWS_BULK
}

// multigrid levels: all rows of the rank (interior tiles + boundary CTAs) through the bulk kernel, Laplace control
// function, optional FAS right-hand side
template <int MODE, int STATS>
void launch_rows_mg(tm_mesh* m, RankMesh& r, const double2* u, double2* out, double omega, const double2* rhs, RowPart part = RowPart()) {
    const int t_first = part.first, t_count = part.count < 0 ? r.n_tiles : part.count, b_ctas = part.bnd ? r.n_bnd_ctas : 0;
    const BndArgs bnd{r.d_srows.p, r.d_jrows.p, r.d_lrows.p, r.d_slaves.p, r.part_bnd.p, int(r.L.smoothed.size()),
                      int(r.L.junction_rows.size()), int(r.L.sliding.size()), b_ctas};
    if (t_count + b_ctas == 0) return;
    if (rhs)
        LAUNCH((winslow_interior_bulk_kernel<MODE, false, STATS, true>), t_count + b_ctas, TILE_J, part.stream ? part.stream : m->stream, (const Tile*)r.d_tiles.p + t_first,
               (const DevBlock*)r.d_blocks.p, u, (const double2*)nullptr, out, omega, (const double2*)nullptr, r.part_int.p + size_t(t_first) * 5, bnd, rhs);
    else
        LAUNCH((winslow_interior_bulk_kernel<MODE, false, STATS, false>), t_count + b_ctas, TILE_J, part.stream ? part.stream : m->stream, (const Tile*)r.d_tiles.p + t_first,
               (const DevBlock*)r.d_blocks.p, u, (const double2*)nullptr, out, omega, (const double2*)nullptr, r.part_int.p + size_t(t_first) * 5, bnd,
               (const double2*)nullptr);
}

// One damped-Jacobi sweep of every rank held by this process (X[cur] -> X[1-cur], swap) followed by the halo exchange of
// the new iterate.  A real multi-GPU rank overlaps the two: the rim tiles and the boundary rows run on a second,
// high-priority stream, immediately followed there by the exchange (push over NVLink / NCCL, wait, copies of ghost
// roots), while the bulk of the interior is swept on the main stream at the same time:
//     comm:  wait(bulk k-1)  rim(k)  exchange(k)                    main:  wait(exchange k-1)  bulk(k)
// rim(k) may overwrite what bulk(k-1) still reads and bulk(k) what rim(k-1) read -- the two waits order exactly that.
// `mg`: Laplace rows with optional FAS rhs.
template <int STATS>
void relax_sweep(tm_mesh* m, RankList& R, double omega, bool mg, bool with_rhs) {
    auto xcur = [](RankMesh& r) { return r.X[r.cur].p; };
    auto rows = [&](RankMesh& r, const double2* u, double2* out, RowPart part) {
        if (mg) launch_rows_mg<MODE_RELAX, STATS>(m, r, u, out, omega, with_rhs ? (const double2*)r.mg_rhs.p : nullptr, part);
        else launch_rows<MODE_RELAX, STATS>(m, r, false, u, u, out, omega, nullptr, part);
    };
    const bool overlap = m->overlap && m->n_ranks > 1 && !m->emulated && R.size() == 1 && !R[0]->replicated && (mg || m->use_bulk) && R[0]->n_rim_tiles * 4 < R[0]->n_tiles;
    if (!overlap) {
        for (auto& rp : R) {
            RankMesh& r = *rp;
            rows(r, r.X[r.cur].p, r.X[1 - r.cur].p, RowPart());
            r.cur = 1 - r.cur;
        }
        exchange_on(m, R, xcur, false, 1);
        return;
    }
    RankMesh& r = *R[0];
    const double2* u = r.X[r.cur].p;
    double2* out = r.X[1 - r.cur].p;
    // everything queued on the main stream so far (the previous sweep's bulk in particular) precedes the rim
    CUDA_TRY(cudaEventRecord(m->ev_rim, m->stream));
    CUDA_TRY(cudaStreamWaitEvent(m->comm_stream, m->ev_rim, 0));
    rows(r, u, out, RowPart{0, r.n_rim_tiles, true, m->comm_stream});
    r.cur = 1 - r.cur;
    // the previous exchange (queued on comm before the rim) precedes the bulk: ev_x still holds its record
    if (m->ev_x_valid) CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_x, 0));
    rows(r, u, out, RowPart{r.n_rim_tiles, r.n_tiles - r.n_rim_tiles, false, nullptr});
    exchange_on(m, R, xcur, false, 1, m->comm_stream);
    CUDA_TRY(cudaEventRecord(m->ev_x, m->comm_stream));
    m->ev_x_valid = true;
    m->overlap_pending = true;
}
// after a run of overlapped sweeps: the main stream joins the last exchange
void relax_join(tm_mesh* m) {
    if (!m->overlap_pending) return;
    CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_x, 0));
    m->overlap_pending = false;
}

// rank-local reduction of the per-CTA partials, all-reduce over ranks, solver scalars
void launch_reduce(tm_mesh* m, int op, const tm_smooth_options* o, bool from_rows) {
    const int max_it = o->max_inner_iterations > 0x7fffffffull ? 0x7fffffff : int(o->max_inner_iterations);
    const int single = m->n_ranks == 1 ? 1 : 0;
    cudaStream_t s = m->stream;
    for (auto& rp : m->ranks) {
        RankMesh& r = *rp;
        if (from_rows)
            LAUNCH((reduce_kernel<256>), 1, 256, s, (const double*)r.part_int.p, r.n_tiles, (const double*)r.part_bnd.p, r.n_bnd_ctas, r.red.p,
                   (const double*)r.bconst.p, single, op, r.d_ctl.p, o->rtol, o->atol, max_it);
        else
            LAUNCH((reduce_kernel<256>), 1, 256, s, (const double*)r.part_vec.p, r.vec_grid, (const double*)nullptr, 0, r.red.p, (const double*)r.bconst.p,
                   single, op, r.d_ctl.p, o->rtol, o->atol, max_it);
    }
    if (single) return;
    if (m->emulated) {
        RedPtrs ptrs{};
        for (size_t k = 0; k < m->ranks.size(); ++k) ptrs.p[k] = m->ranks[k]->red.p;
        LAUNCH(combine_red_kernel, 1, 32, s, ptrs, int(m->ranks.size()));
    } else {
        RankMesh& r = *m->ranks[0];
        NCCL_TRY(g_nccl.AllReduce(r.red.p, r.red.p, 4, ncclDouble, ncclSum, m->comm, s));
        NCCL_TRY(g_nccl.AllReduce(r.red.p + 4, r.red.p + 4, 1, ncclDouble, ncclMax, m->comm, s));
    }
    for (auto& rp : m->ranks) LAUNCH(finalize_kernel, 1, 32, s, (const double*)rp->red.p, op, rp->d_ctl.p, o->rtol, o->atol, max_it);
}

void fetch_ctl(tm_mesh* m) {  // the solver scalars are identical on all ranks after the all-reduce
    CUDA_TRY(cudaMemcpyAsync(m->h_ctl, m->ranks[0]->d_ctl.p, sizeof(SolveCtl), cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
}

void white_step(tm_mesh* m, bool update, double ds_target, double theta_target) {
    for (auto& rp : m->ranks) {
        RankMesh& r = *rp;
        if (!r.has_pq || r.n_wnodes == 0) continue;
        LAUNCH(white_wall_kernel, (r.n_wnodes + 127) / 128, 128, m->stream, (const WhiteParams*)r.d_wgroups.p, (const WhiteNode*)r.d_wnodes.p, r.n_wnodes,
               ds_target, theta_target, (const double2*)r.X[r.cur].p, r.wall_pq.p, update ? 1 : 0);
        LAUNCH(white_blend_kernel, r.n_wnodes, 64, m->stream, (const WhiteParams*)r.d_wgroups.p, (const WhiteNode*)r.d_wnodes.p, r.n_wnodes,
               (const double2*)r.wall_pq.p, r.pq.p);
    }
}

void validate_options(const tm_smooth_options* o) {
    if (!o) TM_THROW(TM_ERR_INVALID_ARGUMENT, "options are NULL");
    if (o->struct_size != sizeof(tm_smooth_options)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "tm_smooth_options.struct_size mismatch (ABI version?)");
    if (o->solver > TM_SOLVER_FAS_MULTIGRID) TM_THROW(TM_ERR_INVALID_ARGUMENT, "unknown solver %u", o->solver);
    if (o->control_function > TM_CF_WHITE) TM_THROW(TM_ERR_INVALID_ARGUMENT, "unknown control function %u", o->control_function);
    if (!(o->omega > 0.0 && o->omega <= 1.0)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "omega must be in (0, 1]");
    if (o->solver != TM_SOLVER_PICARD_BICGSTAB && o->sweeps_per_iteration == 0) TM_THROW(TM_ERR_INVALID_ARGUMENT, "sweeps_per_iteration must be > 0");
    if (!(o->rtol >= 0.0) || !(o->atol >= 0.0)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "tolerances must be non-negative");
}

// ---- the two ways of advancing one outer iteration -------------------------------------------------
void run_relax(tm_mesh* m, const tm_smooth_options* o, tm_smooth_stats* st) {
    for (uint64_t it = 0; it < o->iterations; ++it) {
        if (m->cf == TM_CF_WHITE && m->outer_done > 0) white_step(m, true, o->white_ds_target, o->white_theta_target);
        for (uint64_t sw = 0; sw < o->sweeps_per_iteration; ++sw) {
            const bool last = sw + 1 == o->sweeps_per_iteration;
            if (last) relax_sweep<1>(m, m->ranks, o->omega, false, false);
            else relax_sweep<0>(m, m->ranks, o->omega, false, false);
            if (last) relax_join(m);
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

// v = D^-1 A p on every rank: refresh ghosts of p, make its copies consistent, apply the rows
template <int STATS, class GetIn, class GetOut, class GetDot>
void apply_operator(tm_mesh* m, GetIn in, GetOut out, GetDot dot) {
    exchange(m, in);
    for (auto& rp : m->ranks) {
        RankMesh& r = *rp;
        sync_slaves(m, r, in(r), 0);
        launch_rows<MODE_APPLY, STATS>(m, r, true, in(r), r.X[r.cur].p, out(r), 1.0, dot(r));
    }
}

// BiCGStab iterations (BiCGStab.zig:303-366) on the row-scaled system until both components report done.
// one BiCGStab iteration (BiCGStab.zig:303-366): ~12 small launches, all scalars on the device
void bicgstab_iteration(tm_mesh* m, const tm_smooth_options* o) {
    cudaStream_t s = m->stream;
    {
        for (auto& rp : m->ranks) {
            RankMesh& r = *rp;
            LAUNCH(bicg_p_kernel, r.vec_grid, VEC_THREADS, s, r.L.n_own, (const SolveCtl*)r.d_ctl.p, (const double2*)r.kr.p, r.kp.p, (const double2*)r.kv.p);
        }
        apply_operator<2>(m, [](RankMesh& r) { return r.kp.p; }, [](RankMesh& r) { return r.kv.p; }, [](RankMesh& r) { return (const double2*)r.krhat.p; });  // v = A p, rhat.v
        launch_reduce(m, RED_ALPHA, o, true);
        for (auto& rp : m->ranks) {
            RankMesh& r = *rp;
            LAUNCH(bicg_s_kernel, r.vec_grid, VEC_THREADS, s, r.L.n_own, (const SolveCtl*)r.d_ctl.p, (const double2*)r.kr.p, (const double2*)r.kv.p, r.ks.p,
                   r.kd.p, (const double2*)r.kp.p, r.part_vec.p);
        }
        launch_reduce(m, RED_NORM_S, o, false);
        apply_operator<3>(m, [](RankMesh& r) { return r.ks.p; }, [](RankMesh& r) { return r.kt.p; }, [](RankMesh& r) { return (const double2*)r.ks.p; });  // t = A s, t.s, t.t
        for (auto& rp : m->ranks) sync_slaves(m, *rp, rp->ks.p, 2);
        launch_reduce(m, RED_OMEGA, o, true);
        for (auto& rp : m->ranks) {
            RankMesh& r = *rp;
            LAUNCH(bicg_r_kernel, r.vec_grid, VEC_THREADS, s, r.L.n_own, (const SolveCtl*)r.d_ctl.p, (const double2*)r.ks.p, (const double2*)r.kt.p, r.kr.p,
                   r.kd.p, (const double2*)r.krhat.p, r.part_vec.p);
        }
        launch_reduce(m, RED_NORM_R, o, false);
    }
}

// Iterates until both components report done.  On one GPU the iteration is captured once into a CUDA graph and replayed:
// on the reference's own mesh sizes (1e4..1e5 nodes) the solver is bound by launch cadence, not by bandwidth.
void bicgstab_cycle(tm_mesh* m, const tm_smooth_options* o, tm_smooth_stats* st) {
    cudaStream_t s = m->stream;
    const int check_every = 8;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches_per_iteration = 0;
    if (m->n_ranks == 1 && m->use_graph) {
        const uint64_t before = g_launches.load();
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        try {
            bicgstab_iteration(m, o);
        } catch (...) {
            cudaStreamEndCapture(s, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        CUDA_TRY(cudaStreamEndCapture(s, &graph));
        launches_per_iteration = g_launches.load() - before;
        g_launches.store(before);  // captured, not run
        const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        CUDA_TRY(e);
    }
    try {
        for (uint64_t k = 0;; ++k) {
            if (k % check_every == 0) {
                fetch_ctl(m);
                if (m->h_ctl->done[0] && m->h_ctl->done[1]) break;
            }
            if (exec) {
                CUDA_TRY(cudaGraphLaunch(exec, s));
                g_launches.fetch_add(launches_per_iteration, std::memory_order_relaxed);
            } else {
                bicgstab_iteration(m, o);
            }
            st->operator_applications += 2;
        }
    } catch (...) {
        if (exec) cudaGraphExecDestroy(exec);
        throw;
    }
    if (exec) cudaGraphExecDestroy(exec);
}

#include "krylov.inl"  // krylov_solve_persistent: all inner solves of an outer iteration in one cooperative launch

// One outer (Picard) iteration = the reference's fill + solve(x) + solve(y) (smooth.zig:104-154), with the two
// solves advanced in lock-step by a matrix-free BiCGStab on the row-scaled system.  One extension over BiCGStab.zig
// that only matters when the tolerance is tighter than the reference's: whenever the solver reports convergence or
// breaks down, the recursive residual is replaced by the true one and, if that is still above the tolerance, the
// iteration restarts from there (at most `max_restarts` times).
void run_picard_bicgstab(tm_mesh* m, const tm_smooth_options* o, tm_smooth_stats* st) {
    ensure_krylov(m);
    cudaStream_t s = m->stream;
    const int max_restarts = 60;
    st->converged = 1;
    auto xnew = [](RankMesh& r) { return r.X[1 - r.cur].p; };
    auto refresh_x = [&]() {  // ghosts and copies of the iterate
        exchange(m, xnew);
        for (auto& rp : m->ranks) sync_slaves(m, *rp, xnew(*rp), 1);
    };
    for (uint64_t it = 0; it < o->iterations; ++it) {
        if (m->cf == TM_CF_WHITE && m->outer_done > 0) white_step(m, true, o->white_ds_target, o->white_theta_target);
        // X[cur] = lagged coordinates (the mesh before this iteration); X[1-cur] = x_new / y_new, warm-started from the
        // mesh (GMRES.zig:157-174)
        for (auto& rp : m->ranks) {
            RankMesh& r = *rp;
            CUDA_TRY(cudaMemcpyAsync(r.X[1 - r.cur].p, r.X[r.cur].p, size_t(r.N) * sizeof(double2), cudaMemcpyDeviceToDevice, s));
        }
        const bool persistent = krylov_persistent_possible(m);
        if (persistent) {
            // One or a few systems: the persistent kernel (a group of CTAs per system, everything L2-resident, barriers instead
            // of launches).  A batch of many systems: one launch per phase over all of them (HBM-bound, high occupancy).
            const char* e = std::getenv("TM_KRYLOV");
            const bool batch = e ? std::strcmp(e, "phased") == 0 : m->topo.n_comp > 8;
            if (batch) krylov_solve_phased(m, *m->ranks[0], o, st);
            else krylov_solve_persistent(m, *m->ranks[0], o, st);
        }
        for (int cycle = 0; !persistent; ++cycle) {
            if (cycle > 0) refresh_x();
            for (auto& rp : m->ranks) {
                RankMesh& r = *rp;
                launch_rows<MODE_RESID, 4>(m, r, true, xnew(r), r.X[r.cur].p, r.kr.p, 1.0, nullptr);  // r = D^-1 (b - A x)
            }
            st->operator_applications += 1;
            launch_reduce(m, cycle == 0 ? RED_INIT : RED_RESTART, o, true);
            fetch_ctl(m);
            const int d0 = m->h_ctl->done[0], d1 = m->h_ctl->done[1];
            if ((d0 == 1 && d1 == 1) || d0 == 3 || d1 == 3 || cycle > max_restarts) break;
            for (auto& rp : m->ranks) {
                RankMesh& r = *rp;
                CUDA_TRY(cudaMemcpyAsync(r.krhat.p, r.kr.p, size_t(r.N) * sizeof(double2), cudaMemcpyDeviceToDevice, s));
                r.kp.zero(s);
                r.kv.zero(s);
            }
            bicgstab_cycle(m, o, st);  // accumulates the correction d (from zero) with A d ~ r
            for (auto& rp : m->ranks) {
                RankMesh& r = *rp;
                LAUNCH(add_correction_kernel, r.vec_grid, VEC_THREADS, s, r.L.n_own, xnew(r), r.kd.p);
            }
        }
        if (!persistent) {
            refresh_x();
            st->inner_iterations += uint64_t(m->h_ctl->iters[0]) + uint64_t(m->h_ctl->iters[1]);
            st->last_inner_residual = std::fmax(m->h_ctl->norm_r[0], m->h_ctl->norm_r[1]);
            if (m->h_ctl->done[0] != 1 || m->h_ctl->done[1] != 1) st->converged = 0;  // log.warn "did not converge", BiCGStab.zig:368-369
        }
        for (auto& rp : m->ranks) {
            RankMesh& r = *rp;
            LAUNCH(diff_stats_kernel, r.vec_grid, VEC_THREADS, s, r.L.n_own, (const double2*)r.X[r.cur].p, (const double2*)r.X[1 - r.cur].p, r.part_vec.p);
        }
        launch_reduce(m, RED_UPDATE_STATS, o, false);
        for (auto& rp : m->ranks) rp->cur = 1 - rp->cur;  // copy-back (smooth.zig:139-153) is a buffer swap
        m->outer_done += 1;
        st->outer_iterations += 1;
        if (o->stop_max_update > 0.0) {
            fetch_ctl(m);
            if (m->h_ctl->max_update <= o->stop_max_update) break;
        }
    }
}

#include "multigrid.inl"  // run_fas_multigrid: single-block and multi-block hierarchies, Anderson step

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
// the rank (held by this process) that owns a global block, and the block's position in its own-block list
RankMesh& owner_of_block(tm_mesh* m, size_t block, size_t* pos = nullptr) {
    check_mesh(m);
    if (block >= m->topo.blocks.size()) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block index %zu out of range", block);
    for (auto& rp : m->ranks) {
        if (rp->L.rank != m->owner[block]) continue;
        const auto& ob = rp->L.own_blocks;
        const auto it = std::lower_bound(ob.begin(), ob.end(), int32_t(block));
        if (pos) *pos = size_t(it - ob.begin());
        return *rp;
    }
    TM_THROW(TM_ERR_INVALID_ARGUMENT, "block %zu is owned by rank %d, not by this process", block, m->owner[block]);
}

void create_common(tm_mesh* m, const tm_block* blocks, size_t n_blocks, const tm_connection* connections, size_t n_connections,
                   const tm_condition* conditions, size_t n_conditions, int device, void* stream) {
    if ((n_connections && !connections) || (n_conditions && !conditions)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "NULL connection / condition array");
    require_device(device);
    if (device < 0) CUDA_TRY(cudaGetDevice(&m->device)); else m->device = device;
    CUDA_TRY(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device));
    m->topo.build(blocks, n_blocks, connections, n_connections, conditions, n_conditions);
    m->h_blocks.assign(blocks, blocks + n_blocks);
    for (auto& b : m->h_blocks) b.xy = nullptr;
    if (n_connections) m->h_conns.assign(connections, connections + n_connections);
    if (n_conditions) m->h_bcs.assign(conditions, conditions + n_conditions);
    if (const char* e = std::getenv("TM_MG_AA")) m->mg_aa = std::atoi(e) != 0;
    if (const char* e = std::getenv("TM_MG_AA_WINDOW")) m->mg_aa_window = std::min(AA_MAX, std::max(1, std::atoi(e)));
    if (const char* e = std::getenv("TM_MG_REPLICATE_NODES")) m->mg_replicate_nodes = std::atoll(e);
    if (const char* e = std::getenv("TM_MG_SMALL_NODES")) m->mg_small_nodes = std::atoll(e);
    if (const char* e = std::getenv("TM_TILE_ROWS")) m->tile_rows = std::max(4, std::atoi(e));
    if (const char* e = std::getenv("TM_INTERIOR")) m->use_bulk = std::strcmp(e, "regs") != 0;
    if (stream) m->stream = (cudaStream_t)stream;
    else { CUDA_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking)); m->own_stream = true; }
    {
        int lo = 0, hi = 0;  // numerically lowest = highest priority: the few exchange CTAs must not queue behind the sweep's
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&m->comm_stream, cudaStreamNonBlocking, hi));
    }
    {
        int lo = 0, hi = 0;  // highest priority: its few CTAs must not queue behind the waves of a sweep
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&m->copy_stream, cudaStreamNonBlocking, hi));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_snap, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_copied, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_rim, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_x, cudaEventDisableTiming));
    if (const char* e = std::getenv("TM_OVERLAP")) m->overlap = std::atoi(e) != 0;
    if (const char* e = std::getenv("TM_GRAPH")) m->use_graph = std::atoi(e) != 0;
    CUDA_TRY(cudaEventCreate(&m->ev0));
    CUDA_TRY(cudaEventCreate(&m->ev1));
    m->h_ctl = acquire_pinned_ctl();
    std::memset(m->h_ctl, 0, sizeof(SolveCtl));
}

void upload_initial(tm_mesh* m, const tm_block* blocks, size_t n_blocks) {
    for (size_t b = 0; b < n_blocks; ++b) {
        if (!blocks[b].xy) continue;
        bool mine = false;
        for (auto& rp : m->ranks) mine = mine || rp->L.rank == m->owner[b];
        if (!mine) continue;
        size_t pos = 0;
        RankMesh& r = owner_of_block(m, b, &pos);
        CUDA_TRY(cudaMemcpyAsync(r.X[r.cur].p + r.L.loff[b], blocks[b].xy, size_t(blocks[b].ni * blocks[b].nj) * sizeof(double2), cudaMemcpyHostToDevice, m->stream));
        r.have_coords[pos] = 1;
    }
    CUDA_TRY(cudaStreamSynchronize(m->stream));
}

void tfi_launch(tm_mesh* m, RankMesh& r, size_t block, size_t pos) {
    const auto& B = m->topo.blocks[block];
    const int ni = int(B.ni), nj = int(B.nj);
    const double* e = r.edges[pos].buf.p;
    // layout of the cache: x_i_min[2ni] x_i_max[2ni] x_j_min[2nj] x_j_max[2nj] s1[ni] s2[ni] t1[nj] t2[nj]
    const double2* x_i_min = (const double2*)e;
    const double2* x_i_max = x_i_min + ni;
    const double2* x_j_min = x_i_max + ni;
    const double2* x_j_max = x_j_min + nj;
    const double* s1 = (const double*)(x_j_max + nj);
    const double *s2 = s1 + ni, *t1 = s2 + ni, *t2 = t1 + nj;
    dim3 grid((nj + TILE_J - 1) / TILE_J, (ni + TFI_ROWS - 1) / TFI_ROWS);
    LAUNCH(tfi_kernel, grid, TILE_J, m->stream, ni, nj, x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2, r.X[r.cur].p + r.L.loff[block]);
    r.have_coords[pos] = 1;
    m->begun = false;
}

void tfi_validate_host(uint64_t ni, uint64_t nj, const double* x_i_min, const double* x_i_max, const double* x_j_min, const double* x_j_max,
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

#include "streamed.inl"  // smooth_streamed: tm_smooth_mesh on a large single block, copies overlapped with the sweeps

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
#pragma GCC visibility push(default)
extern "C" {

const char* tm_last_error(void) { return g_last_error.c_str(); }
int tm_abi_version(void) { return TM_ABI_VERSION; }
void tm_release_cached_memory(void) {
    release_parked_slots();  // the window meshes of the streamed tm_smooth_mesh (streamed.inl)
    g_cache.clear();
    std::lock_guard<std::mutex> lock(g_pinned_mu);
    for (SolveCtl* q : g_pinned_free) cudaFreeHost(q);
    g_pinned_free.clear();
}
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
        m = new tm_mesh();
        create_common(m, blocks, n_blocks, connections, n_connections, conditions, n_conditions, device, stream);
        m->n_ranks = 1;
        m->owner.assign(n_blocks, 0);
        m->ranks.emplace_back(new RankMesh());
        build_rank(m, m->topo, *m->ranks[0], 0);
        upload_initial(m, blocks, n_blocks);
        *out = m;
    });
    if (rc != TM_OK) delete m;
    return rc;
}

int tm_dist_get_unique_id(uint8_t* id) {
    return guarded([&] {
        if (!id) TM_THROW(TM_ERR_INVALID_ARGUMENT, "id is NULL");
        static_assert(sizeof(ncclUniqueId) <= TM_UNIQUE_ID_BYTES, "unique id does not fit");
        g_nccl.load();
        ncclUniqueId uid;
        NCCL_TRY(g_nccl.GetUniqueId(&uid));
        std::memset(id, 0, TM_UNIQUE_ID_BYTES);
        std::memcpy(id, &uid, sizeof uid);
    });
}

int tm_mesh_create_distributed(const tm_block* blocks, size_t n_blocks, const tm_connection* connections, size_t n_connections,
                               const tm_condition* conditions, size_t n_conditions, const int32_t* block_owner, int rank, int n_ranks,
                               const uint8_t* unique_id, int device, void* stream, tm_mesh** out) {
    if (out) *out = nullptr;
    tm_mesh* m = nullptr;
    int rc = guarded([&] {
        if (!out || !block_owner) TM_THROW(TM_ERR_INVALID_ARGUMENT, "out / block_owner is NULL");
        if (n_ranks < 1 || n_ranks > 16) TM_THROW(TM_ERR_INVALID_ARGUMENT, "n_ranks must be in [1, 16]");
        m = new tm_mesh();
        create_common(m, blocks, n_blocks, connections, n_connections, conditions, n_conditions, device, stream);
        m->n_ranks = n_ranks;
        m->owner.assign(block_owner, block_owner + n_blocks);
        validate_owner(m->topo, m->owner, n_ranks);
        if (rank < 0) {  // all ranks emulated in this process on one GPU
            m->emulated = n_ranks > 1;
            for (int r = 0; r < n_ranks; ++r) {
                m->ranks.emplace_back(new RankMesh());
                build_rank(m, m->topo, *m->ranks.back(), r);
            }
        } else {
            if (rank >= n_ranks) TM_THROW(TM_ERR_INVALID_ARGUMENT, "rank %d out of range", rank);
            if (n_ranks > 1) {
                if (!unique_id) TM_THROW(TM_ERR_INVALID_ARGUMENT, "unique_id is NULL");
                g_nccl.load();
                ncclUniqueId uid;
                std::memcpy(&uid, unique_id, sizeof uid);
                NCCL_TRY(g_nccl.CommInitRank(&m->comm, n_ranks, uid, rank));
            }
            m->ranks.emplace_back(new RankMesh());
            build_rank(m, m->topo, *m->ranks[0], rank);
            p2p_setup(m, *m->ranks[0]);
        }
        upload_initial(m, blocks, n_blocks);
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
        size_t pos = 0;
        RankMesh& r = owner_of_block(m, block, &pos);
        if (!xy) TM_THROW(TM_ERR_INVALID_ARGUMENT, "xy is NULL");
        CUDA_TRY(cudaSetDevice(m->device));
        const auto& B = m->topo.blocks[block];
        CUDA_TRY(cudaMemcpyAsync(r.X[r.cur].p + r.L.loff[block], xy, size_t(B.ni * B.nj) * sizeof(double2), cudaMemcpyHostToDevice, m->stream));
        CUDA_TRY(cudaStreamSynchronize(m->stream));
        r.have_coords[pos] = 1;
        m->begun = false;
    });
}

int tm_mesh_download_block(tm_mesh* m, size_t block, double* xy) {
    return guarded([&] {
        RankMesh& r = owner_of_block(m, block);
        if (!xy) TM_THROW(TM_ERR_INVALID_ARGUMENT, "xy is NULL");
        CUDA_TRY(cudaSetDevice(m->device));
        const auto& B = m->topo.blocks[block];
        CUDA_TRY(cudaMemcpyAsync(xy, r.X[r.cur].p + r.L.loff[block], size_t(B.ni * B.nj) * sizeof(double2), cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(cudaStreamSynchronize(m->stream));
    });
}

// Asynchronous read-back: the block is snapshot on the device (a device-to-device copy on the mesh's stream, ~0.1 ms per
// GB) and the snapshot goes to the host on a stream of its own, so the mesh can be overwritten -- the next TFI, the next
// smoothing call -- while the previous result is still on its way.  The blocks of one step are snapshot into one buffer.
int tm_mesh_download_block_async(tm_mesh* m, size_t block, double* xy) {
    return guarded([&] {
        RankMesh& r = owner_of_block(m, block);
        if (!xy) TM_THROW(TM_ERR_INVALID_ARGUMENT, "xy is NULL");
        CUDA_TRY(cudaSetDevice(m->device));
        const auto& B = m->topo.blocks[block];
        if (r.snapshot.n != size_t(r.N)) r.snapshot.alloc(size_t(r.N));
        const size_t off = size_t(r.L.loff[block]), bytes = size_t(B.ni * B.nj) * sizeof(double2);
        if (r.snap_copied.size() != m->topo.blocks.size()) r.snap_copied.assign(m->topo.blocks.size(), nullptr);
        cudaEvent_t& done = r.snap_copied[block];
        // this block's part of the snapshot may still be read by the copy an earlier call started (other blocks' copies do not matter)
        if (done) CUDA_TRY(cudaStreamWaitEvent(m->stream, done, 0));
        else CUDA_TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        CUDA_TRY(cudaMemcpyAsync(r.snapshot.p + off, r.X[r.cur].p + off, bytes, cudaMemcpyDeviceToDevice, m->stream));
        CUDA_TRY(cudaEventRecord(m->ev_snap, m->stream));
        CUDA_TRY(cudaStreamWaitEvent(m->copy_stream, m->ev_snap, 0));
        // Page-locked destination: a few CTAs store the snapshot straight into the host memory (it is mapped into the device's
        // address space), not the copy engine -- the mesh's own small read-backs (statistics, checks) use the device-to-host
        // engine in submission order and would wait for the whole block behind a DMA copy of it.
        cudaPointerAttributes attr{};
        void* mapped = nullptr;
        if (cudaPointerGetAttributes(&attr, xy) == cudaSuccess && attr.type == cudaMemoryTypeHost && cudaHostGetDevicePointer(&mapped, xy, 0) == cudaSuccess && mapped &&
            (reinterpret_cast<uintptr_t>(mapped) & 15) == 0) {
            static const int store_ctas = [] { const char* e = std::getenv("TM_HOST_STORE_CTAS"); return e ? std::max(1, std::atoi(e)) : 16; }();
            LAUNCH(host_store_kernel, unsigned(store_ctas), 256, m->copy_stream, reinterpret_cast<double2*>(mapped), (const double2*)(r.snapshot.p + off), int64_t(B.ni * B.nj));
        } else {
            (void)cudaGetLastError();
            CUDA_TRY(cudaMemcpyAsync(xy, r.snapshot.p + off, bytes, cudaMemcpyDeviceToHost, m->copy_stream));
        }
        CUDA_TRY(cudaEventRecord(done, m->copy_stream));
        m->copies_pending = true;
    });
}
int tm_mesh_download_wait(tm_mesh* m) {
    return guarded([&] {
        if (!m) TM_THROW(TM_ERR_INVALID_ARGUMENT, "mesh is NULL");
        CUDA_TRY(cudaSetDevice(m->device));
        CUDA_TRY(cudaStreamSynchronize(m->copy_stream));
        m->copies_pending = false;
    });
}

int tm_mesh_tfi_block(tm_mesh* m, size_t block, const double* x_i_min, const double* x_i_max, const double* x_j_min, const double* x_j_max,
                      const double* s1, const double* s2, const double* t1, const double* t2) {
    return guarded([&] {
        size_t pos = 0;
        RankMesh& r = owner_of_block(m, block, &pos);
        CUDA_TRY(cudaSetDevice(m->device));
        const auto& B = m->topo.blocks[block];
        const size_t ni = size_t(B.ni), nj = size_t(B.nj);
        tfi_validate_host(ni, nj, x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2);
        EdgeCache& ec = r.edges[pos];
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
        tfi_launch(m, r, block, pos);
        CUDA_TRY(cudaStreamSynchronize(m->stream));  // the host edge arrays may be freed after return
    });
}

int tm_mesh_tfi_block_resident(tm_mesh* m, size_t block) {
    return guarded([&] {
        size_t pos = 0;
        RankMesh& r = owner_of_block(m, block, &pos);
        CUDA_TRY(cudaSetDevice(m->device));
        if (!r.edges[pos].valid) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block %zu has no cached edges; call tm_mesh_tfi_block first", block);
        tfi_launch(m, r, block, pos);
    });
}

int tm_mesh_begin_smoothing(tm_mesh* m, const tm_smooth_options* o) {
    return guarded([&] {
        check_mesh(m);
        validate_options(o);
        CUDA_TRY(cudaSetDevice(m->device));
        cudaStream_t s = m->stream;
        if (const char* e = std::getenv("TM_MG_AA")) m->mg_aa = std::atoi(e) != 0;   // (may be switched per smoothing run; the buffers exist if it was on at creation)
        else m->mg_aa = true;
        for (auto& rp : m->ranks)
            for (size_t k = 0; k < rp->have_coords.size(); ++k)
                if (!rp->have_coords[k]) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block %d has no coordinates yet", int(rp->L.own_blocks[k]));
        auto xcur = [](RankMesh& r) { return r.X[r.cur].p; };
        exchange(m, xcur, true);  // raw side-0 coordinates of cross-rank interface pairs
        // connectionDataCheck (smooth.zig:220-275); each pair is checked by the rank that owns its side-1 node
        for (auto& rp : m->ranks) {
            RankMesh& r = *rp;
            CUDA_TRY(cudaMemsetAsync(r.d_worst.p, 0, sizeof(unsigned long long), s));
            const int np = int(r.L.pairs.size());
            if (np > 0) LAUNCH(pair_check_kernel, (np + 255) / 256, 256, s, (const PairCheck*)r.d_pairs.p, np, (const double2*)xcur(r), 1e-15, r.d_worst.p);
        }
        int bad_conn = -1, bad_point = -1;
        for (auto& rp : m->ranks) {
            unsigned long long worst = 0;
            CUDA_TRY(cudaMemcpyAsync(&worst, rp->d_worst.p, sizeof worst, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            if (worst != 0 && bad_conn < 0) {
                const PairCheck& p = rp->L.pairs[size_t(worst & 0xffffffffull)];
                bad_conn = p.conn; bad_point = p.point;
            }
        }
        // every rank must fail together: a rank that threw alone would leave its peers waiting in the next exchange
        bool elsewhere = false;
        if (m->n_ranks > 1 && !m->emulated) {
            RankMesh& r0 = *m->ranks[0];
            const double mine_bad = bad_conn >= 0 ? 1.0 : 0.0;
            double any_bad = 0.0;
            CUDA_TRY(cudaMemcpyAsync(r0.red.p, &mine_bad, sizeof(double), cudaMemcpyHostToDevice, s));
            NCCL_TRY(g_nccl.AllReduce(r0.red.p, r0.red.p, 1, ncclDouble, ncclMax, m->comm, s));
            CUDA_TRY(cudaMemcpyAsync(&any_bad, r0.red.p, sizeof(double), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            elsewhere = any_bad > 0.0 && bad_conn < 0;
        }
        if (bad_conn >= 0) TM_THROW(TM_ERR_TOPOLOGY, "non matching points for connection %d point %d (tolerance 1e-15 abs, smooth.zig:220-275)", bad_conn, bad_point);
        if (elsewhere) TM_THROW(TM_ERR_TOPOLOGY, "non matching interface points were found by another rank (connectionDataCheck, smooth.zig:220-275)");
        // rhs of fixed / sliding rows is captured from the initial mesh (smooth.zig:790-796, 853-858); then every copy of
        // a node is made exactly consistent with its root (x_copy = x_root + shift), constants included
        for (auto& rp : m->ranks) {
            RankMesh& r = *rp;
            const int n_l = int(r.L.sliding.size()), n_fo = int(r.L.fixed_overrides.size());
            if (n_l + n_fo > 0) LAUNCH(capture_boundary_kernel, (n_l + n_fo + 127) / 128, 128, s, r.d_lrows.p, n_l, (const FixedOverride*)r.d_fo.p, n_fo, xcur(r));
        }
        exchange(m, xcur);
        for (auto& rp : m->ranks) {
            RankMesh& r = *rp;
            const int n_cs = int(r.L.const_slaves.size());
            if (n_cs > 0) LAUNCH(sync_slaves_kernel, (n_cs + 127) / 128, 128, s, (const SlaveRow*)r.d_cslaves.p, n_cs, xcur(r), 1);
            sync_slaves(m, r, xcur(r), 1);
            CUDA_TRY(cudaMemcpyAsync(r.X[1 - r.cur].p, xcur(r), size_t(r.N) * sizeof(double2), cudaMemcpyDeviceToDevice, s));
            // this rank's constant part of ||b||^2
            if (r.d_rhs_terms.n > 0) LAUNCH(rhs_const_kernel, 1, VEC_THREADS, s, (const RhsTerm*)r.d_rhs_terms.p, int(r.d_rhs_terms.n), (const double2*)xcur(r), r.bconst.p);
            else r.bconst.zero(s);
        }
        // control function (ControlFunction.init, wall_control_function.zig:27-42)
        m->cf = int(o->control_function);
        for (auto& rp : m->ranks) rp->has_pq = false;
        if (m->cf == TM_CF_WHITE) {
            if (!m->topo.white_ok) TM_THROW(TM_ERR_UNSUPPORTED, "%s", m->topo.white_why.c_str());
            for (const auto& g : m->topo.white_groups)
                if (m->owner[size_t(g.first)] != m->owner[size_t(g.second)])
                    TM_THROW(TM_ERR_UNSUPPORTED, "White control function: blocks %lld and %lld must be owned by the same rank", (long long)g.first, (long long)g.second);
            for (auto& rp : m->ranks) {
                RankMesh& r = *rp;
                std::vector<WhiteParams> groups;
                std::vector<WhiteNode> nodes;
                int32_t wall_base = 0;
                for (const auto& g : m->topo.white_groups) {
                    if (m->owner[size_t(g.first)] != r.L.rank) continue;
                    const auto& B0 = m->topo.blocks[size_t(g.first)];
                    const auto& B1 = m->topo.blocks[size_t(g.second)];
                    WhiteParams w{};
                    w.off0 = r.L.loff[size_t(g.first)]; w.off1 = r.L.loff[size_t(g.second)];
                    w.ni0 = int32_t(B0.ni); w.nj0 = int32_t(B0.nj); w.ni1 = int32_t(B1.ni); w.nj1 = int32_t(B1.nj);
                    w.c_in0 = int32_t(B0.nj); w.c_in1 = int32_t(B1.nj); w.c_al0 = 1;  // j_min sides starting at node 0, running towards +j
                    w.wall_base = wall_base;
                    for (int32_t t = 0; t < w.ni0 + w.ni1; ++t) nodes.push_back(WhiteNode{int32_t(groups.size()), t});
                    wall_base += w.ni0 + w.ni1;
                    groups.push_back(w);
                }
                r.n_wnodes = int(nodes.size());
                if (groups.empty()) continue;
                r.d_wgroups.upload(groups, s);
                r.d_wnodes.upload(nodes, s);
                if (r.pq.n != size_t(r.N)) r.pq.alloc(size_t(r.N));
                r.pq.zero(s);
                r.wall_pq.alloc(size_t(wall_base));
                r.has_pq = true;
            }
            white_step(m, false, o->white_ds_target, o->white_theta_target);
        }
        if (o->solver == TM_SOLVER_FAS_MULTIGRID && m->n_ranks == 1 && m->topo.blocks.size() == 1) mg_build(m, *m->ranks[0]);
        for (auto& rp : m->ranks)
            for (auto& lv : rp->mg) { lv->aa_head = -1; lv->aa_count = 0; lv->aa_have_x = false; }
        for (auto& lv : m->mgb)
            for (auto& rp : lv->ranks) {
                rp->mg_primed = false;
                rp->mg_E.zero(s);
                rp->aa_head = -1; rp->aa_count = 0; rp->aa_have_x = false;
            }
        for (auto& rp : m->ranks) { if (rp->kplan) rp->kplan->coarse_age = -1; if (rp->pplan) rp->pplan->coarse_age = -1; }   // the coarse operator is rebuilt from the new mesh
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
        st.nodes = uint64_t(m->topo.n_nodes);
        st.converged = 1;
        CUDA_TRY(cudaEventRecord(m->ev0, m->stream));
        if (o->solver == TM_SOLVER_RELAX) run_relax(m, o, &st);
        else if (o->solver == TM_SOLVER_FAS_MULTIGRID) run_fas_multigrid(m, o, &st);
        else run_picard_bicgstab(m, o, &st);
        CUDA_TRY(cudaEventRecord(m->ev1, m->stream));
        fetch_ctl(m);
        for (auto& rp : m->ranks) p2p_check(m, *rp);
        for (auto& lv : m->mgb) for (auto& rp : lv->ranks) p2p_check(m, *rp);
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

int tm_mesh_set_white_groups(tm_mesh* m, const uint64_t* block_pairs, size_t n_groups) {
    return guarded([&] {
        check_mesh(m);
        if (n_groups && !block_pairs) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block_pairs is NULL");
        m->topo.set_white_groups(block_pairs, n_groups);
        m->begun = false;
    });
}

uint64_t tm_mesh_component_count(const tm_mesh* m) { return m ? uint64_t(m->topo.n_comp) : 0; }
int tm_mesh_component_of_block(const tm_mesh* m, size_t block, uint64_t* component) {
    return guarded([&] {
        check_mesh(m);
        if (block >= m->topo.blocks.size() || !component) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block index out of range / component is NULL");
        *component = uint64_t(m->topo.comp_of_block[block]);
    });
}
int tm_mesh_component_stats(const tm_mesh* m, size_t component, tm_component_stats* out) {
    return guarded([&] {
        check_mesh(m);
        if (!out || component >= size_t(m->topo.n_comp)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "component index out of range / out is NULL");
        if (m->ranks.size() != 1 || (!m->ranks[0]->kplan && !m->ranks[0]->pplan))
            TM_THROW(TM_ERR_UNSUPPORTED, "per-component records exist after a TM_SOLVER_PICARD_BICGSTAB solve of a single-process mesh");
        std::memset(out, 0, sizeof *out);
        if (m->ranks[0]->pplan) {
            const PhasedPlan& P = *m->ranks[0]->pplan;
            const KState& k = P.h_state[component];
            out->nodes = uint64_t(P.h_comps[component].nodes);
            for (int c = 0; c < 2; ++c) {
                out->iterations[c] = uint64_t(k.iters[c]); out->tolerance[c] = k.tol[c]; out->norm_b[c] = k.norm_b[c]; out->norm_r[c] = k.norm_r[c];
                out->status[c] = k.done[c];
            }
            out->operator_applications = uint64_t(k.applications);
            out->restarts = uint64_t(k.cycles);
            return;
        }
        const KrylovPlan& P = *m->ranks[0]->kplan;
        const KCtl& k = P.h_ctl[component];
        out->nodes = uint64_t(P.h_comps[component].nodes);
        for (int c = 0; c < 2; ++c) {
            out->iterations[c] = uint64_t(k.iters[c]); out->tolerance[c] = k.tol[c]; out->norm_b[c] = k.norm_b[c]; out->norm_r[c] = k.norm_r[c];
            out->status[c] = k.done[c];
        }
        out->operator_applications = uint64_t(k.applications);
        out->restarts = uint64_t(k.cycles);
    });
}
uint64_t tm_mesh_block_count(const tm_mesh* m) { return m ? m->topo.blocks.size() : 0; }
uint64_t tm_mesh_node_count(const tm_mesh* m) { return m ? uint64_t(m->topo.n_nodes) : 0; }
int tm_mesh_halo_path(const tm_mesh* m) {
    if (!m || m->n_ranks < 2) return TM_HALO_NONE;
    if (m->emulated) return TM_HALO_EMULATED;
    return (!m->ranks.empty() && m->ranks[0]->p2p.ready) ? TM_HALO_PEER_MEMORY : TM_HALO_NCCL;
}
uint64_t tm_mesh_local_node_count(const tm_mesh* m) {
    uint64_t n = 0;
    if (m) for (const auto& rp : m->ranks) n += uint64_t(rp->L.n_own);
    return n;
}
int tm_mesh_block_size(const tm_mesh* m, size_t block, uint64_t* ni, uint64_t* nj) {
    return guarded([&] {
        check_mesh(m);
        if (block >= m->topo.blocks.size()) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block index %zu out of range", block);
        if (ni) *ni = uint64_t(m->topo.blocks[block].ni);
        if (nj) *nj = uint64_t(m->topo.blocks[block].nj);
    });
}
double* tm_mesh_block_device_ptr(tm_mesh* m, size_t block) {
    if (!m || block >= m->topo.blocks.size()) return nullptr;
    for (auto& rp : m->ranks)
        if (rp->L.rank == m->owner[block]) return reinterpret_cast<double*>(rp->X[rp->cur].p + rp->L.loff[block]);
    return nullptr;
}
int tm_mesh_download_control_function(tm_mesh* m, size_t block, double* pqv) {
    return guarded([&] {
        RankMesh& r = owner_of_block(m, block);
        if (!pqv) TM_THROW(TM_ERR_INVALID_ARGUMENT, "pq is NULL");
        CUDA_TRY(cudaSetDevice(m->device));
        const auto& B = m->topo.blocks[block];
        const size_t bytes = size_t(B.ni * B.nj) * sizeof(double2);
        if (!r.has_pq || !r.pq.p) { std::memset(pqv, 0, bytes); return; }  // laplace: all zero (wall_control_function.zig:29-33)
        CUDA_TRY(cudaMemcpyAsync(pqv, r.pq.p + r.L.loff[block], bytes, cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(cudaStreamSynchronize(m->stream));
    });
}
int tm_mesh_download_block_soa(tm_mesh* m, size_t block, int field, double* x, double* y) {
    return guarded([&] {
        RankMesh& r = owner_of_block(m, block);
        if (!x || !y) TM_THROW(TM_ERR_INVALID_ARGUMENT, "x / y is NULL");
        if (field != TM_FIELD_COORDINATES && field != TM_FIELD_CONTROL_FUNCTION) TM_THROW(TM_ERR_INVALID_ARGUMENT, "unknown field %d", field);
        CUDA_TRY(cudaSetDevice(m->device));
        const auto& B = m->topo.blocks[block];
        const size_t n = size_t(B.ni * B.nj);
        if (field == TM_FIELD_CONTROL_FUNCTION && (!r.has_pq || !r.pq.p)) {  // laplace: all zero (wall_control_function.zig:29-33)
            std::memset(x, 0, n * sizeof(double));
            std::memset(y, 0, n * sizeof(double));
            return;
        }
        const double2* src = (field == TM_FIELD_COORDINATES ? r.X[r.cur].p : r.pq.p) + r.L.loff[block];
        if (r.soa_stage.n < 2 * n) r.soa_stage.alloc(2 * n);
        dim3 grid(unsigned((B.nj + 31) / 32), unsigned((B.ni + 31) / 32));
        LAUNCH(aos_to_soa_kernel, grid, 256, m->stream, int(B.ni), int(B.nj), src, r.soa_stage.p, r.soa_stage.p + n);
        CUDA_TRY(cudaMemcpyAsync(x, r.soa_stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(cudaMemcpyAsync(y, r.soa_stage.p + n, n * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        CUDA_TRY(cudaStreamSynchronize(m->stream));
    });
}
int tm_mesh_write_plot3d(tm_mesh* m, const char* grid_path, const char* function_path) {
    return guarded([&] {
        check_mesh(m);
        if (!grid_path) TM_THROW(TM_ERR_INVALID_ARGUMENT, "grid_path is NULL");
        CUDA_TRY(cudaSetDevice(m->device));
        struct File {
            FILE* f = nullptr;
            ~File() { if (f) std::fclose(f); }
        };
        std::vector<std::pair<RankMesh*, int32_t>> own;   // blocks held by this process, in global block order
        for (size_t b = 0; b < m->topo.blocks.size(); ++b)
            for (auto& rp : m->ranks)
                if (rp->L.rank == m->owner[b]) own.push_back({rp.get(), int32_t(b)});
        std::vector<double> host;
        for (int pass = 0; pass < (function_path ? 2 : 1); ++pass) {
            File out;
            out.f = std::fopen(pass == 0 ? grid_path : function_path, "wb");
            if (!out.f) TM_THROW(TM_ERR_INVALID_ARGUMENT, "cannot open %s for writing", pass == 0 ? grid_path : function_path);
            auto put = [&](const void* p, size_t bytes) { if (std::fwrite(p, 1, bytes, out.f) != bytes) TM_THROW(TM_ERR_INVALID_ARGUMENT, "short write (disk full?)"); };
            const int32_t nb = int32_t(own.size());
            put(&nb, sizeof nb);
            for (const auto& ob : own) {
                const auto& B = m->topo.blocks[size_t(ob.second)];
                const int32_t dims[3] = {int32_t(B.ni), int32_t(B.nj), 2};
                put(dims, (pass == 0 ? 2 : 3) * sizeof(int32_t));
            }
            for (const auto& ob : own) {
                RankMesh& r = *ob.first;
                const auto& B = m->topo.blocks[size_t(ob.second)];
                const size_t n = size_t(B.ni * B.nj);
                host.resize(2 * n);
                if (pass == 1 && (!r.has_pq || !r.pq.p)) {   // laplace: all zero (wall_control_function.zig:29-33)
                    std::fill(host.begin(), host.end(), 0.0);
                } else {
                    const double2* src = (pass == 0 ? r.X[r.cur].p : r.pq.p) + r.L.loff[size_t(ob.second)];
                    if (r.soa_stage.n < 2 * n) r.soa_stage.alloc(2 * n);
                    dim3 grid(unsigned((B.nj + 31) / 32), unsigned((B.ni + 31) / 32));
                    LAUNCH(aos_to_soa_kernel, grid, 256, m->stream, int(B.ni), int(B.nj), src, r.soa_stage.p, r.soa_stage.p + n);
                    CUDA_TRY(cudaMemcpyAsync(host.data(), r.soa_stage.p, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
                    CUDA_TRY(cudaStreamSynchronize(m->stream));
                }
                put(host.data(), 2 * n * sizeof(double));
            }
        }
    });
}
int tm_mesh_viewer_sizes(const tm_mesh* m, uint64_t* n_points, uint64_t* n_indices) {
    return guarded([&] {
        check_mesh(m);
        uint64_t np = 0, nx = 0;
        for (const auto& B : m->topo.blocks) { np += uint64_t(B.ni * B.nj); nx += uint64_t(B.ni * (B.nj - 1) * 2 + B.nj * (B.ni - 1) * 2); }
        if (n_points) *n_points = np;
        if (n_indices) *n_indices = nx;
    });
}
int tm_mesh_viewer_buffers(tm_mesh* m, float* points, float* ranges, uint32_t* indices) {
    return guarded([&] {
        check_mesh(m);
        if (m->n_ranks != 1) TM_THROW(TM_ERR_UNSUPPORTED, "viewer buffers are built for single-GPU meshes");
        if (m->topo.n_nodes >= (int64_t(1) << 32)) TM_THROW(TM_ERR_UNSUPPORTED, "more than 2^32 points: 32-bit line indices (gl.uint, gui/lib.zig:267) cannot address them");
        CUDA_TRY(cudaSetDevice(m->device));
        RankMesh& r = *m->ranks[0];
        cudaStream_t s = m->stream;
        const int64_t n = r.L.n_own;  // one rank: the local field is all blocks in the reference's order
        if (points || ranges) {
            DevBuf<float2> d_pts;
            DevBuf<float> d_part, d_rng;
            d_pts.alloc(size_t(n));
            d_part.alloc(size_t(r.vec_grid) * 4);
            d_rng.alloc(4);
            LAUNCH(viewer_points_kernel, r.vec_grid, 256, s, n, (const double2*)r.X[r.cur].p, d_pts.p, d_part.p);
            LAUNCH(viewer_ranges_kernel, 1, 32, s, (const float*)d_part.p, r.vec_grid, d_rng.p);
            // cudaMemcpyDefault: the destinations may be host memory or device memory (a mapped GL buffer)
            if (points) CUDA_TRY(cudaMemcpyAsync(points, d_pts.p, size_t(n) * sizeof(float2), cudaMemcpyDefault, s));
            if (ranges) CUDA_TRY(cudaMemcpyAsync(ranges, d_rng.p, 4 * sizeof(float), cudaMemcpyDefault, s));
            CUDA_TRY(cudaStreamSynchronize(s));
        }
        if (indices) {
            std::vector<ViewerBlock> vb;
            int64_t poff = 0, ioff = 0, most = 0;
            for (const auto& B : m->topo.blocks) {
                vb.push_back(ViewerBlock{poff, ioff, int32_t(B.ni), int32_t(B.nj)});
                const int64_t segs = B.ni * (B.nj - 1) + B.nj * (B.ni - 1);
                most = std::max(most, segs);
                poff += B.ni * B.nj;
                ioff += 2 * segs;
            }
            if (vb.size() > 65535) TM_THROW(TM_ERR_UNSUPPORTED, "more than 65535 blocks");
            DevBuf<ViewerBlock> d_vb;
            DevBuf<uint2> d_idx;
            d_vb.upload(vb, s);
            d_idx.alloc(size_t(ioff / 2));
            dim3 grid(unsigned(std::max<int64_t>(1, std::min<int64_t>((most + 255) / 256, 4096))), unsigned(vb.size()));
            LAUNCH(viewer_wireframe_kernel, grid, 256, s, (const ViewerBlock*)d_vb.p, d_idx.p);
            CUDA_TRY(cudaMemcpyAsync(indices, d_idx.p, size_t(ioff) * sizeof(uint32_t), cudaMemcpyDefault, s));
            CUDA_TRY(cudaStreamSynchronize(s));
        }
    });
}
int tm_mesh_download_boundary_kinds(tm_mesh* m, size_t block, uint8_t* kinds) {
    return guarded([&] {
        check_mesh(m);
        if (block >= m->topo.blocks.size()) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block index %zu out of range", block);
        if (!kinds) TM_THROW(TM_ERR_INVALID_ARGUMENT, "kinds is NULL");
        const auto& B = m->topo.blocks[block];
        std::memcpy(kinds, m->topo.kind.data() + B.bbuf, size_t(2 * (B.ni + B.nj - 2)));
    });
}

// host-only view of the partition (no GPU needed): what rank `rank` owns, receives and sends
int tm_dist_plan(const tm_block* blocks, size_t n_blocks, const tm_connection* connections, size_t n_connections, const tm_condition* conditions,
                 size_t n_conditions, const int32_t* block_owner, int rank, int n_ranks, tm_dist_plan_info* info, int64_t* ghost_ids, int64_t* send_ids,
                 int64_t* counts) {
    return guarded([&] {
        if (!block_owner || !info) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block_owner / info is NULL");
        if ((n_connections && !connections) || (n_conditions && !conditions)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "NULL connection / condition array");
        Topology T;
        T.build(blocks, n_blocks, connections, n_connections, conditions, n_conditions);
        std::vector<int32_t> owner(block_owner, block_owner + n_blocks);
        const LocalTables L = localize(T, owner, rank, n_ranks);
        info->n_own = uint64_t(L.n_own); info->n_ghost = uint64_t(L.n_ghost); info->n_synth = uint64_t(L.n_synth);
        info->n_send = uint64_t(L.send_lidx.size());
        info->n_smoothed = uint64_t(L.smoothed.size()); info->n_junction = uint64_t(L.junction_rows.size());
        info->n_sliding = uint64_t(L.sliding.size()); info->n_slaves = uint64_t(L.slaves.size());
        if (counts)
            for (int p = 0; p < n_ranks; ++p) { counts[2 * p] = int64_t(L.ghost_ids[size_t(p)].size()); counts[2 * p + 1] = int64_t(L.send_ids[size_t(p)].size()); }
        if (ghost_ids) { size_t k = 0; for (const auto& v : L.ghost_ids) for (int64_t g : v) ghost_ids[k++] = g; }
        if (send_ids) { size_t k = 0; for (const auto& v : L.send_ids) for (int64_t g : v) send_ids[k++] = g; }
    });
}

int tm_edges_discretize(const tm_edge_job* jobs, size_t n_jobs, int device) {
    return guarded([&] {
        if (n_jobs && !jobs) TM_THROW(TM_ERR_INVALID_ARGUMENT, "jobs is NULL");
        if (n_jobs == 0) return;
        require_device(device);
        std::vector<EdgeJob> dev_jobs(n_jobs);
        std::vector<double> tables;
        std::vector<std::pair<const tm_spline*, int64_t>> seen;  // every spline is uploaded once
        int64_t total = 0;
        for (size_t k = 0; k < n_jobs; ++k) {
            const tm_edge_job& j = jobs[k];
            if (j.n < 2 || j.n > 0x7fffffffull || !j.points || !j.clustering) TM_THROW(TM_ERR_INVALID_ARGUMENT, "edge %zu: needs n >= 2 and output arrays", k);
            if (j.curve_kind > TM_CURVE_SPLINE || j.clustering_kind > TM_CLUSTERING_SINGLE_HYPERBOLIC) TM_THROW(TM_ERR_INVALID_ARGUMENT, "edge %zu: unknown curve / clustering kind", k);
            EdgeJob e{};
            e.out_off = total; e.n = int32_t(j.n); e.curve = int32_t(j.curve_kind); e.clustering = int32_t(j.clustering_kind);
            e.line[0] = j.line_start[0]; e.line[1] = j.line_start[1]; e.line[2] = j.line_end[0]; e.line[3] = j.line_end[1];
            e.alpha = j.alpha; e.beta = j.beta;
            if (j.clustering_kind == TM_CLUSTERING_SINGLE_HYPERBOLIC) {  // Vinokur's inversion of sinh(d)/d = 1/B, clustering.zig:60-81
                const double y = 1.0 / (double(j.n - 1) * j.delta_s);
                if (!(y >= 1.0)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "edge %zu: single hyperbolic clustering needs (n-1)*delta_s <= 1 (clustering.zig:68-76)", k);
                if (y < 2.7829681) {
                    const double y_bar = y - 1.0;
                    e.delta = std::sqrt(6.0 * y_bar) * (1.0 + y_bar * (-0.15 + y_bar * (0.057321429 + y_bar * (-0.024907295 + y_bar * (0.0077424461 - 0.0010794123 * y_bar)))));
                } else {
                    const double w = 1.0 / y - 0.028527431, v = std::log(y);
                    e.delta = v + (1.0 + 1.0 / v) * std::log(2.0 * v) - 0.02041793 + w * (0.24902722 + w * (1.9496443 + w * (-2.6294547 + 8.56795911 * w)));
                }
            }
            if (j.curve_kind == TM_CURVE_SPLINE) {
                const tm_spline* sp = j.spline;
                if (!sp || sp->n_points < 2 || sp->n_samples < 2 || !sp->params || !sp->points || !sp->second_derivs_x || !sp->second_derivs_y || !sp->sample_arc)
                    TM_THROW(TM_ERR_INVALID_ARGUMENT, "edge %zu: incomplete spline", k);
                int64_t off = -1;
                for (const auto& pr : seen) if (pr.first == sp) off = pr.second;
                if (off < 0) {
                    off = int64_t(tables.size());
                    const size_t m = size_t(sp->n_points);
                    tables.insert(tables.end(), sp->params, sp->params + m);
                    tables.insert(tables.end(), sp->points, sp->points + 2 * m);
                    tables.insert(tables.end(), sp->second_derivs_x, sp->second_derivs_x + m);
                    tables.insert(tables.end(), sp->second_derivs_y, sp->second_derivs_y + m);
                    tables.insert(tables.end(), sp->sample_arc, sp->sample_arc + sp->n_samples);
                    seen.push_back({sp, off});
                }
                e.spline_off = off; e.spline_m = int32_t(sp->n_points); e.n_samples = int32_t(sp->n_samples); e.total_length = sp->total_length;
            }
            dev_jobs[k] = e;
            total += int64_t(j.n);
        }
        cudaStream_t s = nullptr;
        CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        try {
            DevBuf<EdgeJob> d_jobs;
            DevBuf<double> d_tables, d_cl;
            DevBuf<double2> d_pts;
            d_jobs.upload(dev_jobs, s);
            if (tables.empty()) tables.push_back(0.0);
            d_tables.upload(tables, s);
            d_pts.alloc(size_t(total));
            d_cl.alloc(size_t(total));
            LAUNCH(edge_discretize_kernel, unsigned(n_jobs), 128, s, (const EdgeJob*)d_jobs.p, (const double*)d_tables.p, d_pts.p, d_cl.p);
            copy_out_runs(n_jobs, (const double2*)d_pts.p, [&](size_t k) { return reinterpret_cast<double2*>(jobs[k].points); },
                          [&](size_t k) { return dev_jobs[k].out_off; }, [&](size_t k) { return size_t(jobs[k].n); }, s);
            copy_out_runs(n_jobs, (const double*)d_cl.p, [&](size_t k) { return jobs[k].clustering; }, [&](size_t k) { return dev_jobs[k].out_off; },
                          [&](size_t k) { return size_t(jobs[k].n); }, s);
            CUDA_TRY(cudaStreamSynchronize(s));
        } catch (...) {
            cudaStreamDestroy(s);
            throw;
        }
        cudaStreamDestroy(s);
    });
}

int tm_splines_fit(const tm_spline_fit_job* jobs, size_t n_jobs, int device) {
    return guarded([&] {
        if (n_jobs && !jobs) TM_THROW(TM_ERR_INVALID_ARGUMENT, "jobs is NULL");
        if (n_jobs == 0) return;
        std::vector<SplineFitJob> dj(n_jobs);
        std::vector<double> pts;
        int64_t n_pts = 0, n_arc = 0;
        for (size_t k = 0; k < n_jobs; ++k) {
            const tm_spline_fit_job& j = jobs[k];
            if (j.n_points < 2 || j.n_points > 0x7fffffffull || j.n_samples < 2 || j.n_samples > 0x7fffffffull || !j.points || !j.params || !j.second_derivs_x ||
                !j.second_derivs_y || !j.sample_arc || !j.total_length)
                TM_THROW(TM_ERR_INVALID_ARGUMENT, "spline %zu: needs >= 2 points, >= 2 samples and all output arrays (spline.zig:41-45)", k);
            dj[k] = SplineFitJob{n_pts, n_arc, int32_t(j.n_points), int32_t(j.n_samples)};
            pts.insert(pts.end(), j.points, j.points + 2 * j.n_points);
            n_pts += int64_t(j.n_points);
            n_arc += int64_t(j.n_samples);
        }
        require_device(device);
        cudaStream_t s = nullptr;
        CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        try {
            DevBuf<SplineFitJob> d_jobs;
            DevBuf<double> d_pts, d_params, d_zx, d_zy, d_tmp, d_arc, d_len;
            DevBuf<double2> d_samples;
            DevBuf<int> d_err;
            d_jobs.upload(dj, s); d_pts.upload(pts, s);
            d_params.alloc(size_t(n_pts)); d_zx.alloc(size_t(n_pts)); d_zy.alloc(size_t(n_pts)); d_tmp.alloc(size_t(2 * n_pts));
            d_arc.alloc(size_t(n_arc)); d_samples.alloc(size_t(n_arc)); d_len.alloc(n_jobs);
            d_err.alloc(1); d_err.zero(s);
            LAUNCH(spline_fit_kernel, unsigned(n_jobs), 128, s, (const SplineFitJob*)d_jobs.p, (const double2*)d_pts.p, d_params.p, d_zx.p, d_zy.p, d_tmp.p, d_arc.p,
                   d_samples.p, d_len.p, d_err.p);
            int err = 0;
            CUDA_TRY(cudaMemcpyAsync(&err, d_err.p, sizeof err, cudaMemcpyDeviceToHost, s));
            auto n_of = [&](size_t k) { return size_t(dj[k].n); };
            auto pt_off = [&](size_t k) { return dj[k].pt_off; };
            copy_out_runs(n_jobs, (const double*)d_params.p, [&](size_t k) { return jobs[k].params; }, pt_off, n_of, s);
            copy_out_runs(n_jobs, (const double*)d_zx.p, [&](size_t k) { return jobs[k].second_derivs_x; }, pt_off, n_of, s);
            copy_out_runs(n_jobs, (const double*)d_zy.p, [&](size_t k) { return jobs[k].second_derivs_y; }, pt_off, n_of, s);
            copy_out_runs(n_jobs, (const double*)d_arc.p, [&](size_t k) { return jobs[k].sample_arc; }, [&](size_t k) { return dj[k].arc_off; },
                          [&](size_t k) { return size_t(dj[k].n_samples); }, s);
            copy_out_runs(n_jobs, (const double*)d_len.p, [&](size_t k) { return jobs[k].total_length; }, [&](size_t k) { return int64_t(k); }, [&](size_t) { return size_t(1); }, s);
            CUDA_TRY(cudaStreamSynchronize(s));
            if (err) TM_THROW(TM_ERR_INVALID_ARGUMENT, "spline fit: coincident consecutive points (CoincidentParameters, spline.zig:176-178)");
        } catch (...) {
            cudaStreamDestroy(s);
            throw;
        }
        cudaStreamDestroy(s);
    });
}

int tm_edges_combine(const tm_combine_job* jobs, size_t n_jobs, int device) {
    return guarded([&] {
        if (n_jobs && !jobs) TM_THROW(TM_ERR_INVALID_ARGUMENT, "jobs is NULL");
        if (n_jobs == 0) return;
        std::vector<CombineJob> dj(n_jobs);
        std::vector<CombineView> dv;
        std::vector<double> src_pts, src_cl;
        std::vector<std::pair<const double*, int64_t>> seen;  // every source edge is uploaded once
        int64_t total = 0;
        for (size_t k = 0; k < n_jobs; ++k) {
            const tm_combine_job& j = jobs[k];
            if (!j.views || j.n_views < 2 || j.n_views > 16 || !j.points || !j.clustering) TM_THROW(TM_ERR_INVALID_ARGUMENT, "combine %zu: needs 2..16 views and output arrays", k);
            CombineJob c{};
            c.out_off = total; c.view_begin = int32_t(dv.size()); c.n_views = int32_t(j.n_views);
            int64_t n = 0;
            for (size_t v = 0; v < j.n_views; ++v) {
                const tm_edge_view& w = j.views[v];
                if (!w.points || !w.clustering || w.start >= w.n || w.end >= w.n || w.n > 0x7fffffffull) TM_THROW(TM_ERR_INVALID_ARGUMENT, "combine %zu: view %zu out of range", k, v);
                if (v > 0) {  // discrete.zig:43-56: end point of the previous view == start point of this one within 1e-10
                    const tm_edge_view& q = j.views[v - 1];
                    for (int d = 0; d < 2; ++d)
                        if (!(std::fabs(q.points[2 * q.end + d] - w.points[2 * w.start + d]) <= 1e-10))
                            TM_THROW(TM_ERR_INVALID_ARGUMENT, "combine %zu: edges %zu and %zu cannot be combined as end points do not match (discrete.zig:43-56)", k, v, v + 1);
                }
                int64_t off = -1;
                for (const auto& pr : seen) if (pr.first == w.points) off = pr.second;
                if (off < 0) {
                    off = int64_t(src_cl.size());
                    src_pts.insert(src_pts.end(), w.points, w.points + 2 * w.n);
                    src_cl.insert(src_cl.end(), w.clustering, w.clustering + w.n);
                    seen.push_back({w.points, off});
                }
                dv.push_back(CombineView{off, int32_t(w.start), int32_t(w.end)});
                n += int64_t(w.start > w.end ? w.start - w.end : w.end - w.start) + 1;
            }
            c.n = int32_t(n - int64_t(j.n_views - 1));
            total += c.n;
            dj[k] = c;
        }
        require_device(device);
        cudaStream_t s = nullptr;
        CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        try {
            DevBuf<CombineJob> d_jobs;
            DevBuf<CombineView> d_views;
            DevBuf<double> d_src_pts, d_src_cl, d_cl;
            DevBuf<double2> d_pts;
            d_jobs.upload(dj, s); d_views.upload(dv, s); d_src_pts.upload(src_pts, s); d_src_cl.upload(src_cl, s);
            d_pts.alloc(size_t(total)); d_cl.alloc(size_t(total));
            LAUNCH(edge_combine_kernel, unsigned(n_jobs), 128, s, (const CombineJob*)d_jobs.p, (const CombineView*)d_views.p, (const double2*)d_src_pts.p,
                   (const double*)d_src_cl.p, d_pts.p, d_cl.p);
            copy_out_runs(n_jobs, (const double2*)d_pts.p, [&](size_t k) { return reinterpret_cast<double2*>(jobs[k].points); }, [&](size_t k) { return dj[k].out_off; },
                          [&](size_t k) { return size_t(dj[k].n); }, s);
            copy_out_runs(n_jobs, (const double*)d_cl.p, [&](size_t k) { return jobs[k].clustering; }, [&](size_t k) { return dj[k].out_off; },
                          [&](size_t k) { return size_t(dj[k].n); }, s);
            CUDA_TRY(cudaStreamSynchronize(s));
        } catch (...) {
            cudaStreamDestroy(s);
            throw;
        }
        cudaStreamDestroy(s);
    });
}

int tm_edges_project_normal(const tm_project_job* jobs, size_t n_jobs, int device) {
    return guarded([&] {
        if (n_jobs && !jobs) TM_THROW(TM_ERR_INVALID_ARGUMENT, "jobs is NULL");
        if (n_jobs == 0) return;
        std::vector<ProjectJob> dj(n_jobs);
        std::vector<double> in;
        int64_t total = 0;
        for (size_t k = 0; k < n_jobs; ++k) {
            const tm_project_job& j = jobs[k];
            if (!j.points || !j.out || j.n < 2 || j.n > 0x7fffffffull) TM_THROW(TM_ERR_INVALID_ARGUMENT, "projectNormal %zu: needs n >= 2 and input / output arrays", k);
            dj[k] = ProjectJob{total, int32_t(j.n), 0, j.distance};
            in.insert(in.end(), j.points, j.points + 2 * j.n);
            total += int64_t(j.n);
        }
        require_device(device);
        cudaStream_t s = nullptr;
        CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        try {
            DevBuf<ProjectJob> d_jobs;
            DevBuf<double> d_in;
            DevBuf<double2> d_out;
            d_jobs.upload(dj, s); d_in.upload(in, s);
            d_out.alloc(size_t(total));
            LAUNCH(project_normal_kernel, unsigned(n_jobs), 128, s, (const ProjectJob*)d_jobs.p, (const double2*)d_in.p, d_out.p);
            copy_out_runs(n_jobs, (const double2*)d_out.p, [&](size_t k) { return reinterpret_cast<double2*>(jobs[k].out); }, [&](size_t k) { return dj[k].off; },
                          [&](size_t k) { return size_t(dj[k].n); }, s);
            CUDA_TRY(cudaStreamSynchronize(s));
        } catch (...) {
            cudaStreamDestroy(s);
            throw;
        }
        cudaStreamDestroy(s);
    });
}

int tm_mg_plan(const tm_block* blocks, size_t n_blocks, const tm_connection* connections, size_t n_connections, const tm_condition* conditions,
               size_t n_conditions, const double* cell_size, size_t max_levels, uint64_t* n_levels, uint64_t* sizes) {
    return guarded([&] {
        if (!blocks || !n_levels) TM_THROW(TM_ERR_INVALID_ARGUMENT, "blocks / n_levels is NULL");
        if ((n_connections && !connections) || (n_conditions && !conditions)) TM_THROW(TM_ERR_INVALID_ARGUMENT, "NULL connection / condition array");
        Topology T;  // validates the fine topology exactly like tm_mesh_create
        T.build(blocks, n_blocks, connections, n_connections, conditions, n_conditions);
        std::vector<tm_block> b(blocks, blocks + n_blocks);
        std::vector<tm_connection> c(connections, connections + n_connections);
        std::vector<tm_condition> k(conditions, conditions + n_conditions);
        std::vector<double> h;
        if (cell_size) h.assign(cell_size, cell_size + 2 * n_blocks);
        const std::vector<MgPlanLevel> plan = plan_multigrid(b, c, k, h);
        *n_levels = plan.size();
        if (sizes)
            for (size_t l = 0; l < plan.size() && l < max_levels; ++l)
                for (size_t q = 0; q < n_blocks; ++q) {
                    sizes[(l * n_blocks + q) * 2] = plan[l].blocks[q].ni;
                    sizes[(l * n_blocks + q) * 2 + 1] = plan[l].blocks[q].nj;
                }
    });
}

int tm_smooth_stream_plan(uint64_t ni, uint64_t nj, uint64_t sweeps, uint64_t* n_chunks, uint64_t* window_rows, uint64_t* window_first,
                          uint64_t* owned_first) {
    return guarded([&] {
        if (!n_chunks || !window_rows) TM_THROW(TM_ERR_INVALID_ARGUMENT, "n_chunks / window_rows is NULL");
        *n_chunks = 0; *window_rows = 0;
        if (ni < 3 || nj < 3 || ni >= (uint64_t(1) << 31) || nj >= (uint64_t(1) << 31) || sweeps > (uint64_t(1) << 40)) return;
        StreamPlan P;
        if (!plan_streaming(int64_t(ni), int64_t(nj), int64_t(sweeps), P)) return;
        *n_chunks = uint64_t(P.K()); *window_rows = uint64_t(P.W);
        for (int k = 0; k < P.K(); ++k) {
            if (window_first) window_first[k] = uint64_t(P.w0[size_t(k)]);
            if (owned_first) owned_first[k] = uint64_t(P.o0[size_t(k)]);
        }
        if (owned_first) owned_first[P.K()] = ni;
    });
}

int tm_tfi_block(uint64_t ni, uint64_t nj, const double* x_i_min, const double* x_i_max, const double* x_j_min, const double* x_j_max, const double* s1,
                 const double* s2, const double* t1, const double* t2, double* out_xy) {
    int rc = guarded([&] {
        if (!out_xy) TM_THROW(TM_ERR_INVALID_ARGUMENT, "out_xy is NULL");
        if (ni < 3 || nj < 3) TM_THROW(TM_ERR_INVALID_ARGUMENT, "tfi: block smaller than 3x3 nodes");
    });
    if (rc != TM_OK) return rc;
    // No device mesh is needed for a single TFI: edges up (one packed buffer), kernel, block down, on a stream of this
    // call.  Both device buffers come from the allocation cache, so a call costs its copies and little else.
    return guarded([&] {
        tfi_validate_host(ni, nj, x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2);
        if (ni * nj >= (uint64_t(1) << 31)) TM_THROW(TM_ERR_UNSUPPORTED, "tfi: block has 2^31 or more nodes");
        require_device(-1);
        const bool trace = std::getenv("TM_STREAM_TRACE") != nullptr;  // host-side phase times on stderr (tuning aid)
        const auto t_start = std::chrono::steady_clock::now();
        auto ms_since_start = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
        DevBuf<double> edges;
        DevBuf<double2> out;
        struct Stream {  // declared after the buffers: drained and destroyed before they return to the cache
            cudaStream_t s = nullptr;
            ~Stream() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }
        } st;
        CUDA_TRY(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
        edges.alloc(size_t(6 * (ni + nj)));
        out.alloc(size_t(ni * nj));
        // layout: x_i_min[2ni] x_i_max[2ni] x_j_min[2nj] x_j_max[2nj] s1[ni] s2[ni] t1[nj] t2[nj]  (as in tfi_launch)
        const double* src[8] = {x_i_min, x_i_max, x_j_min, x_j_max, s1, s2, t1, t2};
        const size_t cnt[8] = {size_t(2 * ni), size_t(2 * ni), size_t(2 * nj), size_t(2 * nj), size_t(ni), size_t(ni), size_t(nj), size_t(nj)};
        double* d = edges.p;
        const double* dev[8];
        for (int k = 0; k < 8; ++k) {
            CUDA_TRY(cudaMemcpyAsync(d, src[k], cnt[k] * sizeof(double), cudaMemcpyHostToDevice, st.s));
            dev[k] = d;
            d += cnt[k];
        }
        dim3 grid(unsigned((nj + TILE_J - 1) / TILE_J), unsigned((ni + TFI_ROWS - 1) / TFI_ROWS));
        LAUNCH(tfi_kernel, grid, TILE_J, st.s, int(ni), int(nj), (const double2*)dev[0], (const double2*)dev[1], (const double2*)dev[2], (const double2*)dev[3],
               dev[4], dev[5], dev[6], dev[7], out.p);
        const double t_queued = ms_since_start();
        CUDA_TRY(cudaMemcpyAsync(out_xy, out.p, size_t(ni * nj) * sizeof(double2), cudaMemcpyDeviceToHost, st.s));
        CUDA_TRY(cudaStreamSynchronize(st.s));
        if (trace) std::fprintf(stderr, "[tfi] queued %.2f ms, downloaded %.2f ms\n", t_queued, ms_since_start());
    });
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
    bool streamed = false;
    rc = guarded([&] {
        StreamPlan plan;
        if (!can_stream(blocks, n_blocks, n_connections, conditions, n_conditions, opts, plan)) return;
        smooth_streamed(blocks, opts, plan, stats);
        streamed = true;
    });
    if (streamed || rc != TM_OK) return rc;
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
