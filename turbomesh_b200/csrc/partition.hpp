// partition.hpp -- host-side localisation of the global topology for one rank (one process per GPU).
//
// The reference has no distributed runtime; this is the B200-native sharding of its one global
// system: whole blocks are assigned to ranks, every rank keeps its own blocks plus *ghost* copies of the few remote
// nodes its rows read -- per cross-rank connection the partner's first interior line (for the 9-point interface row,
// smooth.zig:994-1105) and the interface line itself (for the `connected` copies, smooth.zig:804-812), plus the
// diagonal neighbours of junction rows (smooth.zig:1457-1511).  Once per sweep / operator application each rank packs
// the owned nodes its peers ghost and receives its ghosts straight into the tail of the field (no unpack).
//
// A remote node that is itself a `connected` copy is never received: the rank keeps a *synthesised* slot for it and
// derives it from the copy's root (received or owned) exactly like its own copies, so every copy of a node is updated
// at the same point of the algorithm as on a single GPU.
//
// Every rank builds the same global Topology and derives the lists of ALL ranks from it, so the send list of rank a
// towards b is by construction the ghost list of b from a (both sorted by global id) -- no handshake is needed.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <map>
#include <vector>

#include "topology.hpp"

namespace tmesh {

struct LocalTables {
    int rank = 0, n_ranks = 1;
    std::vector<int32_t> own_blocks;             // global block ids owned by this rank, ascending
    std::vector<int64_t> loff;                   // per global block: offset of its node (0,0) in the local field, -1 if remote
    int64_t n_own = 0, n_ghost = 0, n_synth = 0, n_check = 0, n_local = 0; // local field = [own][ghosts from rank 0][from rank 1]...[synthesised copies][check ghosts]
    std::vector<int64_t> synth_ids;              // sorted global ids of remote `connected` copies kept as synthesised slots
    std::vector<std::vector<int64_t>> ghost_ids; // per peer: sorted global ids read here, owned there
    std::vector<std::vector<int64_t>> send_ids;  // per peer: sorted global ids owned here, read there
    std::vector<int64_t> ghost_base, send_base;  // per peer (n_ranks+1 entries): offsets in the ghost region / send buffer
    std::vector<int64_t> send_lidx;              // local indices to pack, concatenated over peers
    std::vector<int64_t> peer_ghost_offset;      // per peer p: where this rank's segment starts in p's local field (peer-memory push)
    // one-time exchange of raw coordinates for connectionDataCheck (smooth.zig:220-275) across ranks
    std::vector<std::vector<int64_t>> check_ghost_ids, check_send_ids;
    std::vector<int64_t> check_ghost_base, check_send_base, check_send_lidx;
    // rows owned by this rank, all node references are local indices
    std::vector<SmoothedRow> smoothed;
    std::vector<JunctionRow> junction_rows;
    std::vector<SlidingRow> sliding;
    std::vector<SlaveRow> slaves;                // [0, n_slaves_local_root): root row owned here (written by the root's
    int64_t n_slaves_local_root = 0;             //  thread in sweeps); the rest have ghost roots (synced after exchange)
    std::vector<SlaveRow> const_slaves;
    std::vector<FixedOverride> fixed_overrides;
    std::vector<PairCheck> pairs;                // owned by the rank of the side-1 node
    std::vector<RhsTerm> rhs_terms;
    bool owns_white = false;                     // blocks 0 and 1 (the O-grid halves White works on) live here
};

inline int owner_of_node(const Topology& T, const std::vector<int32_t>& owner, int64_t g) { return owner[T.block_of(g)]; }

struct ReadSets {
    std::vector<std::vector<int64_t>> recv;  // per owning rank: remote nodes received every exchange (sorted, unique)
    std::vector<int64_t> synth;              // remote `connected` copies derived locally from their root (sorted, unique)
};

// copy -> root for the copies whose root is a free row, sorted by copy id (binary-searched)
struct RootTable {
    std::vector<std::pair<int64_t, int64_t>> v;
    explicit RootTable(const Topology& T) {
        v.reserve(T.slaves.size());
        for (const auto& s : T.slaves) v.push_back({s.self, s.root});
        std::sort(v.begin(), v.end());
    }
    bool find(int64_t g, int64_t& root) const {
        const auto it = std::lower_bound(v.begin(), v.end(), std::make_pair(g, std::numeric_limits<int64_t>::min()));
        if (it == v.end() || it->first != g) return false;
        root = it->second;
        return true;
    }
};

// remote nodes read by the rows of EVERY rank, in one pass over the rows (a row is evaluated by the owner of its node)
inline std::vector<ReadSets> read_sets_all(const Topology& T, const std::vector<int32_t>& owner, int n_ranks, const RootTable& root_of) {
    std::vector<ReadSets> out;
    out.resize(size_t(n_ranks));
    for (auto& rs : out) rs.recv.assign(size_t(n_ranks), {});
    auto need = [&](int r, int64_t g) {  // rank r reads node g
        const int o = owner_of_node(T, owner, g);
        if (o == r) return;
        int64_t root;
        if (root_of.find(g, root)) {  // a copy: keep a slot, fetch its root instead
            out[size_t(r)].synth.push_back(g);
            const int ro = owner_of_node(T, owner, root);
            if (ro != r) out[size_t(r)].recv[size_t(ro)].push_back(root);
        } else {
            out[size_t(r)].recv[size_t(o)].push_back(g);
        }
    };
    for (const auto& row : T.smoothed) {
        const int r = owner_of_node(T, owner, row.g0);
        need(r, row.iN); need(r, row.iNW); need(r, row.iNE);
    }
    for (const auto& row : T.junction_rows) {
        const int r = owner_of_node(T, owner, row.self);
        for (int k = 0; k < row.n; ++k) need(r, row.nbr[k]);
    }
    for (const auto* list : {&T.slaves, &T.const_slaves})
        for (const auto& s : *list) need(owner_of_node(T, owner, s.self), s.root);
    for (auto& rs : out) {
        for (auto& v : rs.recv) {
            std::sort(v.begin(), v.end());
            v.erase(std::unique(v.begin(), v.end()), v.end());
        }
        std::sort(rs.synth.begin(), rs.synth.end());
        rs.synth.erase(std::unique(rs.synth.begin(), rs.synth.end()), rs.synth.end());
    }
    return out;
}

// side-0 nodes of interface pairs whose side-1 node belongs to rank r but which live elsewhere (raw coordinates,
// fetched once when smoothing begins): out[r][o] = sorted ids rank r fetches from rank o, for every r in one pass
inline std::vector<std::vector<std::vector<int64_t>>> check_sets_all(const Topology& T, const std::vector<int32_t>& owner, int n_ranks) {
    std::vector<std::vector<std::vector<int64_t>>> out;
    out.assign(size_t(n_ranks), std::vector<std::vector<int64_t>>(size_t(n_ranks)));
    for (const auto& p : T.pairs) {
        const int r = owner_of_node(T, owner, p.g1), o = owner_of_node(T, owner, p.g0);
        if (o != r) out[size_t(r)][size_t(o)].push_back(p.g0);
    }
    for (auto& per_rank : out)
        for (auto& v : per_rank) {
            std::sort(v.begin(), v.end());
            v.erase(std::unique(v.begin(), v.end()), v.end());
        }
    return out;
}

// local index (position in the rank's field) of global node g: owned, synthesised copy or ghost
inline int64_t local_index(const Topology& T, const std::vector<int32_t>& owner, const LocalTables& L, int64_t g) {
    const size_t b = T.block_of(g);
    if (owner[b] == L.rank) return L.loff[b] + (g - T.blocks[b].off);
    {
        const auto it = std::lower_bound(L.synth_ids.begin(), L.synth_ids.end(), g);
        if (it != L.synth_ids.end() && *it == g) return L.n_own + L.n_ghost + int64_t(it - L.synth_ids.begin());
    }
    const auto& v = L.ghost_ids[size_t(owner[b])];
    const auto it = std::lower_bound(v.begin(), v.end(), g);
    if (it == v.end() || *it != g) TM_THROW(TM_ERR_TOPOLOGY, "internal: node %lld is not in the ghost set of rank %d", (long long)g, L.rank);
    return L.n_own + L.ghost_base[size_t(owner[b])] + int64_t(it - v.begin());
}

inline void validate_owner(const Topology& T, const std::vector<int32_t>& owner, int n_ranks) {
    if (owner.size() != T.blocks.size()) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block_owner must have one entry per block");
    for (size_t b = 0; b < owner.size(); ++b)
        if (owner[b] < 0 || owner[b] >= n_ranks) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block %zu: owner %d out of range (%d ranks)", b, owner[b], n_ranks);
}

inline LocalTables localize(const Topology& T, const std::vector<int32_t>& owner, int rank, int n_ranks) {
    validate_owner(T, owner, n_ranks);
    if (rank < 0 || rank >= n_ranks) TM_THROW(TM_ERR_INVALID_ARGUMENT, "rank %d out of range (%d ranks)", rank, n_ranks);
    LocalTables L;
    L.rank = rank; L.n_ranks = n_ranks;
    L.loff.assign(T.blocks.size(), -1);
    for (size_t b = 0; b < T.blocks.size(); ++b) {
        if (owner[b] != rank) continue;
        L.own_blocks.push_back(int32_t(b));
        L.loff[b] = L.n_own;
        L.n_own += T.blocks[b].ni * T.blocks[b].nj;
    }
    const RootTable root_of(T);
    std::vector<ReadSets> all = read_sets_all(T, owner, n_ranks, root_of);
    L.send_ids.assign(size_t(n_ranks), {});
    L.peer_ghost_offset.assign(size_t(n_ranks), 0);
    for (int p = 0; p < n_ranks; ++p) {
        if (p == rank) continue;
        // p's local field is [own nodes of p | ghosts from rank 0 | from rank 1 | ...]: my segment starts after p's own nodes
        // and the ghosts p receives from the ranks below me
        int64_t off = 0;
        for (size_t b = 0; b < T.blocks.size(); ++b)
            if (owner[b] == p) off += T.blocks[b].ni * T.blocks[b].nj;
        for (int q = 0; q < rank; ++q) off += int64_t(all[size_t(p)].recv[size_t(q)].size());
        L.peer_ghost_offset[size_t(p)] = off;
        L.send_ids[size_t(p)] = std::move(all[size_t(p)].recv[size_t(rank)]);
    }
    L.ghost_ids = std::move(all[size_t(rank)].recv);
    L.synth_ids = std::move(all[size_t(rank)].synth);
    L.ghost_base.assign(size_t(n_ranks) + 1, 0);
    L.send_base.assign(size_t(n_ranks) + 1, 0);
    for (int p = 0; p < n_ranks; ++p) {
        L.ghost_base[size_t(p) + 1] = L.ghost_base[size_t(p)] + int64_t(L.ghost_ids[size_t(p)].size());
        L.send_base[size_t(p) + 1] = L.send_base[size_t(p)] + int64_t(L.send_ids[size_t(p)].size());
    }
    L.n_ghost = L.ghost_base[size_t(n_ranks)];
    L.n_synth = int64_t(L.synth_ids.size());
    {
        auto checks = check_sets_all(T, owner, n_ranks);
        L.check_send_ids.assign(size_t(n_ranks), {});
        for (int p = 0; p < n_ranks; ++p)
            if (p != rank) L.check_send_ids[size_t(p)] = std::move(checks[size_t(p)][size_t(rank)]);
        L.check_ghost_ids = std::move(checks[size_t(rank)]);
    }
    L.check_ghost_base.assign(size_t(n_ranks) + 1, 0);
    L.check_send_base.assign(size_t(n_ranks) + 1, 0);
    for (int p = 0; p < n_ranks; ++p) {
        L.check_ghost_base[size_t(p) + 1] = L.check_ghost_base[size_t(p)] + int64_t(L.check_ghost_ids[size_t(p)].size());
        L.check_send_base[size_t(p) + 1] = L.check_send_base[size_t(p)] + int64_t(L.check_send_ids[size_t(p)].size());
    }
    L.n_check = L.check_ghost_base[size_t(n_ranks)];
    L.n_local = L.n_own + L.n_ghost + L.n_synth + L.n_check;

    auto lidx = [&](int64_t g) -> int64_t { return local_index(T, owner, L, g); };
    auto mine = [&](int64_t g) { return owner_of_node(T, owner, g) == rank; };

    for (int p = 0; p < n_ranks; ++p)
        for (int64_t g : L.send_ids[size_t(p)]) L.send_lidx.push_back(lidx(g));
    for (int p = 0; p < n_ranks; ++p)
        for (int64_t g : L.check_send_ids[size_t(p)]) L.check_send_lidx.push_back(lidx(g));
    auto check_idx = [&](int64_t g) -> int64_t {  // raw copy of a remote side-0 node
        const int o = owner_of_node(T, owner, g);
        const auto& v = L.check_ghost_ids[size_t(o)];
        const auto it = std::lower_bound(v.begin(), v.end(), g);
        return L.n_own + L.n_ghost + L.n_synth + L.check_ghost_base[size_t(o)] + int64_t(it - v.begin());
    };

    // slaves owned here: those whose root row is also here are written by the root's thread
    struct Rooted { int64_t root; SlaveRow row; };  // keyed by the root's global id; stable order within a root
    std::vector<Rooted> by_root;
    std::vector<SlaveRow> remote_root;
    for (const auto& s : T.slaves) {
        const bool synth = std::binary_search(L.synth_ids.begin(), L.synth_ids.end(), s.self);
        if (!mine(s.self) && !synth) continue;
        SlaveRow l{lidx(s.self), lidx(s.root), s.sx, s.sy};
        if (mine(s.root)) by_root.push_back(Rooted{s.root, l});
        else remote_root.push_back(l);
    }
    std::stable_sort(by_root.begin(), by_root.end(), [](const Rooted& a, const Rooted& b) { return a.root < b.root; });
    size_t n_attached = 0;
    auto attach = [&](int64_t root_g, int32_t& bgn, int32_t& end) {
        bgn = end = int32_t(L.slaves.size());
        auto it = std::lower_bound(by_root.begin(), by_root.end(), root_g, [](const Rooted& a, int64_t g) { return a.root < g; });
        for (; it != by_root.end() && it->root == root_g; ++it) { L.slaves.push_back(it->row); ++n_attached; }
        end = int32_t(L.slaves.size());
    };
    for (const auto& row : T.smoothed) {
        if (!mine(row.g0)) continue;
        SmoothedRow l = row;
        const int64_t g0 = row.g0;
        l.g0 = lidx(g0); l.iN = lidx(row.iN); l.iNW = lidx(row.iNW); l.iNE = lidx(row.iNE);
        attach(g0, l.slave_begin, l.slave_end);
        L.smoothed.push_back(l);
    }
    for (const auto& row : T.junction_rows) {
        if (!mine(row.self)) continue;
        JunctionRow l = row;
        l.self = lidx(row.self);
        for (int k = 0; k < row.n; ++k) l.nbr[k] = lidx(row.nbr[k]);
        attach(row.self, l.slave_begin, l.slave_end);
        L.junction_rows.push_back(l);
    }
    for (const auto& row : T.sliding) {
        if (!mine(row.self)) continue;
        SlidingRow l = row;
        l.self = lidx(row.self); l.inner = lidx(row.inner);
        attach(row.self, l.slave_begin, l.slave_end);
        L.sliding.push_back(l);
    }
    if (n_attached != by_root.size()) TM_THROW(TM_ERR_TOPOLOGY, "internal: %zu slaves without a root row", by_root.size() - n_attached);
    L.n_slaves_local_root = int64_t(L.slaves.size());
    for (const auto& s : remote_root) L.slaves.push_back(s);

    for (const auto& s : T.const_slaves)
        if (mine(s.self)) L.const_slaves.push_back({lidx(s.self), lidx(s.root), s.sx, s.sy});
    for (const auto& f : T.fixed_overrides)
        if (mine(f.self)) L.fixed_overrides.push_back({lidx(f.self), f.x, f.y});
    for (const auto& p : T.pairs)
        if (mine(p.g1)) L.pairs.push_back({mine(p.g0) ? lidx(p.g0) : check_idx(p.g0), lidx(p.g1), p.px, p.py, p.conn, p.point});

    // rows of the reference system with a non-zero rhs (smooth.zig:780-921), restricted to this rank
    std::vector<uint8_t> over(size_t(T.n_boundary), 0);
    for (const auto& f : T.fixed_overrides) {
        over[size_t(T.bid_of_global(f.self))] = 1;
        if (mine(f.self)) L.rhs_terms.push_back(RhsTerm{lidx(f.self), f.x, f.y, 0, 0});
    }
    for (int32_t b : L.own_blocks) {
        const auto& B = T.blocks[size_t(b)];
        auto visit = [&](int64_t i, int64_t j) {
            const int64_t local = i * B.nj + j;
            const size_t id = size_t(T.bid(size_t(b), local));
            if (T.kind[id] == K_FIXED && !over[id]) L.rhs_terms.push_back(RhsTerm{L.loff[size_t(b)] + local, 0.0, 0.0, 1, 1});
        };
        for (int64_t j = 0; j < B.nj; ++j) { visit(0, j); visit(B.ni - 1, j); }
        for (int64_t i = 1; i + 1 < B.ni; ++i) { visit(i, 0); visit(i, B.nj - 1); }
    }
    for (const auto& s : T.sliding)
        if (mine(s.self)) L.rhs_terms.push_back(RhsTerm{lidx(s.self), s.rhs_x, s.rhs_y, s.rhs_x_from_initial, 0});
    for (const auto& j : T.junction_rows)
        if (mine(j.self)) L.rhs_terms.push_back(RhsTerm{lidx(j.self), j.rhs_x, j.rhs_y, 0, 0});
    for (const auto& c : T.connected_rhs)
        if (mine(c.self)) L.rhs_terms.push_back(RhsTerm{lidx(c.self), c.x, c.y, 0, 0});  // smooth.zig:904-915

    L.owns_white = T.blocks.size() >= 2 && owner[0] == rank && owner[1] == rank;
    return L;
}

}  // namespace tmesh
