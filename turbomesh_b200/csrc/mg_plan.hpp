// mg_plan.hpp -- host-only planning of the multi-block multigrid hierarchy (no device code, testable without a GPU).
//
// Coarse levels are complete multi-block meshes: block sizes and every connection / condition range are halved (nested
// coarsening).  Directions are tied into classes by the connections -- the along and the normal direction of the two
// sides of a connection must coarsen together -- and a class is halved when every extent and every range end point in
// it is even (and long enough), but only while its mean cell size is not much larger than the smallest one
// (semi-coarsening).  A level whose topology the row construction cannot express (Topology::build throws) ends the
// hierarchy.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "topology.hpp"

namespace tmesh {

inline int along_dir(uint32_t side) { return (side == TM_SIDE_I_MIN || side == TM_SIDE_I_MAX) ? 0 : 1; }

struct MgPlanLevel {
    std::vector<tm_block> blocks;          // xy = NULL
    std::vector<tm_connection> conns;
    std::vector<tm_condition> bcs;
    std::vector<int> fi, fj;               // per block: coarsening factors (1 or 2) towards the next level; empty on the coarsest
    Topology topo;                         // levels >= 1 (level 0 is the mesh's own topology)
};

// cell_size: mean cell size per (block, direction), 2 * n_blocks entries (empty: all equal)
inline std::vector<MgPlanLevel> plan_multigrid(const std::vector<tm_block>& blocks, const std::vector<tm_connection>& conns,
                                               const std::vector<tm_condition>& bcs, const std::vector<double>& cell_size, int max_levels = 20) {
    const size_t nb = blocks.size();
    std::vector<int> parent(2 * nb);
    for (size_t k = 0; k < parent.size(); ++k) parent[k] = int(k);
    auto find = [&](int k) { while (parent[size_t(k)] != k) { parent[size_t(k)] = parent[size_t(parent[size_t(k)])]; k = parent[size_t(k)]; } return k; };
    auto unite = [&](int a, int b) { a = find(a); b = find(b); if (a != b) parent[size_t(std::max(a, b))] = std::min(a, b); };
    for (const auto& c : conns) {
        const int b0 = int(c.ranges[0].block), b1 = int(c.ranges[1].block), a0 = along_dir(c.ranges[0].side), a1 = along_dir(c.ranges[1].side);
        unite(2 * b0 + a0, 2 * b1 + a1);
        unite(2 * b0 + 1 - a0, 2 * b1 + 1 - a1);
    }
    std::vector<double> hc(2 * nb, 0.0);  // per class root: mean cell size
    {
        std::vector<int> cnt(2 * nb, 0);
        for (size_t k = 0; k < 2 * nb; ++k) { hc[size_t(find(int(k)))] += cell_size.empty() ? 1.0 : cell_size[k]; cnt[size_t(find(int(k)))] += 1; }
        for (size_t k = 0; k < 2 * nb; ++k) if (cnt[k]) hc[k] /= double(cnt[k]);
    }
    std::vector<MgPlanLevel> levels(1);
    levels[0].blocks = blocks;
    for (auto& b : levels[0].blocks) b.xy = nullptr;
    levels[0].conns = conns;
    levels[0].bcs = bcs;
    for (int level = 0; level + 1 < max_levels; ++level) {
        MgPlanLevel& F = levels.back();
        std::vector<uint8_t> ok(2 * nb, 1);
        auto veto = [&](size_t block, int dir) { ok[size_t(find(int(2 * block) + dir))] = 0; };
        for (size_t b = 0; b < nb; ++b) {
            if ((F.blocks[b].ni - 1) % 2 || (F.blocks[b].ni - 1) / 2 < 2) veto(b, 0);
            if ((F.blocks[b].nj - 1) % 2 || (F.blocks[b].nj - 1) / 2 < 2) veto(b, 1);
        }
        auto check_range = [&](const tm_range& r, uint64_t min_span) {
            const uint64_t span = r.start > r.end ? r.start - r.end : r.end - r.start;
            if (r.start % 2 || r.end % 2 || span < min_span) veto(size_t(r.block), along_dir(r.side));
        };
        for (const auto& c : F.conns) { check_range(c.ranges[0], 4); check_range(c.ranges[1], 4); }
        for (const auto& c : F.bcs) check_range(c.range, 2);
        double h_min = 0.0;
        for (size_t k = 0; k < 2 * nb; ++k)
            if (find(int(k)) == int(k) && ok[k] && (h_min == 0.0 || hc[k] < h_min)) h_min = hc[k];
        if (h_min == 0.0) break;  // nothing can be coarsened any further
        std::vector<uint8_t> go(2 * nb, 0);
        for (size_t k = 0; k < 2 * nb; ++k)
            if (find(int(k)) == int(k) && ok[k] && hc[k] <= h_min / 0.6) go[k] = 1;
        MgPlanLevel C;
        std::vector<int> fi(nb), fj(nb);
        C.blocks.resize(nb);
        for (size_t b = 0; b < nb; ++b) {
            fi[b] = go[size_t(find(int(2 * b)))] ? 2 : 1;
            fj[b] = go[size_t(find(int(2 * b) + 1))] ? 2 : 1;
            C.blocks[b] = tm_block{(F.blocks[b].ni - 1) / uint64_t(fi[b]) + 1, (F.blocks[b].nj - 1) / uint64_t(fj[b]) + 1, nullptr};
        }
        auto coarse_range = [&](tm_range r) {
            const uint64_t f = uint64_t(along_dir(r.side) == 0 ? fi[size_t(r.block)] : fj[size_t(r.block)]);
            r.start /= f; r.end /= f;
            return r;
        };
        C.conns = F.conns;
        for (auto& c : C.conns) { c.ranges[0] = coarse_range(c.ranges[0]); c.ranges[1] = coarse_range(c.ranges[1]); }
        C.bcs = F.bcs;
        for (auto& c : C.bcs) c.range = coarse_range(c.range);
        C.topo.min_conn_nodes = 3;
        try {
            C.topo.build(C.blocks.data(), nb, C.conns.data(), C.conns.size(), C.bcs.data(), C.bcs.size());
        } catch (const Error&) {  // a topology the row construction cannot express at this resolution: stop coarsening here
            break;
        }
        F.fi = fi; F.fj = fj;
        for (size_t k = 0; k < 2 * nb; ++k) if (go[k]) hc[k] *= 2.0;
        levels.push_back(std::move(C));
    }
    return levels;
}

}  // namespace tmesh
