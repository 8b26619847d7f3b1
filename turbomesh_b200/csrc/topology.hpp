// topology.hpp -- host-side analysis of the multi-block topology for the CUDA smoother.
//
// Reproduces the *behaviour* of the reference's boundary classification and row construction
//   BlockBoundaryPoints.init / initLaplacianPoints      src/core/smoothing/smooth.zig:1234-1529
//   initNonZeroMatrixEntries* / initBoundaryData         src/core/smoothing/smooth.zig:421-921
//   Range / RangeFillMatrixIterator                      src/core/boundary.zig:15-117, smooth.zig:1531-1599
// but emits flat device tables for a matrix-free solver instead of a CSR matrix: every block-boundary node
// ends up as exactly one of
//   fixed      value frozen at its initial coordinate                                   (smooth.zig:790-796)
//   smoothed   9-point Winslow row straddling two blocks (side-0 copy of an interface)   (smooth.zig:994-1105)
//   junction   "laplacian_smoothed" primary copy of a multi-block corner / T-junction    (smooth.zig:813-836)
//   sliding    inlet/outlet node: x Dirichlet, y Neumann                                 (smooth.zig:837-859, 1115-1165)
//   slave      "connected" copy: x = x_root + shift, master chains resolved here         (smooth.zig:804-812, 904-915)
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/turbomesh_gpu.h"

namespace tmesh {

struct Error {
    int code;
    std::string msg;
};
#define TM_THROW(code, ...)                              \
    do {                                                 \
        char _b[512];                                    \
        std::snprintf(_b, sizeof _b, __VA_ARGS__);       \
        throw ::tmesh::Error{(code), std::string(_b)};   \
    } while (0)

enum Kind : uint8_t { K_FIXED = 0, K_SMOOTHED = 1, K_CONNECTED = 2, K_LAPLACIAN = 3, K_SLIDING = 4 }; // smooth.zig:1168-1174

struct BlockInfo {
    int64_t off;  // global id of node (0,0): IndexConverter, smooth.zig:1623-1637
    int64_t ni, nj;
    int64_t bbuf; // first flat boundary id: PointDataBufferIndexConverter, boundary.zig:218-239
};

// ---- device-visible PODs (copied verbatim to the GPU) -------------------------------------------
struct SmoothedRow {       // interior interface node, side-0 copy g0 with partner g1
    int64_t g0;             // the row's own node; its block-0 neighbours are g0 -/+ d0 (along) and + n0 (inward)
    int64_t iN, iNW, iNE;   // block-1 nodes g1+n1, g1-d1+n1, g1+d1+n1 (explicit: they may be ghosts of another rank)
    int32_t d0, n0;         // along (in_connection_direction_shift) / inward (first_internal_point_shift) of side 0
    double px, py;          // periodicity mapping side 0 onto side 1 (0 when not periodic)
    int32_t periodic;       // periodic rows use (P,Q), the others (Q,P): smooth.zig:1040-1041 vs 1082-1083
    int32_t slave_begin, slave_end;  // this rank's copies written by the row's thread
    int32_t n_copies;       // all `connected` copies of the node in the whole mesh (weight of the row in sum dx^2)
};
struct JunctionRow {        // sum_k x_k - n x_self = rhs
    int64_t self;
    int64_t nbr[5];
    double rhs_x, rhs_y;
    int32_t n;
    int32_t slave_begin, slave_end;
    int32_t n_copies;
};
struct SlidingRow {         // x-solve: x_self = rhs_x ; y-solve: ysign*(y_self - y_inner) = rhs_y
    int64_t self, inner;
    double rhs_x, rhs_y;
    int32_t rhs_x_from_initial; // 1: rhs_x is captured from the initial mesh (smooth.zig:853-857)
    int32_t ysign;
    int32_t slave_begin, slave_end;
    int32_t n_copies, _pad;
};
struct SlaveRow {           // x_self = x_root + shift
    int64_t self, root;
    double sx, sy;
};
struct FixedOverride {      // a fixed node whose rhs was overwritten by the periodic loop (smooth.zig:904-915)
    int64_t self;
    double x, y;
};
struct RhsTerm {            // a row of the reference system with a non-zero right-hand side (for ||b||, BiCGStab.zig:289-291)
    int64_t g;              // the row's node; its coordinate is the rhs where from_x / from_y is set
    double cx, cy;          // constant rhs
    int32_t from_x, from_y;
};
struct PairCheck {          // connectionDataCheck, smooth.zig:220-275
    int64_t g0, g1;
    double px, py;
    int32_t conn, point;
};

struct Junction {           // BlockBoundaryPoints.LaplacianPoint, smooth.zig:1219-1232
    struct Copy { int64_t g; double px, py; };
    std::vector<Copy> copies;       // sorted by global id, [0] is the laplacian_smoothed primary
    std::vector<int64_t> stencil;   // sorted; includes the primary itself
    double rhs_x = 0, rhs_y = 0;
};

struct Topology {
    std::vector<BlockInfo> blocks;
    int64_t n_nodes = 0, n_boundary = 0;
    std::vector<uint8_t> kind;            // per flat boundary id
    std::vector<Junction> junctions;
    std::vector<SmoothedRow> smoothed;
    std::vector<JunctionRow> junction_rows;
    std::vector<SlidingRow> sliding;
    std::vector<SlaveRow> slaves;         // every connected node whose root is a free row (any order)
    std::vector<SlaveRow> const_slaves;   // slaves of fixed nodes: constants, applied once
    std::vector<FixedOverride> fixed_overrides;
    std::vector<FixedOverride> connected_rhs; // non-zero rhs of `connected` rows (periodic copies), for ||b||
    std::vector<PairCheck> pairs;
    // White control function groups: pairs of O-grid half blocks (A, B) joined by a connection A:j_min[0..] <-> B:j_min[0..].
    // Default = the reference's hard-coded single group (blocks 0 and 1 with connection 0).
    std::vector<std::pair<int64_t, int64_t>> white_groups;
    bool white_ok = false;
    std::string white_why;
    std::vector<tm_connection> conns_copy;  // kept for set_white_groups
    // connected components of the block graph (blocks joined by connections): independent linear systems -- the cuts of a
    // batch --, numbered in the order of their lowest block
    std::vector<int32_t> comp_of_block;
    int32_t n_comp = 0;
    int64_t min_conn_nodes = 6;             // smooth.zig:631 (lenInternal() > 3); coarse multigrid levels relax this to 3

    // ---- helpers -------------------------------------------------------------------------------
    static int64_t range_len(const tm_range& r) { return r.start > r.end ? int64_t(r.start - r.end) + 1 : int64_t(r.end - r.start) + 1; }

    void walk(const tm_range& r, int64_t& base, int64_t& along, int64_t& inward) const {
        const int64_t ni = blocks[r.block].ni, nj = blocks[r.block].nj;
        switch (r.side) {
            case TM_SIDE_I_MIN: base = int64_t(r.start) * nj; along = nj; inward = 1; break;
            case TM_SIDE_I_MAX: base = int64_t(r.start) * nj + nj - 1; along = nj; inward = -1; break;
            case TM_SIDE_J_MIN: base = int64_t(r.start); along = 1; inward = nj; break;
            default: base = (ni - 1) * nj + int64_t(r.start); along = 1; inward = -nj; break;
        }
        if (r.start > r.end) along = -along;
    }
    size_t block_of(int64_t g) const {  // offsets ascend: binary search (a batch of cuts has thousands of blocks)
        size_t lo = 0, hi = blocks.size() - 1;
        while (lo < hi) {
            const size_t mid = (lo + hi + 1) / 2;
            if (blocks[mid].off <= g) lo = mid; else hi = mid - 1;
        }
        return lo;
    }
    // boundary.zig:248-285; -1 when (block, local) is not on the block boundary
    int64_t bid(size_t block, int64_t local) const {
        const int64_t ni = blocks[block].ni, nj = blocks[block].nj;
        const int64_t i = local / nj, j = local - i * nj;
        int64_t k;
        if (i == 0) k = j;
        else if (i == ni - 1) k = nj + 2 * (ni - 2) + j;
        else if (j == 0) k = nj + (i - 1) * 2;
        else if (j == nj - 1) k = nj - 1 + i * 2;
        else return -1;
        return blocks[block].bbuf + k;
    }
    int64_t bid_of_global(int64_t g) const { const size_t b = block_of(g); return bid(b, g - blocks[b].off); }
    bool range_ok(const tm_range& r) const {
        if (r.block >= blocks.size() || r.side > 3) return false;
        const int64_t ext = (r.side == TM_SIDE_I_MIN || r.side == TM_SIDE_I_MAX) ? blocks[r.block].ni : blocks[r.block].nj;
        return r.start < uint64_t(ext) && r.end < uint64_t(ext);  // unsigned: values >= 2^63 must not pass as negative numbers
    }

    // ---- construction --------------------------------------------------------------------------
    void build(const tm_block* blk, size_t nb, const tm_connection* conns, size_t nc, const tm_condition* bcs, size_t nbc) {
        if (nb == 0 || !blk) TM_THROW(TM_ERR_INVALID_ARGUMENT, "mesh without blocks");
        blocks.resize(nb);
        for (size_t b = 0; b < nb; ++b) {
            if (blk[b].ni < 3 || blk[b].nj < 3) TM_THROW(TM_ERR_INVALID_ARGUMENT, "block %zu is smaller than 3x3 nodes", b);
            if (blk[b].ni * blk[b].nj >= (uint64_t(1) << 31)) TM_THROW(TM_ERR_UNSUPPORTED, "block %zu has 2^31 or more nodes", b);
            blocks[b] = BlockInfo{n_nodes, int64_t(blk[b].ni), int64_t(blk[b].nj), n_boundary};
            n_nodes += int64_t(blk[b].ni * blk[b].nj);
            n_boundary += 2 * int64_t(blk[b].ni + blk[b].nj - 2);
        }
        for (size_t c = 0; c < nc; ++c) {
            if (!range_ok(conns[c].ranges[0]) || !range_ok(conns[c].ranges[1])) TM_THROW(TM_ERR_TOPOLOGY, "connection %zu: range out of bounds", c);
            if (range_len(conns[c].ranges[0]) != range_len(conns[c].ranges[1])) TM_THROW(TM_ERR_TOPOLOGY, "connection %zu: the two ranges differ in length", c);
        }
        for (size_t c = 0; c < nbc; ++c) {
            if (!range_ok(bcs[c].range) || bcs[c].kind > TM_BC_OUTLET) TM_THROW(TM_ERR_TOPOLOGY, "condition %zu: invalid range or kind", c);
        }
        const bool timing = std::getenv("TM_TOPO_TIMING") != nullptr;
        auto t0 = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) {
            if (!timing) return;
            const auto t1 = std::chrono::steady_clock::now();
            std::fprintf(stderr, "topology %-16s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
            t0 = t1;
        };
        find_components(conns, nc); lap("components");
        find_junctions(conns, nc); lap("junctions");
        classify(conns, nc, bcs, nbc); lap("classify");
        build_rows(conns, nc, bcs, nbc); lap("rows");
        build_pairs(conns, nc); lap("pairs");
        check_white(conns, nc); lap("white");
    }

  private:
    void find_components(const tm_connection* conns, size_t nc) {
        std::vector<int32_t> parent(blocks.size());
        for (size_t b = 0; b < parent.size(); ++b) parent[b] = int32_t(b);
        auto find = [&](int32_t k) { while (parent[size_t(k)] != k) { parent[size_t(k)] = parent[size_t(parent[size_t(k)])]; k = parent[size_t(k)]; } return k; };
        for (size_t c = 0; c < nc; ++c) {
            const int32_t a = find(int32_t(conns[c].ranges[0].block)), b = find(int32_t(conns[c].ranges[1].block));
            if (a != b) parent[size_t(std::max(a, b))] = std::min(a, b);   // the root is the lowest block of the component
        }
        comp_of_block.assign(blocks.size(), -1);
        n_comp = 0;
        for (size_t b = 0; b < blocks.size(); ++b) {
            const int32_t root = find(int32_t(b));
            if (comp_of_block[size_t(root)] < 0) comp_of_block[size_t(root)] = n_comp++;   // roots are met first (root <= b)
            comp_of_block[b] = comp_of_block[size_t(root)];
        }
    }

    static void periodicity_of(const tm_connection& c, double& px, double& py) {
        px = c.has_periodicity ? c.periodicity[0] : 0.0;
        py = c.has_periodicity ? c.periodicity[1] : 0.0;
    }
    static void append_if_unique(Junction& jn, int64_t g, double px, double py) { // smooth.zig:1516-1522
        for (const auto& cp : jn.copies) if (cp.g == g) return;
        if (jn.copies.size() >= 4) TM_THROW(TM_ERR_UNSUPPORTED, "junction with more than 4 overlapping points (smooth.zig:1221)");
        jn.copies.push_back({g, px, py});
    }

    // initLaplacianPoints, smooth.zig:1340-1514.  The pair scan and the periodicity bookkeeping are
    // order dependent; they are reproduced step by step so that degenerate inputs classify identically.
    void find_junctions(const tm_connection* conns, size_t nc) {
        junctions.clear();
        if (nc == 0) return; // the reference underflows here (smooth.zig:1364); a block without connections has no junctions
        std::vector<int64_t> ids(4 * nc);
        for (size_t c = 0; c < nc; ++c) { // smooth.zig:1349-1356 with Range.endpoints (boundary.zig:65-77)
            for (int s = 0; s < 2; ++s) {
                const tm_range& r = conns[c].ranges[s];
                tm_range first = r, last = r;
                first.end = r.start; last.start = r.end;
                int64_t b0, a0, n0, b1, a1, n1;
                walk(first, b0, a0, n0); walk(last, b1, a1, n1);
                ids[4 * c + s] = blocks[r.block].off + b0;
                ids[4 * c + 2 + s] = blocks[r.block].off + b1;
            }
        }
        const size_t ne = ids.size();
        // The reference scans all pairs a < b with ids[a] == ids[b] in lexicographic order (O(ne^2)) and, per pair, all
        // junctions found so far.  Same visiting order here, but the candidates come from an index: positions grouped by
        // id, junctions looked up by the node ids they contain (a batch of cuts has ~1e5 end points).
        std::vector<size_t> order(ne);
        for (size_t k = 0; k < ne; ++k) order[k] = k;
        std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return ids[x] != ids[y] ? ids[x] < ids[y] : x < y; });
        std::vector<size_t> rank_of(ne);  // position of k inside `order`
        for (size_t k = 0; k < ne; ++k) rank_of[order[k]] = k;
        std::unordered_map<int64_t, std::vector<size_t>> junctions_with;  // node id -> junctions holding it, ascending
        auto note_copy = [&](size_t j, int64_t g) {  // sorted, unique
            auto& v = junctions_with[g];
            const auto it = std::lower_bound(v.begin(), v.end(), j);
            if (it == v.end() || *it != j) v.insert(it, j);
        };
        for (size_t a = 0; a + 1 < ne; ++a) {
            for (size_t q = rank_of[a] + 1; q < ne && ids[order[q]] == ids[a]; ++q) {
                const size_t b = order[q];  // ascending b > a with the same id
                bool found = false;
                const auto hit = junctions_with.find(ids[a]);
                if (hit != junctions_with.end()) {
                    const std::vector<size_t> holders = hit->second;  // copy: appending below may touch the map
                    for (size_t j : holders) {
                        Junction& jn = junctions[j];
                        const size_t n_before = jn.copies.size(); // the reference iterates a slice taken before appending
                        for (size_t k = 0; k < n_before; ++k) {
                            if (jn.copies[k].g != ids[a]) continue;
                            found = true;
                            const size_t add = (b % 2 == 0) ? b + 1 : b - 1; // smooth.zig:1378
                            double px, py; periodicity_of(conns[add / 4], px, py);
                            append_if_unique(jn, ids[add], px, py);
                            note_copy(j, ids[add]);
                        }
                    }
                }
                if (found) continue;
                const size_t p0 = a / 2, p1 = b / 2; // smooth.zig:1390
                if (p0 == p1) TM_THROW(TM_ERR_TOPOLOGY, "connection %zu joins a node with itself", a / 4);
                Junction jn;
                double px, py;
                periodicity_of(conns[p0 / 2], px, py);
                jn.copies.push_back({ids[2 * p0], 0.0, 0.0});
                jn.copies.push_back({ids[2 * p0 + 1], px, py});
                if (jn.copies[0].g == jn.copies[1].g) TM_THROW(TM_ERR_TOPOLOGY, "connection %zu joins a node with itself", p0 / 2);
                periodicity_of(conns[p1 / 2], px, py);
                append_if_unique(jn, ids[2 * p1], px, py);
                append_if_unique(jn, ids[2 * p1 + 1], px, py);
                for (const auto& cp : jn.copies) note_copy(junctions.size(), cp.g);
                junctions.push_back(std::move(jn));
            }
        }
        for (auto& jn : junctions) // smooth.zig:1441-1448
            std::stable_sort(jn.copies.begin(), jn.copies.end(), [](const Junction::Copy& x, const Junction::Copy& y) { return x.g < y.g; });
        std::stable_sort(junctions.begin(), junctions.end(), [](const Junction& x, const Junction& y) { return x.copies[0].g < y.copies[0].g; });
        for (auto& jn : junctions) { // smooth.zig:1457-1511
            jn.stencil.push_back(jn.copies[0].g);
            for (const auto& cp : jn.copies) {
                const size_t blk = block_of(cp.g);
                const int64_t ni = blocks[blk].ni, nj = blocks[blk].nj, local = cp.g - blocks[blk].off;
                const int64_t i = local / nj, j = local - i * nj;
                int64_t pi[2], pj[2];
                int np = 0;
                auto add = [&](int64_t ii, int64_t jj) { pi[np] = ii; pj[np] = jj; ++np; };
                if (i == 0) {
                    if (j == 0) add(1, 1);
                    else if (j == nj - 1) add(1, nj - 2);
                    else { add(1, j - 1); add(1, j + 1); }
                } else if (i == ni - 1) {
                    if (j == 0) add(ni - 2, 1);
                    else if (j == nj - 1) add(ni - 2, nj - 2);
                    else { add(ni - 2, j - 1); add(ni - 2, j + 1); }
                } else if (j == 0) { add(i - 1, 1); add(i + 1, 1); }
                else if (j == nj - 1) { add(i - 1, j - 1); add(i + 1, j - 1); }
                else TM_THROW(TM_ERR_TOPOLOGY, "junction copy %lld is not a block-boundary node", (long long)cp.g);
                for (int q = 0; q < np; ++q) {
                    if (jn.stencil.size() >= 6) TM_THROW(TM_ERR_UNSUPPORTED, "junction stencil with more than 6 ids (smooth.zig:1224)");
                    jn.stencil.push_back(blocks[blk].off + pi[q] * nj + pj[q]);
                    jn.rhs_x += cp.px; jn.rhs_y += cp.py;
                }
            }
            std::sort(jn.stencil.begin(), jn.stencil.end());
        }
    }

    // BlockBoundaryPoints.init, smooth.zig:1243-1329: fixed -> junction groups -> sliding -> per connection
    void classify(const tm_connection* conns, size_t nc, const tm_condition* bcs, size_t nbc) {
        kind.assign(size_t(n_boundary), K_FIXED);
        for (const auto& jn : junctions)
            for (size_t k = 0; k < jn.copies.size(); ++k) {
                const int64_t id = bid_of_global(jn.copies[k].g);
                if (id < 0) TM_THROW(TM_ERR_TOPOLOGY, "junction point is not a block-boundary node");
                kind[size_t(id)] = k == 0 ? K_LAPLACIAN : K_CONNECTED;
            }
        for (size_t c = 0; c < nbc; ++c) {
            if (bcs[c].kind == TM_BC_WALL) continue; // smooth.zig:1275
            int64_t base, along, inward;
            walk(bcs[c].range, base, along, inward);
            for (int64_t k = 0, n = range_len(bcs[c].range); k < n; ++k) kind[size_t(bid(bcs[c].range.block, base + k * along))] = K_SLIDING;
        }
        for (size_t c = 0; c < nc; ++c) {
            int64_t b0, a0, n0, b1, a1, n1;
            walk(conns[c].ranges[0], b0, a0, n0);
            walk(conns[c].ranges[1], b1, a1, n1);
            const int64_t n = range_len(conns[c].ranges[0]);
            for (int64_t k = 0; k < n; ++k) {
                const int64_t i0 = bid(conns[c].ranges[0].block, b0 + k * a0), i1 = bid(conns[c].ranges[1].block, b1 + k * a1);
                if (k == 0 || k == n - 1) {
                    if (kind[size_t(i0)] == K_FIXED || kind[size_t(i0)] == K_SLIDING) kind[size_t(i1)] = K_CONNECTED;
                } else {
                    kind[size_t(i0)] = K_SMOOTHED;
                    kind[size_t(i1)] = K_CONNECTED;
                }
            }
        }
    }

    // Row construction in the reference's order (smooth.zig:723-778, 867-921), then elimination of the
    // `connected` rows (x_master - x_self = rhs) by following master chains to a non-connected root.
    void build_rows(const tm_connection* conns, size_t nc, const tm_condition* bcs, size_t nbc) {
        const size_t nbn = size_t(n_boundary);
        std::vector<int64_t> master(nbn, -1);         // per flat boundary id of a connected node: global id of its master
        std::vector<double> rhs(2 * nbn, 0.0);        // rhs of constant rows (connected / sliding-y / fixed overrides)
        std::vector<uint8_t> rhs_over(nbn, 0);        // rhs overwritten by the periodic loop
        std::vector<int32_t> smoothed_of(nbn, -1), sliding_of(nbn, -1);

        // junction copies are tied to the primary (smooth.zig:738-747)
        for (size_t l = 0; l < junctions.size(); ++l) {
            const auto& jn = junctions[l];
            for (size_t k = 1; k < jn.copies.size(); ++k) {
                const int64_t id = bid_of_global(jn.copies[k].g);
                if (kind[size_t(id)] != K_CONNECTED) TM_THROW(TM_ERR_TOPOLOGY, "junction copy %lld was re-classified (inconsistent topology)", (long long)jn.copies[k].g);
                master[size_t(id)] = jn.copies[0].g;
            }
        }
        // connections (smooth.zig:618-721)
        for (size_t c = 0; c < nc; ++c) {
            const tm_connection& cn = conns[c];
            if (!(cn.ranges[0].block <= cn.ranges[1].block)) TM_THROW(TM_ERR_TOPOLOGY, "connection %zu: ranges[0].block must be <= ranges[1].block (smooth.zig:627)", c);
            if (cn.ranges[0].block == cn.ranges[1].block && !(cn.ranges[0].side == TM_SIDE_I_MIN && cn.ranges[1].side == TM_SIDE_I_MAX))
                TM_THROW(TM_ERR_UNSUPPORTED, "connection %zu: same-block connections are only supported as i_min -> i_max (smooth.zig:522-559)", c);
            const int64_t n = range_len(cn.ranges[0]);
            if (n < min_conn_nodes) TM_THROW(TM_ERR_UNSUPPORTED, "connection %zu: needs at least %lld nodes (smooth.zig:631)", c, (long long)min_conn_nodes);
            int64_t b0, a0, n0, b1, a1, n1;
            walk(cn.ranges[0], b0, a0, n0);
            walk(cn.ranges[1], b1, a1, n1);
            const int64_t off0 = blocks[cn.ranges[0].block].off, off1 = blocks[cn.ranges[1].block].off;
            double px, py; periodicity_of(cn, px, py);
            for (int64_t k = 0; k < n; ++k) {
                const int64_t l0 = b0 + k * a0, l1 = b1 + k * a1, g0 = off0 + l0, g1 = off1 + l1;
                const int64_t i0 = bid(cn.ranges[0].block, l0), i1 = bid(cn.ranges[1].block, l1);
                if (k == 0 || k == n - 1) { // initNonZeroMatrixForConnectionEndpoint, smooth.zig:695-721
                    switch (kind[size_t(i0)]) {
                        case K_FIXED: case K_SLIDING:
                            if (!(g0 < g1)) TM_THROW(TM_ERR_TOPOLOGY, "connection %zu: side-0 ids must be below side-1 ids (smooth.zig:712)", c);
                            master[size_t(i1)] = g0;
                            break;
                        case K_LAPLACIAN: case K_CONNECTED: break;
                        default: TM_THROW(TM_ERR_UNSUPPORTED, "connection %zu ends inside another connection (smooth.zig:719)", c);
                    }
                } else {
                    if (!(g0 < g1)) TM_THROW(TM_ERR_TOPOLOGY, "connection %zu: side-0 ids must be below side-1 ids (smooth.zig:651)", c);
                    if (kind[size_t(i1)] != K_CONNECTED || kind[size_t(i0)] != K_SMOOTHED)
                        TM_THROW(TM_ERR_TOPOLOGY, "connection %zu overlaps another connection", c);
                    master[size_t(i1)] = g0;
                    SmoothedRow row{};
                    row.g0 = g0;
                    row.iN = g1 + n1; row.iNW = g1 - a1 + n1; row.iNE = g1 + a1 + n1;
                    row.d0 = int32_t(a0); row.n0 = int32_t(n0);
                    row.px = px; row.py = py; row.periodic = cn.has_periodicity ? 1 : 0;
                    if (smoothed_of[size_t(i0)] >= 0) smoothed[size_t(smoothed_of[size_t(i0)])] = row; // later connection overwrites
                    else { smoothed_of[size_t(i0)] = int32_t(smoothed.size()); smoothed.push_back(row); }
                }
            }
        }
        // sliding rows (smooth.zig:751-777, 837-859)
        for (size_t c = 0; c < nbc; ++c) {
            if (bcs[c].kind == TM_BC_WALL) continue; // the reference hits `unreachable` here (smooth.zig:775); walls stay fixed
            int64_t base, along, inward;
            walk(bcs[c].range, base, along, inward);
            const int64_t off = blocks[bcs[c].range.block].off;
            for (int64_t k = 0, n = range_len(bcs[c].range); k < n; ++k) {
                const int64_t local = base + k * along, id = bid(bcs[c].range.block, local);
                if (kind[size_t(id)] != K_SLIDING) continue;
                SlidingRow row{};
                row.self = off + local; row.inner = off + local + inward;
                row.rhs_x = 0; row.rhs_y = 0; row.rhs_x_from_initial = 1; row.ysign = inward > 0 ? 1 : -1;
                if (sliding_of[size_t(id)] >= 0) sliding[size_t(sliding_of[size_t(id)])] = row;
                else { sliding_of[size_t(id)] = int32_t(sliding.size()); sliding.push_back(row); }
            }
        }
        for (size_t id = 0; id < nbn; ++id)
            if (kind[id] == K_SLIDING && sliding_of[id] < 0) TM_THROW(TM_ERR_TOPOLOGY, "internal: sliding node without a row");
        // periodic right-hand sides: every side-1 node of a periodic connection (smooth.zig:903-915)
        for (size_t c = 0; c < nc; ++c) {
            if (!conns[c].has_periodicity) continue;
            int64_t b1, a1, n1;
            walk(conns[c].ranges[1], b1, a1, n1);
            for (int64_t k = 0, n = range_len(conns[c].ranges[1]); k < n; ++k) {
                const int64_t id = bid(conns[c].ranges[1].block, b1 + k * a1);
                rhs[2 * size_t(id)] = -conns[c].periodicity[0];
                rhs[2 * size_t(id) + 1] = -conns[c].periodicity[1];
                rhs_over[size_t(id)] = 1;
            }
        }
        // junction rows; their rhs is set last (smooth.zig:917-920)
        for (const auto& jn : junctions) {
            JunctionRow row{};
            row.self = jn.copies[0].g;
            row.n = 0;
            for (int64_t g : jn.stencil) if (g != row.self) row.nbr[row.n++] = g;
            if (row.n < 1) TM_THROW(TM_ERR_TOPOLOGY, "junction %lld without neighbours", (long long)row.self);
            row.rhs_x = jn.rhs_x; row.rhs_y = jn.rhs_y;
            junction_rows.push_back(row);
        }
        // per-kind consequences of an overwritten rhs
        for (size_t b = 0; b < blocks.size(); ++b) {
            const int64_t nbb = 2 * (blocks[b].ni + blocks[b].nj - 2);
            for (int64_t q = 0; q < nbb; ++q) {
                const size_t id = size_t(blocks[b].bbuf + q);
                if (!rhs_over[id]) continue;
                const int64_t g = blocks[b].off + local_of_bid(b, q);
                if (kind[id] == K_FIXED) fixed_overrides.push_back({g, rhs[2 * id], rhs[2 * id + 1]});
                else if (kind[id] == K_CONNECTED) connected_rhs.push_back({g, rhs[2 * id], rhs[2 * id + 1]});
                else if (kind[id] == K_SLIDING) {
                    auto& row = sliding[size_t(sliding_of[id])];
                    row.rhs_x = rhs[2 * id]; row.rhs_y = rhs[2 * id + 1]; row.rhs_x_from_initial = 0;
                }
            }
        }
        // eliminate connected rows: x_self = x_master - rhs_self, chains followed to the root
        std::vector<int32_t> copies(nbn, 0);  // per flat boundary id of a free root: copies hanging off it
        for (size_t b = 0; b < blocks.size(); ++b) {
            const int64_t nbb = 2 * (blocks[b].ni + blocks[b].nj - 2);
            for (int64_t q = 0; q < nbb; ++q) {
                const size_t id = size_t(blocks[b].bbuf + q);
                if (kind[id] != K_CONNECTED) continue;
                const int64_t self = blocks[b].off + local_of_bid(b, q);
                double sx = 0, sy = 0;
                size_t cur = id;
                int64_t root = -1;
                for (int hops = 0;; ++hops) {
                    if (hops > 64) TM_THROW(TM_ERR_TOPOLOGY, "cyclic `connected` chain at node %lld", (long long)self);
                    if (master[cur] < 0) TM_THROW(TM_ERR_TOPOLOGY, "connected node %lld has no master (inconsistent topology)", (long long)self);
                    sx -= rhs[2 * cur]; sy -= rhs[2 * cur + 1];
                    root = master[cur];
                    const int64_t rid = bid_of_global(root);
                    if (rid < 0) TM_THROW(TM_ERR_TOPOLOGY, "master of node %lld is not a boundary node", (long long)self);
                    cur = size_t(rid);
                    if (kind[cur] != K_CONNECTED) break;
                }
                if (kind[cur] == K_FIXED) const_slaves.push_back({self, root, sx, sy});
                else { slaves.push_back({self, root, sx, sy}); copies[cur] += 1; }  // cur = the root's boundary id
            }
        }
        // per root: how many copies hang off it anywhere in the mesh
        auto copies_of = [&](int64_t g) { return copies[size_t(bid_of_global(g))]; };
        for (auto& r : smoothed) { r.slave_begin = r.slave_end = 0; r.n_copies = copies_of(r.g0); }
        for (auto& r : junction_rows) { r.slave_begin = r.slave_end = 0; r.n_copies = copies_of(r.self); }
        for (auto& r : sliding) { r.slave_begin = r.slave_end = 0; r.n_copies = copies_of(r.self); r._pad = 0; }
    }

    int64_t local_of_bid(size_t block, int64_t q) const { // inverse of bid() within a block
        const int64_t ni = blocks[block].ni, nj = blocks[block].nj;
        if (q < nj) return q;                                   // j_min line (i = 0)
        if (q >= nj + 2 * (ni - 2)) return (ni - 1) * nj + (q - (nj + 2 * (ni - 2))); // j_max line
        const int64_t r = q - nj, i = r / 2 + 1;
        return (r % 2 == 0) ? i * nj : i * nj + nj - 1;         // i_min / i_max
    }

    void build_pairs(const tm_connection* conns, size_t nc) {
        for (size_t c = 0; c < nc; ++c) {
            int64_t b0, a0, n0, b1, a1, n1;
            walk(conns[c].ranges[0], b0, a0, n0);
            walk(conns[c].ranges[1], b1, a1, n1);
            double px, py; periodicity_of(conns[c], px, py);
            for (int64_t k = 0, n = range_len(conns[c].ranges[0]); k < n; ++k)
                pairs.push_back({blocks[conns[c].ranges[0].block].off + b0 + k * a0, blocks[conns[c].ranges[1].block].off + b1 + k * a1, px, py, int32_t(c), int32_t(k)});
        }
    }

    static bool le_connection(const tm_connection& c, int64_t a, int64_t b) {  // wall_control_function.zig:212-217
        return int64_t(c.ranges[0].block) == a && c.ranges[0].start == 0 && c.ranges[0].side == TM_SIDE_J_MIN && int64_t(c.ranges[1].block) == b &&
               c.ranges[1].start == 0 && c.ranges[1].side == TM_SIDE_J_MIN && !c.has_periodicity && c.ranges[0].end > 1;
    }

    void check_white(const tm_connection* conns, size_t nc) {  // wall_control_function.zig:72, 204-217
        conns_copy.assign(conns, conns + nc);
        white_ok = false;
        white_groups.clear();
        if (blocks.size() < 2 || nc < 1) { white_why = "the White control function needs blocks 0 and 1 (O-grid halves) and connection 0"; return; }
        if (!le_connection(conns[0], 0, 1)) {
            white_why = "White: connection 0 must be block0:j_min[0..] <-> block1:j_min[0..], non periodic (wall_control_function.zig:212-217)";
            return;
        }
        white_groups.push_back({0, 1});
        white_ok = true;
    }

  public:
    // Extension for batches of cuts: several (A, B) groups, each needing its own leading-edge connection.
    void set_white_groups(const uint64_t* pairs, size_t n) {
        std::vector<std::pair<int64_t, int64_t>> g;
        for (size_t k = 0; k < n; ++k) {
            const int64_t a = int64_t(pairs[2 * k]), b = int64_t(pairs[2 * k + 1]);
            if (a < 0 || b < 0 || size_t(a) >= blocks.size() || size_t(b) >= blocks.size() || a == b)
                TM_THROW(TM_ERR_INVALID_ARGUMENT, "white group %zu: invalid block pair", k);
            bool found = false;
            for (const auto& c : conns_copy) found = found || le_connection(c, a, b);
            if (!found) TM_THROW(TM_ERR_UNSUPPORTED, "white group %zu: no connection block%lld:j_min[0..] <-> block%lld:j_min[0..] (non periodic)", k, (long long)a, (long long)b);
            if (blocks[size_t(a)].ni < 3 || blocks[size_t(b)].ni < 3 || blocks[size_t(a)].nj < 3 || blocks[size_t(b)].nj < 3)
                TM_THROW(TM_ERR_UNSUPPORTED, "white group %zu: blocks too small", k);
            g.push_back({a, b});
        }
        white_groups = std::move(g);
        white_ok = !white_groups.empty();
        if (!white_ok) white_why = "no White groups set";
    }
};

} // namespace tmesh
