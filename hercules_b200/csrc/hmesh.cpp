// hmesh.cpp -- host-side octree primitives of the per-rank mesher (hercules_b200/octree_local.py).
//
// The reference meshes with octor (C, distributed: octor_refinetree / octor_balancetree / octor_extractmesh,
// octor.c:4337-4700, 5268-6645).  Here a rank works on CHUNKS of coarse cells (edge S = the largest leaf the
// multi-rank bootstrap admits): the balanced refinement inside a chunk depends on the material model within
// one cell of it only, so refining + balancing the chunk with its 26-neighbour ring gives the exact leaves
// of the whole-domain mesh inside the chunk (octree_local.py has the argument).  Each call below handles one
// chunk and holds no global state: the Python side runs them on a thread pool (ctypes drops the GIL).
//
//   hmesh_chunk_leaves  refine by the vs rule (toexpand / vsrule, psolve.c:2185, quake_util.c:215) from a
//                       material grid + 2:1 balance across faces and edges (18 directions, octor.c:4398) by
//                       ripple propagation from the finest level up; returns the leaves inside the chunk
//   hmesh_chunk_nodes   octor_extractmesh's node side (octor.c:5268-6645) for the nodes that lie in a chunk:
//                       distinct corners in Z-order with the far domain faces pulled in, the smallest leaf at
//                       each node, hanging nodes and their anchors (node_setproperty octor.c:3294, anchor lists
//                       octor.c:5863-5991)
//   hmesh_lnid          elem_t.lnid: the 8 corner node ids of a range of leaves
//
// Plain C ABI, no CUDA: built into libhercules_mesh.so by csrc/Makefile.  Coordinates are integers in units
// of the finest admissible edge h, below 2^16; Morton codes interleave x (least significant), y, z.
#include "hercules_mesh.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

inline uint64_t spread(uint64_t v)
{
    v &= 0x1FFFFF;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
inline uint64_t compact(uint64_t v)
{
    v &= 0x1249249249249249ull;
    v = (v | (v >> 2)) & 0x10C30C30C30C30C3ull;
    v = (v | (v >> 4)) & 0x100F00F00F00F00Full;
    v = (v | (v >> 8)) & 0x1F0000FF0000FFull;
    v = (v | (v >> 16)) & 0x1F00000000FFFFull;
    v = (v | (v >> 32)) & 0x1FFFFF;
    return v;
}
inline uint64_t code3(int64_t x, int64_t y, int64_t z) { return spread((uint64_t)x) | (spread((uint64_t)y) << 1) | (spread((uint64_t)z) << 2); }

// Open-addressing set of octants: key = code << 4 | log2(edge).  Removal leaves a tombstone.
struct OctSet {
    static constexpr uint64_t EMPTY = ~0ull, DEAD = ~0ull - 1;
    std::vector<uint64_t> t;
    uint64_t mask = 0;
    size_t used = 0;
    explicit OctSet(size_t expect)
    {
        size_t cap = 64;
        while (cap < 2 * expect + 16) cap <<= 1;
        t.assign(cap, EMPTY);
        mask = cap - 1;
    }
    static inline uint64_t hash(uint64_t k) { k ^= k >> 31; k *= 0x9E3779B97F4A7C15ull; k ^= k >> 29; return k; }
    void grow()
    {
        std::vector<uint64_t> old;
        old.swap(t);
        t.assign(old.size() * 2, EMPTY);
        mask = t.size() - 1;
        used = 0;
        for (uint64_t k : old) if (k != EMPTY && k != DEAD) insert(k);
    }
    void insert(uint64_t k)
    {
        if (2 * (used + 1) > t.size()) grow();
        uint64_t i = hash(k) & mask;
        while (t[i] != EMPTY && t[i] != DEAD) i = (i + 1) & mask;
        if (t[i] == EMPTY) used++;
        t[i] = k;
    }
    bool contains(uint64_t k) const
    {
        uint64_t i = hash(k) & mask;
        while (t[i] != EMPTY) {
            if (t[i] == k) return true;
            i = (i + 1) & mask;
        }
        return false;
    }
    bool erase(uint64_t k)
    {
        uint64_t i = hash(k) & mask;
        while (t[i] != EMPTY) {
            if (t[i] == k) { t[i] = DEAD; return true; }
            i = (i + 1) & mask;
        }
        return false;
    }
};

struct Oct { int32_t x, y, z; };

struct Model {
    const uint8_t *grid;        // material index per model cell, [gx][gy][gz]
    int64_t gx, gy, gz;
    int32_t cl;                 // edge of a model cell in units of h
    const double *vs;           // Vs per material (already clamped to vs_min)
    double factor_h;            // h * points per wavelength * f_max: an octant of edge s is split when s * factor_h > Vs
    // Vs at the centre (x + s/2, ...) of an octant: the model cell that holds it
    inline double vs_at(int32_t x, int32_t y, int32_t z, int32_t s) const
    {
        const int64_t d = 2 * (int64_t)cl;
        const int64_t ix = (2 * (int64_t)x + s) / d, iy = (2 * (int64_t)y + s) / d, iz = (2 * (int64_t)z + s) / d;
        return vs[grid[(ix * gy + iy) * gz + iz]];
    }
};

}  // namespace

extern "C" {

// One chunk.  reg = the coarse cells to refine + balance together ([nreg][3] lowest corners, edge S; NULL = sub
// and its 26 neighbours inside the domain), sub =
// those of them whose leaves are wanted ([nsub][3], ascending Morton order).  Outputs (malloc'ed, release with
// hmesh_free): codes / sizes of the leaves inside sub in ascending Morton order, and per_cell[nsub] = leaves
// per cell of sub (caller-provided).  want_leaves = 0: only the counts.  Returns 0, or -1 on bad arguments.
int hmesh_chunk_leaves(const int32_t *dims, int32_t S, int32_t smin, const uint8_t *mat_grid, const int64_t *grid_dims,
                       int32_t cl, const double *vs_tab, double factor_h, const int32_t *reg, int64_t nreg,
                       const int32_t *sub, int64_t nsub, int32_t want_leaves, uint64_t **codes_out, int32_t **sizes_out,
                       int64_t *n_out, int64_t *per_cell)
{
    if (!dims || S <= 0 || (S & (S - 1)) || !mat_grid || !grid_dims || cl <= 0 || !vs_tab || !sub || !n_out || !per_cell)
        return -1;
    int K = 0;
    while ((1 << K) < S) K++;
    if (smin < 1) smin = 1;
    std::vector<int32_t> ring;
    if (!reg) {                                      // reg = sub and its 26 neighbours inside the domain
        std::vector<uint64_t> keys;
        for (int64_t i = 0; i < nsub; i++)
            for (int dz = -1; dz <= 1; dz++)
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        const int32_t qx = sub[3 * i] + dx * S, qy = sub[3 * i + 1] + dy * S, qz = sub[3 * i + 2] + dz * S;
                        if (qx < 0 || qx >= dims[0] || qy < 0 || qy >= dims[1] || qz < 0 || qz >= dims[2]) continue;
                        keys.push_back(code3(qx, qy, qz));
                    }
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        ring.resize(3 * keys.size());
        for (size_t i = 0; i < keys.size(); i++) {
            ring[3 * i] = (int32_t)compact(keys[i]); ring[3 * i + 1] = (int32_t)compact(keys[i] >> 1); ring[3 * i + 2] = (int32_t)compact(keys[i] >> 2);
        }
        reg = ring.data();
        nreg = (int64_t)keys.size();
    }
    const Model M{mat_grid, grid_dims[0], grid_dims[1], grid_dims[2], cl, vs_tab, factor_h};
    const int32_t nx = dims[0], ny = dims[1], nz = dims[2];

    // ---- refine: top-down from the cells of reg, by the vs rule ----
    std::vector<std::vector<Oct>> lev(K + 1);       // leaves by log2(edge)
    {
        std::vector<Oct> cur((size_t)nreg), nxt;
        for (int64_t i = 0; i < nreg; i++) cur[i] = Oct{reg[3 * i], reg[3 * i + 1], reg[3 * i + 2]};
        for (int k = K; k >= 0 && !cur.empty(); k--) {
            const int32_t s = 1 << k, hs = s >> 1;
            nxt.clear();
            for (const Oct &o : cur) {
                if (s > smin && s * M.factor_h > M.vs_at(o.x, o.y, o.z, s)) {
                    for (int j = 0; j < 8; j++) nxt.push_back(Oct{o.x + hs * (j & 1), o.y + hs * ((j >> 1) & 1), o.z + hs * ((j >> 2) & 1)});
                } else {
                    lev[k].push_back(o);
                }
            }
            cur.swap(nxt);
        }
    }
    size_t total = 0;
    for (auto &v : lev) total += v.size();
    OctSet set(total + total / 4);
    for (int k = 0; k <= K; k++)
        for (const Oct &o : lev[k]) set.insert(code3(o.x, o.y, o.z) << 4 | (uint64_t)k);

    // ---- balance: finest level first; a leaf of edge s forbids leaves larger than 2 s at its 18 neighbours.
    //      Siblings ask for the same cells of edge 2 s around their parent, so the requests are collected per
    //      parent (leaves sorted by code: siblings are adjacent) and every cell is probed once. ----
    std::vector<uint64_t> order;
    for (int k = 0; k + 2 <= K; k++) {
        const int32_t s = 1 << k, t2 = 2 * s;
        // live leaves of this level (leaves of a level are only created while FINER levels are processed), by parent
        order.clear();
        for (const Oct &o : lev[k]) {
            const uint64_t c = code3(o.x, o.y, o.z);
            if (set.contains(c << 4 | (uint64_t)k)) order.push_back(c);
        }
        std::sort(order.begin(), order.end());
        const uint64_t pmask = ~((8ull << (3 * k)) - 1);            // code of the parent (edge 2 s)
        for (size_t li = 0; li < order.size();) {
            const uint64_t pc = order[li] & pmask;
            uint32_t want = 0;                                    // bit (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1): neighbour cell of the parent
            for (; li < order.size() && (order[li] & pmask) == pc; li++) {
                const unsigned child = (unsigned)((order[li] >> (3 * k)) & 7);
                const int b[3] = {(int)(child & 1), (int)((child >> 1) & 1), (int)((child >> 2) & 1)};
                for (int dz = -1; dz <= 1; dz++)
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dx = -1; dx <= 1; dx++) {
                            const int nzc = (dx != 0) + (dy != 0) + (dz != 0);
                            if (nzc < 1 || nzc > 2) continue;
                            const int ox = (dx < 0 && !b[0]) ? -1 : (dx > 0 && b[0]) ? 1 : 0;
                            const int oy = (dy < 0 && !b[1]) ? -1 : (dy > 0 && b[1]) ? 1 : 0;
                            const int oz = (dz < 0 && !b[2]) ? -1 : (dz > 0 && b[2]) ? 1 : 0;
                            want |= 1u << ((oz + 1) * 9 + (oy + 1) * 3 + (ox + 1));
                        }
            }
            want &= ~(1u << 13);                                  // the parent itself
            int32_t Px, Py, Pz;
            Px = (int32_t)compact(pc); Py = (int32_t)compact(pc >> 1); Pz = (int32_t)compact(pc >> 2);
            for (int bit = 0; bit < 27; bit++) {
                if (!(want >> bit & 1)) continue;
                const int32_t ax = Px + (bit % 3 - 1) * t2, ay = Py + (bit / 3 % 3 - 1) * t2, az = Pz + (bit / 9 - 1) * t2;
                if (ax < 0 || ax >= nx || ay < 0 || ay >= ny || az < 0 || az >= nz) continue;
                // the leaf larger than 2 s that holds the cell (ax, ay, az; 2 s), if any
                for (int kt = k + 2; kt <= K; kt++) {
                    const int32_t t = 1 << kt;
                    int32_t cx = ax & ~(t - 1), cy = ay & ~(t - 1), cz = az & ~(t - 1);
                    if (!set.erase(code3(cx, cy, cz) << 4 | (uint64_t)kt)) continue;
                    // split it down to edge 2 s along the way to the cell; the other children become leaves
                    for (int ku = kt; ku > k + 1; ku--) {
                        const int32_t hu = 1 << (ku - 1);
                        int32_t px = cx, py = cy, pz = cz;
                        for (int j = 0; j < 8; j++) {
                            const int32_t kx = cx + hu * (j & 1), ky = cy + hu * ((j >> 1) & 1), kz = cz + hu * ((j >> 2) & 1);
                            const bool on_path = (ku - 1 > k + 1) && kx == (ax & ~(hu - 1)) && ky == (ay & ~(hu - 1)) &&
                                                 kz == (az & ~(hu - 1));
                            if (on_path) { px = kx; py = ky; pz = kz; continue; }
                            set.insert(code3(kx, ky, kz) << 4 | (uint64_t)(ku - 1));
                            lev[ku - 1].push_back(Oct{kx, ky, kz});
                        }
                        cx = px; cy = py; cz = pz;
                    }
                    break;
                }
            }
        }
    }

    // ---- the leaves inside sub, in Morton order.  lev[] still lists octants that were split: the set decides ----
    const uint64_t shift = 3 * (uint64_t)K;
    std::vector<uint64_t> subkey((size_t)nsub);
    for (int64_t i = 0; i < nsub; i++) subkey[i] = code3(sub[3 * i], sub[3 * i + 1], sub[3 * i + 2]) >> shift;
    for (int64_t i = 1; i < nsub; i++) if (subkey[i] <= subkey[i - 1]) return -1;
    std::vector<uint64_t> out;                        // code << 4 | level: sorts like the code (codes of leaves are distinct)
    std::memset(per_cell, 0, sizeof(int64_t) * (size_t)nsub);
    for (uint64_t key : set.t) {
        if (key == OctSet::EMPTY || key == OctSet::DEAD) continue;
        const uint64_t ck = (key >> 4) >> shift;
        auto it = std::lower_bound(subkey.begin(), subkey.end(), ck);
        if (it == subkey.end() || *it != ck) continue;
        per_cell[it - subkey.begin()]++;
        if (want_leaves) out.push_back(key);
    }
    *n_out = 0;
    if (want_leaves) {
        std::sort(out.begin(), out.end());
        uint64_t *codes = (uint64_t *)std::malloc(sizeof(uint64_t) * (out.size() + 1));
        int32_t *sizes = (int32_t *)std::malloc(sizeof(int32_t) * (out.size() + 1));
        if (!codes || !sizes) { std::free(codes); std::free(sizes); return -2; }
        for (size_t i = 0; i < out.size(); i++) { codes[i] = out[i] >> 4; sizes[i] = 1 << (int)(out[i] & 15); }
        *codes_out = codes; *sizes_out = sizes; *n_out = (int64_t)out.size();
    }
    return 0;
}

// ---- extraction ---------------------------------------------------------------------------------------------
// X = the coarse cells a rank knows exactly (its own and one ring), ascending keys xkeys[nX]; the leaves of X
// in Morton order (lcodes, lsizes) with lstart[nX + 1] = first leaf of every cell.  A node is LOCATED in the
// cell that holds its point, far domain faces pulled in by one tick (octor.c:5466-5475); its code is the
// Morton code of the doubled coordinates (2 g, or 2 n - 1 on a far face): ascending code = octor's node order,
// and all nodes of a cell are contiguous in it.

namespace {

struct XCells {
    const uint64_t *keys; int64_t n; int K; int32_t S; int32_t dims[3];
    inline int64_t find(uint64_t key) const
    {
        const uint64_t *it = std::lower_bound(keys, keys + n, key);
        return (it != keys + n && *it == key) ? (int64_t)(it - keys) : -1;
    }
    inline uint64_t key_of_point(int32_t x, int32_t y, int32_t z) const { return code3(x, y, z) >> (3 * K); }
    inline uint64_t nkey(int32_t g, int c) const { return g == dims[c] ? 2ull * (uint64_t)g - 1 : 2ull * (uint64_t)g; }
    inline uint64_t ncode(int32_t x, int32_t y, int32_t z) const { return code3((int64_t)nkey(x, 0), (int64_t)nkey(y, 1), (int64_t)nkey(z, 2)); }
    inline uint64_t cell_key_of_ncode(uint64_t nc) const { return nc >> (3 * K + 3); }
};

inline void decode3(uint64_t c, int32_t &x, int32_t &y, int32_t &z) { x = (int32_t)compact(c); y = (int32_t)compact(c >> 1); z = (int32_t)compact(c >> 2); }

struct NodeRec { uint64_t code; int32_t size; };

}  // namespace

// Nodes located in the cells X[i0, i1).  Outputs (malloc'ed): ncodes ascending, xyz [n][3] coordinates, holder =
// index (into lcodes) of the leaf whose half-open box holds the node, far faces pulled in (-1: none in X), dang
// (0 = anchored, 2 / 4 = hanging on an edge / a face: the number of anchors), anchors [ndang][4] as node codes
// for the hanging nodes in order (~0 = unused); per_cell[i1 - i0] = nodes per cell.
int hmesh_chunk_nodes(const int32_t *dims, int32_t S, const uint64_t *xkeys, int64_t nX, const int64_t *lstart,
                      const uint64_t *lcodes, const int32_t *lsizes, int64_t i0, int64_t i1, uint64_t **ncodes_out,
                      int32_t **xyz_out, int64_t **holder_out, uint8_t **dang_out, uint64_t **anchors_out, int64_t *n_out,
                      int64_t *ndang_out, int64_t *per_cell)
{
    if (!dims || S <= 0 || (S & (S - 1)) || !xkeys || !lstart || !lcodes || !lsizes || i0 < 0 || i1 > nX || i0 >= i1 || !n_out || !per_cell)
        return -1;
    XCells X{xkeys, nX, 0, S, {dims[0], dims[1], dims[2]}};
    while ((1 << X.K) < S) X.K++;
    const int32_t cdim[3] = {dims[0] / S, dims[1] / S, dims[2] / S};
    // the cells of X around the chunk
    std::vector<int64_t> reg;
    for (int64_t i = i0; i < i1; i++) {
        int32_t cx, cy, cz;
        decode3(xkeys[i], cx, cy, cz);
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const int32_t qx = cx + dx, qy = cy + dy, qz = cz + dz;
                    if (qx < 0 || qx >= cdim[0] || qy < 0 || qy >= cdim[1] || qz < 0 || qz >= cdim[2]) continue;
                    const int64_t j = X.find(code3(qx, qy, qz));
                    if (j >= 0) reg.push_back(j);
                }
    }
    std::sort(reg.begin(), reg.end());
    reg.erase(std::unique(reg.begin(), reg.end()), reg.end());
    // corners of their leaves that are located in the chunk
    const uint64_t klo = xkeys[i0], khi = xkeys[i1 - 1];
    std::vector<NodeRec> rec;
    for (int64_t j : reg)
        for (int64_t l = lstart[j]; l < lstart[j + 1]; l++) {
            int32_t x, y, z;
            decode3(lcodes[l], x, y, z);
            const int32_t s = lsizes[l];
            for (int c = 0; c < 8; c++) {
                const uint64_t nc = X.ncode(x + s * (c & 1), y + s * ((c >> 1) & 1), z + s * ((c >> 2) & 1));
                const uint64_t ck = X.cell_key_of_ncode(nc);
                if (ck < klo || ck > khi) continue;
                if (!std::binary_search(xkeys + i0, xkeys + i1, ck)) continue;
                rec.push_back(NodeRec{nc, s});
            }
        }
    std::sort(rec.begin(), rec.end(), [](const NodeRec &a, const NodeRec &b) { return a.code != b.code ? a.code < b.code : a.size < b.size; });
    size_t n = 0;
    for (size_t i = 0; i < rec.size(); i++) if (i == 0 || rec[i].code != rec[i - 1].code) n++;
    uint64_t *nc = (uint64_t *)std::malloc(sizeof(uint64_t) * (n + 1));
    int32_t *xyz = (int32_t *)std::malloc(sizeof(int32_t) * 3 * (n + 1));
    int64_t *hold = (int64_t *)std::malloc(sizeof(int64_t) * (n + 1));
    uint8_t *dg = (uint8_t *)std::malloc(n + 1);
    std::vector<uint64_t> anv;
    if (!nc || !xyz || !hold || !dg) { std::free(nc); std::free(xyz); std::free(hold); std::free(dg); return -2; }
    std::memset(per_cell, 0, sizeof(int64_t) * (size_t)(i1 - i0));
    auto is_leaf = [&](int32_t x, int32_t y, int32_t z, int32_t t) -> bool {
        if (t > S || x < 0 || y < 0 || z < 0 || x + t > dims[0] || y + t > dims[1] || z + t > dims[2]) return false;
        const int64_t j = X.find(X.key_of_point(x, y, z));
        if (j < 0) return false;
        const uint64_t code = code3(x, y, z);
        const uint64_t *b = lcodes + lstart[j], *e = lcodes + lstart[j + 1];
        const uint64_t *it = std::lower_bound(b, e, code);
        return it != e && *it == code && lsizes[it - lcodes] == t;
    };
    size_t k = 0;
    for (size_t i = 0; i < rec.size();) {
        size_t j = i;
        while (j < rec.size() && rec[j].code == rec[i].code) j++;
        const uint64_t code = rec[i].code;
        const int32_t smin = rec[i].size;
        const size_t touches = j - i;
        nc[k] = code; dg[k] = 0;
        const int64_t cellj = i0 + (std::lower_bound(xkeys + i0, xkeys + i1, X.cell_key_of_ncode(code)) - (xkeys + i0));
        per_cell[cellj - i0]++;
        int32_t kx, ky, kz;
        decode3(code, kx, ky, kz);
        const int32_t P[3] = {(kx & 1) ? dims[0] : kx / 2, (ky & 1) ? dims[1] : ky / 2, (kz & 1) ? dims[2] : kz / 2};
        xyz[3 * k] = P[0]; xyz[3 * k + 1] = P[1]; xyz[3 * k + 2] = P[2];
        {   // the leaf that holds the point (far faces pulled in): the last leaf of the node's cell whose code is not above the point's
            const int32_t q[3] = {std::min(P[0], dims[0] - 1), std::min(P[1], dims[1] - 1), std::min(P[2], dims[2] - 1)};
            const uint64_t qc = code3(q[0], q[1], q[2]);
            const uint64_t *b = lcodes + lstart[cellj], *e = lcodes + lstart[cellj + 1];
            const uint64_t *it = std::upper_bound(b, e, qc);
            hold[k] = it == b ? -1 : (int64_t)(it - 1 - lcodes);
        }
        if (touches < 8) {
            // a leaf twice the size of the node's smallest leaf that holds it inside an edge or a face
            const int32_t t = 2 * smin;
            const bool odd[3] = {P[0] % t != 0, P[1] % t != 0, P[2] % t != 0};
            const int nodd = odd[0] + odd[1] + odd[2];
            if (nodd >= 1 && nodd <= 2) {
                bool hit = false;
                for (int m = 0; m < 8 && !hit; m++) {
                    int32_t cand[3];
                    bool dup = false;
                    for (int c = 0; c < 3; c++) {
                        if (odd[c]) { cand[c] = P[c] - smin; dup |= ((m >> c) & 1) != 0; }     // the bit of an off-grid axis changes nothing
                        else cand[c] = ((m >> c) & 1) ? P[c] - t : P[c];
                    }
                    if (!dup) hit = is_leaf(cand[0], cand[1], cand[2], t);
                }
                if (hit) {
                    // anchors in descending Z-order: +- smin along the off-grid axes, the higher axis varying slowest
                    int ax1 = odd[0] ? 0 : (odd[1] ? 1 : 2), ax2 = odd[2] ? 2 : (odd[1] ? 1 : 0);
                    dg[k] = nodd == 1 ? 2 : 4;
                    for (int a = 0; a < 4; a++) {
                        if (a >= (nodd == 1 ? 2 : 4)) { anv.push_back(~0ull); continue; }
                        int32_t q[3] = {P[0], P[1], P[2]};
                        q[ax1] += (a % 2 == 0 ? 1 : -1) * smin;
                        if (nodd == 2) q[ax2] += (a < 2 ? 1 : -1) * smin;
                        anv.push_back(X.ncode(q[0], q[1], q[2]));
                    }
                }
            }
        }
        k++;
        i = j;
    }
    uint64_t *an = (uint64_t *)std::malloc(sizeof(uint64_t) * (anv.size() + 4));
    if (!an) { std::free(nc); std::free(xyz); std::free(hold); std::free(dg); return -2; }
    std::memcpy(an, anv.data(), sizeof(uint64_t) * anv.size());
    *ncodes_out = nc; *xyz_out = xyz; *holder_out = hold; *dang_out = dg; *anchors_out = an; *n_out = (int64_t)n;
    *ndang_out = (int64_t)(anv.size() / 4);
    return 0;
}

// lnid[(e1 - e0)][8] for the leaves [e0, e1): the global index (position in ncodes) of each corner; `missing`
// for a corner located in a cell that is not in X (outer faces of X).  nstart[nX + 1] = first node of every cell.
// exyz (optional) [(e1 - e0)][3]: the leaves' lowest corners.
int hmesh_lnid(const int32_t *dims, int32_t S, const uint64_t *xkeys, int64_t nX, const int64_t *nstart, const uint64_t *ncodes,
               const uint64_t *lcodes, const int32_t *lsizes, int64_t e0, int64_t e1, int32_t missing, int32_t *lnid, int32_t *exyz)
{
    if (!dims || S <= 0 || (S & (S - 1)) || !xkeys || !nstart || !ncodes || !lcodes || !lsizes || !lnid || e0 > e1) return -1;
    XCells X{xkeys, nX, 0, S, {dims[0], dims[1], dims[2]}};
    while ((1 << X.K) < S) X.K++;
    uint64_t last_key = ~0ull;
    int64_t last_j = -1;
    for (int64_t l = e0; l < e1; l++) {
        int32_t x, y, z;
        decode3(lcodes[l], x, y, z);
        const int32_t s = lsizes[l];
        if (exyz) { exyz[3 * (l - e0)] = x; exyz[3 * (l - e0) + 1] = y; exyz[3 * (l - e0) + 2] = z; }
        for (int c = 0; c < 8; c++) {
            const uint64_t nc = X.ncode(x + s * (c & 1), y + s * ((c >> 1) & 1), z + s * ((c >> 2) & 1));
            const uint64_t ck = X.cell_key_of_ncode(nc);
            if (ck != last_key) { last_key = ck; last_j = X.find(ck); }
            int32_t id = missing;
            if (last_j >= 0) {
                const uint64_t *b = ncodes + nstart[last_j], *e = ncodes + nstart[last_j + 1];
                const uint64_t *it = std::lower_bound(b, e, nc);
                if (it == e || *it != nc) return -3;             // a corner of a leaf of X inside X must be a node
                id = (int32_t)(it - ncodes);
            }
            lnid[8 * (l - e0) + c] = id;
        }
    }
    return 0;
}

// solver_init's lumped-mass sums (psolve.c:3445-3473) for the grouped form meshgen._accumulate uses: for every
// corner column j = 0..7 in turn, col[n] = sum over the elements e (ascending) with lnid[e][j] == n of w[e], then
// out[n] += col[n] -- the summation order of the numpy restatement (np.bincount per column), so that the two give
// the same doubles.  nw weight arrays w[k][E] -> out[k][N] (out is accumulated into); scratch-free for the caller.
int hmesh_corner_sums(int64_t E, const int32_t *lnid, int64_t N, int32_t nw, const double *const *w, double *const *out)
{
    if (E < 0 || N < 0 || nw < 1 || !lnid || !w || !out) return -1;
    std::vector<double> col((size_t)N);
    for (int k = 0; k < nw; k++)
        for (int j = 0; j < 8; j++) {
            std::fill(col.begin(), col.end(), 0.0);
            const double *wk = w[k];
            for (int64_t e = 0; e < E; e++) {
                const int32_t n = lnid[8 * e + j];
                if (n < 0 || n >= N) return -3;
                col[(size_t)n] += wk[e];
            }
            double *o = out[k];
            for (int64_t n = 0; n < N; n++) o[n] += col[(size_t)n];
        }
    return 0;
}

// com_allocpctl's neighbour discovery (octor.c:2639-2742) for a rank: its leaves in Morton order, per leaf 4 x 4 x 4
// probe points half an edge apart starting half an edge below the lowest corner (z outermost, x innermost), points
// outside the domain skipped; the rank of a point = the rank of the leaf that holds it.  cand[ncand] = indices (into
// lcodes, ascending) of the rank's leaves worth probing, key_base[ncand] = their position in the rank's leaf list;
// gidx[l] = global Morton index of leaf l, rank of a leaf = ((gidx + 1) * world - 1) / etotal (octor.c:738-742).
// first[world] (caller's, preset to INT64_MAX) receives, per foreign rank, the smallest key position * 64 + probe
// number at which it was met.
int hmesh_discovery(const int32_t *dims, int32_t S, const uint64_t *xkeys, int64_t nX, const int64_t *lstart,
                    const uint64_t *lcodes, const int32_t *lsizes, const int64_t *gidx, int64_t etotal, int32_t world,
                    int32_t rank, const int64_t *cand, const int64_t *key_base, int64_t ncand, int64_t *first)
{
    if (!dims || S <= 0 || (S & (S - 1)) || !xkeys || !lstart || !lcodes || !lsizes || !gidx || !cand || !key_base || !first ||
        world < 1 || etotal < 1)
        return -1;
    XCells X{xkeys, nX, 0, S, {dims[0], dims[1], dims[2]}};
    while ((1 << X.K) < S) X.K++;
    uint64_t last_key = ~0ull;
    int64_t last_j = -1;
    for (int64_t ci = 0; ci < ncand; ci++) {
        const int64_t l = cand[ci];
        int32_t x, y, z;
        decode3(lcodes[l], x, y, z);
        const int32_t s = lsizes[l];
        for (int k = 0; k < 4; k++)
            for (int j = 0; j < 4; j++)
                for (int i = 0; i < 4; i++) {
                    // doubled coordinates: 2 x - s + s i
                    const int64_t p2[3] = {2 * (int64_t)x - s + (int64_t)s * i, 2 * (int64_t)y - s + (int64_t)s * j,
                                           2 * (int64_t)z - s + (int64_t)s * k};
                    if (p2[0] < 0 || p2[0] >= 2 * (int64_t)dims[0] || p2[1] < 0 || p2[1] >= 2 * (int64_t)dims[1] || p2[2] < 0 ||
                        p2[2] >= 2 * (int64_t)dims[2])
                        continue;
                    const int32_t q[3] = {(int32_t)(p2[0] / 2), (int32_t)(p2[1] / 2), (int32_t)(p2[2] / 2)};
                    const uint64_t ck = X.key_of_point(q[0], q[1], q[2]);
                    if (ck != last_key) { last_key = ck; last_j = X.find(ck); }
                    if (last_j < 0) return -3;                    // a probe of a rank's leaf lies within X by construction
                    const uint64_t qc = code3(q[0], q[1], q[2]);
                    const uint64_t *b = lcodes + lstart[last_j], *e = lcodes + lstart[last_j + 1];
                    const uint64_t *it = std::upper_bound(b, e, qc);
                    if (it == b) return -3;
                    const int64_t h = (int64_t)(it - 1 - lcodes);
                    const int64_t r = ((gidx[h] + 1) * (int64_t)world - 1) / etotal;
                    if (r == rank || r < 0 || r >= world) continue;
                    const int64_t key = key_base[ci] * 64 + (k * 16 + j * 4 + i);
                    if (key < first[r]) first[r] = key;
                }
    }
    return 0;
}

void hmesh_free(void *p) { std::free(p); }

int hmesh_abi_version(void) { return 1; }

}  // extern "C"
