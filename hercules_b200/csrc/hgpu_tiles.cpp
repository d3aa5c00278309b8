// hgpu_tiles.cpp -- host-side index builders for libhercules_gpu.so (no CUDA in this file).
//
// build_tile_plan     tiling that replaces the scatter-add of compute_addforce_effective /
//                     damping_addforce / constant_Q_addforce (stiffness.c:228-235, damping.c:88-98,
//                     406-412) with an atomic-free, fixed-order reduction: see DESIGN.md section 3.
// build_dangling_plan anchor-centric CSR for compute_adjust(DISTRIBUTION) (psolve.c:5943-5987).
#include "hgpu_internal.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

namespace hgpu {

// Tiles are contiguous, even-aligned node ranges.  Octor numbers both elements and nodes in Morton
// order (octor.c:5373-5507, 6166), so the nodes whose highest-numbered incident element falls in
// one block of `elem_block` consecutive elements form one compact, nearly cubic patch (an aligned
// 8x8x8 cell of a uniform region when elem_block = 512), and the elements whose corner 0 lies in
// that patch -- the tile's CORE elements -- are that block itself.  Corner 0 of an element is its
// componentwise lowest node, hence lowest in Z-order: every node a core element touches belongs to
// this tile or to a HIGHER one, so partial forces only ever travel from lower to higher tiles.
// A tile that exceeds one of the capacities is split in half until it fits.
bool build_tile_plan(int32_t E, int32_t N, const int32_t *lnid, const TileCaps &caps,
                     const uint8_t *self_node, const uint8_t *special_node, TilePlan &plan,
                     std::string &err)
{
    if (caps.elem_block <= 0 || caps.max_owned < 2 || caps.max_slots < 16 || caps.max_acc < 16 ||
        caps.max_recs < 2 || caps.max_srcs < 16) { err = "bad tile limits"; return false; }
    const int32_t max_owned = std::min(caps.max_owned, 65535 / 3) & ~1;
    const int32_t max_slots = std::min(caps.max_slots, 65535 / 3), max_acc = std::min(caps.max_acc, max_slots);
    plan = TilePlan();

    // node -> incident elements (CSR, ascending element id) and the highest incident element
    std::vector<int32_t> noff((size_t)N + 1, 0);
    for (int32_t e = 0; e < E; e++)
        for (int j = 0; j < 8; j++) {
            const int32_t n = lnid[8 * (size_t)e + j];
            if (n < 0 || n >= N) { err = "element node id out of range"; return false; }
            noff[(size_t)n + 1]++;
        }
    for (int32_t n = 0; n < N; n++) noff[(size_t)n + 1] += noff[n];
    if ((int64_t)8 * E > (int64_t)INT32_MAX) { err = "more than 2^31 element corners on one rank"; return false; }
    std::vector<int32_t> nelem((size_t)8 * E);
    {
        std::vector<int32_t> cur(noff.begin(), noff.end() - 1);
        for (int32_t e = 0; e < E; e++)
            for (int j = 0; j < 8; j++) nelem[(size_t)cur[lnid[8 * (size_t)e + j]]++] = e;
    }

    // cut points
    int32_t min_tile = std::min(256, max_owned / 2);
    { const char *e = getenv("HGPU_MIN_TILE"); if (e && atoi(e) > 0) min_tile = std::min(atoi(e), max_owned); }
    std::vector<int32_t> cuts;
    cuts.push_back(0);
    {
        // cuts fall on even node ids only (16-byte aligned bulk copies of the owned range); when
        // the block changes at an odd id that one node stays with the tile before it
        int32_t cur_blk = -1, last_blk = -1, start = 0;
        for (int32_t n = 0; n < N; n++) {
            const int32_t blk = noff[(size_t)n + 1] > noff[n]
                                    ? nelem[(size_t)noff[(size_t)n + 1] - 1] / caps.elem_block : last_blk;
            last_blk = blk;
            if (n == start) { cur_blk = blk; continue; }
            // At the interface of two octree levels the Z-ordered node list alternates between
            // nodes whose highest element lies in a block of fine elements and nodes whose highest
            // element lies in a block of coarse ones: cutting at every change would shred the
            // interface into tiles of one to three nodes.  A change of block therefore only ends a
            // tile that has reached a useful size; shorter runs are absorbed into it.
            if (!(n & 1) && ((blk != cur_blk && n - start >= min_tile) || n - start >= max_owned)) {
                cuts.push_back(n);
                start = n;
                cur_blk = blk;
            }
        }
        cuts.push_back(N);
    }

    const char *e1 = getenv("HGPU_PLAN_SORT"), *e2 = getenv("HGPU_PLAN_GREEDY");
    const char *e3 = getenv("HGPU_PLAN_PACK");
    const bool opt_sort = !(e1 && atoi(e1) == 0), opt_greedy = !(e2 && atoi(e2) == 0), opt_pack = !(e3 && atoi(e3) == 0);
    std::unordered_map<std::string, std::vector<int32_t>> pack_cache;
    std::vector<uint8_t> g_cnt;
    std::vector<int32_t> g_roff, g_rval, g_rcur, g_order, g_hslot;
    std::vector<int32_t> freeslots[16];
    std::vector<int32_t> stamp_e((size_t)E, -1), stamp_n((size_t)N, -1), slot_of((size_t)N, 0);
    std::vector<int32_t> tile_of((size_t)N, -1);
    std::vector<int32_t> elems, extras, halo, ftiles;
    std::vector<int32_t> want_recs, want_srcs;       // per tile, counted while the tile is built
    plan.node_off.push_back(0);
    plan.elem_off.push_back(0);
    plan.halo_off.push_back(0);
    // worklist of [a, b) ranges, processed in ascending order (split ranges are re-queued in place)
    std::vector<std::pair<int32_t, int32_t>> work;
    for (size_t i = cuts.size() - 1; i > 0; i--) work.emplace_back(cuts[i - 1], cuts[i]);
    int32_t stamp = 0;
    while (!work.empty()) {
        const int32_t a = work.back().first, b = work.back().second;
        work.pop_back();
        if (a >= b) continue;
        const int32_t nown = b - a;
        const int32_t tile = (int32_t)plan.node_off.size() - 1;
        bool fits = nown <= max_owned;
        bool is_self = false;
        if (self_node) for (int32_t n = a; n < b && !is_self; n++) is_self = self_node[n] != 0;
        elems.clear(); extras.clear(); halo.clear();
        int32_t npub = 0, nrec = 0, nsrc = 0;
        if (fits) {
            for (int32_t n = a; n < b; n++) {
                ftiles.clear();
                for (int32_t k = noff[n]; k < noff[(size_t)n + 1]; k++) {
                    const int32_t e = nelem[k];
                    const int32_t c0 = lnid[8 * (size_t)e];
                    if (c0 == n) elems.push_back(e);                    // core: corner 0 owned
                    else if (c0 < a || c0 >= b) {
                        if (c0 > n) { err = "element whose corner 0 is not its lowest node (node ids are not in Z-order)"; return false; }
                        if (is_self) { if (stamp_e[e] != stamp) { stamp_e[e] = stamp; extras.push_back(e); } }
                        else {
                            const int32_t ft = tile_of[c0];
                            if (ft < 0) { err = "internal: foreign corner-0 node not yet tiled"; return false; }
                            if (std::find(ftiles.begin(), ftiles.end(), ft) == ftiles.end()) ftiles.push_back(ft);
                        }
                    }
                }
                if (ftiles.size() > 8) { err = "a node receives partial forces from more than 8 tiles"; return false; }
                nsrc += (int32_t)ftiles.size();
                if (!ftiles.empty() || (special_node && special_node[n])) nrec++;
            }
            for (int32_t e : elems)
                for (int j = 0; j < 8; j++) {
                    const int32_t n = lnid[8 * (size_t)e + j];
                    if (n >= a && n < b) continue;
                    if (stamp_n[n] != stamp) { stamp_n[n] = stamp; halo.push_back(n); }
                }
            npub = (int32_t)halo.size();
            for (int32_t e : extras)
                for (int j = 0; j < 8; j++) {
                    const int32_t n = lnid[8 * (size_t)e + j];
                    if (n >= a && n < b) continue;
                    if (stamp_n[n] != stamp) { stamp_n[n] = stamp; halo.push_back(n); }
                }
            fits = (int64_t)nown + npub <= (int64_t)max_acc &&
                   (int64_t)nown + (int64_t)halo.size() <= (int64_t)max_slots &&
                   nrec <= caps.max_recs && nsrc <= caps.max_srcs;
        }
        stamp++;
        if (!fits) {
            if (nown <= 2) { err = "a 2-node tile exceeds the staging capacity"; return false; }
            const int32_t mid = a + ((nown / 2 + 1) & ~1);
            work.emplace_back(mid, b);
            work.emplace_back(a, mid);
            continue;
        }
        std::sort(halo.begin(), halo.begin() + npub);
        std::sort(halo.begin() + npub, halo.end());
        // Structured tile?  (hgpu_internal.h)  The 9x9x9 node block is read off the core elements: element k
        // of the Morton-ordered cell sits at (x, y, z) = de-interleave(k), its corner j at (x + jx, y + jy, z + jz).
        bool is_struct = false;
        int32_t sgrid[STRUCT_NODES];
        if (caps.allow_struct && !is_self && nown == STRUCT_OWNED && (int32_t)elems.size() == STRUCT_OWNED &&
            extras.empty() && (int32_t)halo.size() == STRUCT_HALO && npub == STRUCT_HALO) {
            std::vector<int32_t> byc0(elems);
            std::sort(byc0.begin(), byc0.end(), [&](int32_t x, int32_t y) { return lnid[8 * (size_t)x] < lnid[8 * (size_t)y]; });
            std::fill(sgrid, sgrid + STRUCT_NODES, -1);
            is_struct = true;
            for (int32_t k = 0; k < STRUCT_OWNED && is_struct; k++) {
                const int32_t *ln = lnid + 8 * (size_t)byc0[k];
                if (ln[0] != a + k) { is_struct = false; break; }
                int x = 0, y = 0, z = 0;
                for (int bt = 0; bt < 3; bt++) { x |= ((k >> (3 * bt)) & 1) << bt; y |= ((k >> (3 * bt + 1)) & 1) << bt; z |= ((k >> (3 * bt + 2)) & 1) << bt; }
                for (int j = 0; j < 8; j++) {
                    int32_t &g = sgrid[((z + (j >> 2)) * 9 + (y + ((j >> 1) & 1))) * 9 + (x + (j & 1))];
                    if (g == -1) g = ln[j]; else if (g != ln[j]) { is_struct = false; break; }
                }
            }
            for (int z = 0; z < 9 && is_struct; z++) for (int y = 0; y < 9 && is_struct; y++) for (int x = 0; x < 9; x++) {
                const int32_t g = sgrid[(z * 9 + y) * 9 + x];
                const bool inner = x < 8 && y < 8 && z < 8;
                if (g < 0 || (inner ? g != a + struct_morton3(x, y, z) : (g >= a && g < b))) { is_struct = false; break; }
            }
            // 217 distinct gathered nodes = the halo list
            if (is_struct) {
                std::vector<int32_t> hs;
                hs.reserve(STRUCT_HALO);
                for (int z = 0; z < 9; z++) for (int y = 0; y < 9; y++) for (int x = 0; x < 9; x++)
                    if (x == 8 || y == 8 || z == 8) hs.push_back(sgrid[(z * 9 + y) * 9 + x]);
                std::sort(hs.begin(), hs.end());
                if (!std::equal(hs.begin(), hs.end(), halo.begin())) is_struct = false;
            }
        }
        // Entry order.  Threads take consecutive entries, and a shared-memory access of 16 lanes
        // (one half-warp of 8-byte words) is conflict-free when the 16 slots differ modulo 16.
        // Core entries sorted by their corner-0 node are, in a uniform region, the tile's own
        // Morton cell in Morton order: every aligned run of 16 is a 4x2x2 block whose corner-j
        // nodes have 16 different slot residues.  Extra entries (self tiles: the layers shared
        // with lower neighbours) follow, sorted by their lowest owned corner.
        const int32_t ncore = (int32_t)elems.size();
        {
            if (opt_sort) std::sort(elems.begin(), elems.end(), [&](int32_t x, int32_t y) {
                const int32_t kx = lnid[8 * (size_t)x], ky = lnid[8 * (size_t)y];
                return kx != ky ? kx < ky : x < y; });
            auto key = [&](int32_t e) -> int32_t {
                const int32_t *ln = lnid + 8 * (size_t)e;
                int32_t lo = INT32_MAX;
                for (int j = 1; j < 8; j++) if (ln[j] >= a && ln[j] < b) lo = std::min(lo, ln[j] - a);
                return lo;
            };
            std::sort(extras.begin(), extras.end());
            if (opt_sort) std::stable_sort(extras.begin(), extras.end(), [&](int32_t x, int32_t y) { return key(x) < key(y); });
            elems.insert(elems.end(), extras.begin(), extras.end());
            // Extra entries touch owned nodes on one face / edge of the patch only, whose slots
            // share residues: pack them into half-warp groups of 16 greedily so that, corner by
            // corner, the owned slots of a group collide as little as possible.  The packing
            // depends only on the residue pattern, which repeats from tile to tile: memoised.
            const int32_t nrest = (int32_t)elems.size() - ncore;
            if (opt_sort && opt_pack && nrest > 1) {
                // signature: residues (or 255) of the 8 corners of every extra entry + the partial group
                std::string sig;
                sig.reserve((size_t)8 * (nrest + 16) + 4);
                const int32_t g0 = ncore & ~15;                   // first entry of the partially filled group
                for (int32_t k = g0; k < (int32_t)elems.size(); k++)
                    for (int j = 0; j < 8; j++) {
                        const int32_t n = lnid[8 * (size_t)elems[k] + j];
                        sig.push_back((n >= a && n < b) ? (char)((n - a) & 15) : (char)-1);
                    }
                sig.push_back((char)(ncore & 15));
                auto hit = pack_cache.find(sig);
                std::vector<int32_t> perm;
                if (hit != pack_cache.end()) perm = hit->second;
                else {
                    perm.reserve(nrest);
                    std::vector<uint8_t> used((size_t)nrest, 0);
                    int cnt[8][16];
                    int32_t filled = ncore;                        // entries placed so far
                    int32_t first_free = 0;
                    while ((int32_t)perm.size() < nrest) {
                        if ((filled & 15) == 0 || perm.empty()) {
                            memset(cnt, 0, sizeof cnt);
                            if (perm.empty())                      // core entries already in this group
                                for (int32_t k = g0; k < ncore; k++)
                                    for (int j = 0; j < 8; j++) {
                                        const unsigned char r = (unsigned char)sig[(size_t)8 * (k - g0) + j];
                                        if (r != 255) cnt[j][r]++;
                                    }
                        }
                        while (first_free < nrest && used[first_free]) first_free++;
                        int32_t best = -1, best_cost = 0, seen = 0;
                        for (int32_t c = first_free; c < nrest && seen < 100000; c++) {
                            if (used[c]) continue;
                            seen++;
                            int32_t cost = 0;
                            const char *sg = sig.data() + (size_t)8 * (ncore - g0 + c);
                            for (int j = 0; j < 8; j++) { const unsigned char r = (unsigned char)sg[j]; if (r != 255) cost += cnt[j][r]; }
                            if (best < 0 || cost < best_cost) { best = c; best_cost = cost; if (!cost) break; }
                        }
                        used[best] = 1;
                        perm.push_back(best);
                        const char *sg = sig.data() + (size_t)8 * (ncore - g0 + best);
                        for (int j = 0; j < 8; j++) { const unsigned char r = (unsigned char)sg[j]; if (r != 255) cnt[j][r]++; }
                        filled++;
                    }
                    if (pack_cache.size() < 4096) pack_cache.emplace(sig, perm);
                }
                std::vector<int32_t> rest(elems.begin() + ncore, elems.end());
                for (int32_t k = 0; k < nrest; k++) elems[(size_t)ncore + k] = rest[perm[k]];
            }
        }
        // Halo slots: greedy choice of the slot residue (mod 16) that collides least with the other
        // lanes of every half-warp access the node takes part in; most-referenced nodes first.
        // Published halo nodes take the slots right after the owned range (they share the
        // accumulator with it), the others follow.
        int32_t nslots = nown, npub_slots = 0;
        {
            const int32_t ne = (int32_t)elems.size(), ngrp = (ne + 15) / 16;
            // scratch reused from tile to tile: cnt[group][corner][residue], CSR of the accesses
            // (group * 8 + corner) every halo node takes part in
            g_cnt.assign((size_t)ngrp * 8 * 16, 0);
            std::vector<uint8_t> &cnt = g_cnt;
            const size_t nhalo = halo.size();
            g_roff.assign(nhalo + 1, 0);
            for (size_t h = 0; h < nhalo; h++) slot_of[halo[h]] = -1 - (int32_t)h;   // index while unassigned
            for (int32_t k = 0; k < ne; k++) {
                const int32_t *ln = lnid + 8 * (size_t)elems[k];
                for (int j = 0; j < 8; j++) {
                    const int32_t n = ln[j];
                    if (n >= a && n < b) cnt[((size_t)(k / 16) * 8 + j) * 16 + ((n - a) & 15)]++;
                    else g_roff[(size_t)(-1 - slot_of[n]) + 1]++;
                }
            }
            for (size_t h = 0; h < nhalo; h++) g_roff[h + 1] += g_roff[h];
            g_rval.resize((size_t)g_roff[nhalo]);
            g_rcur.assign(g_roff.begin(), g_roff.end() - 1);
            for (int32_t k = 0; k < ne; k++) {
                const int32_t *ln = lnid + 8 * (size_t)elems[k];
                for (int j = 0; j < 8; j++) {
                    const int32_t n = ln[j];
                    if (!(n >= a && n < b)) g_rval[(size_t)g_rcur[(size_t)(-1 - slot_of[n])]++] = (k / 16) * 8 + j;
                }
            }
            struct Refs { const int32_t *b, *e; const int32_t *begin() const { return b; } const int32_t *end() const { return e; }
                          size_t size() const { return (size_t)(e - b); } };
            auto refs_of = [&](int32_t h) { return Refs{g_rval.data() + g_roff[(size_t)h], g_rval.data() + g_roff[(size_t)h + 1]}; };
            std::vector<int32_t> &hslot = g_hslot;
            hslot.assign(nhalo, -1);
            // assign halo nodes [h0, h1) to slots in [lo, hi); returns one past the highest slot used
            auto assign = [&](int32_t h0, int32_t h1, int32_t lo, int32_t hi) -> int32_t {
                std::vector<int32_t> &order = g_order;
                order.resize((size_t)(h1 - h0));
                for (int32_t h = h0; h < h1; h++) order[(size_t)(h - h0)] = h;
                std::stable_sort(order.begin(), order.end(),
                                 [&](int32_t x, int32_t y) { return refs_of(x).size() > refs_of(y).size(); });
                for (int r = 0; r < 16; r++) freeslots[r].clear();
                for (int32_t sl = hi - 1; sl >= lo; sl--) freeslots[sl & 15].push_back(sl);   // pop_back = lowest
                int32_t top = lo;
                for (int32_t h : order) {
                    int best = -1; int64_t best_cost = 0; int32_t best_slot = 0;
                    for (int r = 0; r < 16; r++) {
                        if (freeslots[r].empty()) continue;
                        int64_t cost = 0;
                        if (opt_greedy) for (int32_t gj : refs_of(h)) cost += cnt[(size_t)gj * 16 + r];
                        const int32_t sl = freeslots[r].back();
                        // equal cost: keep the staged range compact
                        if (best < 0 || cost < best_cost || (cost == best_cost && sl < best_slot)) {
                            best = r; best_cost = cost; best_slot = sl;
                        }
                    }
                    if (best < 0) return -1;
                    freeslots[best].pop_back();
                    hslot[h] = best_slot;
                    for (int32_t gj : refs_of(h)) cnt[(size_t)gj * 16 + best]++;
                    top = std::max(top, best_slot + 1);
                }
                return top;
            };
            const int32_t nrest = (int32_t)nhalo - npub;
            if (is_struct) {
                // canonical order of the three far faces (hgpu_internal.h)
                for (int z = 0; z < 9; z++) for (int y = 0; y < 9; y++) for (int x = 0; x < 9; x++) {
                    if (x < 8 && y < 8 && z < 8) continue;
                    const int32_t h = -1 - slot_of[sgrid[(z * 9 + y) * 9 + x]];
                    hslot[(size_t)h] = struct_slot(x, y, z);
                }
                npub_slots = STRUCT_HALO;
                nslots = STRUCT_NODES;
            } else {
            // slack of up to 31 slots for the residue choice, as far as the capacities and the
            // slots the other group still needs allow
            const int32_t hi_pub = std::min(std::min(max_acc, max_slots - nrest), (nown + npub + 31) & ~15);
            const int32_t top_pub = assign(0, npub, nown, hi_pub);
            if (top_pub < 0) { err = "internal: no free halo slot"; return false; }
            npub_slots = top_pub - nown;
            const int32_t hi_all = std::min(max_slots, (top_pub + nrest + 31) & ~15);
            const int32_t top_all = nrest > 0 ? assign(npub, (int32_t)nhalo, top_pub, hi_all) : top_pub;
            if (top_all < 0) { err = "internal: no free halo slot"; return false; }
            nslots = top_all;
            }
            // halo list in slot order, -1 marks an unused slot
            // (each tile's range of halo_id -- hence of the partial-force array -- is padded to a
            // multiple of 16 entries = 384 bytes, so that no 128-byte line of partial forces is
            // written by two tiles: a reader may cache lines in L1)
            const size_t base = plan.halo_id.size();
            plan.halo_id.resize(base + (((size_t)(nslots - nown) + 15) & ~(size_t)15), -1);
            for (size_t h = 0; h < halo.size(); h++) {
                plan.halo_id[base + (size_t)(hslot[h] - nown)] = halo[h];
                slot_of[halo[h]] = hslot[h];
            }
        }
        for (int32_t e : elems) {
            plan.elem_id.push_back(e);
            for (int j = 0; j < 8; j++) {
                const int32_t n = lnid[8 * (size_t)e + j];
                const int32_t sl = (n >= a && n < b) ? n - a : slot_of[n];
                plan.elem_slot.push_back((uint16_t)sl);
            }
        }
        if (plan.elem_id.size() > (size_t)INT32_MAX) { err = "tile plan exceeds 2^31 entries"; return false; }
        for (int32_t n = a; n < b; n++) tile_of[n] = tile;
        plan.node_off.push_back(b);
        plan.tile_self.push_back(is_self ? 1 : 0);
        plan.tile_struct.push_back(is_struct ? 1 : 0);
        plan.elem_off.push_back((int32_t)plan.elem_id.size());
        plan.elem_core.push_back(ncore);
        plan.halo_off.push_back((int32_t)plan.halo_id.size());
        plan.halo_pub.push_back(npub_slots);
        want_recs.push_back(nrec); want_srcs.push_back(nsrc);
        plan.max_tile_owned = std::max(plan.max_tile_owned, nown);
        plan.max_tile_acc = std::max(plan.max_tile_acc, nown + npub_slots);
        plan.max_tile_nodes = std::max(plan.max_tile_nodes, nslots);
        plan.halo_nodes_total += (int64_t)halo.size();
        plan.core_total += ncore;
        plan.max_tile_elems = std::max(plan.max_tile_elems, (int32_t)elems.size());
    }
    plan.ntiles = (int32_t)plan.node_off.size() - 1;
    if (plan.core_total != (int64_t)E) { err = "internal: core elements do not partition the element list"; return false; }

    // ---- who reads which published partial force --------------------------------------------------
    // pubs: (node, partial index) of every published slot, sorted by node then by index (= by tile)
    std::vector<std::pair<int32_t, int32_t>> pubs;
    for (int32_t t = 0; t < plan.ntiles; t++)
        for (int32_t h = 0; h < plan.halo_pub[t]; h++) {
            const int32_t p = plan.halo_off[t] + h, n = plan.halo_id[(size_t)p];
            if (n >= 0) pubs.emplace_back(n, p);
        }
    std::sort(pubs.begin(), pubs.end());
    std::vector<int32_t> poff((size_t)N + 1, 0);
    for (const auto &pr : pubs) poff[(size_t)pr.first + 1]++;
    for (int32_t n = 0; n < N; n++) poff[(size_t)n + 1] += poff[n];
    plan.rec_off.push_back(0); plan.src_off.push_back(0); plan.dep_off.push_back(0);
    std::vector<int32_t> deps;
    for (int32_t t = 0; t < plan.ntiles; t++) {
        const int32_t a = plan.node_off[t], b = plan.node_off[(size_t)t + 1];
        const int32_t src0 = (int32_t)plan.src.size();
        deps.clear();
        for (int32_t n = a; n < b; n++) {
            const int32_t cnt = plan.tile_self[t] ? 0 : poff[(size_t)n + 1] - poff[n];
            const bool special = special_node && special_node[n];
            if (!cnt && !special) continue;
            if (cnt > 8) { err = "a node receives partial forces from more than 8 tiles"; return false; }
            FinishRec r;
            r.slot3 = (uint16_t)(3 * (n - a)); r.cnt = (uint8_t)cnt; r.flags = special ? 1 : 0;
            r.first = (int32_t)plan.src.size() - src0;
            plan.rec.push_back(r);
            for (int32_t k = poff[n]; k < poff[n] + cnt; k++) {
                const int32_t p = pubs[(size_t)k].second;
                plan.src.push_back(p);
                const int32_t st = (int32_t)(std::upper_bound(plan.halo_off.begin(), plan.halo_off.end(), p) - plan.halo_off.begin()) - 1;
                deps.push_back(st);
            }
        }
        std::sort(deps.begin(), deps.end());
        deps.erase(std::unique(deps.begin(), deps.end()), deps.end());
        for (int32_t d : deps) {
            if (d >= t) { err = "internal: a tile depends on a higher-numbered tile"; return false; }
            plan.dep.push_back(d);
        }
        plan.rec_off.push_back((int32_t)plan.rec.size());
        plan.src_off.push_back((int32_t)plan.src.size());
        plan.dep_off.push_back((int32_t)plan.dep.size());
        const int32_t nrec = plan.rec_off[(size_t)t + 1] - plan.rec_off[t], nsrc = plan.src_off[(size_t)t + 1] - plan.src_off[t];
        if (nrec != want_recs[t] || nsrc != want_srcs[t]) { err = "internal: finish-record count differs between the two passes"; return false; }
        plan.max_tile_recs = std::max(plan.max_tile_recs, nrec);
        plan.max_tile_srcs = std::max(plan.max_tile_srcs, nsrc);
    }
    return true;
}

// Independent check of a plan against the mesh: every node is owned by exactly one tile, every
// element is the core element of exactly one tile, every slot decodes to the element's own corner
// node, and the force of every owned node is complete: the tile's own entries plus the partial
// forces it reads account for each incident element exactly once.
bool validate_tile_plan(int32_t E, int32_t N, const int32_t *lnid, const TilePlan &pl, std::string &err)
{
    const int32_t T = pl.ntiles;
    if (T < 0 || (int32_t)pl.node_off.size() != T + 1) { err = "node_off size"; return false; }
    if (pl.node_off.front() != 0 || pl.node_off.back() != N) { err = "tiles do not cover the node range"; return false; }
    if ((int32_t)pl.tile_self.size() != T || (int32_t)pl.elem_core.size() != T || (int32_t)pl.halo_pub.size() != T ||
        (int32_t)pl.rec_off.size() != T + 1 || (int32_t)pl.src_off.size() != T + 1 || (int32_t)pl.dep_off.size() != T + 1) {
        err = "per-tile table size"; return false;
    }
    std::vector<int32_t> degree((size_t)N, 0), seen((size_t)N, 0), stamp((size_t)N, -1), stamp_slot((size_t)N, -1);
    std::vector<int32_t> core_of((size_t)E, -1);
    std::vector<int32_t> pcount(pl.halo_id.size(), 0);     // core entries adding to each published slot
    for (int32_t e = 0; e < E; e++) for (int j = 0; j < 8; j++) degree[lnid[8 * (size_t)e + j]]++;
    for (int32_t t = 0; t < T; t++) {
        const int32_t a = pl.node_off[t], b = pl.node_off[(size_t)t + 1];
        if (b <= a || (a & 1)) { err = "empty or odd-aligned tile"; return false; }
        const int32_t nown = b - a, hb = pl.halo_off[t], nh = pl.halo_off[(size_t)t + 1] - hb, npub = pl.halo_pub[t];
        if (nown > pl.max_tile_owned || nown + nh > pl.max_tile_nodes + 15 || nown + npub > pl.max_tile_acc || npub > nh) {
            err = "tile exceeds recorded maxima"; return false;
        }
        for (int32_t h = 0; h < nh; h++) {
            const int32_t n = pl.halo_id[(size_t)hb + h];
            if (n < -1 || n >= N) { err = "halo id out of range"; return false; }
            if (n >= a && n < b) { err = "owned node listed as halo"; return false; }
        }
        const int32_t eb = pl.elem_off[t], ne = pl.elem_off[(size_t)t + 1] - eb, ncore = pl.elem_core[t];
        if (ncore > ne || (!pl.tile_self[t] && ncore != ne)) { err = "a shared tile lists extra entries"; return false; }
        for (int32_t k = 0; k < ne; k++) {
            const int32_t e = pl.elem_id[(size_t)eb + k];
            if (e < 0 || e >= E) { err = "element id out of range"; return false; }
            const bool core = k < ncore;
            if (core) {
                if (core_of[e] >= 0) { err = "element is the core element of two tiles"; return false; }
                core_of[e] = t;
            }
            bool touches = false;
            for (int j = 0; j < 8; j++) {
                const int32_t sl = pl.elem_slot[8 * (size_t)(eb + k) + j];
                if (sl >= nown + nh) { err = "slot out of range"; return false; }
                const int32_t n = sl < nown ? a + sl : pl.halo_id[(size_t)hb + sl - nown];
                if (n != lnid[8 * (size_t)e + j]) { err = "slot decodes to the wrong node"; return false; }
                if (sl < nown) { seen[n]++; touches = true; }
                else {
                    if (stamp[n] == t && stamp_slot[n] != sl) { err = "halo node staged twice"; return false; }
                    stamp[n] = t; stamp_slot[n] = sl;
                    if (core) {
                        if (sl >= nown + npub) { err = "core element touches an unpublished halo slot"; return false; }
                        pcount[(size_t)hb + sl - nown]++;
                    }
                }
            }
            if (core && (lnid[8 * (size_t)e] < a || lnid[8 * (size_t)e] >= b)) { err = "core element whose corner 0 is not owned"; return false; }
            if (!pl.tile_struct.empty() && pl.tile_struct[t]) {
                // canonical layout: entry k is the element at de-interleave(k), corner j in slot struct_slot(...)
                if (nown != STRUCT_OWNED || ne != STRUCT_OWNED || npub != STRUCT_HALO || nh < STRUCT_HALO) { err = "structured tile with the wrong counts"; return false; }
                int x = 0, y = 0, z = 0;
                for (int bt = 0; bt < 3; bt++) { x |= ((k >> (3 * bt)) & 1) << bt; y |= ((k >> (3 * bt + 1)) & 1) << bt; z |= ((k >> (3 * bt + 2)) & 1) << bt; }
                for (int j = 0; j < 8; j++)
                    if (pl.elem_slot[8 * (size_t)(eb + k) + j] != struct_slot(x + (j & 1), y + ((j >> 1) & 1), z + (j >> 2))) {
                        err = "structured tile whose slots are not canonical"; return false;
                    }
            }
            if (!core && !touches) { err = "tile evaluates an extra element that touches none of its nodes"; return false; }
        }
    }
    for (int32_t e = 0; e < E; e++) if (core_of[e] < 0) { err = "element " + std::to_string(e) + " is nobody's core element"; return false; }
    for (int32_t t = 0; t < T; t++) {
        const int32_t a = pl.node_off[t], b = pl.node_off[(size_t)t + 1];
        const int32_t sb = pl.src_off[t], ns = pl.src_off[(size_t)t + 1] - sb;
        int32_t prev_slot = -1;
        for (int32_t r = pl.rec_off[t]; r < pl.rec_off[(size_t)t + 1]; r++) {
            const FinishRec &fr = pl.rec[(size_t)r];
            if (fr.slot3 % 3 || fr.slot3 / 3 >= b - a || (int32_t)fr.slot3 <= prev_slot) { err = "finish record slot"; return false; }
            prev_slot = fr.slot3;
            if (fr.first < 0 || fr.first + fr.cnt > ns || fr.cnt > 8) { err = "finish record src range"; return false; }
            const int32_t n = a + fr.slot3 / 3;
            int32_t last_tile = -1;
            for (int32_t k = 0; k < fr.cnt; k++) {
                const int32_t p = pl.src[(size_t)sb + fr.first + k];
                if (p < 0 || p >= (int32_t)pl.halo_id.size() || pl.halo_id[(size_t)p] != n) { err = "partial force of another node"; return false; }
                const int32_t st = (int32_t)(std::upper_bound(pl.halo_off.begin(), pl.halo_off.end(), p) - pl.halo_off.begin()) - 1;
                if (st <= last_tile || st >= t) { err = "partial forces not in ascending lower-tile order"; return false; }
                last_tile = st;
                if (p - pl.halo_off[st] >= pl.halo_pub[st]) { err = "reads a slot its tile does not publish"; return false; }
                if (!std::binary_search(pl.dep.begin() + pl.dep_off[t], pl.dep.begin() + pl.dep_off[(size_t)t + 1], st)) {
                    err = "source tile missing from the dependency list"; return false;
                }
                seen[n] += pcount[(size_t)p];
            }
        }
    }
    for (int32_t n = 0; n < N; n++)
        if (seen[n] != degree[n]) { err = "node " + std::to_string(n) + " misses incident elements"; return false; }
    return true;
}

// Shared-memory wavefronts per 8-byte access instruction of the step kernel under the bank model
// "16 lanes per wavefront, 16 banks of 8 bytes": gather = the 8 corner reads of every entry,
// scatter = the accumulator updates (all corners of a core entry, owned corners of an extra one).
// 2.0 per 32-lane instruction is ideal.
void estimate_wavefronts(const TilePlan &pl, double *gather, double *scatter)
{
    int64_t gw = 0, gi = 0, sw = 0, si = 0;
    for (int32_t t = 0; t < pl.ntiles; t++) {
        const int32_t nown = pl.node_off[(size_t)t + 1] - pl.node_off[t];
        const int32_t eb = pl.elem_off[t], ne = pl.elem_off[(size_t)t + 1] - eb, ncore = pl.elem_core[t];
        for (int32_t w0 = 0; w0 < ne; w0 += 32) {
            for (int j = 0; j < 8; j++) {
                int any_acc = 0;
                for (int half = 0; half < 2; half++) {
                    int c_all[16] = {0}, c_acc[16] = {0};
                    for (int l = 0; l < 16; l++) {
                        const int32_t k = w0 + 16 * half + l;
                        if (k >= ne) break;
                        const int32_t sl = pl.elem_slot[8 * (size_t)(eb + k) + j];
                        c_all[sl & 15]++;
                        if (k < ncore || sl < nown) { c_acc[sl & 15]++; any_acc = 1; }
                    }
                    int ma = 0, mo = 0;
                    for (int r = 0; r < 16; r++) { ma = std::max(ma, c_all[r]); mo = std::max(mo, c_acc[r]); }
                    gw += ma; sw += mo;
                }
                gi++; si += any_acc;
            }
        }
    }
    *gather = gi ? (double)gw / (double)gi : 0.0;
    *scatter = si ? (double)sw / (double)si : 0.0;
}

bool build_dangling_plan(int32_t N, int32_t D, const int32_t *dnode, DanglingPlan &plan,
                         std::string &err)
{
    plan = DanglingPlan();
    std::vector<uint8_t> is_dangling((size_t)N, 0);
    for (int32_t d = 0; d < D; d++) {
        const int32_t *dn = dnode + 6 * (size_t)d;
        if (dn[0] < 0 || dn[0] >= N) { err = "dangling node id out of range"; return false; }
        if (dn[1] != 2 && dn[1] != 4) { err = "dangling node deps must be 2 or 4 (octor.h:155)"; return false; }
        int cnt = 0;
        for (int a = 0; a < 4; a++) {
            if (dn[2 + a] < 0) break;
            if (dn[2 + a] >= N) { err = "anchor id out of range"; return false; }
            cnt++;
        }
        if (cnt != dn[1]) { err = "dangling node anchor count differs from deps (psolve.c:5979-5986)"; return false; }
        is_dangling[dn[0]] = 1;
    }
    // (anchor, dnode index) pairs, sorted by anchor then by dnode-table order
    std::vector<std::pair<int32_t, int32_t>> pairs;
    for (int32_t d = 0; d < D; d++) {
        const int32_t *dn = dnode + 6 * (size_t)d;
        for (int a = 0; a < dn[1]; a++) {
            if (is_dangling[dn[2 + a]]) {
                err = "an anchor of a dangling node is itself dangling; the anchor-centric "
                      "distribution would not reproduce the reference's sequential order";
                return false;
            }
            pairs.emplace_back(dn[2 + a], d);
        }
    }
    std::stable_sort(pairs.begin(), pairs.end());
    plan.anchor_off.push_back(0);
    for (size_t i = 0; i < pairs.size(); i++) {
        if (i == 0 || pairs[i].first != pairs[i - 1].first) {
            if (i) plan.anchor_off.push_back((int32_t)i);
            plan.anchor_id.push_back(pairs[i].first);
        }
        const int32_t *dn = dnode + 6 * (size_t)pairs[i].second;
        plan.anchor_dn.push_back(dn[0]);
        plan.anchor_deps.push_back(dn[1]);
    }
    if (!pairs.empty()) plan.anchor_off.push_back((int32_t)pairs.size());
    return true;
}

}  // namespace hgpu
