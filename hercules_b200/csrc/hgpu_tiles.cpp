// hgpu_tiles.cpp -- host-side index builders for libhercules_gpu.so (no CUDA in this file).
//
// build_tile_plan     owner-computes tiling that replaces the scatter-add of
//                     compute_addforce_effective / damping_addforce (stiffness.c:228-235,
//                     damping.c:88-98) with an atomic-free gather: see DESIGN.md section 3.
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
// 8x8x8 cell of a uniform region when elem_block = 512): cutting the node range where that block
// index changes keeps the gathered halo (and the elements evaluated twice) at the geometric
// minimum, whatever the refinement pattern is.  A tile that would need more than max_owned owned
// nodes or max_slots staged nodes is split in half until it fits.
bool build_tile_plan(int32_t E, int32_t N, const int32_t *lnid, int32_t elem_block,
                     int32_t max_owned, int32_t max_slots, TilePlan &plan, std::string &err)
{
    if (elem_block <= 0 || max_owned < 2 || max_slots < 16) { err = "bad tile limits"; return false; }
    max_owned &= ~1;
    plan = TilePlan();
    plan.tile_nodes = max_owned;

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
    std::vector<int32_t> cuts;
    cuts.push_back(0);
    {
        // cuts fall on even node ids only (16-byte aligned bulk copies of the owned range); when
        // the block changes at an odd id that one node stays with the tile before it
        int32_t cur_blk = -1, last_blk = -1, start = 0;
        for (int32_t n = 0; n < N; n++) {
            const int32_t blk = noff[(size_t)n + 1] > noff[n]
                                    ? nelem[(size_t)noff[(size_t)n + 1] - 1] / elem_block : last_blk;
            last_blk = blk;
            if (n == start) { cur_blk = blk; continue; }
            if (!(n & 1) && (blk != cur_blk || n - start >= max_owned)) {
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
    std::vector<int32_t> elems, halo;
    plan.node_off.push_back(0);
    plan.elem_off.push_back(0);
    plan.halo_off.push_back(0);
    // worklist of [a, b) ranges, processed in order (split ranges are re-queued in place)
    std::vector<std::pair<int32_t, int32_t>> work;
    for (size_t i = cuts.size() - 1; i > 0; i--) work.emplace_back(cuts[i - 1], cuts[i]);
    int32_t tile = 0;
    while (!work.empty()) {
        const int32_t a = work.back().first, b = work.back().second;
        work.pop_back();
        if (a >= b) continue;
        const int32_t nown = b - a;
        bool fits = nown <= max_owned;
        elems.clear(); halo.clear();
        if (fits) {
            for (int32_t n = a; n < b; n++)
                for (int32_t k = noff[n]; k < noff[(size_t)n + 1]; k++) {
                    const int32_t e = nelem[k];
                    if (stamp_e[e] != tile) { stamp_e[e] = tile; elems.push_back(e); }
                }
            std::sort(elems.begin(), elems.end());
            for (int32_t e : elems)
                for (int j = 0; j < 8; j++) {
                    const int32_t n = lnid[8 * (size_t)e + j];
                    if (n >= a && n < b) continue;
                    if (stamp_n[n] != tile) { stamp_n[n] = tile; halo.push_back(n); }
                }
            fits = (int64_t)nown + (int64_t)halo.size() <= (int64_t)max_slots;
        }
        if (!fits) {
            if (nown <= 2) { err = "a 2-node tile exceeds the staging capacity"; return false; }
            const int32_t mid = a + ((nown / 2 + 1) & ~1);
            // invalidate the stamps used by this attempt
            tile++;
            work.emplace_back(mid, b);
            work.emplace_back(a, mid);
            continue;
        }
        std::sort(halo.begin(), halo.end());
        // Entry order.  Threads take consecutive entries, and a shared-memory access of 16 lanes
        // (one half-warp of 8-byte words) is conflict-free when the 16 slots differ modulo 16.
        // "Core" entries -- corner 0 is an owned node -- sorted by that node are, in a uniform
        // region, the tile's own Morton cell in Morton order: every aligned run of 16 is a 4x2x2
        // block whose corner-j nodes have 16 different slot residues.  The remaining entries (the
        // layers shared with lower neighbours) follow, sorted by their lowest owned corner.
        {
            auto key = [&](int32_t e) -> int64_t {
                const int32_t *ln = lnid + 8 * (size_t)e;
                if (ln[0] >= a && ln[0] < b) return (int64_t)(ln[0] - a);
                int32_t lo = INT32_MAX;
                for (int j = 1; j < 8; j++) if (ln[j] >= a && ln[j] < b) lo = std::min(lo, ln[j] - a);
                return ((int64_t)1 << 32) + lo;
            };
            if (opt_sort) std::stable_sort(elems.begin(), elems.end(), [&](int32_t x, int32_t y) { return key(x) < key(y); });
            // Non-core entries touch owned nodes on one face / edge of the patch only, whose slots
            // share residues: pack them into half-warp groups of 16 greedily so that, corner by
            // corner, the owned slots of a group collide as little as possible.  The packing
            // depends only on the residue pattern, which repeats from tile to tile: memoised.
            int32_t ncore = 0;
            for (int32_t e : elems) { const int32_t n0c = lnid[8 * (size_t)e]; if (n0c >= a && n0c < b) ncore++; }
            const int32_t nrest = (int32_t)elems.size() - ncore;
            if (opt_sort && opt_pack && nrest > 1) {
                // signature: residues (or 255) of the 8 corners of every non-core entry + the partial group
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
        int32_t nslots = nown + (int32_t)halo.size();
        {
            const int32_t ne = (int32_t)elems.size(), ngrp = (ne + 15) / 16;
            const int32_t limit = std::min(max_slots, (nown + (int32_t)halo.size() + 31) & ~15);
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
            std::vector<int32_t> &order = g_order;
            order.resize(nhalo);
            for (size_t h = 0; h < nhalo; h++) order[h] = (int32_t)h;
            std::stable_sort(order.begin(), order.end(),
                             [&](int32_t x, int32_t y) { return refs_of(x).size() > refs_of(y).size(); });
            for (int r = 0; r < 16; r++) freeslots[r].clear();
            for (int32_t sl = limit - 1; sl >= nown; sl--) freeslots[sl & 15].push_back(sl);   // pop_back = lowest
            std::vector<int32_t> &hslot = g_hslot;
            hslot.assign(nhalo, -1);
            nslots = nown;
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
                if (best < 0) { err = "internal: no free halo slot"; return false; }
                freeslots[best].pop_back();
                hslot[h] = best_slot;
                for (int32_t gj : refs_of(h)) cnt[(size_t)gj * 16 + best]++;
                nslots = std::max(nslots, best_slot + 1);
            }
            // halo list in slot order, -1 marks an unused slot
            const size_t base = plan.halo_id.size();
            plan.halo_id.resize(base + (size_t)(nslots - nown), -1);
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
        plan.node_off.push_back(b);
        plan.elem_off.push_back((int32_t)plan.elem_id.size());
        plan.halo_off.push_back((int32_t)plan.halo_id.size());
        plan.max_tile_owned = std::max(plan.max_tile_owned, nown);
        plan.max_tile_nodes = std::max(plan.max_tile_nodes, nslots);
        plan.halo_nodes_total += (int64_t)halo.size();
        plan.max_tile_elems = std::max(plan.max_tile_elems, (int32_t)elems.size());
        tile++;
    }
    plan.ntiles = (int32_t)plan.node_off.size() - 1;
    return true;
}

// Independent check of a plan against the mesh: every node is owned by exactly one tile, every
// element incident to an owned node is evaluated by that tile exactly once, and every slot decodes
// to the element's own corner node.
bool validate_tile_plan(int32_t E, int32_t N, const int32_t *lnid, const TilePlan &pl, std::string &err)
{
    if (pl.ntiles < 0 || (int32_t)pl.node_off.size() != pl.ntiles + 1) { err = "node_off size"; return false; }
    if (pl.node_off.front() != 0 || pl.node_off.back() != N) { err = "tiles do not cover the node range"; return false; }
    std::vector<int32_t> degree((size_t)N, 0), seen((size_t)N, 0), stamp((size_t)N, -1), stamp_slot((size_t)N, -1);
    for (int32_t e = 0; e < E; e++) for (int j = 0; j < 8; j++) degree[lnid[8 * (size_t)e + j]]++;
    for (int32_t t = 0; t < pl.ntiles; t++) {
        const int32_t a = pl.node_off[t], b = pl.node_off[(size_t)t + 1];
        if (b <= a || (a & 1)) { err = "empty or odd-aligned tile"; return false; }
        const int32_t nown = b - a, hb = pl.halo_off[t], nh = pl.halo_off[(size_t)t + 1] - hb;
        if (nown > pl.max_tile_owned || nown + nh > pl.max_tile_nodes) { err = "tile exceeds recorded maxima"; return false; }
        for (int32_t h = 0; h < nh; h++) {
            const int32_t n = pl.halo_id[(size_t)hb + h];
            if (n < -1 || n >= N) { err = "halo id out of range"; return false; }
            if (n >= a && n < b) { err = "owned node listed as halo"; return false; }
        }
        for (int32_t k = pl.elem_off[t]; k < pl.elem_off[(size_t)t + 1]; k++) {
            const int32_t e = pl.elem_id[k];
            if (e < 0 || e >= E) { err = "element id out of range"; return false; }
            bool touches = false;
            for (int j = 0; j < 8; j++) {
                const int32_t sl = pl.elem_slot[8 * (size_t)k + j];
                if (sl >= nown + nh) { err = "slot out of range"; return false; }
                const int32_t n = sl < nown ? a + sl : pl.halo_id[(size_t)hb + sl - nown];
                if (n != lnid[8 * (size_t)e + j]) { err = "slot decodes to the wrong node"; return false; }
                if (sl < nown) { seen[n]++; touches = true; }
                else if (stamp[n] == t && stamp_slot[n] != sl) { err = "halo node staged twice"; return false; }
                else { stamp[n] = t; stamp_slot[n] = sl; }
            }
            if (!touches) { err = "tile evaluates an element that touches none of its nodes"; return false; }
        }
    }
    for (int32_t n = 0; n < N; n++)
        if (seen[n] != degree[n]) { err = "node " + std::to_string(n) + " misses incident elements"; return false; }
    return true;
}

// Shared-memory wavefronts per 8-byte access instruction of the step kernel under the bank model
// "16 lanes per wavefront, 16 banks of 8 bytes": gather = the 8 corner reads of every entry,
// scatter = the accumulator updates of owned corners.  2.0 per 32-lane instruction is ideal.
void estimate_wavefronts(const TilePlan &pl, double *gather, double *scatter)
{
    int64_t gw = 0, gi = 0, sw = 0, si = 0;
    const char *dbg = getenv("HGPU_PLAN_DEBUG");
    bool printed = false;
    for (int32_t t = 0; t < pl.ntiles; t++) {
        const int32_t nown = pl.node_off[(size_t)t + 1] - pl.node_off[t];
        const int32_t eb = pl.elem_off[t], ne = pl.elem_off[(size_t)t + 1] - eb;
        if (dbg && atoi(dbg) > 0 && ne != atoi(dbg)) continue;      // only tiles with that many entries
        if (t > 0 && gi > 0) printed = true;
        for (int32_t w0 = 0; w0 < ne; w0 += 32) {
            for (int j = 0; j < 8; j++) {
                int any_owned = 0;
                for (int half = 0; half < 2; half++) {
                    int c_all[16] = {0}, c_own[16] = {0};
                    for (int l = 0; l < 16; l++) {
                        const int32_t k = w0 + 16 * half + l;
                        if (k >= ne) break;
                        const int32_t sl = pl.elem_slot[8 * (size_t)(eb + k) + j];
                        c_all[sl & 15]++;
                        if (sl < nown) { c_own[sl & 15]++; any_owned = 1; }
                    }
                    int ma = 0, mo = 0;
                    for (int r = 0; r < 16; r++) { ma = std::max(ma, c_all[r]); mo = std::max(mo, c_own[r]); }
                    gw += ma; sw += mo;
                }
                gi++; si += any_owned;
            }
            if (dbg && atoi(dbg) > 0 && !printed) {
                fprintf(stderr, "tile %d warp %d: per corner gather:", t, w0 / 32);
                for (int j = 0; j < 8; j++) {
                    int tot = 0;
                    for (int half = 0; half < 2; half++) {
                        int c_all[16] = {0};
                        for (int l = 0; l < 16; l++) { const int32_t k = w0 + 16 * half + l; if (k >= ne) break; c_all[pl.elem_slot[8 * (size_t)(eb + k) + j] & 15]++; }
                        int ma = 0; for (int r = 0; r < 16; r++) ma = std::max(ma, c_all[r]);
                        tot += ma;
                    }
                    fprintf(stderr, " %d", tot);
                }
                fprintf(stderr, "\n");
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
