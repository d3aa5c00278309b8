// hgpu_tiles.cpp -- host-side index builders for libhercules_gpu.so (no CUDA in this file).
//
// build_tile_plan     owner-computes tiling that replaces the scatter-add of
//                     compute_addforce_effective / damping_addforce (stiffness.c:228-235,
//                     damping.c:88-98) with an atomic-free gather: see DESIGN.md section 3.
// build_dangling_plan anchor-centric CSR for compute_adjust(DISTRIBUTION) (psolve.c:5943-5987).
#include "hgpu_internal.h"

#include <algorithm>
#include <cstring>

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
        for (size_t h = 0; h < halo.size(); h++) slot_of[halo[h]] = nown + (int32_t)h;
        for (int32_t e : elems) {
            plan.elem_id.push_back(e);
            for (int j = 0; j < 8; j++) {
                const int32_t n = lnid[8 * (size_t)e + j];
                const int32_t sl = (n >= a && n < b) ? n - a : slot_of[n];
                plan.elem_slot.push_back((uint16_t)sl);
            }
        }
        if (plan.elem_id.size() > (size_t)INT32_MAX) { err = "tile plan exceeds 2^31 entries"; return false; }
        plan.halo_id.insert(plan.halo_id.end(), halo.begin(), halo.end());
        plan.node_off.push_back(b);
        plan.elem_off.push_back((int32_t)plan.elem_id.size());
        plan.halo_off.push_back((int32_t)plan.halo_id.size());
        plan.max_tile_owned = std::max(plan.max_tile_owned, nown);
        plan.max_tile_nodes = std::max(plan.max_tile_nodes, nown + (int32_t)halo.size());
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
    std::vector<int32_t> degree((size_t)N, 0), seen((size_t)N, 0);
    for (int32_t e = 0; e < E; e++) for (int j = 0; j < 8; j++) degree[lnid[8 * (size_t)e + j]]++;
    for (int32_t t = 0; t < pl.ntiles; t++) {
        const int32_t a = pl.node_off[t], b = pl.node_off[(size_t)t + 1];
        if (b <= a || (a & 1)) { err = "empty or odd-aligned tile"; return false; }
        const int32_t nown = b - a, hb = pl.halo_off[t], nh = pl.halo_off[(size_t)t + 1] - hb;
        if (nown > pl.max_tile_owned || nown + nh > pl.max_tile_nodes) { err = "tile exceeds recorded maxima"; return false; }
        for (int32_t h = 0; h < nh; h++) {
            const int32_t n = pl.halo_id[(size_t)hb + h];
            if (n >= a && n < b) { err = "owned node listed as halo"; return false; }
            if (h && n <= pl.halo_id[(size_t)hb + h - 1]) { err = "halo list not ascending"; return false; }
        }
        for (int32_t k = pl.elem_off[t]; k < pl.elem_off[(size_t)t + 1]; k++) {
            const int32_t e = pl.elem_id[k];
            if (k > pl.elem_off[t] && e <= pl.elem_id[(size_t)k - 1]) { err = "element list not ascending"; return false; }
            bool touches = false;
            for (int j = 0; j < 8; j++) {
                const int32_t sl = pl.elem_slot[8 * (size_t)k + j];
                if (sl >= nown + nh) { err = "slot out of range"; return false; }
                const int32_t n = sl < nown ? a + sl : pl.halo_id[(size_t)hb + sl - nown];
                if (n != lnid[8 * (size_t)e + j]) { err = "slot decodes to the wrong node"; return false; }
                if (sl < nown) { seen[n]++; touches = true; }
            }
            if (!touches) { err = "tile evaluates an element that touches none of its nodes"; return false; }
        }
    }
    for (int32_t n = 0; n < N; n++)
        if (seen[n] != degree[n]) { err = "node " + std::to_string(n) + " misses incident elements"; return false; }
    return true;
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
