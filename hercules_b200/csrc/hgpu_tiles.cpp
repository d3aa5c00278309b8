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

bool build_tile_plan(int32_t E, int32_t N, const int32_t *lnid, int32_t tile_nodes,
                     int32_t max_slots, TilePlan &plan, std::string &err)
{
    if (tile_nodes <= 0 || (tile_nodes & 1)) { err = "tile_nodes must be positive and even"; return false; }
    const int32_t T = N > 0 ? (N + tile_nodes - 1) / tile_nodes : 0;
    plan = TilePlan();
    plan.tile_nodes = tile_nodes;
    plan.ntiles = T;
    plan.elem_off.assign((size_t)T + 1, 0);
    plan.halo_off.assign((size_t)T + 1, 0);

    // pass 1: an element is evaluated by every tile that owns one of its 8 corner nodes
    for (int32_t e = 0; e < E; e++) {
        const int32_t *ln = lnid + 8 * (size_t)e;
        int32_t seen[8]; int ns = 0;
        for (int j = 0; j < 8; j++) {
            if (ln[j] < 0 || ln[j] >= N) { err = "element node id out of range"; return false; }
            int32_t t = ln[j] / tile_nodes;
            bool dup = false;
            for (int k = 0; k < ns; k++) dup |= (seen[k] == t);
            if (!dup) { seen[ns++] = t; plan.elem_off[(size_t)t + 1]++; }
        }
    }
    for (int32_t t = 0; t < T; t++) plan.elem_off[(size_t)t + 1] += plan.elem_off[t];
    const size_t entries = T ? (size_t)plan.elem_off[T] : 0;
    if (entries > (size_t)INT32_MAX) { err = "tile plan exceeds 2^31 entries"; return false; }
    plan.elem_id.resize(entries);
    plan.elem_slot.resize(entries * 8);

    // pass 2: fill element lists (ascending element id inside each tile)
    {
        std::vector<int32_t> cursor(plan.elem_off.begin(), plan.elem_off.end() - (T ? 1 : 0));
        for (int32_t e = 0; e < E; e++) {
            const int32_t *ln = lnid + 8 * (size_t)e;
            int32_t seen[8]; int ns = 0;
            for (int j = 0; j < 8; j++) {
                int32_t t = ln[j] / tile_nodes;
                bool dup = false;
                for (int k = 0; k < ns; k++) dup |= (seen[k] == t);
                if (!dup) { seen[ns++] = t; plan.elem_id[(size_t)cursor[t]++] = e; }
            }
        }
    }

    // pass 3: per tile, gathered-node list (ascending id) and the 8 local slots of every entry
    std::vector<int32_t> stamp((size_t)N, -1), slot_of((size_t)N, 0);
    std::vector<int32_t> halo;
    for (int32_t t = 0; t < T; t++) {
        const int32_t n0 = t * tile_nodes;
        const int32_t nown = std::min(tile_nodes, N - n0);
        const int32_t b = plan.elem_off[t], en = plan.elem_off[(size_t)t + 1];
        halo.clear();
        for (int32_t k = b; k < en; k++) {
            const int32_t *ln = lnid + 8 * (size_t)plan.elem_id[k];
            for (int j = 0; j < 8; j++) {
                int32_t n = ln[j];
                if (n >= n0 && n < n0 + nown) continue;
                if (stamp[n] != t) { stamp[n] = t; halo.push_back(n); }
            }
        }
        std::sort(halo.begin(), halo.end());
        if ((int64_t)nown + (int64_t)halo.size() > (int64_t)max_slots) {
            err = "a tile needs " + std::to_string(nown + halo.size()) + " node slots (limit " +
                  std::to_string(max_slots) + "); use a smaller tile_nodes";
            return false;
        }
        for (size_t h = 0; h < halo.size(); h++) slot_of[halo[h]] = nown + (int32_t)h;
        for (int32_t k = b; k < en; k++) {
            const int32_t *ln = lnid + 8 * (size_t)plan.elem_id[k];
            for (int j = 0; j < 8; j++) {
                int32_t n = ln[j];
                int32_t s = (n >= n0 && n < n0 + nown) ? n - n0 : slot_of[n];
                plan.elem_slot[8 * (size_t)k + j] = (uint16_t)s;
            }
        }
        plan.halo_off[(size_t)t + 1] = plan.halo_off[t] + (int32_t)halo.size();
        plan.halo_id.insert(plan.halo_id.end(), halo.begin(), halo.end());
        plan.max_tile_nodes = std::max(plan.max_tile_nodes, nown + (int32_t)halo.size());
        plan.max_tile_elems = std::max(plan.max_tile_elems, en - b);
    }
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
