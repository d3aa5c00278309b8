// hgpu_internal.h -- private declarations shared by the host-side builders and the CUDA
// translation unit of libhercules_gpu.so.  Public ABI: include/hercules_gpu.h.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "hercules_gpu.h"

namespace hgpu {

// Node classes.  REGULAR nodes are advanced inside the fused tile kernel; SPECIAL nodes go through
// the force array and the reference's adjust/exchange/update sequence.
enum : uint8_t { NODE_REGULAR = 0, NODE_SPECIAL = 1 };

// Owner-computes tiling of the node range.  Tile t owns the contiguous node
// ids [node_off[t], node_off[t+1]) (even start); it evaluates every element incident to an owned
// node, so the force on an owned node is complete inside the tile and no atomics are needed.
struct TilePlan {
    int32_t tile_nodes = 0;       // cap on owned nodes per tile
    int32_t ntiles = 0;
    int32_t max_tile_owned = 0;
    int32_t max_tile_nodes = 0;   // owned + gathered
    int32_t max_tile_elems = 0;
    std::vector<int32_t>  node_off;   // [ntiles+1]
    std::vector<int32_t>  elem_off;   // [ntiles+1] into elem_id / elem_slot
    std::vector<int32_t>  elem_id;    // element evaluated by this tile entry
    std::vector<uint16_t> elem_slot;  // [entries][8] tile-local slot of each corner node
    std::vector<int32_t>  halo_off;   // [ntiles+1] into halo_id
    std::vector<int32_t>  halo_id;    // gathered (non-owned) node ids; slot = owned_count + index; -1 = unused slot
    int64_t halo_nodes_total = 0;     // gathered nodes over all tiles (unused slots not counted)
};

// Builds the plan; returns false and sets err on inconsistent input.
bool build_tile_plan(int32_t E, int32_t N, const int32_t *lnid, int32_t elem_block,
                     int32_t max_owned, int32_t max_slots, TilePlan &plan, std::string &err);

bool validate_tile_plan(int32_t E, int32_t N, const int32_t *lnid, const TilePlan &plan, std::string &err);

void estimate_wavefronts(const TilePlan &plan, double *gather, double *scatter);

// Hanging-node lists (flattened dnode_t, octor.h:153-158).
struct DanglingPlan {
    // anchor-centric CSR for the DISTRIBUTION pass: anchor a receives force[dn]/deps from every
    // dangling node that lists it, in ascending dnode-table order (the reference's order,
    // psolve.c:5943-5987), which makes the sum bit-reproducible.
    std::vector<int32_t> anchor_id;     // [nA] anchored node ids (ascending)
    std::vector<int32_t> anchor_off;    // [nA+1]
    std::vector<int32_t> anchor_dn;     // dangling node id of each contribution
    std::vector<int32_t> anchor_deps;   // its deps
};
bool build_dangling_plan(int32_t N, int32_t D, const int32_t *dnode, DanglingPlan &plan,
                         std::string &err);

}  // namespace hgpu
