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

// One owned node of a tile that needs more than "advance what the tile accumulated": it receives
// partial forces published by other tiles (cnt of them, src[first .. first+cnt) relative to the
// tile's src range) and / or it is a SPECIAL node whose force is handed to the force array.
struct FinishRec {
    uint16_t slot3;     // 3 * tile-local slot of the owned node
    uint8_t  cnt;       // incoming partial forces (0..8)
    uint8_t  flags;     // bit 0: SPECIAL node
    int32_t  first;     // into the tile's src range
};
static_assert(sizeof(FinishRec) == 8, "FinishRec is staged as 8-byte records");

// Tiling of the mesh (DESIGN.md sections 3-4).  Tile t OWNS the contiguous node ids
// [node_off[t], node_off[t+1]) (even start) and is the one tile that evaluates its CORE elements
// -- those whose corner 0 it owns -- so every element is evaluated exactly once.  What a core
// element adds to nodes of other tiles is PUBLISHED as a per-(tile, node) partial force and summed
// by the owner ("shared" tiles).  A "self" tile does not wait for anybody: it additionally
// evaluates the foreign elements incident to its nodes (EXTRA entries, accumulated on owned nodes
// only) and ignores what others publish for it; tiles whose forces the halo exchange needs first
// are built that way.
struct TilePlan {
    int32_t ntiles = 0;
    int32_t max_tile_owned = 0;
    int32_t max_tile_acc = 0;     // owned + publishable halo slots
    int32_t max_tile_nodes = 0;   // all staged slots
    int32_t max_tile_elems = 0;
    int32_t max_tile_recs = 0, max_tile_srcs = 0;
    std::vector<int32_t>  node_off;   // [ntiles+1]
    std::vector<uint8_t>  tile_self;  // [ntiles]
    std::vector<uint8_t>  tile_struct; // [ntiles] 1: an aligned 8x8x8 cell of equal elements in canonical layout (see struct_*)
    std::vector<int32_t>  elem_off;   // [ntiles+1] into elem_id / elem_slot
    std::vector<int32_t>  elem_core;  // [ntiles] the first elem_core[t] entries of a tile are its core elements
    std::vector<int32_t>  elem_id;    // element evaluated by this tile entry
    std::vector<uint16_t> elem_slot;  // [entries][8] tile-local slot of each corner node
    std::vector<int32_t>  halo_off;   // [ntiles+1] into halo_id; also the numbering of published partial forces
    std::vector<int32_t>  halo_pub;   // [ntiles] the first halo_pub[t] halo slots are published
    std::vector<int32_t>  halo_id;    // gathered (non-owned) node ids; slot = owned_count + index; -1 = unused slot
    std::vector<int32_t>  rec_off;    // [ntiles+1]
    std::vector<FinishRec> rec;
    std::vector<int32_t>  src_off;    // [ntiles+1]
    std::vector<int32_t>  src;        // index (halo_id numbering) of a published partial force
    std::vector<int32_t>  dep_off;    // [ntiles+1]
    std::vector<int32_t>  dep;        // tiles whose partial forces this tile reads (all lower-numbered)
    int64_t halo_nodes_total = 0;     // gathered nodes over all tiles (unused slots not counted)
    int64_t core_total = 0;           // = number of elements
};

struct TileCaps {
    int32_t elem_block;   // elements per block of the Morton-ordered element list that seeds a tile
    int32_t max_owned;    // owned nodes per tile
    int32_t max_acc;      // owned + published slots (shared-memory accumulator)
    int32_t max_slots;    // staged nodes
    int32_t max_recs, max_srcs;
    int32_t allow_struct; // recognise aligned uniform 8x8x8 cells and give them the canonical halo order
};

// ---- structured tiles ---------------------------------------------------------------------------------
// A tile is STRUCTURED when it is exactly one aligned 8x8x8 cell of equal elements: it owns the 512
// nodes at the lower corners of its 512 core elements, in Morton order (slot = interleave of x, y, z
// bits, x lowest), has no extra entries, and its 217 gathered nodes -- the x = 8, y = 8 and z = 8 faces
// of the 9x9x9 node block -- take the slots 512 + h in this canonical order:
//   h in [0, 81)    : z = 8 face,          x = h % 9,  y = h / 9
//   h in [81, 153)  : y = 8 face, z < 8,   x = k % 9,  z = k / 9     (k = h - 81)
//   h in [153, 217) : x = 8 face, y, z < 8, y = k % 8, z = k / 8     (k = h - 153)
// The step kernel then needs no per-element slot table for such a tile (hgpu_kernels.cuh).
constexpr int STRUCT_OWNED = 512, STRUCT_HALO = 217, STRUCT_NODES = 729;
inline int struct_morton3(int x, int y, int z)
{
    int m = 0;
    for (int b = 0; b < 3; b++) m |= (((x >> b) & 1) << (3 * b)) | (((y >> b) & 1) << (3 * b + 1)) | (((z >> b) & 1) << (3 * b + 2));
    return m;
}
inline int struct_slot(int x, int y, int z)     // x, y, z in 0..8
{
    if (z == 8) return STRUCT_OWNED + 9 * y + x;
    if (y == 8) return STRUCT_OWNED + 81 + 9 * z + x;
    if (x == 8) return STRUCT_OWNED + 153 + 8 * z + y;
    return struct_morton3(x, y, z);
}

// Builds the plan; returns false and sets err on inconsistent input.  self_node (may be null):
// tiles owning a flagged node become "self" tiles.  special_node (may be null): SPECIAL flags.
bool build_tile_plan(int32_t E, int32_t N, const int32_t *lnid, const TileCaps &caps,
                     const uint8_t *self_node, const uint8_t *special_node, TilePlan &plan,
                     std::string &err);

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
