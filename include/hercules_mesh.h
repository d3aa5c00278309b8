/*
 * hercules_mesh.h -- C ABI of libhercules_mesh.so: host-side octree primitives of the per-rank mesher
 * (hercules_b200/octree_local.py).  No CUDA.  Built from hercules_b200/csrc/hmesh.cpp.
 *
 * What they replace.  The reference meshes with octor, distributed over MPI ranks:
 *   octor_refinetree   octor.c:4337   with the application's toexpand callback (psolve.c:2185 -> vsrule,
 *                                     quake_util.c:215): split an octant while its edge exceeds
 *                                     Vs(centre) / (points per wavelength * f_max)
 *   octor_balancetree  octor.c:4398   2:1 across faces and edges (18 directions), ripple propagation
 *   octor_extractmesh  octor.c:5268   leaves in Morton order, nodes in Z-order with the far domain faces
 *                                     pulled in (octor.c:5466-5475), elem_t.lnid, hanging nodes and their
 *                                     anchors (node_setproperty octor.c:3294, anchor lists octor.c:5863-5991)
 * A rank works on CHUNKS of coarse cells (edge S = the largest leaf octor's multi-rank bootstrap admits,
 * octor.c:4170-4200): the balanced refinement inside a chunk depends on the material model within one cell of
 * it, so each call is self-contained and stateless; the caller runs the calls of many chunks in parallel.
 *
 * Conventions: coordinates are integers in units of the finest admissible edge h, below 2^16; Morton codes
 * interleave x (least significant bit), y, z; a node's code is the Morton code of its DOUBLED coordinates
 * (2 g, or 2 n - 1 on a far face of the domain), so ascending node code = octor's node order.  Functions
 * return 0, -1 (bad argument), -2 (out of memory) or -3 (inconsistent input).  Output arrays are malloc'ed by
 * the library and released with hmesh_free.
 */
#ifndef HERCULES_MESH_H
#define HERCULES_MESH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int hmesh_abi_version(void);
void hmesh_free(void *p);

/* Refine + balance one chunk.  dims[3] = domain in h; S = coarse edge (power of two dividing dims); smin =
 * smallest admissible edge (1).  Material model: mat_grid[gx][gy][gz] (grid_dims) = material index of the model
 * cell of edge cl h that holds a point -- what cvm_query returns for a CVM etree leaf -- and vs_tab[material];
 * an octant of edge s is split when s * factor_h > Vs at its centre (factor_h = h * ppw * f_max).
 * sub[nsub][3] = lowest corners of the chunk's cells in ascending Morton order; reg[nreg][3] = the cells to
 * refine and balance together (NULL: sub and its 26 neighbours inside the domain).
 * Out: the leaves of the WHOLE-domain mesh inside sub, ascending: codes[n], sizes[n] (edge in h); per_cell[nsub]
 * (caller's array) = leaves per cell.  want_leaves = 0: counts only. */
int hmesh_chunk_leaves(const int32_t *dims, int32_t S, int32_t smin, const uint8_t *mat_grid, const int64_t *grid_dims,
                       int32_t cl, const double *vs_tab, double factor_h, const int32_t *reg, int64_t nreg,
                       const int32_t *sub, int64_t nsub, int32_t want_leaves, uint64_t **codes_out, int32_t **sizes_out,
                       int64_t *n_out, int64_t *per_cell);

/* Nodes located in the cells X[i0, i1) of a cell set X (xkeys[nX] = Morton keys of the cells, ascending) whose
 * leaves are known exactly: lcodes / lsizes in Morton order, lstart[nX + 1] = first leaf of every cell.
 * Out: ncodes[n] ascending; xyz[n][3]; holder[n] = index of the leaf whose half-open box holds the node (-1: none);
 * dang[n] = 0 (anchored) or the number of anchors 2 / 4 (hanging on an edge / a face); anchors[ndang][4] = node
 * codes of the anchors of the hanging nodes, in order (~0 = unused); per_cell[i1 - i0] = nodes per cell. */
int hmesh_chunk_nodes(const int32_t *dims, int32_t S, const uint64_t *xkeys, int64_t nX, const int64_t *lstart,
                      const uint64_t *lcodes, const int32_t *lsizes, int64_t i0, int64_t i1, uint64_t **ncodes_out,
                      int32_t **xyz_out, int64_t **holder_out, uint8_t **dang_out, uint64_t **anchors_out, int64_t *n_out,
                      int64_t *ndang_out, int64_t *per_cell);

/* elem_t.lnid of the leaves [e0, e1): lnid[(e1 - e0)][8] = position in ncodes of every corner (x fastest, then
 * y, then z: octor.c:6449-6470), `missing` for a corner located in a cell outside X; nstart[nX + 1] = first node
 * of every cell; exyz (may be NULL) [(e1 - e0)][3] = the leaves' lowest corners. */
int hmesh_lnid(const int32_t *dims, int32_t S, const uint64_t *xkeys, int64_t nX, const int64_t *nstart, const uint64_t *ncodes,
               const uint64_t *lcodes, const int32_t *lsizes, int64_t e0, int64_t e1, int32_t missing, int32_t *lnid, int32_t *exyz);

/* The per-node sums of solver_init's lumped terms (psolve.c:3445-3473: mass_simple += M, ... -= dt a M) in grouped
 * form: out[k][n] += sum over corner columns j = 0..7 (in turn) of the sum over elements e (ascending) with
 * lnid[e][j] == n of w[k][e].  Fixed summation order = the order of the numpy restatement (one np.bincount per
 * column), so both give the same doubles.  nw weight arrays of E doubles, nw output arrays of N doubles. */
int hmesh_corner_sums(int64_t E, const int32_t *lnid, int64_t N, int32_t nw, const double *const *w, double *const *out);

/* com_allocpctl's neighbour discovery order (octor.c:2639-2742), which fixes the order of the sharers in a node's
 * share list and with it the messenger order of the schedules: for the rank's leaves cand[ncand] (indices into
 * lcodes, ascending; key_base = their position in the rank's leaf list) 4 x 4 x 4 probe points half an edge apart,
 * starting half an edge below the lowest corner (z outermost, x innermost); first[world] (preset to INT64_MAX by
 * the caller) = per foreign rank the smallest key_base * 64 + probe number at which one of its leaves was met.
 * gidx[l] = global Morton index of leaf l of X; rank of a leaf = ((gidx + 1) world - 1) / etotal (octor.c:738-742). */
int hmesh_discovery(const int32_t *dims, int32_t S, const uint64_t *xkeys, int64_t nX, const int64_t *lstart,
                    const uint64_t *lcodes, const int32_t *lsizes, const int64_t *gidx, int64_t etotal, int32_t world,
                    int32_t rank, const int64_t *cand, const int64_t *key_base, int64_t ncand, int64_t *first);

#ifdef __cplusplus
}
#endif

#endif /* HERCULES_MESH_H */
