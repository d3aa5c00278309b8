// hgpu_kernels.cuh -- sm_100a device code of libhercules_gpu.so.
//
// Kernels (DESIGN.md sections 3-4):
//   step_kernel        per-element internal force (stiffness + Rayleigh damping) gathered per
//                      owned node inside an owner-computes tile, optionally fused with the
//                      central-difference update of the tile's REGULAR nodes; persistent CTAs with
//                      a two-stage cp.async pipeline
//   source_kernel      compute_addforce_s            (psolve.c:5912-5928)
//   adjust_dist_kernel compute_adjust(DISTRIBUTION)  (psolve.c:5943-5987), anchor-centric
//   update_list_kernel solver_compute_displacement   (psolve.c:4072-4114) on the SPECIAL nodes
//   update_all_kernel  solver_compute_displacement on every node (unfused path)
//   adjust_asgn_kernel compute_adjust(ASSIGNMENT)    (psolve.c:5992-6035)
//   pack / unpack      schedule_senddata pack and unpack loops (psolve.c:4985-5011, 5035-5073)
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace hgpu {

// ------------------------------------------------------------------------------------------
// Element operator in the factored ("effective") form of stiffness.c:180-237.
//
// The 8 corner values of one displacement component are taken to 8 "modes" by the sign matrix of
// aTransposeU (stiffness.c:260-288): mode 0 = sum (unused, forced to 0 by the reference),
// 1 = z, 2 = y, 3 = x, 4 = yz, 5 = xz, 6 = xy, 7 = xyz differences.  Rows of that matrix are
// products of the corner signs x_j, y_j, z_j (psolve.c:5451-5453), so it is a 2x2x2 Walsh-
// Hadamard transform: 3 butterfly stages = 24 add/sub per component instead of 49.
// ------------------------------------------------------------------------------------------

// forward: w[j], j = jx + 2 jy + 4 jz  ->  t[k], k in the reference's mode numbering
__device__ __forceinline__ void wht_forward(const double (&w)[8], double (&t)[8])
{
    // x stage: s = sum, d = (x=+1) - (x=-1)
    double sx0 = w[0] + w[1], dx0 = w[1] - w[0];
    double sx1 = w[2] + w[3], dx1 = w[3] - w[2];
    double sx2 = w[4] + w[5], dx2 = w[5] - w[4];
    double sx3 = w[6] + w[7], dx3 = w[7] - w[6];
    // y stage on (jy, jz) pairs
    double s_s0 = sx0 + sx1, s_d0 = sx1 - sx0;   // jz = 0 : x-sum   -> y-sum, y-diff
    double s_s1 = sx2 + sx3, s_d1 = sx3 - sx2;   // jz = 1
    double d_s0 = dx0 + dx1, d_d0 = dx1 - dx0;   // jz = 0 : x-diff  -> y-sum, y-diff
    double d_s1 = dx2 + dx3, d_d1 = dx3 - dx2;   // jz = 1
    // z stage
    t[0] = 0.0;                 // the reference zeroes the rigid-translation mode (stiffness.c:261)
    t[1] = s_s1 - s_s0;         // z
    t[2] = s_d0 + s_d1;         // y
    t[3] = d_s0 + d_s1;         // x
    t[4] = s_d1 - s_d0;         // yz
    t[5] = d_s1 - d_s0;         // xz
    t[6] = d_d0 + d_d1;         // xy
    t[7] = d_d1 - d_d0;         // xyz
}

// inverse (au, stiffness.c:388-413): f[j] = sum_k S[k][j] v[k]
__device__ __forceinline__ void wht_inverse(const double (&v)[8], double (&f)[8])
{
    // z stage: combine each (bx,by) pair of modes into jz = 0 / 1 values
    double a0 = v[0] - v[1], a1 = v[0] + v[1];   // (0,0): 1 , z
    double b0 = v[2] - v[4], b1 = v[2] + v[4];   // (0,1): y , yz
    double c0 = v[3] - v[5], c1 = v[3] + v[5];   // (1,0): x , xz
    double d0 = v[6] - v[7], d1 = v[6] + v[7];   // (1,1): xy, xyz
    // y stage
    double p00 = a0 - b0, p01 = a0 + b0;         // bx = 0, jz = 0 : jy = 0 / 1
    double p10 = a1 - b1, p11 = a1 + b1;         // bx = 0, jz = 1
    double q00 = c0 - d0, q01 = c0 + d0;         // bx = 1, jz = 0
    double q10 = c1 - d1, q11 = c1 + d1;         // bx = 1, jz = 1
    // x stage
    f[0] = p00 - q00; f[1] = p00 + q00;
    f[2] = p01 - q01; f[3] = p01 + q01;
    f[4] = p10 - q10; f[5] = p10 + q10;
    f[6] = p11 - q11; f[7] = p11 + q11;
}

// firstVector (stiffness.c:291-319) with a = -0.5625 (c2 + 2 c1), c = -0.5625 c2, b = -0.5625 c1.
// Divisions by 3 and 9 are multiplications by the rounded reciprocals (<= 1 ulp apart).
__device__ __forceinline__ void scale_modes(const double (&tx)[8], const double (&ty)[8],
                                            const double (&tz)[8], double a, double c, double b,
                                            double (&vx)[8], double (&vy)[8], double (&vz)[8])
{
    const double third = 1.0 / 3.0, ninth = 1.0 / 9.0;
    const double ab3 = (a + b) * third, c3 = c * third, b3 = b * third;
    const double a2b9 = (a + 2.0 * b) * ninth;
    vx[0] = 0.0; vy[0] = 0.0; vz[0] = 0.0;
    vx[1] = b * (tz[3] + tx[1]);
    vx[2] = b * (ty[3] + tx[2]);
    vx[3] = a * tx[3] + c * (ty[2] + tz[1]);
    vx[4] = b3 * (ty[5] + tz[6] + 2.0 * tx[4]);
    vx[5] = ab3 * tx[5] + c3 * ty[4];
    vx[6] = ab3 * tx[6] + c3 * tz[4];
    vx[7] = a2b9 * tx[7];

    vy[1] = b * (tz[2] + ty[1]);
    vy[2] = a * ty[2] + c * (tx[3] + tz[1]);
    vy[3] = vx[2];
    vy[4] = ab3 * ty[4] + c3 * tx[5];
    vy[5] = b3 * (tx[4] + tz[6] + 2.0 * ty[5]);
    vy[6] = ab3 * ty[6] + c3 * tz[5];
    vy[7] = a2b9 * ty[7];

    vz[1] = a * tz[1] + c * (tx[3] + ty[2]);
    vz[2] = vy[1];
    vz[3] = vx[1];
    vz[4] = ab3 * tz[4] + c3 * tx[6];
    vz[5] = ab3 * tz[5] + c3 * ty[6];
    vz[6] = b3 * (tx[4] + ty[5] + 2.0 * tz[6]);
    vz[7] = a2b9 * tz[7];
}

// ------------------------------------------------------------------------------------------
// step_kernel: persistent, software-pipelined tile kernel.
//
// One CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... of the processing order.  Per tile:
//
//   stage     displacements of the tile's nodes arrive in shared memory by cp.async (LDGSTS), one
//             tile ahead: 16-byte copies for the owned node range (contiguous, even start), 8-byte
//             copies for the gathered halo nodes
//   elements  one thread per entry: gather 8 corners, factored operator, accumulate into the
//             tile's shared-memory accumulator in eight corner passes (a node is corner j of at
//             most one element, so no two threads touch one accumulator within a pass)
//   publish   what the tile's core elements added to nodes of HIGHER tiles goes to partial[]
//             (one slot per (tile, node)), then the tile's flag is raised to this pass's epoch
//   finish    one tile LATER (so nobody waits in practice): wait for the flags of the lower tiles
//             that publish for this tile's nodes, add their partial forces in a fixed order, and
//             advance the owned nodes (central difference, psolve.c:4078-4108) -- or hand the
//             force of SPECIAL nodes to the force array
//
// Every element is evaluated once, every sum has a fixed order: no atomics, bit-reproducible.
//
//   MODE 0: w = u1                     stiffness term only            (damping none / mass)
//   MODE 1: w = u1 + beta (u1 - u2)    stiffness + Rayleigh damping   beta = c3/c1 = c4/c2 = b/dt
//   MODE 2: w = beta (u1 - u2)         Rayleigh damping only          (psolve.c:3387-3409)
// ------------------------------------------------------------------------------------------
struct StepArgs {
    const double *__restrict__ u1;      // tm1  [N][3]
    const double *__restrict__ u2;      // tm2  [N][3]
    double *__restrict__ unext;         // u(t+dt) target (fused update) [N][3]
    double *__restrict__ force;         // [N][3]
    const double *__restrict__ nt3;     // [N][3] {1/mass_simple, mass2_minusaM, mass_minusaM} of nodes the
                                        // fused update may advance; first entry negative = hand the force on
    const double *__restrict__ Kd;      // dense K1|K2 as [2][24][24] (conventional only)
    const int4 *__restrict__ tile_meta; // per tile, in processing order: 4 x int4, see TileMeta
    const int32_t *__restrict__ halo_id;
    const uint4 *__restrict__ ent_slot; // per entry 8 x uint16: 3 * tile-local slot of each corner
    const double *__restrict__ ent_coef;// per entry c1, c2, beta
    const uint2 *__restrict__ rec;      // FinishRec {slot3 | cnt << 16 | flags << 24, first}
    const int32_t *__restrict__ src;    // partial index per incoming contribution
    const int32_t *__restrict__ dep;    // tile ids whose flags a tile waits for
    double *partial;                    // [halo slots][3] published partial forces
    unsigned int *flag;                 // [ntiles] epoch of the tile's last publish
    unsigned int epoch;
    int32_t tile_begin, ntiles;         // this launch processes tile_meta[tile_begin .. ntiles)
    int32_t cap_slots;                  // staged nodes per stage
    int32_t cap_acc;                    // accumulator nodes (owned + published)
    int32_t cap_owned;                  // owned nodes (pending buffer)
    int32_t cap_recs, cap_srcs;         // finish records / sources staged per tile
    int32_t fuse_update;                // 1: advance owned REGULAR nodes here
};

constexpr int META_INTS = 16;           // per tile
constexpr int META_RING = 8;
constexpr int CAP_DEPS = 64;            // dependency ids staged in shared memory per tile

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

struct Entry { uint4 s; double c1, c2, beta; };

// Loads that must be ISSUED where they are written (register prefetch one round / one phase ahead
// of their use): volatile asm keeps the compiler from sinking them next to the consumer.
__device__ __forceinline__ double ldg_f64_pinned(const double *p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_u4_pinned(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_i32_pinned(const int32_t *p)
{
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// Flags are polled with relaxed loads and raised with a relaxed store after a gpu-scope fence;
// the data they guard (partial forces) is written with st.cg and read with cp.async.cg / ld.cg,
// i.e. at L2, where the fence has made it visible -- no L1 invalidation is needed on either side.
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void publish_flag(unsigned int *p, unsigned int v)
{
    asm volatile("fence.acq_rel.gpu;\n\tst.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Raw offsets as loaded (differences are taken where they are used, so that nothing consumes a
// freshly requested value early).
struct TileMeta {
    int n0, n1, hb, h1;         // owned node range; halo_id range
    int eb, e1, ncore, npub;    // entry range; core entries; published halo slots
    int rb, r1, sb, s1;         // finish records; sources
    int db, d1, id;             // dependencies; tile id (flag index)
    __device__ __forceinline__ int nown() const { return n1 - n0; }
    __device__ __forceinline__ int nh() const { return h1 - hb; }
    __device__ __forceinline__ int ne() const { return e1 - eb; }
};

// Tile offsets travel through a shared-memory ring filled by cp.async (no registers, no
// scoreboard): slot i & (META_RING-1) holds the offsets of the CTA's i-th tile.
__device__ __forceinline__ void fetch_meta_async(const StepArgs &A, int t, int *slot, int tid)
{
    if (tid < 4) cp_async16(slot + 4 * tid, A.tile_meta + 4 * (size_t)t + tid);
}
__device__ __forceinline__ TileMeta read_meta(const int *slot)
{
    const volatile int *v = slot;
    TileMeta m;
    m.n0 = v[0]; m.n1 = v[1]; m.hb = v[2]; m.h1 = v[3];
    m.eb = v[4]; m.e1 = v[5]; m.ncore = v[6]; m.npub = v[7];
    m.rb = v[8]; m.r1 = v[9]; m.sb = v[10]; m.s1 = v[11];
    m.db = v[12]; m.d1 = v[13]; m.id = v[14];
    return m;
}

template <bool NEED_BETA>
__device__ __forceinline__ Entry load_entry(const StepArgs &A, int idx)
{
    Entry e;
    e.s = ldg_u4_pinned(A.ent_slot + idx);
    const double *c = A.ent_coef + 3 * (size_t)idx;
    e.c1 = ldg_f64_pinned(c); e.c2 = ldg_f64_pinned(c + 1);
    e.beta = NEED_BETA ? ldg_f64_pinned(c + 2) : 0.0;
    return e;
}

// Stage the displacements of one tile: su1 (and su2) <- owned range + gathered halo nodes.
// U2_OWNED: copy the owned part of u2; U2_HALO: also its halo part.  The ids of the first
// HALO_PRE * blockDim.x halo nodes are loaded by the caller ahead of time (hid[]).
constexpr int HALO_PRE = 2;

__device__ __forceinline__ void load_halo_ids(const StepArgs &A, const TileMeta &m, int tid, int nthr,
                                              int (&hid)[HALO_PRE])
{
#pragma unroll
    for (int q = 0; q < HALO_PRE; q++) {
        const int h = tid + q * nthr;
        hid[q] = h < m.nh() ? ldg_i32_pinned(A.halo_id + m.hb + h) : -1;
    }
}

template <bool U2_OWNED, bool U2_HALO>
__device__ __forceinline__ void stage_tile(const StepArgs &A, const TileMeta &m, double *su1, double *su2,
                                           int tid, int nthr, const int (&hid)[HALO_PRE])
{
    const int nd = 3 * m.nown(), nv = nd >> 1;
    const double *g1 = A.u1 + 3 * (size_t)m.n0, *g2 = A.u2 + 3 * (size_t)m.n0;
    for (int i = tid; i < nv; i += nthr) {
        cp_async16(su1 + 2 * i, g1 + 2 * i);
        if (U2_OWNED) cp_async16(su2 + 2 * i, g2 + 2 * i);
    }
    if ((nd & 1) && tid == 0) {
        cp_async8(su1 + nd - 1, g1 + nd - 1);
        if (U2_OWNED) cp_async8(su2 + nd - 1, g2 + nd - 1);
    }
    // gathered nodes: one thread per node, three 8-byte copies per array
#pragma unroll
    for (int q = 0; q < HALO_PRE; q++) {
        const int h = tid + q * nthr;
        if (h < m.nh() && hid[q] >= 0) {        // -1 = slot left unused by the plan
            const size_t g = 3 * (size_t)hid[q];
            double *d1 = su1 + nd + 3 * h, *d2 = su2 + nd + 3 * h;
            cp_async8(d1, A.u1 + g); cp_async8(d1 + 1, A.u1 + g + 1); cp_async8(d1 + 2, A.u1 + g + 2);
            if (U2_HALO) { cp_async8(d2, A.u2 + g); cp_async8(d2 + 1, A.u2 + g + 1); cp_async8(d2 + 2, A.u2 + g + 2); }
        }
    }
    for (int h = tid + HALO_PRE * nthr; h < m.nh(); h += nthr) {
        const int id = __ldg(A.halo_id + m.hb + h);
        if (id < 0) continue;
        const size_t g = 3 * (size_t)id;
        double *d1 = su1 + nd + 3 * h, *d2 = su2 + nd + 3 * h;
        cp_async8(d1, A.u1 + g); cp_async8(d1 + 1, A.u1 + g + 1); cp_async8(d1 + 2, A.u1 + g + 2);
        if (U2_HALO) { cp_async8(d2, A.u2 + g); cp_async8(d2 + 1, A.u2 + g + 1); cp_async8(d2 + 2, A.u2 + g + 2); }
    }
}

// Finish data of one tile (records, sources, dependency ids) -> shared memory, by cp.async.
// Layout of one buffer: uint2 rec[cap_recs] | int src[cap_srcs] | int dep[CAP_DEPS].
__device__ __forceinline__ void stage_finish(const StepArgs &A, const TileMeta &m, char *buf, int tid, int nthr)
{
    uint2 *srec = reinterpret_cast<uint2 *>(buf);
    int *ssrc = reinterpret_cast<int *>(buf + 8 * (size_t)A.cap_recs);
    int *sdep = ssrc + A.cap_srcs;
    const int nr = min(m.r1 - m.rb, A.cap_recs), ns = min(m.s1 - m.sb, A.cap_srcs), ndp = min(m.d1 - m.db, CAP_DEPS);
    for (int i = tid; i < nr; i += nthr) cp_async8(srec + i, A.rec + m.rb + i);
    for (int i = tid; i < ns; i += nthr) cp_async4(ssrc + i, A.src + m.sb + i);
    if (tid < ndp) cp_async4(sdep + tid, A.dep + m.db + tid);
}

constexpr int NT_PRE = 3;

__device__ __forceinline__ void load_node_tables(const StepArgs &A, const TileMeta &m, int tid, int nthr,
                                                 double (&ntv)[NT_PRE][3])
{
#pragma unroll
    for (int q = 0; q < NT_PRE; q++) {
        const int i = tid + q * nthr;
        if (i < m.nown()) {
            const double *nt = A.nt3 + 3 * (size_t)(m.n0 + i);
            ntv[q][0] = ldg_f64_pinned(nt); ntv[q][1] = ldg_f64_pinned(nt + 1); ntv[q][2] = ldg_f64_pinned(nt + 2);
        } else {
            ntv[q][0] = ntv[q][1] = ntv[q][2] = 0.0;
        }
    }
}

// The tile's own share of solver_compute_displacement (psolve.c:4078-4108) for one owned node, in
// place: acc <- (acc + m2 u1 - m1 u2) / mass for a REGULAR node (rm > 0), the plain force otherwise;
// partial forces of other tiles are added, scaled alike, when the tile is finished.
__device__ __forceinline__ void settle_node(double *acc, const double *su1, const double *su2, int i,
                                            double rm, double m2, double m1)
{
    const int k = 3 * i;
    if (rm > 0.0) {
#pragma unroll
        for (int c = 0; c < 3; c++) acc[k + c] = (acc[k + c] + (m2 * su1[k + c] - m1 * su2[k + c])) * rm;
    }
}

// Shared-memory buffers of the finish phase.
struct FinishBufs {
    double *pend;       // [cap_recs][3] own share of the record nodes of the tile being finished
    double *spart;      // [cap_srcs][3] partial forces published by lower tiles
    double *srm;        // [cap_recs]    nt3[node][0] of the record nodes
};

// Poll the flags of the lower tiles that publish for tile m's nodes (dependency ids in buf).
__device__ __forceinline__ void poll_deps(const StepArgs &A, const TileMeta &m, const char *buf, int tid, int nthr)
{
    const int *sdep = reinterpret_cast<const int *>(buf + 8 * (size_t)A.cap_recs) + A.cap_srcs;
    const int ndep = m.d1 - m.db;
    for (int d = tid; d < ndep; d += nthr) {
        const int dt = d < CAP_DEPS ? sdep[d] : __ldg(A.dep + m.db + d);
        const unsigned int *f = A.flag + dt;
        while ((int)(ld_relaxed_u32(f) - A.epoch) < 0) __nanosleep(32);
    }
}

// Request the partial forces tile m reads, and the node-table entry of its record nodes
// (cp.async.cg: served by L2, where the publishers' fences made them visible).
__device__ __forceinline__ void request_partials(const StepArgs &A, const TileMeta &m, const char *buf,
                                                 const FinishBufs &fb, int tid, int nthr)
{
    const uint2 *srec = reinterpret_cast<const uint2 *>(buf);
    const int *ssrc = reinterpret_cast<const int *>(buf + 8 * (size_t)A.cap_recs);
    const int ns = min(m.s1 - m.sb, A.cap_srcs), nr = min(m.r1 - m.rb, A.cap_recs);
    for (int q = tid; q < ns; q += nthr) {
        const double *pp = A.partial + 3 * (size_t)ssrc[q];
        cp_async8(fb.spart + 3 * q, pp); cp_async8(fb.spart + 3 * q + 1, pp + 1); cp_async8(fb.spart + 3 * q + 2, pp + 2);
    }
    const double *nt = A.nt3 + 3 * (size_t)m.n0;
    for (int r = tid; r < nr; r += nthr) cp_async8(fb.srm + r, nt + (srec[r].x & 0xffff));
}

// Finish a tile: add the partial forces of the lower tiles to the record nodes' own share (pend)
// in a fixed order and store the result -- u(t+dt) of a REGULAR node, or the force of a SPECIAL
// node (source term, hanging-node transfer, halo exchange and the list update follow).
__device__ __forceinline__ void finish_tile(const StepArgs &A, const TileMeta &m, const char *buf,
                                            const FinishBufs &fb, int tid, int nthr, bool fuse)
{
    const uint2 *srec = reinterpret_cast<const uint2 *>(buf);
    const int *ssrc = reinterpret_cast<const int *>(buf + 8 * (size_t)A.cap_recs);
    const int nrec = m.r1 - m.rb;
    const size_t g0 = 3 * (size_t)m.n0;
    for (int r = tid; r < nrec; r += nthr) {
        const bool in_smem = r < A.cap_recs;
        const uint2 rc = in_smem ? srec[r] : __ldg(A.rec + m.rb + r);
        const int slot3 = rc.x & 0xffff, cnt = (rc.x >> 16) & 0xff, first = (int)rc.y;
        const double rmv = in_smem ? fb.srm[r] : __ldg(A.nt3 + g0 + slot3);
        const bool regular = fuse && rmv > 0.0;
        const double scale = regular ? rmv : 1.0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        if (in_smem && fuse) { s0 = fb.pend[3 * r]; s1 = fb.pend[3 * r + 1]; s2 = fb.pend[3 * r + 2]; }
        for (int k = 0; k < cnt; k++) {
            const int q = first + k;
            double v0, v1, v2;
            if (q < A.cap_srcs) { v0 = fb.spart[3 * q]; v1 = fb.spart[3 * q + 1]; v2 = fb.spart[3 * q + 2]; }
            else {
                const double *pp = A.partial + 3 * (size_t)__ldg(A.src + m.sb + q);
                v0 = __ldcg(pp); v1 = __ldcg(pp + 1); v2 = __ldcg(pp + 2);
            }
            s0 = fma(v0, scale, s0); s1 = fma(v1, scale, s1); s2 = fma(v2, scale, s2);
        }
        if (regular) {
            double *o = A.unext + g0 + slot3;
            // a record beyond the staged ones kept its own share in unext
            if (!in_smem) { s0 += o[0]; s1 += o[1]; s2 += o[2]; }
            o[0] = s0; o[1] = s1; o[2] = s2;
        } else {
            // fused launch: the own share waited in pend; unfused: it is in the force array already
            double *fo = A.force + g0 + slot3;
            if (fuse && !in_smem) { const double *o = A.unext + g0 + slot3; s0 += o[0]; s1 += o[1]; s2 += o[2]; }
            fo[0] += s0; fo[1] += s1; fo[2] += s2;
        }
    }
}

template <int MODE, bool DENSE, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) step_kernel(const StepArgs A)
{
    constexpr bool U2E = MODE != 0;          // elements read u2
    extern __shared__ double smem[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int S3 = 3 * A.cap_slots, O3 = 3 * A.cap_owned, A3 = 3 * A.cap_acc;
    const int stage_doubles = S3 + (U2E ? S3 : O3);
    double *acc = smem + 2 * stage_doubles;
    FinishBufs fb;
    fb.pend = acc + A3;
    fb.spart = fb.pend + 3 * A.cap_recs;
    fb.srm = fb.spart + 3 * A.cap_srcs;
    char *fbuf = reinterpret_cast<char *>(fb.srm + A.cap_recs);
    const int fbuf_bytes = 8 * A.cap_recs + 4 * A.cap_srcs + 4 * CAP_DEPS;
    const bool fuse = A.fuse_update != 0;

    // Pipeline.  All global loads of a warp share one hardware scoreboard, so a consumer waits
    // for EVERY load issued before it, however young.  Each register prefetch is therefore
    // consumed right BEFORE the next batch of loads is issued, and everything that can avoid
    // registers does:
    //   displacements : cp.async into the other stage, one tile ahead
    //   tile offsets  : cp.async into a ring, three tiles ahead
    //   finish data   : cp.async when the tile's element phase starts (records, sources, deps)
    //   partial forces: cp.async during the accumulation passes of the NEXT tile's last round,
    //                   after the publishers' flags have been seen
    //   entries       : registers; enext -> ecur at the top of a round, then the following round's
    //                   entry is requested
    //   halo ids      : registers; requested before the accumulation passes of a tile's last round
    //                   for the tile that is staged at the top of the next iteration
    //   node tables   : registers; requested before the accumulation passes of the last round
    __shared__ __align__(16) int smeta[META_RING][META_INTS];
    const int G = gridDim.x;
    int t = A.tile_begin + blockIdx.x;
    if (t >= A.ntiles) return;
    for (int k = tid; k < A3; k += nthr) acc[k] = 0.0;
    fetch_meta_async(A, t, smeta[0], tid);
    if (t + G < A.ntiles) fetch_meta_async(A, t + G, smeta[1], tid);
    if (t + 2 * G < A.ntiles) fetch_meta_async(A, t + 2 * G, smeta[2], tid);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    int hid[HALO_PRE];
    {
        const TileMeta cur = read_meta(smeta[0]);
        load_halo_ids(A, cur, tid, nthr, hid);
        if (U2E || fuse) stage_tile<true, U2E>(A, cur, smem, smem + S3, tid, nthr, hid);
        else             stage_tile<false, false>(A, cur, smem, smem + S3, tid, nthr, hid);
        cp_async_commit();
        if (t + G < A.ntiles) load_halo_ids(A, read_meta(smeta[1]), tid, nthr, hid);
    }
    Entry ecur, enext;
    enext.s = make_uint4(0, 0, 0, 0); enext.c1 = enext.c2 = enext.beta = 0.0;
    {
        const TileMeta cur = read_meta(smeta[0]);
        if (tid < cur.ne()) enext = load_entry<U2E>(A, cur.eb + tid);
    }

    for (int it = 0;; it++) {
        double *su1 = smem + (it & 1) * stage_doubles;
        double *su2 = su1 + S3;
        const int tn = t + G;
        const bool has_next = tn < A.ntiles;
        const bool has_nn = tn + G < A.ntiles;
        const int *m_cur = smeta[it & (META_RING - 1)], *m_nxt = smeta[(it + 1) & (META_RING - 1)];
        const int *m_nn = smeta[(it + 2) & (META_RING - 1)], *m_prv = smeta[(it - 1) & (META_RING - 1)];
        const char *fb_prv = fbuf + ((it - 1) & 1) * fbuf_bytes;
        cp_async_wait_all();
        __syncthreads();                      // tile `it` has landed; everyone is done with tile it-1's stage
        stage_finish(A, read_meta(m_cur), fbuf + (it & 1) * fbuf_bytes, tid, nthr);
        if (has_next) {
            double *n1 = smem + ((it + 1) & 1) * stage_doubles;
            const TileMeta nxt = read_meta(m_nxt);
            if (U2E || fuse) stage_tile<true, U2E>(A, nxt, n1, n1 + S3, tid, nthr, hid);
            else             stage_tile<false, false>(A, nxt, n1, n1 + S3, tid, nthr, hid);
            if (tn + 2 * G < A.ntiles) fetch_meta_async(A, tn + 2 * G, smeta[(it + 3) & (META_RING - 1)], tid);
        }
        cp_async_commit();
        // only what the element phase needs stays in registers; the rest of the tile's offsets is
        // read again from the ring where it is used
        struct { int n0, n1, eb, e1, ncore;
                 __device__ __forceinline__ int nown() const { return n1 - n0; }
                 __device__ __forceinline__ int ne() const { return e1 - eb; } } cur;
        {
            const volatile int *v = m_cur;
            cur.n0 = v[0]; cur.n1 = v[1]; cur.eb = v[4]; cur.e1 = v[5]; cur.ncore = v[6];
        }
        const int nown3 = 3 * cur.nown();
        const int nxt_eb = has_next ? ((const volatile int *)m_nxt)[4] : 0;
        const int nxt_ne = has_next ? ((const volatile int *)m_nxt)[5] - nxt_eb : 0;
        // the previous tile is finished during this one: its publishers' flags are polled and its
        // partial forces requested between the accumulation passes of the last round
        bool prv_pending = it > 0;
        double ntv[NT_PRE][3];

        // ---- element forces, accumulated per staged node ----------------------------------------
        for (int base = 0; base < cur.ne(); base += nthr) {
            const bool act = base + tid < cur.ne();
            const bool last = base + nthr >= cur.ne();
            // a core entry adds to every corner (owned or published); an extra entry of a self
            // tile only to the owned ones
            const uint32_t lim = base + tid < cur.ncore ? 0xffffffffu : (uint32_t)nown3;
            ecur = enext;
            // next round's entry (or the first round of the next tile) rides along with the math
            if (!last) {
                if (base + nthr + tid < cur.ne()) enext = load_entry<U2E>(A, cur.eb + base + nthr + tid);
            } else if (tid < nxt_ne) {
                enext = load_entry<U2E>(A, nxt_eb + tid);
            }
            double fx[8], fy[8], fz[8];
            uint32_t sl[8];
            if (act) {
                sl[0] = ecur.s.x & 0xffffu; sl[1] = ecur.s.x >> 16; sl[2] = ecur.s.y & 0xffffu; sl[3] = ecur.s.y >> 16;
                sl[4] = ecur.s.z & 0xffffu; sl[5] = ecur.s.z >> 16; sl[6] = ecur.s.w & 0xffffu; sl[7] = ecur.s.w >> 16;
                double wx[8], wy[8], wz[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int o = sl[j];
                    const double ax = su1[o], ay = su1[o + 1], az = su1[o + 2];
                    if (MODE == 0) { wx[j] = ax; wy[j] = ay; wz[j] = az; }
                    else {
                        const double dx = ax - su2[o], dy = ay - su2[o + 1], dz = az - su2[o + 2];
                        if (MODE == 1) { wx[j] = fma(ecur.beta, dx, ax); wy[j] = fma(ecur.beta, dy, ay); wz[j] = fma(ecur.beta, dz, az); }
                        else           { wx[j] = ecur.beta * dx; wy[j] = ecur.beta * dy; wz[j] = ecur.beta * dz; }
                    }
                }
                if (!DENSE) {
                    double tx[8], ty[8], tz[8];
                    wht_forward(wx, tx); wht_forward(wy, ty); wht_forward(wz, tz);
                    const double a = -0.5625 * (ecur.c2 + 2.0 * ecur.c1);
                    const double c = -0.5625 * ecur.c2;
                    const double b = -0.5625 * ecur.c1;
                    scale_modes(tx, ty, tz, a, c, b, wx, wy, wz);   // reuse w* as the scaled modes
                    wht_inverse(wx, fx); wht_inverse(wy, fy); wht_inverse(wz, fz);
                } else {
                    // conventional form (stiffness.c:143-162): f_i = -c1 K1[i][j] w_j - c2 K2[i][j] w_j,
                    // K1|K2 stored as two 24x24 row-major matrices (row = 3 i + k, col = 3 j + l)
                    const double *K1 = A.Kd, *K2 = A.Kd + 576;
#pragma unroll 1
                    for (int i = 0; i < 8; i++) {
                        double r[3];
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const double *r1 = K1 + 24 * (3 * i + k), *r2 = K2 + 24 * (3 * i + k);
                            double s1 = 0.0, s2 = 0.0;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                s1 = fma(__ldg(r1 + 3 * j), wx[j], s1); s1 = fma(__ldg(r1 + 3 * j + 1), wy[j], s1);
                                s1 = fma(__ldg(r1 + 3 * j + 2), wz[j], s1);
                                s2 = fma(__ldg(r2 + 3 * j), wx[j], s2); s2 = fma(__ldg(r2 + 3 * j + 1), wy[j], s2);
                                s2 = fma(__ldg(r2 + 3 * j + 2), wz[j], s2);
                            }
                            r[k] = -ecur.c1 * s1 - ecur.c2 * s2;
                        }
#pragma unroll
                        for (int jj = 0; jj < 8; jj++)
                            if (jj == i) { fx[jj] = r[0]; fy[jj] = r[1]; fz[jj] = r[2]; }
                    }
                }
            }
            if (last) {
                if (fuse) load_node_tables(A, read_meta(m_cur), tid, nthr, ntv);
                if (has_nn) load_halo_ids(A, read_meta(m_nn), tid, nthr, hid);
            }
            // A node is corner j of at most one element (leaf octants do not overlap), so within
            // pass j every accumulator is touched by at most one thread: no atomics, fixed order.
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (act && sl[j] < lim) {
                    const int o = sl[j];
                    acc[o] += fx[j]; acc[o + 1] += fy[j]; acc[o + 2] += fz[j];
                }
                if (j == 2 && last && prv_pending) poll_deps(A, read_meta(m_prv), fb_prv, tid, nthr);
                __syncthreads();
                if (j == 2 && last && prv_pending) {
                    request_partials(A, read_meta(m_prv), fb_prv, fb, tid, nthr);
                    cp_async_commit();
                }
            }
        }
        if (cur.ne() == 0) {                    // a tile of element-less nodes: keep the pipeline fed
            if (tid < nxt_ne) enext = load_entry<U2E>(A, nxt_eb + tid);
            if (has_nn) load_halo_ids(A, read_meta(m_nn), tid, nthr, hid);
            if (fuse) load_node_tables(A, read_meta(m_cur), tid, nthr, ntv);
            if (prv_pending) {
                poll_deps(A, read_meta(m_prv), fb_prv, tid, nthr);
                __syncthreads();
                request_partials(A, read_meta(m_prv), fb_prv, fb, tid, nthr);
                cp_async_commit();
            }
        }

        // ---- publish what the core elements added to nodes of higher tiles ----------------------
        {
            const volatile int *v = m_cur;
            const int np3 = 3 * v[7];
            double *dst = A.partial + 3 * (size_t)v[2];
            for (int k = tid; k < np3; k += nthr) {
                __stcg(dst + k, acc[nown3 + k]);
                acc[nown3 + k] = 0.0;
            }
        }
        // ---- this tile's own share of the update, in place ---------------------------------------
        if (fuse) {
#pragma unroll
            for (int q = 0; q < NT_PRE; q++) {
                const int i = tid + q * nthr;
                if (i < cur.nown()) settle_node(acc, su1, su2, i, ntv[q][0], ntv[q][1], ntv[q][2]);
            }
            for (int i = tid + NT_PRE * nthr; i < cur.nown(); i += nthr) {
                const double *nt = A.nt3 + 3 * (size_t)(cur.n0 + i);
                settle_node(acc, su1, su2, i, __ldg(nt), __ldg(nt + 1), __ldg(nt + 2));
            }
        }
        cp_async_wait_all();                    // partial forces of the previous tile, this tile's finish data
        __syncthreads();                        // ... and every thread's published partial forces are written
        if (tid == 0) publish_flag(A.flag + ((const volatile int *)m_cur)[14], A.epoch);

        // ---- finish the previous tile -------------------------------------------------------------
        if (prv_pending) {
            finish_tile(A, read_meta(m_prv), fb_prv, fb, tid, nthr, fuse);
            __syncthreads();                    // pend is free again
        }
        // ---- hand this tile on: record nodes keep their own share in pend, the rest is final ----
        {
            const uint2 *srec = reinterpret_cast<const uint2 *>(fbuf + (it & 1) * fbuf_bytes);
            const volatile int *v = m_cur;
            const int nrec = min(v[9] - v[8], A.cap_recs);
            if (fuse)
                for (int r = tid; r < nrec; r += nthr) {
                    const int slot3 = srec[r].x & 0xffff;
                    fb.pend[3 * r] = acc[slot3]; fb.pend[3 * r + 1] = acc[slot3 + 1]; fb.pend[3 * r + 2] = acc[slot3 + 2];
                }
            const size_t g0 = 3 * (size_t)cur.n0;
            if (fuse) {
                // coalesced 128-bit copy-out (g0 is even); rows of record nodes are rewritten when
                // the tile is finished, rows of SPECIAL nodes by the special-node update
                double2 *dst = reinterpret_cast<double2 *>(A.unext + g0);
                for (int k = tid; k < (nown3 >> 1); k += nthr) dst[k] = make_double2(acc[2 * k], acc[2 * k + 1]);
                if ((nown3 & 1) && tid == 0) A.unext[g0 + nown3 - 1] = acc[nown3 - 1];
            } else {
                for (int k = tid; k < nown3; k += nthr) A.force[g0 + k] += acc[k];
            }
            __syncthreads();
            for (int k = tid; k < nown3; k += nthr) acc[k] = 0.0;
        }
        if (!has_next) {
            // the last tile of this CTA is finished right away
            const TileMeta me = read_meta(m_cur);
            const char *fb_cur = fbuf + (it & 1) * fbuf_bytes;
            poll_deps(A, me, fb_cur, tid, nthr);
            __syncthreads();
            request_partials(A, me, fb_cur, fb, tid, nthr);
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
            finish_tile(A, me, fb_cur, fb, tid, nthr, fuse);
            break;
        }
        t = tn;
    }
}

// compute_addforce_s (psolve.c:5912-5928): assignment, runs before the element forces
__global__ void source_kernel(int n, const int32_t *__restrict__ lnid, const double *__restrict__ F,
                              double dt2, double *__restrict__ force)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 3 * n) force[3 * (size_t)lnid[k / 3] + (k % 3)] = F[k] * dt2;
}

// compute_adjust(DISTRIBUTION) (psolve.c:5943-5987), one thread per (anchor, component); the
// contributions arrive in dnode-table order exactly as in the reference's sequential loop.
__global__ void adjust_dist_kernel(int nA, const int32_t *__restrict__ anchor_id,
                                   const int32_t *__restrict__ anchor_off,
                                   const int32_t *__restrict__ anchor_dn,
                                   const int32_t *__restrict__ anchor_deps,
                                   double *__restrict__ v)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 3 * nA) return;
    const int a = k / 3, c = k - 3 * a;
    double s = v[3 * (size_t)anchor_id[a] + c];
    for (int i = anchor_off[a]; i < anchor_off[a + 1]; i++)
        s += v[3 * (size_t)anchor_dn[i] + c] / (double)(uint32_t)anchor_deps[i];
    v[3 * (size_t)anchor_id[a] + c] = s;
}

// compute_adjust(ASSIGNMENT) (psolve.c:5992-6035): dangling = sum over anchors of value/deps
__global__ void adjust_asgn_kernel(int D, const int32_t *__restrict__ dnode, double *__restrict__ v)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 3 * D) return;
    const int d = k / 3, c = k - 3 * d;
    const int32_t *dn = dnode + 6 * (size_t)d;
    const double deps = (double)(uint32_t)dn[1];
    double s = 0.0;
    for (int a = 0; a < 4; a++) {
        const int p = dn[2 + a];
        if (p < 0) break;
        s += v[3 * (size_t)p + c] / deps;
    }
    v[3 * (size_t)dn[0] + c] = s;
}

// solver_compute_displacement (psolve.c:4072-4114) on a node list; force is zeroed afterwards
__global__ void update_list_kernel(int n, const int32_t *__restrict__ list,
                                   const double *__restrict__ u1, const double *__restrict__ u2,
                                   double *__restrict__ unext, double *__restrict__ force,
                                   const double *__restrict__ mass, const double *__restrict__ m2,
                                   const double *__restrict__ m1)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 3 * n) return;
    const int node = list[k / 3];
    const size_t g = 3 * (size_t)node + (k % 3);
    const double nf = force[g] + (m2[g] * u1[g] - m1[g] * u2[g]);
    unext[g] = nf / mass[node];
    force[g] = 0.0;
}

// the same over every harbored node (unfused path): pure streaming, 176 B per node
__global__ void update_all_kernel(long long n3, const double *__restrict__ u1,
                                  const double *__restrict__ u2, double *__restrict__ unext,
                                  double *__restrict__ force, const double *__restrict__ mass,
                                  const double *__restrict__ m2, const double *__restrict__ m1)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n3; g += stride) {
        const double nf = force[g] + (__ldg(m2 + g) * __ldg(u1 + g) - __ldg(m1 + g) * __ldg(u2 + g));
        unext[g] = nf / __ldg(mass + g / 3);
        force[g] = 0.0;
    }
}

// schedule_senddata pack loop (psolve.c:4985-5011): buf[i] = v[mapping[i]]
__global__ void pack_kernel(int n, const int32_t *__restrict__ mapping, const double *__restrict__ v,
                            double *__restrict__ buf)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 3 * n) buf[k] = v[3 * (size_t)mapping[k / 3] + (k % 3)];
}

// schedule_senddata unpack loop (psolve.c:5035-5073).  add = 1: CONTRIBUTION (+=), 0: SHARING (=).
// With CONTRIBUTION a node shared with several neighbours appears in several messengers; the
// messengers are applied one after another (segments processed by successive launches), as in
// the reference, so the sum order is fixed.
__global__ void unpack_kernel(int n, const int32_t *__restrict__ mapping, const double *__restrict__ buf,
                              double *__restrict__ v, int add)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 3 * n) return;
    const size_t g = 3 * (size_t)mapping[k / 3] + (k % 3);
    v[g] = add ? v[g] + buf[k] : buf[k];
}

// ------------------------------------------------------------------------------------------
// Halo exchange over peer memory (CUDA IPC mailboxes, NVLink between GPUs): the sender packs
// straight into the receiver's mailbox and raises a sequence flag; the receiver's kernel waits
// for the flag and applies the data.  One kernel on each side per exchange, no host round trip.
// ------------------------------------------------------------------------------------------
struct PushSeg {
    const int32_t *mapping;     // local node ids, messenger order (psolve.c:4806-4860)
    double *remote;             // peer mailbox segment for this exchange parity
    unsigned long long *remote_flag;
    unsigned int *counter;      // local: blocks done
    int32_t n;
};

// schedule_senddata pack + send (psolve.c:4985-5025): blockIdx.y = messenger
__global__ void p2p_push_kernel(const PushSeg *__restrict__ segs, const double *__restrict__ v,
                                unsigned long long seq)
{
    const PushSeg sg = segs[blockIdx.y];
    const int n3 = 3 * sg.n;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n3; k += gridDim.x * blockDim.x)
        sg.remote[k] = v[3 * (size_t)sg.mapping[k / 3] + (k % 3)];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(sg.counter, 1u);
        if (prev == gridDim.x - 1) {
            *sg.counter = 0;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(sg.remote_flag) = seq;
        }
    }
}

struct PullSeg {
    const int32_t *mapping;
    const double *local;        // my mailbox segment for this exchange parity
    const unsigned long long *flag;
    int32_t n;
};

// recv + unpack (psolve.c:5032-5073).  add = 1: CONTRIBUTION (+=), 0: SHARING (=).  blockIdx.y =
// messenger (sharing only: the overwrite lists of different owners are disjoint; contributions are
// applied one messenger per launch, in list order, so sums keep the reference's order).
__global__ void p2p_pull_kernel(const PullSeg *__restrict__ segs, double *__restrict__ v,
                                unsigned long long seq, int add, int *__restrict__ err)
{
    const PullSeg sg = segs[blockIdx.y];
    if (threadIdx.x == 0) {
        const volatile unsigned long long *f = sg.flag;
        const long long t0 = clock64();
        while (*f < seq) {
            __nanosleep(200);
            if (clock64() - t0 > 20000000000LL) { atomicExch(err, 1); break; }   // ~10 s: peer lost
        }
    }
    __syncthreads();
    const int n3 = 3 * sg.n;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n3; k += gridDim.x * blockDim.x) {
        const size_t g = 3 * (size_t)sg.mapping[k / 3] + (k % 3);
        const double x = __ldcg(sg.local + k);
        v[g] = add ? v[g] + x : x;
    }
}

__global__ void gather_nodes_kernel(int n, const int32_t *__restrict__ lnid, const double *__restrict__ v,
                                    double *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 3 * n) out[k] = v[3 * (size_t)lnid[k / 3] + (k % 3)];
}

}  // namespace hgpu
