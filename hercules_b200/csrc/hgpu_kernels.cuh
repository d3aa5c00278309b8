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

#define HGPU_HD __host__ __device__ __forceinline__

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
HGPU_HD void wht_forward(const double (&w)[8], double (&t)[8])
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
HGPU_HD void wht_inverse(const double (&v)[8], double (&f)[8])
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
HGPU_HD void scale_modes(const double (&tx)[8], const double (&ty)[8],
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
//   MODE 3: BKT: memory variables advanced (calc_conv) and the constant-Q force, which includes
//           the elastic term (constant_Q_addforce), in one pass over the element's 768 B of state
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
    int32_t tile_begin, ntiles;         // this launch processes tile_meta[tile_begin .. ntiles) (slot-table tiles)
    int32_t struct_begin, struct_end;   // ... and tile_meta[struct_begin .. struct_end) (structured tiles, STRUCT launches)
    int32_t grid_struct;                // STRUCT launches: CTAs [0, grid_struct) walk the structured tiles
    int *queue;                         // STRUCT launches: {next slot-table tile, next structured tile} handed out dynamically, or null
    int32_t cap_slots;                  // staged nodes per stage
    int32_t cap_acc;                    // accumulator nodes (owned + published)
    int32_t cap_owned;                  // owned nodes (pending buffer)
    int32_t cap_recs, cap_srcs;         // finish records / sources staged per tile
    int32_t fuse_update;                // 1: advance owned REGULAR nodes here
    // BKT (MODE 3): per-entry memory variables and coefficients
    double *conv;                       // conv_shear_1|2, conv_kappa_1|2 in entry-chunked layout, see conv_index
    const double *__restrict__ ent_bkt; // per entry 8 doubles: c1, c2, then 10 floats a0s a1s bs g0s g1s a0k a1k bk g0k g1k, pad
    double rmax;                        // 2 pi f_max dt (damping.c:114)
    // WPASS variant (MODE 1): per tile, in processing order, the Rayleigh ratio beta = c3/c1 shared by
    // all its entries, or NaN when they differ (the tile then takes the per-corner path)
    const double *__restrict__ tile_beta;
    // STRUCT variant: per tile, in processing order, {c1, c2, beta, -} of a structured tile of one material
    const double *__restrict__ tile_coef;
    int *err;                           // error word (mapped host memory), see report_error
};

// BKT memory variables (psolve.h:308-311: conv_shear_1, conv_shear_2, conv_kappa_1, conv_kappa_2,
// each [8 E][3] in the reference) are stored per tile ENTRY in chunks of 32 entries so that the 32
// lanes of a warp read 256 contiguous bytes per value:
//   index(entry, k) = ((entry / 32) * 96 + k) * 32 + entry % 32,  k = family * 48 + which * 24 + 3 * node + comp
// (family 0 = shear, 1 = kappa; which 0 = conv_*_1, 1 = conv_*_2).
constexpr int CONV_PER_ENTRY = 96;
__host__ __device__ __forceinline__ size_t conv_index(size_t entry, int k)
{
    return ((entry >> 5) * CONV_PER_ENTRY + (size_t)k) * 32 + (entry & 31);
}

constexpr int META_INTS = 16;           // per tile
constexpr int META_RING = 8;
constexpr int CAP_DEPS = 64;            // dependency ids staged in shared memory per tile
constexpr long long WAIT_DEPS_CYCLES = 40000000000LL;   // ~20 s at 1.9 GHz

// The error word lives in page-locked host memory mapped into the device (the host reads it after
// every synchronisation without a copy): 1 = a halo peer never arrived, 2 = a tile's publishers never did.
__device__ __forceinline__ void report_error(int *err, int code)
{
    *reinterpret_cast<volatile int *>(err) = code;
    __threadfence_system();
}

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

// ---- bulk asynchronous copies (TMA, 1-D) with mbarrier completion: the contiguous owned range of a
//      structured tile (2 x 12 288 bytes) is ONE instruction per array instead of 48 cp.async per warp ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy accesses to shared memory (ld/st.shared) before, async-proxy accesses (bulk copies) after
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
// Flags are polled with relaxed loads and raised with a relaxed store after a gpu-scope fence.
// The data they guard
// (partial forces) is written with st.cg and read with 8-byte cp.async.ca: a line of partial[] has
// one writer per pass and is read once per pass, and L1 does not survive a kernel boundary, so a
// reader never finds a stale copy of it in its own L1.
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

// Tile offsets travel through a shared-memory ring filled by cp.async (no registers, no
// scoreboard): slot i & (META_RING-1) holds the offsets of the CTA's i-th tile, four int4 groups:
//   0: n0, n1, hb, h1        owned node range; halo_id range
//   1: eb, e1, ncore, npub   entry range; core entries; published halo slots
//   2: rb, r1, sb, s1        finish records; sources
//   3: db, d1, id, -         dependencies; tile id (flag index)
// A group is read (one 128-bit broadcast load) where it is used instead of being kept in registers.
__device__ __forceinline__ void fetch_meta_async(const StepArgs &A, int t, int *slot, int tid)
{
    if (tid < 4) cp_async16(slot + 4 * tid, A.tile_meta + 4 * (size_t)t + tid);
}
__device__ __forceinline__ int4 meta_group(const int *slot, int g)
{
    int4 v;
    const unsigned a = (unsigned)__cvta_generic_to_shared(slot + 4 * g);
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// MODE 0: slots + c1, c2; MODE 1, 2: + beta; MODE 3 (BKT): slots only, the coefficient record is
// read where the round starts
template <int MODE>
__device__ __forceinline__ Entry load_entry(const StepArgs &A, int idx)
{
    Entry e;
    e.s = ldg_u4_pinned(A.ent_slot + idx);
    e.c1 = e.c2 = e.beta = 0.0;
    if (MODE != 3) {
        const double *c = A.ent_coef + 3 * (size_t)idx;
        e.c1 = ldg_f64_pinned(c); e.c2 = ldg_f64_pinned(c + 1);
        if (MODE != 0) e.beta = ldg_f64_pinned(c + 2);
    }
    return e;
}

// Stage the displacements of one tile: su1 (and su2) <- owned range + gathered halo nodes.
// U2_OWNED: copy the owned part of u2; U2_HALO: also its halo part.  The ids of the first
// HALO_PRE * blockDim.x halo nodes are loaded by the caller ahead of time (hid[]).
// m = meta group 0 {n0, n1, hb, h1}.
constexpr int HALO_PRE = 2;

__device__ __forceinline__ void load_halo_ids(const StepArgs &A, const int4 m, int tid, int nthr,
                                              int (&hid)[HALO_PRE])
{
    const int nh = m.w - m.z;
#pragma unroll
    for (int q = 0; q < HALO_PRE; q++) {
        const int h = tid + q * nthr;
        hid[q] = h < nh ? ldg_i32_pinned(A.halo_id + m.z + h) : -1;
    }
}

template <bool U2_OWNED, bool U2_HALO>
__device__ __forceinline__ void stage_tile(const StepArgs &A, const int4 m, double *su1, double *su2,
                                           int tid, int nthr, const int (&hid)[HALO_PRE])
{
    const int nd = 3 * (m.y - m.x), nv = nd >> 1, nh = m.w - m.z;
    const double *g1 = A.u1 + 3 * (size_t)m.x, *g2 = A.u2 + 3 * (size_t)m.x;
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
        if (h < nh && hid[q] >= 0) {            // -1 = slot left unused by the plan
            const size_t g = 3 * (size_t)hid[q];
            double *d1 = su1 + nd + 3 * h, *d2 = su2 + nd + 3 * h;
            cp_async8(d1, A.u1 + g); cp_async8(d1 + 1, A.u1 + g + 1); cp_async8(d1 + 2, A.u1 + g + 2);
            if (U2_HALO) { cp_async8(d2, A.u2 + g); cp_async8(d2 + 1, A.u2 + g + 1); cp_async8(d2 + 2, A.u2 + g + 2); }
        }
    }
    for (int h = tid + HALO_PRE * nthr; h < nh; h += nthr) {
        const int id = __ldg(A.halo_id + m.z + h);
        if (id < 0) continue;
        const size_t g = 3 * (size_t)id;
        double *d1 = su1 + nd + 3 * h, *d2 = su2 + nd + 3 * h;
        cp_async8(d1, A.u1 + g); cp_async8(d1 + 1, A.u1 + g + 1); cp_async8(d1 + 2, A.u1 + g + 2);
        if (U2_HALO) { cp_async8(d2, A.u2 + g); cp_async8(d2 + 1, A.u2 + g + 1); cp_async8(d2 + 2, A.u2 + g + 2); }
    }
}

// The same for a STRUCTURED tile: its 512 owned nodes are one bulk copy per array (thread 0, completion
// counted on `bar`), its 217 far-face nodes one thread each as above.
template <bool U2_HALO>
__device__ __forceinline__ void stage_tile_struct(const StepArgs &A, const int4 m, double *su1, double *su2,
                                                  unsigned long long *bar, int tid, const int (&hid)[HALO_PRE])
{
    constexpr unsigned OWNED_BYTES = 512 * 3 * sizeof(double);
    if (tid == 0) {
        mbar_expect_tx(bar, 2 * OWNED_BYTES);
        bulk_g2s(su1, A.u1 + 3 * (size_t)m.x, OWNED_BYTES, bar);
        bulk_g2s(su2, A.u2 + 3 * (size_t)m.x, OWNED_BYTES, bar);
    }
    if (tid < 217 && hid[0] >= 0) {
        const size_t g = 3 * (size_t)hid[0];
        double *d1 = su1 + 1536 + 3 * tid, *d2 = su2 + 1536 + 3 * tid;
        cp_async8(d1, A.u1 + g); cp_async8(d1 + 1, A.u1 + g + 1); cp_async8(d1 + 2, A.u1 + g + 2);
        if (U2_HALO) { cp_async8(d2, A.u2 + g); cp_async8(d2 + 1, A.u2 + g + 1); cp_async8(d2 + 2, A.u2 + g + 2); }
    }
}

// Finish data of one tile (records, sources, dependency ids) -> shared memory, by cp.async.
// Layout of one buffer: uint2 rec[cap_recs] | int src[cap_srcs] | int dep[CAP_DEPS].
// mc = meta group 2 {rb, r1, sb, s1}, md = group 3 {db, d1, id, -}.
__device__ __forceinline__ void stage_finish(const StepArgs &A, const int4 mc, const int4 md, char *buf, int tid, int nthr)
{
    uint2 *srec = reinterpret_cast<uint2 *>(buf);
    int *ssrc = reinterpret_cast<int *>(buf + 8 * (size_t)A.cap_recs);
    int *sdep = ssrc + A.cap_srcs;
    const int nr = min(mc.y - mc.x, A.cap_recs), ns = min(mc.w - mc.z, A.cap_srcs), ndp = min(md.y - md.x, CAP_DEPS);
    for (int i = tid; i < nr; i += nthr) cp_async8(srec + i, A.rec + mc.x + i);
    for (int i = tid; i < ns; i += nthr) cp_async4(ssrc + i, A.src + mc.z + i);
    if (tid < ndp) cp_async4(sdep + tid, A.dep + md.x + tid);
}

constexpr int NT_PRE = 3;

__device__ __forceinline__ void load_node_tables(const StepArgs &A, int n0, int nown, int tid, int nthr,
                                                 double (&ntv)[NT_PRE][3])
{
#pragma unroll
    for (int q = 0; q < NT_PRE; q++) {
        const int i = tid + q * nthr;
        if (i < nown) {
            const double *nt = A.nt3 + 3 * (size_t)(n0 + i);
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

// Flags of the lower tiles that publish for a tile's nodes (dependency ids in buf; md = meta
// group 3): request this thread's flag / wait until every flag this thread watches is raised.
__device__ __forceinline__ const unsigned int *dep_flag(const StepArgs &A, const int4 md, const char *buf, int d)
{
    const int *sdep = reinterpret_cast<const int *>(buf + 8 * (size_t)A.cap_recs) + A.cap_srcs;
    return A.flag + (d < CAP_DEPS ? sdep[d] : __ldg(A.dep + md.x + d));
}
__device__ __forceinline__ void wait_deps(const StepArgs &A, const int4 md, const char *buf, int tid, int nthr,
                                          unsigned int first_value)
{
    const int ndep = md.y - md.x;
    unsigned int v = first_value;
    for (int d = tid; d < ndep; d += nthr) {
        const unsigned int *f = dep_flag(A, md, buf, d);
        if (d != tid) v = ld_relaxed_u32(f);
        if ((int)(v - A.epoch) < 0) {
            // Lower tiles are started before this one and every CTA of the launch is resident
            // (cooperative launch), so this wait is short; the bound only turns a broken
            // invariant (a lost CTA, a foreign context holding the SMs) into an error the host
            // sees at its next hgpu_sync instead of a hang.
            const long long t0 = clock64();
            do {
                __nanosleep(32);
                v = ld_relaxed_u32(f);
                if (clock64() - t0 > WAIT_DEPS_CYCLES) { report_error(A.err, 2); break; }
            } while ((int)(v - A.epoch) < 0);
        }
    }
    // No fence on this side: a gpu-scope fence makes ptxas invalidate the SM's whole L1 (CCTL.IVALL), once
    // per tile, which cost 4 % of the kernel when measured (r02 call 1).  The partial forces read after the
    // next barrier cannot be stale without it: see the comment above ld_relaxed_u32.
}

// Request the partial forces a tile reads, and the node-table entry of its record nodes
// (cp.async from L2, where the publishers' fences made them visible; no 128-byte line of
// partial[] is shared by two publishers, so a line cached in L1 is never stale).
__device__ __forceinline__ void request_partials(const StepArgs &A, int n0, const int4 mc, const char *buf,
                                                 const FinishBufs &fb, int tid, int nthr)
{
    const uint2 *srec = reinterpret_cast<const uint2 *>(buf);
    const int *ssrc = reinterpret_cast<const int *>(buf + 8 * (size_t)A.cap_recs);
    const int ns = min(mc.w - mc.z, A.cap_srcs), nr = min(mc.y - mc.x, A.cap_recs);
    for (int q = tid; q < ns; q += nthr) {
        const double *pp = A.partial + 3 * (size_t)ssrc[q];
        cp_async8(fb.spart + 3 * q, pp); cp_async8(fb.spart + 3 * q + 1, pp + 1); cp_async8(fb.spart + 3 * q + 2, pp + 2);
    }
    const double *nt = A.nt3 + 3 * (size_t)n0;
    for (int r = tid; r < nr; r += nthr) cp_async8(fb.srm + r, nt + (srec[r].x & 0xffff));
}

// Finish record r of a tile: add the partial forces of the lower tiles to the node's own share
// (own[]) in a fixed order and store the result -- u(t+dt) of a REGULAR node, or the force of a
// SPECIAL node (source term, hanging-node transfer, halo exchange and the list update follow).
__device__ __forceinline__ void finish_record(const StepArgs &A, int n0, const int4 mc, const char *buf,
                                              const FinishBufs &fb, int r, const double (&own)[3], bool fuse)
{
    const uint2 *srec = reinterpret_cast<const uint2 *>(buf);
    const size_t g0 = 3 * (size_t)n0;
    const uint2 rc = srec[r];
    const int slot3 = rc.x & 0xffff, cnt = (rc.x >> 16) & 0xff, first = (int)rc.y;
    const double rmv = fb.srm[r];
    const bool regular = fuse && rmv > 0.0;
    const double scale = regular ? rmv : 1.0;
    double s0 = own[0], s1 = own[1], s2 = own[2];
    for (int k = 0; k < cnt; k++) {
        const int q = first + k;
        s0 = fma(fb.spart[3 * q], scale, s0); s1 = fma(fb.spart[3 * q + 1], scale, s1); s2 = fma(fb.spart[3 * q + 2], scale, s2);
    }
    if (regular) {
        double *o = A.unext + g0 + slot3;
        o[0] = s0; o[1] = s1; o[2] = s2;
    } else {
        // fused launch: own[] is the plain force the tile accumulated; unfused: the tile's share is
        // in the force array already and own[] is zero
        double *fo = A.force + g0 + slot3;
        fo[0] += s0; fo[1] += s1; fo[2] += s2;
    }
}

// firstVector_mu (stiffness.c:351-379): v += shear part of the scaled modes, b = -0.5625 c1
__device__ __forceinline__ void scale_modes_mu_add(const double (&tx)[8], const double (&ty)[8],
                                                   const double (&tz)[8], double b,
                                                   double (&vx)[8], double (&vy)[8], double (&vz)[8])
{
    const double b3 = b * (1.0 / 3.0), b9 = b * (1.0 / 9.0), b27 = b * (10.0 / 27.0);
    const double sxz = b * (tz[3] + tx[1]), sxy = b * (ty[3] + tx[2]), syz = b * (tz[2] + ty[1]);
    const double m4 = b3 * (ty[5] + tz[6] + 2.0 * tx[4]);
    const double m5 = b3 * (tx[4] + tz[6] + 2.0 * ty[5]);
    const double m6 = b3 * (tx[4] + ty[5] + 2.0 * tz[6]);
    vx[1] += sxz; vx[2] += sxy; vx[3] += b3 * (4.0 * tx[3] - 2.0 * (ty[2] + tz[1]));
    vx[4] += m4;  vx[5] += b9 * (7.0 * tx[5] - 2.0 * ty[4]); vx[6] += b9 * (7.0 * tx[6] - 2.0 * tz[4]);
    vx[7] += b27 * tx[7];
    vy[1] += syz; vy[2] += b3 * (4.0 * ty[2] - 2.0 * (tx[3] + tz[1])); vy[3] += sxy;
    vy[4] += b9 * (7.0 * ty[4] - 2.0 * tx[5]); vy[5] += m5; vy[6] += b9 * (7.0 * ty[6] - 2.0 * tz[5]);
    vy[7] += b27 * ty[7];
    vz[1] += b3 * (4.0 * tz[1] - 2.0 * (tx[3] + ty[2])); vz[2] += syz; vz[3] += sxz;
    vz[4] += b9 * (7.0 * tz[4] - 2.0 * tx[6]); vz[5] += b9 * (7.0 * tz[5] - 2.0 * ty[6]); vz[6] += m6;
    vz[7] += b27 * tz[7];
}

// firstVector_kappa (stiffness.c:321-349): v += volumetric part, kappa = -0.5625 (c2 + 2/3 c1)
__device__ __forceinline__ void scale_modes_kappa_add(const double (&tx)[8], const double (&ty)[8],
                                                      const double (&tz)[8], double kap,
                                                      double (&vx)[8], double (&vy)[8], double (&vz)[8])
{
    const double k3 = kap * (1.0 / 3.0), k9 = kap * (1.0 / 9.0);
    const double div = kap * (tx[3] + ty[2] + tz[1]);
    const double exy = k3 * (tx[5] + ty[4]), exz = k3 * (tx[6] + tz[4]), eyz = k3 * (ty[6] + tz[5]);
    vx[3] += div; vx[5] += exy; vx[6] += exz; vx[7] += k9 * tx[7];
    vy[2] += div; vy[4] += exy; vy[6] += eyz; vy[7] += k9 * ty[7];
    vz[1] += div; vz[4] += exz; vz[5] += eyz; vz[7] += k9 * tz[7];
}

// One family (shear or kappa) of calc_conv + the damping vector of constant_Q_addforce
// (damping.c:126-169 / 173-216 and 256-311 / 315-371) for one element: the memory variables are
// advanced in place (global memory, entry-chunked layout) and the 8 x 3 damping vector
//   d = (b / rmax) (u1 - u2) - (a0 f0 + a1 f1) + u1     (or u1 when a0 + a1 + b = 0)
// is taken straight to mode space, t[c][m] = sum_i S[m][i] d[i][c] (the sign matrix of aTransposeU,
// stiffness.c:260-288; row 0 is dropped as the reference zeroes it), without ever being stored.
__device__ __forceinline__ void bkt_family(double *conv, size_t entry, int fam, float a0f, float a1f, float bf,
                                           float g0f, float g1f, double rmax, const double *su1, const double *su2,
                                           const uint4 slots, double (&tx)[8], double (&ty)[8], double (&tz)[8])
{
    const bool advance = g0f != 0.f && g1f != 0.f;           // damping.c:126, 173
    const double a0 = a0f, a1 = a1f, b = bf;
    const bool damped = (a0 + a1 + b) != 0.0;                 // damping.c:262, 321
    const double g0 = (double)g0f * rmax, g1 = (double)g1f * rmax;
    const double k1 = 0.5 * g0, k2 = k1 * (1.0 - g0), k3 = 0.5 * g1, k4 = k3 * (1.0 - g1);
    const double e0 = exp(-g0), e1 = exp(-g1), cb = b / rmax;
#pragma unroll
    for (int m = 0; m < 8; m++) { tx[m] = 0.0; ty[m] = 0.0; tz[m] = 0.0; }
    double *p0 = conv + conv_index(entry, fam * 48), *p1 = p0 + 24 * 32;
    // two nodes (an x pair: ix = 0, 1) per trip, not unrolled further: 12 loads of state in flight
    // per thread keep HBM busy without the whole element's 48 values sitting in registers
#pragma unroll 1
    for (int h = 0; h < 4; h++) {
        const uint32_t w = h == 0 ? slots.x : h == 1 ? slots.y : h == 2 ? slots.z : slots.w;
        const uint32_t o[2] = {w & 0xffffu, w >> 16};
        double f0[2][3], f1[2][3];
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                f0[q][c] = p0[(6 * h + 3 * q + c) * 32];
                f1[q][c] = p1[(6 * h + 3 * q + c) * 32];
            }
        const bool py = h & 1, pz = h & 2;                    // node = q + 2 h: iy = h & 1, iz = h >> 1
#pragma unroll
        for (int q = 0; q < 2; q++) {
            double d[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double x1 = su1[o[q] + c], x2 = su2[o[q] + c];
                double v0 = f0[q][c], v1 = f1[q][c];
                if (advance) {
                    v0 = fma(e0, v0, k2 * x1 + k1 * x2);
                    v1 = fma(e1, v1, k4 * x1 + k3 * x2);
                    p0[(6 * h + 3 * q + c) * 32] = v0; p1[(6 * h + 3 * q + c) * 32] = v1;
                }
                d[c] = damped ? (cb * (x1 - x2) - (a0 * v0 + a1 * v1)) + x1 : x1;
            }
            // mode signs (see wht_forward): 1 = z, 2 = y, 3 = x, 4 = yz, 5 = xz, 6 = xy, 7 = xyz
            const bool px = q == 1;
#define HGPU_ACC(T, v)                                                                     \
            T[1] += pz ? (v) : -(v); T[2] += py ? (v) : -(v); T[3] += px ? (v) : -(v);          \
            T[4] += (py == pz) ? (v) : -(v); T[5] += (px == pz) ? (v) : -(v);                   \
            T[6] += (px == py) ? (v) : -(v); T[7] += ((px != py) != pz) ? (v) : -(v);
            HGPU_ACC(tx, d[0]) HGPU_ACC(ty, d[1]) HGPU_ACC(tz, d[2])
#undef HGPU_ACC
        }
    }
}


// ------------------------------------------------------------------------------------------
// STRUCTURED tiles (hgpu_internal.h): one aligned 8x8x8 cell of equal elements of one material, walked by
// tile_loop<..., KIND = 1> with 512 threads and one CTA per SM.  No slot table is needed:
//   pre-pass       the damped displacement w = u1 + beta (u1 - u2) is formed ONCE PER NODE (729 nodes) right
//                  after the tile has landed, into three padded planes [component][z][y][x] with row stride 12
//                  and plane stride 108: lanes = (x & 3, y), so the 16 lanes of a half-warp (4 x values, 4
//                  consecutive rows) always hit 16 different 8-byte banks (12 y mod 16 = 0, 12, 8, 4)
//   element phase  thread = element (x, y, z): 24 conflict-free gathers, the factored operator with the tile's
//                  coefficients, 24 plain stores F[(3 j + c) * 512 + thread]  (j = corner, c = component)
//   node phase     thread = node (x, y, z) (and far-face node tid < 217): the forces of its (up to) eight
//                  elements are summed in registers in a fixed order -- no read-modify-write, no accumulator --
//                  far-face nodes are published as partial forces, owned ones go to a slot-order pass that adds
//                  the inertia term, scales by 1/mass and stores u(t+dt) as coalesced 24-byte rows
// Variants that ran before this one (z pairs with register carry, four ordered accumulator passes, a side
// array for the x = 4 column, shuffle-combined updates, tiles from shared counters, dependency-level order)
// are recorded with their measurements in profiles/README.md; none of them, nor this one, beats the
// slot-table path on the bench, so the path is opt-in.
// ------------------------------------------------------------------------------------------
constexpr int SP_ROW = 12, SP_Z = 108, SP_C = 972, SP_TOTAL = 3 * SP_C;     // doubles
constexpr int SF_TOTAL = 24 * 512;                                          // element forces of one structured tile

// tile-local slot (Morton for the 512 owned nodes, canonical far-face order for the 217 others,
// hgpu_internal.h) -> offset of the node inside one component plane
HGPU_HD int sp_of_slot(int s)
{
    int x, y, z;
    if (s < 512) {
        x = (s & 1) | ((s >> 2) & 2) | ((s >> 4) & 4);
        y = ((s >> 1) & 1) | ((s >> 3) & 2) | ((s >> 5) & 4);
        z = ((s >> 2) & 1) | ((s >> 4) & 2) | ((s >> 6) & 4);
    } else {
        const int h = s - 512;
        if (h < 81)       { z = 8; y = h / 9; x = h - 9 * y; }
        else if (h < 153) { const int k = h - 81; y = 8; z = k / 9; x = k - 9 * z; }
        else              { const int k = h - 153; x = 8; z = k >> 3; y = k & 7; }
    }
    return z * SP_Z + y * SP_ROW + x;
}

// the four nodes (x + dx, y + dy) of one level of a thread's column, one component plane
HGPU_HD void gather_face(const double *plane, int o, double &w0, double &w1, double &w2, double &w3)
{
    w0 = plane[o]; w1 = plane[o + 1]; w2 = plane[o + SP_ROW]; w3 = plane[o + SP_ROW + 1];
}


// WPASS (MODE 1, fused update; opt-in, HGPU_FLAG_WPASS): on a tile whose entries share one beta, the
// damped displacement w = u1 + beta (u1 - u2) is formed ONCE PER STAGED NODE right after the tile has
// landed (in place, over the u1 stage) instead of once per element corner, so the element phase
// gathers 24 values per element instead of 48.  The inertia term m2 u1 - m1 u2 of the owned REGULAR
// nodes, which needs u1 and u2, is put into the accumulator in the same pass (the forces are then
// added on top of it and the sum is scaled by 1/mass where the default path adds the inertia term
// last): m2, m1 of a tile's nodes are therefore prefetched one tile ahead.
// tile_loop: one CTA walks tiles t0, t0 + G, t0 + 2 G, ... (< tend) of the processing order.
//   KIND 0: slot-table tiles (any shape)      KIND 1: structured tiles (fused update, MODE 0 or 1)
// A STRUCT launch runs both loops, on different CTAs (step_kernel below): each loop then carries only its
// own registers between tiles -- the two paths in ONE loop spilled (r02 call 2).
// Tiles are handed out either statically (t0, t0 + G, ...; queue == nullptr) or from a counter shared by
// the CTAs of the launch (queue: next tile = t0 + atomicAdd(queue, 1); G unused): in both cases a CTA's tiles
// ascend, which is all the deadlock argument needs.
template <int MODE, bool DENSE, int THREADS, bool WPASS, int KIND>
__device__ __forceinline__ void tile_loop(const StepArgs &A, const int t0, const int G, const int tend, int *queue)
{
    constexpr bool STRUCT = KIND == 1;
    static_assert(!WPASS || (MODE == 1 && !DENSE), "WPASS is a variant of the Rayleigh + effective kernel");
    static_assert(!STRUCT || ((MODE == 0 || MODE == 1) && !DENSE && !WPASS && THREADS == 512),
                  "STRUCT: effective stiffness with or without Rayleigh damping, 512 threads");
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
    //   flags         : of the PREVIOUS tile's publishers; requested at accumulation pass 0 of this
    //                   tile's last round, checked at pass 3
    //   partial forces: cp.async after pass 3, used when the previous tile is finished at the end
    //                   of this iteration
    //   entries       : registers; enext -> ecur at the top of a round, then the following round's
    //                   entry is requested
    //   halo ids      : registers; requested before the accumulation passes of a tile's last round
    //                   for the tile that is staged at the top of the next iteration
    //   node tables   : registers; requested before the accumulation passes of the last round
    __shared__ __align__(16) int smeta[META_RING][META_INTS];
    // ids of this CTA's tiles, position p in slot p & (META_RING - 1); >= tend = none.  Positions 0..3 now,
    // position it + 4 during iteration it (the offsets of position it + 3 are requested at its top).
    __shared__ int stile[META_RING];
    __shared__ __align__(8) unsigned long long sbar[2];      // STRUCT: completion of the bulk copies into either stage
    if (tid == 0) {
#pragma unroll
        for (int p = 0; p < 4; p++) stile[p] = queue ? t0 + atomicAdd(queue, 1) : t0 + p * G;
        if (STRUCT) { mbar_init(&sbar[0], 1); mbar_init(&sbar[1], 1); }
    }
    if (STRUCT) fence_proxy_async();
    __syncthreads();
    int t = stile[0];
    if (t >= tend) return;
    for (int k = tid; k < A3; k += nthr) acc[k] = 0.0;
    fetch_meta_async(A, t, smeta[0], tid);
    if (stile[1] < tend) fetch_meta_async(A, stile[1], smeta[1], tid);
    if (stile[2] < tend) fetch_meta_async(A, stile[2], smeta[2], tid);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    int hid[HALO_PRE];
    {
        const int4 ma = meta_group(smeta[0], 0);
        load_halo_ids(A, ma, tid, nthr, hid);
        if (STRUCT)           stage_tile_struct<U2E>(A, ma, smem, smem + S3, &sbar[0], tid, hid);
        else if (U2E || fuse) stage_tile<true, U2E>(A, ma, smem, smem + S3, tid, nthr, hid);
        else                  stage_tile<false, false>(A, ma, smem, smem + S3, tid, nthr, hid);
        cp_async_commit();
        if (stile[1] < tend) load_halo_ids(A, meta_group(smeta[1], 0), tid, nthr, hid);
    }
    Entry ecur, enext;
    enext.s = make_uint4(0, 0, 0, 0); enext.c1 = enext.c2 = enext.beta = 0.0;
    if (!STRUCT) {
        const int4 mb = meta_group(smeta[0], 1);
        if (tid < mb.y - mb.x) enext = load_entry<MODE>(A, mb.x + tid);
    }

    // WPASS: m2, m1 of the owned nodes and beta of the tile that is processed NEXT (registers)
    double ntm[NT_PRE][2];
    double beta_pref = 0.0;
    if (WPASS) {
        const int4 ma = meta_group(smeta[0], 0);
#pragma unroll
        for (int q = 0; q < NT_PRE; q++) {
            const int i = tid + q * nthr;
            ntm[q][0] = ntm[q][1] = 0.0;
            if (i < ma.y - ma.x) {
                const double *nt = A.nt3 + 3 * (size_t)(ma.x + i);
                ntm[q][0] = ldg_f64_pinned(nt + 1); ntm[q][1] = ldg_f64_pinned(nt + 2);
            }
        }
        beta_pref = ldg_f64_pinned(A.tile_beta + t);
    }

    double spre[3], cpre[3];             // STRUCT: node-table row (slot tid) and coefficients of the NEXT structured tile
    bool pre_ok = false;
    for (int it = 0;; it++) {
        double *su1 = smem + (it & 1) * stage_doubles;
        double *su2 = su1 + S3;
        const int tn = stile[(it + 1) & (META_RING - 1)];
        int popped = 0;
        const bool has_next = tn < tend;
        const bool has_nn = stile[(it + 2) & (META_RING - 1)] < tend;
        const int *m_cur = smeta[it & (META_RING - 1)], *m_nxt = smeta[(it + 1) & (META_RING - 1)];
        const int *m_nn = smeta[(it + 2) & (META_RING - 1)], *m_prv = smeta[(it - 1) & (META_RING - 1)];
        const char *fb_prv = fbuf + ((it - 1) & 1) * fbuf_bytes;
        char *fb_cur = fbuf + (it & 1) * fbuf_bytes;
        // STRUCT: node-table rows of this thread's two owned nodes and the tile's coefficients travel in
        // registers from the previous structured tile's second round (spre / cpre); a structured tile that
        // follows a slot-table tile (or is the CTA's first) requests them here, before the landing wait
        constexpr bool st = STRUCT;           // every tile of this loop is a structured one (host-side lists)
        double snt[3], scf[3];
        if (STRUCT && st) {
            if (!pre_ok) {
                const double *nt = A.nt3 + 3 * (size_t)(meta_group(m_cur, 0).x + tid);
                spre[0] = ldg_f64_pinned(nt); spre[1] = ldg_f64_pinned(nt + 1); spre[2] = ldg_f64_pinned(nt + 2);
                const double *tc = A.tile_coef + 4 * (size_t)t;
                cpre[0] = ldg_f64_pinned(tc); cpre[1] = ldg_f64_pinned(tc + 1); cpre[2] = ldg_f64_pinned(tc + 2);
            }
            snt[0] = spre[0]; snt[1] = spre[1]; snt[2] = spre[2];
            scf[0] = cpre[0]; scf[1] = cpre[1]; scf[2] = cpre[2];
        }
        pre_ok = false;
        cp_async_wait_all();
        if (STRUCT) {
            mbar_wait(&sbar[it & 1], (unsigned)(it >> 1) & 1u);     // the bulk part of tile `it`
            fence_proxy_async();              // this thread's ld/st.shared of the other stage, before its next bulk copy
        }
        __syncthreads();                      // tile `it` has landed; everyone is done with tile it-1's stage
        if (it > 0 && tid == 0) publish_flag(A.flag + meta_group(m_prv, 3).z, A.epoch);
        stage_finish(A, meta_group(m_cur, 2), meta_group(m_cur, 3), fb_cur, tid, nthr);
        if (has_next) {
            double *n1 = smem + ((it + 1) & 1) * stage_doubles;
            const int4 nxt = meta_group(m_nxt, 0);
            if (STRUCT)           stage_tile_struct<U2E>(A, nxt, n1, n1 + S3, &sbar[(it + 1) & 1], tid, hid);
            else if (U2E || fuse) stage_tile<true, U2E>(A, nxt, n1, n1 + S3, tid, nthr, hid);
            else                  stage_tile<false, false>(A, nxt, n1, n1 + S3, tid, nthr, hid);
            const int t3 = stile[(it + 3) & (META_RING - 1)];
            if (t3 < tend) fetch_meta_async(A, t3, smeta[(it + 3) & (META_RING - 1)], tid);
            // position it + 4: requested here, stored before the first barrier of the tail (so that thread 0
            // does not wait for the atomic), read from the top of iteration it + 1 on
            if (tid == 0) popped = queue ? t0 + atomicAdd(queue, 1) : t0 + (it + 4) * G;
        }
        cp_async_commit();
        // what the element phase needs of this tile's offsets
        int n0, nown, eb, ne, ncore, nxt_eb = 0, nxt_ne = 0;
        {
            const int4 ma = meta_group(m_cur, 0), mb = meta_group(m_cur, 1);
            n0 = ma.x; nown = ma.y - ma.x; eb = mb.x; ne = mb.y - mb.x; ncore = mb.z;
            if (has_next) { const int4 nb = meta_group(m_nxt, 1); nxt_eb = nb.x; nxt_ne = nb.y - nb.x; }
        }
        const int nown3 = 3 * nown;
        // the previous tile is finished during this one: its publishers' flags are polled and its
        // partial forces requested between the accumulation passes of the last round
        const bool prv_pending = it > 0;
        unsigned int flag_value = A.epoch;
        if (STRUCT && st) {
            // ---- structured tile, 512 threads (see the comment above sp_of_slot) -------------------------
            // thread = element (x, y, z) of the cell in the element phase, = node (x, y, z) (and, tid < 217, far-
            // face node tid) in the node phase, = staged slot tid (and 512 + tid) where data is in slot order
            const int x = (tid & 3) | ((tid >> 3) & 4), y = (tid >> 2) & 7, z = tid >> 6;
            double *W = acc;                    // padded planes of w; later the owned nodes' forces in slot order
            double *F = acc + SP_TOTAL;         // element forces, F[(3 j + c) * 512 + element thread]
            // pre-pass: w = u1 + beta (u1 - u2) of slot tid and of far-face slot 512 + tid, into the planes
            // (the raw stage stays: the inertia term is taken from it when the tile is settled)
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (q == 0 || tid < 217) {
                    const int sl = q == 0 ? tid : 512 + tid, k = 3 * sl, sp = sp_of_slot(sl);
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const double x1 = su1[k + c];
                        W[sp + c * SP_C] = MODE == 1 ? fma(scf[2], x1 - su2[k + c], x1) : x1;
                    }
                }
            }
            if (prv_pending) {                  // flags of the previous tile's publishers: checked before the node phase
                const int4 md = meta_group(m_prv, 3);
                if (tid < md.y - md.x) flag_value = ld_relaxed_u32(dep_flag(A, md, fb_prv, tid));
            }
            __syncthreads();
            // ---- element phase: gather 24, evaluate, store the 24 corner forces (no accumulator, no order) ----
            {
                const double ca = -0.5625 * (scf[1] + 2.0 * scf[0]), cc = -0.5625 * scf[1], cb = -0.5625 * scf[0];
                const int o = z * SP_Z + y * SP_ROW + x;
                double wx[8], wy[8], wz[8], tx[8], ty[8], tz[8], fx[8], fy[8], fz[8];
                gather_face(W, o, wx[0], wx[1], wx[2], wx[3]);
                gather_face(W + SP_C, o, wy[0], wy[1], wy[2], wy[3]);
                gather_face(W + 2 * SP_C, o, wz[0], wz[1], wz[2], wz[3]);
                gather_face(W, o + SP_Z, wx[4], wx[5], wx[6], wx[7]);
                gather_face(W + SP_C, o + SP_Z, wy[4], wy[5], wy[6], wy[7]);
                gather_face(W + 2 * SP_C, o + SP_Z, wz[4], wz[5], wz[6], wz[7]);
                wht_forward(wx, tx); wht_forward(wy, ty); wht_forward(wz, tz);
                scale_modes(tx, ty, tz, ca, cc, cb, wx, wy, wz);            // w* reused as the scaled modes
                wht_inverse(wx, fx); wht_inverse(wy, fy); wht_inverse(wz, fz);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    F[(3 * j) * 512 + tid] = fx[j]; F[(3 * j + 1) * 512 + tid] = fy[j]; F[(3 * j + 2) * 512 + tid] = fz[j];
                }
            }
            // what the next tile needs in registers
            if (has_nn) load_halo_ids(A, meta_group(m_nn, 0), tid, nthr, hid);
            if (has_next) {
                const double *nt = A.nt3 + 3 * (size_t)(meta_group(m_nxt, 0).x + tid);
                spre[0] = ldg_f64_pinned(nt); spre[1] = ldg_f64_pinned(nt + 1); spre[2] = ldg_f64_pinned(nt + 2);
                const double *tc = A.tile_coef + 4 * (size_t)tn;
                cpre[0] = ldg_f64_pinned(tc); cpre[1] = ldg_f64_pinned(tc + 1); cpre[2] = ldg_f64_pinned(tc + 2);
                pre_ok = true;
            }
            if (prv_pending) wait_deps(A, meta_group(m_prv, 3), fb_prv, tid, nthr, flag_value);
            __syncthreads();                    // every element's forces are in F; W is free
            if (prv_pending) {
                request_partials(A, meta_group(m_prv, 0).x, meta_group(m_prv, 2), fb_prv, fb, tid, nthr);
                cp_async_commit();
            }
            // ---- node phase: every node sums the forces of its (up to) eight elements, in a fixed order ----
            {
                double T0 = 0.0, T1 = 0.0, T2 = 0.0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int dx = j & 1, dy = (j >> 1) & 1, dz = j >> 2;
                    if (x >= dx && y >= dy && z >= dz) {
                        const int ex = x - dx;
                        const int te = (ex & 3) | ((ex & 4) << 3) | ((y - dy) << 2) | ((z - dz) << 6);
                        T0 += F[(3 * j) * 512 + te]; T1 += F[(3 * j + 1) * 512 + te]; T2 += F[(3 * j + 2) * 512 + te];
                    }
                }
                const int k = 3 * ((x & 1) | ((y & 1) << 1) | ((z & 1) << 2) | ((x & 2) << 2) | ((y & 2) << 3) | ((z & 2) << 4) |
                                   ((x & 4) << 4) | ((y & 4) << 5) | ((z & 4) << 6));          // Morton slot of node (x, y, z)
                W[k] = T0; W[k + 1] = T1; W[k + 2] = T2;
            }
            if (tid < 217) {                    // far-face node tid (hgpu_internal.h): published as partial force tid
                int X, Y, Z;
                if (tid < 81)       { Z = 8; Y = tid / 9; X = tid - 9 * Y; }
                else if (tid < 153) { const int q = tid - 81; Y = 8; Z = q / 9; X = q - 9 * Z; }
                else                { const int q = tid - 153; X = 8; Z = q >> 3; Y = q & 7; }
                double T0 = 0.0, T1 = 0.0, T2 = 0.0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int ex = X - (j & 1), ey = Y - ((j >> 1) & 1), ez = Z - (j >> 2);
                    if (ex >= 0 && ex < 8 && ey >= 0 && ey < 8 && ez >= 0 && ez < 8) {
                        const int te = (ex & 3) | ((ex & 4) << 3) | (ey << 2) | (ez << 6);
                        T0 += F[(3 * j) * 512 + te]; T1 += F[(3 * j + 1) * 512 + te]; T2 += F[(3 * j + 2) * 512 + te];
                    }
                }
                double *dst = A.partial + 3 * ((size_t)meta_group(m_cur, 0).z + tid);
                __stcg(dst, T0); __stcg(dst + 1, T1); __stcg(dst + 2, T2);
            }
            if (has_next && tid == 0) stile[(it + 4) & (META_RING - 1)] = popped;
            cp_async_wait_all();
            __syncthreads();                    // the owned nodes' forces are in W, in slot order
            // ---- settle slot tid: central difference of a REGULAR node, the plain force otherwise; hand it on ----
            {
                const int k = 3 * tid;
                double *o = A.unext + 3 * (size_t)(n0 + tid);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    double v = W[k + c];
                    if (snt[0] > 0.0) v = (v + (snt[1] * su1[k + c] - snt[2] * su2[k + c])) * snt[0];
                    W[k + c] = v;               // record nodes take it from here (pend)
                    o[c] = v;
                }
            }
        } else {
        double ntv[NT_PRE][3];
        bool uni = false;                     // WPASS: this tile's entries share one beta
        if (WPASS) {
            const double beta_t = beta_pref;
            uni = beta_t == beta_t;
            const double *nt_own = A.nt3 + 3 * (size_t)n0;
            // owned nodes: inertia term into the (zeroed) accumulator -- nt3 holds m2 = m1 = 0 for
            // SPECIAL nodes, whose accumulator must end up as the plain force -- and w over u1
#pragma unroll
            for (int q = 0; q < NT_PRE; q++) {
                const int i = tid + q * nthr;
                if (i < nown) {
                    const int k = 3 * i;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const double x1 = su1[k + c], x2 = su2[k + c];
                        acc[k + c] = ntm[q][0] * x1 - ntm[q][1] * x2;
                        if (uni) su1[k + c] = fma(beta_t, x1 - x2, x1);
                    }
                }
            }
            for (int i = tid + NT_PRE * nthr; i < nown; i += nthr) {
                const double m2 = __ldg(nt_own + 3 * i + 1), m1 = __ldg(nt_own + 3 * i + 2);
                const int k = 3 * i;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double x1 = su1[k + c], x2 = su2[k + c];
                    acc[k + c] = m2 * x1 - m1 * x2;
                    if (uni) su1[k + c] = fma(beta_t, x1 - x2, x1);
                }
            }
            // gathered nodes
            if (uni) {
                const int4 ma = meta_group(m_cur, 0);
                const int nh3 = 3 * (ma.w - ma.z);
                for (int k = tid; k < nh3; k += nthr) {
                    const double x1 = su1[nown3 + k], x2 = su2[nown3 + k];
                    su1[nown3 + k] = fma(beta_t, x1 - x2, x1);
                }
            }
            __syncthreads();
        }

        // ---- element forces, accumulated per staged node ----------------------------------------
        for (int base = 0; base < ne; base += nthr) {
            const bool act = base + tid < ne;
            const bool last = base + nthr >= ne;
            // a core entry adds to every corner (owned or published); an extra entry of a self
            // tile only to the owned ones
            const uint32_t lim = base + tid < ncore ? 0xffffffffu : (uint32_t)nown3;
            ecur = enext;
            if (last && prv_pending) {           // flags of the previous tile's publishers: needed at pass 3
                const int4 md = meta_group(m_prv, 3);
                if (tid < md.y - md.x) flag_value = ld_relaxed_u32(dep_flag(A, md, fb_prv, tid));
            }
            // next round's entry (or the first round of the next tile) rides along with the math
            if (!last) {
                if (base + nthr + tid < ne) enext = load_entry<MODE>(A, eb + base + nthr + tid);
            } else if (tid < nxt_ne) {
                enext = load_entry<MODE>(A, nxt_eb + tid);
            }
            double fx[8], fy[8], fz[8];
            uint32_t sl[8];
            if (act) {
                sl[0] = ecur.s.x & 0xffffu; sl[1] = ecur.s.x >> 16; sl[2] = ecur.s.y & 0xffffu; sl[3] = ecur.s.y >> 16;
                sl[4] = ecur.s.z & 0xffffu; sl[5] = ecur.s.z >> 16; sl[6] = ecur.s.w & 0xffffu; sl[7] = ecur.s.w >> 16;
                double wx[8], wy[8], wz[8];
                if (MODE == 3) {
                    // BKT: calc_conv + constant_Q_addforce (damping.c:110-416), elastic term included
                    const size_t entry = (size_t)(eb + base + tid);
                    {
                        // the warp's 24 KB of memory variables: start all of it towards L2 now (192
                        // lines, 6 per lane), the per-pair loads below then find it there or in flight
                        const int lane = tid & 31;
#pragma unroll
                        for (int q = 0; q < 3; q++) {
                            const double *row = A.conv + conv_index(entry & ~(size_t)31, lane + 32 * q);
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 16));
                        }
                    }
                    const double *cr = A.ent_bkt + 8 * entry;
                    const double c1 = __ldg(cr), c2 = __ldg(cr + 1);
                    const float2 q0 = __ldg(reinterpret_cast<const float2 *>(cr + 2)), q1 = __ldg(reinterpret_cast<const float2 *>(cr + 3));
                    const float2 q2 = __ldg(reinterpret_cast<const float2 *>(cr + 4)), q3 = __ldg(reinterpret_cast<const float2 *>(cr + 5));
                    const float2 q4 = __ldg(reinterpret_cast<const float2 *>(cr + 6));
                    double tx[8], ty[8], tz[8];
#pragma unroll
                    for (int m = 0; m < 8; m++) { wx[m] = 0.0; wy[m] = 0.0; wz[m] = 0.0; }
                    bkt_family(A.conv, entry, 0, q0.x, q0.y, q1.x, q1.y, q2.x, A.rmax, su1, su2, ecur.s, tx, ty, tz);
                    scale_modes_mu_add(tx, ty, tz, -0.5625 * c1, wx, wy, wz);
                    bkt_family(A.conv, entry, 1, q2.y, q3.x, q3.y, q4.x, q4.y, A.rmax, su1, su2, ecur.s, tx, ty, tz);
                    scale_modes_kappa_add(tx, ty, tz, -0.5625 * (c2 + (2.0 / 3.0) * c1), wx, wy, wz);
                    wht_inverse(wx, fx); wht_inverse(wy, fy); wht_inverse(wz, fz);
                } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int o = sl[j];
                    const double ax = su1[o], ay = su1[o + 1], az = su1[o + 2];
                    if (MODE == 0 || (WPASS && uni)) { wx[j] = ax; wy[j] = ay; wz[j] = az; }
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
            }
            if (last) {
                if (WPASS) {                     // settle only scales: 1/mass of this tile's nodes
#pragma unroll
                    for (int q = 0; q < NT_PRE; q++) {
                        const int i = tid + q * nthr;
                        ntv[q][0] = i < nown ? ldg_f64_pinned(A.nt3 + 3 * (size_t)(n0 + i)) : 0.0;
                    }
                } else if (fuse) load_node_tables(A, n0, nown, tid, nthr, ntv);
                if (has_nn) load_halo_ids(A, meta_group(m_nn, 0), tid, nthr, hid);
                if (WPASS && has_next) {         // m2, m1 and beta of the next tile
                    const int4 na = meta_group(m_nxt, 0);
#pragma unroll
                    for (int q = 0; q < NT_PRE; q++) {
                        const int i = tid + q * nthr;
                        if (i < na.y - na.x) {
                            const double *nt = A.nt3 + 3 * (size_t)(na.x + i);
                            ntm[q][0] = ldg_f64_pinned(nt + 1); ntm[q][1] = ldg_f64_pinned(nt + 2);
                        }
                    }
                    beta_pref = ldg_f64_pinned(A.tile_beta + tn);
                }
            }
            // A node is corner j of at most one element (leaf octants do not overlap), so within
            // pass j every accumulator is touched by at most one thread: no atomics, fixed order.
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (act && sl[j] < lim) {
                    const int o = sl[j];
                    acc[o] += fx[j]; acc[o + 1] += fy[j]; acc[o + 2] += fz[j];
                }
                if (j == 3 && last && prv_pending) wait_deps(A, meta_group(m_prv, 3), fb_prv, tid, nthr, flag_value);
                __syncthreads();
                if (j == 3 && last && prv_pending) {
                    request_partials(A, meta_group(m_prv, 0).x, meta_group(m_prv, 2), fb_prv, fb, tid, nthr);
                    cp_async_commit();
                }
            }
        }
        if (ne == 0) {                          // a tile of element-less nodes: keep the pipeline fed
            if (tid < nxt_ne) enext = load_entry<MODE>(A, nxt_eb + tid);
            if (has_nn) load_halo_ids(A, meta_group(m_nn, 0), tid, nthr, hid);
            if (WPASS) {
#pragma unroll
                for (int q = 0; q < NT_PRE; q++) {
                    const int i = tid + q * nthr;
                    ntv[q][0] = i < nown ? ldg_f64_pinned(A.nt3 + 3 * (size_t)(n0 + i)) : 0.0;
                }
            } else if (fuse) load_node_tables(A, n0, nown, tid, nthr, ntv);
            if (WPASS && has_next) {
                const int4 na = meta_group(m_nxt, 0);
#pragma unroll
                for (int q = 0; q < NT_PRE; q++) {
                    const int i = tid + q * nthr;
                    if (i < na.y - na.x) {
                        const double *nt = A.nt3 + 3 * (size_t)(na.x + i);
                        ntm[q][0] = ldg_f64_pinned(nt + 1); ntm[q][1] = ldg_f64_pinned(nt + 2);
                    }
                }
                beta_pref = ldg_f64_pinned(A.tile_beta + tn);
            }
            if (prv_pending) {
                const int4 md = meta_group(m_prv, 3);
                if (tid < md.y - md.x) flag_value = ld_relaxed_u32(dep_flag(A, md, fb_prv, tid));
                wait_deps(A, md, fb_prv, tid, nthr, flag_value);
                __syncthreads();
                request_partials(A, meta_group(m_prv, 0).x, meta_group(m_prv, 2), fb_prv, fb, tid, nthr);
                cp_async_commit();
            }
        }

        // ---- publish what the core elements added to nodes of higher tiles ----------------------
        {
            const int np3 = 3 * meta_group(m_cur, 1).w;
            double *dst = A.partial + 3 * (size_t)meta_group(m_cur, 0).z;
            for (int k = tid; k < np3; k += nthr) {
                __stcg(dst + k, acc[nown3 + k]);
                acc[nown3 + k] = 0.0;
            }
        }
        // ---- this tile's own share of the update, in place ---------------------------------------
        if (fuse && WPASS) {
            // the inertia term is in the accumulator already: scale by 1/mass
#pragma unroll
            for (int q = 0; q < NT_PRE; q++) {
                const int i = tid + q * nthr;
                if (i < nown && ntv[q][0] > 0.0) {
                    acc[3 * i] *= ntv[q][0]; acc[3 * i + 1] *= ntv[q][0]; acc[3 * i + 2] *= ntv[q][0];
                }
            }
            for (int i = tid + NT_PRE * nthr; i < nown; i += nthr) {
                const double rm = __ldg(A.nt3 + 3 * (size_t)(n0 + i));
                if (rm > 0.0) { acc[3 * i] *= rm; acc[3 * i + 1] *= rm; acc[3 * i + 2] *= rm; }
            }
        } else if (fuse) {
#pragma unroll
            for (int q = 0; q < NT_PRE; q++) {
                const int i = tid + q * nthr;
                if (i < nown) settle_node(acc, su1, su2, i, ntv[q][0], ntv[q][1], ntv[q][2]);
            }
            for (int i = tid + NT_PRE * nthr; i < nown; i += nthr) {
                const double *nt = A.nt3 + 3 * (size_t)(n0 + i);
                settle_node(acc, su1, su2, i, __ldg(nt), __ldg(nt + 1), __ldg(nt + 2));
            }
        }
        }
        if (!(STRUCT && st) && has_next && tid == 0) stile[(it + 4) & (META_RING - 1)] = popped;
        if (!(STRUCT && st)) cp_async_wait_all();   // partial forces of the previous tile, this tile's finish data
        __syncthreads();                        // ... and every thread's published partial forces are written
        // (the flag is raised at the top of the next iteration: by then the stores have long been
        // acknowledged and the fence in front of the flag costs thread 0 next to nothing)

        // ---- record nodes: the previous tile's own share leaves pend, this tile's enters ---------
        // (records are at most one per thread: TileCaps.max_recs <= blockDim.x)
        double own_prv[3] = {0.0, 0.0, 0.0};
        const int nrec_prv = prv_pending ? min(meta_group(m_prv, 2).y - meta_group(m_prv, 2).x, A.cap_recs) : 0;
        {
            const int4 mc = meta_group(m_cur, 2);
            const int nrec = min(mc.y - mc.x, A.cap_recs);
            if (fuse) {
                if (tid < nrec_prv) { own_prv[0] = fb.pend[3 * tid]; own_prv[1] = fb.pend[3 * tid + 1]; own_prv[2] = fb.pend[3 * tid + 2]; }
                if (tid < nrec) {
                    const int slot3 = reinterpret_cast<const uint2 *>(fb_cur)[tid].x & 0xffff;
                    // (STRUCT: acc = the planes' region, which holds the settled owned nodes in slot order here)
                    fb.pend[3 * tid] = acc[slot3]; fb.pend[3 * tid + 1] = acc[slot3 + 1]; fb.pend[3 * tid + 2] = acc[slot3 + 2];
                }
            }
        }
        __syncthreads();
        // ---- hand this tile on: rows of record nodes are rewritten when the tile is finished,
        //      rows of SPECIAL nodes by the special-node update --------------------------------
        {
            const size_t g0 = 3 * (size_t)n0;
            if (STRUCT && st) {
                // handed on when it was settled; nothing accumulates in shared memory on this path
            } else if (fuse) {
                double2 *dst = reinterpret_cast<double2 *>(A.unext + g0);         // g0 is even
                for (int k = tid; k < (nown3 >> 1); k += nthr) {
                    dst[k] = make_double2(acc[2 * k], acc[2 * k + 1]);
                    acc[2 * k] = 0.0; acc[2 * k + 1] = 0.0;
                }
                if ((nown3 & 1) && tid == 0) { A.unext[g0 + nown3 - 1] = acc[nown3 - 1]; acc[nown3 - 1] = 0.0; }
            } else {
                for (int k = tid; k < nown3; k += nthr) { A.force[g0 + k] += acc[k]; acc[k] = 0.0; }
            }
        }
        // ---- finish the previous tile (nobody waits for this) ------------------------------------
        if (tid < nrec_prv)
            finish_record(A, meta_group(m_prv, 0).x, meta_group(m_prv, 2), fb_prv, fb, tid, own_prv, fuse);
        if (!has_next) {
            // the last tile of this CTA is finished right away
            const int4 md = meta_group(m_cur, 3), mc = meta_group(m_cur, 2);
            if (tid == 0) publish_flag(A.flag + md.z, A.epoch);
            if (tid < md.y - md.x) flag_value = ld_relaxed_u32(dep_flag(A, md, fb_cur, tid));
            wait_deps(A, md, fb_cur, tid, nthr, flag_value);
            __syncthreads();                    // ... also: the previous tile's records are done with spart
            request_partials(A, n0, mc, fb_cur, fb, tid, nthr);
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
            const int nrec = min(mc.y - mc.x, A.cap_recs);
            if (tid < nrec) {
                double own[3] = {0.0, 0.0, 0.0};
                if (fuse) { own[0] = fb.pend[3 * tid]; own[1] = fb.pend[3 * tid + 1]; own[2] = fb.pend[3 * tid + 2]; }
                finish_record(A, n0, mc, fb_cur, fb, tid, own, fuse);
            }
            break;
        }
        t = tn;
    }
}

// step_kernel: a launch over tiles [tile_begin, ntiles) of the slot-table list and, with STRUCT, over
// [struct_begin, struct_end) of the structured list.  STRUCT: CTAs [0, grid_struct) walk the structured
// tiles, the others the slot-table tiles.  Every CTA of the launch is resident (cooperative launch) and
// each walks ITS tiles in ascending order, so the lowest tile whose flag is not raised yet never waits:
// the argument that makes the single loop deadlock-free holds for any such split.
template <int MODE, bool DENSE, int THREADS, bool WPASS = false, bool STRUCT = false>
__global__ void __launch_bounds__(THREADS, THREADS >= 512 ? 1 : 2) step_kernel(const StepArgs A)
{
    if (STRUCT) {
        const int gs = A.grid_struct;
        const bool sfirst = (int)blockIdx.x < gs;
        // static split: one pass (a CTA walks the list it was started on).  Dynamic (A.queue): two shared
        // counters; a CTA drains the list it was started on, then helps with the other one.  The loop keeps
        // ONE inlined copy of either tile_loop in the kernel.
        if (A.queue) {
            if (sfirst) {
                tile_loop<MODE, DENSE, THREADS, false, STRUCT ? 1 : 0>(A, A.struct_begin, 0, A.struct_end, A.queue + 1);
                __syncthreads();
                tile_loop<MODE, DENSE, THREADS, false, 0>(A, A.tile_begin, 0, A.ntiles, A.queue);
            } else {
                tile_loop<MODE, DENSE, THREADS, false, 0>(A, A.tile_begin, 0, A.ntiles, A.queue);
                __syncthreads();
                tile_loop<MODE, DENSE, THREADS, false, STRUCT ? 1 : 0>(A, A.struct_begin, 0, A.struct_end, A.queue + 1);
            }
        } else if (sfirst) tile_loop<MODE, DENSE, THREADS, false, STRUCT ? 1 : 0>(A, A.struct_begin + (int)blockIdx.x, gs, A.struct_end, nullptr);
        else tile_loop<MODE, DENSE, THREADS, false, 0>(A, A.tile_begin + ((int)blockIdx.x - gs), (int)gridDim.x - gs, A.ntiles, nullptr);
    } else {
        tile_loop<MODE, DENSE, THREADS, WPASS, 0>(A, A.tile_begin + blockIdx.x, gridDim.x, A.ntiles, nullptr);
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
                                unsigned long long seq, int add, int *__restrict__ err, long long timeout_cycles)
{
    const PullSeg sg = segs[blockIdx.y];
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const volatile unsigned long long *f = sg.flag;
        const long long t0 = clock64();
        int good = 1;
        while (*f < seq) {
            __nanosleep(200);
            // a peer that never arrives: report, and do NOT apply whatever the mailbox holds
            if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) { report_error(err, 1); good = 0; break; }
        }
        ok = good;
    }
    __syncthreads();
    if (!ok) return;
    const int n3 = 3 * sg.n;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n3; k += gridDim.x * blockDim.x) {
        const size_t g = 3 * (size_t)sg.mapping[k / 3] + (k % 3);
        const double x = __ldcg(sg.local + k);
        v[g] = add ? v[g] + x : x;
    }
}

// The same for ALL messengers of a contribution list in one launch (the reference applies them one
// after another, psolve.c:5035-5073, so that a node shared with several neighbours receives its
// contributions in list order): one thread per (entry of the node-major CSR, component) walks the
// node's contributions in messenger order.  csr_node[i] = local node, csr_off[i..i+1) into csr_src,
// csr_src[j] = index of the double triple inside this list's receive area for this parity
// (p2p_alloc_mailbox: messenger i occupies triples [2 off_i, 2 off_i + 2 n_i), parity p its second half).
struct PullAll {
    const int32_t *csr_node, *csr_off, *csr_src;
    const double *local;                    // receive area of this list and parity
    const unsigned long long *flags;        // [nmsg][2 parities] sequence flags
    int32_t nnodes, nmsg, parity;
};
__global__ void p2p_pull_all_kernel(const PullAll pa, double *__restrict__ v, unsigned long long seq,
                                    int *__restrict__ err, long long timeout_cycles)
{
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    for (int m = threadIdx.x; m < pa.nmsg; m += blockDim.x) {
        const volatile unsigned long long *f = pa.flags + 2 * m + pa.parity;
        const long long t0 = clock64();
        while (*f < seq) {
            __nanosleep(200);
            if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) { report_error(err, 1); ok = 0; break; }
        }
    }
    __syncthreads();
    if (!ok) return;
    const int n3 = 3 * pa.nnodes;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n3; k += gridDim.x * blockDim.x) {
        const int i = k / 3, c = k - 3 * i;
        const size_t g = 3 * (size_t)pa.csr_node[i] + c;
        double s = v[g];
        for (int j = pa.csr_off[i]; j < pa.csr_off[i + 1]; j++) s += __ldcg(pa.local + 3 * (size_t)pa.csr_src[j] + c);
        v[g] = s;
    }
}

// BKT memory variables between the reference's layout ([8 E][3] per array, psolve.h:308-311) and the
// entry-chunked device layout.  which_k0 = family * 48 + which * 24; to_ref = 1: device -> ref.
__global__ void conv_convert_kernel(int E, const int32_t *__restrict__ entry_of_elem, int which_k0,
                                    double *conv, double *ref, int to_ref)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 24LL * E) return;
    const int e = (int)(k / 24), ic = (int)(k % 24);
    const size_t idx = conv_index((size_t)entry_of_elem[e], which_k0 + ic);
    if (to_ref) ref[k] = conv[idx]; else conv[idx] = ref[k];
}

__global__ void gather_nodes_kernel(int n, const int32_t *__restrict__ lnid, const double *__restrict__ v,
                                    double *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 3 * n) out[k] = v[3 * (size_t)lnid[k / 3] + (k % 3)];
}

// interpolate_station_displacements (psolve.c:6680-6795) for every station of this rank, one thread
// per station: trilinear shape functions phi_i = (1 + xi_i lx)(1 + eta_i ly)(1 + zeta_i lz) / 8 of the
// station's local coordinates, applied to tm1 (displacement), then -tm2 (velocity = (u1 - u2) / dt),
// then -tm2 + tm3 (acceleration = (u1 - 2 u2 + u3) / dt2).  Explicit round-to-nearest multiplies and
// adds in the reference's order (no FMA contraction), so that a row equals the reference's doubles
// bit for bit when the displacement field does.  row = [nst][9]: dis, vel, acc (unused = 0).
__global__ void station_kernel(int nst, const int32_t *__restrict__ nodes, const double *__restrict__ local,
                               const double *__restrict__ tm1, const double *__restrict__ tm2,
                               const double *__restrict__ tm3, int vel, int acc, double dt, double dt2,
                               double *__restrict__ row)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nst) return;
    const double lx = local[3 * s], ly = local[3 * s + 1], lz = local[3 * s + 2];
    double phi[8];
    size_t nd[8];
    double d[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double sx = (i & 1) ? 1.0 : -1.0, sy = (i & 2) ? 1.0 : -1.0, sz = (i & 4) ? 1.0 : -1.0;
        phi[i] = __dmul_rn(__dmul_rn(__dadd_rn(1.0, __dmul_rn(sx, lx)), __dadd_rn(1.0, __dmul_rn(sy, ly))),
                           __dadd_rn(1.0, __dmul_rn(sz, lz))) * 0.125;      // / 8 is exact
        nd[i] = 3 * (size_t)nodes[8 * s + i];
#pragma unroll
        for (int c = 0; c < 3; c++) d[c] = __dadd_rn(d[c], __dmul_rn(phi[i], tm1[nd[i] + c]));
    }
    double *o = row + 9 * (size_t)s;
#pragma unroll
    for (int c = 0; c < 9; c++) o[c] = c < 3 ? d[c] : 0.0;
    if (vel || acc) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int c = 0; c < 3; c++) d[c] = __dsub_rn(d[c], __dmul_rn(phi[i], tm2[nd[i] + c]));
#pragma unroll
        for (int c = 0; c < 3; c++) o[3 + c] = __ddiv_rn(d[c], dt);
    }
    if (acc) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                d[c] = __dsub_rn(d[c], __dmul_rn(phi[i], tm2[nd[i] + c]));
                d[c] = __dadd_rn(d[c], __dmul_rn(phi[i], tm3[nd[i] + c]));
            }
#pragma unroll
        for (int c = 0; c < 3; c++) o[6 + c] = __ddiv_rn(d[c], dt2);
    }
}

// Old_planes_print's interpolation (io_planes.c:168-191) for every plane point of this rank, one thread
// per point: phi_i = (1 + xi_i lx)(1 + eta_i ly)(1 + zeta_i lz) / 8 evaluated left to right, the three
// sums over the element's 8 nodes of tm1 in the reference's order, explicitly rounded multiplies and adds
// (no FMA contraction): a row equals the reference's doubles bit for bit when the field does.
// row = [npoints][3], the layout of the reference's strip buffers.
__global__ void plane_kernel(long long npoints, const int32_t *__restrict__ nodes, const double *__restrict__ local,
                             const double *__restrict__ tm1, double *__restrict__ row)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npoints) return;
    const double lx = local[3 * p], ly = local[3 * p + 1], lz = local[3 * p + 2];
    double d[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double sx = (i & 1) ? 1.0 : -1.0, sy = (i & 2) ? 1.0 : -1.0, sz = (i & 4) ? 1.0 : -1.0;
        const double phi = __dmul_rn(__dmul_rn(__dadd_rn(1.0, __dmul_rn(sx, lx)), __dadd_rn(1.0, __dmul_rn(sy, ly))),
                                     __dadd_rn(1.0, __dmul_rn(sz, lz))) * 0.125;      // / 8 is exact
        const size_t nd = 3 * (size_t)nodes[8 * p + i];
#pragma unroll
        for (int c = 0; c < 3; c++) d[c] = __dadd_rn(d[c], __dmul_rn(phi, tm1[nd + c]));
    }
#pragma unroll
    for (int c = 0; c < 3; c++) row[3 * p + c] = d[c];
}

}  // namespace hgpu
