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
// step_kernel: persistent, software-pipelined owner-computes kernel.
//
// One CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  While tile i is being computed out
// of shared-memory stage i&1, the displacements of tile i+1 stream into the other stage with
// cp.async (LDGSTS): 16-byte copies for the owned node range (contiguous, even start), 8-byte
// copies for the gathered halo nodes.  The per-entry element data (8 slot offsets + c1, c2, beta)
// is prefetched into registers one round ahead.  Two CTAs per SM hide each other's barriers.
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
    const int4 *__restrict__ tile_meta; // per tile, in processing order: {n0, n1, hb, h1}, {eb, e1, -, -}
    const int32_t *__restrict__ halo_id;
    const uint4 *__restrict__ ent_slot; // per entry 8 x uint16: 3 * tile-local slot of each corner
    const double *__restrict__ ent_coef;// per entry c1, c2, beta
    int32_t tile_begin, ntiles;         // this launch processes tile_meta[tile_begin .. ntiles)
    int32_t cap_slots;                  // staged nodes per stage
    int32_t cap_owned;                  // accumulator nodes
    int32_t fuse_update;                // 1: advance owned nodes flagged in nt3 here
};

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc)
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

// Raw offsets as loaded (differences are taken where they are used, so that nothing consumes a
// freshly requested value early).
struct TileMeta {
    int n0, n1, hb, h1, eb, e1;
    __device__ __forceinline__ int nown() const { return n1 - n0; }
    __device__ __forceinline__ int nh() const { return h1 - hb; }
    __device__ __forceinline__ int ne() const { return e1 - eb; }
};

// Tile offsets travel through a 4-deep shared-memory ring filled by cp.async (no registers, no
// scoreboard): slot i & 3 holds the offsets of the CTA's i-th tile.
__device__ __forceinline__ void fetch_meta_async(const StepArgs &A, int t, int *slot, int tid)
{
    if (tid < 2) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(slot + 4 * tid);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(A.tile_meta + 2 * (size_t)t + tid) : "memory");
    }
}
__device__ __forceinline__ TileMeta read_meta(const int *slot)
{
    const volatile int *v = slot;
    TileMeta m;
    m.n0 = v[0]; m.n1 = v[1]; m.hb = v[2]; m.h1 = v[3]; m.eb = v[4]; m.e1 = v[5];
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

// solver_compute_displacement (psolve.c:4078-4108) for one owned node whose force is complete in
// acc: acc <- u(t+dt).  rm <= 0 flags a node that is advanced later from the force array.
__device__ __forceinline__ void advance_node(const StepArgs &A, double *acc, const double *su1,
                                             const double *su2, size_t g0, int i, double rm, double m2,
                                             double m1)
{
    const int k = 3 * i;
    if (rm > 0.0) {
#pragma unroll
        for (int c = 0; c < 3; c++)
            acc[k + c] = (acc[k + c] + (m2 * su1[k + c] - m1 * su2[k + c])) * rm;
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) A.force[g0 + k + c] += acc[k + c];
    }
}

template <int MODE, bool DENSE, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) step_kernel(const StepArgs A)
{
    constexpr bool U2E = MODE != 0;          // elements read u2
    extern __shared__ double smem[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int S3 = 3 * A.cap_slots, O3 = 3 * A.cap_owned;
    const int stage_doubles = S3 + (U2E ? S3 : O3);
    double *acc = smem + 2 * stage_doubles;
    const bool fuse = A.fuse_update != 0;

    // Pipeline.  All global loads of a warp share one hardware scoreboard, so a consumer waits
    // for EVERY load issued before it, however young.  Each register prefetch is therefore
    // consumed right BEFORE the next batch of loads is issued, and everything that can avoid
    // registers does:
    //   displacements : cp.async into the other stage, one tile ahead
    //   tile offsets  : cp.async into a 4-slot ring, three tiles ahead
    //   entries       : registers; enext -> ecur at the top of a round, then the following round's
    //                   entry is requested
    //   halo ids      : registers; requested before the accumulation passes of a tile's last round
    //                   for the tile that is staged at the top of the next iteration
    //   node tables   : registers; requested before the accumulation passes of the last round
    __shared__ __align__(16) int smeta[4][8];
    const int G = gridDim.x;
    int t = A.tile_begin + blockIdx.x;
    if (t >= A.ntiles) return;
    for (int k = tid; k < O3; k += nthr) acc[k] = 0.0;
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
        const int *m_cur = smeta[it & 3], *m_nxt = smeta[(it + 1) & 3], *m_nn = smeta[(it + 2) & 3];
        cp_async_wait_all();
        __syncthreads();                      // tile `it` has landed; everyone is done with tile it-1
        if (has_next) {
            double *n1 = smem + ((it + 1) & 1) * stage_doubles;
            const TileMeta nxt = read_meta(m_nxt);
            if (U2E || fuse) stage_tile<true, U2E>(A, nxt, n1, n1 + S3, tid, nthr, hid);
            else             stage_tile<false, false>(A, nxt, n1, n1 + S3, tid, nthr, hid);
            if (tn + 2 * G < A.ntiles) fetch_meta_async(A, tn + 2 * G, smeta[(it + 3) & 3], tid);
            cp_async_commit();
        }
        const TileMeta cur = read_meta(m_cur);
        const int nown3 = 3 * cur.nown();
        const int nxt_eb = has_next ? ((const volatile int *)m_nxt)[4] : 0;
        const int nxt_ne = has_next ? ((const volatile int *)m_nxt)[5] - nxt_eb : 0;

        // node tables of the (up to NT_PRE) owned nodes this thread advances, prefetched into
        // registers before the last round's accumulation passes
        double ntv[NT_PRE][3];
        bool nt_loaded = false;

        // ---- element forces, accumulated per owned node ------------------------------------
        for (int base = 0; base < cur.ne(); base += nthr) {
            const bool act = base + tid < cur.ne();
            ecur = enext;
            // next round's entry (or the first round of the next tile) rides along with the math
            if (base + nthr < cur.ne()) {
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
            if (base + nthr >= cur.ne()) {
                if (fuse) { load_node_tables(A, cur, tid, nthr, ntv); nt_loaded = true; }
                if (has_nn) load_halo_ids(A, read_meta(m_nn), tid, nthr, hid);
            }
            // A node is corner j of at most one element (leaf octants do not overlap), so within
            // pass j every accumulator is touched by at most one thread: no atomics, fixed order.
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (act && sl[j] < (uint32_t)nown3) {
                    const int o = sl[j];
                    acc[o] += fx[j]; acc[o + 1] += fy[j]; acc[o + 2] += fz[j];
                }
                __syncthreads();
            }
        }
        if (cur.ne() == 0) {                    // a tile of element-less nodes: keep the pipeline fed
            if (tid < nxt_ne) enext = load_entry<U2E>(A, nxt_eb + tid);
            if (has_nn) load_halo_ids(A, read_meta(m_nn), tid, nthr, hid);
        }

        // ---- owned nodes: fused update, or hand the force on ------------------------------------
        {
            const size_t g0 = 3 * (size_t)cur.n0;
            if (fuse) {
                if (!nt_loaded) load_node_tables(A, cur, tid, nthr, ntv);
                // node-wise: u(t+dt) replaces the force in acc; nodes flagged in nt3 keep their force
#pragma unroll
                for (int q = 0; q < NT_PRE; q++) {
                    const int i = tid + q * nthr;
                    if (i < cur.nown()) advance_node(A, acc, su1, su2, g0, i, ntv[q][0], ntv[q][1], ntv[q][2]);
                }
                for (int i = tid + NT_PRE * nthr; i < cur.nown(); i += nthr) {
                    const double *nt = A.nt3 + 3 * (size_t)(cur.n0 + i);
                    advance_node(A, acc, su1, su2, g0, i, __ldg(nt), __ldg(nt + 1), __ldg(nt + 2));
                }
                __syncthreads();
                // coalesced 128-bit copy-out (g0 is even); rows of flagged nodes carry their force,
                // which the special-node update overwrites afterwards
                double2 *dst = reinterpret_cast<double2 *>(A.unext + g0);
                for (int v = tid; v < (nown3 >> 1); v += nthr) {
                    dst[v] = make_double2(acc[2 * v], acc[2 * v + 1]);
                    acc[2 * v] = 0.0; acc[2 * v + 1] = 0.0;
                }
                if ((nown3 & 1) && tid == 0) { A.unext[g0 + nown3 - 1] = acc[nown3 - 1]; acc[nown3 - 1] = 0.0; }
            } else {
                for (int k = tid; k < nown3; k += nthr) {
                    A.force[g0 + k] += acc[k];
                    acc[k] = 0.0;
                }
            }
        }
        if (!has_next) break;
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
