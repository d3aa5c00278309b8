// hgpu_kernels.cuh -- sm_100a device code of libhercules_gpu.so.
//
// Kernels (DESIGN.md sections 3-4):
//   tile_kernel        per-element internal force (stiffness + Rayleigh damping) gathered per
//                      owned node inside an owner-computes tile, optionally fused with the
//                      central-difference update of the tile's REGULAR nodes
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

struct TileArgs {
    const double *__restrict__ u1;      // tm1  [N][3]
    const double *__restrict__ u2;      // tm2  [N][3]
    double *__restrict__ unext;         // u(t+dt) target (fused update) [N][3]
    double *__restrict__ force;         // [N][3]
    const double *__restrict__ mass;    // n_t.mass_simple   [N]
    const double *__restrict__ m2;      // n_t.mass2_minusaM [N][3]
    const double *__restrict__ m1;      // n_t.mass_minusaM  [N][3]
    const uint8_t *__restrict__ ncls;   // node class [N]
    const double *__restrict__ etab;    // e_t [E][4]
    const double *__restrict__ Kd;      // dense K1|K2 as [2][24][24] (conventional only)
    const int32_t *__restrict__ elem_off;
    const int32_t *__restrict__ elem_id;
    const uint4 *__restrict__ elem_slot;   // 8 x uint16 per entry
    const int32_t *__restrict__ halo_off;
    const int32_t *__restrict__ halo_id;
    int32_t N, tile_nodes, ntiles, tile_begin;
    int32_t stage_nodes;                // smem capacity in node slots
    double s_u1;                        // 1: include -K(c1,c2) u1          (stiffness term)
    double s_du;                        // 1: include -K(c3,c4) (u1 - u2)   (Rayleigh term)
    int32_t fuse_update;                // 1: advance REGULAR owned nodes here
};

// One CTA per tile.  Shared memory: su1[3*S] | su2[3*S] (only when NEED_U2) | acc[3*tile_nodes].
template <bool NEED_U2, bool DENSE>
__global__ void __launch_bounds__(256, 2) tile_kernel(const TileArgs A)
{
    extern __shared__ double smem[];
    const int t = A.tile_begin + blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n0 = t * A.tile_nodes;
    const int nown = min(A.tile_nodes, A.N - n0);
    const int hb = A.halo_off[t], nh = A.halo_off[t + 1] - hb;
    double *su1 = smem;
    double *su2 = smem + 3 * A.stage_nodes;
    double *acc = smem + (NEED_U2 ? 6 : 3) * A.stage_nodes;

    // ---- phase 0: stage the tile's displacements --------------------------------------------
    {
        // owned range: contiguous, 16-byte aligned (tile_nodes is even) -> 128-bit loads
        const int nd = 3 * nown, nv = nd >> 1;
        const double2 *g1 = reinterpret_cast<const double2 *>(A.u1 + 3 * (size_t)n0);
        const double2 *g2 = reinterpret_cast<const double2 *>(A.u2 + 3 * (size_t)n0);
        for (int i = tid; i < nv; i += nthr) {
            double2 v = __ldg(g1 + i);
            su1[2 * i] = v.x; su1[2 * i + 1] = v.y;
            if (NEED_U2) { double2 w = __ldg(g2 + i); su2[2 * i] = w.x; su2[2 * i + 1] = w.y; }
            acc[2 * i] = 0.0; acc[2 * i + 1] = 0.0;
        }
        if ((nd & 1) && tid == 0) {
            su1[nd - 1] = A.u1[3 * (size_t)n0 + nd - 1];
            if (NEED_U2) su2[nd - 1] = A.u2[3 * (size_t)n0 + nd - 1];
            acc[nd - 1] = 0.0;
        }
        // gathered nodes: 3 consecutive lanes read the 24 contiguous bytes of one node
        for (int i = tid; i < 3 * nh; i += nthr) {
            const int h = i / 3, c = i - 3 * h;
            const size_t g = 3 * (size_t)A.halo_id[hb + h] + c;
            su1[3 * nown + i] = __ldg(A.u1 + g);
            if (NEED_U2) su2[3 * nown + i] = __ldg(A.u2 + g);
        }
    }
    __syncthreads();

    // ---- phase 1: element forces, accumulated per owned node ------------------------------
    const int eb = A.elem_off[t], ne = A.elem_off[t + 1] - eb;
    for (int base = 0; base < ne; base += nthr) {
        const int le = base + tid;
        const bool act = le < ne;
        double fx[8], fy[8], fz[8];
        uint32_t sl[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { sl[j] = 0xffffu; fx[j] = fy[j] = fz[j] = 0.0; }
        if (act) {
            const uint4 s4 = __ldg(A.elem_slot + eb + le);
            sl[0] = s4.x & 0xffffu; sl[1] = s4.x >> 16; sl[2] = s4.y & 0xffffu; sl[3] = s4.y >> 16;
            sl[4] = s4.z & 0xffffu; sl[5] = s4.z >> 16; sl[6] = s4.w & 0xffffu; sl[7] = s4.w >> 16;
            const int e = __ldg(A.elem_id + eb + le);
            const double2 c12 = __ldg(reinterpret_cast<const double2 *>(A.etab + 4 * (size_t)e));
            double beta = 0.0;
            if (NEED_U2) {
                const double2 c34 = __ldg(reinterpret_cast<const double2 *>(A.etab + 4 * (size_t)e) + 1);
                // c3/c1 == c4/c2 == b/dt by construction (psolve.c:3387-3409)
                beta = (c12.x != 0.0) ? A.s_du * (c34.x / c12.x) : 0.0;
            }
            // w = s_u1 * u1 + beta * (u1 - u2), per corner and component
            double wx[8], wy[8], wz[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int o = 3 * sl[j];
                const double ax = su1[o], ay = su1[o + 1], az = su1[o + 2];
                if (NEED_U2) {
                    wx[j] = fma(beta, ax - su2[o], A.s_u1 * ax);
                    wy[j] = fma(beta, ay - su2[o + 1], A.s_u1 * ay);
                    wz[j] = fma(beta, az - su2[o + 2], A.s_u1 * az);
                } else {
                    wx[j] = A.s_u1 * ax; wy[j] = A.s_u1 * ay; wz[j] = A.s_u1 * az;
                }
            }
            if (!DENSE) {
                double tx[8], ty[8], tz[8];
                wht_forward(wx, tx); wht_forward(wy, ty); wht_forward(wz, tz);
                const double a = -0.5625 * (c12.y + 2.0 * c12.x);
                const double c = -0.5625 * c12.y;
                const double b = -0.5625 * c12.x;
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
                        r[k] = -c12.x * s1 - c12.y * s2;
                    }
                    // i is a runtime index here; write through a switch-free select
#pragma unroll
                    for (int jj = 0; jj < 8; jj++)
                        if (jj == i) { fx[jj] = r[0]; fy[jj] = r[1]; fz[jj] = r[2]; }
                }
            }
        }
        // A node is corner j of at most one element (leaf octants do not overlap), so within
        // pass j every accumulator is touched by at most one thread: no atomics, fixed order.
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (sl[j] < (uint32_t)nown) {
                const int o = 3 * sl[j];
                acc[o] += fx[j]; acc[o + 1] += fy[j]; acc[o + 2] += fz[j];
            }
            __syncthreads();
        }
    }

    // ---- phase 2: per owned node component: fused update, or hand the force on --------------
    for (int k = tid; k < 3 * nown; k += nthr) {
        const int i = k / 3;
        const size_t g = 3 * (size_t)n0 + k;
        const double F = acc[k];
        if (A.fuse_update && A.ncls[n0 + i] == NODE_REGULAR) {
            const double p = NEED_U2 ? su2[k] : __ldg(A.u2 + g);
            const double nf = F + (__ldg(A.m2 + g) * su1[k] - __ldg(A.m1 + g) * p);
            A.unext[g] = nf / __ldg(A.mass + n0 + i);
        } else {
            A.force[g] += F;
        }
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

__global__ void gather_nodes_kernel(int n, const int32_t *__restrict__ lnid, const double *__restrict__ v,
                                    double *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 3 * n) out[k] = v[3 * (size_t)lnid[k / 3] + (k % 3)];
}

}  // namespace hgpu
