/*
 * hercules_oracle.c -- CPU restatement of the Hercules explicit time-stepping hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library (as the checker, never as the product).  The
 * product path (hercules_b200/, include/hercules_gpu.h) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks every function below against the
 * unmodified reference compiled from /root/reference (oracle/_ref: libref_kernels.so for the
 * per-element functions on random inputs, ref_dump for whole runs on octor meshes with hanging
 * nodes) and tests/golden/ holds the committed vectors those runs produced.
 *
 * Each function cites the reference lines it follows.  Arithmetic keeps the reference's
 * operation order and operand types (float vs double) so results agree to the last bit when
 * compiled with -O2 -ffp-contract=off, like the parity build of the reference.
 *
 * Layouts are the reference's: fvector_t = double[3] (psolve.h:102-104), e_t = {c1,c2,c3,c4}
 * (psolve.h:196-198), n_t = {mass_simple, mass2_minusaM[3], mass_minusaM[3]} (psolve.h:210-214),
 * edata_t = 14 floats (psolve.h:95-97), fmatrix_t = double[3][3] (psolve.h:221-223).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HO_UNDERFLOW_CAP 1e-20 /* quake_util.c:36 */

/* corner sign table, x fastest (psolve.c:5451-5453; octor.c:6449-6470) */
static const double XI[3][8] = {{-1, 1, -1, 1, -1, 1, -1, 1},
                                {-1, -1, 1, 1, -1, -1, 1, 1},
                                {-1, -1, -1, -1, 1, 1, 1, 1}};

/* Row k of the 8x8 sign matrix used by aTransposeU/au (stiffness.c:260-288, 388-413):
 * k = 0:1  1:z  2:y  3:x  4:yz  5:xz  6:xy  7:xyz, evaluated at corner j. */
static double mode_sign(int k, int j)
{
    double x = XI[0][j], y = XI[1][j], z = XI[2][j];
    switch (k) {
    case 0: return 1;
    case 1: return z;
    case 2: return y;
    case 3: return x;
    case 4: return y * z;
    case 5: return x * z;
    case 6: return x * y;
    default: return x * y * z;
    }
}

/* quake_util.c:49-68 : 1 if any of the 24 components exceeds the cap */
static int any_nonzero24(const double v[8][3])
{
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 3; j++)
            if (fabs(v[i][j]) > HO_UNDERFLOW_CAP) return 1;
    return 0;
}

/* quake_util.c:79-96 */
static int any_nonzero3(const double v[3])
{
    for (int i = 0; i < 3; i++)
        if (fabs(v[i]) > HO_UNDERFLOW_CAP) return 1;
    return 0;
}

/* stiffness.c:245-289 (aTransposeU + reformU): t[c][k] = sum_j S[k][j] * un[j][c], left to right;
 * the k = 0 entries are forced to 0. */
static void modes_from_nodes(const double un[8][3], double t[3][8])
{
    for (int c = 0; c < 3; c++) {
        t[c][0] = 0;
        for (int k = 1; k < 8; k++) {
            double s = mode_sign(k, 0) * un[0][c];
            for (int j = 1; j < 8; j++) s += mode_sign(k, j) * un[j][c];
            t[c][k] = s;
        }
    }
}

/* stiffness.c:381-424 (au + reformF): res[j][c] += sum_k S[k][j] * v[c][k], left to right */
static void nodes_from_modes_add(double res[8][3], const double v[3][8])
{
    for (int j = 0; j < 8; j++)
        for (int c = 0; c < 3; c++) {
            double s = v[c][0];
            for (int k = 1; k < 8; k++) s += mode_sign(k, j) * v[c][k];
            res[j][c] += s;
        }
}

/* stiffness.c:291-319 firstVector(atu, out, a, c, b) with t[c][k] = atu[8c+k] */
static void scale_modes(const double t[3][8], double f[3][8], double a, double c, double b)
{
    f[0][0] = 0;
    f[0][1] = b * (t[2][3] + t[0][1]);
    f[0][2] = b * (t[1][3] + t[0][2]);
    f[0][3] = a * t[0][3] + c * (t[1][2] + t[2][1]);
    f[0][4] = b * (t[1][5] + t[2][6] + 2. * t[0][4]) / 3.;
    f[0][5] = ((a + b) * t[0][5] + c * t[1][4]) / 3.;
    f[0][6] = ((a + b) * t[0][6] + c * t[2][4]) / 3.;
    f[0][7] = ((a + 2. * b) * t[0][7]) / 9.;

    f[1][0] = 0;
    f[1][1] = b * (t[2][2] + t[1][1]);
    f[1][2] = a * t[1][2] + c * (t[0][3] + t[2][1]);
    f[1][3] = b * (t[1][3] + t[0][2]);
    f[1][4] = ((a + b) * t[1][4] + c * t[0][5]) / 3.;
    f[1][5] = b * (t[0][4] + t[2][6] + 2. * t[1][5]) / 3.;
    f[1][6] = ((a + b) * t[1][6] + c * t[2][5]) / 3.;
    f[1][7] = (a + 2. * b) * t[1][7] / 9.;

    f[2][0] = 0;
    f[2][1] = a * t[2][1] + c * (t[0][3] + t[1][2]);
    f[2][2] = b * (t[2][2] + t[1][1]);
    f[2][3] = b * (t[2][3] + t[0][1]);
    f[2][4] = ((a + b) * t[2][4] + c * t[0][6]) / 3.;
    f[2][5] = ((a + b) * t[2][5] + c * t[1][6]) / 3.;
    f[2][6] = b * (t[0][4] + t[1][5] + 2. * t[2][6]) / 3.;
    f[2][7] = (a + 2. * b) * t[2][7] / 9.;
}

/* stiffness.c:351-379 firstVector_mu: f += shear part */
static void scale_modes_mu_add(const double t[3][8], double f[3][8], double b)
{
    f[0][1] += b * (t[2][3] + t[0][1]);
    f[0][2] += b * (t[1][3] + t[0][2]);
    f[0][3] += b * (4. * t[0][3] - 2. * (t[1][2] + t[2][1])) / 3.;
    f[0][4] += b * (t[1][5] + t[2][6] + 2. * t[0][4]) / 3.;
    f[0][5] += b * (7. * t[0][5] - 2. * t[1][4]) / 9.;
    f[0][6] += b * (7. * t[0][6] - 2. * t[2][4]) / 9.;
    f[0][7] += (10. * b * t[0][7]) / 27.;

    f[1][1] += b * (t[2][2] + t[1][1]);
    f[1][2] += b * (4. * t[1][2] - 2. * (t[0][3] + t[2][1])) / 3.;
    f[1][3] += b * (t[1][3] + t[0][2]);
    f[1][4] += b * (7. * t[1][4] - 2. * t[0][5]) / 9.;
    f[1][5] += b * (t[0][4] + t[2][6] + 2. * t[1][5]) / 3.;
    f[1][6] += b * (7. * t[1][6] - 2. * t[2][5]) / 9.;
    f[1][7] += (10. * b * t[1][7]) / 27.;

    f[2][1] += b * (4. * t[2][1] - 2. * (t[0][3] + t[1][2])) / 3.;
    f[2][2] += b * (t[2][2] + t[1][1]);
    f[2][3] += b * (t[2][3] + t[0][1]);
    f[2][4] += b * (7. * t[2][4] - 2. * t[0][6]) / 9.;
    f[2][5] += b * (7. * t[2][5] - 2. * t[1][6]) / 9.;
    f[2][6] += b * (t[0][4] + t[1][5] + 2. * t[2][6]) / 3.;
    f[2][7] += (10. * b * t[2][7]) / 27.;
}

/* stiffness.c:321-349 firstVector_kappa: f += volumetric part */
static void scale_modes_kappa_add(const double t[3][8], double f[3][8], double kappa)
{
    f[0][3] += kappa * (t[0][3] + t[1][2] + t[2][1]);
    f[0][5] += kappa * (t[0][5] + t[1][4]) / 3.;
    f[0][6] += kappa * (t[0][6] + t[2][4]) / 3.;
    f[0][7] += kappa * t[0][7] / 9.;

    f[1][2] += kappa * (t[1][2] + t[0][3] + t[2][1]);
    f[1][4] += kappa * (t[1][4] + t[0][5]) / 3.;
    f[1][6] += kappa * (t[1][6] + t[2][5]) / 3.;
    f[1][7] += kappa * t[1][7] / 9.;

    f[2][1] += kappa * (t[2][1] + t[0][3] + t[1][2]);
    f[2][4] += kappa * (t[2][4] + t[0][6]) / 3.;
    f[2][5] += kappa * (t[2][5] + t[1][6]) / 3.;
    f[2][7] += kappa * t[2][7] / 9.;
}

static void scatter_add(const int32_t *ln, const double lf[8][3], double *force)
{
    for (int i = 0; i < 8; i++)
        for (int c = 0; c < 3; c++) force[3 * (size_t)ln[i] + c] += lf[i][c];
}

/* ---- a2: compute_addforce_effective, stiffness.c:180-237 --------------------------------- */
void ho_addforce_effective(int32_t E, const int32_t *lnid, const double *eTable,
                           const double *tm1, double *force)
{
    for (int32_t e = 0; e < E; e++) {
        const int32_t *ln = lnid + 8 * (size_t)e;
        const double *ep = eTable + 4 * (size_t)e;
        double lf[8][3], u[8][3];
        memset(lf, 0, sizeof lf);
        for (int i = 0; i < 8; i++)
            for (int c = 0; c < 3; c++) u[i][c] = tm1[3 * (size_t)ln[i] + c];
        if (any_nonzero24(u)) {
            double a = -0.5625 * (ep[1] + 2 * ep[0]);
            double cc = -0.5625 * (ep[1]);
            double b = -0.5625 * (ep[0]);
            double t[3][8], f[3][8];
            modes_from_nodes(u, t);
            scale_modes(t, f, a, cc, b);
            nodes_from_modes_add(lf, f);
        }
        scatter_add(ln, lf, force);
    }
}

/* quake_util.c:107-122 MultAddMatVec: V2 += c * (M * V1) */
static void mult_add_mat_vec(const double *M, const double *v1, double c, double *v2)
{
    double tmp[3] = {0, 0, 0};
    for (int r = 0; r < 3; r++)
        for (int col = 0; col < 3; col++) tmp[r] += M[3 * r + col] * v1[col];
    for (int r = 0; r < 3; r++) v2[r] += c * tmp[r];
}

/* ---- a3: compute_addforce_conventional, stiffness.c:121-174.  K1,K2 = [8][8][3][3] ------- */
void ho_addforce_conventional(int32_t E, const int32_t *lnid, const double *eTable,
                              const double *K1, const double *K2, const double *tm1,
                              double *force)
{
    for (int32_t e = 0; e < E; e++) {
        const int32_t *ln = lnid + 8 * (size_t)e;
        const double *ep = eTable + 4 * (size_t)e;
        double lf[8][3];
        memset(lf, 0, sizeof lf);
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) {
                const double *d = tm1 + 3 * (size_t)ln[j];
                if (any_nonzero3(d)) {
                    mult_add_mat_vec(K1 + 9 * (8 * i + j), d, -ep[0], lf[i]);
                    mult_add_mat_vec(K2 + 9 * (8 * i + j), d, -ep[1], lf[i]);
                }
            }
        scatter_add(ln, lf, force);
    }
}

/* ---- a5: damping_addforce (Rayleigh and MASS), damping.c:29-103 -------------------------- */
void ho_damping_addforce(int32_t E, const int32_t *lnid, const double *eTable,
                         const double *K1, const double *K2, const double *tm1,
                         const double *tm2, double *force)
{
    for (int32_t e = 0; e < E; e++) {
        const int32_t *ln = lnid + 8 * (size_t)e;
        const double *ep = eTable + 4 * (size_t)e;
        double lf[8][3], du[8][3];
        memset(lf, 0, sizeof lf);
        for (int i = 0; i < 8; i++)
            for (int c = 0; c < 3; c++)
                du[i][c] = tm1[3 * (size_t)ln[i] + c] - tm2[3 * (size_t)ln[i] + c];
        if (any_nonzero24(du)) {
            for (int i = 0; i < 8; i++)
                for (int j = 0; j < 8; j++) {
                    mult_add_mat_vec(K1 + 9 * (8 * i + j), du[j], -ep[2], lf[i]);
                    mult_add_mat_vec(K2 + 9 * (8 * i + j), du[j], -ep[3], lf[i]);
                }
        }
        scatter_add(ln, lf, force);
    }
}

/* edata_t field offsets (psolve.h:95-97) */
enum { ED_EDGE, ED_VP, ED_VS, ED_RHO, ED_A0S, ED_A1S, ED_BS, ED_G0S, ED_G1S,
       ED_A0K, ED_A1K, ED_BK, ED_G0K, ED_G1K, ED_N };

/* one family (shear or kappa) of calc_conv, damping.c:126-169 / 173-216 */
static void conv_family(const int32_t *ln, size_t e, float g0f, float g1f, double rmax,
                        const double *tm1, const double *tm2, double *f0, double *f1)
{
    if ((g0f != 0) && (g1f != 0)) {
        double c0 = g0f, c1 = g1f;
        double g0 = c0 * rmax, g1 = c1 * rmax;
        double coef_1 = g0 / 2.;
        double coef_2 = coef_1 * (1. - g0);
        double coef_3 = g1 / 2.;
        double coef_4 = coef_3 * (1. - g1);
        double exp0 = exp(-g0), exp1 = exp(-g1);
        for (int i = 0; i < 8; i++) {
            size_t n = 3 * (size_t)ln[i], ci = 3 * (8 * e + i);
            for (int c = 0; c < 3; c++) {
                f0[ci + c] = coef_2 * tm1[n + c] + coef_1 * tm2[n + c] + exp0 * f0[ci + c];
                f1[ci + c] = coef_4 * tm1[n + c] + coef_3 * tm2[n + c] + exp1 * f1[ci + c];
            }
        }
    }
}

/* ---- a6: calc_conv, damping.c:110-222.  conv_* = [8E][3] --------------------------------- */
void ho_calc_conv(int32_t E, const int32_t *lnid, const float *edata, const double *tm1,
                  const double *tm2, double *conv_shear_1, double *conv_shear_2,
                  double *conv_kappa_1, double *conv_kappa_2, double freq, double dt)
{
    double rmax = 2. * M_PI * freq * dt;
    for (int32_t e = 0; e < E; e++) {
        const int32_t *ln = lnid + 8 * (size_t)e;
        const float *ed = edata + ED_N * (size_t)e;
        conv_family(ln, e, ed[ED_G0S], ed[ED_G1S], rmax, tm1, tm2, conv_shear_1, conv_shear_2);
        conv_family(ln, e, ed[ED_G0K], ed[ED_G1K], rmax, tm1, tm2, conv_kappa_1, conv_kappa_2);
    }
}

/* damping vector of one family, damping.c:256-311 / 315-371 */
static void damping_vector(const int32_t *ln, size_t e, float a0f, float a1f, float bf,
                           double rmax, const double *tm1, const double *tm2,
                           const double *f0, const double *f1, double dv[8][3])
{
    double a0 = a0f, a1 = a1f, b = bf;
    double csum = a0 + a1 + b;
    if (csum != 0) {
        double coef = b / rmax;
        for (int i = 0; i < 8; i++) {
            size_t n = 3 * (size_t)ln[i], ci = 3 * (8 * e + i);
            for (int c = 0; c < 3; c++)
                dv[i][c] = coef * (tm1[n + c] - tm2[n + c])
                         - (a0 * f0[ci + c] + a1 * f1[ci + c])
                         + tm1[n + c];
        }
    } else {
        for (int i = 0; i < 8; i++)
            for (int c = 0; c < 3; c++) dv[i][c] = tm1[3 * (size_t)ln[i] + c];
    }
}

/* ---- a7: constant_Q_addforce, damping.c:228-416 ------------------------------------------ */
void ho_constant_Q_addforce(int32_t E, const int32_t *lnid, const double *eTable,
                            const float *edata, const double *tm1, const double *tm2,
                            const double *conv_shear_1, const double *conv_shear_2,
                            const double *conv_kappa_1, const double *conv_kappa_2,
                            double *force, double freq, double dt)
{
    double rmax = 2. * M_PI * freq * dt;
    for (int32_t e = 0; e < E; e++) {
        const int32_t *ln = lnid + 8 * (size_t)e;
        const double *ep = eTable + 4 * (size_t)e;
        const float *ed = edata + ED_N * (size_t)e;
        double dvs[8][3], dvk[8][3], lf[8][3], t[3][8], f[3][8];
        damping_vector(ln, e, ed[ED_A0S], ed[ED_A1S], ed[ED_BS], rmax, tm1, tm2,
                       conv_shear_1, conv_shear_2, dvs);
        damping_vector(ln, e, ed[ED_A0K], ed[ED_A1K], ed[ED_BK], rmax, tm1, tm2,
                       conv_kappa_1, conv_kappa_2, dvk);
        double kappa = -0.5625 * (ep[1] + 2. / 3. * ep[0]);
        double mu = -0.5625 * ep[0];
        memset(f, 0, sizeof f);
        memset(lf, 0, sizeof lf);
        if (any_nonzero24(dvs)) {
            modes_from_nodes(dvs, t);
            scale_modes_mu_add(t, f, mu);
        }
        if (any_nonzero24(dvk)) {
            modes_from_nodes(dvk, t);
            scale_modes_kappa_add(t, f, kappa);
        }
        nodes_from_modes_add(lf, f);
        scatter_add(ln, lf, force);
    }
}

/* ---- a8: compute_addforce_s, psolve.c:5912-5928 (assignment, not add) --------------------- */
void ho_addforce_s(int32_t n, const int32_t *loaded_lnid, const double *F, double dt2,
                   double *force)
{
    for (int32_t i = 0; i < n; i++)
        for (int c = 0; c < 3; c++)
            force[3 * (size_t)loaded_lnid[i] + c] = F[3 * (size_t)i + c] * dt2;
}

/* ---- a9: solver_compute_displacement, psolve.c:4072-4114 ---------------------------------- */
void ho_compute_displacement(int32_t N, const double *nTable, const double *tm1, double *tm2,
                             double *tm3, double *force)
{
    for (int32_t n = 0; n < N; n++) {
        const double *np = nTable + 7 * (size_t)n;
        double nf[3];
        for (int c = 0; c < 3; c++) {
            nf[c] = force[3 * (size_t)n + c];
            nf[c] += np[1 + c] * tm1[3 * (size_t)n + c] - np[4 + c] * tm2[3 * (size_t)n + c];
        }
        if (tm3)
            for (int c = 0; c < 3; c++) tm3[3 * (size_t)n + c] = tm2[3 * (size_t)n + c];
        for (int c = 0; c < 3; c++) tm2[3 * (size_t)n + c] = nf[c] / np[0];
    }
    memset(force, 0, sizeof(double) * 3 * (size_t)N);
}

/* ---- a10: compute_adjust, psolve.c:5936-6039.  dnode = [D][6] {ldnid, deps, anchor x4 (-1 pad)}
 * how: 0 = DISTRIBUTION (dangling value / deps added to each anchor), 1 = ASSIGNMENT. ---------- */
void ho_compute_adjust(int32_t D, const int32_t *dnode, double *values, int32_t items, int32_t how)
{
    for (int32_t d = 0; d < D; d++) {
        const int32_t *dn = dnode + 6 * (size_t)d;
        double *mine = values + (size_t)dn[0] * items;
        uint32_t deps = (uint32_t)dn[1];
        if (how == 0) {
            double part[7];
            for (int k = 0; k < items; k++) part[k] = mine[k] / deps;
            for (int a = 0; a < 4 && dn[2 + a] >= 0; a++) {
                double *pv = values + (size_t)dn[2 + a] * items;
                for (int k = 0; k < items; k++) pv[k] += part[k];
            }
        } else {
            memset(mine, 0, sizeof(double) * items);
            for (int a = 0; a < 4 && dn[2 + a] >= 0; a++) {
                const double *pv = values + (size_t)dn[2 + a] * items;
                for (int k = 0; k < items; k++) mine[k] += pv[k] / deps;
            }
        }
    }
}

/* ---- a12: compute_K, psolve.c:5446-5573 with INTEGRAL_1/2 (psolve.c:2574-2578) ------------ */
static double integral_1(double xki, double xkj, double xli, double xlj, double xmi, double xmj)
{
    return 4.5 * xki * xkj * (1 + xli * xlj / 3) * (1 + xmi * xmj / 3) / 8;
}
static double integral_2(double xki, double xlj, double xmi, double xmj)
{
    return 4.5 * xki * xlj * (1 + xmi * xmj / 3) / 8;
}

void ho_compute_K(double *K1, double *K2)
{
    double K3[8][8][3][3];
    memset(K3, 0, sizeof K3);
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++)
            for (int k = 0; k < 3; k++) {
                int k0 = k % 3, k1 = (k + 1) % 3, k2 = (k + 2) % 3;
                double I1 = integral_1(XI[k0][i], XI[k0][j], XI[k1][i], XI[k1][j], XI[k2][i], XI[k2][j]);
                double I2 = integral_1(XI[k1][i], XI[k1][j], XI[k2][i], XI[k2][j], XI[k0][i], XI[k0][j]);
                double I3 = integral_1(XI[k2][i], XI[k2][j], XI[k0][i], XI[k0][j], XI[k1][i], XI[k1][j]);
                K3[i][j][k][k] = I1 + I2 + I3;
            }
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            double *m0 = K1 + 9 * (8 * i + j), *m1 = K2 + 9 * (8 * i + j);
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++) {
                    if (k == l) {
                        int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
                        m0[3 * k + k] = integral_1(XI[k][i], XI[k][j], XI[k1][i], XI[k1][j], XI[k2][i], XI[k2][j]);
                        m1[3 * k + k] = integral_1(XI[k][j], XI[k][i], XI[k1][j], XI[k1][i], XI[k2][j], XI[k2][i]);
                    } else {
                        int m = 3 - (k + l);
                        m0[3 * k + l] = integral_2(XI[k][j], XI[l][i], XI[m][j], XI[m][i]);
                        m1[3 * k + l] = integral_2(XI[k][i], XI[l][j], XI[m][i], XI[m][j]);
                    }
                }
        }
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++)
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++) K1[9 * (8 * i + j) + 3 * k + l] += K3[i][j][k][l];
}

/* ---- a12: compute_setab, psolve.c:5813-5876.  damping: 0 rayleigh 1 mass 2 none 3 bkt ----- */
void ho_compute_setab(int32_t damping, double freq, double *abase, double *bbase)
{
    const double PI = 3.14159265358979323846264338327;
    if (damping == 0) {
        double w1 = 2 * PI * freq * .2, w2 = 2 * PI * freq * 1;
        double lw1 = log(w1), lw2 = log(w2);
        double sw1 = w1 * w1, sw2 = w2 * w2;
        double cw1 = w1 * w1 * w1, cw2 = w2 * w2 * w2;
        double numer = w1 * w2 *
            (-2 * sw1 * lw2 + 2 * sw1 * lw1 - 2 * w1 * w2 * lw2
             + 2 * w1 * w2 * lw1 + 3 * sw2 - 3 * sw1
             - 2 * sw2 * lw2 + 2 * sw2 * lw1);
        double denom = (cw1 - cw2 + 3 * sw2 * w1 - 3 * sw1 * w2);
        *abase = numer / denom;
        numer = 3 * (2 * w1 * w2 * lw2 - 2 * w1 * w2 * lw1 + sw1 - sw2);
        *bbase = numer / denom;
    } else if (damping == 1) {
        double w1 = 2 * PI * freq * .1, w2 = 2 * PI * freq * 8;
        double numer = 2 * w2 * w1 * log(w2 / w1);
        double denom = w2 - w1;
        *abase = 1.3 * numer / denom;
        *bbase = 0;
    } else {
        *abase = 0;
        *bbase = 0;
    }
}

/* compute_setflag, psolve.c:5629-5714: later tests override earlier ones, so the most specific
 * (corner > edge > face) match wins.  Encoded here as flag = 9*fz + 3*fy + fx with
 * f = 0 (touches the near end), 2 (touches the far end), 1 (neither); near wins over far only
 * where the reference's cascade says so (it cannot touch both unless the mesh is one element). */
static int boundary_flag(const int64_t ldb[3], const int64_t ruf[3], const int64_t p1[3],
                         const int64_t p2[3])
{
    int f[3];
    for (int a = 0; a < 3; a++) f[a] = (ruf[a] == p2[a]) ? 2 : (ldb[a] == p1[a]) ? 0 : 1;
    return 9 * f[2] + 3 * f[1] + f[0];
}

/* theIDBoundaryMatrix, psolve.c:5718-5746, regenerated from its rule: corner j gets bit a set
 * when the element touches a domain face normal to axis a and corner j lies on that face. */
static int boundary_bitmark(int flag, int j)
{
    int f[3] = {flag % 3, (flag / 3) % 3, flag / 9};
    int bits = 0;
    for (int a = 0; a < 3; a++) {
        int far_side = (j >> a) & 1;
        if ((f[a] == 0 && !far_side) || (f[a] == 2 && far_side)) bits |= 1 << a;
    }
    return bits;
}

/* compute_setboundary, psolve.c:5752-5804 (float arguments, HALFSPACE: flags < 9 -> +9) */
static void set_boundary(float size, float Vp, float Vs, float rho, int flag, double dashpot[8][3])
{
    memset(dashpot, 0, sizeof(double) * 24);
    flag = (flag < 9) ? flag + 9 : flag;
    double scale = rho * (size / 2) * (size / 2);
    for (int j = 0; j < 8; j++) {
        int bm = boundary_bitmark(flag, j);
        int nb = (bm & 1) + ((bm >> 1) & 1) + ((bm >> 2) & 1);
        if (nb == 3) {
            dashpot[j][0] = dashpot[j][1] = dashpot[j][2] = (Vp + 2 * Vs) * scale;
        } else if (nb == 2) {
            for (int c = 0; c < 3; c++) dashpot[j][c] = (Vs + ((bm & (1 << c)) ? Vp : Vs)) * scale;
        } else if (nb == 1) {
            for (int c = 0; c < 3; c++) dashpot[j][c] = ((bm & (1 << c)) ? Vp : Vs) * scale;
        }
    }
}

/* ---- a12: the per-element part of solver_init, psolve.c:3360-3473, and mu_and_lambda,
 * psolve.c:3236-3272 (which may overwrite edata Vp -- edata is therefore in/out).
 * node_ticks = [N][3] int64, level = [E] int8, domain = {near x,y,z, far x,y,z} ticks.
 * Returns -1 on a negative lambda (the reference aborts). ----------------------------------- */
int ho_solver_init_tables(int32_t E, int32_t N, const int32_t *lnid, const int8_t *level,
                          float *edata, const int64_t *node_ticks, const int64_t *domain,
                          double dt, double dt2, double abase, double bbase,
                          double thr_damping, double thr_vpvs, double *eTable, double *nTable)
{
    memset(nTable, 0, sizeof(double) * 7 * (size_t)N);
    for (int32_t e = 0; e < E; e++) {
        float *ed = edata + ED_N * (size_t)e;
        const int32_t *ln = lnid + 8 * (size_t)e;
        double *ep = eTable + 4 * (size_t)e;
        double mu, lambda;

        mu = ed[ED_RHO] * ed[ED_VS] * ed[ED_VS];
        if (ed[ED_VP] > (ed[ED_VS] * thr_vpvs)) {
            lambda = ed[ED_RHO] * ed[ED_VS] * ed[ED_VS] * thr_vpvs * thr_vpvs - 2 * mu;
        } else {
            lambda = ed[ED_RHO] * ed[ED_VP] * ed[ED_VP] - 2 * mu;
        }
        if (lambda < 0) {
            if (ed[ED_VS] < 500) ed[ED_VP] = 2.45 * ed[ED_VS];
            else if (ed[ED_VS] < 1200) ed[ED_VP] = 2 * ed[ED_VS];
            else ed[ED_VP] = 1.87 * ed[ED_VS];
            lambda = ed[ED_RHO] * ed[ED_VP] * ed[ED_VP];
        }
        if (lambda < 0) return -1;

        ep[0] = dt2 * ed[ED_EDGE] * mu / 9;
        ep[1] = dt2 * ed[ED_EDGE] * lambda / 9;

        double zeta = 10 / ed[ED_VS];
        if (zeta > thr_damping) zeta = thr_damping;
        double a = zeta * abase, b = zeta * bbase;
        ep[2] = b * dt * ed[ED_EDGE] * mu / 9;
        ep[3] = b * dt * ed[ED_EDGE] * lambda / 9;

        int64_t ldb[3], ruf[3];
        int64_t edgeticks = (int64_t)1 << (30 - level[e]); /* PIXELLEVEL = 30, octor.h:37 */
        for (int a3 = 0; a3 < 3; a3++) {
            ldb[a3] = node_ticks[3 * (size_t)ln[0] + a3];
            ruf[a3] = ldb[a3] + edgeticks;
        }
        int flag = boundary_flag(ldb, ruf, domain, domain + 3);
        double dashpot[8][3];
        if (flag != 13) set_boundary(ed[ED_EDGE], ed[ED_VP], ed[ED_VS], ed[ED_RHO], flag, dashpot);

        double mass = ed[ED_RHO] * ed[ED_EDGE] * ed[ED_EDGE] * ed[ED_EDGE];
        double M = mass / 8;
        for (int j = 0; j < 8; j++) {
            double *np = nTable + 7 * (size_t)ln[j];
            np[0] += M;
            for (int axis = 0; axis < 3; axis++) {
                np[4 + axis] -= (dt * a * M);
                np[1 + axis] -= (dt * a * M);
                if (flag != 13) {
                    np[4 + axis] -= (dt * dashpot[j][axis]);
                    np[1 + axis] -= (dt * dashpot[j][axis]);
                }
                np[4 + axis] += M;
                np[1 + axis] += (M * 2);
            }
        }
    }
    return 0;
}

/* ---- a1: one time step of solver_run on ONE rank without neighbours, psolve.c:4265-4319.
 * damping: 0 rayleigh 1 mass 2 none 3 bkt; stiffness: 0 conventional 1 effective.
 * On entry tm1 = u(t), tm2 = u(t - dt) (i.e. AFTER the swap of psolve.c:4271-4273);
 * on return tm2 = u(t + dt).  conv may be NULL unless damping == 3. ------------------------- */
void ho_step(int32_t E, int32_t N, int32_t D, const int32_t *lnid, const double *eTable,
             const double *nTable, const float *edata, const int32_t *dnode,
             const double *K1, const double *K2, int32_t damping, int32_t stiffness,
             double freq, double dt, double dt2, int32_t nloaded, const int32_t *loaded_lnid,
             const double *F, double *tm1, double *tm2, double *tm3, double *force,
             double *conv)
{
    ho_addforce_s(nloaded, loaded_lnid, F, dt2, force);
    if (damping != 3) {
        if (stiffness == 1) ho_addforce_effective(E, lnid, eTable, tm1, force);
        else ho_addforce_conventional(E, lnid, eTable, K1, K2, tm1, force);
    }
    if (damping == 0 || damping == 1) {
        ho_damping_addforce(E, lnid, eTable, K1, K2, tm1, tm2, force);
    } else if (damping == 3) {
        size_t s = 24 * (size_t)E;
        ho_calc_conv(E, lnid, edata, tm1, tm2, conv, conv + s, conv + 2 * s, conv + 3 * s, freq, dt);
        ho_constant_Q_addforce(E, lnid, eTable, edata, tm1, tm2, conv, conv + s, conv + 2 * s,
                               conv + 3 * s, force, freq, dt);
    }
    ho_compute_adjust(D, dnode, force, 3, 0);
    ho_compute_displacement(N, nTable, tm1, tm2, tm3, force);
    ho_compute_adjust(D, dnode, tm2, 3, 1);
}
