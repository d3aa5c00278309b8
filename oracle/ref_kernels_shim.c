/*
 * ref_kernels_shim.c -- flat-array entry points onto the UNMODIFIED reference kernels.
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/libref_kernels.so together
 * with the reference's own stiffness.c, damping.c and quake_util.c (compiled where they lie).
 *
 * The reference functions take mesh_t / mysolver_t (octor.h:166-179, psolve.h:295-312).  Each
 * refk_* wrapper below assembles those structs around caller-owned flat arrays and calls the
 * reference function unchanged, so tests can run the real compute_addforce_effective,
 * compute_addforce_conventional, damping_addforce, calc_conv and constant_Q_addforce on random
 * inputs and compare them with oracle/hercules_oracle.c.
 *
 * Three symbols those objects import are supplied here:
 *   isThisElementNonLinear -> NO for every element, which is what nonlinear.c:82-95 returns when
 *                             nonlinear analysis is off (both Vs bounds are 0, nonlinear.c:45-46)
 *   hu_xmalloc             -> malloc-or-abort (util.c)
 *   MPI_Abort              -> abort()
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "psolve.h"
#include "stiffness.h"
#include "damping.h"
#include "quake_util.h"

noyesflag_t isThisElementNonLinear(mesh_t *myMesh, int32_t eindex)
{
    (void)myMesh; (void)eindex;
    return NO;
}

void *hu_xmalloc(size_t size, const char *varname)
{
    void *p = malloc(size ? size : 1);
    if (!p) { fprintf(stderr, "refk: out of memory for %s\n", varname ? varname : "?"); abort(); }
    return p;
}

int MPI_Abort(MPI_Comm comm, int errorcode)
{
    (void)comm; (void)errorcode;
    abort();
    return 0;
}

typedef struct {
    mesh_t mesh;
    mysolver_t solver;
} refk_ctx_t;

static void ctx_open(refk_ctx_t *c, int32_t E, int32_t N, const int32_t *lnid, const double *eTable,
                     const float *edata)
{
    memset(c, 0, sizeof *c);
    c->mesh.lenum = E;
    c->mesh.nharbored = N;
    c->mesh.elemTable = calloc((size_t)(E > 0 ? E : 1), sizeof(elem_t));
    for (int32_t e = 0; e < E; e++) {
        memcpy(c->mesh.elemTable[e].lnid, lnid + 8 * (size_t)e, 8 * sizeof(int32_t));
        c->mesh.elemTable[e].geid = e;
        c->mesh.elemTable[e].data = edata ? (void *)(edata + 14 * (size_t)e) : NULL;
    }
    c->solver.eTable = (e_t *)eTable;
}

static void ctx_close(refk_ctx_t *c) { free(c->mesh.elemTable); }

void refk_addforce_effective(int32_t E, int32_t N, const int32_t *lnid, const double *eTable,
                             const double *tm1, double *force)
{
    refk_ctx_t c;
    ctx_open(&c, E, N, lnid, eTable, NULL);
    c.solver.tm1 = (fvector_t *)tm1;
    c.solver.force = (fvector_t *)force;
    stiffness_init(0, &c.mesh);
    compute_addforce_effective(&c.mesh, &c.solver);
    ctx_close(&c);
}

void refk_addforce_conventional(int32_t E, int32_t N, const int32_t *lnid, const double *eTable,
                                const double *K1, const double *K2, const double *tm1,
                                double *force)
{
    refk_ctx_t c;
    ctx_open(&c, E, N, lnid, eTable, NULL);
    c.solver.tm1 = (fvector_t *)tm1;
    c.solver.force = (fvector_t *)force;
    stiffness_init(0, &c.mesh);
    compute_addforce_conventional(&c.mesh, &c.solver, (fmatrix_t(*)[8])K1, (fmatrix_t(*)[8])K2);
    ctx_close(&c);
}

void refk_damping_addforce(int32_t E, int32_t N, const int32_t *lnid, const double *eTable,
                           const double *K1, const double *K2, const double *tm1,
                           const double *tm2, double *force)
{
    refk_ctx_t c;
    ctx_open(&c, E, N, lnid, eTable, NULL);
    c.solver.tm1 = (fvector_t *)tm1;
    c.solver.tm2 = (fvector_t *)tm2;
    c.solver.force = (fvector_t *)force;
    damping_addforce(&c.mesh, &c.solver, (fmatrix_t(*)[8])K1, (fmatrix_t(*)[8])K2);
    ctx_close(&c);
}

void refk_calc_conv(int32_t E, int32_t N, const int32_t *lnid, const float *edata,
                    const double *tm1, const double *tm2, double *cs1, double *cs2,
                    double *ck1, double *ck2, double freq, double dt)
{
    refk_ctx_t c;
    ctx_open(&c, E, N, lnid, NULL, edata);
    c.solver.tm1 = (fvector_t *)tm1;
    c.solver.tm2 = (fvector_t *)tm2;
    c.solver.conv_shear_1 = (fvector_t *)cs1;
    c.solver.conv_shear_2 = (fvector_t *)cs2;
    c.solver.conv_kappa_1 = (fvector_t *)ck1;
    c.solver.conv_kappa_2 = (fvector_t *)ck2;
    calc_conv(&c.mesh, &c.solver, freq, dt, dt * dt);
    ctx_close(&c);
}

void refk_constant_Q_addforce(int32_t E, int32_t N, const int32_t *lnid, const double *eTable,
                              const float *edata, const double *tm1, const double *tm2,
                              const double *cs1, const double *cs2, const double *ck1,
                              const double *ck2, double *force, double freq, double dt)
{
    refk_ctx_t c;
    ctx_open(&c, E, N, lnid, eTable, edata);
    c.solver.tm1 = (fvector_t *)tm1;
    c.solver.tm2 = (fvector_t *)tm2;
    c.solver.force = (fvector_t *)force;
    c.solver.conv_shear_1 = (fvector_t *)cs1;
    c.solver.conv_shear_2 = (fvector_t *)cs2;
    c.solver.conv_kappa_1 = (fvector_t *)ck1;
    c.solver.conv_kappa_2 = (fvector_t *)ck2;
    constant_Q_addforce(&c.mesh, &c.solver, freq, dt, dt * dt);
    ctx_close(&c);
}
