/*
 * ref_dump.c -- run the UNMODIFIED reference host code and tap its tables.
 * TEST INFRASTRUCTURE ONLY (fixtures for tests/golden/, bench cpu_baseline inputs).
 *
 * psolve.c keeps everything the hot path consumes in file-static structs (`Param`, `Global`,
 * quake/forward/psolve.c:193-337) that no other translation unit can reach.  This file
 * therefore compiles psolve.c *where it lies* as part of this translation unit
 * (#include REF_PSOLVE_C, set by oracle/Makefile; no reference text is copied), renames its
 * `main`, and interposes -- with function-like macros that also rewrite the prototypes in the
 * reference headers -- on two EXTERNAL functions psolve.c calls:
 *
 *   stiffness_init(id, mesh)   psolve.c:7505, the last call before solver_run: mesh_generate,
 *                              solver_init and source_init are done, so mesh_t, eTable,
 *                              nTable, K1/K2, the schedules, the loaded-node list and the
 *                              stations are final.  The hook dumps them, then calls the real
 *                              stiffness_init.
 *   Timer_Start(name)          name == "Solver I/O" marks the top of every time step right
 *                              after the tm1/tm2 swap (psolve.c:4271-4275): tm1 = u(t_step).
 *                              The hook dumps tm1 (and tm2) every HDUMP_EVERY steps and at the
 *                              last step; name == "Solver" resets the step counter.
 *
 * Everything else is the reference running unchanged, so the stations, planes and timing
 * output of a ref_dump run equal those of psolve_ref_O2.
 *
 * Output: <HDUMP_DIR or ".">/dump.<rank>.bin, a flat list of named arrays:
 *   char name[48]; int32 dtype (0=i32 1=i64 2=f32 3=f64 4=i8 5=u32); int32 ndim;
 *   int64 dims[4]; raw little-endian data.   Reader: oracle/refdump.py
 */
#include <stdio.h>
#include <stdint.h>

/* hook prototypes are produced by the macro-rewritten reference prototypes */
#define stiffness_init(id, mesh)  hdump_stiffness_init(id, mesh)
#define Timer_Start(name)         hdump_Timer_Start(name)
#define main                      ref_psolve_main

#include REF_PSOLVE_C

#undef main
#undef Timer_Start
#undef stiffness_init

void stiffness_init(int32_t myID, mesh_t *myMesh);
void Timer_Start(char *name);

static FILE *hd_fp;
static int   hd_step = -1;
static int   hd_every = 0;

static void hd_put(const char *name, int dtype, int ndim, int64_t d0, int64_t d1, int64_t d2,
                   const void *data)
{
    static const size_t sz[] = {4, 8, 4, 8, 1, 4};
    char nm[48];
    int32_t hdr[2] = {dtype, ndim};
    int64_t dims[4] = {d0, d1, d2, 1};
    size_t n = 1;
    memset(nm, 0, sizeof nm);
    strncpy(nm, name, sizeof nm - 1);
    for (int i = 0; i < ndim; i++) n *= (size_t)dims[i];
    fwrite(nm, 1, sizeof nm, hd_fp);
    fwrite(hdr, sizeof hdr, 1, hd_fp);
    fwrite(dims, sizeof dims, 1, hd_fp);
    if (n) fwrite(data, sz[dtype], n, hd_fp);
}

static void hd_sched(const char *prefix, messenger_t *first)
{
    /* per messenger: procid, nodecount ; then the concatenated mapping[] lists */
    int count = 0, total = 0, k = 0, t = 0;
    char nm[48];
    for (messenger_t *m = first; m; m = m->next) { count++; total += m->nodecount; }
    int32_t *hdr = malloc(sizeof(int32_t) * 2 * (count + 1));
    int32_t *map = malloc(sizeof(int32_t) * (total + 1));
    for (messenger_t *m = first; m; m = m->next) {
        hdr[2 * k] = m->procid; hdr[2 * k + 1] = m->nodecount; k++;
        for (int i = 0; i < m->nodecount; i++) map[t++] = m->mapping[i];
    }
    snprintf(nm, sizeof nm, "%s_hdr", prefix);
    hd_put(nm, 0, 2, count, 2, 1, hdr);
    snprintf(nm, sizeof nm, "%s_map", prefix);
    hd_put(nm, 0, 1, total, 1, 1, map);
    free(hdr); free(map);
}

void hdump_stiffness_init(int32_t myID, mesh_t *mesh)
{
    const char *dir = getenv("HDUMP_DIR");
    const char *ev  = getenv("HDUMP_EVERY");
    char path[512];
    mysolver_t *s = Global.mySolver;
    int32_t E = mesh->lenum, N = mesh->nharbored, D = mesh->ldnnum;

    hd_every = ev ? atoi(ev) : 0;
    snprintf(path, sizeof path, "%s/dump.%d.bin", dir ? dir : ".", (int)myID);
    hd_fp = fopen(path, "wb");
    if (!hd_fp) { perror(path); MPI_Abort(MPI_COMM_WORLD, ERROR); exit(1); }

    double par[16] = {Param.theDeltaT, Param.theDeltaTSquared, Param.theFreq,
                      (double)Param.theTypeOfDamping, (double)Param.theStiffness,
                      (double)Param.theTotalSteps, (double)Global.myID,
                      (double)Global.theGroupSize,
                      (double)(Param.printStationAccelerations == YES),
                      Global.theABase, Global.theBBase, mesh->ticksize,
                      Param.theThresholdDamping, Param.theThresholdVpVs,
                      (double)Global.theETotal, (double)Global.theNTotal};
    hd_put("params", 3, 1, 16, 1, 1, par);
    int32_t counts[4] = {E, mesh->lnnum, D, N};
    hd_put("counts", 0, 1, 4, 1, 1, counts);
    int64_t ends[6] = {Global.myOctree->nearendp[0], Global.myOctree->nearendp[1],
                       Global.myOctree->nearendp[2], Global.myOctree->farendp[0],
                       Global.myOctree->farendp[1], Global.myOctree->farendp[2]};
    hd_put("domain_ticks", 1, 1, 6, 1, 1, ends);

    /* elements (octor.h:110-115) + edata_t (psolve.h:95-97) + e_t (psolve.h:196-198) */
    int32_t *lnid  = malloc(sizeof(int32_t) * 8 * (E + 1));
    int64_t *geid  = malloc(sizeof(int64_t) * (E + 1));
    int8_t  *level = malloc(E + 1);
    float   *ed    = malloc(sizeof(float) * 14 * (E + 1));
    for (int32_t e = 0; e < E; e++) {
        elem_t *ep = &mesh->elemTable[e];
        memcpy(lnid + 8 * e, ep->lnid, sizeof(int32_t) * 8);
        geid[e] = ep->geid; level[e] = ep->level;
        memcpy(ed + 14 * e, ep->data, sizeof(float) * 14);
    }
    hd_put("elem_lnid", 0, 2, E, 8, 1, lnid);
    hd_put("elem_geid", 1, 1, E, 1, 1, geid);
    hd_put("elem_level", 4, 1, E, 1, 1, level);
    hd_put("elem_edata", 2, 2, E, 14, 1, ed);
    hd_put("eTable", 3, 2, E, 4, 1, s->eTable);
    free(lnid); free(geid); free(level); free(ed);

    /* nodes (octor.h:133-144) + n_t (psolve.h:210-214) */
    int64_t *xyz   = malloc(sizeof(int64_t) * 3 * (N + 1));
    int64_t *gnid  = malloc(sizeof(int64_t) * (N + 1));
    int8_t  *flags = malloc(2 * (N + 1));
    int32_t *owner = malloc(sizeof(int32_t) * (N + 1));
    int nshare = 0;
    for (int32_t n = 0; n < N; n++) {
        node_t *np = &mesh->nodeTable[n];
        xyz[3 * n] = np->x; xyz[3 * n + 1] = np->y; xyz[3 * n + 2] = np->z;
        gnid[n] = np->gnid; flags[2 * n] = np->ismine; flags[2 * n + 1] = np->isanchored;
        owner[n] = np->ismine ? Global.myID : np->proc.ownerid;
        if (np->ismine) for (int32link_t *l = np->proc.share; l; l = l->next) nshare++;
    }
    int32_t *share = malloc(sizeof(int32_t) * 2 * (nshare + 1));
    nshare = 0;
    for (int32_t n = 0; n < N; n++) {
        node_t *np = &mesh->nodeTable[n];
        if (np->ismine) for (int32link_t *l = np->proc.share; l; l = l->next) {
            share[2 * nshare] = n; share[2 * nshare + 1] = l->id; nshare++;
        }
    }
    hd_put("node_ticks", 1, 2, N, 3, 1, xyz);
    hd_put("node_gnid", 1, 1, N, 1, 1, gnid);
    hd_put("node_flags", 4, 2, N, 2, 1, flags);
    hd_put("node_owner", 0, 1, N, 1, 1, owner);
    hd_put("node_share", 0, 2, nshare, 2, 1, share);
    hd_put("nTable", 3, 2, N, 7, 1, s->nTable);
    free(xyz); free(gnid); free(flags); free(owner); free(share);

    /* owned dangling nodes (octor.h:153-158), anchors in list order */
    int32_t *dn = malloc(sizeof(int32_t) * 6 * (D + 1));
    for (int32_t d = 0; d < D; d++) {
        dnode_t *dp = &mesh->dnodeTable[d];
        int k = 0;
        dn[6 * d] = dp->ldnid; dn[6 * d + 1] = (int32_t)dp->deps;
        for (int i = 0; i < 4; i++) dn[6 * d + 2 + i] = -1;
        for (int32link_t *l = dp->lanid; l && k < 4; l = l->next) dn[6 * d + 2 + k++] = l->id;
    }
    hd_put("dnode", 0, 2, D, 6, 1, dn);
    free(dn);

    hd_put("K1", 3, 3, 8, 8, 9, Global.theK1);
    hd_put("K2", 3, 3, 8, 8, 9, Global.theK2);

    /* halo schedules (psolve.h:235-272) */
    hd_sched("dn_c", s->dn_sched->first_c);
    hd_sched("dn_s", s->dn_sched->first_s);
    hd_sched("an_c", s->an_sched->first_c);
    hd_sched("an_s", s->an_sched->first_s);

    /* source: loaded-node list (psolve.c:6275-6330); forces are in force_process.<rank> */
    hd_put("loaded_lnid", 0, 1, Global.theNodesLoaded > 0 ? Global.theNodesLoaded : 0, 1, 1,
           Global.theNodesLoadedList);

    /* stations on this rank (psolve.h:333-342) */
    int ns = Param.myNumberOfStations;
    int32_t *sn = malloc(sizeof(int32_t) * 9 * (ns + 1));
    double  *sc = malloc(sizeof(double) * 3 * (ns + 1));
    for (int i = 0; i < ns; i++) {
        sn[9 * i] = Param.myStations[i].id;
        memcpy(sn + 9 * i + 1, Param.myStations[i].nodestointerpolate, sizeof(int32_t) * 8);
        memcpy(sc + 3 * i, Param.myStations[i].localcoords.x, sizeof(double) * 3);
    }
    hd_put("station_nodes", 0, 2, ns, 9, 1, sn);
    hd_put("station_local", 3, 2, ns, 3, 1, sc);
    free(sn); free(sc);
    fflush(hd_fp);

    /* fixtures that only need the tables: cut the time loop to one step (the timer report needs one) */
    if (getenv("HDUMP_STOP_AFTER_INIT")) Param.theTotalSteps = 1;

    stiffness_init(myID, mesh);
}

void hdump_Timer_Start(char *name)
{
    if (hd_fp && strcmp(name, "Solver") == 0) {
        hd_step = -1;
    } else if (hd_fp && strcmp(name, "Solver I/O") == 0) {
        hd_step++;
        int last = hd_step == Param.theTotalSteps - 1;
        if ((hd_every > 0 && hd_step % hd_every == 0) || last) {
            char nm[48];
            int32_t N = Global.myMesh->nharbored;
            snprintf(nm, sizeof nm, "tm1_step%d", hd_step);
            hd_put(nm, 3, 2, N, 3, 1, Global.mySolver->tm1);
            snprintf(nm, sizeof nm, "tm2_step%d", hd_step);
            hd_put(nm, 3, 2, N, 3, 1, Global.mySolver->tm2);
            fflush(hd_fp);
        }
    }
    Timer_Start(name);
}

int main(int argc, char **argv)
{
    int rc = ref_psolve_main(argc, argv);
    if (hd_fp) fclose(hd_fp);
    return rc;
}
