/*
 * psolve_gpu.c -- the reference's own psolve with its time loop executed by libhercules_gpu.so.
 *
 * This is the binding INTEGRATION.md describes, as a working program.  psolve.c keeps the state
 * the hot path needs in file-static structs (`Param`, `Global`, quake/forward/psolve.c:193-337),
 * so the binding has to live in psolve.c's translation unit: this file compiles psolve.c WHERE
 * IT LIES (#include REF_PSOLVE_C, set by integration/Makefile; no reference text is copied) and
 * interposes on two external functions psolve.c calls, exactly as oracle/ref_dump.c does:
 *
 *   Timer_Start("Solver")   psolve.c:7514, immediately before solver_run(): parameters, octor
 *                           mesh and partition, solver_init, source_init, stiffness_init and
 *                           the station/plane set-up are done.  The hook runs the whole time
 *                           loop on the GPU (gpu_solver_run below, a transcription of the loop
 *                           body of solver_run, psolve.c:4265-4319, with every hot-path wrapper
 *                           replaced by its hgpu_* call) and then sets Param.theTotalSteps to
 *                           the starting step so that the reference's own loop has nothing left
 *                           to do.
 *   Timer_Stop("Solver")    psolve.c:7516: restores Param.theTotalSteps for print_timing_stat.
 *
 * Everything the reference does outside the loop -- and, inside it, the checkpoint, status,
 * 4D-output, plane and station writers and read_myForces -- runs unchanged and is CALLED from
 * here (they are static functions of the same translation unit): the station files of a
 * psolve_gpu run are produced by the reference's interpolate_station_displacements from
 * displacements fetched off the GPU.
 *
 * Multi-rank: one process per GPU; the four schedule_senddata calls per step become the
 * library's peer-memory exchange; the mailbox descriptors are all-gathered over comm_solver.
 */
#include <stdio.h>
#include <stdint.h>

#define Timer_Start(name) hgpu_hook_Timer_Start(name)
#define Timer_Stop(name)  hgpu_hook_Timer_Stop(name)

#include REF_PSOLVE_C

#undef Timer_Start
#undef Timer_Stop

void Timer_Start(char *name);
void Timer_Stop(char *name);

#include "hercules_gpu.h"

int64_t hgpu_planes_node_list(int theNumberOfPlanes, int32_t *out);      /* io_planes_gpu.c */
int64_t hgpu_planes_point_tables(int theNumberOfPlanes, int32_t *nodes, double *local);
int hgpu_planes_print_rows(int32_t myID, int theNumberOfPlanes, const double *rows);

#include <pthread.h>

static hgpu_solver_t *theGpu;
static int32_t        theSavedTotalSteps = -1;

#define GPU(call)                                                                          \
    do {                                                                                   \
        if ((call) != 0) {                                                                 \
            fprintf(stderr, "psolve_gpu: %s failed: %s\n", #call, hgpu_last_error());      \
            MPI_Abort(MPI_COMM_WORLD, ERROR);                                              \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

/* ---- asynchronous checkpoints ------------------------------------------------------------
 * checkpoint_write (io_checkpoint.c:29-117) stops the reference for as long as two whole fields take to
 * reach the disk.  Here the fields are snapshotted on the device (hgpu_fetch_all_async), travel to two
 * page-locked shadow buffers while the time loop goes on, and a writer thread puts them into the file in the
 * reference's layout: header {group size, step, nharboredmax} written by rank 0, then for rank r at byte
 * 12 + 2 r nharboredmax sizeof(fvector_t): the tm2 array followed by the tm1 array (io_checkpoint.c:93-112),
 * files checkpoint.out0 / checkpoint.out1 in turn.  The main thread only creates the file (rank 0) and meets
 * the other ranks at a barrier, as checkpoint_write does, so that every writer finds the header in place. */
typedef struct ckpt_job {
    char      filename[256];
    off_t     offset;
    size_t    rows;
    double   *tm1, *tm2;             /* shadow buffers (page-locked) */
    pthread_t thread;
    int       running, failed;
} ckpt_job_t;
static ckpt_job_t theCkpt;
static int        theCkptNumber = 0;     /* io_checkpoint.c:38 CheckpointNumber */

static void *ckpt_writer(void *arg)
{
    ckpt_job_t *j = arg;
    if (hgpu_fetch_wait(theGpu) != 0) { j->failed = 1; return NULL; }
    FILE *fp = fopen(j->filename, "rb+");
    if (!fp) { j->failed = 1; return NULL; }
    const size_t row = 3 * sizeof(double);
    if (fseeko(fp, j->offset, SEEK_SET) != 0 || fwrite(j->tm2, row, j->rows, fp) != j->rows ||
        fseeko(fp, j->offset + (off_t)(j->rows * row), SEEK_SET) != 0 || fwrite(j->tm1, row, j->rows, fp) != j->rows)
        j->failed = 1;
    if (fclose(fp) != 0) j->failed = 1;
    return NULL;
}

static void ckpt_join(void)
{
    if (!theCkpt.running) return;
    pthread_join(theCkpt.thread, NULL);
    theCkpt.running = 0;
    if (theCkpt.failed) solver_abort("gpu_solver_run", NULL, "asynchronous checkpoint to %s failed: %s\n",
                                     theCkpt.filename, hgpu_last_error());
}

/* checkpoint_write's job for `step`, without stopping the loop */
static void ckpt_write_async(int32_t step, int nharboredmax)
{
    ckpt_join();                                            /* the previous checkpoint is on disk first */
    const size_t rows = (size_t)Global.myMesh->nharbored;
    if (!theCkpt.tm1) {
        theCkpt.tm1 = hgpu_host_alloc(sizeof(double) * 3 * (rows + 1));
        theCkpt.tm2 = hgpu_host_alloc(sizeof(double) * 3 * (rows + 1));
        if (!theCkpt.tm1 || !theCkpt.tm2) solver_abort("gpu_solver_run", NULL, "hgpu_host_alloc: %s\n", hgpu_last_error());
    }
    GPU(hgpu_fetch_all_async(theGpu, HGPU_TM1, theCkpt.tm1));
    GPU(hgpu_fetch_all_async(theGpu, HGPU_TM2, theCkpt.tm2));
    sprintf(theCkpt.filename, "%s%s%d", Param.theCheckPointingDirOut, "/checkpoint.out", theCkptNumber);
    if (Global.myID == 0) {
        FILE *fp = fopen(theCkpt.filename, "wb");
        if (!fp) solver_abort("gpu_solver_run", NULL, "cannot create %s\n", theCkpt.filename);
        int hdr[3] = {Global.theGroupSize, step, nharboredmax};
        fwrite(hdr, sizeof(int), 3, fp);
        fclose(fp);
    }
    MPI_Barrier(comm_solver);
    theCkpt.offset = (off_t)(3 * sizeof(int)) + (off_t)2 * Global.myID * nharboredmax * (off_t)sizeof(fvector_t);
    theCkpt.rows = rows;
    theCkpt.failed = 0;
    if (pthread_create(&theCkpt.thread, NULL, ckpt_writer, &theCkpt) != 0)
        solver_abort("gpu_solver_run", NULL, "pthread_create failed\n");
    theCkpt.running = 1;
    theCkptNumber = theCkptNumber ? 0 : 1;
}


/* messenger_t list (psolve.h:235-272) -> flat arrays; list order = the reference's unpack order */
static void flatten_messengers(messenger_t *first, hgpu_msglist_t *out)
{
    int32_t k = 0, t = 0, count = 0, total = 0;
    for (messenger_t *m = first; m; m = m->next) { count++; total += m->nodecount; }
    int32_t *peer = malloc(sizeof(int32_t) * (count + 1)), *nodes = malloc(sizeof(int32_t) * (count + 1));
    int32_t *map = malloc(sizeof(int32_t) * (total + 1));
    for (messenger_t *m = first; m; m = m->next) {
        peer[k] = m->procid; nodes[k++] = m->nodecount;
        for (int32_t i = 0; i < m->nodecount; i++) map[t++] = m->mapping[i];
    }
    out->count = count; out->peer = peer; out->nodes = nodes; out->mapping = map;
}

static void gpu_attach(void)
{
    mesh_t *mesh = Global.myMesh;
    mysolver_t *sv = Global.mySolver;
    hgpu_mesh_t m;
    hgpu_params_t p;
    memset(&m, 0, sizeof m);
    memset(&p, 0, sizeof p);
    int32_t *lnid = malloc(sizeof(int32_t) * 8 * (mesh->lenum + 1));
    float *edata = malloc(sizeof(float) * 14 * (mesh->lenum + 1));
    int32_t *dn = malloc(sizeof(int32_t) * 6 * (mesh->ldnnum + 1));
    for (int32_t e = 0; e < mesh->lenum; e++) {                 /* elem_t, octor.h:110-115 */
        memcpy(lnid + 8 * e, mesh->elemTable[e].lnid, 8 * sizeof(int32_t));
        memcpy(edata + 14 * e, mesh->elemTable[e].data, 14 * sizeof(float));   /* edata_t, psolve.h:95-97 */
    }
    for (int32_t d = 0; d < mesh->ldnnum; d++) {                /* dnode_t, octor.h:153-158 */
        dnode_t *q = &mesh->dnodeTable[d];
        int a = 0;
        dn[6 * d] = q->ldnid; dn[6 * d + 1] = (int32_t)q->deps;
        for (int32link_t *l = q->lanid; l && a < 4; l = l->next) dn[6 * d + 2 + a++] = l->id;
        for (; a < 4; a++) dn[6 * d + 2 + a] = -1;
    }
    m.lenum = mesh->lenum; m.nharbored = mesh->nharbored; m.ldnnum = mesh->ldnnum;
    m.elem_lnid = lnid; m.edata = edata; m.dnode = dn;
    m.eTable = (const double *)sv->eTable;                      /* e_t = 4 doubles, psolve.h:196-198 */
    m.nTable = (const double *)sv->nTable;                      /* n_t = 7 doubles, psolve.h:210-214 */
    m.K1 = (const double *)Global.theK1; m.K2 = (const double *)Global.theK2;
    flatten_messengers(sv->dn_sched->first_c, &m.dn_c); flatten_messengers(sv->dn_sched->first_s, &m.dn_s);
    flatten_messengers(sv->an_sched->first_c, &m.an_c); flatten_messengers(sv->an_sched->first_s, &m.an_s);
    p.dt = Param.theDeltaT; p.dt2 = Param.theDeltaTSquared; p.freq = Param.theFreq;
    p.damping = (int32_t)Param.theTypeOfDamping;                /* same enum values, damping.h:28 */
    p.stiffness = (int32_t)Param.theStiffness;                  /* stiffness.h:24 */
    p.print_accel = Param.printStationAccelerations == YES;
    p.rank = Global.myID; p.nranks = Global.theGroupSize;
    p.nloaded = Global.theNodesLoaded; p.loaded_lnid = Global.theNodesLoadedList;
    p.device = -1;
    p.flags = HGPU_FLAG_TIMERS;
    /* stiffness_calculation_method = conventional: the same operator as effective, applied in its factored form
     * unless the literal dense K1 / K2 products are asked for (include/hercules_gpu.h, HGPU_FLAG_DENSE_K) */
    if (getenv("PSOLVE_GPU_DENSE_K") && atoi(getenv("PSOLVE_GPU_DENSE_K"))) p.flags |= HGPU_FLAG_DENSE_K;
    GPU(hgpu_init(&theGpu, &m, &p));
    if (Global.theGroupSize > 1) {
        /* every rank's mailbox descriptor to every rank (fixed-size slots: MPI_Allgather) */
        int32_t n = 0, nmax = 0;
        GPU(hgpu_comm_p2p_export(theGpu, NULL, 0, &n));
        MPI_Allreduce(&n, &nmax, 1, MPI_INT, MPI_MAX, comm_solver);
        char *mine = calloc(1, nmax), *all = calloc((size_t)Global.theGroupSize, nmax);
        int32_t *sizes = malloc(sizeof(int32_t) * Global.theGroupSize);
        const void **blobs = malloc(sizeof(void *) * Global.theGroupSize);
        GPU(hgpu_comm_p2p_export(theGpu, mine, nmax, &n));
        MPI_Allgather(&n, 1, MPI_INT, sizes, 1, MPI_INT, comm_solver);
        MPI_Allgather(mine, nmax, MPI_CHAR, all, nmax, MPI_CHAR, comm_solver);
        for (int r = 0; r < Global.theGroupSize; r++) blobs[r] = all + (size_t)r * nmax;
        GPU(hgpu_comm_p2p_connect(theGpu, blobs, sizes));
        free(mine); free(all); free(sizes); free(blobs);
    }
    free(lnid); free(edata); free(dn);
}

/* rows of a device array at the stations' nodes -> the same rows of the host array the
 * reference's writers read (psolve.c:6680-6795) */
static void fetch_station_rows(int32_t which, fvector_t *host, const int32_t *ids, int32_t n, double *tmp)
{
    if (n == 0 || host == NULL) return;
    GPU(hgpu_fetch_nodes(theGpu, which, ids, n, tmp));
    for (int32_t i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) host[ids[i]].f[c] = tmp[3 * i + c];
}

/* Station rows interpolated on the device (hgpu_stations_*) -> the reference's "station.N" text:
 * the row layout of interpolate_station_displacements (psolve.c:6729-6786): time, displacement,
 * then velocity and acceleration when asked for. */
static void write_station_rows(const double *rows, const int32_t *steps, int32_t nrows, int vel, int acc)
{
    const int32_t nst = Param.myNumberOfStations;
    for (int32_t r = 0; r < nrows; r++)
        for (int32_t i = 0; i < nst; i++) {
            const double *q = rows + 9 * ((size_t)r * nst + i);
            FILE *fp = Param.myStations[i].fpoutputfile;
            fprintf(fp, "\n%10.6f % 8e % 8e % 8e", Param.theDeltaT * steps[r], q[0], q[1], q[2]);
            if (vel) fprintf(fp, " % 8e % 8e % 8e", q[3], q[4], q[5]);
            if (acc) fprintf(fp, " % 8e % 8e % 8e", q[6], q[7], q[8]);
        }
}

/* Device times under the reference's timer names.  The host timers around the hgpu_* calls hold launch
 * time only (the calls enqueue work); what print_timing_stat (psolve.c:6041-6266) and
 * solver_run_collect_timers (psolve.c:4186-4235) should report is the time the DEVICE spent in each
 * phase (CUDA events, hgpu_get_timers).  timers.c keeps its table in a non-static array of
 * { char name[128]; double starttime, elapsed, max, min, average; int running; enum flags } entries
 * (timers.c:10-25), so the cumulative `elapsed` of an existing timer can be replaced from here.
 * The fused launches (element forces + update of the regular nodes in one kernel) have no timer of
 * their own in the reference: they are booked under "Compute addforces e".  Phases that ran on the
 * communication stream beside the late tiles are counted in full, so the parts can add up to more
 * than "Solver". */
struct hgpu_ref_timer { char name[128]; double starttime, elapsed, max, min, average; int running; int flags; };
extern struct hgpu_ref_timer Timers[];

static void set_ref_timer(const char *name, double seconds)
{
    for (int i = 0; i < 100 && Timers[i].name[0]; i++)          /* MAXTIMERS, timers.c:8 */
        if (strcmp(Timers[i].name, name) == 0) { Timers[i].elapsed = seconds; return; }
}

static void device_times_to_reference_timers(const hgpu_timers_t *tm)
{
    const double physics = tm->addforce_s + tm->addforce_e + tm->damping + tm->fused_step + tm->new_disp;
    const double comm = tm->send_dn_force + tm->adjust_force + tm->send_an_force +
                        tm->send_an_disp + tm->adjust_disp + tm->send_dn_disp;
    set_ref_timer("Compute addforces s", tm->addforce_s);
    set_ref_timer("Compute addforces e", tm->addforce_e + tm->fused_step);
    set_ref_timer("Damping addforce", tm->damping);
    set_ref_timer("1st schedule send data (contribution)", tm->send_dn_force);
    set_ref_timer("1st compute adjust (distribution)", tm->adjust_force);
    set_ref_timer("2nd schedule send data (contribution)", tm->send_an_force);
    set_ref_timer("Compute new displacement", tm->new_disp);
    set_ref_timer("3rd schedule send data (sharing)", tm->send_an_disp);
    set_ref_timer("2nd compute adjust (assignment)", tm->adjust_disp);
    set_ref_timer("4th schadule send data (sharing)", tm->send_dn_disp);
    set_ref_timer("Compute Physics", physics);
    set_ref_timer("Communication", comm);
}

static void gpu_solver_run(void)
{
    int32_t step, startingStep;
    mysolver_t *sv = Global.mySolver;

    /* What this loop does NOT carry over from solver_run (psolve.c:4281-4307): the nonlinear state and
     * forces, gravity / geostatic fix (solver_nonlinear_state, solver_compute_force_gravity/_nonlinear,
     * solver_geostatic_fix: all behind Param.includeNonlinearAnalysis, psolve.c:3911, 4016, 4026, 4119), the
     * buildings' fixed-base displacements (solver_load_fixedbase_displacements, behind get_fixedbase_flag(),
     * psolve.c:3942-3949) and the DRM taps (solver_output_drm_nodes, solver_read_drm_displacements,
     * solver_compute_effective_drm_force: Param.drmImplement).  It also treats every element as linear,
     * where the reference walks myLinearElementsMapper (the identity map without nonlinear analysis or
     * buildings, stiffness.c:101-118).  An input deck that switches any of those on must not run here
     * and produce silently different results: abort, as the reference does on every other unsupported
     * configuration (solver_abort, util.h:128). */
    if (Param.includeNonlinearAnalysis == YES)
        solver_abort("gpu_solver_run", NULL, "nonlinear analysis (include_nonlinear_analysis = yes) is not "
                     "carried by the GPU time loop; run the CPU psolve for this deck\n");
    if (Param.includeBuildings == YES)
        solver_abort("gpu_solver_run", NULL, "buildings (include_buildings = yes: fixed-base displacements, "
                     "non-identity linear-element map) are not carried by the GPU time loop\n");
    if (Param.drmImplement == YES)
        solver_abort("gpu_solver_run", NULL, "DRM (implement_drm = yes) is not carried by the GPU time loop\n");
    if (sizeof(solver_float) != sizeof(double))
        solver_abort("gpu_solver_run", NULL, "libhercules_gpu.so takes the double-precision tables "
                     "(build without -DSINGLE_PRECISION_SOLVER, psolve.h:60-64)\n");

    gpu_attach();
    if (Param.theUseCheckPoint == 1) {                          /* psolve.c:4248-4253 */
        startingStep = checkpoint_read(Global.myID, Global.myMesh, Param.theCheckPointingDirOut,
                                       Global.theGroupSize, Global.mySolver, comm_solver);
        GPU(hgpu_store_all(theGpu, HGPU_TM1, (const double *)sv->tm1));
        GPU(hgpu_store_all(theGpu, HGPU_TM2, (const double *)sv->tm2));
        Param.theUseCheckPoint = 0;                             /* solver_run must not read it again */
    } else {
        startingStep = 0;
    }
    if (Global.myID == 0)
        monitor_print("gpu_solver_run() start (libhercules_gpu.so)\nStarting time step = %d\n\n", startingStep);

    /* the 8 interpolation nodes of every station of this rank (psolve.c:6697-6701) */
    int32_t nst = 8 * Param.myNumberOfStations;
    int32_t *st_ids = malloc(sizeof(int32_t) * (nst + 1));
    double *st_tmp = malloc(sizeof(double) * 3 * (nst + 1));
    for (int32_t s = 0; s < Param.myNumberOfStations; s++)
        for (int k = 0; k < 8; k++) st_ids[8 * s + k] = Param.myStations[s].nodestointerpolate[k];
    const int vel = (Param.printStationVelocities == YES) || (Param.printStationAccelerations == YES);
    const int acc = Param.printStationAccelerations == YES;
    /* Stations on the device (default): the interpolation runs as a kernel at the station steps and
     * the rows come back ST_RING steps at a time, so a station step no longer drains the device
     * pipeline (single-rank runs; PSOLVE_GPU_HOST_STATIONS=0 turns it on for multi-rank runs too).
     * PSOLVE_GPU_HOST_STATIONS=1 keeps the reference's interpolate_station_displacements
     * on displacements fetched off the GPU (the two produce identical files, tests/test_integration.py). */
    enum { ST_RING = 256 };
    const int host_stations = getenv("PSOLVE_GPU_HOST_STATIONS") ? atoi(getenv("PSOLVE_GPU_HOST_STATIONS"))
                                                                 : Global.theGroupSize > 1;   /* multi-rank: not yet validated */
    const int dev_stations = Param.myNumberOfStations > 0 && !host_stations;
    double *st_rows = NULL;
    int32_t *st_steps = NULL;
    if (dev_stations) {
        double *loc = malloc(sizeof(double) * 3 * Param.myNumberOfStations);
        for (int32_t s = 0; s < Param.myNumberOfStations; s++)
            for (int c = 0; c < 3; c++) loc[3 * s + c] = Param.myStations[s].localcoords.x[c];
        GPU(hgpu_stations_attach(theGpu, Param.myNumberOfStations, st_ids, loc, vel, acc, 0, ST_RING));
        free(loc);
        st_rows = hgpu_host_alloc(sizeof(double) * 9 * (size_t)Param.myNumberOfStations * ST_RING);
        st_steps = malloc(sizeof(int32_t) * ST_RING);
    }
    const int32_t saved_nstations = Param.theNumberOfStations;

    /* checkpoints: asynchronous by default (PSOLVE_GPU_ASYNC_CKPT=0: the reference's checkpoint_write on
     * fields fetched synchronously) */
    const int async_ckpt = Param.theCheckPointingRate != 0 &&
                           !(getenv("PSOLVE_GPU_ASYNC_CKPT") && atoi(getenv("PSOLVE_GPU_ASYNC_CKPT")) == 0);
    int nharboredmax = Global.myMesh->nharbored;
    if (async_ckpt) {
        int mine = Global.myMesh->nharbored;
        MPI_Allreduce(&mine, &nharboredmax, 1, MPI_INT, MPI_MAX, comm_solver);       /* io_checkpoint.c:44-48 */
    }
    const int32_t saved_ckpt_rate = Param.theCheckPointingRate;

    /* Planes: planes_print reads tm1 at the 8 nodes of every plane point (io_planes.c:151-250).
     * Their sorted, duplicate-free list is fetched on a plane step instead of the whole field
     * (PSOLVE_GPU_PLANES_FULL=1 keeps the whole-field copy). */
    int32_t npl = 0, *pl_ids = NULL;
    double *pl_tmp = NULL;
    /* Planes on the device (default on one rank; PSOLVE_GPU_DEVICE_PLANES=1/0 forces it on or off): the
     * interpolation itself runs as a kernel (hgpu_planes_*), a plane step moves 3 doubles per plane point, and
     * the strips travel and are printed as the reference does it (hgpu_planes_print_rows, io_planes_gpu.c).
     * Byte-identical files (tests/test_zz_planes_gpu.py; the multi-rank transport: tests/test_planes_host.py). */
    const int want_dev_planes = getenv("PSOLVE_GPU_DEVICE_PLANES") ? atoi(getenv("PSOLVE_GPU_DEVICE_PLANES"))
                                                                   : Global.theGroupSize == 1;
    const int dev_planes = Param.theNumberOfPlanes != 0 && Param.IO_pool_pe_count == 0 && want_dev_planes;
    double *pl_rows = NULL;
    if (Param.theNumberOfPlanes != 0 && Global.myID == 0)
        monitor_print("gpu_solver_run() planes: %s\n", dev_planes ? "interpolated on the device" : "reference planes_print on fetched rows");
    if (dev_planes) {
        const int64_t npts = hgpu_planes_point_tables(Param.theNumberOfPlanes, NULL, NULL);
        int32_t *nd = malloc(sizeof(int32_t) * 8 * (size_t)(npts + 1));
        double *lc = malloc(sizeof(double) * 3 * (size_t)(npts + 1));
        hgpu_planes_point_tables(Param.theNumberOfPlanes, nd, lc);
        GPU(hgpu_planes_attach(theGpu, npts, nd, lc));
        free(nd); free(lc);
        pl_rows = hgpu_host_alloc(sizeof(double) * 3 * (size_t)(npts + 1));
        if (!pl_rows) solver_abort("gpu_solver_run", NULL, "hgpu_host_alloc: %s\n", hgpu_last_error());
    }
    if (Param.theNumberOfPlanes != 0 && Param.IO_pool_pe_count == 0 && !dev_planes &&
        !(getenv("PSOLVE_GPU_PLANES_FULL") && atoi(getenv("PSOLVE_GPU_PLANES_FULL")))) {
        const int64_t nraw = hgpu_planes_node_list(Param.theNumberOfPlanes, NULL);
        int32_t *raw = malloc(sizeof(int32_t) * (size_t)(nraw + 1));
        unsigned char *seen = calloc((size_t)Global.myMesh->nharbored + 1, 1);
        hgpu_planes_node_list(Param.theNumberOfPlanes, raw);
        for (int64_t i = 0; i < nraw; i++) seen[raw[i]] = 1;
        for (int32_t n = 0; n < Global.myMesh->nharbored; n++) npl += seen[n];
        pl_ids = malloc(sizeof(int32_t) * (size_t)(npl + 1));
        pl_tmp = hgpu_host_alloc(sizeof(double) * 3 * (size_t)(npl + 1));
        npl = 0;
        for (int32_t n = 0; n < Global.myMesh->nharbored; n++) if (seen[n]) pl_ids[npl++] = n;
        free(raw); free(seen);
    }

    /* Source streaming: read_myForces does one fseeko + fread of this rank's row of
     * force_process.<rank> per step (psolve.c:3651-3667).  Here the rows of the next SRC_WIN steps are read
     * with ONE fread into a page-locked buffer and copied to HBM (hgpu_source_preload); each step then
     * applies its row on the device (hgpu_force_source_resident).  The window keeps the host buffer below
     * 256 MB for extended sources with many loaded nodes.  PSOLVE_GPU_HOST_SOURCE=1 keeps the per-step
     * read_myForces + hgpu_force_source pair. */
    const int host_source = getenv("PSOLVE_GPU_HOST_SOURCE") ? atoi(getenv("PSOLVE_GPU_HOST_SOURCE")) : 0;
    int32_t src_win = 0, src_have0 = 0, src_have = 0;
    double *src_rows = NULL;
    if (Global.theNodesLoaded > 0 && !host_source) {
        const size_t row = sizeof(double) * 3 * (size_t)Global.theNodesLoaded;
        size_t w = ((size_t)256 << 20) / row;
        if (w < 1) w = 1;
        if (w > (size_t)(Param.theTotalSteps - startingStep)) w = (size_t)(Param.theTotalSteps - startingStep);
        src_win = (int32_t)(w > 0 ? w : 1);
        src_rows = hgpu_host_alloc(row * (size_t)src_win);
        if (!src_rows) solver_abort("gpu_solver_run", NULL, "hgpu_host_alloc: %s\n", hgpu_last_error());
    }

    MPI_Barrier(comm_solver);
    double loop_t0 = MPI_Wtime();
    double host_io = 0, host_tap = 0, host_gpu = 0, tmark;      /* where the host spends the loop */
    /* PSOLVE_GPU_WARMUP=w: the loop clock (and the device timers) restart after w steps, so that the
     * reported rate excludes one-time costs (CUDA module load, first allocations, file opens) */
    const int warm = getenv("PSOLVE_GPU_WARMUP") ? atoi(getenv("PSOLVE_GPU_WARMUP")) : 0;
    hgpu_timers_t tm_warm;
    memset(&tm_warm, 0, sizeof tm_warm);
    int32_t timed_from = startingStep;
    for (step = startingStep; step < Param.theTotalSteps; step++) {
        if (warm > 0 && step == startingStep + warm) {
            GPU(hgpu_sync(theGpu));
            GPU(hgpu_get_timers(theGpu, &tm_warm));
            host_io = host_tap = host_gpu = 0;
            timed_from = step;
            loop_t0 = MPI_Wtime();
        }
        fvector_t *tmpvector = sv->tm2;                         /* psolve.c:4271-4273 */
        sv->tm2 = sv->tm1;
        sv->tm1 = tmpvector;
        GPU(hgpu_step_begin(theGpu, step));
        tmark = MPI_Wtime();

        /* Solver I/O (psolve.c:4275-4284): the reference's writers read host tm1 (tm2, tm3);
         * fetch whole fields only on the steps a writer needs them, station rows otherwise */
        const int ckpt = (Param.theCheckPointingRate != 0) && (step != startingStep) &&
                         ((step % Param.theCheckPointingRate) == 0);
        const int wave = DO_OUTPUT && (step % Param.theRate == 0);
        const int plane = (Param.theNumberOfPlanes != 0) && (step % Param.thePlanePrintRate == 0);
        const int stat = (Param.theNumberOfStations != 0) && (step % Param.theStationsPrintRate == 0);
        const int plane_sparse = plane && pl_ids != NULL;
        const int plane_dev = plane && dev_planes;
        const int ckpt_sync = ckpt && !async_ckpt;
        if (ckpt && async_ckpt) ckpt_write_async(step, nharboredmax);
        if (plane_dev) GPU(hgpu_planes_record(theGpu, pl_rows));           /* kernel + copy, no whole field */
        if (ckpt_sync || wave || (plane && !plane_sparse && !plane_dev)) {
            GPU(hgpu_fetch_all(theGpu, HGPU_TM1, (double *)sv->tm1));
            if (ckpt_sync || wave) GPU(hgpu_fetch_all(theGpu, HGPU_TM2, (double *)sv->tm2));
        } else if (plane_sparse) {
            fetch_station_rows(HGPU_TM1, sv->tm1, pl_ids, npl, pl_tmp);       /* rows of tm1 planes_print reads */
            if (stat && !dev_stations) fetch_station_rows(HGPU_TM1, sv->tm1, st_ids, nst, st_tmp);
        } else if (stat && !dev_stations) {
            fetch_station_rows(HGPU_TM1, sv->tm1, st_ids, nst, st_tmp);
        }
        if (stat && dev_stations) {
            if (hgpu_stations_pending(theGpu) == ST_RING) {
                int32_t nr = 0;
                GPU(hgpu_stations_drain(theGpu, st_rows, st_steps, ST_RING, &nr));
                write_station_rows(st_rows, st_steps, nr, vel, acc);
            }
            GPU(hgpu_stations_record(theGpu, step));
        } else {
            if (stat && vel) fetch_station_rows(HGPU_TM2, sv->tm2, st_ids, nst, st_tmp);
            if (stat && acc) fetch_station_rows(HGPU_TM3, sv->tm3, st_ids, nst, st_tmp);
        }
        host_tap += MPI_Wtime() - tmark; tmark = MPI_Wtime();
        Timer_Start("Solver I/O");
        if (async_ckpt) Param.theCheckPointingRate = 0;        /* written by ckpt_write_async above */
        solver_write_checkpoint(step, startingStep);
        Param.theCheckPointingRate = saved_ckpt_rate;
        solver_update_status(step, startingStep);
        solver_output_wavefield(step);
        if (plane_dev) {                                        /* solver_output_planes, psolve.c:3871-3880 */
            Timer_Start("Print Planes");
            GPU(hgpu_planes_wait(theGpu));
            hgpu_planes_print_rows(Global.myID, Param.theNumberOfPlanes, pl_rows);
            Timer_Stop("Print Planes");
        } else {
            solver_output_planes(Global.mySolver, Global.myID, step);
        }
        if (dev_stations) Param.theNumberOfStations = 0;       /* rows are written by write_station_rows */
        solver_output_stations(step);
        Param.theNumberOfStations = saved_nstations;
        if (src_rows == NULL) {
            solver_read_source_forces(step);                    /* read_myForces, psolve.c:3651 */
        } else if (step >= src_have0 + src_have) {
            /* next window of rows: same file, same offsets as read_myForces, one read */
            const int32_t nrows = (Param.theTotalSteps - step) < src_win ? (Param.theTotalSteps - step) : src_win;
            const off_t where = ((off_t)sizeof(int32_t)) + Global.theNodesLoaded * sizeof(int32_t)
                              + (off_t)Global.theNodesLoaded * step * sizeof(double) * 3;
            Timer_Start("Read My Forces");
            hu_fseeko(Global.fpsource, where, SEEK_SET);
            hu_fread(src_rows, sizeof(double), (size_t)Global.theNodesLoaded * 3 * (size_t)nrows, Global.fpsource);
            Timer_Stop("Read My Forces");
            GPU(hgpu_source_preload(theGpu, step, nrows, src_rows));
            src_have0 = step; src_have = nrows;
        }
        Timer_Stop("Solver I/O");
        host_io += MPI_Wtime() - tmark; tmark = MPI_Wtime();

        /* Compute Physics / Communication (psolve.c:4286-4316), under the reference's timer names
         * (solver_run_collect_timers reduces them, psolve.c:4186-4235).  The calls only enqueue
         * device work, so these host timers hold launch time; the device times are reported from
         * hgpu_get_timers below. */
        Timer_Start("Compute Physics");
        Timer_Start("Compute addforces s");
        if (Global.theNodesLoaded > 0) {
            if (src_rows) GPU(hgpu_force_source_resident(theGpu, step));
            else GPU(hgpu_force_source(theGpu, (const double *)Global.myForces));
        }
        Timer_Stop("Compute addforces s");
        Timer_Start("Compute addforces e");
        GPU(hgpu_force_stiffness(theGpu));
        Timer_Stop("Compute addforces e");
        Timer_Start("Damping addforce");
        GPU(hgpu_force_damping(theGpu));
        Timer_Stop("Damping addforce");
        Timer_Stop("Compute Physics");
        Timer_Start("Communication");
        Timer_Start("1st schedule send data (contribution)");
        GPU(hgpu_force_exchange(theGpu));                       /* phases 8-10 in one call */
        Timer_Stop("1st schedule send data (contribution)");
        Timer_Start("1st compute adjust (distribution)"); Timer_Stop("1st compute adjust (distribution)");
        Timer_Start("2nd schedule send data (contribution)"); Timer_Stop("2nd schedule send data (contribution)");
        Timer_Stop("Communication");
        Timer_Start("Compute Physics");
        Timer_Start("Compute new displacement");
        GPU(hgpu_update(theGpu));
        Timer_Stop("Compute new displacement");
        Timer_Stop("Compute Physics");
        Timer_Start("Communication");
        Timer_Start("3rd schedule send data (sharing)");
        GPU(hgpu_disp_exchange(theGpu));                        /* phases 13-15 in one call */
        Timer_Stop("3rd schedule send data (sharing)");
        Timer_Start("2nd compute adjust (assignment)"); Timer_Stop("2nd compute adjust (assignment)");
        Timer_Start("4th schadule send data (sharing)"); Timer_Stop("4th schadule send data (sharing)");
        Timer_Stop("Communication");
        host_gpu += MPI_Wtime() - tmark;
    }
    Timer_Start("Compute Physics");
    GPU(hgpu_sync(theGpu));                                     /* the device finishes the last steps */
    Timer_Stop("Compute Physics");
    if (dev_stations) {
        int32_t nr = 0;
        GPU(hgpu_stations_drain(theGpu, st_rows, st_steps, ST_RING, &nr));
        write_station_rows(st_rows, st_steps, nr, vel, acc);
    }
    ckpt_join();
    const double loop_wall = MPI_Wtime() - loop_t0;
    /* leave the host arrays as the reference's loop would: tm1 = u(t_last), tm2 = u(t_last + dt) */
    GPU(hgpu_fetch_all(theGpu, HGPU_TM1, (double *)sv->tm1));
    GPU(hgpu_fetch_all(theGpu, HGPU_TM2, (double *)sv->tm2));
    {
        hgpu_timers_t tm;
        GPU(hgpu_get_timers(theGpu, &tm));
        if (Global.myID == 0) {
            monitor_print("gpu_solver_run() host time: taps %.6f s, reference I/O block %.6f s, hgpu calls %.6f s\n",
                          host_tap, host_io, host_gpu);
            monitor_print("gpu_solver_run() done: %lld steps, %lld kernel launches, loop wall %.6f s; device time: "
                          "step kernels %.6f s, new displacement %.6f s, adjust %.6f s, exchanges %.6f s\n",
                          (long long)(Param.theTotalSteps - timed_from), (long long)(tm.launches - tm_warm.launches), loop_wall,
                          (tm.fused_step + tm.addforce_e + tm.damping) - (tm_warm.fused_step + tm_warm.addforce_e + tm_warm.damping),
                          tm.new_disp - tm_warm.new_disp,
                          (tm.adjust_force + tm.adjust_disp) - (tm_warm.adjust_force + tm_warm.adjust_disp),
                          (tm.send_dn_force + tm.send_an_force + tm.send_an_disp + tm.send_dn_disp) -
                          (tm_warm.send_dn_force + tm_warm.send_an_force + tm_warm.send_an_disp + tm_warm.send_dn_disp));
        }
        device_times_to_reference_timers(&tm);
    }
    GPU(hgpu_finalize(theGpu));
    theGpu = NULL;
    free(st_ids); free(st_tmp); free(st_steps);
    hgpu_host_free(st_rows);
    hgpu_host_free(src_rows);
    free(pl_ids); hgpu_host_free(pl_tmp); hgpu_host_free(pl_rows);
    hgpu_host_free(theCkpt.tm1); hgpu_host_free(theCkpt.tm2);
    theCkpt.tm1 = theCkpt.tm2 = NULL;
}

void hgpu_hook_Timer_Start(char *name)
{
    Timer_Start(name);
    if (strcmp(name, "Solver") == 0 && theSavedTotalSteps < 0) {
        gpu_solver_run();
        /* nothing left for the reference's loop (psolve.c:4265): it starts at step 0 */
        theSavedTotalSteps = Param.theTotalSteps;
        Param.theTotalSteps = 0;
    }
}

void hgpu_hook_Timer_Stop(char *name)
{
    if (strcmp(name, "Solver") == 0 && theSavedTotalSteps >= 0) {
        Param.theTotalSteps = theSavedTotalSteps;
        theSavedTotalSteps = -2;
    }
    Timer_Stop(name);
}
