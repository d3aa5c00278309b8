/*
 * hercules_gpu.h -- C ABI of libhercules_gpu.so: the B200 (sm_100a) implementation of the
 * explicit time-stepping hot path of Hercules' quake/forward solver.
 *
 * The reference has no plugin interface for this path; the seam is the family of one-line
 * static wrappers solver_run calls once per step (quake/forward/psolve.c:3953-4163, call order
 * psolve.c:4265-4319).  Each entry point below replaces the body of one of them and says which.
 * All floating point is IEEE double; ids are int32 local (psolve.h:66-71).  Array layouts are
 * the reference's own so host code can pass its tables without reshuffling:
 *
 *   fvector_t  double[3]                       psolve.h:102-104
 *   e_t        {c1,c2,c3,c4}                   psolve.h:196-198
 *   n_t        {mass_simple, mass2_minusaM[3], mass_minusaM[3]}   psolve.h:210-214
 *   edata_t    14 floats                       psolve.h:95-97
 *   fmatrix_t  double[3][3]                    psolve.h:221-223
 *
 * Pointer-linked reference structures are passed flattened (INTEGRATION.md shows the loops):
 *   elem_t.lnid[8]  (octor.h:110-115)   -> int32 [lenum][8]
 *   dnode_t + int32link_t (octor.h:121-158) -> int32 [ldnnum][6] = {ldnid, deps, anchor[4]},
 *                                             anchors in list order, unused slots = -1
 *   messenger_t lists (psolve.h:235-272) -> per schedule side: peer[], count[], concatenated
 *                                             mapping[] (ascending lnid per peer, psolve.c:4806-4860)
 *
 * Error behaviour.  The reference functions return void and abort the job on any error
 * (MPI_Abort + exit, util.h:128).  Every function here returns 0 on success or a negative
 * HGPU_E* code and records a message retrievable with hgpu_last_error(); the reference-side
 * shim is expected to call solver_abort on non-zero.  There is no CPU fallback: without a CUDA
 * device hgpu_init fails with HGPU_ENODEVICE.
 *
 * Threading: one host thread per solver handle, one handle per GPU (one MPI rank = one GPU).
 * Memory: every host array handed to hgpu_init is copied; the caller keeps ownership.
 */
#ifndef HERCULES_GPU_H
#define HERCULES_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGPU_ABI_VERSION 1

/* damping_type_t, damping.h:28 */
enum { HGPU_DAMPING_RAYLEIGH = 0, HGPU_DAMPING_MASS = 1, HGPU_DAMPING_NONE = 2, HGPU_DAMPING_BKT = 3 };
/* stiffness_type_t, stiffness.h:24 */
enum { HGPU_STIFFNESS_CONVENTIONAL = 0, HGPU_STIFFNESS_EFFECTIVE = 1 };
/* which array hgpu_fetch_* / hgpu_store_* address */
enum { HGPU_TM1 = 1, HGPU_TM2 = 2, HGPU_TM3 = 3, HGPU_FORCE = 4,
       HGPU_CONV_SHEAR_1 = 5, HGPU_CONV_SHEAR_2 = 6, HGPU_CONV_KAPPA_1 = 7, HGPU_CONV_KAPPA_2 = 8 };

enum {
    HGPU_OK = 0,
    HGPU_EINVAL = -1,     /* bad argument / inconsistent mesh          */
    HGPU_ENODEVICE = -2,  /* no usable CUDA device (never falls back)  */
    HGPU_ECUDA = -3,      /* a CUDA call failed                        */
    HGPU_ENOMEM = -4,
    HGPU_ECOMM = -5,      /* NCCL / halo exchange failure              */
    HGPU_ESTATE = -6      /* call out of sequence                      */
};

/* One side of one schedule_t (psolve.h:255-272): the messengers of a c-list or an s-list. */
typedef struct hgpu_msglist {
    int32_t        count;    /* number of messengers (neighbour ranks)              */
    const int32_t *peer;     /* [count]  messenger_t.procid                          */
    const int32_t *nodes;    /* [count]  messenger_t.nodecount                       */
    const int32_t *mapping;  /* [sum nodes] messenger_t.mapping[], peer after peer   */
} hgpu_msglist_t;

/* Snapshot of one rank's mesh_t (octor.h:166-179) and mysolver_t tables (psolve.h:295-312). */
typedef struct hgpu_mesh {
    int32_t lenum;              /* mesh_t.lenum      */
    int32_t nharbored;          /* mesh_t.nharbored  */
    int32_t ldnnum;             /* mesh_t.ldnnum     */
    const int32_t *elem_lnid;   /* [lenum][8]        */
    const double  *eTable;      /* [lenum][4]        solver_init, psolve.c:3384-3409 */
    const double  *nTable;      /* [nharbored][7]    solver_init, psolve.c:3445-3473 + mass exchange :3498-3507 */
    const float   *edata;       /* [lenum][14] or NULL; required for HGPU_DAMPING_BKT */
    const int32_t *dnode;       /* [ldnnum][6]       */
    const double  *K1;          /* [8][8][3][3] theK1 (= K1+K3, psolve.c:5558-5570); required for CONVENTIONAL */
    const double  *K2;          /* [8][8][3][3] theK2 */
    hgpu_msglist_t dn_c, dn_s;  /* mysolver_t.dn_sched: contribute-to-owner / shared-by-others */
    hgpu_msglist_t an_c, an_s;  /* mysolver_t.an_sched */
} hgpu_mesh_t;

/* The scalars of Param/Global the path reads (psolve.c:193-337). */
typedef struct hgpu_params {
    double  dt;                 /* Param.theDeltaT         */
    double  dt2;                /* Param.theDeltaTSquared  */
    double  freq;               /* Param.theFreq (BKT rmax = 2 pi f dt, damping.c:114) */
    int32_t damping;            /* Param.theTypeOfDamping  */
    int32_t stiffness;          /* Param.theStiffness      */
    int32_t print_accel;        /* Param.printStationAccelerations == YES: keep tm3 (psolve.c:4094-4101) */
    int32_t rank, nranks;       /* Global.myID, Global.theGroupSize */
    int32_t nloaded;            /* Global.theNodesLoaded   */
    const int32_t *loaded_lnid; /* Global.theNodesLoadedList [nloaded] */
    int32_t device;             /* CUDA device ordinal, -1 = rank % device count */
    int32_t tile_nodes;         /* cap on owned nodes per tile, 0 = default */
    int32_t flags;              /* HGPU_FLAG_* */
} hgpu_params_t;

#define HGPU_FLAG_NO_FUSE 1     /* keep force evaluation and update as separate kernels */
#define HGPU_FLAG_TIMERS  2     /* bracket every phase with CUDA events (hgpu_get_timers)   */
#define HGPU_FLAG_NO_OVERLAP 4  /* multi-GPU: run the force exchange after all tiles, on one stream */
#define HGPU_FLAG_TAIL_OVERLAP 8 /* multi-GPU, opt-in: also run the update of the shared nodes and the displacement
                                   exchange on the communication stream, beside the late tiles */

#define HGPU_FLAG_WPASS 16       /* opt-in step-kernel variant (Rayleigh + effective, fused): on tiles of one material
                                   the damped displacement is formed once per staged node, not per element corner */

#define HGPU_FLAG_NO_STRUCT 32   /* never use the structured-tile path of the step kernel */
#define HGPU_FLAG_STRUCT 64      /* opt-in: aligned uniform 8x8x8 cells of one material are evaluated without a slot table,
                                   from a per-node damped displacement in padded conflict-free planes, on their own CTAs
                                   (also HGPU_STRUCT=1 in the environment) */

#define HGPU_FLAG_DENSE_K 128    /* HGPU_STIFFNESS_CONVENTIONAL: evaluate the literal dense 24 x 24 products with the K1 / K2
                                   handed to hgpu_init (stiffness.c:121-176).  Without it a conventional solver applies the
                                   same operator in the factored form of compute_addforce_effective (results agree to
                                   rounding, 5e-16 rel L2 over a run; 7x faster) */

typedef struct hgpu_solver hgpu_solver_t;

/* Named per-phase device times in seconds, accumulated with CUDA events under the reference's
 * timer names (psolve.c:3955-4162; read by solver_run_collect_timers, psolve.c:4186-4235). */
typedef struct hgpu_timers {
    double addforce_s;      /* "Compute addforces s"                    */
    double addforce_e;      /* "Compute addforces e"                    */
    double damping;         /* "Damping addforce"                       */
    double send_dn_force;   /* "1st schedule send data (contribution)"  */
    double adjust_force;    /* "1st compute adjust (distribution)"      */
    double send_an_force;   /* "2nd schedule send data (contribution)"  */
    double new_disp;        /* "Compute new displacement"               */
    double send_an_disp;    /* "3rd schedule send data (sharing)"       */
    double adjust_disp;     /* "2nd compute adjust (assignment)"        */
    double send_dn_disp;    /* "4th schadule send data (sharing)"       */
    double fused_step;      /* time spent in fused force+update launches (not in the reference) */
    int64_t launches;       /* kernels launched by this library so far  */
    int64_t steps;
} hgpu_timers_t;

const char *hgpu_last_error(void);
int hgpu_abi_version(void);
/* number of CUDA devices visible, or a negative HGPU_E* code */
int hgpu_device_count(void);

/* Called once after stiffness_init (psolve.c:7505) and before solver_run (psolve.c:7515).
 * Copies the tables to HBM, builds the tile lists and the hanging-node/halo index lists.
 * tm1 = tm2 = force = 0, as after solver_init's calloc (psolve.c:3317-3325). */
int hgpu_init(hgpu_solver_t **out, const hgpu_mesh_t *mesh, const hgpu_params_t *params);

/* Multi-GPU only: join the NCCL communicator that replaces comm_solver for the four
 * schedule_senddata calls (psolve.c:4036-4163).  unique_id is the 128-byte ncclUniqueId made by
 * hgpu_comm_unique_id on rank 0 and broadcast by the host (MPI_Bcast / torch.distributed). */
int hgpu_comm_unique_id(void *unique_id_128);
int hgpu_comm_init(hgpu_solver_t *s, const void *unique_id_128);

/* Multi-GPU, peer-memory transport (one process per GPU on one node): every rank exports a blob
 * describing its mailbox (CUDA IPC handle + where each neighbour writes), the host all-gathers
 * the blobs (MPI_Allgather / torch.distributed), every rank connects.  After that the four
 * schedule_senddata calls run as device kernels that store straight into the peer's memory over
 * NVLink and wait on sequence flags; NCCL is not used.  hgpu_comm_p2p_export with blob == NULL
 * only reports the size needed. */
int hgpu_comm_p2p_export(hgpu_solver_t *s, void *blob, int32_t capacity, int32_t *size_out);
int hgpu_comm_p2p_connect(hgpu_solver_t *s, const void *const *blobs, const int32_t *sizes);

/* local_finalize / solver_delete (psolve.c:488, 3627) */
int hgpu_finalize(hgpu_solver_t *s);

/* ---- one time step, in the order of psolve.c:4265-4319 ---------------------------------- */

/* psolve.c:4271-4273: swap tm1/tm2 (device pointer rotation, no copy). */
int hgpu_step_begin(hgpu_solver_t *s, int32_t step);
/* solver_compute_force_source -> compute_addforce_s (psolve.c:3953, 5912): force[lnid] = F*dt2.
 * F = [nloaded][3] HOST doubles as read by read_myForces (psolve.c:3651). */
int hgpu_force_source(hgpu_solver_t *s, const double *F);
/* solver_compute_force_stiffness (psolve.c:3962): no-op when damping is BKT.  Effective and conventional are the
 * same operator (see HGPU_FLAG_DENSE_K). */
int hgpu_force_stiffness(hgpu_solver_t *s);
/* solver_compute_force_damping (psolve.c:3983): Rayleigh/MASS damping_addforce, or BKT
 * calc_conv + constant_Q_addforce; no-op for NONE. */
int hgpu_force_damping(hgpu_solver_t *s);
/* solver_send_force_dangling + solver_adjust_forces + solver_send_force_anchored
 * (psolve.c:4036, 4048, 4058). */
int hgpu_force_exchange(hgpu_solver_t *s);
/* solver_compute_displacement (psolve.c:4072): central difference into tm2, force = 0. */
int hgpu_update(hgpu_solver_t *s);
/* solver_send_displacement_anchored + solver_adjust_displacement +
 * solver_send_displacement_dangling (psolve.c:4130, 4144, 4154). */
int hgpu_disp_exchange(hgpu_solver_t *s);

/* The seven calls above in sequence; F may be NULL when nloaded == 0. */
int hgpu_step(hgpu_solver_t *s, int32_t step, const double *F);
/* Source streaming (read_myForces does fseeko+fread per step, psolve.c:3651-3667):
 * copy F_all = [nsteps][nloaded][3] host doubles (rows step0.. of force_process.<rank>) to HBM once. */
int hgpu_source_preload(hgpu_solver_t *s, int32_t step0, int32_t nsteps, const double *F_all);
/* compute_addforce_s (psolve.c:5912) for one step from the rows hgpu_source_preload left in HBM: what
 * hgpu_force_source does, without the per-step host read (read_myForces) and host->device copy. */
int hgpu_force_source_resident(hgpu_solver_t *s, int32_t step);
/* nsteps steps starting at step0 with the source history resident in HBM.  F_all non-NULL =
 * preload rows for [step0, step0+nsteps) first; NULL = use what hgpu_source_preload left
 * resident (it must cover those steps). */
int hgpu_run(hgpu_solver_t *s, int32_t step0, int32_t nsteps, const double *F_all);

/* ---- output taps and restart (host <-> device only when the host asks) ------------------- */

/* Sparse read for stations and planes (psolve.c:6680, io_planes.c:151): out[i] = which[lnid[i]]. */
int hgpu_fetch_nodes(hgpu_solver_t *s, int32_t which, const int32_t *lnid, int32_t n, double *out);
/* Full read for 4D output and checkpoints (output.c:1233, io_checkpoint.c:29): [nharbored][3]
 * (conv arrays: [8*lenum][3]). */
int hgpu_fetch_all(hgpu_solver_t *s, int32_t which, double *out);
/* The same without stopping the time loop (checkpoints and 4D frames): the array (HGPU_TM1..3) is
 * snapshotted on the device in stream order and copied to `out` (page-locked memory from hgpu_host_alloc for a
 * truly asynchronous copy) on a separate stream; the call returns at once.  hgpu_fetch_wait blocks until every
 * such read has landed; it only waits on events and may be called from a writer thread. */
int hgpu_fetch_all_async(hgpu_solver_t *s, int32_t which, double *out);
int hgpu_fetch_wait(hgpu_solver_t *s);
/* Restart after checkpoint_read (psolve.c:4249), and test set-up: overwrite a device array. */
int hgpu_store_all(hgpu_solver_t *s, int32_t which, const double *in);

/* ---- stations on the device ------------------------------------------------
 * interpolate_station_displacements (psolve.c:6680-6795) evaluated by a kernel, so that a station
 * step neither drains the device pipeline nor moves 8 nodes per station to the host: the rows
 * are collected in a device ring and read back many steps at a time.  The reference's writer keeps
 * the "station.N" text format (psolve.c:6733-6776); INTEGRATION.md shows the loop.
 *
 * attach: nodes = station_t.nodestointerpolate [nstations][8], localcoords = station_t.localcoords
 *   [nstations][3] (psolve.c:6597-6660); print_vel / print_acc = Param.printStationVelocities /
 *   Accelerations (acc needs hgpu_params_t.print_accel); rate = Param.theStationsPrintRate: hgpu_run
 *   then records by itself on every step with step % rate == 0 (0 = explicit hgpu_stations_record
 *   calls only); capacity = rows the ring holds between two drains.
 * record: at the point of the step where solver_output_stations runs (psolve.c:4281, after the
 *   swap): row = for every station dis[3], vel[3] = (u1-u2)/dt, acc[3] = (u1-2 u2+u3)/dt2, computed
 *   with the reference's operation order, unfused -- bit-equal to the reference's doubles for equal
 *   fields.  Asynchronous.  HGPU_ESTATE when the ring is full.
 * drain: waits for the recorded rows and copies them out: rows[r][station][9], steps[r] (may be
 *   NULL); max_rows must cover hgpu_stations_pending(). */
int hgpu_stations_attach(hgpu_solver_t *s, int32_t nstations, const int32_t *nodes, const double *localcoords,
                         int32_t print_vel, int32_t print_acc, int32_t rate, int32_t capacity);
int hgpu_stations_record(hgpu_solver_t *s, int32_t step);
int hgpu_stations_pending(hgpu_solver_t *s);
int hgpu_stations_drain(hgpu_solver_t *s, double *rows, int32_t *steps, int32_t max_rows, int32_t *nrows);

/* ---- planes on the device ----------------------------------------------------
 * The interpolation of Old_planes_print (io_planes.c:168-191) evaluated by a kernel: a plane step moves 3
 * doubles per plane point to the host instead of the 8 nodes it is interpolated from, and does not stop the
 * time loop.  The strip transport to the printing rank and the file format stay the reference's
 * (io_planes.c:193-250); integration/io_planes_gpu.c shows them on these rows.
 *
 * attach: the plane points of this rank in plane / strip / point order: nodes = plane_strip_element_t.
 *   nodestointerpolate [npoints][8], localcoords = .localcoords [npoints][3] (io_planes.c:83-88).
 * record: at the point of the step where solver_output_planes runs (psolve.c:4283, after the swap):
 *   out[point][3] = sum_i phi_i(localcoords) tm1[node_i], the reference's operation order, unfused --
 *   bit-equal to the reference's doubles for equal fields.  Asynchronous (kernel in stream order, copy on
 *   the copy stream; `out` page-locked from hgpu_host_alloc for a truly asynchronous copy); two records may
 *   be in flight, each into its own `out`.
 * wait: blocks until every recorded step has landed; only waits on events (writer-thread safe). */
int hgpu_planes_attach(hgpu_solver_t *s, int64_t npoints, const int32_t *nodes, const double *localcoords);
int hgpu_planes_record(hgpu_solver_t *s, double *out);
int hgpu_planes_wait(hgpu_solver_t *s);

/* Page-locked host memory for the buffers handed to hgpu_fetch_all / hgpu_store_all / hgpu_fetch_nodes /
 * hgpu_force_source (the calloc'ed tm1/tm2 of solver_init, psolve.c:3317-3325, on the host side):
 * copies to and from it run at full PCIe/C2C speed and without a bounce buffer.  Any host pointer
 * is accepted by those calls; memory from here is just faster.  NULL on failure. */
void *hgpu_host_alloc(size_t bytes);
void hgpu_host_free(void *p);

int hgpu_sync(hgpu_solver_t *s);
int hgpu_get_timers(hgpu_solver_t *s, hgpu_timers_t *out);
/* Raw CUDA stream the kernels are launched on (cudaStream_t), for event timing by the caller. */
void *hgpu_stream(hgpu_solver_t *s);

/* Sizes of the device data structures, for DESIGN.md / bench reporting. */
typedef struct hgpu_layout {
    int32_t tile_nodes, ntiles, max_tile_nodes, max_tile_elems;
    int64_t tile_elems_total;   /* sum over tiles of elements evaluated (= lenum + the extra entries of self tiles) */
    int64_t tile_halo_total;    /* sum over tiles of gathered non-owned nodes */
    int64_t n_regular, n_special;
    int64_t device_bytes;
    int32_t smem_bytes, block_threads;
    int32_t grid_ctas, ctas_per_sm;   /* persistent step kernel: CTAs launched, resident per SM */
    int32_t early_tiles;              /* self tiles: evaluated before the halo exchange starts (multi-GPU) */
    /* modelled shared-memory wavefronts per 32-lane 8-byte access (2.0 = conflict-free) */
    double est_gather_wavefronts, est_scatter_wavefronts;
    int32_t max_tile_acc;             /* owned + published nodes of the largest tile (shared-memory accumulator) */
    int32_t max_tile_recs, max_tile_srcs;
    int32_t struct_tiles;             /* tiles that take the structured path (aligned uniform 8x8x8 cells of one material) */
    int64_t partial_slots;            /* (tile, node) partial forces exchanged between tiles per pass */
    int64_t deps_total;               /* sum over tiles of the lower tiles they wait for */
} hgpu_layout_t;
int hgpu_get_layout(hgpu_solver_t *s, hgpu_layout_t *out);
/* Host-only (no device needed): build and self-check the tile plan hgpu_init would use for this
 * mesh on a B200 and report its sizes.  Only lenum, nharbored and elem_lnid are read. */
int hgpu_plan_build(const hgpu_mesh_t *mesh, int32_t tile_nodes, hgpu_layout_t *out);

#ifdef __cplusplus
}
#endif

#endif /* HERCULES_GPU_H */
