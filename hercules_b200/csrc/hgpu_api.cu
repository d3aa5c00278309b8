// hgpu_api.cu -- C ABI of libhercules_gpu.so (include/hercules_gpu.h): solver handle, device
// data layout, the per-step call sequence of solver_run (psolve.c:4265-4319) and the halo
// exchange that replaces schedule_senddata (psolve.c:4945-5079).
//
// There is no CPU fallback anywhere in this file: every entry point either runs CUDA kernels
// or fails with a negative HGPU_E* code.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "hgpu_internal.h"
#include "hgpu_kernels.cuh"

using namespace hgpu;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(HGPU_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                              \
    } while (0)

// ---- minimal NCCL binding, resolved at run time so single-GPU use needs no libnccl ----------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl()
{
    if (g_nccl.h) return HGPU_OK;
    const char *names[] = {getenv("HGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    }
    if (!h) return fail(HGPU_ECOMM, "cannot dlopen libnccl (set HGPU_NCCL_LIB): %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.Send = (int (*)(const void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclSend");
    g_nccl.Recv = (int (*)(void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv ||
        !g_nccl.GroupStart || !g_nccl.GroupEnd)
        return fail(HGPU_ECOMM, "libnccl lacks a required symbol");
    g_nccl.h = h;
    return HGPU_OK;
}

#define NK(call)                                                                       \
    do {                                                                               \
        int r_ = (call);                                                               \
        if (r_ != ncclSuccess)                                                         \
            return fail(HGPU_ECOMM, "%s failed: %s", #call,                            \
                        g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?");      \
    } while (0)

// One side of a schedule (c-list or s-list) on the device.
struct MsgList {
    std::vector<int32_t> peer, nodes, off;   // off = prefix sum of nodes
    int32_t total = 0;
    int32_t *d_map = nullptr;                // concatenated mapping[]
    double *d_send = nullptr, *d_recv = nullptr;   // [total][3] staging (NCCL transport)
    // peer-memory transport: this list's receive area inside my mailbox, and where each of my
    // messengers writes in its peer's mailbox
    size_t mb_data_off = 0, mb_flag_off = 0;       // bytes from the mailbox base
    std::vector<double *> remote_data;             // [messenger] peer segment (parity 0)
    std::vector<unsigned long long *> remote_flag; // [messenger] peer flags (2 parities)
    unsigned int *d_counters = nullptr;            // [messenger]
    PushSeg *d_push = nullptr;                     // [2 parities][messenger]
    PullSeg *d_pull = nullptr;                     // [2 parities][messenger]
    unsigned long long seq_out = 0, seq_in = 0;
    // contribution receive side in ONE launch: node-major CSR over all messengers (list order kept)
    int32_t csr_nodes = 0;
    int32_t *d_csr_node = nullptr, *d_csr_off = nullptr, *d_csr_src[2] = {nullptr, nullptr};
};

enum ForceState { F_CLEAN = 0, F_PENDING = 1, F_MATERIALIZED = 2, F_FUSED_DONE = 3 };

// Device-time phases, in the order of hgpu_timers_t's double fields.
enum Phase { PH_ADDFORCE_S = 0, PH_ADDFORCE_E, PH_DAMPING, PH_SEND_DN_FORCE, PH_ADJUST_FORCE,
             PH_SEND_AN_FORCE, PH_NEW_DISP, PH_SEND_AN_DISP, PH_ADJUST_DISP, PH_SEND_DN_DISP,
             PH_FUSED_STEP, PH_COUNT };
struct EvPair { cudaEvent_t a = nullptr, b = nullptr; int phase = 0; };

struct hgpu_solver {
    hgpu_params_t P{};
    int32_t E = 0, N = 0, D = 0;
    int dev = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;      // halo exchange + hanging-node transfer while late tiles run
    cudaEvent_t ev_early = nullptr, ev_comm = nullptr;
    // node arrays
    double *u[3] = {nullptr, nullptr, nullptr};
    int i1 = 0, i2 = 1, i3 = 2;          // which buffer plays tm1 / tm2 / tm3
    double *force = nullptr;
    double *mass = nullptr, *m2 = nullptr, *m1 = nullptr;
    // element arrays
    double *etab = nullptr;
    double *Kd = nullptr;
    // tiles
    TilePlan plan;
    int4 *t_meta = nullptr;              // per tile in processing order (early tiles first)
    int32_t n_early = 0;                 // tiles whose nodes take part in the halo / hanging-node phases
    int32_t *t_halo_id = nullptr;
    uint4 *t_ent_slot = nullptr;         // per entry 8 x uint16 = 3 * slot
    double *t_ent_coef = nullptr;        // per entry c1, c2, beta
    double *t_beta = nullptr;            // per tile (processing order): the entries' common beta, or NaN
    double *t_coef = nullptr;            // per tile (processing order): {c1, c2, beta, -} of a structured tile
    int32_t n_struct = 0;                // structured tiles of one material (take the STRUCT path of the step kernel)
    int4 *t_meta_s = nullptr;            // tiles in the order of the STRUCT launches: slot-table tiles, then structured ones
    int32_t n_generic = 0;               // slot-table tiles (= index of the first structured tile in t_meta_s)
    int64_t struct_entries = 0, generic_entries = 0;   // elements evaluated by structured / by late slot-table tiles
    double generic_cost = 2.6;           // cost of a slot-table element relative to a structured one (CTA split)
    std::vector<int64_t> ent_prefix_s;   // [ntiles + 1] prefix sum of entries over t_meta_s
    int *d_queue = nullptr;              // {next slot-table tile, next structured tile} of a STRUCT launch (HGPU_DYNAMIC=0: unused)
    bool dynamic_tiles = false;          // HGPU_DYNAMIC=1: tiles from two shared counters (measured slower: consecutive tiles on one CTA wait for each other)
    uint2 *t_rec = nullptr;              // finish records
    int32_t *t_src = nullptr, *t_dep = nullptr;
    double *t_partial = nullptr;         // [halo slots][3] partial forces published by lower tiles
    unsigned int *t_flag = nullptr;      // [ntiles] epoch of each tile's last publish
    unsigned int epoch = 0;              // one per pass over all tiles
    int32_t n_self = 0;
    // BKT: memory variables (entry-chunked), per-entry coefficient records, element -> entry
    double *conv = nullptr; double *t_ent_bkt = nullptr; int32_t *entry_of_elem = nullptr;
    double *conv_scratch = nullptr;      // [8 E][3] staging for hgpu_fetch_all / hgpu_store_all of a conv array
    double *nt3 = nullptr;               // [N][3] {+-1/mass, m2, m1} for the fused update
    int smem_struct = 0, cap_acc_struct = 0, grid_struct_max = 0;   // STRUCT launches: one CTA of 512 threads per SM
    int smem_u2 = 0, smem_nou2 = 0, block = 256, grid = 0, grid_late = 0, cap_slots = 0, cap_acc = 0, cap_owned = 0,
        cap_recs = 0, cap_srcs = 0, ctas_per_sm = 0;
    // special-node path
    int32_t nS = 0; int32_t *d_slist = nullptr;
    // the first nS_early entries of d_slist are owned by self tiles: their forces are final once the
    // early tiles and the force exchange are done, so their update and the displacement exchange can
    // run on the communication stream while the late tiles are evaluated (tail_done: they did)
    int32_t nS_early = 0; bool tail_done = false;
    int32_t *d_loaded = nullptr; double *d_F = nullptr; double *h_F = nullptr;
    // per-step source staging: SRC_RING pinned + device slots so hgpu_force_source never has to
    // wait for the previous step (slot k is reusable once its own copy has completed)
    static constexpr int SRC_RING = 8;
    cudaEvent_t src_done[SRC_RING] = {nullptr}; int src_slot = 0;
    double *d_Fall = nullptr; size_t Fall_steps = 0, Fall_loaded = 0; int32_t Fall_step0 = 0;
    int32_t *d_dnode = nullptr;
    int32_t nA = 0; int32_t *d_anchor_id = nullptr, *d_anchor_off = nullptr, *d_anchor_dn = nullptr,
            *d_anchor_deps = nullptr;
    // halo
    MsgList dn_c, dn_s, an_c, an_s;
    ncclComm_t comm = nullptr;
    // peer-memory transport
    char *mailbox = nullptr; size_t mailbox_bytes = 0;
    std::vector<void *> peer_base;           // [rank] IPC-mapped mailbox of each peer
    bool p2p_ready = false;
    bool cooperative = true;                 // step kernels are launched cooperatively (HGPU_COOPERATIVE=0 turns it off)
    int *h_err = nullptr, *d_err = nullptr;  // error word: page-locked host memory mapped into the device
    long long p2p_timeout_cycles = 0;        // halo wait bound in SM cycles (HGPU_P2P_TIMEOUT_S, default 600 s; 0 = none)
    // step state
    bool want_stiff = false, want_damp = false;
    ForceState fstate = F_CLEAN;
    int64_t n_regular = 0, n_special = 0, device_bytes = 0;
    hgpu_timers_t tm{};
    // CUDA-event phase timing (HGPU_FLAG_TIMERS): a pool of event pairs drained at sync points
    std::vector<EvPair> evpool;
    size_t ev_used = 0;
    double phase_s[PH_COUNT] = {0};
    // asynchronous whole-field reads (hgpu_fetch_all_async): device-side snapshots + a copy stream
    struct AsyncFetch { double *snap = nullptr; size_t cap = 0; cudaEvent_t ready = nullptr, done = nullptr; bool busy = false; };
    AsyncFetch af[2];
    cudaStream_t copy_stream = nullptr;
    // scratch for fetch
    int32_t *d_fetch_ids = nullptr; double *d_fetch_out = nullptr; int32_t fetch_cap = 0;
    std::vector<int32_t> fetch_ids_host;     // the list d_fetch_ids holds
    // stations interpolated on the device (hgpu_stations_*): ring of rows [capacity][nst][9]
    int32_t st_n = 0, st_vel = 0, st_acc = 0, st_rate = 0, st_cap = 0, st_count = 0;
    int32_t *d_st_nodes = nullptr; double *d_st_local = nullptr, *d_st_rows = nullptr;
    std::vector<int32_t> st_steps;           // step of each recorded row
    // planes interpolated on the device (hgpu_planes_*): two row buffers [npoints][3], read out on the copy stream
    struct PlaneSlot { double *rows = nullptr; cudaEvent_t ready = nullptr, done = nullptr; bool busy = false; };
    int64_t pl_n = 0;
    int32_t *d_pl_nodes = nullptr; double *d_pl_local = nullptr;
    PlaneSlot pl[2];
    int pl_next = 0;
};

template <typename T>
static int dalloc(hgpu_solver *s, T **p, size_t n)
{
    *p = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e != cudaSuccess)
        return fail(HGPU_ENOMEM, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
    s->device_bytes += (int64_t)(n * sizeof(T));
    return HGPU_OK;
}

template <typename T>
static int upload(hgpu_solver *s, T **p, const T *src, size_t n)
{
    int rc = dalloc(s, p, n);
    if (rc) return rc;
    if (n) CK(cudaMemcpy(*p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return HGPU_OK;
}

static int upload_msglist(hgpu_solver *s, const hgpu_msglist_t &in, MsgList &out, const char *what)
{
    out.off.assign(1, 0);
    for (int32_t i = 0; i < in.count; i++) {
        if (in.peer[i] < 0 || in.peer[i] >= s->P.nranks || in.peer[i] == s->P.rank)
            return fail(HGPU_EINVAL, "%s: bad peer rank %d", what, in.peer[i]);
        out.peer.push_back(in.peer[i]);
        out.nodes.push_back(in.nodes[i]);
        out.off.push_back(out.off.back() + in.nodes[i]);
    }
    out.total = out.off.back();
    for (int32_t i = 0; i < out.total; i++)
        if (in.mapping[i] < 0 || in.mapping[i] >= s->N)
            return fail(HGPU_EINVAL, "%s: mapping entry out of range", what);
    int rc;
    if ((rc = upload(s, &out.d_map, in.mapping, (size_t)out.total))) return rc;
    if ((rc = dalloc(s, &out.d_send, 3 * (size_t)out.total))) return rc;
    if ((rc = dalloc(s, &out.d_recv, 3 * (size_t)out.total))) return rc;
    return HGPU_OK;
}

// Sum the elapsed time of every recorded event pair into phase_s (synchronises the stream).
static void drain_events(hgpu_solver *s)
{
    if (!s->ev_used) return;
    cudaStreamSynchronize(s->stream);
    if (s->comm_stream) cudaStreamSynchronize(s->comm_stream);
    for (size_t i = 0; i < s->ev_used; i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s->evpool[i].a, s->evpool[i].b) == cudaSuccess)
            s->phase_s[s->evpool[i].phase] += 1e-3 * (double)ms;
    }
    s->ev_used = 0;
}

// Brackets the launches of one phase with a CUDA event pair on the solver's stream.
struct PhaseTimer {
    hgpu_solver *s; EvPair *p = nullptr;
    cudaStream_t st;
    PhaseTimer(hgpu_solver *s_, int phase, cudaStream_t stream = nullptr) : s(s_), st(stream ? stream : s_->stream)
    {
        if (!(s->P.flags & HGPU_FLAG_TIMERS)) return;
        if (s->ev_used == s->evpool.size()) {
            if (s->evpool.size() >= 8192) drain_events(s);
            else {
                EvPair e;
                if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
                s->evpool.push_back(e);
            }
        }
        p = &s->evpool[s->ev_used++];
        p->phase = phase;
        cudaEventRecord(p->a, st);
    }
    ~PhaseTimer() { if (p) cudaEventRecord(p->b, st); }
};

// The device error word (mapped host memory), read after a synchronisation of the solver's streams.
static int check_device_error(hgpu_solver *s)
{
    const int e = s->h_err ? *(volatile int *)s->h_err : 0;
    if (e == 1) return fail(HGPU_ECOMM, "halo exchange timed out waiting for a peer (HGPU_P2P_TIMEOUT_S); "
                                        "nothing was applied from that exchange on: the state is not valid");
    if (e == 2) return fail(HGPU_ECUDA, "a tile waited for its lower tiles for too long (a CTA of the step kernel "
                                        "was not resident?); the state is not valid");
    return HGPU_OK;
}

extern "C" const char *hgpu_last_error(void) { return g_err.c_str(); }
extern "C" int hgpu_abi_version(void) { return HGPU_ABI_VERSION; }

extern "C" int hgpu_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(HGPU_ENODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

// Shared memory of the step kernel, in bytes, for given capacities (u2e: the elements read u2).
static int finish_buf_bytes(int32_t cap_recs, int32_t cap_srcs) { return 8 * cap_recs + 4 * cap_srcs + 4 * CAP_DEPS; }
// bytes besides the stages and the accumulator: pend [recs][3] + spart [srcs][3] + srm [recs]
// doubles, and two buffers of finish data
static int finish_smem_bytes(int32_t cap_recs, int32_t cap_srcs)
{
    return (4 * cap_recs + 3 * cap_srcs) * (int)sizeof(double) + 2 * finish_buf_bytes(cap_recs, cap_srcs);
}
static int step_smem_bytes(bool u2e, int32_t cap_slots, int32_t cap_acc, int32_t cap_owned, int32_t cap_recs, int32_t cap_srcs)
{
    const int stage = 3 * cap_slots + 3 * (u2e ? cap_slots : cap_owned);
    return (2 * stage + 3 * cap_acc) * (int)sizeof(double) + finish_smem_bytes(cap_recs, cap_srcs);
}

// Tile capacities for a device with max_smem bytes of opt-in shared memory per CTA.
// Two CTAs per SM: each may use half of the SM's shared memory minus the 1 KB the system reserves
// per CTA.  Per CTA: 2 stages x (u1 + u2) x cap_slots nodes + the accumulator (owned + published
// nodes) + the pending buffer (owned nodes) + 2 finish buffers.
static TileCaps tile_caps(int max_smem, int32_t tile_nodes, bool allow_struct = false)
{
    const int per_cta = std::min(max_smem, (max_smem + 1024) / 2 - 1024);
    TileCaps c;
    c.max_recs = 256; c.max_srcs = 320;      // multiples of 16
    int32_t cap_owned = tile_nodes > 0 ? tile_nodes : 730;
    const char *env = getenv("HGPU_TILE_NODES");
    if (tile_nodes <= 0 && env && atoi(env) > 0) cap_owned = atoi(env);
    cap_owned = std::max(2, cap_owned & ~1);
    c.elem_block = 512;
    const char *eenv = getenv("HGPU_ELEM_BLOCK");
    if (eenv && atoi(eenv) > 0) c.elem_block = atoi(eenv);
    const int budget = (per_cta - finish_smem_bytes(c.max_recs, c.max_srcs)) / 8;     // doubles
    // an interior tile of a uniform region stages and accumulates 9^3 nodes and owns 8^3 of them
    const int32_t rest = budget / 15;                      // 12 S + 3 A with S = A
    c.max_owned = cap_owned;
    c.max_acc = std::max(16, std::min(rest, 65535 / 3) & ~15);
    c.max_slots = c.max_acc;
    c.max_owned = std::min(c.max_owned, c.max_acc);
    c.allow_struct = allow_struct ? 1 : 0;
    return c;
}

static inline int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

extern "C" int hgpu_init(hgpu_solver_t **out, const hgpu_mesh_t *mesh, const hgpu_params_t *params)
{
    if (!out || !mesh || !params) return fail(HGPU_EINVAL, "hgpu_init: null argument");
    *out = nullptr;
    if (mesh->lenum < 0 || mesh->nharbored <= 0 || mesh->ldnnum < 0)
        return fail(HGPU_EINVAL, "hgpu_init: bad mesh counts");
    if (!mesh->elem_lnid || !mesh->eTable || !mesh->nTable)
        return fail(HGPU_EINVAL, "hgpu_init: elem_lnid, eTable and nTable are required");
    if (params->damping < 0 || params->damping > 3 || params->stiffness < 0 || params->stiffness > 1)
        return fail(HGPU_EINVAL, "hgpu_init: bad damping/stiffness type");
    if (params->damping == HGPU_DAMPING_BKT && !mesh->edata)
        return fail(HGPU_EINVAL, "hgpu_init: BKT damping needs edata (psolve.h:95-97)");
    if (params->stiffness == HGPU_STIFFNESS_CONVENTIONAL && (!mesh->K1 || !mesh->K2))
        return fail(HGPU_EINVAL, "hgpu_init: conventional stiffness needs K1 and K2");
    if (params->nranks < 1 || params->rank < 0 || params->rank >= params->nranks)
        return fail(HGPU_EINVAL, "hgpu_init: bad rank/nranks");
    if (params->nloaded < 0 || (params->nloaded > 0 && !params->loaded_lnid))
        return fail(HGPU_EINVAL, "hgpu_init: bad loaded-node list");

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0)
        return fail(HGPU_ENODEVICE, "no CUDA device: %s (this library has no CPU path)",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");

    hgpu_solver *s = new hgpu_solver();
    s->P = *params;
    s->P.loaded_lnid = nullptr;
    s->E = mesh->lenum; s->N = mesh->nharbored; s->D = mesh->ldnnum;
    s->dev = params->device >= 0 ? params->device : params->rank % ndev;
    if (s->dev >= ndev) { delete s; return fail(HGPU_ENODEVICE, "device %d not present", params->device); }
    int rc = HGPU_OK;
#define TRY(x) do { rc = (x); if (rc) { hgpu_finalize(s); return rc; } } while (0)
#define TRYCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { hgpu_finalize(s); \
        return fail(HGPU_ECUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); } } while (0)
    TRYCU(cudaSetDevice(s->dev));
    cudaDeviceProp prop;
    TRYCU(cudaGetDeviceProperties(&prop, s->dev));
    if (prop.major < 10) {
        hgpu_finalize(s);
        return fail(HGPU_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                    s->dev, prop.major, prop.minor);
    }
    TRYCU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    TRYCU(cudaHostAlloc((void **)&s->h_err, sizeof(int), cudaHostAllocMapped));
    *s->h_err = 0;
    TRYCU(cudaHostGetDevicePointer((void **)&s->d_err, s->h_err, 0));
    {
        // halo wait bound: host-side skew between ranks (a peer writing a checkpoint or a 4D frame)
        // is legitimate and unbounded in principle, so the default is long; 0 waits for ever
        double secs = 600.0;
        const char *tenv = getenv("HGPU_P2P_TIMEOUT_S");
        if (tenv) secs = atof(tenv);
        s->p2p_timeout_cycles = secs > 0 ? (long long)(secs * 1e3 * (double)prop.clockRate) : 0;
    }

    const int32_t E = s->E, N = s->N, D = s->D;
    const size_t n3 = 3 * (size_t)N;

    // ---- node arrays: three rotating displacement buffers, force, split n_t -----------------
    for (int b = 0; b < 3; b++) {
        TRY(dalloc(s, &s->u[b], n3));
        TRYCU(cudaMemset(s->u[b], 0, n3 * sizeof(double)));
    }
    TRY(dalloc(s, &s->force, n3));
    TRYCU(cudaMemset(s->force, 0, n3 * sizeof(double)));
    {
        std::vector<double> mass((size_t)N), m2(n3), m1(n3);
        for (int32_t n = 0; n < N; n++) {
            const double *np = mesh->nTable + 7 * (size_t)n;
            mass[n] = np[0];
            for (int c = 0; c < 3; c++) { m2[3 * (size_t)n + c] = np[1 + c]; m1[3 * (size_t)n + c] = np[4 + c]; }
        }
        TRY(upload(s, &s->mass, mass.data(), (size_t)N));
        TRY(upload(s, &s->m2, m2.data(), n3));
        TRY(upload(s, &s->m1, m1.data(), n3));
    }
    TRY(upload(s, &s->etab, mesh->eTable, 4 * (size_t)E));
    if (params->stiffness == HGPU_STIFFNESS_CONVENTIONAL) {
        // [8][8][3][3] block form -> two dense 24x24 row-major matrices
        std::vector<double> kd(2 * 576);
        for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) {
            kd[(size_t)24 * (3 * i + k) + 3 * j + l] = mesh->K1[9 * (8 * i + j) + 3 * k + l];
            kd[576 + (size_t)24 * (3 * i + k) + 3 * j + l] = mesh->K2[9 * (8 * i + j) + 3 * k + l];
        }
        TRY(upload(s, &s->Kd, kd.data(), kd.size()));
    }

    // ---- hanging nodes ------------------------------------------------------------------------
    DanglingPlan dp;
    std::string err;
    if (D > 0 && !mesh->dnode) { hgpu_finalize(s); return fail(HGPU_EINVAL, "ldnnum > 0 but dnode is null"); }
    if (!build_dangling_plan(N, D, mesh->dnode, dp, err)) { hgpu_finalize(s); return fail(HGPU_EINVAL, "%s", err.c_str()); }
    TRY(upload(s, &s->d_dnode, mesh->dnode, 6 * (size_t)D));
    s->nA = (int32_t)dp.anchor_id.size();
    TRY(upload(s, &s->d_anchor_id, dp.anchor_id.data(), dp.anchor_id.size()));
    TRY(upload(s, &s->d_anchor_off, dp.anchor_off.data(), dp.anchor_off.size()));
    TRY(upload(s, &s->d_anchor_dn, dp.anchor_dn.data(), dp.anchor_dn.size()));
    TRY(upload(s, &s->d_anchor_deps, dp.anchor_deps.data(), dp.anchor_deps.size()));

    // ---- halo schedules ------------------------------------------------------------------------
    TRY(upload_msglist(s, mesh->dn_c, s->dn_c, "dn_sched c-list"));
    TRY(upload_msglist(s, mesh->dn_s, s->dn_s, "dn_sched s-list"));
    TRY(upload_msglist(s, mesh->an_c, s->an_c, "an_sched c-list"));
    TRY(upload_msglist(s, mesh->an_s, s->an_s, "an_sched s-list"));

    // ---- source ----------------------------------------------------------------------------------
    for (int32_t i = 0; i < params->nloaded; i++)
        if (params->loaded_lnid[i] < 0 || params->loaded_lnid[i] >= N) {
            hgpu_finalize(s); return fail(HGPU_EINVAL, "loaded node id out of range");
        }
    TRY(upload(s, &s->d_loaded, params->loaded_lnid, (size_t)params->nloaded));
    TRY(dalloc(s, &s->d_F, hgpu_solver::SRC_RING * 3 * (size_t)params->nloaded));
    TRYCU(cudaMallocHost((void **)&s->h_F, std::max<size_t>(1, hgpu_solver::SRC_RING * 3 * (size_t)params->nloaded) * sizeof(double)));
    for (int i = 0; i < hgpu_solver::SRC_RING; i++)
        TRYCU(cudaEventCreateWithFlags(&s->src_done[i], cudaEventDisableTiming));

    std::vector<uint8_t> early_node((size_t)N, 0);   // nodes the exchange / hanging-node phases touch
    std::vector<int32_t> slist;                      // SPECIAL nodes
    // ---- node classes -----------------------------------------------------------------------------
    // A node is advanced inside the fused step kernel unless something else must see or change its
    // force first (source assignment, hanging-node transfer, halo exchange), or unless its
    // mass2_minusaM / mass_minusaM differ between components (absorbing-boundary dashpots,
    // psolve.c:3445-3473): the fused path keeps one scalar of each per node.
    std::vector<uint8_t> cls((size_t)N, NODE_REGULAR);
    {
        if (params->flags & HGPU_FLAG_NO_FUSE) std::fill(cls.begin(), cls.end(), (uint8_t)NODE_SPECIAL);
        for (int32_t i = 0; i < params->nloaded; i++) cls[params->loaded_lnid[i]] = NODE_SPECIAL;
        const bool multi = params->nranks > 1;
        for (int32_t d = 0; d < D; d++) {
            const int32_t *dn = mesh->dnode + 6 * (size_t)d;
            cls[dn[0]] = NODE_SPECIAL;
            if (multi) early_node[dn[0]] = 1;
            for (int a = 0; a < 4 && dn[2 + a] >= 0; a++) { cls[dn[2 + a]] = NODE_SPECIAL; if (multi) early_node[dn[2 + a]] = 1; }
        }
        const hgpu_msglist_t *lists[4] = {&mesh->dn_c, &mesh->dn_s, &mesh->an_c, &mesh->an_s};
        for (const hgpu_msglist_t *l : lists) {
            int32_t tot = 0;
            for (int32_t i = 0; i < l->count; i++) tot += l->nodes[i];
            for (int32_t i = 0; i < tot; i++) { cls[l->mapping[i]] = NODE_SPECIAL; early_node[l->mapping[i]] = 1; }
        }
        std::vector<double> nt3(n3);
        for (int32_t n = 0; n < N; n++) {
            const double *np = mesh->nTable + 7 * (size_t)n;
            const bool iso = np[1] == np[2] && np[1] == np[3] && np[4] == np[5] && np[4] == np[6];
            if (!iso || !(np[0] > 0.0)) cls[n] = NODE_SPECIAL;
            const double rm = np[0] > 0.0 ? 1.0 / np[0] : 1.0;
            nt3[3 * (size_t)n] = cls[n] == NODE_SPECIAL ? -rm : rm;
            // m2, m1 are only ever applied to REGULAR nodes; zero for SPECIAL ones so that the WPASS
            // variant can seed the accumulator with m2 u1 - m1 u2 without looking at the class
            nt3[3 * (size_t)n + 1] = cls[n] == NODE_SPECIAL ? 0.0 : np[1];
            nt3[3 * (size_t)n + 2] = cls[n] == NODE_SPECIAL ? 0.0 : np[4];
        }
        for (int32_t n = 0; n < N; n++) if (cls[n] == NODE_SPECIAL) slist.push_back(n);
        s->nS = (int32_t)slist.size();
        s->n_special = s->nS; s->n_regular = (int64_t)N - s->nS;
        TRY(upload(s, &s->nt3, nt3.data(), n3));
    }

    // ---- tiles ------------------------------------------------------------------------------------
    {
        int max_smem = 0, nsm = 0;
        TRYCU(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->dev));
        TRYCU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, s->dev));
        // tiles owning a node of the exchange / hanging-node phases must not wait for anybody
        // ("self" tiles); without the fused update every node's force goes to the force array and
        // no node needs a record of its own
        const bool fused = !(params->flags & HGPU_FLAG_NO_FUSE);
        // BKT: an element's memory variables are advanced by its one core evaluation, so no tile may
        // re-evaluate foreign elements: no self tiles, the exchange follows the whole pass
        const bool bkt = params->damping == HGPU_DAMPING_BKT;
        // structured tiles (aligned uniform 8x8x8 cells): the fused effective-stiffness kernels only
        bool want_struct = fused && !bkt && params->stiffness == HGPU_STIFFNESS_EFFECTIVE &&
                           !(params->flags & (HGPU_FLAG_WPASS | HGPU_FLAG_NO_STRUCT));
        {   // opt-in while it does not beat the slot-table path on the bench (profiles/README.md): HGPU_STRUCT=1
            const char *senv = getenv("HGPU_STRUCT");
            if (!(senv && atoi(senv) == 1) && !(params->flags & HGPU_FLAG_STRUCT)) want_struct = false;
        }
        TileCaps caps = tile_caps(max_smem, params->tile_nodes, want_struct);
        for (;;) {
            if (!build_tile_plan(E, N, mesh->elem_lnid, caps, (params->nranks > 1 && !bkt) ? early_node.data() : nullptr,
                                 fused ? cls.data() : nullptr, s->plan, err)) {
                hgpu_finalize(s); return fail(HGPU_EINVAL, "tile plan: %s", err.c_str());
            }
            if (!caps.allow_struct) break;
            // STRUCT launches run ONE CTA of 512 threads per SM: two stages + the planes and the element-force
            // buffer (SP_TOTAL + SF_TOTAL doubles) + the finish buffers must fit the SM's shared memory
            const TilePlan &q = s->plan;
            bool any = false;
            for (uint8_t f : q.tile_struct) any = any || f;
            const int need = step_smem_bytes(true, (q.max_tile_nodes + 15) & ~15, (SP_TOTAL + SF_TOTAL + 2) / 3,
                                             (q.max_tile_owned + 15) & ~15, std::min(caps.max_recs, (q.max_tile_recs + 15) & ~15),
                                             std::min(caps.max_srcs, (q.max_tile_srcs + 15) & ~15));
            if (!any || need <= max_smem) break;
            caps.allow_struct = 0;              // does not fit: plain plan
        }
        TilePlan &pl = s->plan;
        const size_t entries = pl.elem_id.size();
        // entry records: slot offsets premultiplied by 3; c1, c2 and beta = c3/c1 (= c4/c2 = b/dt,
        // psolve.c:3387-3409) copied per entry so a tile streams them without indirection
        for (uint16_t &v : pl.elem_slot) v = (uint16_t)(3 * v);
        std::vector<double> coef(3 * entries);
        for (size_t k = 0; k < entries; k++) {
            const double *et = mesh->eTable + 4 * (size_t)pl.elem_id[k];
            coef[3 * k] = et[0]; coef[3 * k + 1] = et[1];
            coef[3 * k + 2] = et[0] != 0.0 ? et[2] / et[0] : 0.0;
        }
        // processing order: self tiles (they own the nodes the exchange / hanging-node phases read
        // or write) first, so those phases can run while the remaining tiles are evaluated; the
        // rest in ascending order, which is also the order of their dependencies
        {
            std::vector<int32_t> order;
            order.reserve((size_t)pl.ntiles);
            for (int32_t t = 0; t < pl.ntiles; t++) if (pl.tile_self[t]) order.push_back(t);
            s->n_early = s->n_self = (int32_t)order.size();
            for (int32_t t = 0; t < pl.ntiles; t++) if (!pl.tile_self[t]) order.push_back(t);
            // HGPU_ORDER=level (experiment, measured SLOWER: profiles/README.md): the rest by dependency LEVEL
            // (0 = reads nobody's partial forces, else 1 + the highest level among the tiles it reads), ascending
            // ids inside a level -- still a topological order of the dependencies; Z-order keeps the halos and the
            // partial forces of neighbouring tiles in L2, which is worth more than the waiting it causes
            {
                const char *oenv = getenv("HGPU_ORDER");
                if (oenv && strcmp(oenv, "level") == 0) {
                    std::vector<int32_t> lvl((size_t)pl.ntiles, 0);
                    for (int32_t t = 0; t < pl.ntiles; t++)
                        for (int32_t k = pl.dep_off[t]; k < pl.dep_off[(size_t)t + 1]; k++)
                            lvl[t] = std::max(lvl[t], lvl[pl.dep[(size_t)k]] + 1);
                    std::stable_sort(order.begin() + s->n_self, order.end(),
                                     [&](int32_t a, int32_t b) { return lvl[a] < lvl[b]; });
                }
            }
            // SPECIAL nodes owned by self tiles first (ascending within each part)
            if (fused) {
                std::vector<int32_t> early_part, late_part;
                int32_t t = 0;
                for (int32_t n : slist) {
                    while (t + 1 < pl.ntiles && pl.node_off[(size_t)t + 1] <= n) t++;
                    (pl.ntiles > 0 && pl.tile_self[t] ? early_part : late_part).push_back(n);
                }
                s->nS_early = (int32_t)early_part.size();
                early_part.insert(early_part.end(), late_part.begin(), late_part.end());
                TRY(upload(s, &s->d_slist, early_part.data(), early_part.size()));
            }
            std::vector<int32_t> meta((size_t)META_INTS * (size_t)pl.ntiles, 0);
            for (int32_t i = 0; i < pl.ntiles; i++) {
                const int32_t t = order[i];
                int32_t *m = meta.data() + (size_t)META_INTS * (size_t)i;
                m[0] = pl.node_off[t]; m[1] = pl.node_off[(size_t)t + 1];
                m[2] = pl.halo_off[t]; m[3] = pl.halo_off[(size_t)t + 1];
                m[4] = pl.elem_off[t]; m[5] = pl.elem_off[(size_t)t + 1];
                m[6] = pl.elem_core[t]; m[7] = pl.halo_pub[t];
                m[8] = pl.rec_off[t]; m[9] = pl.rec_off[(size_t)t + 1];
                m[10] = pl.src_off[t]; m[11] = pl.src_off[(size_t)t + 1];
                m[12] = pl.dep_off[t]; m[13] = pl.dep_off[(size_t)t + 1];
                m[14] = t;
            }
            // the Rayleigh ratio shared by all entries of a tile (one material), NaN otherwise
            std::vector<double> tbeta((size_t)pl.ntiles, 0.0);
            for (int32_t i = 0; i < pl.ntiles; i++) {
                const int32_t t = order[i];
                const int32_t e0 = pl.elem_off[t], e1 = pl.elem_off[(size_t)t + 1];
                double b = e1 > e0 ? coef[3 * (size_t)e0 + 2] : 0.0;
                for (int32_t k = e0 + 1; k < e1; k++)
                    if (coef[3 * (size_t)k + 2] != b) { b = std::numeric_limits<double>::quiet_NaN(); break; }
                tbeta[i] = b;
            }
            TRY(upload(s, &s->t_beta, tbeta.data(), tbeta.size()));
            TRY(upload(s, (int32_t **)&s->t_meta, meta.data(), meta.size()));
            // Structured tiles of ONE material (c1, c2, c3/c1 equal bit for bit over the cell) get their
            // coefficients per tile and a second processing order for the STRUCT launches, in which they come
            // LAST: [self tiles | other slot-table tiles, ascending | structured tiles, ascending].  A STRUCT
            // launch walks the last part with one set of CTAs and the rest with another (step_kernel).
            std::vector<uint8_t> sok((size_t)pl.ntiles, 0);
            s->n_struct = 0;
            for (int32_t t = 0; t < pl.ntiles; t++) {
                if (pl.tile_struct.empty() || !pl.tile_struct[t] || pl.tile_self[t]) continue;
                const int32_t e0 = pl.elem_off[t], e1 = pl.elem_off[(size_t)t + 1];
                bool same = e1 > e0;
                for (int32_t k = e0 + 1; k < e1 && same; k++)
                    same = memcmp(&coef[3 * (size_t)k], &coef[3 * (size_t)e0], 3 * sizeof(double)) == 0;
                if (same) { sok[t] = 1; s->n_struct++; }
            }
            if (s->n_struct > 0) {
                std::vector<int32_t> order_s;
                order_s.reserve((size_t)pl.ntiles);
                std::vector<int32_t> pos((size_t)pl.ntiles, 0);
                for (int32_t i = 0; i < pl.ntiles; i++) pos[order[i]] = i;
                for (int32_t i = 0; i < pl.ntiles; i++) if (!sok[order[i]]) order_s.push_back(order[i]);
                s->n_generic = (int32_t)order_s.size();
                for (int32_t i = 0; i < pl.ntiles; i++) if (sok[order[i]]) order_s.push_back(order[i]);
                std::vector<int32_t> meta_s(meta.size());
                std::vector<double> tcoef(4 * (size_t)pl.ntiles, 0.0);
                s->struct_entries = 0; s->generic_entries = 0;
                s->ent_prefix_s.assign((size_t)pl.ntiles + 1, 0);
                { const char *cenv = getenv("HGPU_GENERIC_COST"); if (cenv && atof(cenv) > 0) s->generic_cost = atof(cenv); }
                for (int32_t i = 0; i < pl.ntiles; i++) {
                    const int32_t t = order_s[i];
                    s->ent_prefix_s[(size_t)i + 1] = s->ent_prefix_s[i] + (pl.elem_off[(size_t)t + 1] - pl.elem_off[t]);
                    memcpy(&meta_s[(size_t)META_INTS * (size_t)i], &meta[(size_t)META_INTS * (size_t)pos[t]], META_INTS * sizeof(int32_t));
                    const int64_t ne = pl.elem_off[(size_t)t + 1] - pl.elem_off[t];
                    if (sok[t]) {
                        meta_s[(size_t)META_INTS * (size_t)i + 15] = 1;
                        for (int c = 0; c < 3; c++) tcoef[4 * (size_t)i + c] = coef[3 * (size_t)pl.elem_off[t] + c];
                        s->struct_entries += ne;
                    } else if (!pl.tile_self[t]) s->generic_entries += ne;
                }
                TRY(upload(s, &s->t_coef, tcoef.data(), tcoef.size()));
                TRY(upload(s, (int32_t **)&s->t_meta_s, meta_s.data(), meta_s.size()));
                TRY(dalloc(s, &s->d_queue, 2));
                { const char *denv = getenv("HGPU_DYNAMIC"); s->dynamic_tiles = denv && atoi(denv) == 1; }
            }
        }
        TRY(upload(s, (uint16_t **)&s->t_ent_slot, pl.elem_slot.data(), pl.elem_slot.size()));
        TRY(upload(s, &s->t_ent_coef, coef.data(), coef.size()));
        TRY(upload(s, &s->t_halo_id, pl.halo_id.data(), pl.halo_id.size()));
        if (bkt) {
            if (entries != (size_t)E) { hgpu_finalize(s); return fail(HGPU_EINVAL, "internal: BKT plan with extra entries"); }
            // record: c1, c2, then edata[4..13] = a0s a1s bs g0s g1s a0k a1k bk g0k g1k (psolve.h:95-97)
            std::vector<double> rec8(8 * entries, 0.0);
            std::vector<int32_t> eoe((size_t)E, -1);
            for (size_t k = 0; k < entries; k++) {
                const int32_t e = pl.elem_id[k];
                const double *et = mesh->eTable + 4 * (size_t)e;
                rec8[8 * k] = et[0]; rec8[8 * k + 1] = et[1];
                memcpy(&rec8[8 * k + 2], mesh->edata + 14 * (size_t)e + 4, 10 * sizeof(float));
                eoe[e] = (int32_t)k;
            }
            TRY(upload(s, &s->t_ent_bkt, rec8.data(), rec8.size()));
            TRY(upload(s, &s->entry_of_elem, eoe.data(), eoe.size()));
            const size_t nconv = ((entries + 31) / 32) * (size_t)CONV_PER_ENTRY * 32;
            TRY(dalloc(s, &s->conv, nconv));
            TRYCU(cudaMemset(s->conv, 0, std::max<size_t>(1, nconv) * sizeof(double)));   // calloc, psolve.c:3322-3325
        }
        {
            std::vector<uint2> rec(pl.rec.size());
            for (size_t i = 0; i < pl.rec.size(); i++)
                rec[i] = make_uint2((uint32_t)pl.rec[i].slot3 | ((uint32_t)pl.rec[i].cnt << 16) | ((uint32_t)pl.rec[i].flags << 24),
                                    (uint32_t)pl.rec[i].first);
            TRY(upload(s, &s->t_rec, rec.data(), rec.size()));
        }
        TRY(upload(s, &s->t_src, pl.src.data(), pl.src.size()));
        TRY(upload(s, &s->t_dep, pl.dep.data(), pl.dep.size()));
        TRY(dalloc(s, &s->t_partial, 3 * pl.halo_id.size()));
        TRYCU(cudaMemset(s->t_partial, 0, std::max<size_t>(1, 3 * pl.halo_id.size()) * sizeof(double)));
        TRY(dalloc(s, &s->t_flag, (size_t)pl.ntiles));
        TRYCU(cudaMemset(s->t_flag, 0, std::max<size_t>(1, (size_t)pl.ntiles) * sizeof(unsigned int)));
        // shared memory actually needed by this plan
        s->cap_slots = (pl.max_tile_nodes + 15) & ~15; s->cap_acc = (pl.max_tile_acc + 15) & ~15;
        s->cap_acc_struct = ((SP_TOTAL + SF_TOTAL + 2) / 3 + 15) & ~15;         // planes + element forces (STRUCT launches)
        s->cap_owned = (pl.max_tile_owned + 15) & ~15;
        s->cap_recs = std::min(caps.max_recs, (pl.max_tile_recs + 15) & ~15);
        s->cap_srcs = std::min(caps.max_srcs, (pl.max_tile_srcs + 15) & ~15);
        s->smem_u2 = step_smem_bytes(true, s->cap_slots, s->cap_acc, s->cap_owned, s->cap_recs, s->cap_srcs);
        s->smem_nou2 = step_smem_bytes(false, s->cap_slots, s->cap_acc, s->cap_owned, s->cap_recs, s->cap_srcs);
        s->smem_struct = step_smem_bytes(true, s->cap_slots, s->cap_acc_struct, s->cap_owned, s->cap_recs, s->cap_srcs);
        const char *benv = getenv("HGPU_BLOCK");
        if (benv && atoi(benv) == 384) s->block = 384;
        int occ = 0;
        // the attribute belongs to the kernel, not to this solver: always ask for the whole per-CTA
        // share so that solvers with different plans can coexist in one process
        const int smem_cap = std::min(max_smem, (max_smem + 1024) / 2 - 1024);
        if (s->smem_u2 > smem_cap) { hgpu_finalize(s); return fail(HGPU_EINVAL, "tile plan needs %d bytes of shared memory (> %d)", s->smem_u2, smem_cap); }
        if (s->n_struct > 0 && s->smem_struct > max_smem) s->n_struct = 0;
#define SETUP(T)                                                                                             \
        do {                                                                                                 \
            TRYCU(cudaFuncSetAttribute(step_kernel<0, false, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap)); \
            TRYCU(cudaFuncSetAttribute(step_kernel<1, false, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap)); \
            TRYCU(cudaFuncSetAttribute(step_kernel<2, false, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap)); \
            TRYCU(cudaFuncSetAttribute(step_kernel<0, true, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap));  \
            TRYCU(cudaFuncSetAttribute(step_kernel<1, true, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap));  \
            TRYCU(cudaFuncSetAttribute(step_kernel<2, true, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap));  \
            TRYCU(cudaFuncSetAttribute(step_kernel<3, false, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap)); \
            if (T == 256) TRYCU(cudaFuncSetAttribute(step_kernel<1, false, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap)); \
            TRYCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, step_kernel<1, false, T>, T, s->smem_u2));            \
        } while (0)
        if (s->block == 384) SETUP(384); else SETUP(256);
#undef SETUP
        if (s->block == 256 && (params->flags & HGPU_FLAG_WPASS)) {
            int occ_w = 0;
            TRYCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_w, step_kernel<1, false, 256, true>, 256, s->smem_u2));
            occ = std::min(occ, occ_w);
        }
        if (s->block != 256) s->n_struct = 0;     // STRUCT launches replace the 256-thread ones only
        if (s->n_struct > 0) {
            // STRUCT launches: one CTA per SM with (almost) the whole shared memory; the opt-in limit counts the
            // kernel's static shared memory too.  Any failure here only switches the structured path off.
            int occ_s = 0;
            cudaFuncAttributes fa0, fa1;
            bool ok = cudaFuncGetAttributes(&fa0, step_kernel<0, false, 512, false, true>) == cudaSuccess &&
                      cudaFuncGetAttributes(&fa1, step_kernel<1, false, 512, false, true>) == cudaSuccess;
            const int dyn_max = ok ? max_smem - (int)std::max(fa0.sharedSizeBytes, fa1.sharedSizeBytes) : 0;
            ok = ok && s->smem_struct <= dyn_max &&
                 cudaFuncSetAttribute(step_kernel<0, false, 512, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max) == cudaSuccess &&
                 cudaFuncSetAttribute(step_kernel<1, false, 512, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max) == cudaSuccess &&
                 cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, step_kernel<1, false, 512, false, true>, 512, s->smem_struct) == cudaSuccess &&
                 occ_s >= 1;
            if (!ok) { cudaGetLastError(); s->n_struct = 0; }
            s->grid_struct_max = nsm * std::max(occ_s, 1);
        }
        if (occ < 1) { hgpu_finalize(s); return fail(HGPU_EINVAL, "step kernel does not fit on an SM"); }
        {
            int coop = 0;
            TRYCU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s->dev));
            const char *cenv = getenv("HGPU_COOPERATIVE");
            s->cooperative = coop != 0 && !(cenv && atoi(cenv) == 0);
        }
        s->ctas_per_sm = occ;
        // every CTA of a launch must be resident: a tile's finish phase spins on flags raised by
        // CTAs of the same launch
        s->grid = std::max(1, std::min(pl.ntiles, nsm * occ));
        {
            int reserve = 1;
            const char *renv = getenv("HGPU_COMM_SMS");
            if (renv && atoi(renv) >= 0) reserve = atoi(renv);
            s->grid_late = std::max(1, (nsm - reserve) * occ);
        }
        const char *genv = getenv("HGPU_GRID");
        if (genv && atoi(genv) > 0) s->grid = std::min(std::min(pl.ntiles, atoi(genv)), nsm * occ);
    }
    TRYCU(cudaStreamSynchronize(s->stream));
    TRYCU(cudaDeviceSynchronize());
#undef TRY
#undef TRYCU
    *out = s;
    return HGPU_OK;
}

template <typename T>
static void dfree(T *&p) { if (p) cudaFree(p); p = nullptr; }

static void free_msglist(MsgList &m)
{
    dfree(m.d_map); dfree(m.d_send); dfree(m.d_recv); dfree(m.d_counters); dfree(m.d_push); dfree(m.d_pull);
    dfree(m.d_csr_node); dfree(m.d_csr_off); dfree(m.d_csr_src[0]); dfree(m.d_csr_src[1]);
}

extern "C" int hgpu_finalize(hgpu_solver_t *s)
{
    if (!s) return HGPU_OK;
    cudaSetDevice(s->dev);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
    for (size_t r = 0; r < s->peer_base.size(); r++) if (s->peer_base[r]) cudaIpcCloseMemHandle(s->peer_base[r]);
    dfree(s->mailbox);
    if (s->h_err) cudaFreeHost(s->h_err);
    for (int b = 0; b < 3; b++) dfree(s->u[b]);
    dfree(s->force); dfree(s->mass); dfree(s->m2); dfree(s->m1); dfree(s->nt3); dfree(s->etab); dfree(s->Kd);
    dfree(s->t_meta); dfree(s->t_ent_slot); dfree(s->t_ent_coef); dfree(s->t_halo_id); dfree(s->t_beta); dfree(s->t_coef); dfree(s->t_meta_s); dfree(s->d_queue);
    dfree(s->conv); dfree(s->t_ent_bkt); dfree(s->entry_of_elem); dfree(s->conv_scratch);
    dfree(s->t_rec); dfree(s->t_src); dfree(s->t_dep); dfree(s->t_partial); dfree(s->t_flag);
    dfree(s->d_slist); dfree(s->d_loaded); dfree(s->d_F); dfree(s->d_Fall); dfree(s->d_dnode);
    dfree(s->d_anchor_id); dfree(s->d_anchor_off); dfree(s->d_anchor_dn); dfree(s->d_anchor_deps);
    dfree(s->d_fetch_ids); dfree(s->d_fetch_out);
    for (auto &f : s->af) {
        if (f.busy && f.done) cudaEventSynchronize(f.done);
        dfree(f.snap);
        if (f.ready) cudaEventDestroy(f.ready);
        if (f.done) cudaEventDestroy(f.done);
    }
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); }
    dfree(s->d_st_nodes); dfree(s->d_st_local); dfree(s->d_st_rows);
    dfree(s->d_pl_nodes); dfree(s->d_pl_local);
    for (auto &q : s->pl) {
        dfree(q.rows);
        if (q.ready) cudaEventDestroy(q.ready);
        if (q.done) cudaEventDestroy(q.done);
    }
    free_msglist(s->dn_c); free_msglist(s->dn_s); free_msglist(s->an_c); free_msglist(s->an_s);
    for (int i = 0; i < hgpu_solver::SRC_RING; i++) if (s->src_done[i]) cudaEventDestroy(s->src_done[i]);
    for (EvPair &e : s->evpool) { if (e.a) cudaEventDestroy(e.a); if (e.b) cudaEventDestroy(e.b); }
    if (s->h_F) cudaFreeHost(s->h_F);
    if (s->ev_early) cudaEventDestroy(s->ev_early);
    if (s->ev_comm) cudaEventDestroy(s->ev_comm);
    if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return HGPU_OK;
}

// ---- force evaluation ---------------------------------------------------------------------------

// The force terms requested since the last update (hgpu_force_stiffness / hgpu_force_damping).
struct Terms { bool stiff, need_u2, bkt; bool any() const { return stiff || need_u2 || bkt; } };

static Terms consume_terms(hgpu_solver *s)
{
    Terms t;
    t.stiff = s->want_stiff;
    // MASS damping has b = 0, hence c3 = c4 = 0 (psolve.c:5866-5867): damping_addforce adds nothing
    t.need_u2 = s->want_damp && s->P.damping == HGPU_DAMPING_RAYLEIGH;
    // BKT: calc_conv + constant_Q_addforce, which carries the elastic term too (psolve.c:3969, 4003-4010)
    t.bkt = s->want_damp && s->P.damping == HGPU_DAMPING_BKT;
    s->want_stiff = s->want_damp = false;
    return t;
}

// Launch the step kernel over tiles [begin, end) of the processing order.
// fuse = advance REGULAR nodes in the same launch (their force never reaches HBM).
static int launch_range(hgpu_solver *s, Terms tm, bool fuse, int32_t begin, int32_t end, int max_grid = 0)
{
    if (end <= begin || !tm.any()) return HGPU_OK;
    StepArgs A{};
    A.u1 = s->u[s->i1]; A.u2 = s->u[s->i2]; A.unext = s->u[s->i3]; A.force = s->force;
    A.nt3 = s->nt3; A.Kd = s->Kd;
    A.tile_meta = s->t_meta; A.halo_id = s->t_halo_id;
    A.ent_slot = s->t_ent_slot; A.ent_coef = s->t_ent_coef;
    A.rec = s->t_rec; A.src = s->t_src; A.dep = s->t_dep; A.partial = s->t_partial; A.flag = s->t_flag;
    A.epoch = s->epoch;
    A.tile_begin = begin; A.ntiles = end; A.cap_slots = s->cap_slots; A.cap_acc = s->cap_acc; A.cap_owned = s->cap_owned;
    A.cap_recs = s->cap_recs; A.cap_srcs = s->cap_srcs;
    A.fuse_update = fuse ? 1 : 0;
    A.conv = s->conv; A.ent_bkt = s->t_ent_bkt;
    A.rmax = 2.0 * M_PI * s->P.freq * s->P.dt;                     // damping.c:114, 234
    A.tile_beta = s->t_beta;
    A.tile_coef = s->t_coef;
    A.err = s->d_err;
    // compute_addforce_conventional (stiffness.c:121-176) applies theK1 x -c1 and theK2 x -c2 as dense 24 x 24
    // matrices; compute_addforce_effective (:180-237) applies the SAME operator in factored form (the reference's
    // two methods differ by rounding only: 5e-16 rel L2 over a whole run, oracle on the conventional golden).
    // A solver created with HGPU_STIFFNESS_CONVENTIONAL therefore runs the factored kernels too (7x faster:
    // profiles/r02k_*); HGPU_FLAG_DENSE_K asks for the literal dense evaluation with the K1 / K2 handed over.
    const bool dense = s->P.stiffness == HGPU_STIFFNESS_CONVENTIONAL && !tm.bkt && (s->P.flags & HGPU_FLAG_DENSE_K);
    const int mode = tm.bkt ? 3 : tm.stiff ? (tm.need_u2 ? 1 : 0) : 2;
    // fused launches have no counterpart among the reference's timers; an unfused launch is
    // booked under "Compute addforces e" when it carries the stiffness term, else under damping
    PhaseTimer pt(s, fuse ? PH_FUSED_STEP : (tm.stiff ? PH_ADDFORCE_E : PH_DAMPING));
    // multi-rank: never ask for every CTA slot of the device (see the launch below)
    if (s->P.nranks > 1) max_grid = max_grid > 0 ? std::min(max_grid, s->grid_late) : s->grid_late;
    int G = std::min(max_grid > 0 ? std::min(max_grid, s->grid) : s->grid, end - begin);
    if (begin == 0) A.epoch = ++s->epoch;       // a new pass over the tiles (a split pass shares one epoch)
    // STRUCT launch: the range is taken from the second processing order (structured tiles last); its
    // structured part goes to CTAs [0, grid_struct), the slot-table part to the others, in proportion to
    // the elements of either kind weighted by their relative cost
    bool use_struct = fuse && !dense && (mode == 0 || mode == 1) && s->block == 256 && s->n_struct > 0 &&
                      !(s->P.flags & HGPU_FLAG_WPASS) && end > s->n_generic;
    int B = s->block;
    if (use_struct) {
        const int32_t gb = std::min(begin, s->n_generic), ge = s->n_generic, sb = std::max(begin, s->n_generic), se = end;
        const int32_t ng = ge - gb, ns = se - sb;
        const double wg = s->generic_cost * (double)(s->ent_prefix_s[ge] - s->ent_prefix_s[gb]);
        const double ws = (double)(s->ent_prefix_s[se] - s->ent_prefix_s[sb]);
        // one CTA of 512 threads per SM (multi-rank: the communication SMs stay free)
        B = 512;
        A.cap_acc = s->cap_acc_struct;
        const int Gmax = s->P.nranks > 1 ? std::max(1, s->grid_struct_max - (s->grid - s->grid_late) / std::max(1, s->ctas_per_sm))
                                         : s->grid_struct_max;
        int Gs, Gg;
        if (ng == 0) { Gs = std::min(Gmax, ns); Gg = 0; }
        else {
            Gs = (int)std::lround((double)Gmax * ws / (ws + wg));
            Gs = std::max(1, std::min(Gs, std::min(ns, Gmax - 1)));
            Gg = std::min(Gmax - Gs, ng);
        }
        G = Gs + Gg;
        if (s->dynamic_tiles && ng > 0) {
            // tiles handed out from two counters: every CTA the device holds takes part, the split only says
            // which list a CTA starts on
            G = std::min(Gmax, ng + ns);
            Gs = std::max(1, std::min(G - 1, (int)std::lround((double)G * ws / (ws + wg))));
            CK(cudaMemsetAsync(s->d_queue, 0, 2 * sizeof(int), s->stream));
            A.queue = s->d_queue;
        }
        A.tile_meta = s->t_meta_s;
        A.tile_begin = gb; A.ntiles = ge; A.struct_begin = sb; A.struct_end = se; A.grid_struct = Gs;
    }
    // Cooperative launch (single rank): a tile's finish spins on flags raised by other CTAs of the same launch,
    // so every CTA must be resident -- the driver then either co-schedules the whole grid or fails the launch
    // (it never starts a part of it), whatever else shares the device.
    // NOT with several ranks: a cooperative grid that needs every slot of the device cannot start while ANY
    // other kernel is resident, and a kernel of another library that spins for its peer -- the NCCL kernel of a
    // host-side barrier -- closes a cycle: this rank's steps wait for that kernel, the peer's steps wait for this
    // rank's halo, the peer's host (launch queue full) never reaches its barrier (found at 120+ queued steps on
    // 2 GPUs, r02 call 5c).  Multi-rank launches are ordinary ones that leave the communication SM free, so
    // the grid fits beside the halo kernels and a foreign kernel or two; the bounded wait in wait_deps reports
    // the case where it still does not.
    const void *fn = nullptr;
    size_t smem = (size_t)s->smem_u2;
#define PICK(T)                                                                                   \
    do {                                                                                          \
        if (dense) {                                                                              \
            if (mode == 0)      { fn = (const void *)step_kernel<0, true, T>; smem = (size_t)s->smem_nou2; } \
            else if (mode == 1) fn = (const void *)step_kernel<1, true, T>;                       \
            else                fn = (const void *)step_kernel<2, true, T>;                       \
        } else if (mode == 3) {                                                                   \
            fn = (const void *)step_kernel<3, false, T>;                                          \
        } else {                                                                                  \
            if (mode == 0)      { fn = (const void *)step_kernel<0, false, T>; smem = (size_t)s->smem_nou2; } \
            else if (mode == 1) fn = (const void *)step_kernel<1, false, T>;                      \
            else                fn = (const void *)step_kernel<2, false, T>;                      \
        }                                                                                         \
    } while (0)
    if (fuse && !dense && mode == 1 && B == 256 && (s->P.flags & HGPU_FLAG_WPASS))
        fn = (const void *)step_kernel<1, false, 256, true>;       // opt-in variant, see hgpu_kernels.cuh
    else if (use_struct) {
        // structured tiles take their own path on their own CTAs, the others the slot-table path, in one launch
        smem = (size_t)s->smem_struct;
        if (mode == 0) fn = (const void *)step_kernel<0, false, 512, false, true>;
        else fn = (const void *)step_kernel<1, false, 512, false, true>;
    }
    else if (B == 384) PICK(384); else PICK(256);
#undef PICK
    {
        void *kargs[] = {(void *)&A};
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)G); cfg.blockDim = dim3((unsigned)B);
        cfg.dynamicSmemBytes = smem; cfg.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = (s->cooperative && s->P.nranks == 1) ? 1 : 0;
        CK(cudaLaunchKernelExC(&cfg, fn, kargs));
    }
    CK(cudaGetLastError());
    s->tm.launches++;
    return HGPU_OK;
}

// All tiles in one launch.  *launched: whether a kernel ran (no term requested = nothing to add).
static int launch_tiles(hgpu_solver *s, bool fuse, bool *launched)
{
    const Terms tm = consume_terms(s);
    *launched = tm.any() && s->plan.ntiles > 0;
    return launch_range(s, tm, fuse, 0, s->plan.ntiles);
}

// Make force[] hold the sum of every requested term for EVERY node (unfused semantics).
static int materialize_forces(hgpu_solver *s)
{
    if (s->fstate == F_PENDING) {
        // temporarily treat all nodes as SPECIAL: run without the fused update
        bool ran;
        int rc = launch_tiles(s, false, &ran);
        if (rc) return rc;
        s->fstate = F_MATERIALIZED;
    }
    return HGPU_OK;
}

extern "C" int hgpu_step_begin(hgpu_solver_t *s, int32_t step)
{
    (void)step;
    if (!s) return fail(HGPU_EINVAL, "null solver");
    if (s->fstate == F_PENDING || s->fstate == F_FUSED_DONE)
        return fail(HGPU_ESTATE, "hgpu_step_begin: previous step's forces were never consumed by hgpu_update");
    std::swap(s->i1, s->i2);   // psolve.c:4271-4273
    s->tail_done = false;
    return HGPU_OK;
}

extern "C" int hgpu_force_source(hgpu_solver_t *s, const double *F)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    const int n = s->P.nloaded;
    if (n == 0) return HGPU_OK;
    if (!F) return fail(HGPU_EINVAL, "hgpu_force_source: F is null but nloaded > 0");
    if (s->fstate != F_CLEAN)
        return fail(HGPU_ESTATE, "hgpu_force_source must precede the element forces (assignment, psolve.c:5921)");
    CK(cudaSetDevice(s->dev));
    // ring slot: wait only for the copy issued SRC_RING steps ago (never recorded = ready)
    const int slot = s->src_slot;
    s->src_slot = (slot + 1) % hgpu_solver::SRC_RING;
    CK(cudaEventSynchronize(s->src_done[slot]));
    double *hF = s->h_F + (size_t)slot * 3 * (size_t)n, *dF = s->d_F + (size_t)slot * 3 * (size_t)n;
    memcpy(hF, F, 3 * (size_t)n * sizeof(double));
    PhaseTimer pt(s, PH_ADDFORCE_S);
    CK(cudaMemcpyAsync(dF, hF, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    source_kernel<<<grid_for(3LL * n, 128), 128, 0, s->stream>>>(n, s->d_loaded, dF, s->P.dt2, s->force);
    CK(cudaGetLastError());
    CK(cudaEventRecord(s->src_done[slot], s->stream));
    s->tm.launches++;
    return HGPU_OK;
}

extern "C" int hgpu_force_stiffness(hgpu_solver_t *s)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    if (s->P.damping == HGPU_DAMPING_BKT) return HGPU_OK;   // psolve.c:3969
    if (s->fstate == F_FUSED_DONE) return fail(HGPU_ESTATE, "forces already consumed for this step");
    CK(cudaSetDevice(s->dev));
    if (s->fstate == F_MATERIALIZED) {
        bool ran;
        s->want_stiff = true;
        return launch_tiles(s, false, &ran);
    }
    s->want_stiff = true;
    s->fstate = F_PENDING;
    return HGPU_OK;
}

extern "C" int hgpu_force_damping(hgpu_solver_t *s)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    if (s->P.damping == HGPU_DAMPING_NONE) return HGPU_OK;   // psolve.c:3999-4000
    if (s->fstate == F_FUSED_DONE) return fail(HGPU_ESTATE, "forces already consumed for this step");
    CK(cudaSetDevice(s->dev));
    if (s->fstate == F_MATERIALIZED) {
        bool ran;
        s->want_damp = true;
        return launch_tiles(s, false, &ran);
    }
    s->want_damp = true;
    s->fstate = F_PENDING;
    return HGPU_OK;
}

// schedule_senddata (psolve.c:4945-5079) for one schedule and one direction.
//   contribution: c-list packs v -> owner; owner's s-list unpacks with += (one messenger after another)
//   sharing     : s-list packs v -> sharers; c-list unpacks with =
static int exchange(hgpu_solver *s, MsgList &c, MsgList &sl, double *v, bool contribution, cudaStream_t st)
{
    // blocks per messenger: the force exchange runs beside the late tiles on the one or two SMs
    // left free for it; the displacement exchange has the whole GPU to itself
    const int max_gx = (st == s->stream) ? 64 : 8;
    if (s->P.nranks == 1 || (c.total == 0 && sl.total == 0)) return HGPU_OK;
    MsgList &snd = contribution ? c : sl;
    MsgList &rcv = contribution ? sl : c;
    if (s->p2p_ready) {
        const unsigned long long so = ++snd.seq_out, si = ++rcv.seq_in;
        const int nsnd = (int)snd.peer.size(), nrcv = (int)rcv.peer.size();
        if (nsnd) {
            int maxn = 1;
            for (int32_t n : snd.nodes) maxn = std::max(maxn, n);
            const int gx = std::max(1, std::min(max_gx, (3 * maxn + 255) / 256));
            p2p_push_kernel<<<dim3(gx, nsnd), 256, 0, st>>>(snd.d_push + (so & 1) * nsnd, v, so);
            CK(cudaGetLastError());
            s->tm.launches++;
        }
        if (nrcv && contribution) {
            // every messenger of the list in ONE launch: a thread owns a (node, component) and adds the
            // node's contributions in messenger order (psolve.c:5035-5073)
            PullAll pa{rcv.d_csr_node, rcv.d_csr_off, rcv.d_csr_src[si & 1],
                       (const double *)(s->mailbox + rcv.mb_data_off),
                       (const unsigned long long *)(s->mailbox + rcv.mb_flag_off), rcv.csr_nodes, nrcv, (int32_t)(si & 1)};
            const int gx = std::max(1, std::min(max_gx, (3 * rcv.csr_nodes + 255) / 256));
            p2p_pull_all_kernel<<<gx, 256, 0, st>>>(pa, v, si, s->d_err, s->p2p_timeout_cycles);
            CK(cudaGetLastError());
            s->tm.launches++;
        } else if (nrcv) {
            int maxn = 1;
            for (int32_t n : rcv.nodes) maxn = std::max(maxn, n);
            const int gx = std::max(1, std::min(max_gx, (3 * maxn + 255) / 256));
            p2p_pull_kernel<<<dim3(gx, nrcv), 256, 0, st>>>(rcv.d_pull + (si & 1) * nrcv, v, si, 0, s->d_err, s->p2p_timeout_cycles);
            CK(cudaGetLastError());
            s->tm.launches++;
        }
        return HGPU_OK;
    }
    if (!s->comm) return fail(HGPU_ECOMM, "halo exchange needs hgpu_comm_init or hgpu_comm_p2p_connect first");
    if (snd.total) {
        pack_kernel<<<grid_for(3LL * snd.total, 256), 256, 0, st>>>(snd.total, snd.d_map, v, snd.d_send);
        CK(cudaGetLastError());
        s->tm.launches++;
    }
    NK(g_nccl.GroupStart());
    for (size_t i = 0; i < rcv.peer.size(); i++)
        NK(g_nccl.Recv(rcv.d_recv + 3 * (size_t)rcv.off[i], 3 * (size_t)rcv.nodes[i], ncclFloat64, rcv.peer[i], s->comm, st));
    for (size_t i = 0; i < snd.peer.size(); i++)
        NK(g_nccl.Send(snd.d_send + 3 * (size_t)snd.off[i], 3 * (size_t)snd.nodes[i], ncclFloat64, snd.peer[i], s->comm, st));
    NK(g_nccl.GroupEnd());
    if (contribution) {
        // messengers applied one after another, as the reference's unpack loop does
        for (size_t i = 0; i < rcv.peer.size(); i++) {
            if (!rcv.nodes[i]) continue;
            unpack_kernel<<<grid_for(3LL * rcv.nodes[i], 256), 256, 0, st>>>(
                rcv.nodes[i], rcv.d_map + rcv.off[i], rcv.d_recv + 3 * (size_t)rcv.off[i], v, 1);
            CK(cudaGetLastError());
            s->tm.launches++;
        }
    } else if (rcv.total) {
        // a harbored node has exactly one owner, so the overwrite lists are disjoint
        unpack_kernel<<<grid_for(3LL * rcv.total, 256), 256, 0, st>>>(rcv.total, rcv.d_map, rcv.d_recv, v, 0);
        CK(cudaGetLastError());
        s->tm.launches++;
    }
    return HGPU_OK;
}

// phases 8-10 on stream st
static int force_phases(hgpu_solver *s, cudaStream_t st)
{
    int rc;
    // phase 8: dangling-node forces to their owners
    {
        PhaseTimer pt(s, PH_SEND_DN_FORCE, st);
        if ((rc = exchange(s, s->dn_c, s->dn_s, s->force, true, st))) return rc;
    }
    // phase 9: owned dangling nodes hand force/deps to their anchors
    if (s->nA > 0) {
        PhaseTimer pt(s, PH_ADJUST_FORCE, st);
        adjust_dist_kernel<<<grid_for(3LL * s->nA, 128), 128, 0, st>>>(
            s->nA, s->d_anchor_id, s->d_anchor_off, s->d_anchor_dn, s->d_anchor_deps, s->force);
        CK(cudaGetLastError());
        s->tm.launches++;
    }
    // phase 10: anchored-node forces to their owners
    PhaseTimer pt(s, PH_SEND_AN_FORCE, st);
    if ((rc = exchange(s, s->an_c, s->an_s, s->force, true, st))) return rc;
    return HGPU_OK;
}

// phases 13-15 on stream st; tm2 = the array holding the new displacements
static int disp_phases(hgpu_solver *s, double *tm2, cudaStream_t st)
{
    int rc;
    // phase 13: owners publish anchored-node displacements
    {
        PhaseTimer pt(s, PH_SEND_AN_DISP, st);
        if ((rc = exchange(s, s->an_c, s->an_s, tm2, false, st))) return rc;
    }
    // phase 14: dangling nodes interpolate from their anchors
    if (s->D > 0) {
        PhaseTimer pt(s, PH_ADJUST_DISP, st);
        adjust_asgn_kernel<<<grid_for(3LL * s->D, 128), 128, 0, st>>>(s->D, s->d_dnode, tm2);
        CK(cudaGetLastError());
        s->tm.launches++;
    }
    // phase 15: owners publish dangling-node displacements
    PhaseTimer pt(s, PH_SEND_DN_DISP, st);
    if ((rc = exchange(s, s->dn_c, s->dn_s, tm2, false, st))) return rc;
    return HGPU_OK;
}

extern "C" int hgpu_force_exchange(hgpu_solver_t *s)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    CK(cudaSetDevice(s->dev));
    int rc;
    if (s->fstate == F_PENDING) {
        const bool fuse = !(s->P.flags & HGPU_FLAG_NO_FUSE);
        const Terms tm = consume_terms(s);
        const int32_t T = s->plan.ntiles, Te = s->n_early;
        const bool ran = tm.any() && T > 0;
        s->fstate = (fuse && ran) ? F_FUSED_DONE : F_MATERIALIZED;
        if (ran && s->P.nranks > 1 && s->comm_stream && Te > 0 && Te < T && !(s->P.flags & HGPU_FLAG_NO_OVERLAP)) {
            // tiles owning nodes of the exchange / hanging-node phases first; those phases then run
            // on the communication stream while the remaining tiles are evaluated
            if ((rc = launch_range(s, tm, fuse, 0, Te))) return rc;
            CK(cudaEventRecord(s->ev_early, s->stream));
            CK(cudaStreamWaitEvent(s->comm_stream, s->ev_early, 0));
            if ((rc = force_phases(s, s->comm_stream))) return rc;
            if (fuse && (s->P.flags & HGPU_FLAG_TAIL_OVERLAP)) {
                // Every node the displacement phases read or write is owned by a self tile, whose
                // forces are final here: advance those nodes and run phases 13-15 on this stream too.
                // They write only the array of the NEW displacements (u[i3]) at nodes no late tile
                // owns, while the late tiles read u[i1], u[i2].
                double *un = s->u[s->i3];
                if (s->nS_early > 0) {
                    PhaseTimer pt(s, PH_NEW_DISP, s->comm_stream);
                    update_list_kernel<<<grid_for(3LL * s->nS_early, 256), 256, 0, s->comm_stream>>>(
                        s->nS_early, s->d_slist, s->u[s->i1], s->u[s->i2], un, s->force, s->mass, s->m2, s->m1);
                    CK(cudaGetLastError());
                    s->tm.launches++;
                }
                if ((rc = disp_phases(s, un, s->comm_stream))) return rc;
                s->tail_done = true;
            }
            CK(cudaEventRecord(s->ev_comm, s->comm_stream));
            // the step kernel fills every SM's registers: leave a few SMs to the exchange kernels
            if ((rc = launch_range(s, tm, fuse, Te, T, s->grid_late))) return rc;
            CK(cudaStreamWaitEvent(s->stream, s->ev_comm, 0));
            return HGPU_OK;
        }
        if ((rc = launch_range(s, tm, fuse, 0, T))) return rc;
    }
    return force_phases(s, s->stream);
}

extern "C" int hgpu_update(hgpu_solver_t *s)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    CK(cudaSetDevice(s->dev));
    int rc;
    if (s->fstate == F_PENDING) {
        // hgpu_force_exchange was skipped (legal on a conforming single-rank mesh)
        const bool fuse = !(s->P.flags & HGPU_FLAG_NO_FUSE);
        bool ran;
        if ((rc = launch_tiles(s, fuse, &ran))) return rc;
        s->fstate = (fuse && ran) ? F_FUSED_DONE : F_MATERIALIZED;
    }
    double *u1 = s->u[s->i1], *u2 = s->u[s->i2], *un = s->u[s->i3];
    PhaseTimer pt(s, PH_NEW_DISP);
    if (s->fstate == F_FUSED_DONE) {
        const int32_t first = s->tail_done ? s->nS_early : 0;     // the self tiles' nodes are done already
        if (s->nS > first) {
            update_list_kernel<<<grid_for(3LL * (s->nS - first), 256), 256, 0, s->stream>>>(
                s->nS - first, s->d_slist + first, u1, u2, un, s->force, s->mass, s->m2, s->m1);
            CK(cudaGetLastError());
            s->tm.launches++;
        }
    } else {
        // F_MATERIALIZED, or F_CLEAN (a step with no element force at all): every node from force[]
        const long long n3 = 3LL * s->N;
        int grid = (int)std::min<long long>((n3 + 255) / 256, 148LL * 16);
        update_all_kernel<<<grid, 256, 0, s->stream>>>(n3, u1, u2, un, s->force, s->mass, s->m2, s->m1);
        CK(cudaGetLastError());
        s->tm.launches++;
    }
    // roles after the update: tm1 unchanged, tm2 = new displacement, tm3 = old tm2 (psolve.c:4094-4106)
    const int old2 = s->i2;
    s->i2 = s->i3;
    s->i3 = old2;
    s->fstate = F_CLEAN;
    s->tm.steps++;
    return HGPU_OK;
}

extern "C" int hgpu_disp_exchange(hgpu_solver_t *s)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    CK(cudaSetDevice(s->dev));
    if (s->tail_done) { s->tail_done = false; return HGPU_OK; }   // ran beside the late tiles of this step
    return disp_phases(s, s->u[s->i2], s->stream);
}

extern "C" int hgpu_step(hgpu_solver_t *s, int32_t step, const double *F)
{
    int rc;
    if ((rc = hgpu_step_begin(s, step))) return rc;
    if ((rc = hgpu_force_source(s, F))) return rc;
    if ((rc = hgpu_force_stiffness(s))) return rc;
    if ((rc = hgpu_force_damping(s))) return rc;
    if ((rc = hgpu_force_exchange(s))) return rc;
    if ((rc = hgpu_update(s))) return rc;
    return hgpu_disp_exchange(s);
}

extern "C" int hgpu_source_preload(hgpu_solver_t *s, int32_t step0, int32_t nsteps, const double *F_all)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    if (nsteps < 0) return fail(HGPU_EINVAL, "hgpu_source_preload: negative step count");
    const int n = s->P.nloaded;
    if (n == 0 || nsteps == 0) return HGPU_OK;
    if (!F_all) return fail(HGPU_EINVAL, "hgpu_source_preload: F_all is null but nloaded > 0");
    CK(cudaSetDevice(s->dev));
    // whole source history resident in HBM: no per-step host traffic
    if (s->Fall_steps < (size_t)nsteps) {
        CK(cudaStreamSynchronize(s->stream));
        dfree(s->d_Fall);
        int rc = dalloc(s, &s->d_Fall, 3 * (size_t)n * (size_t)nsteps);
        if (rc) return rc;
        s->Fall_steps = (size_t)nsteps;
    }
    CK(cudaMemcpyAsync(s->d_Fall, F_all, 3 * (size_t)n * (size_t)nsteps * sizeof(double),
                       cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));     // F_all is the caller's pageable memory
    s->Fall_loaded = (size_t)nsteps;
    s->Fall_step0 = step0;
    return HGPU_OK;
}

extern "C" int hgpu_force_source_resident(hgpu_solver_t *s, int32_t step)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    const int n = s->P.nloaded;
    if (n == 0) return HGPU_OK;
    if (s->fstate != F_CLEAN)
        return fail(HGPU_ESTATE, "hgpu_force_source_resident must precede the element forces (assignment, psolve.c:5921)");
    if (step < s->Fall_step0 || (size_t)(step - s->Fall_step0) >= s->Fall_loaded)
        return fail(HGPU_EINVAL, "hgpu_force_source_resident: step %d is not covered by the preloaded source history [%d, %d)",
                    step, s->Fall_step0, s->Fall_step0 + (int32_t)s->Fall_loaded);
    CK(cudaSetDevice(s->dev));
    PhaseTimer pt(s, PH_ADDFORCE_S);
    source_kernel<<<grid_for(3LL * n, 128), 128, 0, s->stream>>>(
        n, s->d_loaded, s->d_Fall + 3 * (size_t)n * (size_t)(step - s->Fall_step0), s->P.dt2, s->force);
    CK(cudaGetLastError());
    s->tm.launches++;
    return HGPU_OK;
}

extern "C" int hgpu_run(hgpu_solver_t *s, int32_t step0, int32_t nsteps, const double *F_all)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    if (nsteps < 0) return fail(HGPU_EINVAL, "hgpu_run: negative step count");
    const int n = s->P.nloaded;
    CK(cudaSetDevice(s->dev));
    if (n > 0 && F_all) {
        int rc = hgpu_source_preload(s, step0, nsteps, F_all);
        if (rc) return rc;
    } else if (n > 0 && (step0 < s->Fall_step0 ||
                         (size_t)(step0 - s->Fall_step0) + (size_t)nsteps > s->Fall_loaded))
        return fail(HGPU_EINVAL, "hgpu_run: steps [%d, %d) are not covered by the preloaded source history",
                    step0, step0 + nsteps);
    const size_t row0 = n > 0 ? (size_t)(step0 - s->Fall_step0) : 0;
    if (s->st_n > 0 && s->st_rate > 0) {
        // the station ring must take every row of this run: fail BEFORE the first swap, not in the middle of a step
        int64_t rows = 0;
        for (int32_t k = 0; k < nsteps; k++) rows += (step0 + k) % s->st_rate == 0;
        if (s->st_count + rows > s->st_cap)
            return fail(HGPU_ESTATE, "hgpu_run: %lld station rows would be recorded but the device ring has room for %d "
                        "(capacity %d, %d pending): drain it or run fewer steps", (long long)rows, s->st_cap - s->st_count,
                        s->st_cap, s->st_count);
    }
    for (int32_t k = 0; k < nsteps; k++) {
        int rc;
        if ((rc = hgpu_step_begin(s, step0 + k))) return rc;
        if (s->st_n > 0 && s->st_rate > 0 && (step0 + k) % s->st_rate == 0)
            if ((rc = hgpu_stations_record(s, step0 + k))) return rc;       // solver_output_stations, psolve.c:4281
        if (n > 0) {
            PhaseTimer pt(s, PH_ADDFORCE_S);
            source_kernel<<<grid_for(3LL * n, 128), 128, 0, s->stream>>>(
                n, s->d_loaded, s->d_Fall + 3 * (size_t)n * (row0 + (size_t)k), s->P.dt2, s->force);
            CK(cudaGetLastError());
            s->tm.launches++;
        }
        if ((rc = hgpu_force_stiffness(s))) return rc;
        if ((rc = hgpu_force_damping(s))) return rc;
        if ((rc = hgpu_force_exchange(s))) return rc;
        if ((rc = hgpu_update(s))) return rc;
        if ((rc = hgpu_disp_exchange(s))) return rc;
    }
    return HGPU_OK;
}

// ---- taps ---------------------------------------------------------------------------------------

static int resolve(hgpu_solver *s, int32_t which, double **p, size_t *count)
{
    *count = 3 * (size_t)s->N;
    switch (which) {
    case HGPU_TM1: *p = s->u[s->i1]; return HGPU_OK;
    case HGPU_TM2: *p = s->u[s->i2]; return HGPU_OK;
    case HGPU_TM3: *p = s->u[s->i3]; return HGPU_OK;
    case HGPU_FORCE: {
        int rc = materialize_forces(s);
        if (rc) return rc;
        if (s->fstate == F_FUSED_DONE)
            return fail(HGPU_ESTATE, "force of REGULAR nodes is not kept by the fused step; "
                                     "read it before hgpu_force_exchange or use HGPU_FLAG_NO_FUSE");
        *p = s->force; return HGPU_OK;
    }
    case HGPU_CONV_SHEAR_1: case HGPU_CONV_SHEAR_2: case HGPU_CONV_KAPPA_1: case HGPU_CONV_KAPPA_2: {
        if (!s->conv) return fail(HGPU_EINVAL, "conv arrays exist with BKT damping only");
        *count = 24 * (size_t)s->E;
        if (!s->conv_scratch) { int rc = dalloc(s, &s->conv_scratch, *count); if (rc) return rc; }
        *p = s->conv_scratch;       // the caller converts between the layouts (conv_to_ref / conv_from_ref)
        return HGPU_OK;
    }
    default: return fail(HGPU_EINVAL, "unknown array selector %d", which);
    }
}

static bool is_conv(int32_t which) { return which >= HGPU_CONV_SHEAR_1 && which <= HGPU_CONV_KAPPA_2; }
// reference layout <-> entry-chunked device layout of one conv array, through conv_scratch
static int conv_convert(hgpu_solver *s, int32_t which, int to_ref)
{
    const int k0 = (which - HGPU_CONV_SHEAR_1) * 24;        // shear_1, shear_2, kappa_1, kappa_2
    if (s->E == 0) return HGPU_OK;
    conv_convert_kernel<<<grid_for(24LL * s->E, 256), 256, 0, s->stream>>>(s->E, s->entry_of_elem, k0, s->conv,
                                                                            s->conv_scratch, to_ref);
    CK(cudaGetLastError());
    s->tm.launches++;
    return HGPU_OK;
}

extern "C" int hgpu_fetch_all(hgpu_solver_t *s, int32_t which, double *out)
{
    if (!s || !out) return fail(HGPU_EINVAL, "null argument");
    CK(cudaSetDevice(s->dev));
    double *p; size_t cnt;
    int rc = resolve(s, which, &p, &cnt);
    if (rc) return rc;
    if (is_conv(which) && (rc = conv_convert(s, which, 1))) return rc;
    CK(cudaMemcpyAsync(out, p, cnt * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return check_device_error(s);       // never hand out a field computed past a failed exchange
}

// Asynchronous whole-field read (checkpoints io_checkpoint.c:29-117, 4D output output.c:1233):
// the array is snapshotted on the device in stream order (a device-to-device copy: ~0.1 ms per 400 MB), the
// time loop goes on, and the snapshot travels to the host on a separate copy stream.  Up to two reads can be
// in flight (tm1 and tm2 of one checkpoint); a third waits for the older one.
extern "C" int hgpu_fetch_all_async(hgpu_solver_t *s, int32_t which, double *out)
{
    if (!s || !out) return fail(HGPU_EINVAL, "null argument");
    if (which != HGPU_TM1 && which != HGPU_TM2 && which != HGPU_TM3)
        return fail(HGPU_EINVAL, "hgpu_fetch_all_async reads displacement arrays only");
    CK(cudaSetDevice(s->dev));
    double *p; size_t cnt;
    int rc = resolve(s, which, &p, &cnt);
    if (rc) return rc;
    if (!s->copy_stream) CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    hgpu_solver::AsyncFetch *f = !s->af[0].busy ? &s->af[0] : !s->af[1].busy ? &s->af[1] : nullptr;
    if (!f) {                                   // both in flight: the older one (slot 0) first
        CK(cudaEventSynchronize(s->af[0].done));
        s->af[0].busy = false;
        f = &s->af[0];
    }
    if (f->cap < cnt) {
        dfree(f->snap);
        if ((rc = dalloc(s, &f->snap, cnt))) return rc;
        f->cap = cnt;
    }
    if (!f->ready) { CK(cudaEventCreateWithFlags(&f->ready, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&f->done, cudaEventDisableTiming)); }
    CK(cudaMemcpyAsync(f->snap, p, cnt * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    CK(cudaEventRecord(f->ready, s->stream));
    CK(cudaStreamWaitEvent(s->copy_stream, f->ready, 0));
    CK(cudaMemcpyAsync(out, f->snap, cnt * sizeof(double), cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaEventRecord(f->done, s->copy_stream));
    f->busy = true;
    return HGPU_OK;
}

// Waits until every asynchronous read has landed in its host buffer.  Touches only events, so it may be
// called from a writer thread while the solver's own thread keeps enqueueing steps.
extern "C" int hgpu_fetch_wait(hgpu_solver_t *s)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    for (auto &f : s->af)
        if (f.busy) {
            CK(cudaEventSynchronize(f.done));
            f.busy = false;
        }
    return check_device_error(s);
}

extern "C" int hgpu_store_all(hgpu_solver_t *s, int32_t which, const double *in)
{
    if (!s || !in) return fail(HGPU_EINVAL, "null argument");
    CK(cudaSetDevice(s->dev));
    double *p; size_t cnt;
    int rc = resolve(s, which, &p, &cnt);
    if (rc) return rc;
    CK(cudaMemcpyAsync(p, in, cnt * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (is_conv(which) && (rc = conv_convert(s, which, 0))) return rc;
    CK(cudaStreamSynchronize(s->stream));
    return HGPU_OK;
}

extern "C" int hgpu_fetch_nodes(hgpu_solver_t *s, int32_t which, const int32_t *lnid, int32_t n, double *out)
{
    if (!s || (n > 0 && (!lnid || !out)) || n < 0) return fail(HGPU_EINVAL, "bad argument");
    if (n == 0) return HGPU_OK;
    if (is_conv(which)) return fail(HGPU_EINVAL, "hgpu_fetch_nodes addresses node arrays only");
    CK(cudaSetDevice(s->dev));
    double *p; size_t cnt;
    int rc = resolve(s, which, &p, &cnt);
    if (rc) return rc;
    for (int32_t i = 0; i < n; i++)
        if (lnid[i] < 0 || lnid[i] >= s->N) return fail(HGPU_EINVAL, "node id out of range");
    if (n > s->fetch_cap) {
        dfree(s->d_fetch_ids); dfree(s->d_fetch_out);
        if ((rc = dalloc(s, &s->d_fetch_ids, (size_t)n))) return rc;
        if ((rc = dalloc(s, &s->d_fetch_out, 3 * (size_t)n))) return rc;
        s->fetch_cap = n;
        s->fetch_ids_host.clear();
    }
    // stations and planes ask for the same nodes at every output step (psolve.c:6680, io_planes.c:151):
    // the list travels only when it changes
    if ((size_t)n != s->fetch_ids_host.size() || memcmp(s->fetch_ids_host.data(), lnid, (size_t)n * sizeof(int32_t)) != 0) {
        s->fetch_ids_host.assign(lnid, lnid + n);
        CK(cudaMemcpyAsync(s->d_fetch_ids, s->fetch_ids_host.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    }
    gather_nodes_kernel<<<grid_for(3LL * n, 128), 128, 0, s->stream>>>(n, s->d_fetch_ids, p, s->d_fetch_out);
    CK(cudaGetLastError());
    s->tm.launches++;
    CK(cudaMemcpyAsync(out, s->d_fetch_out, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return check_device_error(s);
}

// ---- stations on the device ---------------------------------------------------------

extern "C" int hgpu_stations_attach(hgpu_solver_t *s, int32_t nstations, const int32_t *nodes, const double *localcoords,
                                    int32_t print_vel, int32_t print_acc, int32_t rate, int32_t capacity)
{
    if (!s || nstations < 0 || (nstations > 0 && (!nodes || !localcoords)) || rate < 0 || capacity < 1)
        return fail(HGPU_EINVAL, "hgpu_stations_attach: bad argument");
    if (print_acc && !s->P.print_accel)
        return fail(HGPU_EINVAL, "hgpu_stations_attach: accelerations need hgpu_params_t.print_accel (tm3, psolve.c:4094-4101)");
    for (int64_t i = 0; i < 8LL * nstations; i++)
        if (nodes[i] < 0 || nodes[i] >= s->N) return fail(HGPU_EINVAL, "hgpu_stations_attach: node id out of range");
    CK(cudaSetDevice(s->dev));
    CK(cudaStreamSynchronize(s->stream));
    dfree(s->d_st_nodes); dfree(s->d_st_local); dfree(s->d_st_rows);
    s->d_st_nodes = nullptr; s->d_st_local = nullptr; s->d_st_rows = nullptr;
    s->st_n = nstations; s->st_vel = print_vel != 0; s->st_acc = print_acc != 0; s->st_rate = rate;
    s->st_cap = capacity; s->st_count = 0; s->st_steps.clear();
    if (nstations == 0) return HGPU_OK;
    int rc;
    if ((rc = upload(s, &s->d_st_nodes, nodes, 8 * (size_t)nstations))) return rc;
    if ((rc = upload(s, &s->d_st_local, localcoords, 3 * (size_t)nstations))) return rc;
    if ((rc = dalloc(s, &s->d_st_rows, 9 * (size_t)nstations * (size_t)capacity))) return rc;
    return HGPU_OK;
}

extern "C" int hgpu_stations_record(hgpu_solver_t *s, int32_t step)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    if (s->st_n == 0) return HGPU_OK;
    if (s->st_count >= s->st_cap)
        return fail(HGPU_ESTATE, "hgpu_stations_record: the device ring holds %d rows already; call hgpu_stations_drain", s->st_cap);
    CK(cudaSetDevice(s->dev));
    station_kernel<<<grid_for(s->st_n, 64), 64, 0, s->stream>>>(
        s->st_n, s->d_st_nodes, s->d_st_local, s->u[s->i1], s->u[s->i2], s->u[s->i3], s->st_vel, s->st_acc,
        s->P.dt, s->P.dt2, s->d_st_rows + 9 * (size_t)s->st_n * (size_t)s->st_count);
    CK(cudaGetLastError());
    s->tm.launches++;
    s->st_steps.push_back(step);
    s->st_count++;
    return HGPU_OK;
}

extern "C" int hgpu_stations_pending(hgpu_solver_t *s)
{
    return s ? s->st_count : fail(HGPU_EINVAL, "null solver");
}

extern "C" int hgpu_stations_drain(hgpu_solver_t *s, double *rows, int32_t *steps, int32_t max_rows, int32_t *nrows)
{
    if (!s || !nrows) return fail(HGPU_EINVAL, "hgpu_stations_drain: bad argument");
    *nrows = 0;
    if (s->st_count == 0 || s->st_n == 0) { s->st_count = 0; s->st_steps.clear(); return HGPU_OK; }
    if (!rows || max_rows < s->st_count)
        return fail(HGPU_EINVAL, "hgpu_stations_drain: %d rows are pending, room for %d", s->st_count, max_rows);
    CK(cudaSetDevice(s->dev));
    CK(cudaMemcpyAsync(rows, s->d_st_rows, 9 * (size_t)s->st_n * (size_t)s->st_count * sizeof(double),
                       cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (steps) memcpy(steps, s->st_steps.data(), (size_t)s->st_count * sizeof(int32_t));
    *nrows = s->st_count;
    s->st_count = 0; s->st_steps.clear();
    return check_device_error(s);
}

// ---- planes on the device -----------------------------------------------------------

extern "C" int hgpu_planes_attach(hgpu_solver_t *s, int64_t npoints, const int32_t *nodes, const double *localcoords)
{
    if (!s || npoints < 0 || (npoints > 0 && (!nodes || !localcoords)))
        return fail(HGPU_EINVAL, "hgpu_planes_attach: bad argument");
    for (int64_t i = 0; i < 8 * npoints; i++)
        if (nodes[i] < 0 || nodes[i] >= s->N) return fail(HGPU_EINVAL, "hgpu_planes_attach: node id out of range");
    CK(cudaSetDevice(s->dev));
    CK(cudaStreamSynchronize(s->stream));
    if (s->copy_stream) CK(cudaStreamSynchronize(s->copy_stream));
    dfree(s->d_pl_nodes); dfree(s->d_pl_local);
    s->d_pl_nodes = nullptr; s->d_pl_local = nullptr;
    for (auto &q : s->pl) { dfree(q.rows); q.rows = nullptr; q.busy = false; }
    s->pl_n = npoints; s->pl_next = 0;
    if (npoints == 0) return HGPU_OK;
    int rc;
    if ((rc = upload(s, &s->d_pl_nodes, nodes, 8 * (size_t)npoints))) return rc;
    if ((rc = upload(s, &s->d_pl_local, localcoords, 3 * (size_t)npoints))) return rc;
    for (auto &q : s->pl) {
        if ((rc = dalloc(s, &q.rows, 3 * (size_t)npoints))) return rc;
        if (!q.ready) { CK(cudaEventCreateWithFlags(&q.ready, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&q.done, cudaEventDisableTiming)); }
    }
    if (!s->copy_stream) CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    return HGPU_OK;
}

// One plane step: interpolate in stream order (the rows are those of tm1 as of this point of the step
// sequence), read them out on the copy stream.  Two steps can be in flight; a third waits for the older.
extern "C" int hgpu_planes_record(hgpu_solver_t *s, double *out)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    if (s->pl_n == 0) return HGPU_OK;
    if (!out) return fail(HGPU_EINVAL, "hgpu_planes_record: null output buffer");
    CK(cudaSetDevice(s->dev));
    hgpu_solver::PlaneSlot &q = s->pl[s->pl_next];
    s->pl_next ^= 1;
    if (q.busy) { CK(cudaEventSynchronize(q.done)); q.busy = false; }
    plane_kernel<<<grid_for(s->pl_n, 128), 128, 0, s->stream>>>((long long)s->pl_n, s->d_pl_nodes, s->d_pl_local,
                                                                   s->u[s->i1], q.rows);
    CK(cudaGetLastError());
    s->tm.launches++;
    CK(cudaEventRecord(q.ready, s->stream));
    CK(cudaStreamWaitEvent(s->copy_stream, q.ready, 0));
    CK(cudaMemcpyAsync(out, q.rows, 3 * (size_t)s->pl_n * sizeof(double), cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaEventRecord(q.done, s->copy_stream));
    q.busy = true;
    return HGPU_OK;
}

extern "C" int hgpu_planes_wait(hgpu_solver_t *s)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    for (auto &q : s->pl)
        if (q.busy) { CK(cudaEventSynchronize(q.done)); q.busy = false; }
    return check_device_error(s);
}

extern "C" void *hgpu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(HGPU_ENOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e)); return nullptr; }
    return p;
}

extern "C" void hgpu_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" int hgpu_sync(hgpu_solver_t *s)
{
    if (!s) return fail(HGPU_EINVAL, "null solver");
    CK(cudaSetDevice(s->dev));
    CK(cudaStreamSynchronize(s->stream));
    if (s->comm_stream) CK(cudaStreamSynchronize(s->comm_stream));
    return check_device_error(s);
}

extern "C" int hgpu_get_timers(hgpu_solver_t *s, hgpu_timers_t *out)
{
    if (!s || !out) return fail(HGPU_EINVAL, "null argument");
    CK(cudaSetDevice(s->dev));
    drain_events(s);
    double *dst = &s->tm.addforce_s;   // the eleven double fields are contiguous, in Phase order
    for (int i = 0; i < PH_COUNT; i++) dst[i] = s->phase_s[i];
    *out = s->tm;
    return HGPU_OK;
}

extern "C" void *hgpu_stream(hgpu_solver_t *s) { return s ? (void *)s->stream : nullptr; }

// Host-only: build the tile plan for a mesh with the capacities of a B200 SM (227 KB opt-in shared
// memory, two CTAs per SM), check it against the mesh, and report its sizes.  Needs no device.
extern "C" int hgpu_plan_build(const hgpu_mesh_t *mesh, int32_t tile_nodes, hgpu_layout_t *out)
{
    if (!mesh || !out || !mesh->elem_lnid) return fail(HGPU_EINVAL, "null argument");
    if (mesh->lenum < 0 || mesh->nharbored <= 0) return fail(HGPU_EINVAL, "bad mesh counts");
    TilePlan pl;
    std::string err;
    // structured tiles recognised as hgpu_init does for the fused effective-stiffness kernels (HGPU_STRUCT=0: not)
    const char *senv = getenv("HGPU_STRUCT");
    const TileCaps caps = tile_caps(232448, tile_nodes, senv && atoi(senv) == 1);
    // nodes of the halo schedules and the hanging-node lists (when given) make their tiles "self"
    // tiles, as hgpu_init does on a multi-rank mesh
    std::vector<uint8_t> self_node;
    {
        const hgpu_msglist_t *lists[4] = {&mesh->dn_c, &mesh->dn_s, &mesh->an_c, &mesh->an_s};
        for (const hgpu_msglist_t *l : lists) {
            int32_t tot = 0;
            for (int32_t i = 0; i < l->count; i++) tot += l->nodes[i];
            if (tot > 0 && self_node.empty()) self_node.assign((size_t)mesh->nharbored, 0);
            for (int32_t i = 0; i < tot; i++) {
                if (l->mapping[i] < 0 || l->mapping[i] >= mesh->nharbored) return fail(HGPU_EINVAL, "mapping entry out of range");
                self_node[l->mapping[i]] = 1;
            }
        }
        if (!self_node.empty() && mesh->dnode)
            for (int32_t d = 0; d < mesh->ldnnum; d++) {
                const int32_t *dn = mesh->dnode + 6 * (size_t)d;
                for (int a = 0; a < 6; a++) if (a != 1 && dn[a] >= 0 && dn[a] < mesh->nharbored) self_node[dn[a]] = 1;
            }
    }
    if (!build_tile_plan(mesh->lenum, mesh->nharbored, mesh->elem_lnid, caps,
                         self_node.empty() ? nullptr : self_node.data(), nullptr, pl, err))
        return fail(HGPU_EINVAL, "tile plan: %s", err.c_str());
    if (!validate_tile_plan(mesh->lenum, mesh->nharbored, mesh->elem_lnid, pl, err))
        return fail(HGPU_EINVAL, "tile plan check: %s", err.c_str());
    memset(out, 0, sizeof *out);
    out->tile_nodes = pl.max_tile_owned; out->ntiles = pl.ntiles;
    out->max_tile_nodes = pl.max_tile_nodes; out->max_tile_elems = pl.max_tile_elems;
    out->tile_elems_total = (int64_t)pl.elem_id.size();
    out->tile_halo_total = pl.halo_nodes_total;
    estimate_wavefronts(pl, &out->est_gather_wavefronts, &out->est_scatter_wavefronts);
    out->smem_bytes = step_smem_bytes(true, (pl.max_tile_nodes + 15) & ~15, (pl.max_tile_acc + 15) & ~15,
                                      (pl.max_tile_owned + 15) & ~15, std::min(caps.max_recs, (pl.max_tile_recs + 15) & ~15),
                                      std::min(caps.max_srcs, (pl.max_tile_srcs + 15) & ~15));
    out->block_threads = 256;
    for (int32_t t = 0; t < pl.ntiles; t++) out->early_tiles += pl.tile_self[t];
    out->max_tile_acc = pl.max_tile_acc; out->max_tile_recs = pl.max_tile_recs; out->max_tile_srcs = pl.max_tile_srcs;
    out->partial_slots = (int64_t)pl.halo_id.size(); out->deps_total = (int64_t)pl.dep.size();
    for (uint8_t f : pl.tile_struct) out->struct_tiles += f;       // geometric count (materials are not looked at here)
    return HGPU_OK;
}

extern "C" int hgpu_get_layout(hgpu_solver_t *s, hgpu_layout_t *out)
{
    if (!s || !out) return fail(HGPU_EINVAL, "null argument");
    const TilePlan &pl = s->plan;
    memset(out, 0, sizeof *out);
    out->tile_nodes = pl.max_tile_owned; out->ntiles = pl.ntiles;
    out->max_tile_nodes = pl.max_tile_nodes; out->max_tile_elems = pl.max_tile_elems;
    out->tile_elems_total = (int64_t)pl.elem_id.size();
    out->tile_halo_total = pl.halo_nodes_total;
    estimate_wavefronts(pl, &out->est_gather_wavefronts, &out->est_scatter_wavefronts);
    out->n_regular = s->n_regular; out->n_special = s->n_special;
    out->device_bytes = s->device_bytes;
    out->smem_bytes = s->smem_u2; out->block_threads = s->block;
    out->grid_ctas = s->grid; out->ctas_per_sm = s->ctas_per_sm; out->early_tiles = s->n_early;
    out->max_tile_acc = pl.max_tile_acc; out->max_tile_recs = pl.max_tile_recs; out->max_tile_srcs = pl.max_tile_srcs;
    out->partial_slots = (int64_t)pl.halo_id.size(); out->deps_total = (int64_t)pl.dep.size();
    out->struct_tiles = s->n_struct;
    return HGPU_OK;
}

// ---- multi-GPU ------------------------------------------------------------------------------------

static int make_comm_stream(hgpu_solver *s)
{
    if (!s->comm_stream) {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&s->ev_early, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming));
    }
    return HGPU_OK;
}

// ---- peer-memory transport: CUDA IPC mailboxes ------------------------------------------------
//
// blob = { int32 magic, rank, nranks, nseg; cudaIpcMemHandle_t (64 B); uint64 mailbox_bytes;
//          nseg x { int32 list, peer, count, pad; uint64 data_off, flag_off } }
// list: 0 dn_c, 1 dn_s, 2 an_c, 3 an_s.  A contribution travels c-list -> the peer's s-list
// segment; a sharing travels s-list -> the peer's c-list segment (psolve.c:4945-5079).
struct BlobSeg { int32_t list, peer, count, pad; uint64_t data_off, flag_off; };
static const int32_t BLOB_MAGIC = 0x48475055;

static MsgList *list_of(hgpu_solver *s, int l) { return l == 0 ? &s->dn_c : l == 1 ? &s->dn_s : l == 2 ? &s->an_c : &s->an_s; }

static int p2p_alloc_mailbox(hgpu_solver *s)
{
    if (s->mailbox) return HGPU_OK;
    size_t off = 0;
    for (int l = 0; l < 4; l++) {
        MsgList *m = list_of(s, l);
        m->mb_data_off = off;
        off += 2 * 3 * (size_t)m->total * sizeof(double);       // two parities
        off = (off + 255) & ~(size_t)255;
    }
    for (int l = 0; l < 4; l++) {
        MsgList *m = list_of(s, l);
        m->mb_flag_off = off;
        off += 2 * m->peer.size() * sizeof(unsigned long long);
        off = (off + 255) & ~(size_t)255;
    }
    s->mailbox_bytes = std::max<size_t>(off, 256);
    cudaError_t e = cudaMalloc((void **)&s->mailbox, s->mailbox_bytes);
    if (e != cudaSuccess) return fail(HGPU_ENOMEM, "cudaMalloc(mailbox): %s", cudaGetErrorString(e));
    s->device_bytes += (int64_t)s->mailbox_bytes;
    CK(cudaMemset(s->mailbox, 0, s->mailbox_bytes));
    CK(cudaDeviceSynchronize());
    return HGPU_OK;
}

extern "C" int hgpu_comm_p2p_export(hgpu_solver_t *s, void *blob, int32_t capacity, int32_t *size_out)
{
    if (!s || !size_out) return fail(HGPU_EINVAL, "null argument");
    CK(cudaSetDevice(s->dev));
    int rc = p2p_alloc_mailbox(s);
    if (rc) return rc;
    int32_t nseg = 0;
    for (int l = 0; l < 4; l++) nseg += (int32_t)list_of(s, l)->peer.size();
    const int32_t need = 16 + 64 + 8 + (int32_t)sizeof(BlobSeg) * nseg;
    *size_out = need;
    if (!blob || capacity < need) return blob ? fail(HGPU_EINVAL, "blob buffer too small (%d < %d)", capacity, need) : HGPU_OK;
    char *p = (char *)blob;
    int32_t hdr[4] = {BLOB_MAGIC, s->P.rank, s->P.nranks, nseg};
    memcpy(p, hdr, 16); p += 16;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, s->mailbox));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(p, &h, 64); p += 64;
    const uint64_t mb = s->mailbox_bytes;
    memcpy(p, &mb, 8); p += 8;
    for (int l = 0; l < 4; l++) {
        MsgList *m = list_of(s, l);
        for (size_t i = 0; i < m->peer.size(); i++) {
            BlobSeg sg{l, m->peer[i], m->nodes[i], 0,
                       (uint64_t)(m->mb_data_off + 2 * 3 * (size_t)m->off[i] * sizeof(double)),
                       (uint64_t)(m->mb_flag_off + 2 * i * sizeof(unsigned long long))};
            memcpy(p, &sg, sizeof sg); p += sizeof sg;
        }
    }
    return HGPU_OK;
}

extern "C" int hgpu_comm_p2p_connect(hgpu_solver_t *s, const void *const *blobs, const int32_t *sizes)
{
    if (!s || !blobs || !sizes) return fail(HGPU_EINVAL, "null argument");
    if (s->P.nranks == 1) return HGPU_OK;
    CK(cudaSetDevice(s->dev));
    int rc = p2p_alloc_mailbox(s);
    if (rc) return rc;
    const int R = s->P.nranks;
    s->peer_base.assign((size_t)R, nullptr);
    std::vector<std::vector<BlobSeg>> segs((size_t)R);
    // which peers do I talk to at all
    std::vector<uint8_t> need((size_t)R, 0);
    for (int l = 0; l < 4; l++) for (int32_t p : list_of(s, l)->peer) need[p] = 1;
    for (int r = 0; r < R; r++) {
        if (r == s->P.rank || !need[r]) continue;
        const char *p = (const char *)blobs[r];
        if (!p || sizes[r] < 88) return fail(HGPU_ECOMM, "missing or short p2p blob of rank %d", r);
        int32_t hdr[4];
        memcpy(hdr, p, 16);
        if (hdr[0] != BLOB_MAGIC || hdr[1] != r || hdr[2] != R || sizes[r] < 88 + (int32_t)sizeof(BlobSeg) * hdr[3])
            return fail(HGPU_ECOMM, "malformed p2p blob of rank %d", r);
        cudaIpcMemHandle_t h;
        memcpy(&h, p + 16, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&s->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return fail(HGPU_ECOMM, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        segs[r].resize((size_t)hdr[3]);
        memcpy(segs[r].data(), p + 88, sizeof(BlobSeg) * (size_t)hdr[3]);
    }
    for (int l = 0; l < 4; l++) {
        MsgList *m = list_of(s, l);
        const int comp = l ^ 1;                       // dn_c <-> dn_s, an_c <-> an_s
        const size_t nm = m->peer.size();
        m->remote_data.assign(nm, nullptr);
        m->remote_flag.assign(nm, nullptr);
        for (size_t i = 0; i < nm; i++) {
            const int p = m->peer[i];
            const BlobSeg *hit = nullptr;
            for (const BlobSeg &sg : segs[p]) if (sg.list == comp && sg.peer == s->P.rank) hit = &sg;
            if (!hit || hit->count != m->nodes[i])
                return fail(HGPU_ECOMM, "rank %d has no matching segment for list %d of rank %d (%d nodes)",
                            p, l, s->P.rank, m->nodes[i]);
            m->remote_data[i] = (double *)((char *)s->peer_base[p] + hit->data_off);
            m->remote_flag[i] = (unsigned long long *)((char *)s->peer_base[p] + hit->flag_off);
        }
        if ((rc = dalloc(s, &m->d_counters, std::max<size_t>(nm, 1)))) return rc;
        CK(cudaMemset(m->d_counters, 0, std::max<size_t>(nm, 1) * sizeof(unsigned int)));
        std::vector<PushSeg> push(2 * nm);
        std::vector<PullSeg> pull(2 * nm);
        for (int par = 0; par < 2; par++)
            for (size_t i = 0; i < nm; i++) {
                const size_t n3 = 3 * (size_t)m->nodes[i];
                push[par * nm + i] = PushSeg{m->d_map + m->off[i], m->remote_data[i] + par * n3,
                                             m->remote_flag[i] + par, m->d_counters + i, m->nodes[i]};
                const double *loc = (const double *)(s->mailbox + m->mb_data_off) + 2 * 3 * (size_t)m->off[i] + par * n3;
                const unsigned long long *fl = (const unsigned long long *)(s->mailbox + m->mb_flag_off) + 2 * i + par;
                pull[par * nm + i] = PullSeg{m->d_map + m->off[i], loc, fl, m->nodes[i]};
            }
        if ((rc = upload(s, &m->d_push, push.data(), push.size()))) return rc;
        if ((rc = upload(s, &m->d_pull, pull.data(), pull.size()))) return rc;
    }
    // contribution lists (the s-lists receive): node-major CSR over all messengers, list order kept
    for (MsgList *m : {&s->dn_s, &s->an_s}) {
        std::vector<std::pair<int32_t, int32_t>> ent;      // (node, position in the concatenated mapping)
        std::vector<int32_t> map_host((size_t)m->total);
        if (m->total) CK(cudaMemcpy(map_host.data(), m->d_map, (size_t)m->total * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (int32_t k = 0; k < m->total; k++) ent.emplace_back(map_host[k], k);
        std::stable_sort(ent.begin(), ent.end(), [](const std::pair<int32_t, int32_t> &a, const std::pair<int32_t, int32_t> &b) { return a.first < b.first; });
        std::vector<int32_t> node, off, src0, src1;
        std::vector<int32_t> msg_of((size_t)m->total);
        for (size_t i = 0; i < m->peer.size(); i++) for (int32_t k = m->off[i]; k < m->off[i + 1]; k++) msg_of[k] = (int32_t)i;
        for (size_t j = 0; j < ent.size(); j++) {
            if (j == 0 || ent[j].first != ent[j - 1].first) { node.push_back(ent[j].first); off.push_back((int32_t)j); }
            const int32_t k = ent[j].second, i = msg_of[k], idx = k - m->off[i];
            src0.push_back(2 * m->off[i] + idx);
            src1.push_back(2 * m->off[i] + m->nodes[i] + idx);
        }
        off.push_back((int32_t)ent.size());
        m->csr_nodes = (int32_t)node.size();
        if ((rc = upload(s, &m->d_csr_node, node.data(), node.size()))) return rc;
        if ((rc = upload(s, &m->d_csr_off, off.data(), off.size()))) return rc;
        if ((rc = upload(s, &m->d_csr_src[0], src0.data(), src0.size()))) return rc;
        if ((rc = upload(s, &m->d_csr_src[1], src1.data(), src1.size()))) return rc;
    }
    if ((rc = make_comm_stream(s))) return rc;
    s->p2p_ready = true;
    return HGPU_OK;
}

extern "C" int hgpu_comm_unique_id(void *unique_id_128)
{
    if (!unique_id_128) return fail(HGPU_EINVAL, "null argument");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(unique_id_128, &id, sizeof id);
    return HGPU_OK;
}

extern "C" int hgpu_comm_init(hgpu_solver_t *s, const void *unique_id_128)
{
    if (!s || !unique_id_128) return fail(HGPU_EINVAL, "null argument");
    if (s->P.nranks == 1) return HGPU_OK;
    int rc = load_nccl();
    if (rc) return rc;
    CK(cudaSetDevice(s->dev));
    ncclUniqueId id;
    memcpy(&id, unique_id_128, sizeof id);
    NK(g_nccl.CommInitRank(&s->comm, s->P.nranks, id, s->P.rank));
    return make_comm_stream(s);
}
