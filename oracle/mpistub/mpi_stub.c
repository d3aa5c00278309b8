/*
 * mpi_stub.c -- implementation of "hmpi" (see mpi.h).  TEST INFRASTRUCTURE ONLY.
 *
 * Process model : MPI_Init fork()s HMPI_NP-1 children (default HMPI_NP=1); rank 0
 *                 is the original process and reaps the children at exit.
 * Transport     : one MAP_SHARED arena; per-destination FIFO queues of messages
 *                 guarded by one process-shared mutex + one condvar per rank.
 *                 Sends are always buffered.  Blocks are recycled through
 *                 power-of-two size-class free lists.
 * Matching      : (context id, source rank in that communicator, tag), FIFO per
 *                 sender, MPI_ANY_SOURCE / MPI_ANY_TAG honoured (ANY_TAG only
 *                 matches user tags >= 0; collectives use negative tags).
 * Collectives   : linear algorithms over the point-to-point layer; reductions
 *                 combine in rank order so results are deterministic.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <sched.h>
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#define HMPI_MAXP      64
#define HMPI_MAXCOMM   64
#define HMPI_MAXREQ    4096
#define HMPI_MAXTYPE   64
#define HMPI_NCLASS    40
#define HMPI_TAG_COLL  (-4242)

typedef struct msg_s {
    int    src;       /* sender rank in the communicator of ctx */
    int    tag;
    int    ctx;
    int    cls;       /* size class of this block */
    size_t nbytes;
    size_t next;      /* arena offset of next message in the queue (0 = none) */
} msg_t;

typedef struct shm_s {
    pthread_mutex_t lock;
    pthread_cond_t  cond[HMPI_MAXP];
    size_t qhead[HMPI_MAXP], qtail[HMPI_MAXP];
    size_t freelist[HMPI_NCLASS];
    size_t bump, arena_size;
    pid_t  pids[HMPI_MAXP];
    int    np;
    volatile int aborted;
} shm_t;

typedef struct comm_s {
    int used, ctx, size, rank;
    int world[HMPI_MAXP];  /* comm rank -> world rank */
} comm_t;

typedef struct req_s {
    int active;
    void *buf; int count; MPI_Datatype type; int src, tag; MPI_Comm comm;
} req_t;

static shm_t *g_shm;
static char  *g_arena;
static int    g_rank = 0, g_np = 1, g_inited = 0;
static comm_t g_comm[HMPI_MAXCOMM];
static int    g_next_ctx = 1;
static req_t  g_req[HMPI_MAXREQ];
static size_t g_typesize[HMPI_MAXTYPE] = {0, 1, 1, sizeof(int), sizeof(long),
    sizeof(long long), sizeof(float), sizeof(double), sizeof(unsigned),
    sizeof(unsigned long)};
static int    g_typebase[HMPI_MAXTYPE] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9};
static int    g_ntypes = 10;

struct hmpi_file_s { int fd; MPI_Offset pos; };

static void hmpi_die(const char *what)
{
    fprintf(stderr, "hmpi[%d]: fatal: %s\n", g_rank, what);
    MPI_Abort(MPI_COMM_WORLD, 1);
}

static msg_t *M(size_t off) { return (msg_t *)(g_arena + off); }

static void check_abort(void)
{
    if (g_shm->aborted) {
        _exit(g_shm->aborted > 0 ? g_shm->aborted : 1);
    }
}

static void lock(void)   { pthread_mutex_lock(&g_shm->lock); }
static void unlock(void) { pthread_mutex_unlock(&g_shm->lock); }

static void wait_for_message(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    ts.tv_sec += 1;
    pthread_cond_timedwait(&g_shm->cond[g_rank], &g_shm->lock, &ts);
    if (g_shm->aborted) { unlock(); check_abort(); }
}

/* lock held */
static size_t arena_alloc(size_t payload, int *cls_out)
{
    size_t need = sizeof(msg_t) + payload;
    int cls = 6;
    while (((size_t)1 << cls) < need) cls++;
    if (cls >= HMPI_NCLASS) hmpi_die("message too large");
    *cls_out = cls;
    if (g_shm->freelist[cls]) {
        size_t off = g_shm->freelist[cls];
        g_shm->freelist[cls] = M(off)->next;
        return off;
    }
    size_t sz = (size_t)1 << cls;
    if (g_shm->bump + sz > g_shm->arena_size) {
        unlock();
        hmpi_die("shared arena exhausted (raise HMPI_ARENA_MB)");
    }
    size_t off = g_shm->bump;
    g_shm->bump += sz;
    return off;
}

/* lock held */
static void arena_free(size_t off)
{
    int cls = M(off)->cls;
    M(off)->next = g_shm->freelist[cls];
    g_shm->freelist[cls] = off;
}

static comm_t *C(MPI_Comm comm)
{
    if (comm < 0 || comm >= HMPI_MAXCOMM || !g_comm[comm].used)
        hmpi_die("invalid communicator");
    return &g_comm[comm];
}

static size_t tsize(MPI_Datatype t)
{
    if (t <= 0 || t >= g_ntypes) hmpi_die("invalid datatype");
    return g_typesize[t];
}

static void raw_send(const void *buf, size_t nbytes, int dest, int tag, MPI_Comm comm)
{
    comm_t *c = C(comm);
    if (dest < 0 || dest >= c->size) hmpi_die("send: bad destination rank");
    int wdest = c->world[dest], cls;
    lock();
    size_t off = arena_alloc(nbytes, &cls);
    msg_t *m = M(off);
    m->src = c->rank; m->tag = tag; m->ctx = c->ctx; m->cls = cls;
    m->nbytes = nbytes; m->next = 0;
    if (nbytes) memcpy((char *)(m + 1), buf, nbytes);
    if (g_shm->qtail[wdest]) M(g_shm->qtail[wdest])->next = off;
    else g_shm->qhead[wdest] = off;
    g_shm->qtail[wdest] = off;
    pthread_cond_broadcast(&g_shm->cond[wdest]);
    unlock();
}

static int match(const msg_t *m, int ctx, int src, int tag)
{
    if (m->ctx != ctx) return 0;
    if (src != MPI_ANY_SOURCE && m->src != src) return 0;
    if (tag == MPI_ANY_TAG) return m->tag >= 0;
    return m->tag == tag;
}

/* lock held; returns offset of first match and its predecessor */
static size_t find_msg(int ctx, int src, int tag, size_t *prev_out)
{
    size_t prev = 0, cur = g_shm->qhead[g_rank];
    while (cur) {
        if (match(M(cur), ctx, src, tag)) { *prev_out = prev; return cur; }
        prev = cur; cur = M(cur)->next;
    }
    return 0;
}

static void raw_recv(void *buf, size_t maxbytes, int src, int tag, MPI_Comm comm,
                     MPI_Status *status)
{
    comm_t *c = C(comm);
    lock();
    for (;;) {
        size_t prev, off = find_msg(c->ctx, src, tag, &prev);
        if (off) {
            msg_t *m = M(off);
            if (m->nbytes > maxbytes) { unlock(); hmpi_die("recv: message truncated"); }
            if (m->nbytes) memcpy(buf, (char *)(m + 1), m->nbytes);
            if (status) {
                status->MPI_SOURCE = m->src; status->MPI_TAG = m->tag;
                status->MPI_ERROR = MPI_SUCCESS; status->hmpi_bytes = (int)m->nbytes;
            }
            if (prev) M(prev)->next = m->next; else g_shm->qhead[g_rank] = m->next;
            if (g_shm->qtail[g_rank] == off) g_shm->qtail[g_rank] = prev;
            arena_free(off);
            unlock();
            return;
        }
        wait_for_message();
    }
}

/* ---- environment ------------------------------------------------------- */

static void reap_children(void)
{
    if (g_rank != 0 || !g_shm) return;
    for (int r = 1; r < g_np; r++) {
        int st;
        if (g_shm->pids[r] > 0) waitpid(g_shm->pids[r], &st, 0);
    }
}

static void on_sigchld(int sig)
{
    (void)sig;
    int st; pid_t p;
    while ((p = waitpid(-1, &st, WNOHANG)) > 0) {
        for (int r = 1; r < g_np; r++) if (g_shm->pids[r] == p) g_shm->pids[r] = 0;
        if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0) && !g_shm->aborted)
            g_shm->aborted = 1;
    }
}

int MPI_Init(int *argc, char ***argv)
{
    (void)argc; (void)argv;
    if (g_inited) return MPI_SUCCESS;
    const char *e = getenv("HMPI_NP");
    g_np = e ? atoi(e) : 1;
    if (g_np < 1 || g_np > HMPI_MAXP) { fprintf(stderr, "hmpi: bad HMPI_NP\n"); exit(1); }
    e = getenv("HMPI_ARENA_MB");
    size_t arena_mb = e ? (size_t)atol(e) : (g_np > 1 ? 4096 : 64);
    size_t total = sizeof(shm_t) + 4096 + (arena_mb << 20);
    void *p = mmap(NULL, total, PROT_READ | PROT_WRITE,
                   MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { perror("hmpi: mmap"); exit(1); }
    g_shm = (shm_t *)p;
    g_arena = (char *)p + ((sizeof(shm_t) + 4095) & ~(size_t)4095);
    memset(g_shm, 0, sizeof(shm_t));
    g_shm->arena_size = arena_mb << 20;
    g_shm->bump = 64;
    g_shm->np = g_np;
    pthread_mutexattr_t ma; pthread_mutexattr_init(&ma);
    pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
    pthread_mutex_init(&g_shm->lock, &ma);
    pthread_condattr_t ca; pthread_condattr_init(&ca);
    pthread_condattr_setpshared(&ca, PTHREAD_PROCESS_SHARED);
    for (int r = 0; r < g_np; r++) pthread_cond_init(&g_shm->cond[r], &ca);

    fflush(stdout); fflush(stderr);
    g_shm->pids[0] = getpid();
    if (g_np > 1) {
        struct sigaction sa; memset(&sa, 0, sizeof sa);
        sa.sa_handler = on_sigchld; sa.sa_flags = SA_RESTART | SA_NOCLDSTOP;
        sigaction(SIGCHLD, &sa, NULL);
    }
    for (int r = 1; r < g_np; r++) {
        pid_t pid = fork();
        if (pid < 0) { perror("hmpi: fork"); exit(1); }
        if (pid == 0) {
            g_rank = r;
            signal(SIGCHLD, SIG_DFL);
            break;
        }
        g_shm->pids[r] = pid;
    }
    if (g_rank == 0) atexit(reap_children);
    /* HMPI_PIN=1: rank r stays on the (r mod n)-th cpu this process may use, so that timing runs do not
     * depend on where the scheduler happens to put (and migrate) the forked ranks */
    e = getenv("HMPI_PIN");
    if (e && atoi(e) > 0) {
        cpu_set_t allowed, mine;
        if (sched_getaffinity(0, sizeof allowed, &allowed) == 0) {
            int n = CPU_COUNT(&allowed), want = g_rank % (n > 0 ? n : 1), seen = 0;
            for (int c = 0; c < CPU_SETSIZE; c++) {
                if (!CPU_ISSET(c, &allowed)) continue;
                if (seen++ == want) { CPU_ZERO(&mine); CPU_SET(c, &mine); sched_setaffinity(0, sizeof mine, &mine); break; }
            }
        }
    }

    memset(g_comm, 0, sizeof g_comm);
    g_comm[0].used = 1; g_comm[0].ctx = 0; g_comm[0].size = g_np; g_comm[0].rank = g_rank;
    for (int r = 0; r < g_np; r++) g_comm[0].world[r] = r;
    g_inited = 1;
    return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
    if (g_inited && g_np > 1) MPI_Barrier(MPI_COMM_WORLD);
    fflush(stdout); fflush(stderr);
    return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int errorcode)
{
    (void)comm;
    fflush(stdout); fflush(stderr);
    if (g_shm) {
        g_shm->aborted = errorcode ? (errorcode & 0xff ? errorcode & 0xff : 1) : 1;
        for (int r = 0; r < g_np; r++) {
            pid_t p = g_shm->pids[r];
            if (r != g_rank && p > 0) kill(p, SIGTERM);
        }
    }
    _exit(errorcode ? (errorcode & 0xff ? errorcode & 0xff : 1) : 1);
    return MPI_SUCCESS;
}

double MPI_Wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- communicators ------------------------------------------------------ */

int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = C(comm)->rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { *size = C(comm)->size; return MPI_SUCCESS; }

static int new_comm_slot(void)
{
    for (int i = 1; i < HMPI_MAXCOMM; i++) if (!g_comm[i].used) return i;
    hmpi_die("too many communicators");
    return -1;
}

/* agree on a fresh context id over the parent communicator */
static int agree_ctx(MPI_Comm parent)
{
    int mine = g_next_ctx, top = 0;
    MPI_Allreduce(&mine, &top, 1, MPI_INT, MPI_MAX, parent);
    g_next_ctx = top + 1;
    return top;
}

int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm)
{
    comm_t *c = C(comm);
    int ctx = agree_ctx(comm);
    int s = new_comm_slot();
    g_comm[s] = *c;
    g_comm[s].ctx = ctx * 1024;
    *newcomm = s;
    return MPI_SUCCESS;
}

int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm)
{
    comm_t *c = C(comm);
    int n = c->size;
    int mine[2] = {color, key};
    int *all = (int *)malloc(sizeof(int) * 2 * n);
    MPI_Allgather(mine, 2, MPI_INT, all, 2, MPI_INT, comm);
    int ctx = agree_ctx(comm);
    if (color == MPI_UNDEFINED) { *newcomm = MPI_COMM_NULL; free(all); return MPI_SUCCESS; }
    int s = new_comm_slot();
    comm_t *nc = &g_comm[s];
    memset(nc, 0, sizeof *nc);
    nc->used = 1;
    /* distinct colours get distinct contexts */
    nc->ctx = ctx * 1024 + (color & 1023);
    int members[HMPI_MAXP], m = 0;
    for (int r = 0; r < n; r++) if (all[2 * r] == color) members[m++] = r;
    /* stable sort by key */
    for (int i = 1; i < m; i++) {
        int v = members[i], j = i - 1;
        while (j >= 0 && all[2 * members[j] + 1] > all[2 * v + 1]) { members[j + 1] = members[j]; j--; }
        members[j + 1] = v;
    }
    nc->size = m;
    for (int i = 0; i < m; i++) {
        nc->world[i] = c->world[members[i]];
        if (members[i] == c->rank) nc->rank = i;
    }
    free(all);
    *newcomm = s;
    return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm *comm)
{
    if (*comm > 0 && *comm < HMPI_MAXCOMM) g_comm[*comm].used = 0;
    *comm = MPI_COMM_NULL;
    return MPI_SUCCESS;
}

int MPI_Comm_group(MPI_Comm comm, MPI_Group *group) { (void)comm; *group = 0; hmpi_die("MPI_Comm_group unsupported (DRM only)"); return 0; }
int MPI_Group_incl(MPI_Group g, int n, const int *r, MPI_Group *ng) { (void)g; (void)n; (void)r; (void)ng; hmpi_die("MPI_Group_incl unsupported (DRM only)"); return 0; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *nc) { (void)c; (void)g; (void)nc; hmpi_die("MPI_Comm_create unsupported (DRM only)"); return 0; }

int MPI_Attr_get(MPI_Comm comm, int keyval, void *attr, int *flag)
{
    static int yes = 1;
    (void)comm;
    if (keyval == MPI_WTIME_IS_GLOBAL) { *(int **)attr = &yes; *flag = 1; }
    else *flag = 0;
    return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype)
{
    if (g_ntypes >= HMPI_MAXTYPE) hmpi_die("too many datatypes");
    g_typesize[g_ntypes] = tsize(oldtype) * (size_t)count;
    g_typebase[g_ntypes] = 0;
    *newtype = g_ntypes++;
    return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *type) { (void)type; return MPI_SUCCESS; }

/* ---- point to point ----------------------------------------------------- */

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm)
{
    raw_send(buf, tsize(type) * (size_t)count, dest, tag, comm);
    return MPI_SUCCESS;
}
int MPI_Ssend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm)
{
    return MPI_Send(buf, count, type, dest, tag, comm);
}
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag,
              MPI_Comm comm, MPI_Request *request)
{
    MPI_Send(buf, count, type, dest, tag, comm);
    *request = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}
int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag,
             MPI_Comm comm, MPI_Status *status)
{
    raw_recv(buf, tsize(type) * (size_t)count, source, tag, comm, status);
    return MPI_SUCCESS;
}
int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag,
              MPI_Comm comm, MPI_Request *request)
{
    for (int i = 0; i < HMPI_MAXREQ; i++) {
        if (!g_req[i].active) {
            g_req[i] = (req_t){1, buf, count, type, source, tag, comm};
            *request = i;
            return MPI_SUCCESS;
        }
    }
    hmpi_die("too many outstanding requests");
    return MPI_SUCCESS;
}
int MPI_Waitall(int count, MPI_Request *requests, MPI_Status *statuses)
{
    for (int i = 0; i < count; i++) {
        int r = requests[i];
        if (r == MPI_REQUEST_NULL) continue;
        if (r < 0 || r >= HMPI_MAXREQ || !g_req[r].active) hmpi_die("waitall: bad request");
        req_t *q = &g_req[r];
        MPI_Recv(q->buf, q->count, q->type, q->src, q->tag, q->comm,
                 statuses ? &statuses[i] : NULL);
        q->active = 0;
        requests[i] = MPI_REQUEST_NULL;
    }
    return MPI_SUCCESS;
}
int MPI_Probe(int source, int tag, MPI_Comm comm, MPI_Status *status)
{
    comm_t *c = C(comm);
    lock();
    for (;;) {
        size_t prev, off = find_msg(c->ctx, source, tag, &prev);
        if (off) {
            if (status) {
                status->MPI_SOURCE = M(off)->src; status->MPI_TAG = M(off)->tag;
                status->MPI_ERROR = MPI_SUCCESS; status->hmpi_bytes = (int)M(off)->nbytes;
            }
            unlock();
            return MPI_SUCCESS;
        }
        wait_for_message();
    }
}
int MPI_Iprobe(int source, int tag, MPI_Comm comm, int *flag, MPI_Status *status)
{
    comm_t *c = C(comm);
    check_abort();
    lock();
    size_t prev, off = find_msg(c->ctx, source, tag, &prev);
    if (off && status) {
        status->MPI_SOURCE = M(off)->src; status->MPI_TAG = M(off)->tag;
        status->MPI_ERROR = MPI_SUCCESS; status->hmpi_bytes = (int)M(off)->nbytes;
    }
    unlock();
    *flag = off != 0;
    return MPI_SUCCESS;
}
int MPI_Get_count(const MPI_Status *status, MPI_Datatype type, int *count)
{
    *count = (int)((size_t)status->hmpi_bytes / tsize(type));
    return MPI_SUCCESS;
}

/* ---- collectives -------------------------------------------------------- */

int MPI_Barrier(MPI_Comm comm)
{
    comm_t *c = C(comm);
    char tok = 0;
    if (c->size == 1) return MPI_SUCCESS;
    if (c->rank == 0) {
        for (int r = 1; r < c->size; r++) raw_recv(&tok, 1, r, HMPI_TAG_COLL, comm, NULL);
        for (int r = 1; r < c->size; r++) raw_send(&tok, 1, r, HMPI_TAG_COLL, comm);
    } else {
        raw_send(&tok, 1, 0, HMPI_TAG_COLL, comm);
        raw_recv(&tok, 1, 0, HMPI_TAG_COLL, comm, NULL);
    }
    return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
    comm_t *c = C(comm);
    size_t n = tsize(type) * (size_t)count;
    if (c->size == 1) return MPI_SUCCESS;
    if (c->rank == root) {
        for (int r = 0; r < c->size; r++) if (r != root) raw_send(buf, n, r, HMPI_TAG_COLL, comm);
    } else {
        raw_recv(buf, n, root, HMPI_TAG_COLL, comm, NULL);
    }
    return MPI_SUCCESS;
}

#define COMBINE(T)                                                         \
    do {                                                                   \
        T *a = (T *)acc; const T *b = (const T *)in;                        \
        for (int i = 0; i < count; i++) {                                  \
            if (op == MPI_SUM) a[i] += b[i];                               \
            else if (op == MPI_MAX) { if (b[i] > a[i]) a[i] = b[i]; }      \
            else if (op == MPI_MIN) { if (b[i] < a[i]) a[i] = b[i]; }      \
            else hmpi_die("unsupported reduction op");                     \
        }                                                                  \
    } while (0)

static void combine(void *acc, const void *in, int count, MPI_Datatype type, MPI_Op op)
{
    switch (g_typebase[type]) {
    case 1: case 2: COMBINE(char); break;
    case 3: COMBINE(int); break;
    case 4: COMBINE(long); break;
    case 5: COMBINE(long long); break;
    case 6: COMBINE(float); break;
    case 7: COMBINE(double); break;
    case 8: COMBINE(unsigned); break;
    case 9: COMBINE(unsigned long); break;
    default: hmpi_die("reduce on derived datatype");
    }
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
               MPI_Op op, int root, MPI_Comm comm)
{
    comm_t *c = C(comm);
    size_t n = tsize(type) * (size_t)count;
    if (c->rank != root) {
        raw_send(sendbuf, n, root, HMPI_TAG_COLL, comm);
        return MPI_SUCCESS;
    }
    /* combine in rank order for determinism */
    char *tmp = (char *)malloc(n ? n : 1), *acc = (char *)malloc(n ? n : 1);
    for (int r = 0; r < c->size; r++) {
        const void *in;
        if (r == root) in = sendbuf;
        else { raw_recv(tmp, n, r, HMPI_TAG_COLL, comm, NULL); in = tmp; }
        if (r == 0) memcpy(acc, in, n); else combine(acc, in, count, type, op);
    }
    memcpy(recvbuf, acc, n);
    free(tmp); free(acc);
    return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
                  MPI_Op op, MPI_Comm comm)
{
    MPI_Reduce(sendbuf, recvbuf, count, type, op, 0, comm);
    MPI_Bcast(recvbuf, count, type, 0, comm);
    return MPI_SUCCESS;
}

int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
               void *recvbuf, int recvcount, MPI_Datatype recvtype, int root,
               MPI_Comm comm)
{
    comm_t *c = C(comm);
    size_t ns = tsize(sendtype) * (size_t)sendcount;
    if (c->rank != root) {
        raw_send(sendbuf, ns, root, HMPI_TAG_COLL, comm);
        return MPI_SUCCESS;
    }
    size_t nr = tsize(recvtype) * (size_t)recvcount;
    for (int r = 0; r < c->size; r++) {
        char *dst = (char *)recvbuf + nr * (size_t)r;
        if (r == root) memcpy(dst, sendbuf, ns);
        else raw_recv(dst, nr, r, HMPI_TAG_COLL, comm, NULL);
    }
    return MPI_SUCCESS;
}

int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                  void *recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm)
{
    comm_t *c = C(comm);
    MPI_Gather(sendbuf, sendcount, sendtype, recvbuf, recvcount, recvtype, 0, comm);
    MPI_Bcast(recvbuf, recvcount * c->size, recvtype, 0, comm);
    return MPI_SUCCESS;
}

/* ---- files (DRM only in the reference; plain POSIX I/O) ----------------- */

int MPI_File_open(MPI_Comm comm, const char *filename, int amode, MPI_Info info, MPI_File *fh)
{
    (void)comm; (void)info;
    int flags = 0;
    if (amode & MPI_MODE_RDWR) flags |= O_RDWR;
    else if (amode & MPI_MODE_WRONLY) flags |= O_WRONLY;
    else flags |= O_RDONLY;
    if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
    int fd = open(filename, flags, 0644);
    if (fd < 0) return 1;
    *fh = (MPI_File)malloc(sizeof(struct hmpi_file_s));
    (*fh)->fd = fd; (*fh)->pos = 0;
    return MPI_SUCCESS;
}
int MPI_File_close(MPI_File *fh)
{
    if (*fh) { close((*fh)->fd); free(*fh); *fh = NULL; }
    return MPI_SUCCESS;
}
int MPI_File_seek(MPI_File fh, MPI_Offset offset, int whence)
{
    (void)whence; fh->pos = offset; return MPI_SUCCESS;
}
int MPI_File_read_at(MPI_File fh, MPI_Offset offset, void *buf, int count,
                     MPI_Datatype type, MPI_Status *status)
{
    ssize_t n = pread(fh->fd, buf, tsize(type) * (size_t)count, (off_t)offset);
    if (status) status->hmpi_bytes = n > 0 ? (int)n : 0;
    return n < 0;
}
int MPI_File_read(MPI_File fh, void *buf, int count, MPI_Datatype type, MPI_Status *status)
{
    int rc = MPI_File_read_at(fh, fh->pos, buf, count, type, status);
    fh->pos += (MPI_Offset)(tsize(type) * (size_t)count);
    return rc;
}
int MPI_File_write_at(MPI_File fh, MPI_Offset offset, const void *buf, int count,
                      MPI_Datatype type, MPI_Status *status)
{
    ssize_t n = pwrite(fh->fd, buf, tsize(type) * (size_t)count, (off_t)offset);
    if (status) status->hmpi_bytes = n > 0 ? (int)n : 0;
    return n < 0;
}
int MPI_File_write(MPI_File fh, const void *buf, int count, MPI_Datatype type, MPI_Status *status)
{
    int rc = MPI_File_write_at(fh, fh->pos, buf, count, type, status);
    fh->pos += (MPI_Offset)(tsize(type) * (size_t)count);
    return rc;
}
