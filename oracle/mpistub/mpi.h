/*
 * mpi.h -- "hmpi": a tiny single-node MPI subset for building the UNMODIFIED
 * Hercules reference (etree/octor/quake-forward) as the parity oracle and the
 * CPU baseline.  TEST INFRASTRUCTURE ONLY: nothing in the product path
 * (hercules_b200/, include/) includes or links this.
 *
 * The image has no MPI.  This covers exactly the symbols the reference uses
 * (census in SURVEY.md section 2.2).  Ranks are fork()ed inside MPI_Init when
 * the environment variable HMPI_NP=<P> is set (default 1); messages travel
 * through a MAP_SHARED arena with per-destination queues, so every send is
 * buffered (eager) and can never deadlock against the reference's
 * Irecv -> Send -> Waitall pattern (quake/forward/psolve.c:4945-5079).
 */
#ifndef HMPI_MPI_H
#define HMPI_MPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Group;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct hmpi_file_s *MPI_File;

typedef struct MPI_Status {
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
    int hmpi_bytes;
} MPI_Status;

#define MPI_SUCCESS       0
#define MPI_COMM_WORLD    ((MPI_Comm)0)
#define MPI_COMM_NULL     ((MPI_Comm)-1)
#define MPI_UNDEFINED     (-32766)
#define MPI_ANY_SOURCE    (-1)
#define MPI_ANY_TAG       (-1)
#define MPI_INFO_NULL     ((MPI_Info)0)
#define MPI_REQUEST_NULL  ((MPI_Request)-1)
#define MPI_STATUS_IGNORE   ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

/* basic datatypes: small ids, sizes looked up in the implementation */
#define MPI_CHAR           ((MPI_Datatype)1)
#define MPI_BYTE           ((MPI_Datatype)2)
#define MPI_INT            ((MPI_Datatype)3)
#define MPI_LONG           ((MPI_Datatype)4)
#define MPI_LONG_LONG_INT  ((MPI_Datatype)5)
#define MPI_LONG_LONG      MPI_LONG_LONG_INT
#define MPI_FLOAT          ((MPI_Datatype)6)
#define MPI_DOUBLE         ((MPI_Datatype)7)
#define MPI_UNSIGNED       ((MPI_Datatype)8)
#define MPI_UNSIGNED_LONG  ((MPI_Datatype)9)

#define MPI_SUM  ((MPI_Op)1)
#define MPI_MAX  ((MPI_Op)2)
#define MPI_MIN  ((MPI_Op)3)

#define MPI_WTIME_IS_GLOBAL 7

#define MPI_MODE_RDONLY  1
#define MPI_MODE_WRONLY  2
#define MPI_MODE_RDWR    4
#define MPI_MODE_CREATE  8
#define MPI_SEEK_SET     0

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);

int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group);
int MPI_Group_incl(MPI_Group group, int n, const int *ranks, MPI_Group *newgroup);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *newcomm);
int MPI_Attr_get(MPI_Comm comm, int keyval, void *attr, int *flag);

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *type);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
               MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
                  MPI_Op op, MPI_Comm comm);
int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
               void *recvbuf, int recvcount, MPI_Datatype recvtype, int root,
               MPI_Comm comm);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                  void *recvbuf, int recvcount, MPI_Datatype recvtype,
                  MPI_Comm comm);

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag,
             MPI_Comm comm);
int MPI_Ssend(const void *buf, int count, MPI_Datatype type, int dest, int tag,
              MPI_Comm comm);
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag,
              MPI_Comm comm, MPI_Request *request);
int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag,
             MPI_Comm comm, MPI_Status *status);
int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag,
              MPI_Comm comm, MPI_Request *request);
int MPI_Waitall(int count, MPI_Request *requests, MPI_Status *statuses);
int MPI_Probe(int source, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Iprobe(int source, int tag, MPI_Comm comm, int *flag, MPI_Status *status);
int MPI_Get_count(const MPI_Status *status, MPI_Datatype type, int *count);

int MPI_File_open(MPI_Comm comm, const char *filename, int amode, MPI_Info info,
                  MPI_File *fh);
int MPI_File_close(MPI_File *fh);
int MPI_File_seek(MPI_File fh, MPI_Offset offset, int whence);
int MPI_File_read(MPI_File fh, void *buf, int count, MPI_Datatype type,
                  MPI_Status *status);
int MPI_File_read_at(MPI_File fh, MPI_Offset offset, void *buf, int count,
                     MPI_Datatype type, MPI_Status *status);
int MPI_File_write(MPI_File fh, const void *buf, int count, MPI_Datatype type,
                   MPI_Status *status);
int MPI_File_write_at(MPI_File fh, MPI_Offset offset, const void *buf, int count,
                      MPI_Datatype type, MPI_Status *status);

#ifdef __cplusplus
}
#endif

#endif /* HMPI_MPI_H */
