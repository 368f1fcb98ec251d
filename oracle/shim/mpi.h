/* oracle/shim/mpi.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Stand-in for <mpi.h> so the reference's unmodified sources compile here (no MPI
 * in this image).  Declares exactly the entry points the reference calls
 * (mcpar.cc:37,131,136,228,231; mcout.cc:11-17,67-121;
 * mains).  Semantics are implemented by shim_mpi.cc as a THREAD-based mini-MPI:
 * each "rank" is one std::thread of the harness process.
 */
#ifndef ORACLE_SHIM_MPI_H_
#define ORACLE_SHIM_MPI_H_
#include <stddef.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

#define MPI_COMM_WORLD     ((MPI_Comm)1)
#define MPI_SUCCESS        0
#define MPI_IN_PLACE       ((void*)1)
#define MPI_DATATYPE_NULL  ((MPI_Datatype)0)
#define MPI_FLOAT          ((MPI_Datatype)1)   /* element = the build's `float` (double under prelude64) */
#define MPI_FLOAT_INT      ((MPI_Datatype)2)   /* struct { float val; int rank; } */
#define MPI_MAXLOC         ((MPI_Op)1)

#ifdef __cplusplus
extern "C" {
#endif
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                  void *recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
               void *recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype,
                  MPI_Op op, MPI_Comm comm);
int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
#ifdef __cplusplus
}
#endif

/* harness-side control of the mini-MPI (not part of MPI) */
#ifdef __cplusplus
namespace shim_mpi {
void world_begin(int size);          /* called once before the rank threads start */
void thread_enter(int rank);         /* first call in each rank thread             */
int  my_rank();
int  world_size();
}
#endif
#endif
