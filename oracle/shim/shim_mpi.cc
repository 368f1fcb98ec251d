/* oracle/shim/shim_mpi.cc -- TEST INFRASTRUCTURE, not product code.
 *
 * Thread-based mini-MPI for the reference build: each MPI "rank" is one thread of
 * the harness process.  Implements the ten entry points the reference calls and
 * nothing else.  Collectives are rank-synchronous (a cyclic barrier), which is
 * exactly how the reference uses them (blocking, every rank, same order).
 * Compiled with the same prelude as the reference so sizeof(float) agrees.
 */
#include "mpi.h"
#include <thread>
#include <mutex>
#include <condition_variable>
#include <vector>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace {
struct World {
  int size = 1;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long generation = 0;
  std::vector<const void*> sptr;   /* per-rank pointer slots used by collectives */
  std::vector<void*>       rptr;
} W;
thread_local int t_rank = 0;

void barrier() {
  std::unique_lock<std::mutex> lk(W.m);
  unsigned long gen = W.generation;
  if (++W.arrived == W.size) { W.arrived = 0; ++W.generation; W.cv.notify_all(); }
  else W.cv.wait(lk, [&]{ return gen != W.generation; });
}
size_t elem_size(MPI_Datatype t) {
  if (t == MPI_FLOAT) return sizeof(float);
  if (t == MPI_FLOAT_INT) { struct fi { float v; int r; }; return sizeof(fi); }
  return 0;
}
}

namespace shim_mpi {
void world_begin(int size) {
  W.size = size; W.arrived = 0; W.generation = 0;
  W.sptr.assign(size, nullptr); W.rptr.assign(size, nullptr);
}
void thread_enter(int rank) { t_rank = rank; }
int my_rank() { return t_rank; }
int world_size() { return W.size; }
}

extern "C" {
int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int *size) { *size = W.size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = t_rank; return MPI_SUCCESS; }
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm) { *newcomm = comm; return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int errorcode) {
  fprintf(stderr, "shim MPI_Abort(%d) on rank %d\n", errorcode, t_rank);
  abort();
}

/* Only the reference's form is supported: sendbuf == MPI_IN_PLACE; rank r's
 * contribution already sits in recvbuf[r*recvcount ...] (mcpar.cc:131-132). */
int MPI_Allgather(const void *sendbuf, int, MPI_Datatype, void *recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm) {
  if (sendbuf != MPI_IN_PLACE) return 1;
  const size_t nb = (size_t)recvcount * elem_size(recvtype);
  W.rptr[t_rank] = recvbuf;
  barrier();
  /* pull every other rank's own slot out of that rank's buffer */
  for (int r = 0; r < W.size; ++r) {
    if (r == t_rank) continue;
    memcpy((char*)recvbuf + (size_t)r * nb, (const char*)W.rptr[r] + (size_t)r * nb, nb);
  }
  barrier();   /* nobody may overwrite its own slot until all have copied */
  return MPI_SUCCESS;
}

int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
               int, MPI_Datatype, int root, MPI_Comm) {
  const size_t nb = (size_t)sendcount * elem_size(sendtype);
  W.sptr[t_rank] = sendbuf;
  barrier();
  if (t_rank == root)
    for (int r = 0; r < W.size; ++r)
      memcpy((char*)recvbuf + (size_t)r * nb, W.sptr[r], nb);
  barrier();
  return MPI_SUCCESS;
}

/* Only MPI_FLOAT_INT / MPI_MAXLOC, count 1 (mcout.cc:107).  Ties -> lowest rank. */
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype,
                  MPI_Op op, MPI_Comm) {
  if (datatype != MPI_FLOAT_INT || op != MPI_MAXLOC || count != 1) return 1;
  struct fi { float v; int r; };
  W.sptr[t_rank] = sendbuf;
  barrier();
  fi best = *(const fi*)W.sptr[0];
  for (int r = 1; r < W.size; ++r) {
    const fi &c = *(const fi*)W.sptr[r];
    if (c.v > best.v || (c.v == best.v && c.r < best.r)) best = c;
  }
  *(fi*)recvbuf = best;
  barrier();
  return MPI_SUCCESS;
}

int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm) {
  const size_t nb = (size_t)count * elem_size(datatype);
  W.rptr[t_rank] = buffer;
  barrier();
  if (t_rank != root) memcpy(buffer, W.rptr[root], nb);
  barrier();
  return MPI_SUCCESS;
}
}
