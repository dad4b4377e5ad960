/* Single-rank MPI stand-in for building the reference solver as a parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY. The reference (mnucci32/aither) depends on MPI for
 * moving bytes between ranks; no arithmetic lives in MPI. With one rank every
 * block connection takes the local (non-MPI) swap branch
 * (reference src/gridLevel.cpp:299-303), so point-to-point calls are never
 * reached and simply abort if they are. Collective calls degenerate to no-ops.
 */
#ifndef AITHER_B200_ORACLE_MPI_STUB_H
#define AITHER_B200_ORACLE_MPI_STUB_H
#include <cstddef>
#include <cstdio>
#include <cstdlib>

typedef int MPI_Datatype;
typedef int MPI_Comm;
typedef int MPI_Op;
typedef std::ptrdiff_t MPI_Aint;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };
typedef void(MPI_User_function)(void *, void *, int *, MPI_Datatype *);

enum {
  MPI_COMM_WORLD = 0,
  MPI_INT = 1, MPI_DOUBLE = 2, MPI_CHAR = 3, MPI_CXX_BOOL = 4, MPI_C_BOOL = 5,
  MPI_PACKED = 6, MPI_SUM = 7
};
#define MPI_IN_PLACE (reinterpret_cast<void *>(-1))
#define MPI_STATUS_IGNORE (static_cast<MPI_Status *>(nullptr))
#define MPI_SUCCESS 0

static inline int stub_mpi_unreachable(const char *what) {
  std::fprintf(stderr, "mpi stub: %s reached with a single rank\n", what);
  std::abort();
  return 1;
}

static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Get_version(int *v, int *s) { *v = 3; *s = 1; return 0; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm) { return 0; }
static inline int MPI_Scatter(const void *s, int, MPI_Datatype, void *r, int, MPI_Datatype, int, MPI_Comm) {
  *static_cast<int *>(r) = *static_cast<const int *>(s);
  return 0;
}
static inline int MPI_Type_contiguous(int, MPI_Datatype, MPI_Datatype *t) { *t = 100; return 0; }
static inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
static inline int MPI_Type_free(MPI_Datatype *) { return 0; }
static inline int MPI_Type_create_struct(int, const int *, const MPI_Aint *, const MPI_Datatype *, MPI_Datatype *t) { *t = 101; return 0; }
static inline int MPI_Type_create_resized(MPI_Datatype, MPI_Aint, MPI_Aint, MPI_Datatype *t) { *t = 102; return 0; }
static inline int MPI_Type_get_extent(MPI_Datatype, MPI_Aint *lb, MPI_Aint *ext) { *lb = 0; *ext = 0; return 0; }
static inline int MPI_Get_address(const void *p, MPI_Aint *a) { *a = reinterpret_cast<MPI_Aint>(p); return 0; }
static inline int MPI_Op_create(MPI_User_function *, int, MPI_Op *op) { *op = 200; return 0; }
static inline int MPI_Op_free(MPI_Op *) { return 0; }
static inline int MPI_Pack_size(int, MPI_Datatype, MPI_Comm, int *sz) { *sz = 0; return 0; }

static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return stub_mpi_unreachable("MPI_Send"); }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return stub_mpi_unreachable("MPI_Recv"); }
static inline int MPI_Probe(int, int, MPI_Comm, MPI_Status *) { return stub_mpi_unreachable("MPI_Probe"); }
static inline int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *) { return stub_mpi_unreachable("MPI_Get_count"); }
static inline int MPI_Sendrecv_replace(void *, int, MPI_Datatype, int, int, int, int, MPI_Comm, MPI_Status *) { return stub_mpi_unreachable("MPI_Sendrecv_replace"); }
static inline int MPI_Pack(const void *, int, MPI_Datatype, void *, int, int *, MPI_Comm) { return stub_mpi_unreachable("MPI_Pack"); }
static inline int MPI_Unpack(const void *, int, int *, void *, int, MPI_Datatype, MPI_Comm) { return stub_mpi_unreachable("MPI_Unpack"); }
#endif
