/* Single-process stand-in for <mpi.h> for the Tier-B reference build (oracle/ref_tier_b.cpp): the reference's
 * gradient / reconstruction sources only include utilities/mpiutils.hpp, they make no MPI call. TEST INFRASTRUCTURE ONLY. */
#ifndef FVENS_B200_MPI_LITE
#define FVENS_B200_MPI_LITE
#include <cstddef>
typedef int MPI_Comm;
typedef int MPI_Op;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_DOUBLE 0
#define MPI_INT 1
#define MPI_SUCCESS 0
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-1)
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)1)
/// one process plays every rank in turn: the harness sets these before calling rank-dependent reference code
static int mpi_lite_size = 1, mpi_lite_rank = 0;
static inline int MPI_Comm_size(MPI_Comm, int *size) { *size = mpi_lite_size; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
#include <chrono>
static inline double MPI_Wtime() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline int MPI_Wait(MPI_Request*, MPI_Status*) { return 0; }
static inline int MPI_Waitall(int, MPI_Request*, MPI_Status*) { return 0; }
static inline int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { return 0; }
static inline int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { return 0; }
/// single process: the reduction of one contribution is the contribution (in place, or copied)
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
	if(s != MPI_IN_PLACE) { const size_t sz = (t == MPI_DOUBLE ? sizeof(double) : sizeof(int))*(size_t)n; const char *a = (const char*)s; char *b = (char*)r; for(size_t i = 0; i < sz; i++) b[i] = a[i]; }
	return 0;
}
static inline int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = mpi_lite_rank; return 0; }
static inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return 0; }
#endif
