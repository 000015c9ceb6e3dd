/* Stub <mpi.h>: the reference's utils.h:1 includes <mpi.h> and utils.c's MPI twins
 * reference these names; none of them is on the sequential path the oracle uses.
 * Test infrastructure only (see oracle/Makefile). */
#ifndef ORACLE_STUB_MPI_H
#define ORACLE_STUB_MPI_H
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Group;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_LONG_DOUBLE 1
#define MPI_INT 2
static inline int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
static inline int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { return 0; }
#endif
