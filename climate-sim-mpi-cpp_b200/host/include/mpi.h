/*
 * mpi.h — the slice of MPI that the reference's headers, driver and tests name on the timestep
 * path, for builds where the GPU library replaces MPI as the transport (one process per GPU).
 *
 * The reference includes <mpi.h> from include/decomp.hpp:2 and include/halo.hpp:2 and calls, outside
 * the bodies of decomp.cpp/halo.cpp (which this build replaces): MPI_Init, MPI_Init_thread,
 * MPI_Initialized, MPI_Finalized, MPI_Finalize, MPI_Comm_rank, MPI_Comm_size, MPI_Barrier,
 * MPI_Wtime, MPI_Reduce (src/main.cpp:24-28,82,89,127-128,136 and tests/simulation/unit).
 *
 * Rank and size come from the launcher's environment (CSIM_RANK/CSIM_WORLD_SIZE, else
 * RANK/WORLD_SIZE as set by torchrun, else 0/1).  With more than one rank MPI_Init bootstraps an
 * NCCL communicator on the process-wide csim context through a rendezvous file
 * ($CSIM_RENDEZVOUS, default /tmp/csim_rendezvous_<MASTER_PORT>); exchange_halos and MPI_Reduce
 * then run over NVLink.  A real MPI can be used instead by simply not putting this directory on
 * the include path ahead of it and giving csim_comm_init an id broadcast with MPI_Bcast.
 */
#ifndef CSIM_MPI_SHIM_H
#define CSIM_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_COMM_WORLD ((MPI_Comm)1)
#define MPI_PROC_NULL (-1)
#define MPI_INFO_NULL ((MPI_Info)0)
#define MPI_DOUBLE ((MPI_Datatype)1)
#define MPI_INT ((MPI_Datatype)2)
#define MPI_MAX ((MPI_Op)1)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1

int MPI_Init(int* argc, char*** argv);
int MPI_Init_thread(int* argc, char*** argv, int required, int* provided);
int MPI_Initialized(int* flag);
int MPI_Finalized(int* flag);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Barrier(MPI_Comm comm);
double MPI_Wtime(void);
/* MPI_DOUBLE with MPI_MAX only (what src/main.cpp:127-128 uses); every rank receives the result. */
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root,
               MPI_Comm comm);

#ifdef __cplusplus
}
#endif
#endif /* CSIM_MPI_SHIM_H */
