/*
 * oracle/stub/mpi.h — in-process stand-in for <mpi.h>, TEST INFRASTRUCTURE ONLY.
 *
 * MPI is not installed in this image.  This header lets the reference's own decomp.cpp, halo.cpp,
 * boundary.cpp and init.cpp compile UNMODIFIED (oracle/Makefile) by emulating the handful of MPI
 * calls they make, with one std::thread per rank inside a single process (mini_mpi.cpp).
 * It only implements what include/{decomp,halo}.hpp, src/decomp.cpp:5-39 and src/halo.cpp:6-50 use.
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_NULL   ((MPI_Comm)0)
#define MPI_COMM_WORLD  ((MPI_Comm)1)
#define MPI_PROC_NULL   (-1)
#define MPI_SUCCESS     0
#define MPI_DOUBLE      ((MPI_Datatype)1)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Dims_create(int nnodes, int ndims, int dims[]);
int MPI_Cart_create(MPI_Comm comm_old, int ndims, const int dims[], const int periods[],
                    int reorder, MPI_Comm* comm_cart);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int coords[]);
int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int* rank_source, int* rank_dest);
int MPI_Comm_free(MPI_Comm* comm);

int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype,
                    MPI_Datatype* newtype);
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* datatype);
int MPI_Type_free(MPI_Datatype* datatype);

int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm,
              MPI_Request* request);
int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm,
              MPI_Request* request);
int MPI_Waitall(int count, MPI_Request array_of_requests[], MPI_Status array_of_statuses[]);

/* emulator control (not MPI): world size for the process, rank for the calling thread */
void mini_mpi_set_world(int size);
void mini_mpi_set_rank(int rank);
void mini_mpi_barrier(void);

#ifdef __cplusplus
}
#endif
