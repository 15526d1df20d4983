/* mpi.h -- minimal single-node MPI subset for building AccFFT-style programs against libaccfft_gpu_b200
 * in an image without MPI (this one).  NOT an MPI implementation: just the calls the AccFFT GPU API and its
 * step programs use: MPI_Init/Finalize, Comm_rank/size, Cart_create/get/coords, Dims_create,
 * Comm_split/free, Barrier, Bcast, Reduce, Allreduce, Allgather, Wtime, Abort.  Ranks are processes of ONE node,
 * started by accfft_b200/bin/accfft_mpirun (fork + POSIX shared memory); without the launcher a program runs as a
 * single rank.  With a real MPI installed, put ITS mpi.h first on the include path instead: the AccFFT-compatible
 * layer (accfft_b200/csrc/accfft_gpu_api.cpp) only uses standard MPI calls.
 */
#ifndef ACCFFT_B200_MPI_SHIM_H
#define ACCFFT_B200_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

#define ACCFFT_B200_MPI_SHIM 1

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef long long MPI_Offset;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_NULL (-1)
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_UNDEFINED (-32766)

#define MPI_CHAR 1
#define MPI_BYTE 2
#define MPI_INT 3
#define MPI_LONG 4
#define MPI_LONG_LONG 5
#define MPI_FLOAT 6
#define MPI_DOUBLE 7
#define MPI_UNSIGNED 8
#define MPI_UNSIGNED_LONG 9

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_PROD 4

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Initialized(int* flag);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Dims_create(int nnodes, int ndims, int dims[]);
int MPI_Cart_create(MPI_Comm comm_old, int ndims, const int dims[], const int periods[], int reorder, MPI_Comm* comm_cart);
int MPI_Cart_get(MPI_Comm comm, int maxdims, int dims[], int periods[], int coords[]);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int coords[]);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm);
int MPI_Comm_free(MPI_Comm* comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm);

#ifdef __cplusplus
}
#endif
#endif
