/* mpi.h -- minimal single-node MPI subset for building AccFFT-style programs against libaccfft_gpu_b200
 * in an image without MPI (this one).  NOT an MPI implementation: just the calls the AccFFT GPU API and its
 * step programs use: MPI_Init/Finalize, Comm_rank/size, Cart_create/get/coords, Dims_create,
 * Comm_split/free, Barrier, Bcast, Reduce, Allreduce, Allgather, Wtime, Abort -- plus the point-to-point, all-to-all
 * and derived-datatype calls the reference's CPU transposes use (src/transpose.cpp, include/parUtils.txx: Isend / Irecv /
 * Wait[all] / Test / Sendrecv, Alltoall[v|w], Type_contiguous / vector / hvector), so that oracle/Makefile can build the
 * reference's own CPU library against this header as a parity pin.  Ranks are processes of ONE node,
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
typedef int MPI_Request;
typedef int MPI_Info;
typedef long MPI_Aint;

#define MPI_COMM_NULL (-1)
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_UNDEFINED (-32766)
#define MPI_REQUEST_NULL (-1)
#define MPI_INFO_NULL 0
#define MPI_ANY_TAG (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_DATATYPE_NULL 0

#define MPI_CHAR 1
#define MPI_BYTE 2
#define MPI_INT 3
#define MPI_LONG 4
#define MPI_LONG_LONG 5
#define MPI_FLOAT 6
#define MPI_DOUBLE 7
#define MPI_UNSIGNED 8
#define MPI_UNSIGNED_LONG 9
#define MPI_SHORT 10
#define MPI_UNSIGNED_SHORT 11
#define MPI_LONG_DOUBLE 12
#define MPI_LONG_LONG_INT MPI_LONG_LONG
#define MPI_UNSIGNED_CHAR 13
/* (value, index) pair types: named by the reference's type traits (include/dtypes.h), never communicated */
#define MPI_FLOAT_INT 14
#define MPI_DOUBLE_INT 15
#define MPI_LONG_INT 16
#define MPI_SHORT_INT 17
#define MPI_2INT 18
#define MPI_LONG_DOUBLE_INT 19

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
/* point-to-point: eager, buffered (a send completes once its payload sits in the sender's shared outbox) */
int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Issend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Send(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm);
int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Recv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Wait(MPI_Request* request, MPI_Status* status);
int MPI_Waitall(int count, MPI_Request requests[], MPI_Status statuses[]);
int MPI_Test(MPI_Request* request, int* flag, MPI_Status* status);
int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag, void* recvbuf,
                 int recvcount, MPI_Datatype recvtype, int source, int recvtag, MPI_Comm comm, MPI_Status* status);
int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoallv(const void* sendbuf, const int sendcounts[], const int sdispls[], MPI_Datatype sendtype, void* recvbuf,
                  const int recvcounts[], const int rdispls[], MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoallw(const void* sendbuf, const int sendcounts[], const int sdispls[], const MPI_Datatype sendtypes[],
                  void* recvbuf, const int recvcounts[], const int rdispls[], const MPI_Datatype recvtypes[], MPI_Comm comm);
/* derived datatypes (process-local handles) */
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_hvector(int count, int blocklength, MPI_Aint stride, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_create_struct(int count, const int blocklengths[], const MPI_Aint displacements[], const MPI_Datatype types[],
                           MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* datatype);
int MPI_Type_free(MPI_Datatype* datatype);
int MPI_Type_size(MPI_Datatype datatype, int* size);

/* user-defined reduction operators */
typedef void MPI_User_function(void* invec, void* inoutvec, int* len, MPI_Datatype* datatype);
int MPI_Op_create(MPI_User_function* user_fn, int commute, MPI_Op* op);
int MPI_Op_free(MPI_Op* op);

#ifdef __cplusplus
}
#endif
#endif
