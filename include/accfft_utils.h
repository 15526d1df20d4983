/* accfft_utils.h -- host-side I/O neighbour of the transform path.
 * The reference's utils library reads / writes distributed fields through Parallel-NetCDF
 * (include/accfft_utils.h:24-38: read_pnetcdf / write_pnetcdf, compiled only with USE_PNETCDF; PnetCDF does not exist in
 * this image).  The same two calls are provided here for the NumPy `.npy` container instead -- same argument list, same
 * collective semantics (every rank of c_comm passes the starts / counts of ITS block of the global gsizes[0] x gsizes[1] x
 * gsizes[2] array of doubles and its HOST data) -- which is what the parity tests and debugging sessions of this repo
 * read with numpy.load.  One file holds the whole global array in C order; ranks write / read their rows with pwrite / pread
 * at computed offsets (single node or any shared file system).  Returns nothing, like the reference; errors go to stderr
 * and accfft_npy_last_error().
 */
#ifndef ACCFFT_B200_UTILS_H
#define ACCFFT_B200_UTILS_H
#include <mpi.h>
#include <string>

void write_npy(const std::string& filename, MPI_Offset starts[3], MPI_Offset counts[3], MPI_Comm c_comm, int gsizes[3],
               double* localData);
void read_npy(const std::string& filename, MPI_Offset starts[3], MPI_Offset counts[3], MPI_Comm c_comm, int gsizes[3],
              double* localData);
/* 0 = the last read_npy / write_npy of this process succeeded */
int accfft_npy_last_error(void);

#endif
