// accfft_mpirun -- launcher for programs built against include/mpi_shim/mpi.h (this image has no MPI):
//     accfft_mpirun -n N [--gpus-per-node G] program [args...]
// creates the shared-memory segment of the shim, starts N ranks of `program` on this node and waits for them.
// Each rank gets CUDA_VISIBLE_DEVICES = rank % G unless the variable is already set or G = 0, so that -- like in the
// reference's step programs, which never call cudaSetDevice -- "device 0" is a different GPU for every rank.
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <string>
#include <vector>
#include "mpi_shim_layout.hpp"

int main(int argc, char** argv) {
  int n = 1, gpus = -1, i = 1;
  for (; i < argc; ++i) {
    if (!strcmp(argv[i], "-n") || !strcmp(argv[i], "-np")) n = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--gpus-per-node")) gpus = atoi(argv[++i]);
    else break;
  }
  if (i >= argc || n < 1 || n > 64) {
    fprintf(stderr, "usage: accfft_mpirun -n N [--gpus-per-node G] program [args...]\n");
    return 2;
  }
  if (gpus < 0) gpus = n;   // one GPU per rank by default
  const std::string name = "/accfft_mpi_" + std::to_string((long)getpid());
  const size_t bytes = sizeof(accfft_mpi_shim::ShimShm);   // sparse: pages are committed on first touch
  const int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) {
    perror("accfft_mpirun: shared memory");
    return 2;
  }
  close(fd);
  std::vector<pid_t> kids;
  for (int r = 0; r < n; ++r) {
    const pid_t pid = fork();
    if (pid == 0) {
      setenv("ACCFFT_MPI_SHM", name.c_str(), 1);
      setenv("ACCFFT_MPI_SIZE", std::to_string(n).c_str(), 1);
      setenv("ACCFFT_MPI_RANK", std::to_string(r).c_str(), 1);
      if (gpus > 0 && !getenv("CUDA_VISIBLE_DEVICES")) setenv("CUDA_VISIBLE_DEVICES", std::to_string(r % gpus).c_str(), 1);
      execvp(argv[i], argv + i);
      perror("accfft_mpirun: exec");
      _exit(127);
    }
    kids.push_back(pid);
  }
  int worst = 0;
  for (size_t k = 0; k < kids.size(); ++k) {
    int st = 0;
    waitpid(kids[k], &st, 0);
    const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
    if (code && !worst) {
      worst = code;
      for (pid_t other : kids) kill(other, SIGTERM);   // one rank failed: do not leave the others spinning
    }
  }
  shm_unlink(name.c_str());
  return worst;
}
