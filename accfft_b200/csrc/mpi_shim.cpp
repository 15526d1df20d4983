// mpi_shim.cpp -- implementation of include/mpi_shim/mpi.h: a single-node, process-per-rank MPI subset over POSIX
// shared memory.  Exists only because this image has no MPI; see the header for scope.  Collectives move data
// through one 64 KB slot per world rank and synchronise with a sense-reversing barrier per communicator.
#include "../../include/mpi_shim/mpi.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <vector>

namespace {

constexpr int kMaxRanks = 64;
constexpr int kMaxComms = 256;
constexpr size_t kSlot = 64 * 1024;

struct ShmBarrier {
  std::atomic<int> count;
  std::atomic<int> sense;
};
struct Shm {
  std::atomic<int> next_comm;
  ShmBarrier bar[kMaxComms];
  alignas(64) unsigned char slot[kMaxRanks][kSlot];
};

struct Comm {
  bool used = false;
  int shm_id = 0;                 // index of this communicator's barrier in shared memory
  std::vector<int> members;       // world ranks, in communicator rank order
  int rank = 0;                   // my rank in it
  int local_sense = 0;
  bool cart = false;
  int dims[2] = {1, 1}, periods[2] = {0, 0};
};

Shm* g_shm = nullptr;
Shm g_private;                     // single-rank runs (no launcher)
int g_world_rank = 0, g_world_size = 1;
bool g_init = false;
std::vector<Comm> g_comms;

Comm* get(MPI_Comm c) { return (c >= 0 && c < (int)g_comms.size() && g_comms[c].used) ? &g_comms[c] : nullptr; }

size_t type_size(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_INT: case MPI_FLOAT: case MPI_UNSIGNED: return 4;
    default: return 8;
  }
}

void barrier(Comm* c) {
  const int n = (int)c->members.size();
  if (n == 1) return;
  ShmBarrier& b = g_shm->bar[c->shm_id];
  c->local_sense ^= 1;
  if (b.count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
    b.count.store(0, std::memory_order_relaxed);
    b.sense.store(c->local_sense, std::memory_order_release);
  } else {
    int spins = 0;
    while (b.sense.load(std::memory_order_acquire) != c->local_sense)
      if (++spins > 2000) usleep(50);
  }
}

template <class T>
void combine(T* acc, const T* v, int n, MPI_Op op) {
  for (int i = 0; i < n; ++i) {
    switch (op) {
      case MPI_SUM: acc[i] += v[i]; break;
      case MPI_PROD: acc[i] *= v[i]; break;
      case MPI_MAX: acc[i] = std::max(acc[i], v[i]); break;
      default: acc[i] = std::min(acc[i], v[i]); break;
    }
  }
}
void combine_any(void* acc, const void* v, int n, MPI_Datatype t, MPI_Op op) {
  switch (t) {
    case MPI_DOUBLE: combine((double*)acc, (const double*)v, n, op); break;
    case MPI_FLOAT: combine((float*)acc, (const float*)v, n, op); break;
    case MPI_INT: combine((int*)acc, (const int*)v, n, op); break;
    case MPI_UNSIGNED: combine((unsigned*)acc, (const unsigned*)v, n, op); break;
    case MPI_UNSIGNED_LONG: combine((unsigned long*)acc, (const unsigned long*)v, n, op); break;
    case MPI_CHAR: case MPI_BYTE: combine((unsigned char*)acc, (const unsigned char*)v, n, op); break;
    default: combine((long long*)acc, (const long long*)v, n, op); break;
  }
}

// every member contributes `bytes` bytes; on return out[r*bytes ...] holds member r's contribution (chunked)
void allgather_bytes(Comm* c, const void* mine, size_t bytes, void* out) {
  const int n = (int)c->members.size();
  if (n == 1) {
    if (out != mine) memmove(out, mine, bytes);
    return;
  }
  for (size_t off = 0; off < bytes || off == 0; off += kSlot) {
    const size_t len = std::min(kSlot, bytes - off);
    memcpy(g_shm->slot[g_world_rank], (const char*)mine + off, len);
    barrier(c);
    for (int r = 0; r < n; ++r) memcpy((char*)out + (size_t)r * bytes + off, g_shm->slot[c->members[r]], len);
    barrier(c);
    if (bytes == 0) break;
  }
}

int new_comm_slot() {
  for (size_t i = 0; i < g_comms.size(); ++i)
    if (!g_comms[i].used) return (int)i;
  g_comms.emplace_back();
  return (int)g_comms.size() - 1;
}

}  // namespace

extern "C" {

int MPI_Init(int*, char***) {
  if (g_init) return MPI_SUCCESS;
  const char* name = getenv("ACCFFT_MPI_SHM");
  const char* sz = getenv("ACCFFT_MPI_SIZE");
  const char* rk = getenv("ACCFFT_MPI_RANK");
  if (name && sz && rk) {
    g_world_size = atoi(sz);
    g_world_rank = atoi(rk);
    if (g_world_size < 1 || g_world_size > kMaxRanks || g_world_rank < 0 || g_world_rank >= g_world_size) {
      fprintf(stderr, "mpi_shim: bad ACCFFT_MPI_SIZE / ACCFFT_MPI_RANK\n");
      return MPI_ERR_OTHER;
    }
    const int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) {
      perror("mpi_shim: shm_open");
      return MPI_ERR_OTHER;
    }
    void* p = mmap(nullptr, sizeof(Shm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
      perror("mpi_shim: mmap");
      return MPI_ERR_OTHER;
    }
    g_shm = (Shm*)p;   // zero-filled by the launcher (ftruncate): next_comm = 0, barriers idle
  } else {
    g_world_size = 1;
    g_world_rank = 0;
    g_shm = &g_private;
  }
  g_comms.clear();
  g_comms.emplace_back();
  Comm& w = g_comms[0];
  w.used = true;
  w.shm_id = 0;
  for (int r = 0; r < g_world_size; ++r) w.members.push_back(r);
  w.rank = g_world_rank;
  // slot 0 of the barrier pool belongs to MPI_COMM_WORLD; the pool counter starts after it
  int expected = 0;
  g_shm->next_comm.compare_exchange_strong(expected, 1);
  g_init = true;
  barrier(&w);
  return MPI_SUCCESS;
}

int MPI_Initialized(int* flag) {
  *flag = g_init ? 1 : 0;
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  if (g_init) barrier(&g_comms[0]);
  g_init = false;
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int errorcode) {
  fprintf(stderr, "mpi_shim: MPI_Abort(%d) on rank %d\n", errorcode, g_world_rank);
  _exit(errorcode ? errorcode : 1);
}

double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Comm_rank(MPI_Comm comm, int* rank) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  *rank = c->rank;
  return MPI_SUCCESS;
}

int MPI_Comm_size(MPI_Comm comm, int* size) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  *size = (int)c->members.size();
  return MPI_SUCCESS;
}

// balanced factorisation, non-increasing (what MPI_Dims_create specifies): 2 -> 2x1, 4 -> 2x2, 8 -> 4x2
int MPI_Dims_create(int nnodes, int ndims, int dims[]) {
  if (ndims != 2) {
    if (ndims == 1) {
      dims[0] = nnodes;
      return MPI_SUCCESS;
    }
    return MPI_ERR_OTHER;
  }
  if (dims[0] > 0 && dims[1] > 0) return MPI_SUCCESS;
  if (dims[0] > 0) {
    dims[1] = nnodes / dims[0];
    return MPI_SUCCESS;
  }
  if (dims[1] > 0) {
    dims[0] = nnodes / dims[1];
    return MPI_SUCCESS;
  }
  int a = 1;
  for (int d = 1; d * d <= nnodes; ++d)
    if (nnodes % d == 0) a = d;
  dims[0] = nnodes / a;
  dims[1] = a;
  return MPI_SUCCESS;
}

int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  const int n = (int)c->members.size();
  struct Item { int color, key, world, shm_id; };
  std::vector<Item> all(n);
  Item mine = {color, key, g_world_rank, -1};
  allgather_bytes(c, &mine, sizeof(Item), all.data());
  // my group, ordered by (key, parent rank)
  std::vector<int> idx;
  for (int r = 0; r < n; ++r)
    if (all[r].color == color) idx.push_back(r);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return all[a].key < all[b].key; });
  // the first member of each group allocates the group's barrier from the shared pool and publishes it
  if (color != MPI_UNDEFINED && c->members[idx[0]] == g_world_rank) {
    mine.shm_id = g_shm->next_comm.fetch_add(1);
    if (mine.shm_id >= kMaxComms) {
      fprintf(stderr, "mpi_shim: too many communicators\n");
      _exit(2);
    }
    g_shm->bar[mine.shm_id].count.store(0);
    g_shm->bar[mine.shm_id].sense.store(0);
  }
  allgather_bytes(c, &mine, sizeof(Item), all.data());
  if (color == MPI_UNDEFINED) {
    *newcomm = MPI_COMM_NULL;
    return MPI_SUCCESS;
  }
  const int slot = new_comm_slot();
  Comm nc;
  nc.used = true;
  nc.shm_id = all[idx[0]].shm_id;
  for (size_t i = 0; i < idx.size(); ++i) {
    nc.members.push_back(all[idx[i]].world);
    if (all[idx[i]].world == g_world_rank) nc.rank = (int)i;
  }
  g_comms[slot] = nc;
  *newcomm = slot;
  return MPI_SUCCESS;
}

int MPI_Cart_create(MPI_Comm comm_old, int ndims, const int dims[], const int periods[], int, MPI_Comm* comm_cart) {
  Comm* c = get(comm_old);
  if (!c || ndims != 2 || dims[0] * dims[1] != (int)c->members.size()) return MPI_ERR_OTHER;
  const int rc = MPI_Comm_split(comm_old, 0, c->rank, comm_cart);   // same ranks, row-major coordinates
  if (rc != MPI_SUCCESS) return rc;
  Comm* n = get(*comm_cart);
  n->cart = true;
  n->dims[0] = dims[0];
  n->dims[1] = dims[1];
  n->periods[0] = periods ? periods[0] : 0;
  n->periods[1] = periods ? periods[1] : 0;
  return MPI_SUCCESS;
}

int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int coords[]) {
  Comm* c = get(comm);
  if (!c || !c->cart || maxdims < 2) return MPI_ERR_OTHER;
  coords[0] = rank / c->dims[1];
  coords[1] = rank % c->dims[1];
  return MPI_SUCCESS;
}

int MPI_Cart_get(MPI_Comm comm, int maxdims, int dims[], int periods[], int coords[]) {
  Comm* c = get(comm);
  if (!c || !c->cart || maxdims < 2) return MPI_ERR_OTHER;
  dims[0] = c->dims[0];
  dims[1] = c->dims[1];
  periods[0] = c->periods[0];
  periods[1] = c->periods[1];
  return MPI_Cart_coords(comm, c->rank, maxdims, coords);
}

int MPI_Comm_free(MPI_Comm* comm) {
  Comm* c = get(*comm);
  if (!c || *comm == MPI_COMM_WORLD) return MPI_ERR_OTHER;
  barrier(c);
  c->used = false;
  c->members.clear();
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  barrier(c);
  return MPI_SUCCESS;
}

int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int, MPI_Datatype, MPI_Comm comm) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  allgather_bytes(c, sendbuf, (size_t)sendcount * type_size(sendtype), recvbuf);
  return MPI_SUCCESS;
}

int MPI_Bcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  const size_t bytes = (size_t)count * type_size(datatype);
  std::vector<char> all(bytes * c->members.size());
  allgather_bytes(c, buffer, bytes, all.data());
  memcpy(buffer, all.data() + (size_t)root * bytes, bytes);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  const size_t bytes = (size_t)count * type_size(datatype);
  std::vector<char> all(bytes * c->members.size());
  allgather_bytes(c, sendbuf, bytes, all.data());
  memcpy(recvbuf, all.data(), bytes);
  for (size_t r = 1; r < c->members.size(); ++r) combine_any(recvbuf, all.data() + r * bytes, count, datatype, op);
  return MPI_SUCCESS;
}

int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  const size_t bytes = (size_t)count * type_size(datatype);
  std::vector<char> tmp(bytes);
  const int rc = MPI_Allreduce(sendbuf, tmp.data(), count, datatype, op, comm);
  if (rc == MPI_SUCCESS && c->rank == root) memcpy(recvbuf, tmp.data(), bytes);
  return rc;
}

}  // extern "C"
