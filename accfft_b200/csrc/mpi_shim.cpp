// mpi_shim.cpp -- implementation of include/mpi_shim/mpi.h: a single-node, process-per-rank MPI subset over POSIX
// shared memory.  Exists only because this image has no MPI; see the header for scope.
//   collectives     one 64 KB staging slot per world rank + a sense-reversing barrier per communicator
//   point-to-point  eager and buffered: a send copies its (packed) payload into the sender's outbox and posts a
//                   descriptor into the FIFO of the (source, destination) pair; a receive completes in MPI_Wait / Test by
//                   taking the first descriptor of that FIFO with a matching (communicator, tag)
//   all-to-all      built on the point-to-point layer
//   datatypes       basic types + contiguous / vector / hvector, flattened to (offset, length) runs per element
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
#include <deque>
#include <vector>

#include "mpi_shim_layout.hpp"

namespace {
using namespace accfft_mpi_shim;

struct Comm {
  bool used = false;
  int shm_id = 0;                 // index of this communicator's barrier in shared memory
  std::vector<int> members;       // world ranks, in communicator rank order
  int rank = 0;                   // my rank in it
  int local_sense = 0;
  bool cart = false;
  int dims[2] = {1, 1}, periods[2] = {0, 0};
};

ShimShm* g_shm = nullptr;
int g_world_rank = 0, g_world_size = 1;
bool g_init = false;
std::vector<Comm> g_comms;

Comm* get(MPI_Comm c) { return (c >= 0 && c < (int)g_comms.size() && g_comms[c].used) ? &g_comms[c] : nullptr; }

[[noreturn]] void die(const char* what) {
  fprintf(stderr, "mpi_shim (rank %d): %s\n", g_world_rank, what);
  _exit(3);
}

// ---- datatypes -------------------------------------------------------------------------------------------------
struct Run { long off; size_t len; };
struct Type {
  bool used = false;
  std::vector<Run> runs;          // byte runs of ONE element, relative to its origin
  long extent = 0;                // distance between consecutive elements
  size_t size = 0;                // payload bytes of one element
};
constexpr int kFirstDerived = 100;
std::vector<Type> g_types;

size_t basic_size(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: case MPI_UNSIGNED_CHAR: return 1;
    case MPI_SHORT: case MPI_UNSIGNED_SHORT: return 2;
    case MPI_INT: case MPI_FLOAT: case MPI_UNSIGNED: return 4;
    case MPI_LONG_DOUBLE: return sizeof(long double);
    default: return 8;
  }
}
Type type_of(MPI_Datatype t) {
  if (t >= kFirstDerived) {
    const size_t i = (size_t)(t - kFirstDerived);
    if (i >= g_types.size() || !g_types[i].used) die("invalid derived datatype");
    return g_types[i];
  }
  Type b;
  b.used = true;
  b.size = basic_size(t);
  b.extent = (long)b.size;
  b.runs.push_back({0, b.size});
  return b;
}
size_t type_size(MPI_Datatype t) { return t >= kFirstDerived ? type_of(t).size : basic_size(t); }

void add_run(std::vector<Run>& v, long off, size_t len) {
  if (!len) return;
  if (!v.empty() && v.back().off + (long)v.back().len == off) v.back().len += len;
  else v.push_back({off, len});
}
MPI_Datatype register_type(const Type& t) {
  for (size_t i = 0; i < g_types.size(); ++i)
    if (!g_types[i].used) {
      g_types[i] = t;
      return kFirstDerived + (int)i;
    }
  g_types.push_back(t);
  return kFirstDerived + (int)g_types.size() - 1;
}
// count blocks of blocklength old elements, block b at byte offset b * stride_bytes
MPI_Datatype make_vector(int count, int blocklength, long stride_bytes, MPI_Datatype oldtype) {
  const Type o = type_of(oldtype);
  Type n;
  n.used = true;
  for (int b = 0; b < count; ++b)
    for (int e = 0; e < blocklength; ++e)
      for (const Run& r : o.runs) add_run(n.runs, (long)b * stride_bytes + (long)e * o.extent + r.off, r.len);
  n.size = (size_t)count * blocklength * o.size;
  n.extent = count > 0 ? (long)(count - 1) * stride_bytes + (long)blocklength * o.extent : 0;
  return register_type(n);
}
void pack(const void* buf, int count, const Type& t, unsigned char* out) {
  for (int i = 0; i < count; ++i)
    for (const Run& r : t.runs) {
      memcpy(out, (const char*)buf + (long)i * t.extent + r.off, r.len);
      out += r.len;
    }
}
void unpack(void* buf, int count, const Type& t, const unsigned char* in, size_t avail) {
  for (int i = 0; i < count; ++i)
    for (const Run& r : t.runs) {
      const size_t len = std::min(r.len, avail);
      memcpy((char*)buf + (long)i * t.extent + r.off, in, len);
      in += len;
      avail -= len;
      if (!avail) return;
    }
}

// ---- barrier / staged collectives --------------------------------------------------------------------------------
void backoff(int& spins) {
  if (++spins > 2000) usleep(50);
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
    while (b.sense.load(std::memory_order_acquire) != c->local_sense) backoff(spins);
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
constexpr int kFirstUserOp = 100;
std::vector<MPI_User_function*> g_user_ops;

void combine_any(void* acc, const void* v, int n, MPI_Datatype t, MPI_Op op) {
  if (op >= kFirstUserOp) {   // MPI_Op_create: inout = in (op) inout
    if ((size_t)(op - kFirstUserOp) >= g_user_ops.size() || !g_user_ops[op - kFirstUserOp]) die("invalid user-defined MPI_Op");
    g_user_ops[op - kFirstUserOp](const_cast<void*>(v), acc, &n, &t);
    return;
  }
  switch (t) {
    case MPI_DOUBLE: combine((double*)acc, (const double*)v, n, op); break;
    case MPI_FLOAT: combine((float*)acc, (const float*)v, n, op); break;
    case MPI_INT: combine((int*)acc, (const int*)v, n, op); break;
    case MPI_UNSIGNED: combine((unsigned*)acc, (const unsigned*)v, n, op); break;
    case MPI_UNSIGNED_LONG: combine((unsigned long*)acc, (const unsigned long*)v, n, op); break;
    case MPI_CHAR: case MPI_BYTE: case MPI_UNSIGNED_CHAR: combine((unsigned char*)acc, (const unsigned char*)v, n, op); break;
    case MPI_SHORT: combine((short*)acc, (const short*)v, n, op); break;
    case MPI_UNSIGNED_SHORT: combine((unsigned short*)acc, (const unsigned short*)v, n, op); break;
    case MPI_LONG_DOUBLE: combine((long double*)acc, (const long double*)v, n, op); break;
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

// shared communicator ids are recycled: MPI_Comm_free hands the id back once every member has left its last barrier
int alloc_shm_id() {
  for (int i = 1; i < kMaxComms; ++i) {
    int expected = 0;
    if (g_shm->comm_used[i].compare_exchange_strong(expected, 1)) {
      g_shm->bar[i].count.store(0);
      g_shm->bar[i].sense.store(0);
      g_shm->bar[i].leavers.store(0);
      return i;
    }
  }
  return -1;
}

// ---- point-to-point ------------------------------------------------------------------------------------------------
struct Req {
  bool used = false;
  bool recv = false;
  bool done = false;
  void* buf = nullptr;
  int count = 0;
  MPI_Datatype type = MPI_BYTE;
  int src_world = 0, tag = 0, comm_id = 0, src_rank = 0;
};
std::vector<Req> g_reqs;
// This rank's outbox is a ring: payloads are allocated in order and freed when the receiver has consumed them.
struct Pending { ShmMsg* m; size_t off, need; bool done; };
std::deque<Pending> g_pending;              // my in-flight messages, in allocation order
size_t g_out_head = 0;                      // next allocation position
unsigned g_tail[kMaxRanks] = {0};           // per source: first FIFO slot not yet known to be consumed

void reclaim() {
  for (Pending& e : g_pending)
    if (!e.done && e.m->state.load(std::memory_order_acquire) == 2) {
      e.done = true;
      e.m->state.store(0, std::memory_order_release);   // descriptor slot reusable
    }
  while (!g_pending.empty() && g_pending.front().done) g_pending.pop_front();
}
// returns the outbox offset of `need` contiguous bytes, waiting for receivers to drain the ring if necessary
size_t outbox_alloc(size_t need) {
  int spins = 0;
  const double t0 = MPI_Wtime();
  for (;;) {
    reclaim();
    if (g_pending.empty()) {
      g_out_head = need;
      return 0;
    }
    const size_t tail = g_pending.front().off;
    if (g_out_head > tail) {   // live data is [tail, head): free space at the end, then at the start (head == tail: full)
      if (g_out_head + need <= kOutbox) {
        const size_t off = g_out_head;
        g_out_head += need;
        return off;
      }
      if (need < tail) {
        g_out_head = need;
        return 0;
      }
    } else if (g_out_head < tail && g_out_head + need < tail) {   // wrapped: free space is [head, tail)
      const size_t off = g_out_head;
      g_out_head += need;
      return off;
    }
    backoff(spins);
    if ((spins & 0xfff) == 0 && MPI_Wtime() - t0 > 120.0) die("outbox full for 120 s (a peer is not receiving)");
  }
}

int new_req() {
  for (size_t i = 0; i < g_reqs.size(); ++i)
    if (!g_reqs[i].used) {
      g_reqs[i] = Req();
      g_reqs[i].used = true;
      return (int)i;
    }
  g_reqs.emplace_back();
  g_reqs.back().used = true;
  return (int)g_reqs.size() - 1;
}

int do_send(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, Comm* c) {
  if (dest < 0 || dest >= (int)c->members.size()) return MPI_ERR_OTHER;
  const Type t = type_of(datatype);
  const size_t bytes = (size_t)count * t.size;
  if (bytes + 128 > kOutbox / 2) die("message larger than the shim's outbox (kOutbox in mpi_shim_layout.hpp)");
  const size_t need = ((bytes + 63) & ~(size_t)63) + 64;   // never zero, so that ring positions of live messages differ
  const int dw = c->members[dest];
  ShmChan& ch = g_shm->chan[g_world_rank][dw];
  const unsigned h = ch.head.load(std::memory_order_relaxed);
  ShmMsg& m = ch.q[h % kQueue];
  int spins = 0;
  while (m.state.load(std::memory_order_acquire) != 0) {   // FIFO full: wait for the receiver to consume the oldest entry
    reclaim();
    backoff(spins);
  }
  const size_t off = outbox_alloc(need);
  pack(buf, count, t, g_shm->outbox[g_world_rank] + off);
  m.tag = tag;
  m.comm = c->shm_id;
  m.bytes = bytes;
  m.off = off;
  g_pending.push_back({&m, off, need, false});
  m.state.store(1, std::memory_order_release);
  ch.head.store(h + 1, std::memory_order_release);
  return MPI_SUCCESS;
}

// try to complete a receive: first ready descriptor from its source with matching (communicator, tag)
bool try_recv(Req& r, MPI_Status* st) {
  ShmChan& ch = g_shm->chan[r.src_world][g_world_rank];
  const unsigned head = ch.head.load(std::memory_order_acquire);
  unsigned& tail = g_tail[r.src_world];
  while (tail != head && ch.q[tail % kQueue].state.load(std::memory_order_acquire) != 1) ++tail;   // skip consumed
  for (unsigned i = tail; i != head; ++i) {
    ShmMsg& m = ch.q[i % kQueue];
    if (m.state.load(std::memory_order_acquire) != 1) continue;
    if (m.comm != r.comm_id || (r.tag != MPI_ANY_TAG && m.tag != r.tag)) continue;
    const Type t = type_of(r.type);
    unpack(r.buf, r.count, t, g_shm->outbox[r.src_world] + m.off, (size_t)m.bytes);
    if (st) {
      st->MPI_SOURCE = r.src_rank;
      st->MPI_TAG = m.tag;
      st->MPI_ERROR = MPI_SUCCESS;
    }
    m.state.store(2, std::memory_order_release);
    r.done = true;
    return true;
  }
  return false;
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
    void* p = mmap(nullptr, sizeof(ShimShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
      perror("mpi_shim: mmap");
      return MPI_ERR_OTHER;
    }
    g_shm = (ShimShm*)p;   // zero-filled by the launcher (ftruncate): ids free, barriers idle, FIFOs empty
  } else {
    // single rank, no launcher: a private zero-filled mapping of the same layout
    g_world_size = 1;
    g_world_rank = 0;
    void* p = mmap(nullptr, sizeof(ShimShm), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) {
      perror("mpi_shim: mmap");
      return MPI_ERR_OTHER;
    }
    g_shm = (ShimShm*)p;
  }
  g_comms.clear();
  g_comms.emplace_back();
  Comm& w = g_comms[0];
  w.used = true;
  w.shm_id = 0;   // id 0 of the shared pool belongs to MPI_COMM_WORLD
  for (int r = 0; r < g_world_size; ++r) w.members.push_back(r);
  w.rank = g_world_rank;
  g_shm->comm_used[0].store(1);
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
  // the first member of each group takes a free id from the shared pool and publishes it (-2 = pool exhausted)
  if (color != MPI_UNDEFINED && c->members[idx[0]] == g_world_rank) {
    mine.shm_id = alloc_shm_id();
    if (mine.shm_id < 0) mine.shm_id = -2;
  }
  allgather_bytes(c, &mine, sizeof(Item), all.data());
  bool exhausted = false;
  for (int r = 0; r < n; ++r) exhausted = exhausted || all[r].shm_id == -2;
  if (exhausted) {   // every member of the parent sees the same verdict: fail together, give back what was taken
    if (mine.shm_id > 0) g_shm->comm_used[mine.shm_id].store(0);
    fprintf(stderr, "mpi_shim: too many live communicators (%d)\n", kMaxComms);
    *newcomm = MPI_COMM_NULL;
    return MPI_ERR_OTHER;
  }
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
  // the LAST member to get here has, by construction, left the barrier after everyone else: it recycles the shared id
  const int n = (int)c->members.size();
  ShmBarrier& b = g_shm->bar[c->shm_id];
  if (b.leavers.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
    b.leavers.store(0, std::memory_order_relaxed);
    g_shm->comm_used[c->shm_id].store(0, std::memory_order_release);
  }
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

// ---- point-to-point ------------------------------------------------------------------------------------------------
int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, MPI_Request* request) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  const int rc = do_send(buf, count, datatype, dest, tag, c);
  if (rc != MPI_SUCCESS) return rc;
  const int r = new_req();
  g_reqs[r].done = true;   // buffered: complete as soon as the payload is in the outbox
  *request = r;
  return MPI_SUCCESS;
}
int MPI_Issend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, MPI_Request* request) {
  return MPI_Isend(buf, count, datatype, dest, tag, comm, request);
}
int MPI_Send(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  return do_send(buf, count, datatype, dest, tag, c);
}

int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Request* request) {
  Comm* c = get(comm);
  if (!c || source < 0 || source >= (int)c->members.size()) return MPI_ERR_OTHER;
  const int r = new_req();
  Req& q = g_reqs[r];
  q.recv = true;
  q.buf = buf;
  q.count = count;
  q.type = datatype;
  q.src_rank = source;
  q.src_world = c->members[source];
  q.tag = tag;
  q.comm_id = c->shm_id;
  *request = r;
  return MPI_SUCCESS;
}

int MPI_Test(MPI_Request* request, int* flag, MPI_Status* status) {
  if (*request == MPI_REQUEST_NULL) {
    *flag = 1;
    return MPI_SUCCESS;
  }
  if (*request < 0 || *request >= (int)g_reqs.size() || !g_reqs[*request].used) return MPI_ERR_OTHER;
  Req& r = g_reqs[*request];
  if (!r.done && r.recv) try_recv(r, status);
  *flag = r.done ? 1 : 0;
  if (r.done) {
    r.used = false;
    *request = MPI_REQUEST_NULL;
  }
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request* request, MPI_Status* status) {
  int flag = 0, spins = 0;
  const double t0 = MPI_Wtime();
  for (;;) {
    const int rc = MPI_Test(request, &flag, status);
    if (rc != MPI_SUCCESS || flag) return rc;
    backoff(spins);
    if ((spins & 0xfff) == 0 && MPI_Wtime() - t0 > 300.0) die("MPI_Wait: no matching message for 300 s");
  }
}

int MPI_Waitall(int count, MPI_Request requests[], MPI_Status statuses[]) {
  // progress every pending receive in turn: completion order must not depend on the order of the array
  int spins = 0;
  const double t0 = MPI_Wtime();
  for (;;) {
    bool pending = false;
    for (int i = 0; i < count; ++i) {
      if (requests[i] == MPI_REQUEST_NULL) continue;
      int flag = 0;
      const int rc = MPI_Test(&requests[i], &flag, statuses ? &statuses[i] : nullptr);
      if (rc != MPI_SUCCESS) return rc;
      pending = pending || !flag;
    }
    if (!pending) return MPI_SUCCESS;
    backoff(spins);
    if ((spins & 0xfff) == 0 && MPI_Wtime() - t0 > 300.0) die("MPI_Waitall: no matching message for 300 s");
  }
}

int MPI_Recv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Status* status) {
  MPI_Request r;
  const int rc = MPI_Irecv(buf, count, datatype, source, tag, comm, &r);
  return rc != MPI_SUCCESS ? rc : MPI_Wait(&r, status);
}

int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag, void* recvbuf,
                 int recvcount, MPI_Datatype recvtype, int source, int recvtag, MPI_Comm comm, MPI_Status* status) {
  MPI_Request r;
  int rc = MPI_Irecv(recvbuf, recvcount, recvtype, source, recvtag, comm, &r);
  if (rc == MPI_SUCCESS) rc = MPI_Send(sendbuf, sendcount, sendtype, dest, sendtag, comm);
  return rc != MPI_SUCCESS ? rc : MPI_Wait(&r, status);
}

// ---- all-to-all family (on the point-to-point layer; tag space of its own) ----------------------------------------
static int alltoall_impl(Comm* c, MPI_Comm comm, const void* sendbuf, const int* scounts, const long* sbyte_off,
                         const MPI_Datatype* stypes, void* recvbuf, const int* rcounts, const long* rbyte_off,
                         const MPI_Datatype* rtypes) {
  const int n = (int)c->members.size();
  const int tag = 0x7A110000;
  std::vector<MPI_Request> reqs(n, MPI_REQUEST_NULL);
  for (int p = 0; p < n; ++p) {
    const int rc = MPI_Irecv((char*)recvbuf + rbyte_off[p], rcounts[p], rtypes[p], p, tag, comm, &reqs[p]);
    if (rc != MPI_SUCCESS) return rc;
  }
  for (int i = 0; i < n; ++i) {
    const int p = (c->rank + i) % n;
    const int rc = MPI_Send((const char*)sendbuf + sbyte_off[p], scounts[p], stypes[p], p, tag, comm);
    if (rc != MPI_SUCCESS) return rc;
  }
  const int rc = MPI_Waitall(n, reqs.data(), MPI_STATUSES_IGNORE);
  barrier(c);   // nobody reuses its outbox / send buffer before everybody has received
  return rc;
}

int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, MPI_Comm comm) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  const int n = (int)c->members.size();
  const long se = type_of(sendtype).extent, re = type_of(recvtype).extent;
  std::vector<int> sc(n, sendcount), rc(n, recvcount);
  std::vector<long> so(n), ro(n);
  std::vector<MPI_Datatype> st(n, sendtype), rt(n, recvtype);
  for (int p = 0; p < n; ++p) so[p] = (long)p * sendcount * se, ro[p] = (long)p * recvcount * re;
  return alltoall_impl(c, comm, sendbuf, sc.data(), so.data(), st.data(), recvbuf, rc.data(), ro.data(), rt.data());
}

int MPI_Alltoallv(const void* sendbuf, const int sendcounts[], const int sdispls[], MPI_Datatype sendtype, void* recvbuf,
                  const int recvcounts[], const int rdispls[], MPI_Datatype recvtype, MPI_Comm comm) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  const int n = (int)c->members.size();
  const long se = type_of(sendtype).extent, re = type_of(recvtype).extent;
  std::vector<long> so(n), ro(n);
  std::vector<MPI_Datatype> st(n, sendtype), rt(n, recvtype);
  for (int p = 0; p < n; ++p) so[p] = (long)sdispls[p] * se, ro[p] = (long)rdispls[p] * re;
  return alltoall_impl(c, comm, sendbuf, sendcounts, so.data(), st.data(), recvbuf, recvcounts, ro.data(), rt.data());
}

int MPI_Alltoallw(const void* sendbuf, const int sendcounts[], const int sdispls[], const MPI_Datatype sendtypes[],
                  void* recvbuf, const int recvcounts[], const int rdispls[], const MPI_Datatype recvtypes[], MPI_Comm comm) {
  Comm* c = get(comm);
  if (!c) return MPI_ERR_OTHER;
  const int n = (int)c->members.size();
  std::vector<long> so(n), ro(n);   // Alltoallw displacements are in BYTES
  for (int p = 0; p < n; ++p) so[p] = sdispls[p], ro[p] = rdispls[p];
  return alltoall_impl(c, comm, sendbuf, sendcounts, so.data(), sendtypes, recvbuf, recvcounts, ro.data(), recvtypes);
}

// ---- derived datatypes --------------------------------------------------------------------------------------------
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype) {
  *newtype = make_vector(1, count, 0, oldtype);
  return MPI_SUCCESS;
}
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype* newtype) {
  *newtype = make_vector(count, blocklength, (long)stride * type_of(oldtype).extent, oldtype);
  return MPI_SUCCESS;
}
int MPI_Type_hvector(int count, int blocklength, MPI_Aint stride, MPI_Datatype oldtype, MPI_Datatype* newtype) {
  *newtype = make_vector(count, blocklength, (long)stride, oldtype);
  return MPI_SUCCESS;
}
int MPI_Type_create_struct(int count, const int blocklengths[], const MPI_Aint displacements[], const MPI_Datatype types[],
                           MPI_Datatype* newtype) {
  Type n;
  n.used = true;
  long hi = 0;
  for (int b = 0; b < count; ++b) {
    const Type o = type_of(types[b]);
    for (int e = 0; e < blocklengths[b]; ++e)
      for (const Run& r : o.runs) add_run(n.runs, (long)displacements[b] + (long)e * o.extent + r.off, r.len);
    n.size += (size_t)blocklengths[b] * o.size;
    hi = std::max(hi, (long)displacements[b] + (long)blocklengths[b] * o.extent);
  }
  n.extent = (hi + 7) / 8 * 8;   // padded like a C struct of 8-byte-aligned members
  *newtype = register_type(n);
  return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype* datatype) {
  if (*datatype >= kFirstDerived && (size_t)(*datatype - kFirstDerived) < g_types.size()) {
    g_types[*datatype - kFirstDerived] = Type();
  }
  *datatype = MPI_DATATYPE_NULL;
  return MPI_SUCCESS;
}
int MPI_Type_size(MPI_Datatype datatype, int* size) {
  *size = (int)type_size(datatype);
  return MPI_SUCCESS;
}

int MPI_Op_create(MPI_User_function* user_fn, int, MPI_Op* op) {
  g_user_ops.push_back(user_fn);
  *op = kFirstUserOp + (int)g_user_ops.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Op_free(MPI_Op* op) {
  if (*op >= kFirstUserOp && (size_t)(*op - kFirstUserOp) < g_user_ops.size()) g_user_ops[*op - kFirstUserOp] = nullptr;
  *op = 0;
  return MPI_SUCCESS;
}

}  // extern "C"
