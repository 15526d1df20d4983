// plan.hpp -- backend-agnostic plan: geometry -> per-stage descriptors -> launches.
// Mirrors what accfft_plan_dft_3d_{r2c,c2c}_gpu builds (src/accfft_gpu.cpp:77-442, :704-931) and what
// accfft_execute_gpu / accfft_execute_c2c_gpu run (:444-623, :943-1136), but every pencil transpose
// is fused into the neighbouring FFT stage's stores (see stage.cuh) and the exchange buffers are the
// peers' own stage buffers (NVLink peer stores), not a host-staged MPI_Alltoall.
#pragma once
#include <stddef.h>
#include <string>
#include <vector>
#include "decomp.hpp"
#include "launch.hpp"

namespace accfft_b200 {

// Device services of ONE rank.  The product implements this with CUDA (backend_cuda.cu); tests/emul
// implements it with host memory and a fiber emulator (test infrastructure only).
struct Backend {
  virtual ~Backend() {}
  virtual bool supported(Kind kind, Prec prec, int fftN) = 0;   // a scheduled (fast) kernel exists for this complex length
  // the generic O(N^2) stage kernel can run a 3-D transform axis of this (real or complex) length (generic_dft.cu)
  virtual bool generic_ok(Prec, int /*N*/) { return false; }
  virtual const void* twiddle(Prec prec, int tableN) = 0;   // W_tableN[k], k < tableN
  virtual int launch(const Launch& l) = 0;                   // enqueue on the plan's stream; 0 = ok
  virtual int barrier() = 0;                                 // all ranks of the comm, stream ordered
  // Two-lane software pipelining of adjacent stages (lane 0 = the plan's stream, lane 1 = an auxiliary stream):
  //   pipeline_chunks()  how many chunks the backend wants (<= 1: run the stages back to back on lane 0)
  //   fork()             lane 1 waits for everything enqueued on lane 0 so far
  //   hand(from)         the OTHER lane waits for everything enqueued on lane `from` so far
  //   join()             lane 0 waits for lane 1
  //   barrier_on(lane)   the inter-rank barrier enqueued on that lane
  // Backends without streams (the CPU emulator) execute in program order, which satisfies every dependency.
  virtual int pipeline_chunks() { return 1; }
  virtual int fork() { return 0; }
  virtual int hand(int) { return 0; }
  virtual int join() { return 0; }
  virtual int barrier_on(int) { return barrier(); }
  // ---- copy-engine exchange (slab grids; Plan::forward_slab_ce / inverse_slab_ce) -------------------------------------------
  // The exchanging stage writes its peer-bound output into a LOCAL staging buffer, laid out like the sub-blocks of the peers'
  // buffers it is destined for; asynchronous copies (copy engines over NVLink: no SM involved) move chunk c while chunk c+1
  // computes.  Lane 0 is the plan's stream, lanes 1 .. copy_lanes() are copy lanes.
  //   ce_chunks()            chunks per exchanging stage; <= 0: the backend does not want the copy-engine path
  //   copy_lanes()           number of copy lanes (copies to different peers proceed concurrently on different lanes)
  //   staging()              this rank's staging buffer (at least one exchange buffer long)
  //   link(from, to)         lane `to` waits for everything enqueued on lane `from` so far
  //   copy2d(lane, ...)      enqueue a copy of `height` rows of `width` bytes (pitches in bytes) on a copy lane
  //   local_variant(l)       the kernel variant tuned for this stage when it stores locally (see Launch::variant)
  virtual int ce_chunks() { return 0; }
  virtual int copy_lanes() { return 1; }
  virtual void* staging() { return nullptr; }
  virtual int link(int, int) { return 0; }
  virtual int copy2d(int, void*, size_t, const void*, size_t, size_t, size_t) { return -1; }
  virtual int local_variant(const Launch&) { return 0; }
  // L2-resident hand-off between two local passes (Plan::forward_zy_l2 / inverse_yz_l2): how many i0 planes per chunk;
  // <= 0: one launch per pass over the whole array
  virtual int l2_planes(const struct Plan&) { return 0; }
  // mark [ptr, ptr + bytes) as the L2-persisting window of the stages launched from now on (bytes == 0: no window)
  virtual int l2_window(const void*, size_t) { return 0; }
  // The two passes that are local on a 1 x 1 / slab grid (forward Z then Y, inverse Y then Z) as ONE persistent kernel whose
  // intermediate stays in L2 (fused.cuh).  lz = the contiguous (Z) stage, ly = the strided (Y) stage; z_first = forward.
  // Returns -2 when the backend has no fused kernel for the pair (the caller launches the two stages one after the other).
  virtual int launch_fused(const Launch& /*lz*/, const Launch& /*ly*/, bool /*z_first*/) { return -2; }
};

struct Geometry {
  int n[3];
  int np[2], coords[2];
  bool r2c;
  int n2c;                 // N2/2+1 (r2c) or N2 (c2c)
  Blk i0, i1;              // input block:  [i0.size][i1.size][N2]
  Blk k2;                  // after exchange 1: [i0.size][N1][k2.size]
  Blk k1;                  // after exchange 2: [N0][k1.size][k2.size]  (= output)
  // single-axis REAL paths of r2c plans (src/accfft_common.cpp:150-190: osize_y / osize_x)
  Blk n2b;                 // y path: real [i0.size][N1][n2b.size], N2 split over the row group
  long long M;             // x path: the local (i1, N2) plane flattened, M = i1.size * N2 ...
  Blk mb;                  // ... and split over the column group: real [N0][mb.size]
  LocalSizes L;
};

Geometry make_geometry(const int n[3], const int np[2], const int coords[2], bool r2c, int sizeofReal);

inline int rank_of(const int np[2], int c0, int c1) { return c0 * np[1] + c1; }

struct Plan {
  Geometry g;
  Prec prec;
  bool inplace;
  Backend* be;
  int nprocs;
  // exchange buffers of every rank of the comm as seen from this rank (self = local pointers)
  std::vector<void*> peerA, peerB;
  // Double buffering (multi-rank plans): a second set of exchange buffers; every operation that exchanges starts with
  // flip().  An operation may then write its peers' buffers right away -- the peers last touched this set two exchanging
  // operations ago, and everybody has passed a barrier of the operation in between -- so barriers are needed only where
  // they order DATA (a stage reads what peers wrote), not buffer reuse: one barrier per transform on slab grids instead of
  // two.  Empty alt vectors = single buffering (every barrier stays).
  std::vector<void*> peerA_alt, peerB_alt;
  bool db() const { return nprocs > 1 && !peerA_alt.empty(); }
  void flip() {
    if (db()) {
      peerA.swap(peerA_alt);
      peerB.swap(peerB_alt);
    }
  }
  // Batched fields (the reference's howmany > 1 exchange path, src/transpose_gpu.cpp:4454-4726): a plan made for
  // `batch` fields owns `batch` sets of exchange buffers per rank; field f of rank q lives field_stride[q] * f bytes behind
  // field 0 in every buffer of that rank.  execute_*_many run every stage for all fields back to back and share the
  // barriers: one exchange synchronisation per stage for the whole batch instead of one per field.
  int batch = 1;
  std::vector<long long> field_stride;   // per rank, bytes (empty: batch == 1)
  void shift_fields(int f, int sign) {
    if (f == 0 || field_stride.empty()) return;
    for (std::vector<void*>* v : {&peerA, &peerB, &peerA_alt, &peerB_alt})
      for (size_t q = 0; q < v->size(); ++q)
        if ((*v)[q]) (*v)[q] = (char*)(*v)[q] + (long long)sign * f * field_stride[q];
  }
  struct FieldScope {   // stage descriptors built inside the scope address field f's buffers
    Plan& p;
    int f;
    FieldScope(Plan& p_, int f_) : p(p_), f(f_) { p.shift_fields(f, +1); }
    ~FieldScope() { p.shift_fields(f, -1); }
  };
  int execute_forward_many(int nf, const void* const* in, void* const* out, bool in_padded);
  int execute_inverse_many(int nf, const void* const* in, void* const* out, bool out_padded);
  bool row_exchanges() const { return g.np[1] > 1; }   // Z <-> Y hand-off crosses ranks
  bool col_exchanges() const { return g.np[0] > 1; }   // Y <-> X hand-off crosses ranks
  int bar_if(bool data);   // barrier if it orders data, or (single buffering) buffer reuse
  std::string error;
  // Buffer B uses a padded row pitch when there are peers: its rows hold k2.size (forward) or N2c (inverse)
  // complex elements but start every `pad_elems()` elements (128 bytes), so that the 128-byte row segments the
  // Y stages store over NVLink are line aligned.  Buffer A stays dense: its NVLink writers are the Z stage (long
  // contiguous runs) and the inverse X stage, whose flat (k1,k2) tiles are line aligned exactly when A is dense.
  // Single-rank plans keep everything dense.
  int pad_elems() const { return nprocs > 1 ? (prec == PREC_DOUBLE ? 8 : 16) : 1; }
  long long pitch(long long n) const {
    const long long p = pad_elems();
    return (n + p - 1) / p * p;
  }
  // bytes needed for each of the two exchange buffers of the rank at `coords`
  static long long buffer_bytes(const int n[3], const int np[2], const int coords[2], bool r2c, Prec prec, int nprocs);

  // the three forward and three inverse stages; `in` / `out` are the user's buffers
  Launch fwd_z(const void* in, bool in_padded) const;
  Launch fwd_y() const;
  Launch fwd_x(void* out) const;
  Launch inv_x(const void* in) const;
  Launch inv_y() const;
  Launch inv_z(void* out, bool out_padded) const;

  // ---- single-axis real transforms and the per-axis derivative (r2c plans; axis 0 = x, 1 = y, 2 = z) -------------
  // mode of ax_stage
  enum AxMode : int { AX_R2C = 0, AX_C2R = 1, AX_DERIV = 2 };
  bool ax_exchanges(int axis) const { return axis == 1 ? g.np[1] > 1 : axis == 0 ? g.np[0] > 1 : false; }
  Launch ax_copy_in(int axis, const void* in) const;                 // user real block -> peers' A in the axis layout
  Launch ax_stage(int axis, int mode, const void* in, void* out) const;   // x / y axes (strided kernels)
  Launch ax_copy_out(void* out) const;                               // own B (physical layout) -> user real block
  // reference: accfft_execute_{r2c,c2r}_{x,y,z}_gpu_t (src/accfft_gpu.cpp:1212-1604); layouts osize_y/yi, osize_x/xi
  int execute_r2c_axis(int axis, const void* in, void* out);
  int execute_c2r_axis(int axis, const void* in, void* out);
  // d/dx_axis of a real field, spectral, normalised; accumulate: out += (divergence)
  int execute_deriv_axis(int axis, const void* in, void* out, bool accumulate, void* spec);
  // per-axis schemes of the operators (src/operators_gpu.txx:172-278, :336-553)
  int execute_grad_axes(const void* in, void* const outs[3], int comps, void* spec);
  int execute_divergence_axes(const void* const ins[3], void* out, void* spec);

  bool check_supported();
  // do the fused per-axis kernels (single-axis real transforms, per-axis derivative) exist for this axis?
  bool ax_supported(int axis);
  // Load and configure every kernel the plan can launch (CUDA loads kernels lazily, and a first-use load may
  // synchronise the context -- which must not happen while another rank's barrier kernel is spinning).
  int prepare();
  // xyz: bit 2 = Z, bit 1 = Y, bit 0 = X (the std::bitset<3> of accfft_execute_gpu, src/accfft_gpu.cpp:477,499,527):
  // an axis whose bit is clear is not transformed, but its stage still moves the data into the next layout, exactly
  // as the reference always runs its transposes.  r2c / c2r need the Z bit (the real <-> complex step).
  int execute_forward(const void* in, void* out, bool in_padded, int xyz = 7);
  // slab grids (P1 == 1): the local Z stage and the exchanging Y stage software-pipelined over chunks of i0
  int forward_zy_pipelined(const void* in, bool in_padded, int nch);
  // slab grids: exchange through a local staging buffer + copy-engine transfers, overlapped with the neighbouring stages
  bool use_ce(int xyz) { return xyz == 7 && nprocs > 1 && g.np[1] == 1 && be->ce_chunks() > 0 && be->staging() != nullptr; }
  // the local Z / Y pair as one fused kernel when the backend has one for these lengths (rc -2: it does not)
  int fused_pair(const Launch& z, const Launch& y, bool z_first);
  int forward_zy_l2(const void* in, bool in_padded, int planes);
  int inverse_yz_l2(void* out, bool out_padded, int planes);
  int forward_slab_ce(const void* in, void* out, bool in_padded);
  int inverse_slab_ce(const void* in, void* out, bool out_padded);
  // The copy-engine inverse reads A after its last barrier.  Operations that write peers' A right at their start (the
  // direct peer-store stages) fence against that with one extra barrier: see a_fence().
  bool late_a_reader = false;
  int a_fence();
  int bar();
  int execute_inverse(const void* in, void* out, bool out_padded, int xyz = 7);
  // fused forward -> spectral symbol -> inverse (laplace / biharmonic; src/operators_gpu.txx:280-334, :555-609)
  int execute_spectral(const void* in, void* out, int op, void* spec, double opA = 0.0, double opB = 0.0);
  // grad (src/operators_gpu.txx:72-278): one forward into `spec` (normalised), then per selected component
  // an inverse whose X stage multiplies by i*k_axis on load.  comps: bit 0 = x, bit 1 = y, bit 2 = z.
  // pair_rule: zero the mode pair +-floor(N/2) like the per-axis scheme does (used when that scheme has no kernel for a length)
  int execute_grad(const void* in, void* const outs[3], int comps, void* spec, bool pair_rule = false);
  // divergence (src/operators_gpu.txx:336-553): three forwards whose X stage multiplies by i*k_axis and
  // accumulates into `spec`, then one inverse.
  int execute_divergence(const void* const ins[3], void* out, void* spec, bool pair_rule = false);
};

unsigned blk_magic(int blk);

}  // namespace accfft_b200
