// plan.cpp -- geometry and stage descriptors (host only; no CUDA calls here).
#include "plan.hpp"
#include <string.h>

namespace accfft_b200 {

unsigned blk_magic(int blk) {
  if (blk <= 1) return 0xFFFFFFFFu;   // never used when blk == 1 with nseg > 1? handled in seg_of callers
  return (unsigned)((0x100000000ULL + (unsigned long long)blk - 1) / (unsigned long long)blk);
}

Geometry make_geometry(const int n[3], const int np[2], const int coords[2], bool r2c, int sizeofReal) {
  Geometry g{};
  for (int i = 0; i < 3; ++i) g.n[i] = n[i];
  g.np[0] = np[0];
  g.np[1] = np[1];
  g.coords[0] = coords[0];
  g.coords[1] = coords[1];
  g.r2c = r2c;
  g.n2c = r2c ? n[2] / 2 + 1 : n[2];
  g.i0 = block_of(n[0], np[0], coords[0]);
  g.i1 = block_of(n[1], np[1], coords[1]);
  g.k2 = block_of(g.n2c, np[1], coords[1]);
  g.k1 = block_of(n[1], np[0], coords[0]);
  g.n2b = block_of(n[2], np[1], coords[1]);
  g.M = (long long)g.i1.size * n[2];
  g.mb = block_of((int)g.M, np[0], coords[0]);
  g.L = r2c ? local_size_r2c(n, np, coords, sizeofReal) : local_size_c2c(n, np, coords, sizeofReal);
  return g;
}

static inline void* offs(void* base, long long elems, size_t esize) { return (char*)base + elems * (long long)esize; }

static void set_segs_common(StageDesc& d, int nseg, int blk) {
  d.nseg = nseg;
  d.blk = blk;
  d.blkMagic = blk_magic(blk);
}

static StageDesc blank() {
  StageDesc d;
  memset(&d, 0, sizeof(d));
  d.op = OP_NONE;
  d.scale = 1.0;
  d.div = 1;
  return d;
}

Launch Plan::fwd_z(const void* in, bool in_padded) const {
  Launch l{};
  const size_t cs = prec == PREC_DOUBLE ? 16 : 8;
  StageDesc d = blank();
  const int P1 = g.np[1];
  d.in = in;
  d.nA = g.i0.size;
  d.nB = g.i1.size;
  if (g.r2c) {
    const long long row = in_padded ? 2LL * g.n2c : g.n[2];
    d.inStrideA = row * g.i1.size;
    d.inStrideB = row;
    d.inStrideI = 1;
    l.kind = K_CONTIG_R2C;
    l.fftN = g.n[2] / 2;
  } else {
    d.inStrideA = (long long)g.n[2] * g.i1.size;
    d.inStrideB = g.n[2];
    d.inStrideI = 1;
    l.kind = K_CONTIG_C2C;
    l.fftN = g.n[2];
  }
  d.tw = be->twiddle(prec, g.n[2]);
  set_segs_common(d, P1, ceil_blk(g.n2c, P1));
  for (int p = 0; p < P1; ++p) {
    const Blk kp = block_of(g.n2c, P1, p);
    const int r = rank_of(g.np, g.coords[0], p);
    const long long pp = kp.size;   // A is dense
    d.seg[p].base = offs(peerA[r], (long long)g.i1.start * pp, cs);
    d.seg[p].strideA = (long long)g.n[1] * pp;
    d.seg[p].strideB = pp;
    d.seg[p].strideI = 1;
  }
  l.prec = prec;
  l.inverse = false;
  l.desc = d;
  l.slot = 0;
  l.desc.inv = l.inverse ? 1 : 0;
  return l;
}

Launch Plan::fwd_y() const {
  Launch l{};
  const size_t cs = prec == PREC_DOUBLE ? 16 : 8;
  StageDesc d = blank();
  const int P0 = g.np[0];
  const int self = rank_of(g.np, g.coords[0], g.coords[1]);
  const long long pk = pitch(g.k2.size);
  d.in = peerA[self];
  d.nA = g.i0.size;
  d.nB = g.k2.size;
  d.inStrideA = (long long)g.n[1] * g.k2.size;
  d.inStrideB = 1;
  d.inStrideI = g.k2.size;
  d.tw = be->twiddle(prec, g.n[1]);
  set_segs_common(d, P0, ceil_blk(g.n[1], P0));
  for (int q = 0; q < P0; ++q) {
    const Blk k1q = block_of(g.n[1], P0, q);
    const int r = rank_of(g.np, q, g.coords[1]);
    d.seg[q].base = offs(peerB[r], (long long)g.i0.start * k1q.size * pk, cs);
    d.seg[q].strideA = (long long)k1q.size * pk;
    d.seg[q].strideB = 1;
    d.seg[q].strideI = pk;
  }
  l.kind = K_STRIDED_C2C;
  l.prec = prec;
  l.fftN = g.n[1];
  l.inverse = false;
  l.desc = d;
  l.slot = 1;
  l.desc.inv = l.inverse ? 1 : 0;
  return l;
}

Launch Plan::fwd_x(void* out) const {
  Launch l{};
  StageDesc d = blank();
  const int self = rank_of(g.np, g.coords[0], g.coords[1]);
  const long long plane = (long long)g.k1.size * g.k2.size;
  const long long pk = pitch(g.k2.size);
  d.in = peerB[self];
  d.tw = be->twiddle(prec, g.n[0]);
  set_segs_common(d, 1, g.n[0]);
  d.seg[0].base = out;
  d.n[0] = g.n[0];
  d.n[1] = g.n[1];
  d.n[2] = g.n[2];
  d.kind[0] = 0;
  d.off[0] = 0;
  d.off[1] = g.k1.start;
  d.off[2] = g.k2.start;
  if (pk == g.k2.size) {
    // dense B: the (k1, k2) plane is one flat, line-aligned column index
    d.nA = 1;
    d.nB = (int)plane;
    d.inStrideA = 0;
    d.inStrideB = 1;
    d.inStrideI = plane;
    d.seg[0].strideA = 0;
    d.seg[0].strideB = 1;
    d.seg[0].strideI = plane;
    // global frequency indices of an output element (k0 = transform index, k1/k2 from the flat column index)
    d.kind[1] = 3;
    d.kind[2] = 4;
    d.div = g.k2.size > 0 ? g.k2.size : 1;
  } else {
    // pitched B -> dense user layout: columns are (a = k1, b = k2)
    d.nA = g.k1.size;
    d.nB = g.k2.size;
    d.inStrideA = pk;
    d.inStrideB = 1;
    d.inStrideI = (long long)g.k1.size * pk;
    d.seg[0].strideA = g.k2.size;
    d.seg[0].strideB = 1;
    d.seg[0].strideI = plane;
    d.kind[1] = 1;
    d.kind[2] = 2;
  }
  l.kind = K_STRIDED_C2C;
  l.prec = prec;
  l.fftN = g.n[0];
  l.inverse = false;
  l.desc = d;
  l.slot = 2;
  l.desc.inv = l.inverse ? 1 : 0;
  return l;
}

Launch Plan::inv_x(const void* in) const {
  Launch l{};
  const size_t cs = prec == PREC_DOUBLE ? 16 : 8;
  StageDesc d = blank();
  const int P0 = g.np[0];
  const long long plane = (long long)g.k1.size * g.k2.size;
  d.in = in;
  d.tw = be->twiddle(prec, g.n[0]);
  set_segs_common(d, P0, ceil_blk(g.n[0], P0));
  d.n[0] = g.n[0];
  d.n[1] = g.n[1];
  d.n[2] = g.n[2];
  d.kind[0] = 0;
  d.off[0] = 0;
  d.off[1] = g.k1.start;
  d.off[2] = g.k2.start;
  // user spectrum and A are both dense: the (k1, k2) plane is one flat column index
  d.nA = 1;
  d.nB = (int)plane;
  d.inStrideA = 0;
  d.inStrideB = 1;
  d.inStrideI = plane;
  d.kind[1] = 3;
  d.kind[2] = 4;
  d.div = g.k2.size > 0 ? g.k2.size : 1;
  for (int q = 0; q < P0; ++q) {
    const int r = rank_of(g.np, q, g.coords[1]);
    d.seg[q].base = offs(peerA[r], (long long)g.k1.start * g.k2.size, cs);
    d.seg[q].strideA = 0;
    d.seg[q].strideB = 1;
    d.seg[q].strideI = (long long)g.n[1] * g.k2.size;
  }
  l.kind = K_STRIDED_C2C;
  l.prec = prec;
  l.fftN = g.n[0];
  l.inverse = true;
  l.desc = d;
  l.slot = 3;
  l.desc.inv = l.inverse ? 1 : 0;
  return l;
}

Launch Plan::inv_y() const {
  Launch l{};
  const size_t cs = prec == PREC_DOUBLE ? 16 : 8;
  StageDesc d = blank();
  const int P1 = g.np[1];
  const int self = rank_of(g.np, g.coords[0], g.coords[1]);
  const long long pz = pitch(g.n2c);
  d.in = peerA[self];
  d.nA = g.i0.size;
  d.nB = g.k2.size;
  d.inStrideA = (long long)g.n[1] * g.k2.size;
  d.inStrideB = 1;
  d.inStrideI = g.k2.size;
  d.colShift = g.k2.size > 0 ? g.k2.start % pad_elems() : 0;   // destination rows are entered at column k2.start
  d.tw = be->twiddle(prec, g.n[1]);
  set_segs_common(d, P1, ceil_blk(g.n[1], P1));
  for (int p = 0; p < P1; ++p) {
    const Blk i1p = block_of(g.n[1], P1, p);
    const int r = rank_of(g.np, g.coords[0], p);
    d.seg[p].base = offs(peerB[r], g.k2.start, cs);
    d.seg[p].strideA = (long long)i1p.size * pz;
    d.seg[p].strideB = 1;
    d.seg[p].strideI = pz;
  }
  l.kind = K_STRIDED_C2C;
  l.prec = prec;
  l.fftN = g.n[1];
  l.inverse = true;
  l.desc = d;
  l.slot = 4;
  l.desc.inv = l.inverse ? 1 : 0;
  return l;
}

Launch Plan::inv_z(void* out, bool out_padded) const {
  Launch l{};
  StageDesc d = blank();
  const int self = rank_of(g.np, g.coords[0], g.coords[1]);
  const long long pz = pitch(g.n2c);
  d.in = peerB[self];
  d.nA = g.i0.size;
  d.nB = g.i1.size;
  d.inStrideA = pz * g.i1.size;
  d.inStrideB = pz;
  d.inStrideI = 1;
  d.tw = be->twiddle(prec, g.n[2]);
  set_segs_common(d, 1, g.n[2] + 2);
  d.seg[0].base = out;
  if (g.r2c) {
    const long long row = out_padded ? 2LL * g.n2c : g.n[2];
    d.seg[0].strideA = row * g.i1.size;
    d.seg[0].strideB = row;
    d.seg[0].strideI = 1;
    l.kind = K_CONTIG_C2R;
    l.fftN = g.n[2] / 2;
  } else {
    d.seg[0].strideA = (long long)g.n[2] * g.i1.size;
    d.seg[0].strideB = g.n[2];
    d.seg[0].strideI = 1;
    l.kind = K_CONTIG_C2C;
    l.fftN = g.n[2];
  }
  l.prec = prec;
  l.inverse = true;
  l.desc = d;
  l.slot = 5;
  l.desc.inv = l.inverse ? 1 : 0;
  return l;
}

bool Plan::check_supported() {
  if (g.np[0] > kMaxSeg || g.np[1] > kMaxSeg) {
    error = "process grid dimension exceeds kMaxSeg";
    return false;
  }
  // every axis: a scheduled kernel, else the generic O(N^2) stage kernel (any length it can stage in shared memory; also
  // serves r2c / c2r plans with an odd last dimension, which the scheduled kernels' N2/2-point trick cannot)
  if (g.r2c) {
    const bool fast = g.n[2] % 2 == 0 && be->supported(K_CONTIG_R2C, prec, g.n[2] / 2) && be->supported(K_CONTIG_C2R, prec, g.n[2] / 2);
    if (!fast && !be->generic_ok(prec, g.n[2])) {
      error = "no kernel for N2 = " + std::to_string(g.n[2]);
      return false;
    }
  } else if (!be->supported(K_CONTIG_C2C, prec, g.n[2]) && !be->generic_ok(prec, g.n[2])) {
    error = "no kernel for N2 = " + std::to_string(g.n[2]);
    return false;
  }
  if (!be->supported(K_STRIDED_C2C, prec, g.n[1]) && !be->generic_ok(prec, g.n[1])) {
    error = "no kernel for N1 = " + std::to_string(g.n[1]);
    return false;
  }
  if (!be->supported(K_STRIDED_C2C, prec, g.n[0]) && !be->generic_ok(prec, g.n[0])) {
    error = "no kernel for N0 = " + std::to_string(g.n[0]);
    return false;
  }
  return true;
}

bool Plan::ax_supported(int axis) {
  if (!g.r2c) return false;
  if (axis == 2) return g.n[2] % 2 == 0 && be->supported(K_CONTIG_R2C, prec, g.n[2] / 2) && be->supported(K_CONTIG_C2R, prec, g.n[2] / 2);
  return be->supported(K_STRIDED_R2C, prec, g.n[axis]);
}

long long Plan::buffer_bytes(const int n[3], const int np[2], const int coords[2], bool r2c, Prec prec, int nprocs) {
  Plan tmp;
  tmp.prec = prec;
  tmp.nprocs = nprocs;
  const long long cs = prec == PREC_DOUBLE ? 16 : 8;
  const int n2c = r2c ? n[2] / 2 + 1 : n[2];
  const Blk i0 = block_of(n[0], np[0], coords[0]), i1 = block_of(n[1], np[1], coords[1]);
  const Blk k2 = block_of(n2c, np[1], coords[1]), k1 = block_of(n[1], np[0], coords[0]);
  const long long a = (long long)i0.size * n[1] * k2.size;
  const long long b1 = (long long)n[0] * k1.size * tmp.pitch(k2.size);
  const long long b2 = (long long)i0.size * i1.size * tmp.pitch(n2c);
  long long m = std::max(a, std::max(b1, b2)) * cs;
  if (r2c) {
    // single-axis real paths: [i0][N1][n2b] and [N0][mb] reals in A (the physical layout [i0][i1][N2] in B is covered by b2)
    const long long rs = cs / 2;
    const Blk n2b = block_of(n[2], np[1], coords[1]);
    const Blk mb = block_of((int)((long long)i1.size * n[2]), np[0], coords[0]);
    m = std::max(m, (long long)i0.size * n[1] * n2b.size * rs);
    m = std::max(m, (long long)n[0] * mb.size * rs);
  }
  return std::max<long long>(m, 16);
}

static void no_transform(Launch& l, int n_axis);
static bool empty(const Launch& l) {
  return l.desc.nA <= 0 || l.desc.nB <= 0 || (l.kind == K_COPY_REAL && l.copyN <= 0);
}

int Plan::prepare() {
  Launch ls[6] = {fwd_z(nullptr, false), fwd_y(), fwd_x(nullptr), inv_x(nullptr), inv_y(), inv_z(nullptr, false)};
  for (Launch& l : ls) {
    l.dry = true;
    if (int rc = be->launch(l)) return rc;
  }
  if (g.np[1] == 1) {   // fused local pair (both directions)
    ls[0].dry = ls[1].dry = ls[4].dry = ls[5].dry = true;
    int rc = be->launch_fused(ls[0], ls[1], true);
    if (rc && rc != -2) return rc;
    rc = be->launch_fused(ls[5], ls[4], false);
    if (rc && rc != -2) return rc;
  }
  Launch c = fwd_y();   // the no-transform stage of partial-axis executes
  no_transform(c, g.n[1]);
  c.dry = true;
  if (int rc = be->launch(c)) return rc;
  if (g.r2c) {
    for (int axis = 0; axis < 2; ++axis)
      for (int mode = 0; mode < 3 && ax_supported(axis); ++mode) {
        Launch a = ax_stage(axis, mode, nullptr, nullptr);
        a.dry = true;
        if (int rc = be->launch(a)) return rc;
      }
    Launch ci = ax_copy_out(nullptr);
    ci.dry = true;
    if (int rc = be->launch(ci)) return rc;
    if (be->supported(K_CONTIG_DERIV, prec, g.n[2])) {
      Launch zd = fwd_z(nullptr, false);
      zd.desc.tw = be->twiddle(prec, g.n[2]);
      zd.kind = K_CONTIG_DERIV;
      zd.fftN = g.n[2];
      zd.dry = true;
      if (int rc = be->launch(zd)) return rc;
    }
  }
  return 0;
}

// a stage whose axis is not selected only re-lays the data out (pack-on-store without the transform)
static void no_transform(Launch& l, int n_axis) {
  l.kind = K_COPY;
  l.fftN = n_axis;
}

// restrict a stage to the outer indices [a0, a0 + na): input and every destination advance by a0 outer strides
static Launch restrict_a(Launch l, int a0, int na, size_t in_elem, size_t out_elem) {
  StageDesc& d = l.desc;
  d.in = (const char*)d.in + (long long)a0 * d.inStrideA * (long long)in_elem;
  for (int s = 0; s < d.nseg; ++s) d.seg[s].base = (char*)d.seg[s].base + (long long)a0 * d.seg[s].strideA * (long long)out_elem;
  d.nA = na;
  return l;
}

// Slab grids (P1 == 1): Z writes only this rank's A and Y is the stage whose stores cross NVLink, so the two can be
// software-pipelined over chunks of the outer index i0 with no extra synchronisation between ranks: chunk c of Y (lane 1,
// NVLink-bound) runs while chunk c+1 of Z (lane 0, HBM-bound) computes.  The barrier that opens the operation for writes
// into the peers' B ("B is never written before the first barrier of an operation") moves to lane 1, where
// it overlaps the first Z chunk instead of waiting in front of it.
int Plan::forward_zy_pipelined(const void* in, bool in_padded, int nch) {
  int rc;
  const size_t cs = prec == PREC_DOUBLE ? 16 : 8;
  const size_t in_elem = g.r2c ? cs / 2 : cs;
  const Launch z = fwd_z(in, in_padded), y = fwd_y();
  const int n = g.i0.size;
  if (nch > n) nch = n > 0 ? n : 1;
  if ((rc = be->fork())) return rc;
  if ((rc = be->barrier_on(1))) return rc;
  for (int c = 0; c < nch; ++c) {
    const int a0 = (int)((long long)n * c / nch), a1 = (int)((long long)n * (c + 1) / nch);
    if (a1 <= a0) continue;
    Launch zc = restrict_a(z, a0, a1 - a0, in_elem, cs);
    zc.lane = 0;
    if (!empty(zc) && (rc = be->launch(zc))) return rc;
    if ((rc = be->hand(0))) return rc;
    Launch yc = restrict_a(y, a0, a1 - a0, cs, cs);
    yc.lane = 1;
    if (!empty(yc) && (rc = be->launch(yc))) return rc;
  }
  return be->join();
}

// restrict a stage to the columns [b0, b0 + nb) of its (flat) inner index: unit column stride on both sides
static Launch restrict_b(Launch l, int b0, int nb, size_t in_elem, size_t out_elem) {
  StageDesc& d = l.desc;
  d.in = (const char*)d.in + (long long)b0 * d.inStrideB * (long long)in_elem;
  for (int s = 0; s < d.nseg; ++s) d.seg[s].base = (char*)d.seg[s].base + (long long)b0 * d.seg[s].strideB * (long long)out_elem;
  d.nB = nb;
  for (int ax = 0; ax < 3; ++ax)
    if (d.kind[ax] == 2) d.off[ax] += b0;   // frequency index taken from the column index
  return l;
}

// ================================================================================================================
// L2-resident hand-off between the two LOCAL passes of a transform (single rank; slab grids for the pair that stays local).
// Forward: Z then Y both walk the outer index i0, and Y(i0) needs nothing but Z(i0)'s plane.  Run them chunk by chunk over
// i0 with the intermediate planes in ONE small buffer (the head of A) that every chunk reuses: at `planes` x plane bytes
// well under the 126 MB L2 the Z stage's stores are still resident when the Y stage reads them, and because the next chunk
// overwrites the same lines they are never written back either -- the intermediate array's HBM write AND read disappear
// (2 of the 6 array passes' worth of traffic per transform: R + 5C -> R + 3C).  Inverse: Y then Z, same thing through the
// head of B.  The reference's analogue is the hand-off between cufftExecD2Z and the per-slab Y loop
// (src/accfft_gpu.cpp:477-516), which goes through global memory for the whole array.
// ================================================================================================================
static Launch chunk_a(Launch l, int a0, int na, size_t in_elem, size_t out_elem, bool move_in, bool move_out) {
  StageDesc& d = l.desc;
  if (move_in) d.in = (const char*)d.in + (long long)a0 * d.inStrideA * (long long)in_elem;
  if (move_out)
    for (int s = 0; s < d.nseg; ++s) d.seg[s].base = (char*)d.seg[s].base + (long long)a0 * d.seg[s].strideA * (long long)out_elem;
  d.nA = na;
  return l;
}

int Plan::forward_zy_l2(const void* in, bool in_padded, int planes) {
  int rc;
  const size_t cs = prec == PREC_DOUBLE ? 16 : 8;
  const size_t in_elem = g.r2c ? cs / 2 : cs;
  const Launch z = fwd_z(in, in_padded), y = fwd_y();
  const int n = g.i0.size;
  if ((rc = be->l2_window(y.desc.in, (size_t)planes * (size_t)y.desc.inStrideA * cs))) return rc;
  for (int a0 = 0; a0 < n; a0 += planes) {
    const int na = std::min(planes, n - a0);
    Launch zc = chunk_a(z, a0, na, in_elem, cs, true, false);    // input advances, output = head of A
    if (!empty(zc) && (rc = be->launch(zc))) return rc;
    Launch yc = chunk_a(y, a0, na, cs, cs, false, true);         // input = head of A, output advances
    if (!empty(yc) && (rc = be->launch(yc))) return rc;
  }
  return be->l2_window(nullptr, 0);
}

int Plan::inverse_yz_l2(void* out, bool out_padded, int planes) {
  int rc;
  const size_t cs = prec == PREC_DOUBLE ? 16 : 8;
  const size_t out_elem = g.r2c ? cs / 2 : cs;
  const Launch y = inv_y(), z = inv_z(out, out_padded);
  const int n = g.i0.size;
  if ((rc = be->l2_window(z.desc.in, (size_t)planes * (size_t)z.desc.inStrideA * cs))) return rc;
  for (int a0 = 0; a0 < n; a0 += planes) {
    const int na = std::min(planes, n - a0);
    Launch yc = chunk_a(y, a0, na, cs, cs, true, false);         // input advances through A, output = head of B
    if (!empty(yc) && (rc = be->launch(yc))) return rc;
    Launch zc = chunk_a(z, a0, na, cs, out_elem, false, true);   // input = head of B, output advances
    if (!empty(zc) && (rc = be->launch(zc))) return rc;
  }
  return be->l2_window(nullptr, 0);
}

int Plan::bar_if(bool data) {
  if (nprocs <= 1) return 0;
  if (db() && !data) return 0;
  return bar();
}

int Plan::a_fence() {
  if (db()) return 0;   // double buffering: the peers' late reads concern the other buffer set
  if (!late_a_reader) return 0;
  return nprocs > 1 ? bar() : 0;
}

// inter-rank barrier of the direct (peer-store) paths: any barrier also settles a pending late read of A
int Plan::bar() {
  late_a_reader = false;
  return be->barrier();
}

// ================================================================================================================
// Copy-engine exchange on slab grids (P1 == 1).  Reference analogue: the per-peer pipelined variants of its transposes,
// fast_transpose_cuda_v1_2_h (src/transpose_gpu.cpp:2643-2673: per-peer cudaMemcpyAsync around MPI_Isend / Irecv) and
// par::Mpi_Alltoallv_dense_gpu (include/parUtils.txx:165-384) -- there over PCIe and host memory, here device to device over
// NVLink, with the pack fused into the producing FFT stage and no unpack at all.
//
// forward:  lane 1: [barrier F1]            copies(c=0) ...           copies(c=NCH-1)
//           lane 0: Z(0) Y(0) | Z(1) Y(1) | ... | Z(NCH-1) Y(NCH-1) |        join, barrier F2, X
//   Y(c) transforms the i0 rows of chunk c and stores the k1 block of peer q into staging_q (layout [i0][k1_q][pitch], exactly
//   the sub-block of B_q this rank owns), its own k1 block straight into its own B.  One contiguous copy per (chunk, peer).
// inverse:  lane 1: [barrier I0]      copies(c=0) ...        copies(c=NCH-1)
//           lane 0: X(0) | X(1) | ... | X(NCH-1) |                  join, barrier I1, Y, Z
//   X(c) transforms the columns of chunk c (flat (k1,k2) plane) and stores the i0 block of peer q into staging_q (layout
//   [i0_q][plane]); one 2-D copy per (chunk, peer): i0_q rows of chunk-width bytes into A_q (row pitch N1 * k2).
// Barriers: F1 / I0 ("every peer has finished its previous operation, so its B / A may be overwritten") run on the copy lane
// and hide behind the first chunk's compute; F2 / I1 ("all copies have landed") are the only exposed ones.  On a slab grid
// nothing else crosses ranks (Z and the inverse Y / Z are local), so a transform has 2 barriers instead of the direct path's 2
// exposed ones, and the NVLink transfer overlaps the HBM-bound stages instead of running inside them.
// ================================================================================================================
int Plan::forward_slab_ce(const void* in, void* out, bool in_padded) {
  int rc;
  const size_t cs = prec == PREC_DOUBLE ? 16 : 8;
  const size_t in_elem = g.r2c ? cs / 2 : cs;
  const int P0 = g.np[0];
  const int self = rank_of(g.np, g.coords[0], g.coords[1]);
  const long long pk = pitch(g.k2.size);
  const int n = g.i0.size;
  int nch = be->ce_chunks();
  if (nch > n) nch = n > 0 ? n : 1;
  const int lanes = be->copy_lanes();
  const int pieces = P0 - 1 >= lanes ? 1 : (lanes + P0 - 2) / (P0 - 1);
  const Launch z = fwd_z(in, in_padded);
  Launch y = fwd_y();
  y.variant = be->local_variant(y);
  // peers' k1 blocks go to the staging buffer, packed peer after peer
  std::vector<long long> soff(P0, 0);
  long long pos = 0;
  for (int q = 0; q < P0; ++q) {
    const int r = rank_of(g.np, q, g.coords[1]);
    if (r == self) continue;
    const Blk k1q = block_of(g.n[1], P0, q);
    soff[q] = pos;
    y.desc.seg[q].base = offs(be->staging(), pos, cs);   // strides unchanged: [i0][k1_q][pk]
    pos += (long long)n * k1q.size * pk;
  }
  if ((rc = be->link(0, 1))) return rc;
  if ((rc = be->barrier_on(1))) return rc;                       // F1
  for (int l = 2; l <= lanes; ++l)
    if ((rc = be->link(1, l))) return rc;
  for (int c = 0; c < nch; ++c) {
    const int a0 = (int)((long long)n * c / nch), a1 = (int)((long long)n * (c + 1) / nch);
    if (a1 <= a0) continue;
    Launch zc = restrict_a(z, a0, a1 - a0, in_elem, cs);
    if (!empty(zc) && (rc = be->launch(zc))) return rc;
    Launch yc = restrict_a(y, a0, a1 - a0, cs, cs);
    if (!empty(yc) && (rc = be->launch(yc))) return rc;
    for (int l = 1; l <= lanes; ++l)
      if ((rc = be->link(0, l))) return rc;
    for (int i = 1; i < P0; ++i) {
      const int q = (g.coords[0] + i) % P0;                      // every rank starts with a different peer
      const int r = rank_of(g.np, q, g.coords[1]);
      const Blk k1q = block_of(g.n[1], P0, q);
      const long long row = (long long)k1q.size * pk;            // elements per i0 row of the block
      if (row <= 0) continue;
      // one transfer per (chunk, peer), cut into `pieces` row ranges on different lanes when there are fewer peers than
      // lanes: a single copy engine does not fill the link (measured: ~500 GB/s per transfer on B200 NVLink 5)
      for (int pc = 0; pc < pieces; ++pc) {
        const int r0 = a0 + (int)((long long)(a1 - a0) * pc / pieces), r1 = a0 + (int)((long long)(a1 - a0) * (pc + 1) / pieces);
        if (r1 <= r0) continue;
        void* dst = offs(peerB[r], ((long long)g.i0.start + r0) * row, cs);
        const void* src = offs(be->staging(), soff[q] + (long long)r0 * row, cs);
        const size_t bytes = (size_t)(r1 - r0) * (size_t)row * cs;
        if ((rc = be->copy2d(1 + ((i - 1) * pieces + pc) % lanes, dst, bytes, src, bytes, bytes, 1))) return rc;
      }
    }
  }
  for (int l = 1; l <= lanes; ++l)
    if ((rc = be->link(l, 0))) return rc;
  if ((rc = be->barrier())) return rc;                           // F2
  late_a_reader = false;
  Launch x = fwd_x(out);
  if (!empty(x) && (rc = be->launch(x))) return rc;
  return 0;
}

int Plan::inverse_slab_ce(const void* in, void* out, bool out_padded) {
  int rc;
  const size_t cs = prec == PREC_DOUBLE ? 16 : 8;
  const int P0 = g.np[0];
  const int self = rank_of(g.np, g.coords[0], g.coords[1]);
  const long long plane = (long long)g.k1.size * g.k2.size;
  const long long arow = (long long)g.n[1] * g.k2.size;          // elements per i0 row of A
  int nch = be->ce_chunks();
  const int lanes = be->copy_lanes();
  const int pieces = P0 - 1 >= lanes ? 1 : (lanes + P0 - 2) / (P0 - 1);
  Launch x = inv_x(in);
  x.variant = be->local_variant(x);
  std::vector<long long> soff(P0, 0);
  long long pos = 0;
  for (int q = 0; q < P0; ++q) {
    const int r = rank_of(g.np, q, g.coords[1]);
    if (r == self) continue;
    const Blk i0q = block_of(g.n[0], P0, q);
    soff[q] = pos;
    x.desc.seg[q].base = offs(be->staging(), pos, cs);
    x.desc.seg[q].strideI = plane;                               // staging rows are dense: [i0_q][plane]
    pos += (long long)i0q.size * plane;
  }
  // chunk boundaries on multiples of 64 columns: tiles stay line aligned
  const long long nblk = (plane + 63) / 64;
  if (nch > nblk) nch = nblk > 0 ? (int)nblk : 1;
  if ((rc = be->link(0, 1))) return rc;
  if ((rc = be->barrier_on(1))) return rc;                       // I0
  for (int l = 2; l <= lanes; ++l)
    if ((rc = be->link(1, l))) return rc;
  for (int c = 0; c < nch; ++c) {
    const long long b0 = std::min(plane, nblk * c / nch * 64), b1 = std::min(plane, nblk * (c + 1) / nch * 64);
    if (b1 <= b0) continue;
    Launch xc = restrict_b(x, (int)b0, (int)(b1 - b0), cs, cs);
    if (!empty(xc) && (rc = be->launch(xc))) return rc;
    for (int l = 1; l <= lanes; ++l)
      if ((rc = be->link(0, l))) return rc;
    for (int i = 1; i < P0; ++i) {
      const int q = (g.coords[0] + i) % P0;
      const int r = rank_of(g.np, q, g.coords[1]);
      const Blk i0q = block_of(g.n[0], P0, q);
      if (i0q.size <= 0) continue;
      for (int pc = 0; pc < pieces; ++pc) {   // row ranges of the 2-D transfer on different lanes (see forward_slab_ce)
        const int r0 = (int)((long long)i0q.size * pc / pieces), r1 = (int)((long long)i0q.size * (pc + 1) / pieces);
        if (r1 <= r0) continue;
        void* dst = offs(peerA[r], (long long)r0 * arow + (long long)g.k1.start * g.k2.size + b0, cs);
        const void* src = offs(be->staging(), soff[q] + (long long)r0 * plane + b0, cs);
        if ((rc = be->copy2d(1 + ((i - 1) * pieces + pc) % lanes, dst, (size_t)arow * cs, src, (size_t)plane * cs,
                             (size_t)(b1 - b0) * cs, (size_t)(r1 - r0))))
          return rc;
      }
    }
  }
  for (int l = 1; l <= lanes; ++l)
    if ((rc = be->link(l, 0))) return rc;
  if ((rc = be->barrier())) return rc;                           // I1
  Launch y = inv_y();
  if (!empty(y) && (rc = be->launch(y))) return rc;
  Launch z = inv_z(out, out_padded);
  if (!empty(z) && (rc = be->launch(z))) return rc;
  late_a_reader = true;   // Y read A after the last barrier: see a_fence()
  return 0;
}

// Z and Y are both local to the rank when the row group has one member: try the fused kernel, else one after the other
int Plan::fused_pair(const Launch& z, const Launch& y, bool z_first) {
  if (g.np[1] != 1 || empty(z) || empty(y)) return -2;
  return be->launch_fused(z, y, z_first);
}

int Plan::execute_forward(const void* in, void* out, bool in_padded, int xyz) {
  int rc;
  flip();
  if (use_ce(xyz)) return forward_slab_ce(in, out, in_padded);
  if (xyz == 7 && nprocs == 1 && be->l2_planes(*this) > 0) {
    if ((rc = forward_zy_l2(in, in_padded, be->l2_planes(*this)))) return rc;
    Launch x = fwd_x(out);
    if (!empty(x) && (rc = be->launch(x))) return rc;
    return 0;
  }
  if (xyz == 7 && nprocs > 1 && g.np[1] == 1 && be->pipeline_chunks() > 1) {
    if ((rc = forward_zy_pipelined(in, in_padded, be->pipeline_chunks()))) return rc;
    if ((rc = bar())) return rc;
    Launch x = fwd_x(out);
    if (!empty(x) && (rc = be->launch(x))) return rc;
    return 0;
  }
  Launch z = fwd_z(in, in_padded);
  Launch y = fwd_y();
  rc = xyz == 7 ? fused_pair(z, y, true) : -2;
  if (rc == -2) {
    if (!(xyz & 4)) no_transform(z, g.n[2]);
    if (!empty(z) && (rc = be->launch(z))) return rc;
    if ((rc = bar_if(row_exchanges()))) return rc;
    if (!(xyz & 2)) no_transform(y, g.n[1]);
    if (!empty(y) && (rc = be->launch(y))) return rc;
  } else if (rc) {
    return rc;
  }
  if ((rc = bar_if(col_exchanges()))) return rc;
  Launch x = fwd_x(out);
  if (!(xyz & 1)) no_transform(x, g.n[0]);
  if (!empty(x) && (rc = be->launch(x))) return rc;
  return 0;
}

int Plan::execute_inverse(const void* in, void* out, bool out_padded, int xyz) {
  int rc;
  flip();
  if (use_ce(xyz)) return inverse_slab_ce(in, out, out_padded);
  if (xyz == 7 && nprocs == 1 && be->l2_planes(*this) > 0) {
    Launch x = inv_x(in);
    if (!empty(x) && (rc = be->launch(x))) return rc;
    return inverse_yz_l2(out, out_padded, be->l2_planes(*this));
  }
  if ((rc = a_fence())) return rc;
  Launch x = inv_x(in);
  if (!(xyz & 1)) no_transform(x, g.n[0]);
  if (!empty(x) && (rc = be->launch(x))) return rc;
  if ((rc = bar_if(col_exchanges()))) return rc;
  Launch y = inv_y();
  Launch z = inv_z(out, out_padded);
  rc = xyz == 7 ? fused_pair(z, y, false) : -2;
  if (rc == -2) {
    if (!(xyz & 2)) no_transform(y, g.n[1]);
    if (!empty(y) && (rc = be->launch(y))) return rc;
    if ((rc = bar_if(row_exchanges()))) return rc;
    if (!(xyz & 4)) no_transform(z, g.n[2]);
    if (!empty(z) && (rc = be->launch(z))) return rc;
  } else if (rc) {
    return rc;
  }
  return 0;
}

// ---- batched fields: every stage for all fields, then ONE barrier ----------------------------------------------------
int Plan::execute_forward_many(int nf, const void* const* in, void* const* out, bool in_padded) {
  if (nf < 1 || nf > batch) {
    error = "execute_*_many: the plan was made for fewer fields (accfftb_plan_create_many)";
    return -3;
  }
  if (nf == 1) return execute_forward(in[0], out[0], in_padded);
  int rc;
  flip();
  for (int f = 0; f < nf; ++f) {
    FieldScope fs(*this, f);
    Launch z = fwd_z(in[f], in_padded);
    if (!empty(z) && (rc = be->launch(z))) return rc;
  }
  if ((rc = bar_if(row_exchanges()))) return rc;
  for (int f = 0; f < nf; ++f) {
    FieldScope fs(*this, f);
    Launch y = fwd_y();
    if (!empty(y) && (rc = be->launch(y))) return rc;
  }
  if ((rc = bar_if(col_exchanges()))) return rc;
  for (int f = 0; f < nf; ++f) {
    FieldScope fs(*this, f);
    Launch x = fwd_x(out[f]);
    if (!empty(x) && (rc = be->launch(x))) return rc;
  }
  return 0;
}

int Plan::execute_inverse_many(int nf, const void* const* in, void* const* out, bool out_padded) {
  if (nf < 1 || nf > batch) {
    error = "execute_*_many: the plan was made for fewer fields (accfftb_plan_create_many)";
    return -3;
  }
  if (nf == 1) return execute_inverse(in[0], out[0], out_padded);
  int rc;
  flip();
  if ((rc = a_fence())) return rc;
  for (int f = 0; f < nf; ++f) {
    FieldScope fs(*this, f);
    Launch x = inv_x(in[f]);
    if (!empty(x) && (rc = be->launch(x))) return rc;
  }
  if ((rc = bar_if(col_exchanges()))) return rc;
  for (int f = 0; f < nf; ++f) {
    FieldScope fs(*this, f);
    Launch y = inv_y();
    if (!empty(y) && (rc = be->launch(y))) return rc;
  }
  if ((rc = bar_if(row_exchanges()))) return rc;
  for (int f = 0; f < nf; ++f) {
    FieldScope fs(*this, f);
    Launch z = inv_z(out[f], out_padded);
    if (!empty(z) && (rc = be->launch(z))) return rc;
  }
  return 0;
}

int Plan::execute_spectral(const void* in, void* out, int op, void* spec, double opA, double opB) {
  int rc;
  flip();
  Launch z = fwd_z(in, false);
  Launch y = fwd_y();
  if ((rc = fused_pair(z, y, true)) == -2) {
    if (!empty(z) && (rc = be->launch(z))) return rc;
    if ((rc = bar_if(row_exchanges()))) return rc;
    if (!empty(y) && (rc = be->launch(y))) return rc;
  } else if (rc) {
    return rc;
  }
  if ((rc = bar_if(col_exchanges()))) return rc;
  // X forward with the symbol multiply and the 1/N normalisation fused into its stores, into `spec`
  // (user spectrum layout), which the inverse X stage reads back
  Launch x = fwd_x(spec);
  x.desc.op = op;
  x.desc.opA = opA;
  x.desc.opB = opB;
  x.desc.scale = 1.0 / ((double)g.n[0] * (double)g.n[1] * (double)g.n[2]);
  if (!empty(x) && (rc = be->launch(x))) return rc;
  flip();
  Launch xi = inv_x(spec);
  if (!empty(xi) && (rc = be->launch(xi))) return rc;
  if ((rc = bar_if(col_exchanges()))) return rc;
  Launch yi = inv_y();
  Launch zi = inv_z(out, false);
  if ((rc = fused_pair(zi, yi, false)) == -2) {
    if (!empty(yi) && (rc = be->launch(yi))) return rc;
    if ((rc = bar_if(row_exchanges()))) return rc;
    if (!empty(zi) && (rc = be->launch(zi))) return rc;
  } else if (rc) {
    return rc;
  }
  return 0;
}

int Plan::execute_grad(const void* in, void* const outs[3], int comps, void* spec, bool pair_rule) {
  int rc;
  const double inv_n = 1.0 / ((double)g.n[0] * (double)g.n[1] * (double)g.n[2]);
  flip();
  Launch z = fwd_z(in, false);
  if (!empty(z) && (rc = be->launch(z))) return rc;
  if ((rc = bar_if(row_exchanges()))) return rc;
  Launch y = fwd_y();
  if (!empty(y) && (rc = be->launch(y))) return rc;
  if ((rc = bar_if(col_exchanges()))) return rc;
  Launch x = fwd_x(spec);
  x.desc.op = OP_SCALE;
  x.desc.scale = inv_n;
  if (!empty(x) && (rc = be->launch(x))) return rc;
  if (batch >= 3 && nprocs > 1) {
    // vector-field batch: the inverses of the selected components go through ONE set of exchanges (field set `ax` each)
    flip();
    for (int ax = 0; ax < 3; ++ax) {
      if (!(comps & (1 << ax))) continue;
      FieldScope fs(*this, ax);
      Launch xi = inv_x(spec);
      xi.desc.op = OP_GRAD;
      xi.desc.opAxis = ax;
      xi.desc.scale = 1.0;
      xi.desc.opOnLoad = 1;
      xi.desc.opFullAxis = pair_rule ? 0 : 1;
      if (!empty(xi) && (rc = be->launch(xi))) return rc;
    }
    if ((rc = bar_if(col_exchanges()))) return rc;
    for (int ax = 0; ax < 3; ++ax) {
      if (!(comps & (1 << ax))) continue;
      FieldScope fs(*this, ax);
      Launch yi = inv_y();
      if (!empty(yi) && (rc = be->launch(yi))) return rc;
    }
    if ((rc = bar_if(row_exchanges()))) return rc;
    for (int ax = 0; ax < 3; ++ax) {
      if (!(comps & (1 << ax))) continue;
      FieldScope fs(*this, ax);
      Launch zi = inv_z(outs[ax], false);
      if (!empty(zi) && (rc = be->launch(zi))) return rc;
    }
    return 0;
  }
  for (int ax = 0; ax < 3; ++ax) {
    if (!(comps & (1 << ax))) continue;
    flip();
    Launch xi = inv_x(spec);
    xi.desc.op = OP_GRAD;
    xi.desc.opAxis = ax;
    xi.desc.scale = 1.0;
    xi.desc.opOnLoad = 1;
    xi.desc.opFullAxis = pair_rule ? 0 : 1;
    if (!empty(xi) && (rc = be->launch(xi))) return rc;
    if ((rc = bar_if(col_exchanges()))) return rc;
    Launch yi = inv_y();
    if (!empty(yi) && (rc = be->launch(yi))) return rc;
    if ((rc = bar_if(row_exchanges()))) return rc;
    Launch zi = inv_z(outs[ax], false);
    if (!empty(zi) && (rc = be->launch(zi))) return rc;
  }
  return 0;
}

int Plan::execute_divergence(const void* const ins[3], void* out, void* spec, bool pair_rule) {
  int rc;
  const double inv_n = 1.0 / ((double)g.n[0] * (double)g.n[1] * (double)g.n[2]);
  if (batch >= 3 && nprocs > 1) {
    // vector-field batch: the three forward transforms share one set of exchanges; the X stages accumulate in stream order
    flip();
    for (int ax = 0; ax < 3; ++ax) {
      FieldScope fs(*this, ax);
      Launch z = fwd_z(ins[ax], false);
      if (!empty(z) && (rc = be->launch(z))) return rc;
    }
    if ((rc = bar_if(row_exchanges()))) return rc;
    for (int ax = 0; ax < 3; ++ax) {
      FieldScope fs(*this, ax);
      Launch y = fwd_y();
      if (!empty(y) && (rc = be->launch(y))) return rc;
    }
    if ((rc = bar_if(col_exchanges()))) return rc;
    for (int ax = 0; ax < 3; ++ax) {
      FieldScope fs(*this, ax);
      Launch x = fwd_x(spec);
      x.desc.op = OP_GRAD;
      x.desc.opAxis = ax;
      x.desc.scale = inv_n;
      x.desc.opFullAxis = pair_rule ? 0 : 1;
      x.desc.accumulate = ax > 0;
      if (!empty(x) && (rc = be->launch(x))) return rc;
    }
  } else
  for (int ax = 0; ax < 3; ++ax) {
    flip();
    Launch z = fwd_z(ins[ax], false);
    if (!empty(z) && (rc = be->launch(z))) return rc;
    if ((rc = bar_if(row_exchanges()))) return rc;
    Launch y = fwd_y();
    if (!empty(y) && (rc = be->launch(y))) return rc;
    if ((rc = bar_if(col_exchanges()))) return rc;
    Launch x = fwd_x(spec);
    x.desc.op = OP_GRAD;
    x.desc.opAxis = ax;
    x.desc.scale = inv_n;
    x.desc.opFullAxis = pair_rule ? 0 : 1;
    x.desc.accumulate = ax > 0;
    if (!empty(x) && (rc = be->launch(x))) return rc;
  }
  flip();
  Launch xi = inv_x(spec);
  if (!empty(xi) && (rc = be->launch(xi))) return rc;
  if ((rc = bar_if(col_exchanges()))) return rc;
  Launch yi = inv_y();
  if (!empty(yi) && (rc = be->launch(yi))) return rc;
  if ((rc = bar_if(row_exchanges()))) return rc;
  Launch zi = inv_z(out, false);
  if (!empty(zi) && (rc = be->launch(zi))) return rc;
  return 0;
}

// ================================================================================================================
// single-axis real transforms (reference: accfft_execute_{r2c,c2r}_{x,y,z}_gpu_t, src/accfft_gpu.cpp:1212-1604) and the
// per-axis spectral derivative built from them (src/operators_gpu.txx:172-278).
//   y: real [i0][i1][N2] --row group, N2 split--> [i0][N1][n2b]  -> spectrum [i0][N1/2+1][n2b]      (osize_y / osize_yi)
//   x: real [i0][M = i1*N2] --column group, M split--> [N0][mb]  -> spectrum [N0/2+1][mb]           (osize_x / osize_xi)
//   z: local, spectrum [i0][i1][N2c]                                                                  (osize_0)
// Buffer discipline (same invariant as the 3-D path, section 3 of DESIGN.md): peers' A may be written at any time but is
// never read after the LAST barrier of an operation; peers' B is never written before the FIRST barrier of an operation.
// ================================================================================================================
Launch Plan::ax_copy_in(int axis, const void* in) const {
  Launch l{};
  const size_t rs = prec == PREC_DOUBLE ? 8 : 4;
  StageDesc d = blank();
  d.in = in;
  d.nA = g.i0.size;
  d.inStrideI = 1;
  if (axis == 1) {
    const int P1 = g.np[1];
    d.nB = g.i1.size;
    d.inStrideA = (long long)g.i1.size * g.n[2];
    d.inStrideB = g.n[2];
    l.copyN = g.n[2];
    set_segs_common(d, P1, ceil_blk(g.n[2], P1));
    for (int p = 0; p < P1; ++p) {
      const Blk n2p = block_of(g.n[2], P1, p);
      const int r = rank_of(g.np, g.coords[0], p);
      d.seg[p].base = offs(peerA[r], (long long)g.i1.start * n2p.size, rs);
      d.seg[p].strideA = (long long)g.n[1] * n2p.size;
      d.seg[p].strideB = n2p.size;
      d.seg[p].strideI = 1;
    }
  } else {
    const int P0 = g.np[0];
    d.nB = g.M > 0 ? 1 : 0;
    d.inStrideA = g.M;
    d.inStrideB = 0;
    l.copyN = g.M;
    set_segs_common(d, P0, std::max(1, ceil_blk((int)g.M, P0)));
    for (int q = 0; q < P0; ++q) {
      const Blk mq = block_of((int)g.M, P0, q);
      const int r = rank_of(g.np, q, g.coords[1]);
      d.seg[q].base = offs(peerA[r], (long long)g.i0.start * mq.size, rs);
      d.seg[q].strideA = mq.size;
      d.seg[q].strideB = 0;
      d.seg[q].strideI = 1;
    }
  }
  l.kind = K_COPY_REAL;
  l.prec = prec;
  l.desc = d;
  l.slot = 7;
  return l;
}

Launch Plan::ax_copy_out(void* out) const {
  Launch l{};
  StageDesc d = blank();
  const int self = rank_of(g.np, g.coords[0], g.coords[1]);
  d.in = peerB.empty() ? nullptr : peerB[self];
  d.nA = g.i0.size;
  d.nB = g.i1.size;
  d.inStrideA = (long long)g.i1.size * g.n[2];
  d.inStrideB = g.n[2];
  d.inStrideI = 1;
  l.copyN = g.n[2];
  set_segs_common(d, 1, g.n[2]);
  d.seg[0].base = out;
  d.seg[0].strideA = d.inStrideA;
  d.seg[0].strideB = d.inStrideB;
  d.seg[0].strideI = 1;
  l.kind = K_COPY_REAL;
  l.prec = prec;
  l.desc = d;
  l.slot = 7;
  return l;
}

Launch Plan::ax_stage(int axis, int mode, const void* in, void* out) const {
  Launch l{};
  const size_t rs = prec == PREC_DOUBLE ? 8 : 4;
  StageDesc d = blank();
  const int self = rank_of(g.np, g.coords[0], g.coords[1]);
  const int N = g.n[axis];
  const bool ex = ax_exchanges(axis);
  const long long cols = axis == 1 ? (ex ? g.n2b.size : g.n[2]) : (ex ? (long long)g.mb.size : g.M);
  d.nA = axis == 1 ? g.i0.size : 1;
  d.nB = (int)cols;
  d.tw = be->twiddle(prec, N);
  d.inStrideB = 1;
  if (mode == AX_C2R) {   // complex side in: the caller's half spectrum [a][N/2+1][cols]
    d.in = in;
    d.inStrideA = axis == 1 ? (long long)(N / 2 + 1) * cols : 0;
    d.inStrideI = cols;
  } else {                // real side in: own A after the exchange, or the caller's block when the axis is local
    d.in = ex ? (peerA.empty() ? nullptr : peerA[self]) : in;
    d.inStrideA = axis == 1 ? (long long)N * cols : 0;
    d.inStrideI = cols;
  }
  if (mode == AX_R2C) {   // complex side out: the caller's buffer
    set_segs_common(d, 1, N);
    d.seg[0].base = out;
    d.seg[0].strideA = axis == 1 ? (long long)(N / 2 + 1) * cols : 0;
    d.seg[0].strideB = 1;
    d.seg[0].strideI = cols;
  } else if (!ex) {       // real side out, local
    set_segs_common(d, 1, N);
    d.seg[0].base = out;
    d.seg[0].strideA = axis == 1 ? (long long)N * cols : 0;
    d.seg[0].strideB = 1;
    d.seg[0].strideI = cols;
  } else if (axis == 1) { // real side out: row peers' B in the physical layout [i0][i1][N2], entered at column n2b.start
    const int P1 = g.np[1];
    set_segs_common(d, P1, ceil_blk(N, P1));
    for (int p = 0; p < P1; ++p) {
      const Blk i1p = block_of(N, P1, p);
      const int r = rank_of(g.np, g.coords[0], p);
      d.seg[p].base = offs(peerB[r], g.n2b.start, rs);
      d.seg[p].strideA = (long long)i1p.size * g.n[2];
      d.seg[p].strideB = 1;
      d.seg[p].strideI = g.n[2];
    }
  } else {                // column peers' B, physical layout [i0][M], entered at flat index mb.start
    const int P0 = g.np[0];
    set_segs_common(d, P0, ceil_blk(N, P0));
    for (int q = 0; q < P0; ++q) {
      const int r = rank_of(g.np, q, g.coords[1]);
      d.seg[q].base = offs(peerB[r], g.mb.start, rs);
      d.seg[q].strideA = 0;
      d.seg[q].strideB = 1;
      d.seg[q].strideI = g.M;
    }
  }
  if (mode == AX_DERIV) {
    d.op = OP_GRAD;
    d.opAxis = axis;
    d.kind[axis] = 0;
    d.n[axis] = N;
    d.scale = 1.0 / (double)N;
  }
  l.kind = mode == AX_R2C ? K_STRIDED_R2C : mode == AX_C2R ? K_STRIDED_C2R : K_STRIDED_DERIV;
  l.prec = prec;
  l.fftN = N;
  l.inverse = mode == AX_C2R;
  l.desc = d;
  l.slot = 8;
  return l;
}

// Z pass alone, caller's buffers on both sides: dense [i0][i1][N2] real <-> dense [i0][i1][N2c] complex
static Launch z_local(const Plan& p, bool inverse, const void* in, void* out) {
  Launch l = inverse ? p.inv_z(out, false) : p.fwd_z(in, false);
  StageDesc& d = l.desc;
  const Geometry& g = p.g;
  if (inverse) {
    d.in = in;
    d.inStrideA = (long long)g.n2c * g.i1.size;
    d.inStrideB = g.n2c;
  } else {
    d.nseg = 1;
    d.blk = g.n2c;
    d.blkMagic = blk_magic(d.blk);
    d.seg[0].base = out;
    d.seg[0].strideA = (long long)g.n2c * g.i1.size;
    d.seg[0].strideB = g.n2c;
    d.seg[0].strideI = 1;
  }
  l.slot = 8;
  return l;
}

int Plan::execute_r2c_axis(int axis, const void* in, void* out) {
  int rc;
  if (!ax_supported(axis)) {
    error = "single-axis transforms need a scheduled kernel for the axis length (accfftb_length_fast) and an even N2";
    return -3;
  }
  if (axis == 2) {
    Launch z = z_local(*this, false, in, out);
    return empty(z) ? 0 : be->launch(z);
  }
  const bool ex = ax_exchanges(axis);
  if (ex) {
    flip();
    if ((rc = a_fence())) return rc;
    Launch c = ax_copy_in(axis, in);
    if (!empty(c) && (rc = be->launch(c))) return rc;
    if ((rc = bar())) return rc;
  }
  Launch s = ax_stage(axis, AX_R2C, in, out);
  if (!empty(s) && (rc = be->launch(s))) return rc;
  if (ex && (rc = bar_if(false))) return rc;   // A was read after the exchange barrier: fence it from the next operation (reuse only)
  return 0;
}

int Plan::execute_c2r_axis(int axis, const void* in, void* out) {
  int rc;
  if (!ax_supported(axis)) {
    error = "single-axis transforms need a scheduled kernel for the axis length (accfftb_length_fast) and an even N2";
    return -3;
  }
  if (axis == 2) {
    Launch z = z_local(*this, true, in, out);
    return empty(z) ? 0 : be->launch(z);
  }
  const bool ex = ax_exchanges(axis);
  if (ex) flip();
  if (ex && (rc = bar_if(false))) return rc;   // B is written first thing: wait for the peers' previous operation (reuse only)
  Launch s = ax_stage(axis, AX_C2R, in, out);
  if (!empty(s) && (rc = be->launch(s))) return rc;
  if (ex) {
    if ((rc = bar())) return rc;
    Launch c = ax_copy_out(out);
    if (!empty(c) && (rc = be->launch(c))) return rc;
  }
  return 0;
}

int Plan::execute_deriv_axis(int axis, const void* in, void* out, bool accumulate, void* spec) {
  int rc;
  if (axis == 2) {
    if (be->supported(K_CONTIG_DERIV, prec, g.n[2])) {
      // one kernel: pairs of pencils as complex rows, forward -> i*k_z/N2 -> inverse (stage_real.cuh: contig_deriv_body)
      Launch z = fwd_z(in, false);
      StageDesc& d = z.desc;
      d.tw = be->twiddle(prec, g.n[2]);
      d.nseg = 1;
      d.blk = g.n[2];
      d.blkMagic = blk_magic(d.blk);
      d.seg[0].base = out;
      d.seg[0].strideA = d.inStrideA;   // real [i0][i1][N2] on both sides
      d.seg[0].strideB = d.inStrideB;
      d.seg[0].strideI = 1;
      d.op = OP_GRAD;
      d.opAxis = 2;
      d.kind[2] = 0;
      d.n[2] = g.n[2];
      d.scale = 1.0 / (double)g.n[2];
      d.accumulate = accumulate ? 1 : 0;
      z.kind = K_CONTIG_DERIV;
      z.fftN = g.n[2];
      z.slot = 8;
      return empty(z) ? 0 : be->launch(z);
    }
    // N2 beyond the longest complex schedule: r2c into the scratch spectrum, then c2r with i*k_z/N2 applied to its loads
    Launch z = z_local(*this, false, in, spec);
    if (!empty(z) && (rc = be->launch(z))) return rc;
    Launch zi = z_local(*this, true, spec, out);
    zi.desc.op = OP_GRAD;
    zi.desc.opAxis = 2;
    zi.desc.kind[2] = 0;
    zi.desc.n[2] = g.n[2];
    zi.desc.scale = 1.0 / (double)g.n[2];
    zi.desc.accumulate = accumulate ? 1 : 0;
    if (!empty(zi) && (rc = be->launch(zi))) return rc;
    return 0;
  }
  const bool ex = ax_exchanges(axis);
  if (ex) {
    flip();
    if ((rc = a_fence())) return rc;
    Launch c = ax_copy_in(axis, in);
    if (!empty(c) && (rc = be->launch(c))) return rc;
    if ((rc = bar())) return rc;
  }
  Launch s = ax_stage(axis, AX_DERIV, in, out);
  if (!ex) s.desc.accumulate = accumulate ? 1 : 0;
  if (!empty(s) && (rc = be->launch(s))) return rc;
  if (ex) {
    if ((rc = bar())) return rc;
    Launch c = ax_copy_out(out);
    c.desc.accumulate = accumulate ? 1 : 0;
    if (!empty(c) && (rc = be->launch(c))) return rc;
  }
  return 0;
}

int Plan::execute_grad_axes(const void* in, void* const outs[3], int comps, void* spec) {
  // an axis length without per-axis kernels: the full 3-D scheme with the per-axis scheme's Nyquist rule is the same operator
  for (int ax = 0; ax < 3; ++ax)
    if (ax < 2 && (comps & (1 << ax)) && !ax_supported(ax)) return execute_grad(in, outs, comps, spec, true);   // (z is always local)
  for (int ax = 0; ax < 3; ++ax) {
    if (!(comps & (1 << ax))) continue;
    if (int rc = execute_deriv_axis(ax, in, outs[ax], false, spec)) return rc;
  }
  return 0;
}

int Plan::execute_divergence_axes(const void* const ins[3], void* out, void* spec) {
  for (int ax = 0; ax < 2; ++ax)
    if (!ax_supported(ax)) return execute_divergence(ins, out, spec, true);
  // z first: its fused kernel runs few warps per SM and would expose the latency of the read-modify-write that
  // accumulation needs (measured: 3.11 ms with z last vs 2.5 ms at 512^3 double); the strided x / y kernels hide it
  const int order[3] = {2, 0, 1};
  for (int i = 0; i < 3; ++i)
    if (int rc = execute_deriv_axis(order[i], ins[order[i]], out, i > 0, spec)) return rc;
  return 0;
}

}  // namespace accfft_b200
