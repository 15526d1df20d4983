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
  if (g.r2c) {
    if (g.n[2] % 2 != 0) {
      error = "r2c plans need an even N2 (odd-length real transforms are not implemented)";
      return false;
    }
    if (!be->supported(K_CONTIG_R2C, prec, g.n[2] / 2) || !be->supported(K_CONTIG_C2R, prec, g.n[2] / 2)) {
      error = "no kernel schedule for N2/2 = " + std::to_string(g.n[2] / 2);
      return false;
    }
  } else if (!be->supported(K_CONTIG_C2C, prec, g.n[2])) {
    error = "no kernel schedule for N2 = " + std::to_string(g.n[2]);
    return false;
  }
  if (!be->supported(K_STRIDED_C2C, prec, g.n[1])) {
    error = "no kernel schedule for N1 = " + std::to_string(g.n[1]);
    return false;
  }
  if (!be->supported(K_STRIDED_C2C, prec, g.n[0])) {
    error = "no kernel schedule for N0 = " + std::to_string(g.n[0]);
    return false;
  }
  return true;
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
  const long long m = std::max(a, std::max(b1, b2));
  return std::max<long long>(m * cs, 16);
}

static void no_transform(Launch& l, int n_axis);
static bool empty(const Launch& l) { return l.desc.nA <= 0 || l.desc.nB <= 0; }

int Plan::prepare() {
  Launch ls[6] = {fwd_z(nullptr, false), fwd_y(), fwd_x(nullptr), inv_x(nullptr), inv_y(), inv_z(nullptr, false)};
  for (Launch& l : ls) {
    l.dry = true;
    if (int rc = be->launch(l)) return rc;
  }
  Launch c = fwd_y();   // the no-transform stage of partial-axis executes
  no_transform(c, g.n[1]);
  c.dry = true;
  return be->launch(c);
}

// a stage whose axis is not selected only re-lays the data out (pack-on-store without the transform)
static void no_transform(Launch& l, int n_axis) {
  l.kind = K_COPY;
  l.fftN = n_axis;
}

int Plan::execute_forward(const void* in, void* out, bool in_padded, int xyz) {
  int rc;
  Launch z = fwd_z(in, in_padded);
  if (!(xyz & 4)) no_transform(z, g.n[2]);
  if (!empty(z) && (rc = be->launch(z))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch y = fwd_y();
  if (!(xyz & 2)) no_transform(y, g.n[1]);
  if (!empty(y) && (rc = be->launch(y))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch x = fwd_x(out);
  if (!(xyz & 1)) no_transform(x, g.n[0]);
  if (!empty(x) && (rc = be->launch(x))) return rc;
  return 0;
}

int Plan::execute_inverse(const void* in, void* out, bool out_padded, int xyz) {
  int rc;
  Launch x = inv_x(in);
  if (!(xyz & 1)) no_transform(x, g.n[0]);
  if (!empty(x) && (rc = be->launch(x))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch y = inv_y();
  if (!(xyz & 2)) no_transform(y, g.n[1]);
  if (!empty(y) && (rc = be->launch(y))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch z = inv_z(out, out_padded);
  if (!(xyz & 4)) no_transform(z, g.n[2]);
  if (!empty(z) && (rc = be->launch(z))) return rc;
  return 0;
}

int Plan::execute_spectral(const void* in, void* out, int op, void* spec) {
  int rc;
  Launch z = fwd_z(in, false);
  if (!empty(z) && (rc = be->launch(z))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch y = fwd_y();
  if (!empty(y) && (rc = be->launch(y))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  // X forward with the symbol multiply and the 1/N normalisation fused into its stores, into `spec`
  // (user spectrum layout), which the inverse X stage reads back
  Launch x = fwd_x(spec);
  x.desc.op = op;
  x.desc.scale = 1.0 / ((double)g.n[0] * (double)g.n[1] * (double)g.n[2]);
  if (!empty(x) && (rc = be->launch(x))) return rc;
  Launch xi = inv_x(spec);
  if (!empty(xi) && (rc = be->launch(xi))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch yi = inv_y();
  if (!empty(yi) && (rc = be->launch(yi))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch zi = inv_z(out, false);
  if (!empty(zi) && (rc = be->launch(zi))) return rc;
  return 0;
}

int Plan::execute_grad(const void* in, void* const outs[3], int comps, void* spec) {
  int rc;
  const double inv_n = 1.0 / ((double)g.n[0] * (double)g.n[1] * (double)g.n[2]);
  Launch z = fwd_z(in, false);
  if (!empty(z) && (rc = be->launch(z))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch y = fwd_y();
  if (!empty(y) && (rc = be->launch(y))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch x = fwd_x(spec);
  x.desc.op = OP_SCALE;
  x.desc.scale = inv_n;
  if (!empty(x) && (rc = be->launch(x))) return rc;
  for (int ax = 0; ax < 3; ++ax) {
    if (!(comps & (1 << ax))) continue;
    Launch xi = inv_x(spec);
    xi.desc.op = OP_GRAD;
    xi.desc.opAxis = ax;
    xi.desc.scale = 1.0;
    xi.desc.opOnLoad = 1;
    if (!empty(xi) && (rc = be->launch(xi))) return rc;
    if (nprocs > 1 && (rc = be->barrier())) return rc;
    Launch yi = inv_y();
    if (!empty(yi) && (rc = be->launch(yi))) return rc;
    if (nprocs > 1 && (rc = be->barrier())) return rc;
    Launch zi = inv_z(outs[ax], false);
    if (!empty(zi) && (rc = be->launch(zi))) return rc;
  }
  return 0;
}

int Plan::execute_divergence(const void* const ins[3], void* out, void* spec) {
  int rc;
  const double inv_n = 1.0 / ((double)g.n[0] * (double)g.n[1] * (double)g.n[2]);
  for (int ax = 0; ax < 3; ++ax) {
    Launch z = fwd_z(ins[ax], false);
    if (!empty(z) && (rc = be->launch(z))) return rc;
    if (nprocs > 1 && (rc = be->barrier())) return rc;
    Launch y = fwd_y();
    if (!empty(y) && (rc = be->launch(y))) return rc;
    if (nprocs > 1 && (rc = be->barrier())) return rc;
    Launch x = fwd_x(spec);
    x.desc.op = OP_GRAD;
    x.desc.opAxis = ax;
    x.desc.scale = inv_n;
    x.desc.accumulate = ax > 0;
    if (!empty(x) && (rc = be->launch(x))) return rc;
  }
  Launch xi = inv_x(spec);
  if (!empty(xi) && (rc = be->launch(xi))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch yi = inv_y();
  if (!empty(yi) && (rc = be->launch(yi))) return rc;
  if (nprocs > 1 && (rc = be->barrier())) return rc;
  Launch zi = inv_z(out, false);
  if (!empty(zi) && (rc = be->launch(zi))) return rc;
  return 0;
}

}  // namespace accfft_b200
