// fused_sched.hpp -- the work list of the fused two-pass kernel (fused.cuh): its parameters, how a work item decodes into a tile
// and its dependency, and how the parameters follow from a stage pair.  Host + device (the CPU test tests/test_fused_protocol.py
// drives exactly this code through a simulation of the kernel's claim / signal / wait protocol).
#pragma once
#include "fft_engine.cuh"   // ACC_HD

namespace accfft_b200 {

struct FusedArgs {
  int* ctr;          // [0] work counter, [1 .. nch] pass-1 tiles done per chunk, [1 + nch .. 2 nch] pass-2 tiles done per chunk
  int nch, cp;       // chunks, planes per chunk
  int lag, ns;       // pass 2 runs `lag` chunks behind pass 1; ring slots (in chunks), ns > lag
  int t1, t2;        // tiles per chunk of pass 1 / pass 2
  int tpa;           // strided pass: tiles per plane
  int tz;            // contiguous pass: tiles per chunk (= cp * pencils per plane / PCZ)
  int zfirst;        // 1: pass 1 is the contiguous (Z) pass (forward), 0: the strided one (inverse)
  int ztiles, yplanes;   // tiles of the contiguous pass / planes of the strided pass that exist (the last chunk may be ragged)
  int flags;         // debugging: bit 0 = signal a tile's completion right away (not from the next tile's hook),
                     //            bit 1 = no early dependency check (every tile polls)
                     //            bit 2 = TIMING EXPERIMENT, results invalid: no dependency waits, no signals
                     //            bit 3 = static work assignment (item = CTA + k * grid) instead of the global counter
  long long total;   // work items = (nch + lag) * (t1 + t2)
};

// what the scheduler publishes for a work item
enum FusedKind : int { FK_END = 0, FK_DEAD = 1, FK_EMPTY = 2, FK_Z = 3, FK_Y = 4 };
struct FusedSlot {
  int kind;    // FK_END: list exhausted; FK_DEAD: chunk out of range, nothing to do; FK_EMPTY: a tile beyond the ragged end (signals
               // only); FK_Z / FK_Y: run the contiguous / strided body on tile `bid`
  int ready;   // its dependency was found satisfied by the early check
  int bid;     // tile index in the body's own numbering
  int sig;     // index into ctr of the counter the tile increments when done
  int dep;     // index into ctr of the counter it waits for (0: none) ...
  int target;  // ... and the value that counter must reach
};

ACC_HD FusedSlot fused_decode(const FusedArgs& fa, int w) {
  FusedSlot it;
  it.ready = 0, it.bid = 0, it.sig = 0, it.dep = 0, it.target = 0;
  if ((long long)w >= fa.total) {
    it.kind = FK_END;
    return it;
  }
  const int per = fa.t1 + fa.t2;
  const int s = w / per, i = w - s * per;
  const bool ph1 = i < fa.t1;
  const int c = ph1 ? s : s - fa.lag;
  const int t = ph1 ? i : i - fa.t1;
  if (c < 0 || c >= fa.nch) {
    it.kind = FK_DEAD;
    return it;
  }
  if (ph1) {
    it.sig = 1 + c;
    if (c >= fa.ns) it.dep = 1 + fa.nch + (c - fa.ns), it.target = fa.t2;
  } else {
    it.sig = 1 + fa.nch + c;
    it.dep = 1 + c, it.target = fa.t1;
  }
  const bool ztype = ph1 == (fa.zfirst != 0);
  if (ztype) {
    it.bid = c * fa.tz + t;
    it.kind = it.bid < fa.ztiles ? FK_Z : FK_EMPTY;
  } else {
    it.bid = c * fa.cp * fa.tpa + t;
    it.kind = it.bid / fa.tpa < fa.yplanes ? FK_Y : FK_EMPTY;
  }
  return it;
}

// Host side: work-list parameters for a stage pair.  nA planes; nBz pencils per plane of the contiguous pass, PCZ of them per tile;
// the strided pass has tpa tiles per plane; `grid` CTAs will be resident.  Returns false when the pair cannot be fused.
struct FusedShape {
  int cp, lag, ns, ring_planes;   // ring_planes == 0: the ring spans the whole buffer (plain addressing)
};
inline bool fused_make_args(int nA, int nBz, int PCZ, int tpa, long long grid, int tune_cp, int tune_lag, int tune_ns, int flags,
                            bool zfirst, int ctr_ints, FusedArgs& fa, FusedShape& sh) {
  if (nA <= 0 || nBz <= 0 || PCZ <= 0 || tpa <= 0 || grid <= 0) return false;
  // planes per chunk: a power of two (the ring is addressed with a mask) whose pencils fill whole contiguous-pass tiles
  int cp = 1;
  while (cp < (tune_cp > 0 ? tune_cp : 2)) cp *= 2;
  while (cp <= 64 && ((long long)cp * nBz) % PCZ != 0) cp *= 2;
  if (cp > 64) return false;
  fa.cp = cp;
  fa.nch = (nA + cp - 1) / cp;
  fa.tpa = tpa;
  fa.tz = (int)(((long long)cp * nBz) / PCZ);
  const int ty = cp * tpa;
  fa.t1 = zfirst ? fa.tz : ty;
  fa.t2 = zfirst ? ty : fa.tz;
  const int per = fa.t1 + fa.t2;
  // the second pass runs far enough behind the first that a tile's dependency is complete when it is claimed: 1.5 x the
  // resident CTAs' worth of work items
  int lag = tune_lag > 0 ? tune_lag : (int)((3 * grid / 2 + per - 1) / per);
  if (lag < 1) lag = 1;
  int ns = tune_ns > lag ? tune_ns : lag + 2;
  {   // ring slots: a power of two, and more than `lag` (a slot is rewritten only by tiles claimed after its readers)
    int p2 = 1;
    while (p2 < ns) p2 *= 2;
    ns = p2;
  }
  fa.lag = lag;
  fa.ns = ns;
  fa.zfirst = zfirst ? 1 : 0;
  fa.ztiles = (int)(((long long)nA * nBz + PCZ - 1) / PCZ);
  fa.yplanes = nA;
  fa.flags = flags;
  if ((long long)(fa.nch + lag) * per > 0x3fffffffLL) return false;
  fa.total = (long long)(fa.nch + lag) * per;
  if (1 + 2 * fa.nch > ctr_ints) return false;
  const long long ring = (long long)ns * cp;          // planes, a power of two
  sh.cp = cp, sh.lag = lag, sh.ns = ns;
  sh.ring_planes = ring >= nA ? 0 : (int)ring;
  return true;
}

}  // namespace accfft_b200
