// fused_sim.cpp -- CPU model of the fused two-pass kernel's scheduling protocol (csrc/fused.cuh), driving the REAL work-list code
// (csrc/fused_sched.hpp: fused_make_args, fused_decode).  Test infrastructure only.
//
// `nctas` simulated CTAs run the kernel's loop -- claim the next item from the global counter, decode it, check its dependency
// counter once ("early check"), hold the completion signal of the previous tile back until the next tile's scheduler run, poll
// only after flushing that pending signal -- under a seeded random interleaving (every step advances one randomly chosen CTA by
// one protocol action).  Checked:
//   * a tile's body starts only when its dependency counter has reached its target (read-after-write on the ring, and
//     write-after-read: a ring slot is rewritten only after every reader of its previous contents has signalled),
//   * every tile of every chunk runs exactly once and every counter ends at its target,
//   * the system never stalls with work left (no deadlock), for any number of CTAs.
// Returns 0, or a negative code naming the violated property.
#include <stdint.h>
#include <vector>
#include "fused_sched.hpp"

using namespace accfft_b200;

namespace {
struct Rng {
  uint64_t s;
  uint64_t next() {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
};
enum Phase { P_TOP, P_POLL, P_BODY_LOADS, P_BODY_REST, P_DONE };
struct Cta {
  int phase = P_TOP;
  int pending = 0;       // ctr index of the previous tile's signal, not yet sent
  int claim = 0;         // item whose decode is published for the next iteration
  FusedSlot cur{}, next{};
  bool first = true;
};
}  // namespace

extern "C" int emu_fused_protocol_sim(int nA, int nBz, int PCZ, int tpa, int nctas, int tune_cp, int tune_lag, int tune_ns, int zfirst,
                                      unsigned long long seed, int flags, int* out_shape /* cp, lag, ns, ring, nch */) {
  FusedArgs fa;
  FusedShape sh;
  std::vector<int> ctr;
  if (!fused_make_args(nA, nBz, PCZ, tpa, nctas, tune_cp, tune_lag, tune_ns, flags, zfirst != 0, 1 << 24, fa, sh)) return -1;
  ctr.assign(1 + 2 * fa.nch, 0);
  fa.ctr = ctr.data();
  if (out_shape) out_shape[0] = sh.cp, out_shape[1] = sh.lag, out_shape[2] = sh.ns, out_shape[3] = sh.ring_planes, out_shape[4] = fa.nch;
  if (sh.ns <= sh.lag) return -2;   // the deadlock-freedom argument needs ns > lag
  std::vector<int> ran_z(fa.ztiles, 0), ran_y((size_t)fa.yplanes * fa.tpa, 0);
  std::vector<Cta> c(nctas);
  Rng rng{seed};
  // sched(): decode the claimed item, early dependency check, flush the pending signal, publish
  auto sched = [&](Cta& k) {
    FusedSlot it = fused_decode(fa, k.claim);
    int depv = 0;
    if (it.dep && !(flags & 2)) depv = ctr[it.dep];
    if (k.pending) ctr[k.pending]++, k.pending = 0;
    it.ready = (it.dep == 0 || (depv >= it.target && !(flags & 2))) ? 1 : 0;
    k.next = it;
  };
  for (Cta& k : c) {
    k.claim = ctr[0]++;
    sched(k);
  }
  long long live = nctas, idle_steps = 0;
  const long long max_steps = 64LL * (fa.total + nctas) + 1000000;
  for (long long step = 0; live > 0; ++step) {
    if (step > max_steps) return -3;   // never terminates
    Cta& k = c[rng.next() % nctas];
    bool progressed = true;
    switch (k.phase) {
      case P_TOP:
        k.cur = k.next;
        if (k.cur.kind == FK_END) {
          if (k.pending) ctr[k.pending]++, k.pending = 0;
          k.phase = P_DONE;
          --live;
          break;
        }
        if (k.cur.kind != FK_DEAD && !k.cur.ready) {
          if (k.pending) ctr[k.pending]++, k.pending = 0;   // flush before waiting
          k.phase = P_POLL;
          break;
        }
        k.claim = ctr[0]++;
        k.phase = P_BODY_LOADS;
        break;
      case P_POLL:
        if (ctr[k.cur.dep] >= k.cur.target) {
          k.claim = ctr[0]++;
          k.phase = P_BODY_LOADS;
        } else {
          progressed = false;
        }
        break;
      case P_BODY_LOADS:
        // the tile's loads are issued here: its dependency must hold NOW
        if (k.cur.kind == FK_Z || k.cur.kind == FK_Y || k.cur.kind == FK_EMPTY) {
          if (k.cur.dep && ctr[k.cur.dep] < k.cur.target) return -4;   // started before its input was complete / its slot was free
          if (k.cur.kind == FK_Z) {
            if (k.cur.bid < 0 || k.cur.bid >= fa.ztiles) return -5;
            ran_z[k.cur.bid]++;
          } else if (k.cur.kind == FK_Y) {
            if (k.cur.bid < 0 || k.cur.bid >= (long long)fa.yplanes * fa.tpa) return -5;
            ran_y[k.cur.bid]++;
          }
        }
        sched(k);   // the after_loads hook (or the fall-back right after a skipped tile)
        k.phase = P_BODY_REST;
        break;
      case P_BODY_REST:
        if (k.cur.kind != FK_DEAD) {
          k.pending = k.cur.sig;
          if (flags & 1) ctr[k.pending]++, k.pending = 0;
        }
        k.phase = P_TOP;
        break;
      default:
        progressed = false;
        break;
    }
    idle_steps = progressed ? 0 : idle_steps + 1;
    if (idle_steps > 64LL * nctas + 4096) return -6;   // every CTA is waiting: deadlock
  }
  for (int v : ran_z)
    if (v != 1) return -7;
  for (int v : ran_y)
    if (v != 1) return -7;
  for (int ch = 0; ch < fa.nch; ++ch)
    if (ctr[1 + ch] != fa.t1 || ctr[1 + fa.nch + ch] != fa.t2) return -8;
  return 0;
}
