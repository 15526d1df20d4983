// emul_fiber.hpp -- TEST INFRASTRUCTURE ONLY: the fiber CTA of the CPU emulator (see emul.cpp).  Each CUDA thread of a CTA
// runs as a ucontext fiber; __syncthreads == yield to a round-robin scheduler.  Included by every emulator translation unit.
#pragma once
#include <string.h>
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#include "../../accfft_b200/csrc/kernel_cfg.hpp"
#include "../../accfft_b200/csrc/plan.hpp"

using namespace accfft_b200;

// ---------------------------------------------------------------------------------------------------
// fiber CTA
// ---------------------------------------------------------------------------------------------------
struct CtaEmu {
  int nthreads = 0;
  std::vector<ucontext_t> ctx;
  std::vector<char*> stacks;
  std::vector<char> done;
  ucontext_t sched;
  int cur = 0;
  std::function<void(int)> body;
  static constexpr size_t kStack = 256 * 1024;
  ~CtaEmu() {
    for (char* s : stacks) free(s);
  }
};
static thread_local CtaEmu* g_cta = nullptr;

static void fiber_tramp() {
  CtaEmu* c = g_cta;
  const int tid = c->cur;
  c->body(tid);
  c->done[tid] = 1;
  swapcontext(&c->ctx[tid], &c->sched);
}
static void fiber_sync() {
  CtaEmu* c = g_cta;
  const int tid = c->cur;
  swapcontext(&c->ctx[tid], &c->sched);
}
static void run_cta(CtaEmu& c, int nthreads, std::function<void(int)> body) {
  g_cta = &c;
  c.body = body;
  if ((int)c.stacks.size() < nthreads) {
    for (int i = (int)c.stacks.size(); i < nthreads; ++i) c.stacks.push_back((char*)malloc(CtaEmu::kStack));
  }
  c.ctx.resize(nthreads);
  c.done.assign(nthreads, 0);
  c.nthreads = nthreads;
  for (int i = 0; i < nthreads; ++i) {
    getcontext(&c.ctx[i]);
    c.ctx[i].uc_stack.ss_sp = c.stacks[i];
    c.ctx[i].uc_stack.ss_size = CtaEmu::kStack;
    c.ctx[i].uc_link = &c.sched;
    makecontext(&c.ctx[i], fiber_tramp, 0);
  }
  bool alive = true;
  while (alive) {
    alive = false;
    for (int i = 0; i < nthreads; ++i) {
      if (c.done[i]) continue;
      c.cur = i;
      swapcontext(&c.sched, &c.ctx[i]);
      if (!c.done[i]) alive = true;
    }
  }
}

struct EmuCtx {
  int tid_;
  long long bid_;
  void* smem_;
  int tid() const { return tid_; }
  long long bid() const { return bid_; }
  void* smem() const { return smem_; }
  void sync() const { fiber_sync(); }
  // bulk copies: same 16-byte rules as cp.async.bulk, enforced here so the emulator catches misuse
  void bulk_fence() const {}
  void bulk_store(void* gdst, const void* ssrc, unsigned bytes) const {
    if (((uintptr_t)gdst | (uintptr_t)ssrc | bytes) & 15u) {
      fprintf(stderr, "emul: misaligned bulk copy\n");
      abort();
    }
    memcpy(gdst, ssrc, bytes);
  }
  void bulk_wait() const {}
};

template <class Body>
static void run_grid(long long nblocks, int nthreads, size_t smem_bytes, Body body) {
  CtaEmu cta;
  std::vector<char> smem(smem_bytes + 64);
  for (long long b = 0; b < nblocks; ++b) {
    memset(smem.data(), 0xCD, smem.size());
    run_cta(cta, nthreads, [&](int tid) {
      EmuCtx ctx{tid, b, (void*)smem.data()};
      body(ctx);
    });
  }
}

