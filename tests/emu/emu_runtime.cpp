// emu_runtime.cpp -- TEST INFRASTRUCTURE: runs one "warp" of the engine's device code as 32 ucontext fibers.
// Used only by the CPU (-m "not gpu") tests to exercise the tree / rules / feature device code without a GPU;
// never part of libagz.so.  See alphago.jl_b200/csrc/simt.h.
#define AGZ_EMU 1
#include "../../alphago.jl_b200/csrc/simt.h"

#include <stdlib.h>

namespace simt {
thread_local EmuWarp* g_warp = nullptr;

static const size_t kStack = 512 * 1024;

static void fiber_entry() {
  EmuWarp* w = g_warp;
  w->fn(w->arg);
  int me = w->cur;
  if (me < 31) {
    w->cur = me + 1;
    setcontext(&w->ctx[me + 1]);
  } else {
    setcontext(&w->main_ctx);
  }
}

void emu_run_warp(void (*fn)(void*), void* arg) {
  static thread_local EmuWarp* warp = nullptr;
  static thread_local char* stacks = nullptr;
  if (!warp) {
    warp = new EmuWarp();
    stacks = (char*)malloc(kStack * 32);
  }
  EmuWarp* w = warp;
  g_warp = w;
  w->fn = fn;
  w->arg = arg;
  for (int i = 0; i < 32; ++i) {
    w->ncoll[i] = 0;
    getcontext(&w->ctx[i]);
    w->ctx[i].uc_stack.ss_sp = stacks + kStack * i;
    w->ctx[i].uc_stack.ss_size = kStack;
    w->ctx[i].uc_link = nullptr;
    makecontext(&w->ctx[i], fiber_entry, 0);
  }
  w->cur = 0;
  swapcontext(&w->main_ctx, &w->ctx[0]);
}
}  // namespace simt
