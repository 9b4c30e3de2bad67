// simt.h -- the handful of warp primitives the tree / rules / feature kernels use.
//
// Under nvcc these are the CUDA intrinsics.  Under a plain host compiler with -DAGZ_EMU (used ONLY by
// the CPU unit tests, tests/emu/, to run the very same device code where there is no GPU) a warp is 32
// ucontext fibers run round-robin and every collective is one full rotation.  The product library is
// always built by nvcc and contains no host execution path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(AGZ_EMU)
#define AGZ_CUDA 1
#include <cuda_runtime.h>
#define AGZ_DEV __device__ __forceinline__
#define AGZ_COLD __device__ __noinline__

namespace simt {
#if AGZ_LANE_ASM
AGZ_DEV int lane() { int l; asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }   // not rematerialisable: stays in a register
#else
AGZ_DEV int lane() { return threadIdx.x & 31; }
#endif
AGZ_DEV void sync() { __syncwarp(); }
AGZ_DEV unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
AGZ_DEV bool any(bool p) { return __any_sync(0xffffffffu, p); }
AGZ_DEV int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
AGZ_DEV unsigned shfl(unsigned v, int src) { return __shfl_sync(0xffffffffu, v, src); }
AGZ_DEV float shfl(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
AGZ_DEV double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
AGZ_DEV int shfl_xor(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
AGZ_DEV float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
AGZ_DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
AGZ_DEV int atomic_add(int* p, int v) { return atomicAdd(p, v); }
AGZ_DEV int atomic_or(int* p, int v) { return atomicOr(p, v); }
AGZ_DEV unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
AGZ_DEV unsigned reduce_or(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
AGZ_DEV int reduce_add(int v) { return __reduce_add_sync(0xffffffffu, v); }
AGZ_DEV int popc(unsigned x) { return __popc(x); }
AGZ_DEV int ffs(unsigned x) { return __ffs(x); }
// correctly-rounded ops that can never be contracted into an FMA
AGZ_DEV double dmul(double a, double b) { return __dmul_rn(a, b); }
AGZ_DEV double dadd(double a, double b) { return __dadd_rn(a, b); }
AGZ_DEV double dsub(double a, double b) { return __dsub_rn(a, b); }
AGZ_DEV double ddiv(double a, double b) { return __ddiv_rn(a, b); }
AGZ_DEV double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }   // ONE rounding: only where the algorithm asks for a fused operation
AGZ_DEV unsigned reduce_max(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }
AGZ_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
AGZ_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }
AGZ_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
AGZ_DEV float fsub(float a, float b) { return __fsub_rn(a, b); }
AGZ_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
AGZ_DEV float fsqrt(float a) { return __fsqrt_rn(a); }
AGZ_DEV double dfloor(double a) { return floor(a); }
AGZ_DEV long long dbits(double a) { return __double_as_longlong(a); }
AGZ_DEV unsigned fbits(float a) { return __float_as_uint(a); }
AGZ_DEV double bitsd(long long a) { return __longlong_as_double(a); }
AGZ_DEV uint32_t mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
// ---- kernel timeline trace (debug aid, AGZ_TRACE=<records>): tr[0] = records written, tr[1] = capacity, then 4 words per
// record: tag | block << 8 | grid << 32, start, end (%globaltimer ns), SM id.  Written by the first and last CTA of a launch.
AGZ_DEV unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
AGZ_DEV void trace_rec(unsigned long long* tr, int tag, unsigned long long t0) {
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const unsigned long long i = atomicAdd(tr, 1ULL);
  if (i < tr[1]) {
    tr[2 + 4 * i] = (unsigned long long)tag | ((unsigned long long)blockIdx.x << 8) | ((unsigned long long)gridDim.x << 32);
    tr[3 + 4 * i] = t0;
    tr[4 + 4 * i] = gtimer();
    tr[5 + 4 * i] = smid;
  }
}
AGZ_DEV bool trace_cta() { return threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1); }
}  // namespace simt

#else  // ------------------------------------------------------------------ host emulation (tests only)
#define AGZ_CUDA 0
#define AGZ_DEV inline
#define AGZ_COLD inline
#include <math.h>
#include <string.h>
#include <ucontext.h>

namespace simt {
struct EmuWarp {
  ucontext_t main_ctx;
  ucontext_t ctx[32];
  int cur;
  uint64_t slot[2][32];
  uint32_t ncoll[32];  // collectives issued per lane (parity selects the slot buffer)
  void (*fn)(void*);
  void* arg;
};
extern thread_local EmuWarp* g_warp;
void emu_run_warp(void (*fn)(void*), void* arg);  // runs fn on 32 fibers to completion (emu_runtime.cpp)

inline int lane() { return g_warp->cur; }
inline void barrier() {  // one full rotation: every lane runs up to its next barrier
  EmuWarp* w = g_warp;
  int me = w->cur;
  int nxt = (me + 1) & 31;
  w->cur = nxt;
  swapcontext(&w->ctx[me], &w->ctx[nxt]);
  w->cur = me;
}
inline void sync() { barrier(); }
template <class T>
inline T exchange_(T v, int src) {
  static_assert(sizeof(T) <= 8, "");
  EmuWarp* w = g_warp;
  int me = w->cur;
  int par = w->ncoll[me]++ & 1;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w->slot[par][me] = raw;
  barrier();
  T r;
  memcpy(&r, &w->slot[par][src & 31], sizeof(T));
  return r;
}
inline unsigned ballot(bool p) {
  EmuWarp* w = g_warp;
  int me = w->cur;
  int par = w->ncoll[me]++ & 1;
  w->slot[par][me] = p ? 1 : 0;
  barrier();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)(w->slot[par][i] & 1) << i;
  return m;
}
inline bool any(bool p) { return ballot(p) != 0; }
inline unsigned reduce_or(unsigned v) {
  EmuWarp* w = g_warp;
  int me = w->cur;
  int par = w->ncoll[me]++ & 1;
  w->slot[par][me] = v;
  barrier();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)w->slot[par][i];
  return m;
}
inline int reduce_add(int v) {
  EmuWarp* w = g_warp;
  int me = w->cur;
  int par = w->ncoll[me]++ & 1;
  w->slot[par][me] = (uint64_t)(uint32_t)v;
  barrier();
  int m = 0;
  for (int i = 0; i < 32; ++i) m += (int)(uint32_t)w->slot[par][i];
  return m;
}
inline int shfl(int v, int src) { return exchange_(v, src); }
inline unsigned shfl(unsigned v, int src) { return exchange_(v, src); }
inline float shfl(float v, int src) { return exchange_(v, src); }
inline double shfl(double v, int src) { return exchange_(v, src); }
inline int shfl_xor(int v, int m) { return exchange_(v, lane() ^ m); }
inline float shfl_xor(float v, int m) { return exchange_(v, lane() ^ m); }
inline double shfl_xor(double v, int m) { return exchange_(v, lane() ^ m); }
inline int atomic_add(int* p, int v) { int o = *p; *p = o + v; return o; }
inline int atomic_or(int* p, int v) { int o = *p; *p = o | v; return o; }
inline unsigned long long atomic_add(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int popc(unsigned x) { return __builtin_popcount(x); }
inline int ffs(unsigned x) { return __builtin_ffs((int)x); }
inline double dmul(double a, double b) { volatile double r = a * b; return r; }
inline double dadd(double a, double b) { volatile double r = a + b; return r; }
inline double dsub(double a, double b) { volatile double r = a - b; return r; }
inline double ddiv(double a, double b) { volatile double r = a / b; return r; }
inline double dfma(double a, double b, double c) { return fma(a, b, c); }
inline unsigned reduce_max(unsigned v) {
  EmuWarp* w = g_warp;
  int me = w->cur;
  int par = w->ncoll[me]++ & 1;
  w->slot[par][me] = v;
  barrier();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m = (unsigned)w->slot[par][i] > m ? (unsigned)w->slot[par][i] : m;
  return m;
}
inline void prefetch_l2(const void*) {}
inline float fdiv(float a, float b) { volatile float r = a / b; return r; }
inline float fadd(float a, float b) { volatile float r = a + b; return r; }
inline float fsub(float a, float b) { volatile float r = a - b; return r; }
inline float fmul(float a, float b) { volatile float r = a * b; return r; }
inline float fsqrt(float a) { return sqrtf(a); }
inline double dfloor(double a) { return floor(a); }
inline long long dbits(double a) { long long r; memcpy(&r, &a, 8); return r; }
inline unsigned fbits(float a) { unsigned r; memcpy(&r, &a, 4); return r; }
inline double bitsd(long long a) { double r; memcpy(&r, &a, 8); return r; }
inline uint32_t mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
}  // namespace simt
#endif
