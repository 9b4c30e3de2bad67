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

namespace simt {
AGZ_DEV int lane() { return threadIdx.x & 31; }
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

// ---- lane groups.  G<32> is the warp itself.  G<16> is one half of a warp that plays two trees at once (tree_duo.cuh): the two
// halves run in LOCK STEP -- every collective below is executed by all 32 threads with the full member mask (one SHFL / VOTE / REDUX
// instruction, no WARPSYNC.EXCLUSIVE serialisation of the halves) and returns, to each thread, the result over its own half.  Code
// written against G<16> must therefore keep its control flow warp-uniform wherever a collective is reached: loops and guards use
// any_warp(), per-half conditions become predicates.
template <int W> struct G;
template <> struct G<32> {
  static constexpr int width = 32;
  AGZ_DEV static int lane() { return threadIdx.x & 31; }
  AGZ_DEV static int half() { return 0; }
  AGZ_DEV static void sync() { __syncwarp(); }
  AGZ_DEV static unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
  AGZ_DEV static bool any(bool p) { return __any_sync(0xffffffffu, p); }
  AGZ_DEV static bool any_warp(bool p) { return __any_sync(0xffffffffu, p); }
  AGZ_DEV static bool nz_warp(unsigned group_uniform) { return group_uniform != 0u; }   // "non-zero in some group of the warp"
  template <class T> AGZ_DEV static T shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
  template <class T> AGZ_DEV static T shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
  AGZ_DEV static unsigned reduce_or(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
  AGZ_DEV static int reduce_add(int v) { return __reduce_add_sync(0xffffffffu, v); }
  AGZ_DEV static unsigned reduce_max(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }
};
template <> struct G<16> {
  static constexpr int width = 16;
  AGZ_DEV static int lane() { return threadIdx.x & 15; }
  AGZ_DEV static int half() { return (threadIdx.x >> 4) & 1; }
  AGZ_DEV static void sync() { __syncwarp(); }
  AGZ_DEV static unsigned ballot(bool p) { return (__ballot_sync(0xffffffffu, p) >> (threadIdx.x & 16)) & 0xffffu; }
  AGZ_DEV static bool any(bool p) { return ballot(p) != 0u; }
  AGZ_DEV static bool any_warp(bool p) { return __any_sync(0xffffffffu, p); }
  AGZ_DEV static bool nz_warp(unsigned group_uniform) { return __any_sync(0xffffffffu, group_uniform != 0u); }
  template <class T> AGZ_DEV static T shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src, 16); }
  template <class T> AGZ_DEV static T shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m, 16); }
  // REDUX has no segmented form: one full-warp reduction per half over values that are neutral in the other half
  AGZ_DEV static unsigned reduce_or(unsigned v) {
    const bool h = (threadIdx.x & 16) != 0;
    const unsigned a = __reduce_or_sync(0xffffffffu, h ? 0u : v), b = __reduce_or_sync(0xffffffffu, h ? v : 0u);
    return h ? b : a;
  }
  AGZ_DEV static int reduce_add(int v) {
    const bool h = (threadIdx.x & 16) != 0;
    const int a = __reduce_add_sync(0xffffffffu, h ? 0 : v), b = __reduce_add_sync(0xffffffffu, h ? v : 0);
    return h ? b : a;
  }
  AGZ_DEV static unsigned reduce_max(unsigned v) {
    const bool h = (threadIdx.x & 16) != 0;
    const unsigned a = __reduce_max_sync(0xffffffffu, h ? 0u : v), b = __reduce_max_sync(0xffffffffu, h ? v : 0u);
    return h ? b : a;
  }
};
}  // namespace simt

#else  // ------------------------------------------------------------------ host emulation (tests only)
#define AGZ_CUDA 0
#define AGZ_DEV inline
#include <math.h>
#include <string.h>
#include <ucontext.h>

namespace simt {
struct EmuWarp {
  ucontext_t main_ctx;
  ucontext_t ctx[32];
  int cur;
  uint64_t slot[2][32];
  int tag[2][32];      // op kind of the collective each lane deposited (divergence check of the half-warp groups)
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
  w->tag[par][me] = 0;
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
  w->tag[par][me] = 0;
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
  w->tag[par][me] = 0;
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
  w->tag[par][me] = 0;
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
  w->tag[par][me] = 0;
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

// lane groups (see the CUDA half of this file): G<16> collectives are issued by all 32 fibers in lock step and return the result
// over the caller's own half.  Every collective records an op tag; a mismatch between the lanes of a rotation means the device
// code let the halves' control flow diverge around a collective, which the CUDA build would hang or miscompute on.
void emu_divergence(const char* what);
template <int W> struct G;
template <> struct G<32> {
  static constexpr int width = 32;
  static int lane() { return simt::lane(); }
  static int half() { return 0; }
  static void sync() { simt::sync(); }
  static unsigned ballot(bool p) { return simt::ballot(p); }
  static bool any(bool p) { return simt::any(p); }
  static bool any_warp(bool p) { return simt::any(p); }
  static bool nz_warp(unsigned group_uniform) { return group_uniform != 0u; }
  template <class T> static T shfl(T v, int src) { return simt::exchange_(v, src); }
  template <class T> static T shfl_xor(T v, int m) { return simt::exchange_(v, simt::lane() ^ m); }
  static unsigned reduce_or(unsigned v) { return simt::reduce_or(v); }
  static int reduce_add(int v) { return simt::reduce_add(v); }
  static unsigned reduce_max(unsigned v) { return simt::reduce_max(v); }
};
template <> struct G<16> {
  static constexpr int width = 16;
  static int lane() { return g_warp->cur & 15; }
  static int half() { return (g_warp->cur >> 4) & 1; }
  // all 32 lanes deposit (tag, value), rotate, and read the 16 values of their own half
  static void gather_(int tag, uint64_t mine, uint64_t (&out)[16]) {
    EmuWarp* w = g_warp;
    const int me = w->cur;
    const int par = w->ncoll[me]++ & 1;
    w->slot[par][me] = mine;
    w->tag[par][me] = tag;
    barrier();
    for (int i = 0; i < 32; ++i)
      if (w->tag[par][i] != tag) emu_divergence("the two halves of a duo warp reached different collectives");
    for (int i = 0; i < 16; ++i) out[i] = w->slot[par][(me & 16) | i];
  }
  static void sync() { uint64_t o[16]; gather_(1, 0, o); }
  static unsigned ballot(bool p) {
    uint64_t o[16];
    gather_(2, p ? 1 : 0, o);
    unsigned m = 0;
    for (int i = 0; i < 16; ++i) m |= (unsigned)(o[i] & 1) << i;
    return m;
  }
  static bool any(bool p) { return ballot(p) != 0u; }
  static bool any_warp(bool p) {
    EmuWarp* w = g_warp;
    const int me = w->cur;
    const int par = w->ncoll[me]++ & 1;
    w->slot[par][me] = p ? 1 : 0;
    w->tag[par][me] = 3;
    barrier();
    bool r = false;
    for (int i = 0; i < 32; ++i) {
      if (w->tag[par][i] != 3) emu_divergence("the two halves of a duo warp reached different collectives");
      r = r || (w->slot[par][i] & 1);
    }
    return r;
  }
  static bool nz_warp(unsigned group_uniform) { return any_warp(group_uniform != 0u); }
  template <class T> static T shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "");
    uint64_t raw = 0, o[16];
    memcpy(&raw, &v, sizeof(T));
    gather_(4, raw, o);
    T r;
    memcpy(&r, &o[src & 15], sizeof(T));
    return r;
  }
  template <class T> static T shfl_xor(T v, int m) { return shfl(v, (lane() ^ m) & 15); }
  static unsigned reduce_or(unsigned v) {
    uint64_t o[16];
    gather_(5, v, o);
    unsigned m = 0;
    for (int i = 0; i < 16; ++i) m |= (unsigned)o[i];
    return m;
  }
  static int reduce_add(int v) {
    uint64_t o[16];
    gather_(6, (uint64_t)(uint32_t)v, o);
    int m = 0;
    for (int i = 0; i < 16; ++i) m += (int)(uint32_t)o[i];
    return m;
  }
  static unsigned reduce_max(unsigned v) {
    uint64_t o[16];
    gather_(7, v, o);
    unsigned m = 0;
    for (int i = 0; i < 16; ++i) m = (unsigned)o[i] > m ? (unsigned)o[i] : m;
    return m;
  }
};
}  // namespace simt
#endif
