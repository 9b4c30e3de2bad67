// devrt.h -- thin device-runtime layer used by engine.cu: CUDA in the product build; in the CPU test
// build (-DAGZ_EMU, tests/emu) "device memory" is host memory and a kernel launch runs the functor on
// the fiber warp emulator of simt.h.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "simt.h"

#if AGZ_CUDA
namespace devrt {
typedef cudaStream_t stream_t;
inline const char* last_error_string(int rc) { return cudaGetErrorString((cudaError_t)rc); }
inline int dmalloc(void** p, size_t n) { return (int)cudaMalloc(p, n ? n : 1); }
inline void dfree(void* p) { if (p) cudaFree(p); }
inline int dmemset(void* p, int v, size_t n, stream_t s) { return (int)cudaMemsetAsync(p, v, n, s); }
inline int h2d(void* d, const void* h, size_t n, stream_t s) {
  int rc = (int)cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s);
  if (rc) return rc;
  return (int)cudaStreamSynchronize(s);
}
inline int d2h(void* h, const void* d, size_t n, stream_t s) {
  int rc = (int)cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s);
  if (rc) return rc;
  return (int)cudaStreamSynchronize(s);
}
inline int sync(stream_t s) { return (int)cudaStreamSynchronize(s); }

// minimum resident CTAs per SM the compiler must allow for an op (register cap = 65536 / (128 * v)); the tree kernels
// are latency bound with one warp per game, so occupancy is what buys throughput (ncu: 14 resident warps/SM at 128 regs)
template <class Op>
struct MinBlocks {
  static const int v = 4;
};

template <class Op>
struct TraceTag {
  static const int v = 9;
};

template <class Op>
__global__ void __launch_bounds__(128, MinBlocks<Op>::v) k_warps(const __grid_constant__ Op op, int n_warps, int smem_per_warp) {
  extern __shared__ __align__(16) char smem[];
  const int wib = threadIdx.x >> 5;
  const int w = blockIdx.x * 4 + wib;
  unsigned long long t0 = 0;
  if (op.v.trace) t0 = simt::gtimer();
  if (w < n_warps) op(w, smem + (size_t)wib * smem_per_warp);
  if (op.v.trace) {
    __syncthreads();
    if (simt::trace_cta()) simt::trace_rec(op.v.trace, TraceTag<Op>::v, t0);
  }
}

// When the tree kernels are meant to run underneath the persistent conv kernel (half-batch pipelining) they must ask for
// the same shared-memory carveout: an SM cannot host CTAs of kernels with different L1/shared splits at the same time.
// Otherwise they keep the default split: they like a large L1 (measured: +20 % select time with the max-shared carveout).
inline int& prefer_max_smem() {
  static int v = 0;
  return v;
}

template <class Op>
inline int launch_warps(const Op& op, int n_warps, int smem_per_warp, stream_t s) {
  if (n_warps <= 0) return 0;
  static int carveout = -2;
  const int want = prefer_max_smem() ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault;
  if (carveout != want) {
    cudaFuncSetAttribute(k_warps<Op>, cudaFuncAttributePreferredSharedMemoryCarveout, want);
    carveout = want;
  }
  k_warps<Op><<<(n_warps + 3) / 4, 128, 4 * smem_per_warp, s>>>(op, n_warps, smem_per_warp);
  return (int)cudaGetLastError();
}
}  // namespace devrt

#else
namespace devrt {
typedef int stream_t;
inline const char* last_error_string(int) { return "emulation error"; }
inline int dmalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 1; }
inline void dfree(void* p) { free(p); }
inline int dmemset(void* p, int v, size_t n, stream_t) { memset(p, v, n); return 0; }
inline int h2d(void* d, const void* h, size_t n, stream_t) { memcpy(d, h, n); return 0; }
inline int d2h(void* h, const void* d, size_t n, stream_t) { memcpy(h, d, n); return 0; }
inline int sync(stream_t) { return 0; }

template <class Op>
struct EmuCall {
  const Op* op;
  int w;
  char* smem;
  static void run(void* p) {
    EmuCall* c = (EmuCall*)p;
    (*c->op)(c->w, c->smem);
  }
};

template <class Op>
inline int launch_warps(const Op& op, int n_warps, int smem_per_warp, stream_t) {
  char* smem = (char*)calloc((size_t)smem_per_warp + 64, 1);
  for (int w = 0; w < n_warps; ++w) {
    EmuCall<Op> c{&op, w, smem};
    simt::emu_run_warp(&EmuCall<Op>::run, &c);
  }
  free(smem);
  return 0;
}
}  // namespace devrt
#endif
