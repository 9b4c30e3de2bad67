// nn_tc.cu -- the residual tower on Blackwell tensor cores (replaces the Flux Conv/BatchNorm/relu calls of
// src/neural_net.jl:16-21 and src/resnet.jl:26-32, which the reference sends to NNlib / cuDNN).
//
// A 3x3 convolution is an implicit GEMM  D[M, 256] = sum over 9 taps of A_tap[M, Cin] * W_tap[Cin, 256].
// Activations live in HBM as dense fp16 NHWC rows (row = b*N^2 + N*j + i, 256 channels; the stem input has 64).
// The A tile of a tap is ONE TMA im2col load (the TMA unit walks 128 consecutive output pixels across board rows
// and boards and zero-fills the halo), so M = B*N^2 exactly and no border rows exist.
//
// Kernels conv3x3_tc5_kernel / conv3x3_tc6_kernel (persistent, clusters of 2 CTAs = one TPC, 256 threads, warp-specialised):
//   warp 0   : TMA producer -- per (tap, 64-channel chunk): A im2col box 128 pixels x 64 ch, W box 128 cout x 64 ch (this
//              CTA's half of the weight tile), both SWIZZLE_128B, mbarrier ring of 6 (tc5) / 4 (tc6) stages of 32 KB
//   warp 1   : one elected lane of the leader CTA issues tcgen05.mma.cta_group::2.kind::f16  M=256, N=256, K=16 (4 per stage),
//              fp32 accumulators in the TMEM of both CTAs; tcgen05.commit (multicast) releases the stage / publishes the tile
//   warp 2   : allocates / frees the 512 TMEM columns (2 accumulator stages of 256 columns)
//   warps 4-7: epilogue -- software-pipelined tcgen05.ld 32 lanes x 32 columns, fused (conv bias + BatchNorm) scale/shift,
//              residual add, ReLU, fp16 pack, 32-byte global stores; overlaps the next tile's MMAs
// (Earlier variants -- per-tap 2-D boxes on zero-bordered boards, a shared-memory slab with row-shifted descriptor views, the
// first CTA-pair kernel -- are in the git history of round 1 and in profiles/r01_conv3x3_tc_ncu_full.md; they were removed.)
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "devrt.h"
#include "nn_state.h"
#include "ops.cuh"

namespace agz {

static const int BM = 128, BK = 64;
static const int A_BYTES = BM * BK * 2;
static const int CIN0 = 64;  // stem input channels: 17 planes zero-padded to one 64-channel K chunk

struct TCState {
  unsigned long long* trace;  // kernel timeline trace buffer or nullptr (nn_tc_set_trace)
  // options (agz_set_option "conv.*"; all read at launch time)
  int fuse_heads;             // conv.fuse_heads (default 1): head 1x1 convs in the last tower conv's epilogue, trunk not stored
  int conv5_stages;           // conv.stages: 6 (default) or 4 operand stages in conv3x3_tc5_kernel (experiment)
  int l2pf;                   // conv.l2_prefetch: L2 prefetch of the next tile in the conv producers (default 0)
  int pdl;                    // conv.pdl (default 1): tower convolutions use programmatic dependent launch
  int max_pairs;              // conv.max_pairs: cap on the CTA pairs of the persistent conv kernels (0 = all SMs)
  int res_tma;                // conv.res_tma (default 1): residual convs use conv3x3_tc6_kernel (shortcut tile by TMA)
  int split;                  // conv.precision = 2: split-precision activations / weights (hi + lo fp16 pairs), see conv3x3_tc5_kernel
  __half* sact[3];            // split precision: [rows_alloc][512] (hi | lo)
  std::vector<__half*> sw;    // split precision: per conv layer [9*256][2*cin_pad] (hi | lo)
  CUtensorMap tm5_sact[3];
  std::vector<CUtensorMap> tm_sw;
  bool split_weights_ready;
  float4* head_pre;           // [rows_alloc] (value plane, policy plane 0, policy plane 1, 0) written by the fused-heads epilogue
  int N, PP, C, T, max_batch; // PP = N*N rows per board
  CUtensorMap tm5_in64, tm5_act[3];   // im2col maps over the stem input / the three rotating activation buffers
  int groups;                 // 2 when the batch can be split into two half batches (even max_batch)
  long long grp_rows;         // rows per group
  CUtensorMap tm5g_in64[2], tm5g_act[2][3];
  long long rows_alloc;       // rows allocated per activation buffer (whole boards + spare boards for the last 256-row tile)
  __half* in64;               // [rows_alloc][64]
  __half* act[3];             // [rows_alloc][256]
  std::vector<__half*> w;     // per conv layer: [9*256][cin] fp16, tap-major, K (cin) contiguous
  CUtensorMap tm_act[3];      // plain 2-D maps (128 rows x 64 channel boxes): the shortcut tile of conv3x3_tc6_kernel
  std::vector<CUtensorMap> tm_w;   // per conv layer: boxes of 128 output channels x 64 input channels
  int num_sms;
  bool attr_set;
  float* stage;               // device staging copy of the raw Flux parameter list (base chain)
  size_t stage_cap;
};

// ------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Split form for a software-pipelined epilogue: issue the load of the next 32 accumulator columns, work on the current ones, and
// wait before the first use.  tcgen05.wait::ld covers every outstanding load of the thread; the registers are in/out operands of
// the wait so that no use of them can be scheduled above it.
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                 "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// 32-byte store (STG.256, sm_100): one full sector per thread instead of two half-sector 16-byte stores
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
               "r"(b.z), "r"(b.w)
               : "memory");
}

__device__ __forceinline__ uint4 ld_nc_v4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 64 fp16 (128 B), 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;             // leading byte offset (unused for swizzled K-major), 16 B units
  d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}

struct ConvArgs {
  const float* scale;   // [256] folded BatchNorm scale
  const float* shift;   // [256] folded BatchNorm shift (+ scale * conv bias)
  const __half* res;    // residual input [rows][256] or nullptr
  __half* out;          // [rows][256]
  int n_tiles;          // ceil(rows_valid / 128)
  long long rows_valid; // rows of real boards (B * PP)
  int kchunks;          // K chunks of 64 channels per tap (Cin / 64; three times that in split precision)
  int16_t a_off[12];    // activation / weight channel offset of K chunk kc: identity (64*kc) in fp16 mode; in split precision the
  int16_t w_off[12];    // three products a_hi*w_hi, a_lo*w_hi, a_hi*w_lo are laid out as 3*Cin/64 chunks
  int stem;             // first convolution (trace tag only)
  int N, PP;
  int relu;
  unsigned long long* trace;   // kernel timeline trace (simt.h) or nullptr
  int l2pf;                    // prefetch the next tile's activation (and shortcut) rows into L2 (AGZ_CONV_L2PF)
  // last convolution of the tower (conv3x3_tc6_kernel only): the 1x1 convolutions + BatchNorm + relu of the value and policy
  // heads (neural_net.jl:23-24,28-29) are evaluated in the epilogue from the fp32 trunk values and the trunk is not stored
  const float* head_vw;        // [256] value 1x1 conv weights, or nullptr
  const float* head_pw;        // [2][256] policy 1x1 conv weights
  const float* head_aff;       // [6] folded conv bias + BatchNorm: v scale, v shift, p0 scale, p0 shift, p1 scale, p1 shift
  float4* head_out;            // [rows] (value plane, policy plane 0, policy plane 1, 0)
};

// ------------------------------------------------------------------------------------------- CTA pair
// cta_group::2: two CTAs of a cluster (one TPC) work on one 256-row x 256-channel tile.  Each CTA stages only ITS
// 128 activation rows and HALF of the weight tile (128 of the 256 output channels); one tcgen05.mma issued by the
// leader CTA consumes both halves, so per-SM shared-memory operand traffic per FLOP drops by a third versus
// cta_group::1 (A 4 KB + B 4 KB per 128-cycle MMA instead of 4 + 8).  TMA completions of both CTAs land on the
// leader's mbarrier; tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs.
static const int V3_STAGES = 6;   // operand stages of conv3x3_tc5_kernel<6>
static const int V3_STAGE_BYTES = A_BYTES + A_BYTES;   // A 128x64 + B-half 128x64
static const size_t CONV3_SMEM = (size_t)V3_STAGES * V3_STAGE_BYTES + 2 * 256 * sizeof(float) + 256 + 1024;
static const uint32_t IDESC_F16_M256_N256 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t a = smem_u32(bar);
  uint32_t spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (!done && ++spins > 40000000u) asm volatile("trap;");   // a broken pipeline aborts the launch instead of hanging the GPU
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* tm, uint32_t leader_bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// ------------------------------------------------------------------------------------------- v5: CTA pair + TMA im2col
// Dense NHWC activations (row = b*N^2 + N*j + i, no border rows): the A tile of tap (kj, ki) is ONE TMA im2col load --
// 128 consecutive output pixels starting at (w, h, n) = (i0 - 1, j0 - 1, b0) with filter offsets {ki, kj}; the TMA unit
// walks pixels across rows and boards and zero-fills the halo (pixelBoxLowerCorner = upperCorner = -1, i.e. pad 1,
// 3x3).  No MMA work is spent on border rows: M = B*N^2 exactly (v1-v4 issue (N+1)^2/N^2 = 1.235x the MMAs on 9x9).
__device__ __forceinline__ void tma_load_im2col_2sm(void* dst, const CUtensorMap* tm, uint32_t leader_bar, int c, int w, int h, int n,
                                                    uint16_t ow, uint16_t oh) {
  asm volatile("cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(leader_bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh) : "memory");
}

// L2 prefetch of a future tile's operands (no shared memory involved): the demand loads then hit L2 instead of HBM
__device__ __forceinline__ void tma_prefetch_im2col(const CUtensorMap* tm, int c, int w, int h, int n, uint16_t ow, uint16_t oh) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.im2col [%0, {%1, %2, %3, %4}], {%5, %6};"
               ::"l"(tm), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}

// SPLIT = split precision (option conv.precision = 2): activations and weights are pairs of fp16 numbers x = hi + lo (rows of 512
// channels: hi | lo), the three significant products are three times the K chunks of the same MMA loop (the producer's chunk
// tables), and the epilogue writes hi = fp16(y), lo = fp16(y - hi): ~21 significant bits instead of 11 at a third of the speed.
template <int NSTAGES, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
conv3x3_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const ConvArgs a) {
  const unsigned long long trace_t0 = a.trace ? simt::gtimer() : 0ULL;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next convolution may start its prologue as SMs free up
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  float* s_scale = reinterpret_cast<float*>(smem + (size_t)NSTAGES * V3_STAGE_BYTES);
  float* s_shift = s_scale + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 256);
  uint64_t* full = bars;
  uint64_t* empty = bars + NSTAGES;
  uint64_t* tfull = bars + 2 * NSTAGES;
  uint64_t* tempty = bars + 2 * NSTAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_ptiles = (a.n_tiles + 1) >> 1;
  s_scale[2 * threadIdx.x] = a.scale[threadIdx.x];       // interleaved (scale, shift) pairs: one 16-byte read per two channels
  s_scale[2 * threadIdx.x + 1] = a.shift[threadIdx.x];
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  // programmatic dependent launch: everything above overlapped the previous convolution's tail; its output (this kernel's
  // input / shortcut) and the buffer this kernel overwrites are only touched after the wait (no-op for a normal launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int iters = 9 * a.kchunks;
  const int N2 = a.N * a.N;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < n_ptiles; t += n_pairs) {
        const int m0 = t * 256 + (int)rank * 128;
        const int b0 = m0 / N2, rem = m0 - b0 * N2, j0 = rem / a.N, i0 = rem - j0 * a.N;
        if (a.l2pf && t + n_pairs < n_ptiles) {   // this pair's next tile: pull its rows (centre tap = the rows themselves) into L2 now
          const int m1 = (t + n_pairs) * 256 + (int)rank * 128;
          const int b1 = m1 / N2, rem1 = m1 - b1 * N2, j1 = rem1 / a.N, i1 = rem1 - j1 * a.N;
          for (int kc = 0; kc < a.kchunks; ++kc) tma_prefetch_im2col(&tmA, a.a_off[kc], i1 - 1, j1 - 1, b1, 1, 1);
        }
        for (int tap = 0; tap < 9; ++tap) {
          const uint16_t oh = (uint16_t)(tap / 3), ow = (uint16_t)(tap % 3);
          for (int kc = 0; kc < a.kchunks; ++kc) {
            mbar_wait_guard(&empty[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * V3_STAGE_BYTES);
            const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
            uint8_t* sa = tiles + (size_t)stage * V3_STAGE_BYTES;
            tma_load_im2col_2sm(sa, &tmA, lbar, a.a_off[kc], i0 - 1, j0 - 1, b0, ow, oh);
            tma_load_2d_2sm(sa + A_BYTES, &tmW, lbar, a.w_off[kc], tap * 256 + (int)rank * 128);
            if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int titer = 0;
      for (int t = pair; t < n_ptiles; t += n_pairs, ++titer) {
        const int as = titer & 1;
        const uint32_t aphase = (titer >> 1) & 1;
        mbar_wait_guard(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 256;
        for (int it = 0; it < iters; ++it) {
          mbar_wait_guard(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * V3_STAGE_BYTES);
          const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC_F16_M256_N256, (it > 0 || k > 0) ? 1u : 0u);
          tc_commit_2sm(&empty[stage]);
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_2sm(&tfull[as]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int titer = 0;
    for (int t = pair; t < n_ptiles; t += n_pairs, ++titer) {
      const int as = titer & 1;
      const uint32_t aphase = (titer >> 1) & 1;
      const long long row = (long long)t * 256 + rank * 128 + q * 32 + lane;
      const bool valid = row < a.rows_valid;
      constexpr int ROWC = SPLIT ? 512 : 256;   // channels per stored row
      __half* orow = a.out + row * ROWC;
      const bool addres = a.res != nullptr && valid;
      // residual rows do not depend on the MMAs: chunks 0-1 are fetched before waiting for the accumulator and chunk
      // cc+2 while chunk cc is processed (2 register buffers; keeps the CTA small enough for tree CTAs to co-reside)
      uint4 rv[2][4];
      const uint4* rrow = reinterpret_cast<const uint4*>(a.res + row * ROWC);
      if (addres && !SPLIT) {
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
          for (int j = 0; j < 4; ++j) rv[c2][j] = ld_nc_v4(rrow + c2 * 4 + j);
      }
      mbar_wait_guard(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)as * 256 + ((uint32_t)(q * 32) << 16);
      uint32_t vv[2][32];   // the load of columns cc+1 is in flight while columns cc are scaled, packed and stored
      tc_ld32_issue(tacc, vv[0]);
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t (&v)[32] = vv[cc & 1];
        tc_ld_wait(v);
        if (cc < 7) tc_ld32_issue(tacc + (uint32_t)(cc + 1) * 32, vv[(cc + 1) & 1]);
        if (valid) {
          uint4 o[4];
          uint32_t* ow = reinterpret_cast<uint32_t*>(o);
          if (SPLIT) {
            uint4 ol[4], rh4[4], rl4[4];
            uint32_t* owl = reinterpret_cast<uint32_t*>(ol);
            if (addres) {
#pragma unroll
              for (int j = 0; j < 4; ++j) { rh4[j] = ld_nc_v4(rrow + cc * 4 + j); rl4[j] = ld_nc_v4(rrow + 32 + cc * 4 + j); }
            }
            const __half2* rhh = reinterpret_cast<const __half2*>(rh4);
            const __half2* rll = reinterpret_cast<const __half2*>(rl4);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c0 = cc * 32 + 2 * j;
              const float4 ss = *reinterpret_cast<const float4*>(s_scale + 2 * c0);
              float y0 = fmaf(__uint_as_float(v[2 * j]), ss.x, ss.y);
              float y1 = fmaf(__uint_as_float(v[2 * j + 1]), ss.z, ss.w);
              if (addres) {
                const float2 a2 = __half22float2(rhh[j]), b2 = __half22float2(rll[j]);
                y0 += a2.x + b2.x;
                y1 += a2.y + b2.y;
              }
              if (a.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
              const __half2 h = __floats2half2_rn(y0, y1);
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
              ow[j] = *reinterpret_cast<const uint32_t*>(&h);
              owl[j] = *reinterpret_cast<const uint32_t*>(&l);
            }
            st_global_256(orow + cc * 32, o[0], o[1]);
            st_global_256(orow + cc * 32 + 16, o[2], o[3]);
            st_global_256(orow + 256 + cc * 32, ol[0], ol[1]);
            st_global_256(orow + 256 + cc * 32 + 16, ol[2], ol[3]);
          } else {
          const __half2* rh = reinterpret_cast<const __half2*>(rv[cc & 1]);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c0 = cc * 32 + 2 * j;
            const float4 ss = *reinterpret_cast<const float4*>(s_scale + 2 * c0);
            float y0 = fmaf(__uint_as_float(v[2 * j]), ss.x, ss.y);
            float y1 = fmaf(__uint_as_float(v[2 * j + 1]), ss.z, ss.w);
            if (addres) {
              float2 rr = __half22float2(rh[j]);
              y0 += rr.x;
              y1 += rr.y;
            }
            if (a.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
            __half2 h = __floats2half2_rn(y0, y1);
            ow[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          if (addres && cc < 6) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rv[cc & 1][j] = ld_nc_v4(rrow + (cc + 2) * 4 + j);
          }
          st_global_256(orow + cc * 32, o[0], o[1]);
          st_global_256(orow + cc * 32 + 16, o[2], o[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        const uint32_t lb = mapa_u32(smem_u32(&tempty[as]), 0);
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(lb) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
  if (a.trace && simt::trace_cta()) simt::trace_rec(a.trace, a.stem ? 4 : 5, trace_t0);
}


// ------------------------------------------------------------------------------------------- v6: v5 + residual tile by TMA
// For the second convolution of a residual block the shortcut rows (128 x 256 fp16 = 64 KB per CTA and tile) are
// fetched by TMA into shared memory half way through the tile's K loop -- long before the epilogue needs them -- instead
// of by per-thread global loads inside the epilogue (ncu on v5: the residual-carrying launches were 19 % slower, the
// epilogue waiting on 64-byte global loads).  4 operand stages (128 KB) + 64 KB residual tile.
static const int V6_STAGES = 4;
static const int V6_RES_BYTES = BM * 256 * 2;
static const size_t CONV6_SMEM = (size_t)V6_STAGES * V3_STAGE_BYTES + V6_RES_BYTES + 2 * 256 * sizeof(float) + 256 * sizeof(float4) + 256 + 1024;

template <bool HEADS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
conv3x3_tc6_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmR,
                   const ConvArgs a, const int res_row0) {
  const unsigned long long trace_t0 = a.trace ? simt::gtimer() : 0ULL;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next convolution may start its prologue as SMs free up
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;
  uint8_t* resbuf = smem + (size_t)V6_STAGES * V3_STAGE_BYTES;      // 4 boxes of 128 rows x 64 channels, SWIZZLE_128B
  float* s_scale = reinterpret_cast<float*>(resbuf + V6_RES_BYTES);
  float* s_shift = s_scale + 256;
  float4* s_hw = reinterpret_cast<float4*>(s_shift + 256);   // per channel: (value w, policy-0 w, policy-1 w, 0)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_hw + 256);
  uint64_t* full = bars;
  uint64_t* empty = bars + V6_STAGES;
  uint64_t* tfull = bars + 2 * V6_STAGES;
  uint64_t* tempty = bars + 2 * V6_STAGES + 2;
  uint64_t* rfull = bars + 2 * V6_STAGES + 4;    // local: residual tile landed
  uint64_t* rempty = bars + 2 * V6_STAGES + 5;   // local: 4 epilogue warps are done with it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * V6_STAGES + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_ptiles = (a.n_tiles + 1) >> 1;
  s_scale[2 * threadIdx.x] = a.scale[threadIdx.x];       // interleaved (scale, shift) pairs
  s_scale[2 * threadIdx.x + 1] = a.shift[threadIdx.x];
  constexpr bool heads = HEADS;
  if (heads) s_hw[threadIdx.x] = make_float4(a.head_vw[threadIdx.x], a.head_pw[threadIdx.x], a.head_pw[256 + threadIdx.x], 0.f);
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < V6_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
    mbar_init(rfull, 1);
    mbar_init(rempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  // programmatic dependent launch: everything above overlapped the previous convolution's tail; its output (this kernel's
  // input / shortcut) and the buffer this kernel overwrites are only touched after the wait (no-op for a normal launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int iters = 9 * a.kchunks;
  const int N2 = a.N * a.N;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, rphase = 0;
      for (int t = pair; t < n_ptiles; t += n_pairs) {
        const int m0 = t * 256 + (int)rank * 128;
        const int b0 = m0 / N2, rem = m0 - b0 * N2, j0 = rem / a.N, i0 = rem - j0 * a.N;
        if (a.l2pf && t + n_pairs < n_ptiles) {   // this pair's next tile: its rows and its shortcut rows into L2 now
          const int m1 = (t + n_pairs) * 256 + (int)rank * 128;
          const int b1 = m1 / N2, rem1 = m1 - b1 * N2, j1 = rem1 / a.N, i1 = rem1 - j1 * a.N;
          for (int kc = 0; kc < a.kchunks; ++kc) tma_prefetch_im2col(&tmA, a.a_off[kc], i1 - 1, j1 - 1, b1, 1, 1);
          for (int bx = 0; bx < 4; ++bx) tma_prefetch_2d(&tmR, bx * BK, res_row0 + m1);
        }
        int it = 0;
        for (int tap = 0; tap < 9; ++tap) {
          const uint16_t oh = (uint16_t)(tap / 3), ow = (uint16_t)(tap % 3);
          for (int kc = 0; kc < a.kchunks; ++kc, ++it) {
            mbar_wait_guard(&empty[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * V3_STAGE_BYTES);
            const uint32_t lbar = mapa_u32(smem_u32(&full[stage]), 0);
            uint8_t* sa = tiles + (size_t)stage * V3_STAGE_BYTES;
            tma_load_im2col_2sm(sa, &tmA, lbar, a.a_off[kc], i0 - 1, j0 - 1, b0, ow, oh);
            tma_load_2d_2sm(sa + A_BYTES, &tmW, lbar, a.w_off[kc], tap * 256 + (int)rank * 128);
            if (++stage == V6_STAGES) { stage = 0; phase ^= 1; }
            if (it == iters / 2) {  // the previous tile's epilogue is long done by now: fetch this tile's shortcut rows
              mbar_wait_guard(rempty, rphase ^ 1);
              mbar_expect_tx(rfull, V6_RES_BYTES);
#pragma unroll
              for (int bx = 0; bx < 4; ++bx) tma_load_2d(resbuf + (size_t)bx * A_BYTES, &tmR, rfull, bx * BK, res_row0 + m0);
              rphase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int titer = 0;
      for (int t = pair; t < n_ptiles; t += n_pairs, ++titer) {
        const int as = titer & 1;
        const uint32_t aphase = (titer >> 1) & 1;
        mbar_wait_guard(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 256;
        for (int it = 0; it < iters; ++it) {
          mbar_wait_guard(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)stage * V3_STAGE_BYTES);
          const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC_F16_M256_N256, (it > 0 || k > 0) ? 1u : 0u);
          tc_commit_2sm(&empty[stage]);
          if (++stage == V6_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_2sm(&tfull[as]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int r = q * 32 + lane;   // row of this thread inside the CTA's 128-row tile
    int titer = 0;
    for (int t = pair; t < n_ptiles; t += n_pairs, ++titer) {
      const int as = titer & 1;
      const uint32_t aphase = (titer >> 1) & 1;
      const long long row = (long long)t * 256 + rank * 128 + r;
      const bool valid = row < a.rows_valid;
      __half* orow = a.out + row * 256;
      mbar_wait_guard(&tfull[as], aphase);
      mbar_wait_guard(rfull, (uint32_t)(titer & 1));
      tc_fence_after();
      float h0 = 0.f, h1 = 0.f, h2 = 0.f;   // this row's value / policy 1x1 convolutions (fixed channel order: batch invariant)
      const uint32_t tacc = tmem_base + (uint32_t)as * 256 + ((uint32_t)(q * 32) << 16);
      uint32_t vv[2][32];   // without the heads: the load of columns cc+1 is in flight while columns cc are processed
      if (!heads) tc_ld32_issue(tacc, vv[0]);
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t (&v)[32] = vv[heads ? 0 : (cc & 1)];
        if (heads) {
          tc_ld32(tacc + (uint32_t)cc * 32, v);
        } else {
          tc_ld_wait(v);
          if (cc < 7) tc_ld32_issue(tacc + (uint32_t)(cc + 1) * 32, vv[(cc + 1) & 1]);
        }
        // shortcut values of columns cc*32 .. +31: box cc/2, 16-byte chunks (cc%2)*4 + j, swizzled with the row
        uint4 rv[4];
        const uint8_t* rb = resbuf + (size_t)(cc >> 1) * A_BYTES + (size_t)r * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) rv[j] = *reinterpret_cast<const uint4*>(rb + ((((cc & 1) * 4 + j) ^ (r & 7)) << 4));
        if (valid) {
          uint4 o[4];
          uint32_t* ow = reinterpret_cast<uint32_t*>(o);
          const __half2* rh = reinterpret_cast<const __half2*>(rv);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c0 = cc * 32 + 2 * j;
            float2 rr = __half22float2(rh[j]);
            const float4 ss = *reinterpret_cast<const float4*>(s_scale + 2 * c0);
            float y0 = fmaf(__uint_as_float(v[2 * j]), ss.x, ss.y) + rr.x;
            float y1 = fmaf(__uint_as_float(v[2 * j + 1]), ss.z, ss.w) + rr.y;
            if (a.relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
            if (heads) {
              const float4 w0 = s_hw[c0], w1 = s_hw[c0 + 1];
              h0 = fmaf(y0, w0.x, h0); h1 = fmaf(y0, w0.y, h1); h2 = fmaf(y0, w0.z, h2);
              h0 = fmaf(y1, w1.x, h0); h1 = fmaf(y1, w1.y, h1); h2 = fmaf(y1, w1.z, h2);
            } else {
              __half2 h = __floats2half2_rn(y0, y1);
              ow[j] = *reinterpret_cast<uint32_t*>(&h);
            }
          }
          if (!heads) {
            st_global_256(orow + cc * 32, o[0], o[1]);
            st_global_256(orow + cc * 32 + 16, o[2], o[3]);
          }
        }
      }
      if (heads && valid) {
        const float* af = a.head_aff;
        a.head_out[row] = make_float4(fmaxf(fmaf(h0, af[0], af[1]), 0.f), fmaxf(fmaf(h1, af[2], af[3]), 0.f), fmaxf(fmaf(h2, af[4], af[5]), 0.f), 0.f);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(rempty);
        const uint32_t lb = mapa_u32(smem_u32(&tempty[as]), 0);
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(lb) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
  if (a.trace && simt::trace_cta()) simt::trace_rec(a.trace, 6, trace_t0);
}

// ------------------------------------------------------------------------------------------- heads (fp16 trunk)
// neural_net.jl:23-30.  HPB positions per CTA so that the dense-layer weights (83 KB + 53 KB on 9x9, 370 KB + 1 MB on
// 19x19) are read from L2 once per 8 positions instead of once per position.
static const int HPB = 8;
static size_t heads_smem(int N2, int A) { return (size_t)HPB * (3 * N2 + 256 + A + (A <= 128 ? (256 / A) * A : 0)) * sizeof(float); }

__global__ void __launch_bounds__(256) heads_tc_kernel(const __half* __restrict__ trunk, const float* __restrict__ vw,
                                                       const float* __restrict__ pw, const float* __restrict__ aff,
                                                       const float* __restrict__ D1W, const float* __restrict__ D1b,
                                                       const float* __restrict__ D2W, const float* __restrict__ D2b,
                                                       const float* __restrict__ PW, const float* __restrict__ Pb, float* __restrict__ pi,
                                                       float* __restrict__ v, int B, int N, int A,
                                                       unsigned long long* trace, const float4* __restrict__ pre, float* __restrict__ raw, int split) {
  const unsigned long long trace_t0 = trace ? simt::gtimer() : 0ULL;
  extern __shared__ float sm[];
  const int N2 = N * N;
  float* vf = sm;                       // [HPB][N2]
  float* pf = vf + HPB * N2;            // [HPB][2*N2], index p + N2*c
  float* hid = pf + HPB * 2 * N2;       // [HPB][256]
  float* lg = hid + HPB * 256;          // [HPB][A]
  const int b0 = blockIdx.x * HPB, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = min(HPB, B - b0);
  if (pre != nullptr) {
    // the last tower convolution already evaluated the 1x1 convolutions + BatchNorm + relu in its epilogue (dense layout)
    for (int it = tid; it < nb * N2; it += 256) {
      const int pb = it / N2, p = it - pb * N2;
      const float4 h = pre[(size_t)(b0 + pb) * N2 + p];
      vf[pb * N2 + p] = h.x;
      pf[pb * 2 * N2 + p] = h.y;
      pf[pb * 2 * N2 + N2 + p] = h.z;
    }
  } else {
  // 1x1 convolutions + BatchNorm + relu: a warp per (position, point), 8 channels per lane (one 16-byte load)
  float wv[8], wp0[8], wp1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    wv[j] = vw[lane * 8 + j];
    wp0[j] = pw[lane * 8 + j];
    wp1[j] = pw[256 + lane * 8 + j];
  }
  for (int it = warp; it < nb * N2; it += 8) {
    const int pb = it / N2, p = it - pb * N2;
    const __half* trow = trunk + ((size_t)(b0 + pb) * N2 + p) * (split ? 512 : 256) + lane * 8;
    const uint4 rawh = *reinterpret_cast<const uint4*>(trow);
    const __half2* h = reinterpret_cast<const __half2*>(&rawh);
    uint4 rawl = make_uint4(0u, 0u, 0u, 0u);
    if (split) rawl = *reinterpret_cast<const uint4*>(trow + 256);   // split precision: trunk value = hi + lo
    const __half2* hl = reinterpret_cast<const __half2*>(&rawl);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 x = __half22float2(h[j]);
      const float2 xl = __half22float2(hl[j]);
      x.x += xl.x;
      x.y += xl.y;
      a0 = fmaf(wv[2 * j], x.x, a0); a0 = fmaf(wv[2 * j + 1], x.y, a0);
      a1 = fmaf(wp0[2 * j], x.x, a1); a1 = fmaf(wp0[2 * j + 1], x.y, a1);
      a2 = fmaf(wp1[2 * j], x.x, a2); a2 = fmaf(wp1[2 * j + 1], x.y, a2);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, off);
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
      a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    }
    if (lane == 0) {
      vf[pb * N2 + p] = fmaxf(a0 * aff[0] + aff[1], 0.f);
      pf[pb * 2 * N2 + p] = fmaxf(a1 * aff[2] + aff[3], 0.f);
      pf[pb * 2 * N2 + N2 + p] = fmaxf(a2 * aff[4] + aff[5], 0.f);
    }
  }
  }
  __syncthreads();
  {  // Dense(N2 -> 256, relu): thread o, weights streamed once for all HPB positions
    float acc[HPB];
    const float bias = D1b[tid];
#pragma unroll
    for (int pb = 0; pb < HPB; ++pb) acc[pb] = bias;
#pragma unroll 4
    for (int i = 0; i < N2; ++i) {   // 4 independent L2 loads in flight per thread
      const float w = D1W[tid + 256 * i];
#pragma unroll
      for (int pb = 0; pb < HPB; ++pb) acc[pb] = fmaf(w, vf[pb * N2 + i], acc[pb]);
    }
#pragma unroll
    for (int pb = 0; pb < HPB; ++pb) hid[pb * 256 + tid] = fmaxf(acc[pb], 0.f);
  }
  __syncthreads();
  if (warp < nb) {  // Dense(256 -> 1, tanh): one warp per position
    float acc = 0.f;
    for (int o = lane; o < 256; o += 32) acc = fmaf(D2W[o], hid[warp * 256 + o], acc);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      v[b0 + warp] = tanhf(acc + D2b[0]);
      if (raw) raw[(size_t)(b0 + warp) * (A + 1) + A] = acc + D2b[0];   // debug: the value before tanh
    }
  }
  // Dense(2*N2 -> A).  Small boards (A <= 128): the K range is split over S = 256 / A thread groups so that all 256
  // threads stream weights (partials are summed through shared memory in a fixed order); otherwise thread a, a + 256.
  const int S = A <= 128 ? 256 / A : 1;
  if (S > 1) {
    float* part = lg + HPB * A;                 // [S][HPB][A]
    const int slice = tid / A, a0i = tid - slice * A;
    if (slice < S) {
      const int chunk = (2 * N2 + S - 1) / S, i0 = slice * chunk, i1 = min(2 * N2, i0 + chunk);
      float acc[HPB];
#pragma unroll
      for (int pb = 0; pb < HPB; ++pb) acc[pb] = 0.f;
#pragma unroll 6
      for (int i = i0; i < i1; ++i) {
        const float w = PW[a0i + (size_t)A * i];
#pragma unroll
        for (int pb = 0; pb < HPB; ++pb) acc[pb] = fmaf(w, pf[pb * 2 * N2 + i], acc[pb]);
      }
#pragma unroll
      for (int pb = 0; pb < HPB; ++pb) part[(slice * HPB + pb) * A + a0i] = acc[pb];
    }
    __syncthreads();
    for (int idx = tid; idx < HPB * A; idx += 256) {
      const int a1 = idx % A;
      float acc = Pb[a1];
      for (int sl = 0; sl < S; ++sl) acc += part[sl * HPB * A + idx];
      lg[idx] = acc;
    }
  } else {
    for (int a0i = tid; a0i < A; a0i += 256) {
      float acc[HPB];
      const float bias = Pb[a0i];
#pragma unroll
      for (int pb = 0; pb < HPB; ++pb) acc[pb] = bias;
#pragma unroll 6
      for (int i = 0; i < 2 * N2; ++i) {
        const float w = PW[a0i + (size_t)A * i];
#pragma unroll
        for (int pb = 0; pb < HPB; ++pb) acc[pb] = fmaf(w, pf[pb * 2 * N2 + i], acc[pb]);
      }
#pragma unroll
      for (int pb = 0; pb < HPB; ++pb) lg[pb * A + a0i] = acc[pb];
    }
  }
  __syncthreads();
  if (warp < nb) {  // softmax over all A actions (no legality masking, as the reference): one warp per position
    const float* l = lg + warp * A;
    float mx = -INFINITY;
    for (int a0i = lane; a0i < A; a0i += 32) mx = fmaxf(mx, l[a0i]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
    for (int a0i = lane; a0i < A; a0i += 32) sum += expf(l[a0i] - mx);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    float* out = pi + (size_t)(b0 + warp) * A;
    for (int a0i = lane; a0i < A; a0i += 32) out[a0i] = expf(l[a0i] - mx) / sum;
    if (raw)   // debug: the logits before softmax
      for (int a0i = lane; a0i < A; a0i += 32) raw[(size_t)(b0 + warp) * (A + 1) + a0i] = l[a0i];
  }
  if (trace) {
    __syncthreads();
    if (simt::trace_cta()) simt::trace_rec(trace, 7, trace_t0);
  }
}

// ------------------------------------------------------------------------------------------- feature kernels
// get_feats (features.jl:3-26) written straight into the stem's input rows: channels 0..15 stone planes,
// 16 = colour (+-1), 17..23 zero (24..63 stay zero from initialisation).
template <int KA>
struct LeafFeaturesTCOp {
  Cfg c;
  View v;
  __half* in64;
  int row0;  // first batch row handled by this launch
  __device__ void operator()(int wi, char* smem) const {
    const int b = row0 + wi;
    const int g = b / c.pmax, k = b % c.pmax;
    Warp<KA> w(c, v, g, smem);
    if (k >= w.st.nleaf) return;
    const int lane = threadIdx.x & 31;
    uint32_t planes[KA];
#pragma unroll
    for (int q = 0; q < KA; ++q) planes[q] = 0;
    const int tp = w.gather_features(v.leaf_node[b], [&](int hb, int p, uint32_t mine, uint32_t theirs) {
#pragma unroll
      for (int q = 0; q < KA; ++q)
        if (q == (p >> 5)) planes[q] |= (mine << (2 * hb)) | (theirs << (2 * hb + 1));
    });
    const __half one = __float2half(1.f), zero = __float2half(0.f), tph = __float2half((float)tp);
#pragma unroll
    for (int q = 0; q < KA; ++q) {
      const int p = q * 32 + lane;
      if (p < c.N2) {
        __half* row = in64 + ((size_t)b * c.N2 + p) * CIN0;
        __align__(16) __half h[24];
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) h[ch] = ((planes[q] >> ch) & 1u) ? one : zero;
        h[16] = tph;
#pragma unroll
        for (int ch = 17; ch < 24; ++ch) h[ch] = zero;
#pragma unroll
        for (int j = 0; j < 3; ++j) reinterpret_cast<uint4*>(row)[j] = reinterpret_cast<const uint4*>(h)[j];
      }
    }
  }
};

}  // namespace agz
namespace devrt {
template <int KA> struct TraceTag<agz::LeafFeaturesTCOp<KA>> { static const int v = 3; };
}
namespace agz {

__global__ void host_features_tc_kernel(const int8_t* __restrict__ bh, const int8_t* __restrict__ tp, __half* __restrict__ in64, int B, int N) {
  const int N2 = N * N;
  const size_t total = (size_t)B * N2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx % N2), b = (int)(idx / N2);
    const int t = tp[b];
    __half* row = in64 + ((size_t)b * N2 + p) * CIN0;
    __align__(16) __half h[24];
#pragma unroll
    for (int kq = 0; kq < 8; ++kq) {
      const int s = bh[((size_t)b * 8 + kq) * N2 + p];
      h[2 * kq] = __float2half(s == t ? 1.f : 0.f);
      h[2 * kq + 1] = __float2half(s == -t ? 1.f : 0.f);
    }
    h[16] = __float2half((float)t);
#pragma unroll
    for (int ch = 17; ch < 24; ++ch) h[ch] = __float2half(0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) reinterpret_cast<uint4*>(row)[j] = reinterpret_cast<const uint4*>(h)[j];
  }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col() {
  static EncodeIm2colFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeIm2colFn)p;
  }
  return fn;
}

// NHWC fp16 tensor (C, W = N, H = N, boards) for a 3x3 / pad 1 convolution: 128 pixels x 64 channels per load
static int make_map_im2col(CUtensorMap* tm, void* base, int N, long long boards, int C) {
  EncodeIm2colFn enc = get_encode_im2col();
  if (!enc) return 1;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)boards};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * N, (cuuint64_t)C * 2 * N * N};
  int lower[2] = {-1, -1}, upper[2] = {-1, -1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, lower, upper, (cuuint32_t)BK, (cuuint32_t)BM, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 2;
  // driver quirk mirrored from CUTLASS (copy_traits_sm90_im2col.hpp): small tensors need bit 21 of word 1 cleared on drivers <= 13.1
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && (unsigned long long)C * 2ull * N * N * (unsigned long long)boards < 131072ull)
    reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  return 0;
}

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp16 tensor [rows][cols] row-major, box = box_rows x 64 columns, 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* tm, void* base, long long rows, int cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

int nn_tc_create(NNet* n, char* err, size_t errlen) {
  TCState* t = new TCState();
  n->tc = t;
  t->N = n->s.N;
  t->PP = t->N * t->N;
  t->C = n->C;
  t->T = n->s.tower;
  t->max_batch = n->max_batch;
  t->attr_set = false;
  t->head_pre = nullptr;
  t->trace = nullptr;
  t->stage = nullptr;
  t->stage_cap = 0;
  t->in64 = nullptr;
  t->fuse_heads = 1;
  t->conv5_stages = 6;
  t->l2pf = 0;
  t->pdl = 1;
  t->max_pairs = 0;
  t->res_tma = 1;
  t->split = 0;
  t->split_weights_ready = false;
  for (int i = 0; i < 3; ++i) t->act[i] = nullptr;
  for (int i = 0; i < 3; ++i) t->sact[i] = nullptr;
  if (n->C != 256) {
    snprintf(err, errlen, "the tensor-core path is built for 256 filters");
    return 1;
  }
  // whole boards, with enough spare boards for the last 256-row tile
  const long long n2 = t->PP;
  t->rows_alloc = ((long long)t->max_batch + (256 + n2 - 1) / n2 + 1) * n2;
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  t->num_sms = prop.multiProcessorCount;
  bool ok = cudaMalloc((void**)&t->in64, (size_t)t->rows_alloc * CIN0 * 2) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&t->head_pre, (size_t)t->rows_alloc * sizeof(float4)) == cudaSuccess;
  for (int i = 0; i < 3 && ok; ++i) ok = cudaMalloc((void**)&t->act[i], (size_t)t->rows_alloc * 256 * 2) == cudaSuccess;
  const int nconv = 1 + 2 * t->T;
  t->w.assign(nconv, nullptr);
  t->tm_w.resize(nconv);
  for (int l = 0; l < nconv && ok; ++l) ok = cudaMalloc((void**)&t->w[l], (size_t)9 * 256 * (l == 0 ? CIN0 : 256) * 2) == cudaSuccess;
  if (!ok) {
    snprintf(err, errlen, "cudaMalloc failed for the tensor-core activations (%lld rows)", t->rows_alloc);
    return 1;
  }
  cudaMemset(t->in64, 0, (size_t)t->rows_alloc * CIN0 * 2);
  for (int i = 0; i < 3; ++i) cudaMemset(t->act[i], 0, (size_t)t->rows_alloc * 256 * 2);
  int rc = 0;
  for (int i = 0; i < 3 && !rc; ++i) rc = make_map(&t->tm_act[i], t->act[i], t->rows_alloc, 256, BM);
  for (int l = 0; l < nconv && !rc; ++l) rc = make_map(&t->tm_w[l], t->w[l], 9 * 256, l == 0 ? CIN0 : 256, 128);
  // the boards dimension covers the whole allocation so that tiles running past the batch read zero-initialised rows
  const long long boards = t->rows_alloc / n2;
  if (!rc) rc = make_map_im2col(&t->tm5_in64, t->in64, t->N, boards, CIN0);
  for (int i = 0; i < 3 && !rc; ++i) rc = make_map_im2col(&t->tm5_act[i], t->act[i], t->N, boards, 256);
  t->groups = (t->max_batch % 2 == 0 && t->max_batch >= 2) ? 2 : 1;
  t->grp_rows = (long long)(t->max_batch / 2) * t->PP;
  if (t->groups == 2) {
    for (int g = 0; g < 2 && !rc; ++g) {  // same buffers, base shifted by one group; the boards dimension ends where the allocation ends
      const long long gb = boards - (long long)g * (t->max_batch / 2);
      rc = make_map_im2col(&t->tm5g_in64[g], t->in64 + (size_t)g * t->grp_rows * CIN0, t->N, gb, CIN0);
      for (int i = 0; i < 3 && !rc; ++i) rc = make_map_im2col(&t->tm5g_act[g][i], t->act[i] + (size_t)g * t->grp_rows * 256, t->N, gb, 256);
    }
  }
  if (rc) {
    snprintf(err, errlen, "cuTensorMapEncode failed (%d)", rc);
    return 1;
  }
  return 0;
}

// agz_set_option "conv.*" (include/agz.h): experiment knobs of the tensor-core path, all read at launch time
int nn_tc_set_option(NNet* n, const char* key, long long value) {
  TCState* t = (TCState*)n->tc;
  if (!strcmp(key, "conv.fuse_heads")) t->fuse_heads = value != 0;
  else if (!strcmp(key, "conv.stages")) { if (value != 4 && value != 6) return 2; t->conv5_stages = (int)value; }
  else if (!strcmp(key, "conv.l2_prefetch")) t->l2pf = value != 0;
  else if (!strcmp(key, "conv.pdl")) t->pdl = value != 0;
  else if (!strcmp(key, "conv.max_pairs")) { if (value < 0) return 2; t->max_pairs = (int)value; }
  else if (!strcmp(key, "conv.res_tma")) t->res_tma = value != 0;
  else if (!strcmp(key, "conv.precision")) {
    if (value != 1 && value != 2) return 2;
    if (value == 2 && !t->sact[0]) {   // split precision: its own activation buffers and weight arrays, allocated on first use
      const int nconv = 1 + 2 * t->T;
      bool ok = true;
      for (int i = 0; i < 3 && ok; ++i) ok = cudaMalloc((void**)&t->sact[i], (size_t)t->rows_alloc * 512 * 2) == cudaSuccess;
      t->sw.assign(nconv, nullptr);
      t->tm_sw.resize(nconv);
      for (int l = 0; l < nconv && ok; ++l) ok = cudaMalloc((void**)&t->sw[l], (size_t)9 * 256 * 2 * (l == 0 ? CIN0 : 256) * 2) == cudaSuccess;
      if (!ok) return 2;
      for (int i = 0; i < 3; ++i) cudaMemset(t->sact[i], 0, (size_t)t->rows_alloc * 512 * 2);
      int rc = 0;
      const long long boards = t->rows_alloc / t->PP;
      for (int i = 0; i < 3 && !rc; ++i) rc = make_map_im2col(&t->tm5_sact[i], t->sact[i], t->N, boards, 512);
      for (int l = 0; l < nconv && !rc; ++l) rc = make_map(&t->tm_sw[l], t->sw[l], 9 * 256, 2 * (l == 0 ? CIN0 : 256), 128);
      if (rc) return 2;
      t->split_weights_ready = false;
    }
    t->split = value == 2;
  }
  else return 1;
  return 0;
}

int nn_tc_get_option(const NNet* n, const char* key, long long* value) {
  const TCState* t = (const TCState*)n->tc;
  if (!strcmp(key, "conv.fuse_heads")) *value = t->fuse_heads;
  else if (!strcmp(key, "conv.stages")) *value = t->conv5_stages;
  else if (!strcmp(key, "conv.l2_prefetch")) *value = t->l2pf;
  else if (!strcmp(key, "conv.pdl")) *value = t->pdl;
  else if (!strcmp(key, "conv.max_pairs")) *value = t->max_pairs;
  else if (!strcmp(key, "conv.res_tma")) *value = t->res_tma;
  else if (!strcmp(key, "conv.precision")) *value = t->split ? 2 : 1;
  else return 1;
  return 0;
}

void nn_tc_destroy(NNet* n) {
  TCState* t = (TCState*)n->tc;
  if (!t) return;
  cudaFree(t->in64);
  cudaFree(t->head_pre);
  cudaFree(t->stage);
  for (int i = 0; i < 3; ++i) cudaFree(t->act[i]);
  for (int i = 0; i < 3; ++i) cudaFree(t->sact[i]);
  for (auto p : t->w) cudaFree(p);
  for (auto p : t->sw) cudaFree(p);
  delete t;
  n->tc = nullptr;
}

// Flux (3, 3, Cin, Cout) fp32, column-major -> Wt[tap][co][ci] fp16 with tap = kj*3 + ki for input offset (dj, di) =
// (kj-1, ki-1); Flux Conv is a true convolution, so tap (ki, kj) uses W[2-ki, 2-kj].  Done on the device: the host only
// ships the raw parameter list (what train hands selfplay every iteration).
__global__ void reorder_weights_kernel(const float* __restrict__ flux, __half* __restrict__ out, int cin, int cin_pad) {
  const size_t total = (size_t)9 * 256 * cin_pad;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % cin_pad);
    const int co = (int)((idx / cin_pad) % 256);
    const int tap = (int)(idx / ((size_t)cin_pad * 256));
    const int kj = tap / 3, ki = tap % 3;
    float w = 0.f;
    if (ci < cin) w = flux[(size_t)(2 - ki) + 3 * (2 - kj) + 9 * (size_t)ci + 9 * (size_t)cin * co];
    out[idx] = __float2half(w);
  }
}

// split precision: Wt[tap][co][0..cin_pad) = fp16(w), Wt[tap][co][cin_pad..2*cin_pad) = fp16(w - hi)
__global__ void reorder_weights_split_kernel(const float* __restrict__ flux, __half* __restrict__ out, int cin, int cin_pad) {
  const size_t total = (size_t)9 * 256 * cin_pad;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % cin_pad);
    const int co = (int)((idx / cin_pad) % 256);
    const int tap = (int)(idx / ((size_t)cin_pad * 256));
    const int kj = tap / 3, ki = tap % 3;
    float w = 0.f;
    if (ci < cin) w = flux[(size_t)(2 - ki) + 3 * (2 - kj) + 9 * (size_t)ci + 9 * (size_t)cin * co];
    const __half hi = __float2half(w);
    const size_t o = ((size_t)tap * 256 + co) * (2 * (size_t)cin_pad);
    out[o + ci] = hi;
    out[o + cin_pad + ci] = __float2half(w - __half2float(hi));
  }
}

// conv weights of both precisions from the fp32 base-chain parameter list on the device
static void reorder_all(NNet* n, const float* d_base, cudaStream_t s) {
  TCState* t = (TCState*)n->tc;
  const size_t C = 256, P = (size_t)n->s.planes;
  const int nconv = 1 + 2 * n->s.tower;
  size_t off = 0;
  for (int l = 0; l < nconv; ++l) {
    const int cin = l == 0 ? (int)P : 256, cin_pad = l == 0 ? CIN0 : 256;
    reorder_weights_kernel<<<296, 256, 0, s>>>(d_base + off, t->w[l], cin, cin_pad);
    if (t->sact[0]) reorder_weights_split_kernel<<<296, 256, 0, s>>>(d_base + off, t->sw[l], cin, cin_pad);
    // next conv weight inside the Flux list: stem = W,b,beta,gamma; block = W1,b1,W2,b2,beta1,gamma1,beta2,gamma2
    if (l == 0) off += 9 * P * C + 3 * C;
    else if (l % 2 == 1) off += 9 * C * C + C;          // W1, b1 -> W2
    else off += 9 * C * C + C + 4 * C;                   // W2, b2, beta1, gamma1, beta2, gamma2 -> next block
  }
  t->split_weights_ready = t->sact[0] != nullptr;
}

int nn_tc_commit(NNet* n, const std::vector<ConvLayerHost>& convs, cudaStream_t s, char* err, size_t errlen) {
  TCState* t = (TCState*)n->tc;
  // raw base-chain parameters (Flux order) go up in one transfer; conv weights are located inside it
  const size_t nbase = n->hparams[0].size();
  if (t->stage_cap < nbase) {
    cudaFree(t->stage);
    t->stage = nullptr;
    if (cudaMalloc((void**)&t->stage, nbase * sizeof(float)) != cudaSuccess) { snprintf(err, errlen, "staging buffer allocation failed"); return 1; }
    t->stage_cap = nbase;
  }
  cudaMemcpyAsync(t->stage, n->hparams[0].data(), nbase * sizeof(float), cudaMemcpyHostToDevice, s);
  (void)convs;
  reorder_all(n, t->stage, s);
  if (cudaStreamSynchronize(s) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    snprintf(err, errlen, "uploading / reordering the conv weights failed");
    return 1;
  }
  return 0;
}

// the same reorder with the fp32 base-chain parameters already on the device (training master copy, train.cu train_publish)
int nn_tc_commit_device(NNet* n, const float* d_base, cudaStream_t s, char* err, size_t errlen) {
  TCState* t = (TCState*)n->tc;
  // keep the staging copy current too: a later switch to split precision re-reads the weights from it
  const size_t nbase = nn_param_count(n, 0);
  if (t->stage_cap < nbase) {
    cudaFree(t->stage);
    t->stage = nullptr;
    t->stage_cap = 0;
    if (cudaMalloc((void**)&t->stage, nbase * sizeof(float)) != cudaSuccess) { snprintf(err, errlen, "staging buffer allocation failed"); return 1; }
    t->stage_cap = nbase;
  }
  cudaMemcpyAsync(t->stage, d_base, nbase * sizeof(float), cudaMemcpyDeviceToDevice, s);
  reorder_all(n, t->stage, s);
  if (cudaGetLastError() != cudaSuccess) { snprintf(err, errlen, "reordering the conv weights failed"); return 1; }
  return 0;
}

long long nn_tc_launches_per_forward(const NNet* n) { return 1 + 2 * n->s.tower + 1; }

static int launch_conv5(TCState* t, const CUtensorMap& tmA, const CUtensorMap& tmW, const float* scale, const float* shift, const __half* res,
                        __half* out, int B, int kchunks, cudaStream_t s, const CUtensorMap* res_map = nullptr, int res_row0 = 0, bool pdl = false,
                        const NNet* heads_of = nullptr /* fuse this network's head 1x1 convs into the epilogue */, bool split = false) {
  ConvArgs a;
  a.scale = scale; a.shift = shift; a.res = res; a.out = out;
  a.rows_valid = (long long)B * t->PP;
  a.n_tiles = (int)((a.rows_valid + BM - 1) / BM);
  a.stem = kchunks == 1;
  const int cin = kchunks * BK;   // channels per precision half
  if (!split) {
    a.kchunks = kchunks;
    for (int k = 0; k < kchunks; ++k) { a.a_off[k] = (int16_t)(k * BK); a.w_off[k] = (int16_t)(k * BK); }
  } else if (a.stem) {            // the stem's input planes are exact in fp16: x * w_hi + x * w_lo
    a.kchunks = 2;
    a.a_off[0] = 0; a.w_off[0] = 0;
    a.a_off[1] = 0; a.w_off[1] = (int16_t)cin;
  } else {                        // x_hi * w_hi + x_lo * w_hi + x_hi * w_lo (the lo * lo term is below fp32 resolution)
    a.kchunks = 3 * kchunks;
    for (int k = 0; k < kchunks; ++k) {
      a.a_off[k] = (int16_t)(k * BK);                      a.w_off[k] = (int16_t)(k * BK);
      a.a_off[kchunks + k] = (int16_t)(cin + k * BK);      a.w_off[kchunks + k] = (int16_t)(k * BK);
      a.a_off[2 * kchunks + k] = (int16_t)(k * BK);        a.w_off[2 * kchunks + k] = (int16_t)(cin + k * BK);
    }
  }
  a.N = t->N; a.PP = t->PP;
  a.relu = 1;
  a.trace = t->trace;
  a.head_vw = nullptr; a.head_pw = nullptr; a.head_aff = nullptr; a.head_out = nullptr;
  a.l2pf = t->l2pf;
  const int n_ptiles = (a.n_tiles + 1) / 2;
  // Same number of waves on as few CTA pairs as possible: 9x9 with 8192 (4096) positions is 2592 (1296) pair tiles =
  // 36 (18) waves on 72 pairs exactly, where 74 pairs would idle through a 37th (19th) wave's worth of tail.  The SMs
  // left over run the tree / feature / heads kernels of the other half batch.
  const int max_pairs = t->max_pairs > 0 && t->max_pairs < t->num_sms / 2 ? t->max_pairs : t->num_sms / 2;
  const int waves = (n_ptiles + max_pairs - 1) / max_pairs;
  int pairs = (n_ptiles + waves - 1) / waves;
  if (heads_of && res_map && t->res_tma) {
    a.head_vw = heads_of->f_vw; a.head_pw = heads_of->f_pw; a.head_aff = heads_of->f_head_aff_d;
    a.head_out = t->head_pre + res_row0;
  }
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3(2 * pairs);
  lc.blockDim = dim3(256);
  lc.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = (pdl && t->pdl) ? 1 : 0;
  cudaError_t rc;
  if (split) {
    lc.dynamicSmemBytes = CONV3_SMEM;
    if (a.stem) rc = cudaLaunchKernelEx(&lc, conv3x3_tc5_kernel<4, true>, tmA, tmW, a);
    else rc = cudaLaunchKernelEx(&lc, conv3x3_tc5_kernel<6, true>, tmA, tmW, a);
  } else if (res_map && t->res_tma) {
    lc.dynamicSmemBytes = CONV6_SMEM;
    if (a.head_vw) rc = cudaLaunchKernelEx(&lc, conv3x3_tc6_kernel<true>, tmA, tmW, *res_map, a, res_row0);
    else rc = cudaLaunchKernelEx(&lc, conv3x3_tc6_kernel<false>, tmA, tmW, *res_map, a, res_row0);
  } else {
    lc.dynamicSmemBytes = CONV3_SMEM;
    // the stem (one K chunk per tap, bound by its epilogue) measures 9 % faster with 4 operand stages, the tower convs 1 % slower
    if (t->conv5_stages == 4 || kchunks == 1) rc = cudaLaunchKernelEx(&lc, conv3x3_tc5_kernel<4, false>, tmA, tmW, a);
    else rc = cudaLaunchKernelEx(&lc, conv3x3_tc5_kernel<6, false>, tmA, tmW, a);
  }
  return rc != cudaSuccess ? (int)rc : (int)cudaGetLastError();
}

// debug: trunk rows fp16 [b*N2 + p][256] -> reference layout fp32 [b][c][p]
__global__ void trunk_to_f32_kernel(const __half* __restrict__ act, float* __restrict__ out, int B, int N2, int split) {
  const size_t total = (size_t)B * N2 * 256;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % 256);
    const size_t row = idx / 256;
    const int p = (int)(row % N2), b = (int)(row / N2);
    const float x = split ? __half2float(act[row * 512 + c]) + __half2float(act[row * 512 + 256 + c]) : __half2float(act[idx]);
    out[((size_t)b * 256 + c) * N2 + p] = x;
  }
}

int nn_forward_tc(NNet* n, int B, float* pi, float* v, cudaStream_t s, char* err, size_t errlen, cudaEvent_t* ev, int group, cudaEvent_t convs_done,
                  cudaStream_t heads_stream, const NNDebug* dbg) {
  TCState* t = (TCState*)n->tc;
  if (group >= 0 && (t->groups != 2 || group > 1 || B > t->max_batch / 2)) { snprintf(err, errlen, "bad group"); return 1; }
  const size_t roff = group > 0 ? (size_t)t->grp_rows : 0;   // row offset of this group inside the shared buffers
  if (B > t->max_batch) { snprintf(err, errlen, "batch %d exceeds max_batch %d", B, t->max_batch); return 1; }
  if (!t->attr_set) {
    cudaError_t rc = cudaFuncSetAttribute(heads_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (rc == cudaSuccess) rc = cudaFuncSetAttribute(heads_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heads_smem(n->N2, n->A));
    if (rc == cudaSuccess) rc = cudaFuncSetAttribute(conv3x3_tc5_kernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV3_SMEM);
    if (rc == cudaSuccess) rc = cudaFuncSetAttribute(conv3x3_tc5_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV3_SMEM);
    if (rc == cudaSuccess) rc = cudaFuncSetAttribute(conv3x3_tc5_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV3_SMEM);
    if (rc == cudaSuccess) rc = cudaFuncSetAttribute(conv3x3_tc5_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV3_SMEM);
    if (rc == cudaSuccess) rc = cudaFuncSetAttribute(conv3x3_tc6_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV6_SMEM);
    if (rc == cudaSuccess) rc = cudaFuncSetAttribute(conv3x3_tc6_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV6_SMEM);
    if (rc != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(rc)); return 1; }
    t->attr_set = true;
  }
  if (ev) cudaEventRecord(ev[0], s);
  const int n_blocks = (dbg && dbg->n_blocks >= 0 && dbg->n_blocks < t->T) ? dbg->n_blocks : t->T;   // debug: stop after n_blocks blocks
  const bool split = t->split != 0;
  if (split && group >= 0) { snprintf(err, errlen, "split precision does not run half batches"); return 1; }
  if (split && !t->split_weights_ready) {   // the option was switched on after the last commit: build the hi / lo weights now
    if (!t->stage || t->stage_cap < nn_param_count(n, 0)) { snprintf(err, errlen, "split precision: no parameters committed yet"); return 1; }
    reorder_all(n, t->stage, s);
  }
  const bool fuse = !split && t->fuse_heads && t->res_tma && t->T >= 1 && !(dbg && dbg->trunk) && n_blocks == t->T;
  auto conv = [&](int in_buf /* -1 = stem input */, int layer, int res_buf, int out_buf) {
    const bool pdl = in_buf >= 0;   // tower convolutions directly follow another convolution on the same stream
    const NNet* hf = (fuse && layer == 2 * t->T) ? n : nullptr;   // the last convolution feeds the heads directly
    const int kch = in_buf < 0 ? CIN0 / BK : 4;
    const __half* res = res_buf >= 0 ? t->act[res_buf] + roff * 256 : nullptr;
    const CUtensorMap* rmap = res_buf >= 0 ? &t->tm_act[res_buf] : nullptr;   // plain 2-D map (128 rows x 64 ch boxes) over the shortcut buffer
    if (split)
      return launch_conv5(t, in_buf < 0 ? t->tm5_in64 : t->tm5_sact[in_buf], t->tm_sw[layer], n->f_scale[layer], n->f_shift[layer],
                          res_buf >= 0 ? t->sact[res_buf] : nullptr, t->sact[out_buf], B, kch, s, nullptr, 0, pdl, nullptr, true);
    if (group >= 0)
      return launch_conv5(t, in_buf < 0 ? t->tm5g_in64[group] : t->tm5g_act[group][in_buf], t->tm_w[layer], n->f_scale[layer], n->f_shift[layer], res,
                          t->act[out_buf] + roff * 256, B, kch, s, rmap, (int)roff, pdl, hf);
    return launch_conv5(t, in_buf < 0 ? t->tm5_in64 : t->tm5_act[in_buf], t->tm_w[layer], n->f_scale[layer], n->f_shift[layer], res, t->act[out_buf], B, kch, s,
                        rmap, 0, pdl, hf);
  };
  int rc = conv(-1, 0, -1, 0);
  if (ev) cudaEventRecord(ev[1], s);
  int h = 0, t1 = 1, t2 = 2;
  for (int blk = 0; blk < n_blocks && !rc; ++blk) {
    rc = conv(h, 1 + 2 * blk, -1, t1);
    if (!rc) rc = conv(t1, 2 + 2 * blk, h, t2);
    int tmp = h; h = t2; t2 = tmp;
  }
  if (rc) { snprintf(err, errlen, "conv launch: %s", cudaGetErrorString((cudaError_t)rc)); return 1; }
  if (ev) cudaEventRecord(ev[2], s);
  if (convs_done) cudaEventRecord(convs_done, s);
  if (heads_stream && convs_done) {   // the heads run on another stream, after this group's last convolution
    cudaStreamWaitEvent(heads_stream, convs_done, 0);
    s = heads_stream;
  }
  if (dbg && dbg->trunk) trunk_to_f32_kernel<<<592, 256, 0, s>>>(split ? t->sact[h] : t->act[h] + roff * 256, dbg->trunk, B, n->N2, split ? 1 : 0);
  if (n_blocks < t->T) return cudaGetLastError() == cudaSuccess ? 0 : 1;   // debug: a truncated tower has no heads
  const size_t hsm = heads_smem(n->N2, n->A);
  heads_tc_kernel<<<(B + HPB - 1) / HPB, 256, hsm, s>>>(split ? t->sact[h] : t->act[h] + roff * 256, n->f_vw, n->f_pw, n->f_head_aff_d, n->f_D1W, n->f_D1b, n->f_D2W, n->f_D2b, n->f_PW,
                                                        n->f_Pb, pi, v, B, t->N, n->A, t->trace, fuse ? t->head_pre + roff : nullptr, dbg ? dbg->raw : nullptr, split ? 1 : 0);
  if (ev) cudaEventRecord(ev[3], s);
  cudaError_t e2 = cudaGetLastError();
  if (e2 != cudaSuccess) { snprintf(err, errlen, "heads launch: %s", cudaGetErrorString(e2)); return 1; }
  return 0;
}

int engine_tc_features(const Cfg& c, const View& v, NNet* n, int row0, int nrows, int smem_per_warp, cudaStream_t s) {
  TCState* t = (TCState*)n->tc;
  int rc = 0;
  switch (c.KA) {
    case 3: { LeafFeaturesTCOp<3> op{c, v, t->in64, row0}; rc = devrt::launch_warps(op, nrows, smem_per_warp, s); } break;
    case 6: { LeafFeaturesTCOp<6> op{c, v, t->in64, row0}; rc = devrt::launch_warps(op, nrows, smem_per_warp, s); } break;
    default: { LeafFeaturesTCOp<12> op{c, v, t->in64, row0}; rc = devrt::launch_warps(op, nrows, smem_per_warp, s); } break;
  }
  return rc;
}

int nn_tc_groups(const NNet* n) { return ((TCState*)n->tc)->groups; }
void nn_tc_set_trace(NNet* n, unsigned long long* trace) { ((TCState*)n->tc)->trace = trace; }

int engine_host_features_tc(const Cfg& c, NNet* n, const int8_t* boards_hist, const int8_t* to_play, int B, cudaStream_t s) {
  TCState* t = (TCState*)n->tc;
  int8_t *dbh = nullptr, *dtp = nullptr;
  const size_t nb = (size_t)B * 8 * c.N2;
  cudaError_t rc = cudaMalloc((void**)&dbh, nb);
  if (rc == cudaSuccess) rc = cudaMalloc((void**)&dtp, (size_t)B);
  if (rc == cudaSuccess) rc = cudaMemcpyAsync(dbh, boards_hist, nb, cudaMemcpyHostToDevice, s);
  if (rc == cudaSuccess) rc = cudaMemcpyAsync(dtp, to_play, (size_t)B, cudaMemcpyHostToDevice, s);
  if (rc == cudaSuccess) {
    int blocks = (int)(((size_t)B * c.N2 + 255) / 256);
    host_features_tc_kernel<<<blocks, 256, 0, s>>>(dbh, dtp, t->in64, B, c.N);
    rc = cudaGetLastError();
  }
  cudaError_t rs = cudaStreamSynchronize(s);
  if (rc == cudaSuccess) rc = rs;
  cudaFree(dbh);
  cudaFree(dtp);
  return (int)rc;
}

}  // namespace agz
