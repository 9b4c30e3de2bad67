// nn_f32.cu -- network state, parameter handling, and the fp32 SIMT implementation of the forward pass
// (src/neural_net.jl:57-68; src/resnet.jl:26-32).  The fp32 path is the on-device cross-check of the
// tcgen05 path in nn_tc.cu; it is not the fast path.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "nn_state.h"

namespace agz {

static size_t base_count(const NNShape& s) {
  size_t C = s.filters;
  return 9 * (size_t)s.planes * C + 3 * C + (size_t)s.tower * (2 * (9 * C * C + C) + 4 * C);
}
static size_t value_count(const NNShape& s) {
  size_t C = s.filters, N2 = (size_t)s.N * s.N;
  return C + 3 + 256 * N2 + 256 + 256 + 1;
}
static size_t policy_count(const NNShape& s) {
  size_t C = s.filters, N2 = (size_t)s.N * s.N, A = (size_t)nn_actions(s);
  return 2 * C + 6 + A * 2 * N2 + A;
}

size_t nn_param_count(const NNet* n, int chain) {
  if (!n) return 0;
  return chain == 0 ? base_count(n->s) : (chain == 1 ? value_count(n->s) : policy_count(n->s));
}
size_t nn_bn_count(const NNet* n, int chain) {
  if (!n) return 0;
  return chain == 0 ? (size_t)n->s.filters * (1 + 2 * n->s.tower) : (chain == 1 ? 1 : 2);
}

double nn_flops_per_position(const NNShape& s) {
  double N2 = (double)s.N * s.N, C = s.filters, A = (double)nn_actions(s);
  double stem = 2 * 9 * s.planes * C * N2;
  double tower = (double)s.tower * 2 * (2 * 9 * C * C * N2);
  double heads = 2 * (3 * C * N2) + 2 * (256 * N2 + 256) + 2 * (A * 2 * N2);
  return stem + tower + heads;
}

#define CUDA_TRY(x)                        \
  do {                                     \
    cudaError_t e__ = (x);                 \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)

template <class T>
static bool dmal(T** p, size_t n) { return cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)) == cudaSuccess; }

NNet* nn_create(const NNShape& s, int max_batch, char* err, size_t errlen) {
  NNet* n = new NNet();
  n->s = s;
  n->N2 = s.N * s.N;
  n->A = nn_actions(s);
  n->C = s.filters;
  n->max_batch = max_batch;
  n->ready = false;
  n->f32_weights_ready = false;
  n->tc = nullptr;
  for (int k = 0; k < 3; ++k) {
    n->have[k] = false;
    n->bn_mode[k] = 0;
    n->hmu[k].assign(nn_bn_count(n, k), 0.f);
    n->hsigma[k].assign(nn_bn_count(n, k), 1.f);
  }
  const int nconv = 1 + 2 * s.tower;
  bool ok = true;
  n->f_w.assign(nconv, nullptr);
  n->f_scale.assign(nconv, nullptr);
  n->f_shift.assign(nconv, nullptr);
  for (int l = 0; l < nconv && ok; ++l) {
    size_t cin = l == 0 ? s.planes : s.filters;
    ok = ok && dmal(&n->f_w[l], (size_t)s.filters * cin * 9) && dmal(&n->f_scale[l], (size_t)s.filters) && dmal(&n->f_shift[l], (size_t)s.filters);
  }
  ok = ok && dmal(&n->f_vw, (size_t)s.filters) && dmal(&n->f_pw, (size_t)2 * s.filters) && dmal(&n->f_head_aff_d, (size_t)6);
  ok = ok && dmal(&n->f_D1W, (size_t)256 * n->N2) && dmal(&n->f_D1b, (size_t)256) && dmal(&n->f_D2W, (size_t)256) && dmal(&n->f_D2b, (size_t)1);
  ok = ok && dmal(&n->f_PW, (size_t)n->A * 2 * n->N2) && dmal(&n->f_Pb, (size_t)n->A);
  // fp32 activations are only needed by the cross-check path; cap the batch it supports to bound memory
  for (int i = 0; i < 3; ++i) n->f_act[i] = nullptr;
  if (!ok) {
    snprintf(err, errlen, "cudaMalloc failed for network parameters");
    nn_destroy(n);
    return nullptr;
  }
  if (nn_tc_create(n, err, errlen)) {
    nn_destroy(n);
    return nullptr;
  }
  return n;
}

void nn_destroy(NNet* n) {
  if (!n) return;
  nn_tc_destroy(n);
  for (auto p : n->f_w) cudaFree(p);
  for (auto p : n->f_scale) cudaFree(p);
  for (auto p : n->f_shift) cudaFree(p);
  cudaFree(n->f_vw); cudaFree(n->f_pw); cudaFree(n->f_head_aff_d);
  cudaFree(n->f_D1W); cudaFree(n->f_D1b); cudaFree(n->f_D2W); cudaFree(n->f_D2b); cudaFree(n->f_PW); cudaFree(n->f_Pb);
  for (int i = 0; i < 3; ++i) cudaFree(n->f_act[i]);
  delete n;
}

int nn_set_params(NNet* n, int chain, const float* flat, size_t cnt) {
  if (cnt != nn_param_count(n, chain)) return 1;
  n->hparams[chain].assign(flat, flat + cnt);
  n->have[chain] = true;
  n->ready = false;
  return 0;
}

int nn_set_bn(NNet* n, int chain, const float* mu, const float* sigma, size_t cnt, int mode) {
  if (cnt != nn_bn_count(n, chain)) return 1;
  n->hmu[chain].assign(mu, mu + cnt);
  n->hsigma[chain].assign(sigma, sigma + cnt);
  n->bn_mode[chain] = mode;
  n->ready = false;
  return 0;
}

bool nn_ready(const NNet* n) { return n && n->ready; }

static void fold_bn(const float* beta, const float* gamma, const float* mu, const float* sigma, const float* bias, int C, int mode,
                    float* scale, float* shift) {
  for (int c = 0; c < C; ++c) {
    float den = mode == 0 ? sqrtf(sigma[c] + 1e-5f) : sigma[c];
    float sc = gamma[c] / den;
    scale[c] = sc;
    shift[c] = beta[c] - mu[c] * sc + sc * bias[c];
  }
}

int nn_fold_layers(const NNet* n, std::vector<ConvLayerHost>& convs, char* err, size_t errlen, bool with_weights) {
  if (!n->have[0] || !n->have[1] || !n->have[2]) {
    snprintf(err, errlen, "parameters of chain %d have not been set (agz_net_set_params)", !n->have[0] ? 0 : (!n->have[1] ? 1 : 2));
    return 1;
  }
  const int C = n->C, T = n->s.tower, P = n->s.planes;
  const float* p = n->hparams[0].data();
  const float* mu = n->hmu[0].data();
  const float* sg = n->hsigma[0].data();
  convs.clear();
  auto add = [&](const float* W, const float* b, const float* beta, const float* gamma, int cin, int bn_index) {
    ConvLayerHost L;
    L.cin = cin; L.cout = C;
    if (with_weights) L.w.assign(W, W + (size_t)9 * cin * C);
    L.scale.resize(C); L.shift.resize(C);
    fold_bn(beta, gamma, mu + (size_t)bn_index * C, sg + (size_t)bn_index * C, b, C, n->bn_mode[0], L.scale.data(), L.shift.data());
    convs.push_back(std::move(L));
  };
  // stem: W, b, beta, gamma
  add(p, p + (size_t)9 * P * C, p + (size_t)9 * P * C + C, p + (size_t)9 * P * C + 2 * C, P, 0);
  p += (size_t)9 * P * C + 3 * C;
  for (int t = 0; t < T; ++t) {  // W1, b1, W2, b2, beta1, gamma1, beta2, gamma2
    const float* W1 = p; const float* b1 = W1 + (size_t)9 * C * C;
    const float* W2 = b1 + C; const float* b2 = W2 + (size_t)9 * C * C;
    const float* be1 = b2 + C; const float* ga1 = be1 + C; const float* be2 = ga1 + C; const float* ga2 = be2 + C;
    add(W1, b1, be1, ga1, C, 1 + 2 * t);
    add(W2, b2, be2, ga2, C, 2 + 2 * t);
    p = ga2 + C;
  }
  return 0;
}

int nn_commit(NNet* n, cudaStream_t s, char* err, size_t errlen) {
  std::vector<ConvLayerHost> convs;
  if (nn_fold_layers(n, convs, err, errlen, false)) return 1;
  const int C = n->C, N2 = n->N2, A = n->A;
  n->f32_weights_ready = false;
  for (size_t l = 0; l < convs.size(); ++l) {
    const ConvLayerHost& L = convs[l];
    cudaMemcpyAsync(n->f_scale[l], L.scale.data(), (size_t)C * 4, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(n->f_shift[l], L.shift.data(), (size_t)C * 4, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
  }
  // value head: W(1,1,C,1), b, beta, gamma, D1W(256,N2), D1b, D2W(1,256), D2b
  const float* v = n->hparams[1].data();
  float aff[6];
  fold_bn(v + C + 1, v + C + 2, n->hmu[1].data(), n->hsigma[1].data(), v + C, 1, n->bn_mode[1], &aff[0], &aff[1]);
  cudaMemcpyAsync(n->f_vw, v, (size_t)C * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(n->f_D1W, v + C + 3, (size_t)256 * N2 * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(n->f_D1b, v + C + 3 + (size_t)256 * N2, 256 * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(n->f_D2W, v + C + 3 + (size_t)256 * N2 + 256, 256 * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(n->f_D2b, v + C + 3 + (size_t)256 * N2 + 512, 4, cudaMemcpyHostToDevice, s);
  // policy head: W(1,1,C,2), b(2), beta(2), gamma(2), DW(A, 2*N2), Db(A)
  const float* p = n->hparams[2].data();
  float sc2[2], sh2[2];
  fold_bn(p + 2 * C + 2, p + 2 * C + 4, n->hmu[2].data(), n->hsigma[2].data(), p + 2 * C, 2, n->bn_mode[2], sc2, sh2);
  aff[2] = sc2[0]; aff[3] = sh2[0]; aff[4] = sc2[1]; aff[5] = sh2[1];
  // 1x1 conv weight W[0,0,ci,co] at ci + C*co -> f_pw[co][ci]
  cudaMemcpyAsync(n->f_pw, p, (size_t)2 * C * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(n->f_PW, p + 2 * C + 6, (size_t)A * 2 * N2 * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(n->f_Pb, p + 2 * C + 6 + (size_t)A * 2 * N2, (size_t)A * 4, cudaMemcpyHostToDevice, s);
  memcpy(n->f_head_aff, aff, sizeof(aff));
  cudaMemcpyAsync(n->f_head_aff_d, aff, sizeof(aff), cudaMemcpyHostToDevice, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) {
    snprintf(err, errlen, "uploading network parameters failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  if (nn_tc_commit(n, convs, s, err, errlen)) return 1;
  n->ready = true;
  return 0;
}

// ----------------------------------------------------------------------------------------------------------
// fp32 kernels.  Activations: [B][C][N2] with point p = N*j + i (i fastest) = the reference's W x H x C x B.

// Register-tiled direct convolution: a CTA of 128 threads computes COT = 32 output channels x PT = 48 consecutive points of one
// board; thread (ty, tx) owns channels co0 + 4*ty .. +3 and points p0 + tx + 16*q, q < 3 (12 accumulators; a training batch is a
// small problem, and small thread tiles are what keeps enough warps on 148 SMs).  Per input channel and
// tap it reads one float4 of weights (layout [ci][tap][co] in global and shared memory, a broadcast within a half warp) and 6 inputs from the zero-haloed
// board in shared memory: 4 shared loads per 12 FMAs.  Input channels arrive in chunks of CIT through a two-stage cp.async pipeline
// (the halo is zeroed once; only interior points are ever written).  Every accumulator still sums its (ci, tap) products in ascending
// order with one fmaf each -- the same arithmetic, bit for bit, as a one-output-per-thread loop (which this kernel replaced: 1 shared
// load per FMA, 9 TFLOP/s).
static const int COT = 32, CIT = 8, PT = 96, PQ = 6, CONV_THREADS = 128;

static size_t conv_f32_stage(int N) { return (size_t)((CIT * (N + 2) * (N + 2) + 3) / 4 * 4 + CIT * 9 * COT); }   // floats per stage
static size_t conv_f32_smem(int N) { return 2 * conv_f32_stage(N) * sizeof(float); }
static int conv_f32_chunks(int N) { return (N * N + PT - 1) / PT; }

__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int K>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(K)); }
// r / n for 0 <= r < 2^20 and small n, without the integer-division sequence: (r + 0.5) / n is at least 0.5 / n away from an integer
__device__ __forceinline__ int small_div(int r, float inv_n) { return __float2int_rz(((float)r + 0.5f) * inv_n); }

__global__ void __launch_bounds__(CONV_THREADS) conv3x3_f32_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                                   const float* __restrict__ res, float* __restrict__ out, int Cin, int Cout,
                                                                   int N, int relu, int chunks) {
  extern __shared__ __align__(16) float sm[];
  const int N2 = N * N, NP = N + 2, NPP = NP * NP;
  const int in_floats = (CIT * NPP + 3) / 4 * 4, stage = in_floats + CIT * 9 * COT;
  const float inv_n = 1.0f / (float)N;
  const int b = blockIdx.x / chunks, p0 = (blockIdx.x % chunks) * PT, co0 = blockIdx.y * COT, tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  for (int x = tid; x < 2 * stage; x += CONV_THREADS) sm[x] = 0.f;   // halos (and channels past Cin / Cout) stay zero
  __syncthreads();
  auto fill = [&](int buf, int ci0) {
    float* sin = sm + buf * stage;     // [CIT][NPP]
    float* sw = sin + in_floats;       // [CIT][9][COT]
    for (int ci = 0; ci < CIT && ci0 + ci < Cin; ++ci) {
      const float* src = in + ((size_t)b * Cin + ci0 + ci) * N2;
      for (int p = tid; p < N2; p += CONV_THREADS) {
        const int jj = small_div(p, inv_n), ii = p - jj * N;
        cp_async4(sin + ci * NPP + (jj + 1) * NP + ii + 1, src + p);
      }
    }
    // weights arrive as w[ci][tap][co] (output channel fastest): a row of COT channels is contiguous in both memories
    const bool vec = (Cout & 3) == 0;
    for (int x = tid; x < CIT * 9 * (COT / 4); x += CONV_THREADS) {
      const int row = x / (COT / 4), c4 = (x % (COT / 4)) * 4;   // row = ci * 9 + tap inside the chunk
      if (ci0 * 9 + row < Cin * 9) {
        const float* src = w + ((size_t)ci0 * 9 + row) * Cout + co0 + c4;
        float* dst = sw + row * COT + c4;
        if (vec && co0 + c4 + 3 < Cout) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (co0 + c4 + e < Cout) cp_async4(dst + e, src + e);
        }
      }
    }
    cp_async_commit();
  };
  int base[PQ];                                  // index of the window's corner (jj-1, ii-1) in the padded board = jj*NP + ii
#pragma unroll
  for (int q = 0; q < PQ; ++q) {
    const int p = p0 + tx + 16 * q;
    const int jj = small_div(p, inv_n);
    base[q] = p < N2 ? jj * NP + (p - jj * N) : 0;
  }
  float acc[4][PQ];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int q = 0; q < PQ; ++q) acc[c][q] = 0.f;
  fill(0, 0);
  int buf = 0;
  for (int ci0 = 0; ci0 < Cin; ci0 += CIT, buf ^= 1) {
    if (ci0 + CIT < Cin) {
      fill(buf ^ 1, ci0 + CIT);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* sin = sm + buf * stage;
    const float* sw = sin + in_floats;
    const int nci = Cin - ci0 < CIT ? Cin - ci0 : CIT;
    for (int ci = 0; ci < nci; ++ci) {
      const float* xi = sin + ci * NPP;
      const float4* wp = reinterpret_cast<const float4*>(sw + ci * 9 * COT + 4 * ty);
#pragma unroll
      for (int kj = 0; kj < 3; ++kj)
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) {
          const float4 wv = wp[(kj * 3 + ki) * (COT / 4)];
          const int off = kj * NP + ki;
#pragma unroll
          for (int q = 0; q < PQ; ++q) {
            const float x = xi[base[q] + off];
            acc[0][q] = fmaf(wv.x, x, acc[0][q]);
            acc[1][q] = fmaf(wv.y, x, acc[1][q]);
            acc[2][q] = fmaf(wv.z, x, acc[2][q]);
            acc[3][q] = fmaf(wv.w, x, acc[3][q]);
          }
        }
    }
    __syncthreads();   // this stage is refilled by the next iteration's prefetch
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int co = co0 + 4 * ty + c;
    if (co < Cout) {
      const float sc = scale[co], sh = shift[co];
#pragma unroll
      for (int q = 0; q < PQ; ++q) {
        const int p = p0 + tx + 16 * q;
        if (p < N2) {
          const size_t o = ((size_t)b * Cout + co) * N2 + p;
          float y = acc[c][q] * sc + sh;
          if (res) y += res[o];
          if (relu) y = fmaxf(y, 0.f);
          out[o] = y;
        }
      }
    }
  }
}

// heads (neural_net.jl:23-30): one block per position
__global__ void __launch_bounds__(256) heads_f32_kernel(const float* __restrict__ trunk, const float* __restrict__ vw,
                                                        const float* __restrict__ pw, const float* __restrict__ aff,
                                                        const float* __restrict__ D1W, const float* __restrict__ D1b,
                                                        const float* __restrict__ D2W, const float* __restrict__ D2b,
                                                        const float* __restrict__ PW, const float* __restrict__ Pb, float* __restrict__ pi,
                                                        float* __restrict__ v, int C, int N2, int A, float* __restrict__ raw) {
  extern __shared__ float sm[];
  float* vf = sm;             // [N2]
  float* pf = sm + N2;        // [2*N2], index p + N2*c
  float* hid = sm + 3 * N2;   // [256]
  float* red = hid + 256;     // [256]
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* x = trunk + (size_t)b * C * N2;
  for (int p = tid; p < N2; p += 256) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int c = 0; c < C; ++c) {
      float xv = x[(size_t)c * N2 + p];
      a0 = fmaf(vw[c], xv, a0);
      a1 = fmaf(pw[c], xv, a1);
      a2 = fmaf(pw[C + c], xv, a2);
    }
    vf[p] = fmaxf(a0 * aff[0] + aff[1], 0.f);
    pf[p] = fmaxf(a1 * aff[2] + aff[3], 0.f);
    pf[N2 + p] = fmaxf(a2 * aff[4] + aff[5], 0.f);
  }
  __syncthreads();
  {  // Dense(N2 -> 256, relu): W (256, N2) column-major
    float a = D1b[tid];
    for (int i = 0; i < N2; ++i) a = fmaf(D1W[tid + 256 * i], vf[i], a);
    hid[tid] = fmaxf(a, 0.f);
  }
  __syncthreads();
  red[tid] = D2W[tid] * hid[tid];  // Dense(256 -> 1): W (1, 256)
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  if (tid == 0) {
    v[b] = tanhf(red[0] + D2b[0]);
    if (raw) raw[(size_t)b * (A + 1) + A] = red[0] + D2b[0];
  }
  __syncthreads();
  // Dense(2*N2 -> A) + softmax
  float lg[2];
  float mx = -INFINITY;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    int a = tid + q * 256;
    lg[q] = -INFINITY;
    if (a < A) {
      float acc = Pb[a];
      for (int i = 0; i < 2 * N2; ++i) acc = fmaf(PW[a + (size_t)A * i], pf[i], acc);
      lg[q] = acc;
      mx = fmaxf(mx, acc);
      if (raw) raw[(size_t)b * (A + 1) + a] = acc;
    }
  }
  red[tid] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] = fmaxf(red[tid], red[tid + s]);
    __syncthreads();
  }
  mx = red[0];
  __syncthreads();
  float ex[2], sum = 0.f;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    ex[q] = lg[q] == -INFINITY ? 0.f : expf(lg[q] - mx);
    sum += ex[q];
  }
  red[tid] = sum;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  sum = red[0];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    int a = tid + q * 256;
    if (a < A) pi[(size_t)b * A + a] = ex[q] / sum;
  }
}

long long nn_f32_launches_per_forward(const NNet* n) { return 1 + 2 * n->s.tower + 1; }

// plain launch of the fp32 convolution for other translation units (train.cu: forward and data-gradient convolutions)
int conv3x3_f32_launch(const float* in, const float* w, const float* scale, const float* shift, const float* res, float* out, int B, int Cin,
                       int Cout, int N, int relu, cudaStream_t s) {
  const size_t smem = conv_f32_smem(N);
  static int attr_n = -1;
  if (attr_n != N) {
    if (cudaFuncSetAttribute(conv3x3_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    attr_n = N;
  }
  const int chunks = conv_f32_chunks(N);
  conv3x3_f32_kernel<<<dim3(B * chunks, (Cout + COT - 1) / COT), CONV_THREADS, smem, s>>>(in, w, scale, shift, res, out, Cin, Cout, N, relu, chunks);
  return (int)cudaGetLastError();
}

// fp32 conv weights are reordered and uploaded the first time the cross-check path is used after a commit
static int upload_f32_weights(NNet* n, cudaStream_t s) {
  std::vector<ConvLayerHost> convs;
  char err[128];
  if (nn_fold_layers(n, convs, err, sizeof(err), true)) return (int)cudaErrorInvalidValue;
  for (size_t l = 0; l < convs.size(); ++l) {
    const ConvLayerHost& L = convs[l];
    // device layout w[ci][t][co] (output channel fastest), t = kj*3 + ki for the input offset (dj, di) = (kj-1, ki-1):
    // Flux Conv is a true convolution, so tap (ki, kj) of the correlation uses W[2-ki, 2-kj]
    std::vector<float> w((size_t)L.cout * L.cin * 9);
    for (int co = 0; co < L.cout; ++co)
      for (int ci = 0; ci < L.cin; ++ci)
        for (int kj = 0; kj < 3; ++kj)
          for (int ki = 0; ki < 3; ++ki)
            w[((size_t)ci * 9 + kj * 3 + ki) * L.cout + co] = L.w[(size_t)(2 - ki) + 3 * (2 - kj) + 9 * (size_t)ci + 9 * (size_t)L.cin * co];
    cudaMemcpyAsync(n->f_w[l], w.data(), w.size() * 4, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
  }
  n->f32_weights_ready = true;
  return (int)cudaGetLastError();
}

// debug: trunk [B][C][N2] is already the reference layout in the fp32 path
int nn_forward_f32(NNet* n, const float* feats, int B, float* pi, float* v, cudaStream_t s, cudaEvent_t* ev, const NNDebug* dbg) {
  const int C = n->C, N = n->s.N, N2 = n->N2;
  if (B > n->max_batch) return (int)cudaErrorInvalidValue;
  if (!n->f32_weights_ready) {
    int rc = upload_f32_weights(n, s);
    if (rc) return rc;
  }
  for (int i = 0; i < 3; ++i)
    if (!n->f_act[i]) CUDA_TRY(cudaMalloc((void**)&n->f_act[i], (size_t)n->max_batch * C * N2 * sizeof(float)));
  const size_t smem = conv_f32_smem(N);
  CUDA_TRY(cudaFuncSetAttribute(conv3x3_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int chunks = conv_f32_chunks(N);
  dim3 grid(B * chunks, (C + COT - 1) / COT);
  if (ev) cudaEventRecord(ev[0], s);
  conv3x3_f32_kernel<<<grid, CONV_THREADS, smem, s>>>(feats, n->f_w[0], n->f_scale[0], n->f_shift[0], nullptr, n->f_act[0], n->s.planes, C, N, 1, chunks);
  if (ev) cudaEventRecord(ev[1], s);
  float *h = n->f_act[0], *t1 = n->f_act[1], *t2 = n->f_act[2];
  const int n_blocks = (dbg && dbg->n_blocks >= 0 && dbg->n_blocks < n->s.tower) ? dbg->n_blocks : n->s.tower;
  for (int t = 0; t < n_blocks; ++t) {
    conv3x3_f32_kernel<<<grid, CONV_THREADS, smem, s>>>(h, n->f_w[1 + 2 * t], n->f_scale[1 + 2 * t], n->f_shift[1 + 2 * t], nullptr, t1, C, C, N, 1, chunks);
    conv3x3_f32_kernel<<<grid, CONV_THREADS, smem, s>>>(t1, n->f_w[2 + 2 * t], n->f_scale[2 + 2 * t], n->f_shift[2 + 2 * t], h, t2, C, C, N, 1, chunks);
    float* tmp = h; h = t2; t2 = tmp;
  }
  if (ev) cudaEventRecord(ev[2], s);
  if (dbg && dbg->trunk) cudaMemcpyAsync(dbg->trunk, h, (size_t)B * C * N2 * sizeof(float), cudaMemcpyDeviceToDevice, s);
  if (n_blocks < n->s.tower) return (int)cudaGetLastError();
  const size_t hsm = (size_t)(3 * N2 + 512) * sizeof(float);
  heads_f32_kernel<<<B, 256, hsm, s>>>(h, n->f_vw, n->f_pw, n->f_head_aff_d, n->f_D1W, n->f_D1b, n->f_D2W, n->f_D2b, n->f_PW, n->f_Pb, pi, v, C, N2, n->A,
                                       dbg ? dbg->raw : nullptr);
  if (ev) cudaEventRecord(ev[3], s);
  return (int)cudaGetLastError();
}

}  // namespace agz
