// train.cu -- one optimisation step of the policy/value network on the device, fp32.
//
// Replaces (reference paths): src/neural_net.jl  _train :81-101 (one 32-position minibatch), loss_pi :75, loss_value :77,
// loss_reg :80-83, the train-mode forward `nn(positions, true)` :57-68; src/train.jl:54 `Momentum(2f-2)`.
//   loss = 0.01 * crossentropy(p, pi) + 0.01 * mse(z, v) + 1e-4 * sum(theta .^ 2)          (Flux 0.10.4 definitions:
//          crossentropy = -sum(pi .* log.(p)) / B,  mse = sum((z .- v) .^ 2) / B)
//   BatchNorm in train mode normalises with the batch mean and the biased batch variance (eps = 1e-5) and moves the
//   running statistics with momentum 0.1 (variance unbiased by m / (m - 1));
//   Momentum(eta, rho = 0.9): v = rho * v - eta * grad; theta += v.
// The reference's _train does not run as committed (NamedTuple signature vs tuple call site, undefined loss_avg) and the
// semantics above live in Flux, which is not vendored: parity is against the oracle restatement (oracle/train.py, torch
// autograd), "parity unpinned" by the reference itself.
//
// This is the caller-side row of SURVEY 8(f), not the self-play hot path: a step is one 32-position minibatch per game
// (train.jl:68), ~0.01 % of the loop's arithmetic, so the kernels are fp32 SIMT (the convolutions reuse the register-tiled
// fp32 kernel of the cross-check path for forward and data gradients; the weight gradient has its own tiled kernel).
// Parameters live in three flat device buffers in Flux `params` order (what save_model writes, train.jl:27-33).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "nn_state.h"
#include "train.h"

namespace agz {

// w[ci][t][co] (output channel fastest), t = kj*3 + ki = correlation tap (dj, di) = (kj-1, ki-1)
// w[co][ci][t], t = kj*3 + ki = correlation tap (dj, di) = (kj-1, ki-1)
int conv3x3_f32_launch(const float* in, const float* w, const float* scale, const float* shift, const float* res, float* out, int B, int Cin,
                       int Cout, int N, int relu, cudaStream_t s);

static const float BN_EPS = 1e-5f, BN_MOMENTUM = 0.1f;
static const float W_POLICY = 0.01f, W_VALUE = 0.01f, W_REG = 1e-4f;

struct ConvOff {  // offsets of one Conv + BatchNorm pair inside the base chain's flat parameter buffer
  size_t W, b, beta, gamma;
  int cin;
};

struct TrainState {
  int N, N2, A, C, T, planes, maxB;
  size_t np[3];
  float *P[3], *G[3], *V[3];        // parameters, gradients, momentum
  std::vector<ConvOff> conv;        // stem, then (W1, W2) per block
  float *mu_run, *var_run;          // base BatchNorm running statistics [(1+2T)*C]
  float *hmu_run, *hvar_run;        // heads: [3] value, policy 0, policy 1
  // saved forward state
  float* feats;                     // [B][17][N2]
  std::vector<float*> z, a;         // per conv layer: pre-BatchNorm output, post-activation output [B][C][N2]
  float *mean, *invstd;             // [(1+2T)*C] batch statistics of this step
  float *hz, *ha, *hmean, *hinvstd; // heads: 1x1 conv outputs / activations [B][3][N2] stored as value [B][1][N2] then policy [B][2][N2]
  float *hid, *vout, *logits, *prob;  // [B][256], [B], [B][A], [B][A]
  float *d_pi, *d_z;                // targets
  // backward scratch
  float *dA, *dS, *dT, *dZ;         // [B][C][N2]
  float *dhz, *dha, *dhid, *dlogit, *dvpre;
  float *wc, *wd, *dwc, *dwpart;    // reordered weights / weight gradient of the current layer [C][C][9]; its WG_SPLIT partial sums
  float *ones, *zeros;              // [C]
  float *red;                       // small reduction scratch
  bool dirty;                       // device parameters are newer than the host copy in NNet
};

// ------------------------------------------------------------------------------------------------ small kernels
// Flux W[a + 3b + 9ci + 9Cin*co] (true convolution) -> correlation layout of the fp32 convolution kernel (nn_f32.cu):
// w[ci][kj*3+ki][co] (output channel fastest) with (ki, kj) = (2-a, 2-b)
__global__ void k_flux_to_corr(const float* __restrict__ W, float* __restrict__ wc, int Cin, int Cout) {
  const size_t total = (size_t)Cout * Cin * 9;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(x % Cout), t = (int)((x / Cout) % 9), ci = (int)(x / ((size_t)9 * Cout));
    const int kj = t / 3, ki = t % 3;
    wc[x] = W[(size_t)(2 - ki) + 3 * (2 - kj) + 9 * (size_t)ci + 9 * (size_t)Cin * co];
  }
}
// data-gradient weights: dx = corr(dz, wd), a convolution from Cout to Cin channels whose weight for (input co, tap t, output ci) is
// the forward weight of (ci, tap 8 - t, co): wd[co][t][ci] (its output channel ci fastest)
__global__ void k_flux_to_dgrad(const float* __restrict__ W, float* __restrict__ wd, int Cin, int Cout) {
  const size_t total = (size_t)Cout * Cin * 9;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(x % Cin), t = (int)((x / Cin) % 9), co = (int)(x / ((size_t)9 * Cin));
    const int t2 = 8 - t, kj = t2 / 3, ki = t2 % 3;
    wd[x] = W[(size_t)(2 - ki) + 3 * (2 - kj) + 9 * (size_t)ci + 9 * (size_t)Cin * co];
  }
}
// weight gradient in correlation layout -> Flux layout, written into the gradient buffer
__global__ void k_corr_to_flux(const float* __restrict__ dwc, float* __restrict__ dW, int Cin, int Cout) {
  const size_t total = (size_t)Cout * Cin * 9;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(x % 9), ci = (int)((x / 9) % Cin), co = (int)(x / ((size_t)9 * Cin));
    const int kj = t / 3, ki = t % 3;
    dW[(size_t)(2 - ki) + 3 * (2 - kj) + 9 * (size_t)ci + 9 * (size_t)Cin * co] = dwc[x];
  }
}

__device__ __forceinline__ float block_sum(float v, float* sh) {  // blockDim.x = 256
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) sh[tid] += sh[tid + s];
    __syncthreads();
  }
  const float r = sh[0];
  __syncthreads();
  return r;
}

// x[B][C][P]: per-channel batch mean and biased variance; running statistics moved with momentum 0.1
__global__ void __launch_bounds__(256) k_bn_stats(const float* __restrict__ x, int B, int C, int P, float* __restrict__ mean,
                                                  float* __restrict__ invstd, float* __restrict__ mu_run, float* __restrict__ var_run) {
  __shared__ float sh[256];
  const int c = blockIdx.x, m = B * P;
  float s = 0.f;
  for (int i = threadIdx.x; i < m; i += 256) s += x[((size_t)(i / P) * C + c) * P + (i % P)];
  const float mu = block_sum(s, sh) / (float)m;
  float q = 0.f;
  for (int i = threadIdx.x; i < m; i += 256) {
    const float d = x[((size_t)(i / P) * C + c) * P + (i % P)] - mu;
    q += d * d;
  }
  const float var = block_sum(q, sh) / (float)m;
  if (threadIdx.x == 0) {
    mean[c] = mu;
    invstd[c] = 1.0f / sqrtf(var + BN_EPS);
    mu_run[c] = (1.f - BN_MOMENTUM) * mu_run[c] + BN_MOMENTUM * mu;
    var_run[c] = (1.f - BN_MOMENTUM) * var_run[c] + BN_MOMENTUM * var * ((float)m / (float)(m > 1 ? m - 1 : 1));
  }
}

// y = relu(gamma * (x - mean) * invstd + beta [+ res])
__global__ void k_bn_fwd(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ res,
                         float* __restrict__ y, int B, int C, int P) {
  const size_t total = (size_t)B * C * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / P) % C);
    float v = gamma[c] * ((x[i] - mean[c]) * invstd[c]) + beta[c];
    if (res) v += res[i];
    y[i] = fmaxf(v, 0.f);
  }
}

// g = dOut * (a > 0); dbeta[c] = sum g; dgamma[c] = sum g * xhat
__global__ void __launch_bounds__(256) k_bn_bwd_reduce(const float* __restrict__ dout, const float* __restrict__ a, const float* __restrict__ x,
                                                       const float* __restrict__ mean, const float* __restrict__ invstd, int B, int C, int P,
                                                       float* __restrict__ dbeta, float* __restrict__ dgamma) {
  __shared__ float sh[256];
  const int c = blockIdx.x, m = B * P;
  float sb = 0.f, sg = 0.f;
  for (int i = threadIdx.x; i < m; i += 256) {
    const size_t o = ((size_t)(i / P) * C + c) * P + (i % P);
    const float g = a[o] > 0.f ? dout[o] : 0.f;
    sb += g;
    sg += g * ((x[o] - mean[c]) * invstd[c]);
  }
  const float tb = block_sum(sb, sh), tg = block_sum(sg, sh);
  if (threadIdx.x == 0) { dbeta[c] = tb; dgamma[c] = tg; }
}

// dz = gamma * invstd / m * (m * g - dbeta - xhat * dgamma); optionally also writes g (the shortcut gradient)
__global__ void k_bn_bwd_apply(const float* __restrict__ dout, const float* __restrict__ a, const float* __restrict__ x,
                               const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                               const float* __restrict__ dbeta, const float* __restrict__ dgamma, float* __restrict__ dz,
                               float* __restrict__ gout, int B, int C, int P) {
  const size_t total = (size_t)B * C * P;
  const float m = (float)(B * P);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / P) % C);
    const float g = a[i] > 0.f ? dout[i] : 0.f;
    const float xhat = (x[i] - mean[c]) * invstd[c];
    dz[i] = gamma[c] * invstd[c] / m * (m * g - dbeta[c] - xhat * dgamma[c]);
    if (gout) gout[i] = g;
  }
}

// out[c] = sum over b, p of x[b][c][p]
__global__ void __launch_bounds__(256) k_channel_sum(const float* __restrict__ x, int B, int C, int P, float* __restrict__ out) {
  __shared__ float sh[256];
  const int c = blockIdx.x, m = B * P;
  float s = 0.f;
  for (int i = threadIdx.x; i < m; i += 256) s += x[((size_t)(i / P) * C + c) * P + (i % P)];
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) out[c] = t;
}

__global__ void k_add(float* __restrict__ y, const float* __restrict__ x, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += x[i];
}

// weight gradient of a 3x3 convolution in correlation layout: dwc[co][ci][t] = sum_b sum_p dz[b][co][p] * in[b][ci][p + off(t)]
// A CTA of 128 threads owns 32 output x 16 input channels; a thread owns 2 x 2 channels and all 9 taps (36 accumulators).  Walking
// a board row, the 3 x 3 input window slides: each step loads one new column (3 values per input channel) and one gradient per
// output channel -- 8 shared loads per 36 FMAs (the one-pair-per-thread loop this replaced did 10 per 9).  The tiles of the next
// board travel by cp.async while the current one is accumulated.  590 k outputs at 36 per thread are only 512 warps, so the batch is
// split WG_SPLIT ways over gridDim.z into partial sums that k_wgrad_reduce adds in a fixed order (deterministic).
static const int WG_CO = 32, WG_CI = 16, WG_THREADS = 128, WG_SPLIT = 4;
static size_t wgrad_stage(int N) { return (size_t)((WG_CI * (N + 2) * (N + 2) + 8 + 3) / 4 * 4 + WG_CO * N * N); }   // floats per stage
static size_t wgrad_smem(int N) { return 2 * wgrad_stage(N) * sizeof(float); }

__device__ __forceinline__ void wg_cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}

__global__ void __launch_bounds__(WG_THREADS) k_conv_wgrad(const float* __restrict__ in, const float* __restrict__ dz, float* __restrict__ part,
                                                           int B, int Cin, int Cout, int N) {
  extern __shared__ float sm[];
  const int N2 = N * N, NP = N + 2, NPP = NP * NP;
  const int in_floats = (WG_CI * NPP + 8 + 3) / 4 * 4, stage = in_floats + WG_CO * N2;
  const float inv_n = 1.0f / (float)N;
  const int co0 = blockIdx.x * WG_CO, ci0 = blockIdx.y * WG_CI, tid = threadIdx.x;
  const int per = (B + (int)gridDim.z - 1) / (int)gridDim.z, b_lo = blockIdx.z * per, b_hi = min(B, b_lo + per);
  const int tc = tid >> 3, ti = tid & 7;   // output channels co0 + 2*tc + {0,1}, input channels ci0 + 2*ti + {0,1}
  for (int x = tid; x < 2 * stage; x += WG_THREADS) sm[x] = 0.f;   // halos, the 8 floats past the last board, channels past Cin / Cout
  __syncthreads();
  auto fill = [&](int buf, int b) {
    float* sin = sm + buf * stage;   // [WG_CI][NPP] zero halo (+ 8 zero floats: the sliding window reads up to 2 columns past a row)
    float* sdz = sin + in_floats;    // [WG_CO][N2]
    for (int c = 0; c < WG_CI && ci0 + c < Cin; ++c) {
      const float* src = in + ((size_t)b * Cin + ci0 + c) * N2;
      for (int p = tid; p < N2; p += WG_THREADS) {
        const int jj = __float2int_rz(((float)p + 0.5f) * inv_n), ii = p - jj * N;
        wg_cp_async4(sin + c * NPP + (jj + 1) * NP + ii + 1, src + p);
      }
    }
    for (int c = 0; c < WG_CO && co0 + c < Cout; ++c) {
      const float* src = dz + ((size_t)b * Cout + co0 + c) * N2;
      for (int p = tid; p < N2; p += WG_THREADS) wg_cp_async4(sdz + c * N2 + p, src + p);
    }
    asm volatile("cp.async.commit_group;");
  };
  float acc[2][2][9];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[a][c][t] = 0.f;
  if (b_lo < b_hi) fill(0, b_lo);
  for (int b = b_lo; b < b_hi; ++b) {
    const int buf = (b - b_lo) & 1;
    if (b + 1 < b_hi) {
      fill(buf ^ 1, b + 1);
      asm volatile("cp.async.wait_group 1;");
    } else {
      asm volatile("cp.async.wait_group 0;");
    }
    __syncthreads();
    const float* x0 = sm + buf * stage + (2 * ti) * NPP;
    const float* x1 = x0 + NPP;
    const float* g0 = sm + buf * stage + in_floats + (2 * tc) * N2;
    const float* g1 = g0 + N2;
    for (int jj = 0; jj < N; ++jj) {
      const float* r0 = x0 + jj * NP;
      const float* r1 = x1 + jj * NP;
      float w0[2][3], w1[2][3], w2[2][3];   // window columns [input channel][kj]
#pragma unroll
      for (int kj = 0; kj < 3; ++kj) {
        w0[0][kj] = r0[kj * NP]; w0[1][kj] = r1[kj * NP];
        w1[0][kj] = r0[kj * NP + 1]; w1[1][kj] = r1[kj * NP + 1];
      }
#define AGZ_WG_STEP(II, CA, CB, CC)                                                                \
      {                                                                                            \
        _Pragma("unroll") for (int kj = 0; kj < 3; ++kj) { CC[0][kj] = r0[kj * NP + (II) + 2]; CC[1][kj] = r1[kj * NP + (II) + 2]; } \
        if ((II) < N) {                                                                            \
          const float ga = g0[jj * N + (II)], gb = g1[jj * N + (II)];                              \
          _Pragma("unroll") for (int c = 0; c < 2; ++c)                                            \
          _Pragma("unroll") for (int kj = 0; kj < 3; ++kj) {                                       \
            acc[0][c][kj * 3 + 0] = fmaf(ga, CA[c][kj], acc[0][c][kj * 3 + 0]);                    \
            acc[0][c][kj * 3 + 1] = fmaf(ga, CB[c][kj], acc[0][c][kj * 3 + 1]);                    \
            acc[0][c][kj * 3 + 2] = fmaf(ga, CC[c][kj], acc[0][c][kj * 3 + 2]);                    \
            acc[1][c][kj * 3 + 0] = fmaf(gb, CA[c][kj], acc[1][c][kj * 3 + 0]);                    \
            acc[1][c][kj * 3 + 1] = fmaf(gb, CB[c][kj], acc[1][c][kj * 3 + 1]);                    \
            acc[1][c][kj * 3 + 2] = fmaf(gb, CC[c][kj], acc[1][c][kj * 3 + 2]);                    \
          }                                                                                        \
        }                                                                                          \
      }
      for (int ii = 0; ii < N; ii += 3) {
        AGZ_WG_STEP(ii, w0, w1, w2)
        AGZ_WG_STEP(ii + 1, w1, w2, w0)
        AGZ_WG_STEP(ii + 2, w2, w0, w1)
      }
#undef AGZ_WG_STEP
    }
    __syncthreads();   // this stage is refilled by the next iteration's prefetch
  }
  float* dst = part + (size_t)blockIdx.z * Cout * Cin * 9;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int co = co0 + 2 * tc + a, ci = ci0 + 2 * ti + c;
      if (co < Cout && ci < Cin) {
#pragma unroll
        for (int t = 0; t < 9; ++t) dst[((size_t)co * Cin + ci) * 9 + t] = acc[a][c][t];
      }
    }
}

// dwc = ((part0 + part1) + part2) + part3
__global__ void k_wgrad_reduce(const float* __restrict__ part, float* __restrict__ dwc, size_t n, int splits) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float a = part[i];
    for (int k = 1; k < splits; ++k) a += part[(size_t)k * n + i];
    dwc[i] = a;
  }
}

// 1x1 convolution (Flux weight w[ci + C*k]): out[b][k][p] = bias[k] + sum_c w[c + C*k] * x[b][c][p]
__global__ void k_conv1x1_fwd(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                              int B, int C, int K, int P) {
  const size_t total = (size_t)B * K * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % P), k = (int)((i / P) % K), b = (int)(i / ((size_t)K * P));
    float acc = bias[k];
    for (int c = 0; c < C; ++c) acc = fmaf(w[c + C * k], x[((size_t)b * C + c) * P + p], acc);
    out[i] = acc;
  }
}
// dw[c + C*k] = sum_{b,p} x[b][c][p] * dz[b][k][p]   (one block per (c, k))
__global__ void __launch_bounds__(256) k_conv1x1_wgrad(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ dw, int B,
                                                       int C, int K, int P) {
  __shared__ float sh[256];
  const int c = blockIdx.x, k = blockIdx.y, m = B * P;
  float s = 0.f;
  for (int i = threadIdx.x; i < m; i += 256) {
    const int b = i / P, p = i % P;
    s += x[((size_t)b * C + c) * P + p] * dz[((size_t)b * K + k) * P + p];
  }
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) dw[c + C * k] = t;
}
// dx[b][c][p] (+)= sum_k w[c + C*k] * dz[b][k][p]
__global__ void k_conv1x1_dgrad(const float* __restrict__ dz, const float* __restrict__ w, float* __restrict__ dx, int B, int C, int K, int P,
                                int accumulate) {
  const size_t total = (size_t)B * C * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % P), c = (int)((i / P) % C), b = (int)(i / ((size_t)C * P));
    float acc = accumulate ? dx[i] : 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(w[c + C * k], dz[((size_t)b * K + k) * P + p], acc);
    dx[i] = acc;
  }
}

// Dense, Flux weight W[o + O*i]: y[b][o] = act(bias[o] + sum_i W[o + O*i] * x[b][i]); act 0 none, 1 relu, 2 tanh
__global__ void k_dense_fwd(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ y, int B,
                            int I, int O, int act) {
  const size_t total = (size_t)B * O;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx % O), b = (int)(idx / O);
    float acc = bias[o];
    const float* xb = x + (size_t)b * I;
    for (int i = 0; i < I; ++i) acc = fmaf(W[o + (size_t)O * i], xb[i], acc);
    if (act == 1) acc = fmaxf(acc, 0.f);
    else if (act == 2) acc = tanhf(acc);
    y[idx] = acc;
  }
}
// dW[o + O*i] = sum_b dy[b][o] * x[b][i]; db[o] = sum_b dy[b][o]
__global__ void k_dense_wgrad(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dW, float* __restrict__ db, int B, int I,
                              int O) {
  const size_t total = (size_t)O * (I + 1);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx % O), i = (int)(idx / O);
    float acc = 0.f;
    if (i < I) {
      for (int b = 0; b < B; ++b) acc = fmaf(dy[(size_t)b * O + o], x[(size_t)b * I + i], acc);
      dW[o + (size_t)O * i] = acc;
    } else {
      for (int b = 0; b < B; ++b) acc += dy[(size_t)b * O + o];
      db[o] = acc;
    }
  }
}
// dx[b][i] = sum_o W[o + O*i] * dy[b][o]
__global__ void k_dense_dgrad(const float* __restrict__ dy, const float* __restrict__ W, float* __restrict__ dx, int B, int I, int O) {
  const size_t total = (size_t)B * I;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % I), b = (int)(idx / I);
    float acc = 0.f;
    const float* wi = W + (size_t)O * i;
    const float* dyb = dy + (size_t)b * O;
    for (int o = 0; o < O; ++o) acc = fmaf(wi[o], dyb[o], acc);
    dx[idx] = acc;
  }
}

// softmax over A + the two data-loss terms and their gradients; one block per position.
//   red[0] += -sum_a pi log p (policy term, unweighted), red[1] += (z - v)^2
//   dlogit = W_POLICY / B * (p * sum(pi) - pi);  dvpre = W_VALUE / B * 2 (v - z) (1 - v^2)
__global__ void __launch_bounds__(256) k_loss(const float* __restrict__ logits, const float* __restrict__ v, const float* __restrict__ pi,
                                              const float* __restrict__ z, float* __restrict__ prob, float* __restrict__ dlogit,
                                              float* __restrict__ dvpre, float* __restrict__ red, int B, int A) {
  __shared__ float sh[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* l = logits + (size_t)b * A;
  const float* t = pi + (size_t)b * A;
  float mx = -INFINITY;
  for (int a = tid; a < A; a += 256) mx = fmaxf(mx, l[a]);
  sh[tid] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) sh[tid] = fmaxf(sh[tid], sh[tid + s]);
    __syncthreads();
  }
  mx = sh[0];
  __syncthreads();
  float se = 0.f, sp = 0.f;
  for (int a = tid; a < A; a += 256) { se += expf(l[a] - mx); sp += t[a]; }
  const float sum = block_sum(se, sh), spi = block_sum(sp, sh);
  const float lse = mx + logf(sum);
  float ce = 0.f;
  for (int a = tid; a < A; a += 256) {
    const float p = expf(l[a] - mx) / sum;
    prob[(size_t)b * A + a] = p;
    ce -= t[a] * (l[a] - lse);   // log p through the log-softmax (finite where the reference's log.(p) underflows)
    dlogit[(size_t)b * A + a] = W_POLICY / (float)B * (p * spi - t[a]);
  }
  const float cet = block_sum(ce, sh);
  if (tid == 0) {
    const float d = v[b] - z[b];
    atomicAdd(&red[0], cet);
    atomicAdd(&red[1], d * d);
    dvpre[b] = W_VALUE / (float)B * 2.f * d * (1.f - v[b] * v[b]);
  }
}

__global__ void k_relu_mask(float* __restrict__ g, const float* __restrict__ a, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (!(a[i] > 0.f)) g[i] = 0.f;
}

// sum of squares of a flat buffer into red[slot]
__global__ void __launch_bounds__(256) k_sumsq(const float* __restrict__ x, size_t n, float* __restrict__ red, int slot) {
  __shared__ float sh[256];
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) s += x[i] * x[i];
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(&red[slot], t);
}

// Momentum step with the regulariser's gradient folded in: g = G + 2 * W_REG * P; V = rho * V - eta * g; P += V
__global__ void k_momentum(float* __restrict__ P, const float* __restrict__ G, float* __restrict__ V, size_t n, float eta, float rho) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float g = G[i] + 2.f * W_REG * P[i];
    const float v = rho * V[i] - eta * g;
    V[i] = v;
    P[i] += v;
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int grid_for(size_t n) {
  size_t g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

template <class T>
static bool dm(T** p, size_t n) {
  if (cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return false;
  return cudaMemset(*p, 0, (n ? n : 1) * sizeof(T)) == cudaSuccess;
}

void train_destroy(TrainState* t) {
  if (!t) return;
  for (int k = 0; k < 3; ++k) { cudaFree(t->P[k]); cudaFree(t->G[k]); cudaFree(t->V[k]); }
  cudaFree(t->mu_run); cudaFree(t->var_run); cudaFree(t->hmu_run); cudaFree(t->hvar_run);
  cudaFree(t->feats);
  for (float* p : t->z) cudaFree(p);
  for (float* p : t->a) cudaFree(p);
  cudaFree(t->mean); cudaFree(t->invstd); cudaFree(t->hz); cudaFree(t->ha); cudaFree(t->hmean); cudaFree(t->hinvstd);
  cudaFree(t->hid); cudaFree(t->vout); cudaFree(t->logits); cudaFree(t->prob); cudaFree(t->d_pi); cudaFree(t->d_z);
  cudaFree(t->dA); cudaFree(t->dS); cudaFree(t->dT); cudaFree(t->dZ);
  cudaFree(t->dhz); cudaFree(t->dha); cudaFree(t->dhid); cudaFree(t->dlogit); cudaFree(t->dvpre);
  cudaFree(t->wc); cudaFree(t->wd); cudaFree(t->dwc); cudaFree(t->dwpart); cudaFree(t->ones); cudaFree(t->zeros); cudaFree(t->red);
  delete t;
}

TrainState* train_create(const NNet* n, int max_batch, char* err, size_t errlen) {
  TrainState* t = new TrainState();
  t->N = n->s.N; t->N2 = n->N2; t->A = n->A; t->C = n->C; t->T = n->s.tower; t->planes = n->s.planes; t->maxB = max_batch;
  t->dirty = false;
  for (int k = 0; k < 3; ++k) { t->np[k] = nn_param_count(n, k); t->P[k] = t->G[k] = t->V[k] = nullptr; }
  const size_t C = t->C, N2 = t->N2, A = t->A, B = max_batch, nl = 1 + 2 * (size_t)t->T;
  size_t off = 0;
  for (size_t l = 0; l < nl; ++l) {   // Flux params order of the base chain (see nn_state.h)
    ConvOff o;
    o.cin = l == 0 ? t->planes : (int)C;
    if (l == 0) {
      o.W = off; off += 9 * (size_t)o.cin * C; o.b = off; off += C; o.beta = off; off += C; o.gamma = off; off += C;
      t->conv.push_back(o);
    } else if (l % 2 == 1) {          // a block: W1, b1, W2, b2, beta1, gamma1, beta2, gamma2
      ConvOff o2;
      o2.cin = (int)C;
      o.W = off; off += 9 * C * C; o.b = off; off += C;
      o2.W = off; off += 9 * C * C; o2.b = off; off += C;
      o.beta = off; off += C; o.gamma = off; off += C;
      o2.beta = off; off += C; o2.gamma = off; off += C;
      t->conv.push_back(o);
      t->conv.push_back(o2);
    }
  }
  bool ok = off == t->np[0];
  if (!ok) { snprintf(err, errlen, "base parameter layout mismatch (%zu vs %zu)", off, t->np[0]); train_destroy(t); return nullptr; }
  for (int k = 0; k < 3 && ok; ++k) ok = dm(&t->P[k], t->np[k]) && dm(&t->G[k], t->np[k]) && dm(&t->V[k], t->np[k]);
  ok = ok && dm(&t->mu_run, nl * C) && dm(&t->var_run, nl * C) && dm(&t->hmu_run, (size_t)3) && dm(&t->hvar_run, (size_t)3);
  ok = ok && dm(&t->feats, B * 17 * N2);
  t->z.assign(nl, nullptr);
  t->a.assign(nl, nullptr);
  for (size_t l = 0; l < nl && ok; ++l) ok = dm(&t->z[l], B * C * N2) && dm(&t->a[l], B * C * N2);
  ok = ok && dm(&t->mean, nl * C) && dm(&t->invstd, nl * C) && dm(&t->hz, B * 3 * N2) && dm(&t->ha, B * 3 * N2) && dm(&t->hmean, (size_t)3) &&
       dm(&t->hinvstd, (size_t)3);
  ok = ok && dm(&t->hid, B * 256) && dm(&t->vout, B) && dm(&t->logits, B * A) && dm(&t->prob, B * A) && dm(&t->d_pi, B * A) && dm(&t->d_z, B);
  ok = ok && dm(&t->dA, B * C * N2) && dm(&t->dS, B * C * N2) && dm(&t->dT, B * C * N2) && dm(&t->dZ, B * C * N2);
  ok = ok && dm(&t->dhz, B * 3 * N2) && dm(&t->dha, B * 3 * N2) && dm(&t->dhid, B * 256) && dm(&t->dlogit, B * A) && dm(&t->dvpre, B);
  ok = ok && dm(&t->wc, 9 * C * C) && dm(&t->wd, 9 * C * C) && dm(&t->dwc, 9 * C * C) && dm(&t->dwpart, (size_t)WG_SPLIT * 9 * C * C) && dm(&t->ones, C) && dm(&t->zeros, C) && dm(&t->red, (size_t)8);
  if (!ok) { snprintf(err, errlen, "device allocation for training failed (batch %d)", max_batch); train_destroy(t); return nullptr; }
  std::vector<float> one(C, 1.f);
  cudaMemcpy(t->ones, one.data(), C * sizeof(float), cudaMemcpyHostToDevice);
  return t;
}

int train_max_batch(const TrainState* t) { return t->maxB; }

// host parameters / running statistics of the NNet -> device training buffers (momentum is kept)
int train_load(TrainState* t, const NNet* n, cudaStream_t s, char* err, size_t errlen) {
  for (int k = 0; k < 3; ++k) {
    if (!n->have[k] || n->hparams[k].size() != t->np[k]) { snprintf(err, errlen, "chain %d has no parameters", k); return 1; }
    cudaMemcpyAsync(t->P[k], n->hparams[k].data(), t->np[k] * sizeof(float), cudaMemcpyHostToDevice, s);
  }
  // running statistics: the engine keeps (mu, sigma) with sigma = variance (AGZ_BN_VAR_EPS) or std (AGZ_BN_STD)
  const size_t nb = n->hmu[0].size();
  std::vector<float> var(nb);
  for (size_t i = 0; i < nb; ++i) var[i] = n->bn_mode[0] == 1 ? n->hsigma[0][i] * n->hsigma[0][i] : n->hsigma[0][i];
  cudaMemcpyAsync(t->mu_run, n->hmu[0].data(), nb * sizeof(float), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(t->var_run, var.data(), nb * sizeof(float), cudaMemcpyHostToDevice, s);
  float hm[3] = {n->hmu[1][0], n->hmu[2][0], n->hmu[2][1]};
  float hv[3] = {n->hsigma[1][0], n->hsigma[2][0], n->hsigma[2][1]};
  if (n->bn_mode[1] == 1) hv[0] *= hv[0];
  if (n->bn_mode[2] == 1) { hv[1] *= hv[1]; hv[2] *= hv[2]; }
  cudaMemcpyAsync(t->hmu_run, hm, sizeof(hm), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(t->hvar_run, hv, sizeof(hv), cudaMemcpyHostToDevice, s);
  cudaError_t rc = cudaStreamSynchronize(s);
  if (rc != cudaSuccess) { snprintf(err, errlen, "train_load: %s", cudaGetErrorString(rc)); return 1; }
  t->dirty = false;
  return 0;
}

// ---- hand-over of the updated parameters to the inference paths WITHOUT a host round trip: fold conv bias + BatchNorm (running
// statistics, variance + eps convention) into the per-channel scale / shift of every convolution, refresh the head parameters, and
// let the tensor-core path re-read its fp16 tap-major weights straight from the fp32 master copy (nn_tc_commit_device).
__global__ void k_fold_bn(const float* __restrict__ beta, const float* __restrict__ gamma, const float* __restrict__ mu, const float* __restrict__ var,
                          const float* __restrict__ bias, int C, float* __restrict__ scale, float* __restrict__ shift, int stride_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // same expression, operation order and roundings as the host fold_bn (nn_f32.cu, mode AGZ_BN_VAR_EPS): no FMA contraction
  const float sc = __fdiv_rn(gamma[c], __fsqrt_rn(__fadd_rn(var[c], BN_EPS)));
  scale[(size_t)c * stride_out] = sc;
  shift[(size_t)c * stride_out] = __fadd_rn(__fsub_rn(beta[c], __fmul_rn(mu[c], sc)), __fmul_rn(sc, bias[c]));
}

int train_publish(TrainState* t, NNet* n, cudaStream_t s, char* err, size_t errlen) {
  const int C = t->C, N2 = t->N2, A = t->A, nl = 1 + 2 * t->T;
  const float* Pb = t->P[0];
  for (int l = 0; l < nl; ++l) {
    const ConvOff& o = t->conv[l];
    k_fold_bn<<<(C + 127) / 128, 128, 0, s>>>(Pb + o.beta, Pb + o.gamma, t->mu_run + (size_t)l * C, t->var_run + (size_t)l * C, Pb + o.b, C,
                                               n->f_scale[l], n->f_shift[l], 1);
  }
  const float *Pv = t->P[1], *Pp = t->P[2];
  const size_t vD1W = (size_t)C + 3, vD1b = vD1W + (size_t)256 * N2, vD2W = vD1b + 256, vD2b = vD2W + 256;
  const size_t pb = 2 * (size_t)C, pDW = pb + 6, pDb = pDW + (size_t)A * 2 * N2;
  // f_head_aff_d = (v scale, v shift, p0 scale, p0 shift, p1 scale, p1 shift): stride 2 interleaves scale / shift
  k_fold_bn<<<1, 32, 0, s>>>(Pv + C + 1, Pv + C + 2, t->hmu_run, t->hvar_run, Pv + C, 1, n->f_head_aff_d, n->f_head_aff_d + 1, 2);
  k_fold_bn<<<1, 32, 0, s>>>(Pp + pb + 2, Pp + pb + 4, t->hmu_run + 1, t->hvar_run + 1, Pp + pb, 2, n->f_head_aff_d + 2, n->f_head_aff_d + 3, 2);
  const cudaMemcpyKind dd = cudaMemcpyDeviceToDevice;
  cudaMemcpyAsync(n->f_vw, Pv, (size_t)C * 4, dd, s);
  cudaMemcpyAsync(n->f_D1W, Pv + vD1W, (size_t)256 * N2 * 4, dd, s);
  cudaMemcpyAsync(n->f_D1b, Pv + vD1b, 256 * 4, dd, s);
  cudaMemcpyAsync(n->f_D2W, Pv + vD2W, 256 * 4, dd, s);
  cudaMemcpyAsync(n->f_D2b, Pv + vD2b, 4, dd, s);
  cudaMemcpyAsync(n->f_pw, Pp, (size_t)2 * C * 4, dd, s);
  cudaMemcpyAsync(n->f_PW, Pp + pDW, (size_t)A * 2 * N2 * 4, dd, s);
  cudaMemcpyAsync(n->f_Pb, Pp + pDb, (size_t)A * 4, dd, s);
  if (nn_tc_commit_device(n, Pb, s, err, errlen)) return 1;
  n->f32_weights_ready = false;   // the fp32 cross-check path rebuilds its weights from the host copy (synchronised lazily)
  n->ready = true;
  return 0;
}

// device training buffers -> host copy in the NNet (parameters in Flux order, running statistics as mean / variance); the
// inference paths are NOT invalidated (train_publish has already handed them the same values)
int train_store(TrainState* t, NNet* n, cudaStream_t s, char* err, size_t errlen) {
  for (int k = 0; k < 3; ++k) {
    n->hparams[k].resize(t->np[k]);
    cudaMemcpyAsync(n->hparams[k].data(), t->P[k], t->np[k] * sizeof(float), cudaMemcpyDeviceToHost, s);
    n->have[k] = true;
  }
  const size_t nb = n->hmu[0].size();
  cudaMemcpyAsync(n->hmu[0].data(), t->mu_run, nb * sizeof(float), cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(n->hsigma[0].data(), t->var_run, nb * sizeof(float), cudaMemcpyDeviceToHost, s);
  float hm[3], hv[3];
  cudaMemcpyAsync(hm, t->hmu_run, sizeof(hm), cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(hv, t->hvar_run, sizeof(hv), cudaMemcpyDeviceToHost, s);
  cudaError_t rc = cudaStreamSynchronize(s);
  if (rc != cudaSuccess) { snprintf(err, errlen, "train_store: %s", cudaGetErrorString(rc)); return 1; }
  n->hmu[1][0] = hm[0]; n->hmu[2][0] = hm[1]; n->hmu[2][1] = hm[2];
  n->hsigma[1][0] = hv[0]; n->hsigma[2][0] = hv[1]; n->hsigma[2][1] = hv[2];
  for (int k = 0; k < 3; ++k) n->bn_mode[k] = 0;   // variance + eps convention from here on
  n->f32_weights_ready = false;
  t->dirty = false;
  return 0;
}

bool train_dirty(const TrainState* t) { return t->dirty; }

__global__ void k_scale(float* __restrict__ x, size_t n, float f) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= f;
}

static int train_step_core(TrainState* t, int B, float eta, float rho, float* loss_out, cudaStream_t s, char* err, size_t errlen, int world,
                           train_allreduce_fn allreduce, void* ctx);

int train_step(TrainState* t, const float* d_feats_in /* [B][17][N2] on the device */, const float* h_pi, const float* h_z, int B, float eta,
               float rho, float* loss_out, cudaStream_t s, char* err, size_t errlen, int world, train_allreduce_fn allreduce, void* ctx) {
  if (B < 1 || B > t->maxB) { snprintf(err, errlen, "batch %d outside [1, %d]", B, t->maxB); return 1; }
  cudaMemcpyAsync(t->feats, d_feats_in, (size_t)B * 17 * t->N2 * sizeof(float), cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(t->d_pi, h_pi, (size_t)B * t->A * sizeof(float), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(t->d_z, h_z, (size_t)B * sizeof(float), cudaMemcpyHostToDevice, s);
  return train_step_core(t, B, eta, rho, loss_out, s, err, errlen, world, allreduce, ctx);
}

// packed replay tuples (replay.cu: pi[A] f32 | 8 boards N2 int8 each, current first | to_play | z | pad) -> the step's inputs:
// get_feats planes (features.jl:3-26) [B][17][N2], pi [B][A], z [B]
__global__ void k_unpack_tuples(const unsigned char* __restrict__ stage, size_t stride, int B, int N2, int A, float* __restrict__ feats,
                                float* __restrict__ pi, float* __restrict__ z) {
  const int b = blockIdx.x;
  const unsigned char* t = stage + (size_t)b * stride;
  const float* tpi = reinterpret_cast<const float*>(t);
  const signed char* boards = reinterpret_cast<const signed char*>(t + (size_t)4 * A);
  const int tp = boards[8 * N2];
  for (int a = threadIdx.x; a < A; a += blockDim.x) pi[(size_t)b * A + a] = tpi[a];
  float* f = feats + (size_t)b * 17 * N2;
  for (int i = threadIdx.x; i < 8 * N2; i += blockDim.x) {
    const int k = i / N2, p = i - k * N2, sv = boards[i];
    f[(2 * k) * N2 + p] = sv == tp ? 1.f : 0.f;
    f[(2 * k + 1) * N2 + p] = sv == -tp ? 1.f : 0.f;
  }
  for (int p = threadIdx.x; p < N2; p += blockDim.x) f[16 * N2 + p] = (float)tp;
  if (threadIdx.x == 0) z[b] = (float)boards[8 * N2 + 1];
}

int train_step_from_tuples(TrainState* t, const unsigned char* d_stage, size_t stride, int B, float eta, float rho, float* loss_out, cudaStream_t s,
                           char* err, size_t errlen, int world, train_allreduce_fn allreduce, void* ctx) {
  if (B < 1 || B > t->maxB) { snprintf(err, errlen, "batch %d outside [1, %d]", B, t->maxB); return 1; }
  k_unpack_tuples<<<B, 128, 0, s>>>(d_stage, stride, B, t->N2, t->A, t->feats, t->d_pi, t->d_z);
  return train_step_core(t, B, eta, rho, loss_out, s, err, errlen, world, allreduce, ctx);
}

static int train_step_core(TrainState* t, int B, float eta, float rho, float* loss_out, cudaStream_t s, char* err, size_t errlen, int world,
                           train_allreduce_fn allreduce, void* ctx) {
  const int C = t->C, N = t->N, N2 = t->N2, A = t->A, T = t->T;
  const int nl = 1 + 2 * T;
  const size_t act = (size_t)B * C * N2;
  float* Pb = t->P[0];
  float* Gb = t->G[0];
  cudaMemsetAsync(t->red, 0, 8 * sizeof(float), s);
  const size_t wsm = wgrad_smem(N);
  cudaFuncSetAttribute(k_conv_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm);

  // ---- forward, train mode (neural_net.jl:57-68 with train = true)
  auto conv_fwd = [&](int l, const float* in) {
    const ConvOff& o = t->conv[l];
    k_flux_to_corr<<<grid_for((size_t)9 * o.cin * C), 256, 0, s>>>(Pb + o.W, t->wc, o.cin, C);
    conv3x3_f32_launch(in, t->wc, t->ones, Pb + o.b, nullptr, t->z[l], B, o.cin, C, N, 0, s);
    k_bn_stats<<<C, 256, 0, s>>>(t->z[l], B, C, N2, t->mean + (size_t)l * C, t->invstd + (size_t)l * C, t->mu_run + (size_t)l * C, t->var_run + (size_t)l * C);
  };
  auto bn_fwd = [&](int l, const float* res) {
    const ConvOff& o = t->conv[l];
    k_bn_fwd<<<grid_for(act), 256, 0, s>>>(t->z[l], t->mean + (size_t)l * C, t->invstd + (size_t)l * C, Pb + o.gamma, Pb + o.beta, res, t->a[l], B, C, N2);
  };
  conv_fwd(0, t->feats);
  bn_fwd(0, nullptr);
  for (int blk = 0; blk < T; ++blk) {
    const int l1 = 1 + 2 * blk, l2 = 2 + 2 * blk, l0 = l1 - 1;
    conv_fwd(l1, t->a[l0]);
    bn_fwd(l1, nullptr);
    conv_fwd(l2, t->a[l1]);
    bn_fwd(l2, t->a[l0]);   // relu(bn2(conv2(y)) + shortcut) (resnet.jl:31)
  }
  const float* trunk = t->a[nl - 1];
  // heads: value chain P[1] = W(C), b(1), beta(1), gamma(1), D1W(256*N2), D1b(256), D2W(256), D2b(1);
  //        policy chain P[2] = W(2C), b(2), beta(2), gamma(2), DW(A*2N2), Db(A)
  float *Pv = t->P[1], *Pp = t->P[2], *Gv = t->G[1], *Gp = t->G[2];
  const size_t vW = 0, vb = C, vbeta = C + 1, vgamma = C + 2, vD1W = C + 3, vD1b = vD1W + (size_t)256 * N2, vD2W = vD1b + 256, vD2b = vD2W + 256;
  const size_t pW = 0, pb = 2 * (size_t)C, pbeta = pb + 2, pgamma = pb + 4, pDW = pb + 6, pDb = pDW + (size_t)A * 2 * N2;
  float* hzv = t->hz;                          // [B][1][N2]
  float* hzp = t->hz + (size_t)B * N2;         // [B][2][N2]
  float* hav = t->ha;
  float* hap = t->ha + (size_t)B * N2;
  k_conv1x1_fwd<<<grid_for((size_t)B * N2), 256, 0, s>>>(trunk, Pv + vW, Pv + vb, hzv, B, C, 1, N2);
  k_conv1x1_fwd<<<grid_for((size_t)B * 2 * N2), 256, 0, s>>>(trunk, Pp + pW, Pp + pb, hzp, B, C, 2, N2);
  k_bn_stats<<<1, 256, 0, s>>>(hzv, B, 1, N2, t->hmean, t->hinvstd, t->hmu_run, t->hvar_run);
  k_bn_stats<<<2, 256, 0, s>>>(hzp, B, 2, N2, t->hmean + 1, t->hinvstd + 1, t->hmu_run + 1, t->hvar_run + 1);
  k_bn_fwd<<<grid_for((size_t)B * N2), 256, 0, s>>>(hzv, t->hmean, t->hinvstd, Pv + vgamma, Pv + vbeta, nullptr, hav, B, 1, N2);
  k_bn_fwd<<<grid_for((size_t)B * 2 * N2), 256, 0, s>>>(hzp, t->hmean + 1, t->hinvstd + 1, Pp + pgamma, Pp + pbeta, nullptr, hap, B, 2, N2);
  k_dense_fwd<<<grid_for((size_t)B * 256), 256, 0, s>>>(hav, Pv + vD1W, Pv + vD1b, t->hid, B, N2, 256, 1);
  k_dense_fwd<<<grid_for((size_t)B), 256, 0, s>>>(t->hid, Pv + vD2W, Pv + vD2b, t->vout, B, 256, 1, 2);
  k_dense_fwd<<<grid_for((size_t)B * A), 256, 0, s>>>(hap, Pp + pDW, Pp + pDb, t->logits, B, 2 * N2, A, 0);
  k_loss<<<B, 256, 0, s>>>(t->logits, t->vout, t->d_pi, t->d_z, t->prob, t->dlogit, t->dvpre, t->red, B, A);
  for (int k = 0; k < 3; ++k) k_sumsq<<<grid_for(t->np[k]), 256, 0, s>>>(t->P[k], t->np[k], t->red, 2);

  // ---- backward: heads
  // policy: Dense(2N2 -> A)
  k_dense_wgrad<<<grid_for((size_t)A * (2 * N2 + 1)), 256, 0, s>>>(t->dlogit, hap, Gp + pDW, Gp + pDb, B, 2 * N2, A);
  float* dhav = t->dha;
  float* dhap = t->dha + (size_t)B * N2;
  float* dhzv = t->dhz;
  float* dhzp = t->dhz + (size_t)B * N2;
  k_dense_dgrad<<<grid_for((size_t)B * 2 * N2), 256, 0, s>>>(t->dlogit, Pp + pDW, dhap, B, 2 * N2, A);
  // value: Dense(256 -> 1, tanh) (dvpre already holds the gradient before the tanh), Dense(N2 -> 256, relu)
  k_dense_wgrad<<<grid_for((size_t)257), 256, 0, s>>>(t->dvpre, t->hid, Gv + vD2W, Gv + vD2b, B, 256, 1);
  k_dense_dgrad<<<grid_for((size_t)B * 256), 256, 0, s>>>(t->dvpre, Pv + vD2W, t->dhid, B, 256, 1);
  k_relu_mask<<<grid_for((size_t)B * 256), 256, 0, s>>>(t->dhid, t->hid, (size_t)B * 256);
  k_dense_wgrad<<<grid_for((size_t)256 * (N2 + 1)), 256, 0, s>>>(t->dhid, hav, Gv + vD1W, Gv + vD1b, B, N2, 256);
  k_dense_dgrad<<<grid_for((size_t)B * N2), 256, 0, s>>>(t->dhid, Pv + vD1W, dhav, B, N2, 256);
  // BatchNorm + relu of both heads, then the 1x1 convolutions
  k_bn_bwd_reduce<<<1, 256, 0, s>>>(dhav, hav, hzv, t->hmean, t->hinvstd, B, 1, N2, Gv + vbeta, Gv + vgamma);
  k_bn_bwd_apply<<<grid_for((size_t)B * N2), 256, 0, s>>>(dhav, hav, hzv, t->hmean, t->hinvstd, Pv + vgamma, Gv + vbeta, Gv + vgamma, dhzv, nullptr, B, 1, N2);
  k_bn_bwd_reduce<<<2, 256, 0, s>>>(dhap, hap, hzp, t->hmean + 1, t->hinvstd + 1, B, 2, N2, Gp + pbeta, Gp + pgamma);
  k_bn_bwd_apply<<<grid_for((size_t)B * 2 * N2), 256, 0, s>>>(dhap, hap, hzp, t->hmean + 1, t->hinvstd + 1, Pp + pgamma, Gp + pbeta, Gp + pgamma, dhzp, nullptr, B, 2, N2);
  k_channel_sum<<<1, 256, 0, s>>>(dhzv, B, 1, N2, Gv + vb);
  k_channel_sum<<<2, 256, 0, s>>>(dhzp, B, 2, N2, Gp + pb);
  k_conv1x1_wgrad<<<dim3(C, 1), 256, 0, s>>>(trunk, dhzv, Gv + vW, B, C, 1, N2);
  k_conv1x1_wgrad<<<dim3(C, 2), 256, 0, s>>>(trunk, dhzp, Gp + pW, B, C, 2, N2);
  k_conv1x1_dgrad<<<grid_for(act), 256, 0, s>>>(dhzv, Pv + vW, t->dA, B, C, 1, N2, 0);
  k_conv1x1_dgrad<<<grid_for(act), 256, 0, s>>>(dhzp, Pp + pW, t->dA, B, C, 2, N2, 1);

  // ---- backward: tower and stem.  dA = gradient w.r.t. a[l] (the post-activation output of layer l)
  auto conv_bwd = [&](int l, const float* dout, const float* in, float* gshort /* receives dout * relu' */, float* din /* or nullptr */) {
    const ConvOff& o = t->conv[l];
    const float *mean = t->mean + (size_t)l * C, *inv = t->invstd + (size_t)l * C;
    k_bn_bwd_reduce<<<C, 256, 0, s>>>(dout, t->a[l], t->z[l], mean, inv, B, C, N2, Gb + o.beta, Gb + o.gamma);
    k_bn_bwd_apply<<<grid_for(act), 256, 0, s>>>(dout, t->a[l], t->z[l], mean, inv, Pb + o.gamma, Gb + o.beta, Gb + o.gamma, t->dZ, gshort, B, C, N2);
    k_channel_sum<<<C, 256, 0, s>>>(t->dZ, B, C, N2, Gb + o.b);
    k_conv_wgrad<<<dim3((C + WG_CO - 1) / WG_CO, (o.cin + WG_CI - 1) / WG_CI, WG_SPLIT), WG_THREADS, wsm, s>>>(in, t->dZ, t->dwpart, B, o.cin, C, N);
    k_wgrad_reduce<<<grid_for((size_t)9 * o.cin * C), 256, 0, s>>>(t->dwpart, t->dwc, (size_t)9 * o.cin * C, WG_SPLIT);
    k_corr_to_flux<<<grid_for((size_t)9 * o.cin * C), 256, 0, s>>>(t->dwc, Gb + o.W, o.cin, C);
    if (din) {
      k_flux_to_dgrad<<<grid_for((size_t)9 * o.cin * C), 256, 0, s>>>(Pb + o.W, t->wd, o.cin, C);
      // data gradient = the same correlation kernel with in/out channels swapped and the taps reversed
      conv3x3_f32_launch(t->dZ, t->wd, t->ones, t->zeros, nullptr, din, B, C, o.cin, N, 0, s);
    }
  };
  for (int blk = T - 1; blk >= 0; --blk) {
    const int l1 = 1 + 2 * blk, l2 = 2 + 2 * blk, l0 = l1 - 1;
    conv_bwd(l2, t->dA, t->a[l1], t->dS, t->dT);        // dS = gradient through the shortcut, dT = gradient w.r.t. a[l1]
    conv_bwd(l1, t->dT, t->a[l0], nullptr, t->dA);      // dA = gradient w.r.t. a[l0] through the two convolutions
    k_add<<<grid_for(act), 256, 0, s>>>(t->dA, t->dS, act);
  }
  conv_bwd(0, t->dA, t->feats, nullptr, nullptr);

  // ---- data parallel: average gradients, loss terms and running statistics over the ranks
  if (world > 1 && allreduce) {
    const float inv = 1.0f / (float)world;
    const size_t nstat = (size_t)nl * C;
    int arc = 0;
    for (int k = 0; k < 3; ++k) arc |= allreduce(ctx, t->G[k], t->np[k], s);
    arc |= allreduce(ctx, t->red, 8, s);
    arc |= allreduce(ctx, t->mu_run, nstat, s) | allreduce(ctx, t->var_run, nstat, s) | allreduce(ctx, t->hmu_run, 3, s) | allreduce(ctx, t->hvar_run, 3, s);
    if (arc) { snprintf(err, errlen, "gradient all-reduce failed"); return 1; }
    for (int k = 0; k < 3; ++k) k_scale<<<grid_for(t->np[k]), 256, 0, s>>>(t->G[k], t->np[k], inv);
    k_scale<<<1, 8, 0, s>>>(t->red, 8, inv);
    k_scale<<<grid_for(nstat), 256, 0, s>>>(t->mu_run, nstat, inv);
    k_scale<<<grid_for(nstat), 256, 0, s>>>(t->var_run, nstat, inv);
    k_scale<<<1, 3, 0, s>>>(t->hmu_run, 3, inv);
    k_scale<<<1, 3, 0, s>>>(t->hvar_run, 3, inv);
  }
  // ---- loss value, then the Momentum update (train.jl:54)
  float red[8];
  cudaMemcpyAsync(red, t->red, sizeof(red), cudaMemcpyDeviceToHost, s);
  for (int k = 0; k < 3; ++k) k_momentum<<<grid_for(t->np[k]), 256, 0, s>>>(t->P[k], t->G[k], t->V[k], t->np[k], eta, rho);
  cudaError_t rc = cudaStreamSynchronize(s);
  if (rc == cudaSuccess) rc = cudaGetLastError();
  if (rc != cudaSuccess) { snprintf(err, errlen, "train_step: %s", cudaGetErrorString(rc)); return 1; }
  if (loss_out) *loss_out = W_POLICY * red[0] / (float)B + W_VALUE * red[1] / (float)B + W_REG * red[2];
  t->dirty = true;
  return 0;
}

int train_read_grads(TrainState* t, int chain, float* out, size_t n, cudaStream_t s) {
  if (chain < 0 || chain > 2 || n != t->np[chain]) return 1;
  cudaMemcpyAsync(out, t->G[chain], n * sizeof(float), cudaMemcpyDeviceToHost, s);
  return cudaStreamSynchronize(s) == cudaSuccess ? 0 : 1;
}

}  // namespace agz
