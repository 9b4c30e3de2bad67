// nn_state.h -- internal state of the policy/value network shared by nn_f32.cu (SIMT fp32 path) and
// nn_tc.cu (tcgen05 path).  Parameter order = Flux `params` of the reference chains (src/neural_net.jl:16-30,
// src/resnet.jl:3-5, src/train.jl:27-33).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <vector>

#include "nn.h"

namespace agz {

struct ConvLayerHost {  // one 3x3 conv + BatchNorm, BN folded: y = scale * conv(x) + shift
  std::vector<float> w;      // Flux layout (3, 3, Cin, Cout) column-major: w[a + 3*b + 9*ci + 9*Cin*co]
  std::vector<float> scale;  // [Cout]
  std::vector<float> shift;  // [Cout]
  int cin, cout;
};

struct NNet {
  NNShape s;
  int N2, A, max_batch, C;
  std::vector<float> hparams[3], hmu[3], hsigma[3];
  int bn_mode[3];
  bool have[3];
  bool ready;
  bool f32_weights_ready;

  // ---- fp32 path
  std::vector<float*> f_w, f_scale, f_shift;  // per conv layer (stem, then W1, W2 per block)
  float *f_vw, *f_pw;                          // 1x1 conv weights [C], [2][C]
  float f_head_aff[6];                         // v scale, v shift, p0 scale, p0 shift, p1 scale, p1 shift
  float *f_head_aff_d;
  float *f_D1W, *f_D1b, *f_D2W, *f_D2b, *f_PW, *f_Pb;  // Flux Dense weights, column-major (out, in)
  float* f_act[3];                             // [max_batch][C][N2]

  // ---- tensor-core path (nn_tc.cu)
  void* tc;  // TCState*
};

// shared helpers (nn_f32.cu)
int nn_fold_layers(const NNet* n, std::vector<ConvLayerHost>& convs, char* err, size_t errlen, bool with_weights);
int nn_tc_create(NNet* n, char* err, size_t errlen);
void nn_tc_destroy(NNet* n);
int nn_tc_commit(NNet* n, const std::vector<ConvLayerHost>& convs, cudaStream_t s, char* err, size_t errlen);
int nn_tc_commit_device(NNet* n, const float* d_base, cudaStream_t s, char* err, size_t errlen);   // base-chain parameters already on the device

}  // namespace agz
