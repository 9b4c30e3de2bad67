// nn.h -- policy/value network of the engine (src/neural_net.jl:13-33,57-68; src/resnet.jl:2-32).
// Two device implementations of the same network:
//   NN_F32: fp32 SIMT kernels (cross-check path),
//   NN_TC : TMA-staged tcgen05 implicit-GEMM 3x3 convolutions, fp16 operands / fp32 accumulation in TMEM,
//           fused bias+BatchNorm+ReLU(+residual) epilogue, fused heads.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <cuda_runtime.h>

namespace agz {

struct NNet;

struct NNShape {
  int N, planes, filters, tower;
  int actions;   // outputs of the policy head = env.action_space (neural_net.jl:30): N^2 + 1 for Go, N^2 for Gomoku; 0 = N^2 + 1
};
inline int nn_actions(const NNShape& s) { return s.actions > 0 ? s.actions : s.N * s.N + 1; }

NNet* nn_create(const NNShape& s, int max_batch, char* err, size_t errlen);
void nn_destroy(NNet* n);
size_t nn_param_count(const NNet* n, int chain);
size_t nn_bn_count(const NNet* n, int chain);  // BatchNorm channels of the chain (mu / sigma are this long each)
int nn_set_params(NNet* n, int chain, const float* flat, size_t cnt);
int nn_set_bn(NNet* n, int chain, const float* mu, const float* sigma, size_t cnt, int mode);
// fold BatchNorm, reorder weights for both paths and upload; called lazily before the first forward
int nn_commit(NNet* n, cudaStream_t s, char* err, size_t errlen);
bool nn_ready(const NNet* n);

// Input of both paths: packed history planes written by the feature kernels.
//   feats_f32 : [B][17][N2] float, reference (W x H x C x B) order               (NN_F32)
//   feats_tc  : dense fp16 rows [B*N^2][64], written by engine_tc_features / engine_host_features_tc (NN_TC)
// Output: pi [B][A] float (softmax over all A actions, no legality masking), v [B] float (Black's view).
// Debug outputs of a forward pass (agz_net_forward_debug, test hook): stop after `n_blocks` residual blocks (< 0 or >= tower: the
// whole network), the trunk at that point as fp32 [B][C][N2] (reference layout), and the heads' values before softmax / tanh.
struct NNDebug {
  int n_blocks;
  float* trunk;   // device [B][C][N2] or nullptr
  float* raw;     // device [B][A + 1]: A logits then the value before tanh, or nullptr
};
int nn_forward_f32(NNet* n, const float* feats_f32, int B, float* pi, float* v, cudaStream_t s, cudaEvent_t* ev = nullptr, const NNDebug* dbg = nullptr);

// The tensor-core path (nn_tc.cu) consumes activations as dense fp16 NHWC rows (row = b*N^2 + N*j + i); the feature kernels
// below write the stem input (64 channels per row, 17 used) directly.
// ev (optional): 4 events recorded before the stem, after the stem, after the tower, after the heads
// group >= 0: evaluate only half batch `group` (rows [group*max_batch/2, ...)); pi / v point at that half's first row
int nn_forward_tc(NNet* n, int B, float* pi, float* v, cudaStream_t s, char* err, size_t errlen, cudaEvent_t* ev = nullptr, int group = -1,
                  cudaEvent_t convs_done = nullptr /* recorded after the last convolution, before the heads */,
                  cudaStream_t heads_stream = nullptr /* with convs_done: launch the heads there instead of on s */, const NNDebug* dbg = nullptr);
int nn_tc_groups(const NNet* n);
void nn_tc_set_trace(NNet* n, unsigned long long* trace);   // kernel timeline trace buffer (simt.h), nullptr = off
// agz_set_option / agz_get_option keys "conv.*": 0 ok, 1 unknown key, 2 bad value
int nn_tc_set_option(NNet* n, const char* key, long long value);
int nn_tc_get_option(const NNet* n, const char* key, long long* value);
// feature kernels that write the tensor-core input directly (nn_tc.cu): from the leaves of the current round
// (batch rows [0, row0+nrows)), and from caller-supplied positions.
struct Cfg;
struct View;
int engine_tc_features(const Cfg& c, const View& v, NNet* n, int row0, int nrows, int smem_per_warp, cudaStream_t s);
int engine_host_features_tc(const Cfg& c, NNet* n, const int8_t* boards_hist, const int8_t* to_play, int B, cudaStream_t s);
// 2*MACs of one position through the network (stem + tower + heads), for the roofline
double nn_flops_per_position(const NNShape& s);
long long nn_tc_launches_per_forward(const NNet* n);
long long nn_f32_launches_per_forward(const NNet* n);

}  // namespace agz
