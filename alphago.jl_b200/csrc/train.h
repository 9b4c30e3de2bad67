// train.h -- one optimisation step of the network on the device (train.cu): _train / losses / Momentum of
// src/neural_net.jl:75-101 and src/train.jl:54.
#pragma once
#include <stddef.h>

#include <cuda_runtime.h>

namespace agz {

struct NNet;
struct TrainState;

TrainState* train_create(const NNet* n, int max_batch, char* err, size_t errlen);
void train_destroy(TrainState* t);
int train_max_batch(const TrainState* t);
int train_load(TrainState* t, const NNet* n, cudaStream_t s, char* err, size_t errlen);   // host parameters -> device master copy
int train_store(TrainState* t, NNet* n, cudaStream_t s, char* err, size_t errlen);        // device master copy -> host copy of the parameters (lazy: agz_net_get_params ...)
int train_publish(TrainState* t, NNet* n, cudaStream_t s, char* err, size_t errlen);      // device master copy -> folded / reordered inference weights, all on the device
bool train_dirty(const TrainState* t);
// feats: [B][17][N2] fp32 on the device; pi [B][A], z [B] on the host.  Returns the loss of the batch before the update.
// world > 1 (data parallel, an extension: the reference trains in one process): every rank passes its own minibatch; gradients,
// loss terms and the BatchNorm running statistics are summed over the ranks with `allreduce` and divided by `world` before the
// update, so all ranks hold identical parameters afterwards (batch statistics stay per rank, as in plain data-parallel BatchNorm).
typedef int (*train_allreduce_fn)(void* ctx, float* buf, size_t n, cudaStream_t s);
int train_step(TrainState* t, const float* d_feats, const float* h_pi, const float* h_z, int B, float eta, float rho, float* loss_out,
               cudaStream_t s, char* err, size_t errlen, int world = 1, train_allreduce_fn allreduce = nullptr, void* ctx = nullptr);
// same step with the minibatch already on the device as packed replay tuples (replay.cu: replay_sample_device)
int train_step_from_tuples(TrainState* t, const unsigned char* d_stage, size_t stride, int B, float eta, float rho, float* loss_out, cudaStream_t s,
                           char* err, size_t errlen, int world = 1, train_allreduce_fn allreduce = nullptr, void* ctx = nullptr);
int train_read_grads(TrainState* t, int chain, float* out, size_t n, cudaStream_t s);      // gradients of the last step (data loss only)

}  // namespace agz
