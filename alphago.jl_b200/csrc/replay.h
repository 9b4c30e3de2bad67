// replay.h -- K9: finished-game tuple packing + NCCL all-gather into the device-resident replay ring.
// Stands for extract_data (src/mcts_play.jl:126-139) -> replay_position (src/game/go/board.jl:557-578) and the
// buffer append / trim of src/train.jl:58-65.  Games never exchange data during search; this is the only
// collective of the path.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "tree.cuh"

namespace agz {
struct ReplayState;
ReplayState* replay_create(const Cfg& c, long long capacity /* 0 = memory_size default, src/train.jl:38 */, char* err, size_t errlen);
void replay_destroy(ReplayState* r);
int replay_unique_id(uint8_t id_out[128]);
int replay_nccl_init(ReplayState* r, const uint8_t id[128], int world, int rank, char* err, size_t errlen);
// packs the finished-ring records not gathered yet, all-gathers them, appends to the ring (trim-oldest)
int replay_gather(ReplayState* r, const Cfg& c, const View& v, int smem_per_warp, cudaStream_t s, int64_t* n_total, long long* launches,
                  char* err, size_t errlen);
int replay_read(ReplayState* r, const Cfg& c, int64_t first, int32_t count, int8_t* boards, int8_t* to_play, float* pis, int8_t* zs,
                cudaStream_t s, char* err, size_t errlen, int8_t* boards_hist = nullptr /* count x 8 x N2: the position and the 7 before it */);

int replay_sample(ReplayState* r, const Cfg& c, int32_t batch, uint64_t seed, int8_t* boards, int8_t* to_play, float* pis, int8_t* zs, int64_t* indices,
                  cudaStream_t s, char* err, size_t errlen, int8_t* boards_hist = nullptr);

// the same draw left on the device: stage = [batch][stride] packed tuples (pi | 8 boards | to_play | z), idx = their ring indices
int replay_sample_device(ReplayState* r, int32_t batch, uint64_t seed, cudaStream_t s, const unsigned char** stage_out, const long long** idx_out,
                         char* err, size_t errlen);
// out: [0] bytes per packed tuple, [1] ring capacity, [2] tuples ever appended, [3] tuple bytes appended by the last gather
// (all ranks' payload), [4] tuple bytes appended since creation
void replay_info(const ReplayState* r, int64_t out[5]);
long long replay_capacity(const ReplayState* r);
long long replay_default_capacity();
size_t replay_stride(const ReplayState* r);

// data-parallel training: world size of the initialised communicator (1 without one) and an in-place float sum over the ranks
int replay_world(const ReplayState* r);
int replay_allreduce_sum(ReplayState* r, float* buf, size_t n, cudaStream_t s);

// feature kernels over caller-supplied positions (agz_features / agz_net_forward), features.cu
int engine_host_features(const Cfg& c, const int8_t* boards_hist, const int8_t* to_play, int B, float* out_host, float* out_dev, cudaStream_t s);
}  // namespace agz
