// nn_tc.cu -- placeholder until the tcgen05 path lands (next commit): every entry point reports that clearly.
#include <stdio.h>

#include "nn_state.h"
#include "tree.cuh"

namespace agz {
int nn_tc_create(NNet*, char*, size_t) { return 0; }
void nn_tc_destroy(NNet*) {}
int nn_tc_commit(NNet*, const std::vector<ConvLayerHost>&, cudaStream_t, char*, size_t) { return 0; }
TCInput nn_tc_input(NNet*) { TCInput t{}; return t; }
int nn_forward_tc(NNet*, int, float*, float*, cudaStream_t, char* err, size_t errlen) {
  snprintf(err, errlen, "the tcgen05 network path is not built yet");
  return 1;
}
long long nn_tc_launches_per_forward(const NNet*) { return 0; }
int engine_tc_features(const Cfg&, const View&, NNet*, int, int, int, cudaStream_t) { return (int)cudaErrorNotSupported; }
int engine_host_features_tc(const Cfg&, NNet*, const int8_t*, const int8_t*, int, cudaStream_t) { return (int)cudaErrorNotSupported; }
}  // namespace agz
