// features.cu -- get_feats (src/features.jl:3-26) for caller-supplied positions (agz_features / agz_net_forward).
// Input: boards_hist[b][k][p] int8 = the board k moves ago (what the reference reconstructs from board_deltas,
// features.jl:7-14), to_play[b].  Output (reference layout N x N x 17 x B, row fastest): out[b][c][p].
#include "replay.h"

namespace agz {

__global__ void host_features_kernel(const int8_t* __restrict__ bh, const int8_t* __restrict__ tp, float* __restrict__ out, int B, int N2) {
  const size_t total = (size_t)B * 17 * N2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx % N2);
    const int c = (int)((idx / N2) % 17);
    const int b = (int)(idx / ((size_t)17 * N2));
    const int t = tp[b];
    float val;
    if (c == 16) val = (float)t;                                  // colour plane is +-1 (features.jl:22)
    else {
      const int s = bh[((size_t)b * 8 + (c >> 1)) * N2 + p];
      val = (s == ((c & 1) ? -t : t)) ? 1.f : 0.f;
    }
    out[idx] = val;
  }
}

int engine_host_features(const Cfg& c, const int8_t* boards_hist, const int8_t* to_play, int B, float* out_host, float* out_dev, cudaStream_t s) {
  int8_t *dbh = nullptr, *dtp = nullptr;
  float* dout = out_dev;
  const size_t nb = (size_t)B * 8 * c.N2, no = (size_t)B * 17 * c.N2;
  cudaError_t rc = cudaMalloc((void**)&dbh, nb);
  if (rc == cudaSuccess) rc = cudaMalloc((void**)&dtp, (size_t)B);
  if (rc == cudaSuccess && !dout) rc = cudaMalloc((void**)&dout, no * sizeof(float));
  if (rc == cudaSuccess) rc = cudaMemcpyAsync(dbh, boards_hist, nb, cudaMemcpyHostToDevice, s);
  if (rc == cudaSuccess) rc = cudaMemcpyAsync(dtp, to_play, (size_t)B, cudaMemcpyHostToDevice, s);
  if (rc == cudaSuccess) {
    int blocks = (int)((no + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    host_features_kernel<<<blocks, 256, 0, s>>>(dbh, dtp, dout, B, c.N2);
    rc = cudaGetLastError();
  }
  if (rc == cudaSuccess && out_host) rc = cudaMemcpyAsync(out_host, dout, no * sizeof(float), cudaMemcpyDeviceToHost, s);
  cudaError_t rs = cudaStreamSynchronize(s);
  if (rc == cudaSuccess) rc = rs;
  cudaFree(dbh);
  cudaFree(dtp);
  if (!out_dev) cudaFree(dout);
  return (int)rc;
}

}  // namespace agz
