// replay.cu -- see replay.h.
#include <dlfcn.h>
#include <nccl.h>   // types only: libnccl is resolved at run time (see nccl_api) so that the library loads without it
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "devrt.h"
#include "replay.h"

#define AGZ_HD __host__ __device__

namespace agz {

static const long long kReplayCap = 500000;  // memory_size default (src/train.jl:38)

// NCCL is bound lazily with dlopen: a process that also runs torch.distributed must share torch's bundled
// libnccl.so.2 (a second, older copy in the same process breaks torch), so an already-loaded copy is preferred.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.ok = false;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
      api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.AllReduce && api.CommDestroy && api.GetErrorString;
    }
  }
  return api.ok ? &api : nullptr;
}

struct ReplayState {
  ncclComm_t comm;
  bool have_comm;
  int world, rank;
  size_t stride;            // bytes per packed tuple: pi (4A) | boards[8] (8*N2: the position and the 7 before it) | to_play | z | pad -> multiple of 16
  unsigned char* ring;      // [cap][stride]
  long long cap, total;     // tuples ever appended (ring index = total % cap)
  unsigned long long gather_pos;  // finished-ring records already packed
  unsigned char* send;      // packed local tuples
  unsigned char* recv;      // world * max_count * stride
  size_t send_cap, recv_cap;
  long long* d_counts;      // [world]
  int* d_rec_idx;           // per packed record: ring slot, tuple offset
  size_t rec_cap;
  unsigned char* stage;     // sampled tuples [batch][stride] (replay_sample_device)
  long long* d_idx;         // their ring indices
  size_t stage_cap, idx_cap;
  long long last_payload_bytes, gathered_bytes;   // tuple bytes appended by the last gather (all ranks) / since creation
};

ReplayState* replay_create(const Cfg& c, long long capacity, char* err, size_t errlen) {
  ReplayState* r = new ReplayState();
  memset(r, 0, sizeof(*r));
  r->world = c.world;
  r->rank = c.rank;
  r->stride = ((size_t)4 * c.A + 8 * (size_t)c.N2 + 2 + 15) / 16 * 16;
  r->cap = capacity > 0 ? capacity : kReplayCap;   // option replay.capacity (smaller rings: tests of the trim-oldest wrap-around)
  if (cudaMalloc((void**)&r->ring, (size_t)r->cap * r->stride) != cudaSuccess || cudaMalloc((void**)&r->d_counts, sizeof(long long) * c.world) != cudaSuccess) {
    snprintf(err, errlen, "cudaMalloc of the replay ring failed");
    replay_destroy(r);
    return nullptr;
  }
  return r;
}

void replay_destroy(ReplayState* r) {
  if (!r) return;
  if (r->have_comm && nccl_api()) nccl_api()->CommDestroy(r->comm);
  cudaFree(r->ring); cudaFree(r->send); cudaFree(r->recv); cudaFree(r->d_counts); cudaFree(r->d_rec_idx); cudaFree(r->stage); cudaFree(r->d_idx);
  delete r;
}

int replay_unique_id(uint8_t id_out[128]) {
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (!nccl_api() || nccl_api()->GetUniqueId(&id) != ncclSuccess) return 1;
  memcpy(id_out, &id, 128);
  return 0;
}

int replay_nccl_init(ReplayState* r, const uint8_t idb[128], int world, int rank, char* err, size_t errlen) {
  ncclUniqueId id;
  memcpy(&id, idb, 128);
  if (!nccl_api()) {
    snprintf(err, errlen, "libnccl.so.2 could not be loaded");
    return 1;
  }
  ncclResult_t rc = nccl_api()->CommInitRank(&r->comm, world, id, rank);
  if (rc != ncclSuccess) {
    snprintf(err, errlen, "ncclCommInitRank: %s", nccl_api()->GetErrorString(rc));
    return 1;
  }
  r->have_comm = true;
  r->world = world;
  r->rank = rank;
  return 0;
}

// in-place sum over all ranks (data-parallel training: gradients, loss terms, running statistics)
int replay_world(const ReplayState* r) { return (r && r->have_comm) ? r->world : 1; }

int replay_allreduce_sum(ReplayState* r, float* buf, size_t n, cudaStream_t s) {
  if (!r || !r->have_comm || r->world == 1) return 0;
  ncclResult_t nr = nccl_api()->AllReduce(buf, buf, n, ncclFloat32, ncclSum, r->comm, s);
  return nr == ncclSuccess ? 0 : 1;
}

// One warp per finished game: replay its moves from the empty board (replay_position, board.jl:557-578) and emit,
// for every ply, the packed tuple (pi | board before the move and the 7 boards before that, oldest repeated = what get_feats
// rebuilds from board_deltas, features.jl:7-14 | to_play | z = final result from Black's view).
struct PackOp {
  Cfg c;
  View v;
  const int* rec;  // [n][2]: finished-ring slot, first tuple index
  unsigned char* out;
  size_t stride;
  __device__ void operator()(int wi, char* smem) const {
    (void)smem;
    const int lane = threadIdx.x & 31;
    const BitsCtx B = bits_ctx(c.N, c.KB);   // the game's rules on bitboard lines in registers (go_bits.cuh, tree.cuh game_play)
    Lines pos;
    pos.b = 0;
    pos.w = 0;
    const int rslot = rec[2 * wi], first = rec[2 * wi + 1];
    const RingHeader hd = v.ring_hdr[rslot];
    const size_t L = c.max_game_length + 2;
    int to_play = 1;
    for (int t = 0; t < hd.n_moves; ++t) {
      unsigned char* o = out + (size_t)(first + t) * stride;
      float* opi = reinterpret_cast<float*>(o);
      const float* pi = v.ring_pi + ((size_t)rslot * L + t) * c.A;
      for (int a = lane; a < c.A; a += 32) opi[a] = pi[a];
      int8_t* ob = reinterpret_cast<int8_t*>(o + (size_t)4 * c.A);
      bits_to_bytes(B, pos, ob);
      // boards 1..7 = the previous tuple's boards 0..6 (all empty before the first move)
      const int8_t* prev = t > 0 ? reinterpret_cast<const int8_t*>(o - stride + (size_t)4 * c.A) : nullptr;
      for (int p = lane; p < 7 * c.N2; p += 32) ob[c.N2 + p] = prev ? prev[p] : (int8_t)0;
      if (lane == 0) {
        ob[8 * c.N2] = (int8_t)to_play;
        ob[8 * c.N2 + 1] = (int8_t)hd.result;
      }
      __syncwarp();
      const int mv = v.ring_moves[(size_t)rslot * L + t];
      if (mv != c.pass) {
        int ko, ncap;
        bool ended;
        game_play(c, B, pos, mv, to_play, false, ko, ncap, ended);
      }
      to_play = -to_play;
      __syncwarp();
    }
  }
};

static int grow(unsigned char** p, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  cudaFree(*p);
  *p = nullptr;
  *cap = 0;   // a failed allocation must not leave a stale capacity behind a null pointer
  size_t n = need + need / 2 + 4096;
  if (cudaMalloc((void**)p, n) != cudaSuccess) return 1;
  cudaMemset(*p, 0, n);   // tuples are padded to 16 bytes: the padding travels with every copy and must not be uninitialised memory
  *cap = n;
  return 0;
}

static void ring_append(ReplayState* r, const unsigned char* src, long long n, cudaStream_t s) {
  while (n > 0) {  // trim-oldest ring (train.jl:52,63-65)
    long long pos = r->total % r->cap;
    long long run = n < r->cap - pos ? n : r->cap - pos;
    cudaMemcpyAsync(r->ring + (size_t)pos * r->stride, src, (size_t)run * r->stride, cudaMemcpyDeviceToDevice, s);
    src += (size_t)run * r->stride;
    r->total += run;
    n -= run;
  }
}

int replay_gather(ReplayState* r, const Cfg& c, const View& v, int smem_per_warp, cudaStream_t s, int64_t* n_total, long long* launches,
                  char* err, size_t errlen) {
  *launches = 0;
  r->last_payload_bytes = 0;
  cudaStreamSynchronize(s);
  unsigned long long ctr[CTR_COUNT];
  cudaMemcpy(ctr, v.ctr, sizeof(ctr), cudaMemcpyDeviceToHost);
  unsigned long long head = ctr[CTR_RING_HEAD], tail = ctr[CTR_RING_TAIL];
  if (r->gather_pos < head) r->gather_pos = head;  // records released before being gathered are gone
  // headers of all finished records not gathered yet: the ring is fixed-stride, so they are at most two contiguous ranges
  const unsigned long long npend = tail - r->gather_pos;
  std::vector<RingHeader> hd((size_t)npend);
  for (unsigned long long done = 0; done < npend;) {
    const size_t rs = (size_t)((r->gather_pos + done) % (unsigned long long)c.ring_cap);
    const size_t run = (size_t)std::min<unsigned long long>(npend - done, (unsigned long long)c.ring_cap - rs);
    cudaMemcpy(hd.data() + done, v.ring_hdr + rs, run * sizeof(RingHeader), cudaMemcpyDeviceToHost);
    done += run;
  }
  std::vector<int> rec;
  rec.reserve(2 * (size_t)npend);
  long long n_local = 0;
  for (unsigned long long q = 0; q < npend; ++q) {
    rec.push_back((int)((r->gather_pos + q) % (unsigned long long)c.ring_cap));
    rec.push_back((int)n_local);
    n_local += hd[(size_t)q].n_moves;
  }
  r->gather_pos = tail;
  const int nrec = (int)npend;
  const bool multi = r->have_comm && r->world > 1;
  // Ragged all-gather, step 1: the counts.  They go round BEFORE anything is packed, so the send buffer can be sized once for
  // max(own, largest) tuples -- growing it after the pack kernel would discard this rank's tuples.
  std::vector<long long> counts((size_t)(multi ? r->world : 1), n_local);
  long long mx = n_local;
  if (multi) {
    cudaMemcpyAsync(r->d_counts + r->rank, &n_local, sizeof(long long), cudaMemcpyHostToDevice, s);
    ncclResult_t nr = nccl_api()->AllGather(r->d_counts + r->rank, r->d_counts, 1, ncclInt64, r->comm, s);
    if (nr != ncclSuccess) { snprintf(err, errlen, "ncclAllGather(counts): %s", nccl_api()->GetErrorString(nr)); return 4; }
    cudaMemcpyAsync(counts.data(), r->d_counts, sizeof(long long) * r->world, cudaMemcpyDeviceToHost, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) { snprintf(err, errlen, "replay gather (counts): %s", cudaGetErrorString(cudaGetLastError())); return 3; }
    for (long long x : counts) mx = x > mx ? x : mx;
  }
  if (grow(&r->send, &r->send_cap, (size_t)(mx ? mx : 1) * r->stride)) { snprintf(err, errlen, "replay send buffer allocation failed"); return 3; }
  if (nrec) {
    if (rec.size() * sizeof(int) > r->rec_cap) {
      cudaFree(r->d_rec_idx);
      r->d_rec_idx = nullptr;
      r->rec_cap = 0;
      if (cudaMalloc((void**)&r->d_rec_idx, rec.size() * sizeof(int) * 2) != cudaSuccess) { snprintf(err, errlen, "replay index allocation failed"); return 3; }
      r->rec_cap = rec.size() * sizeof(int) * 2;
    }
    cudaMemcpyAsync(r->d_rec_idx, rec.data(), rec.size() * sizeof(int), cudaMemcpyHostToDevice, s);
    PackOp op{c, v, r->d_rec_idx, r->send, r->stride};
    int rc = devrt::launch_warps(op, nrec, smem_per_warp, s);
    if (rc) { snprintf(err, errlen, "pack kernel: %s", cudaGetErrorString((cudaError_t)rc)); return 3; }
    *launches += 1;
  }
  if (!multi) {
    ring_append(r, r->send, n_local, s);
    r->last_payload_bytes = (long long)n_local * (long long)r->stride;
  } else if (mx > 0) {
    // step 2: fixed-stride padded blocks (every rank sends mx tuples' worth; only counts[k] of block k are appended)
    if (grow(&r->recv, &r->recv_cap, (size_t)mx * r->stride * r->world)) { snprintf(err, errlen, "replay recv buffer allocation failed"); return 3; }
    ncclResult_t nr = nccl_api()->AllGather(r->send, r->recv, (size_t)mx * r->stride, ncclUint8, r->comm, s);
    if (nr != ncclSuccess) { snprintf(err, errlen, "ncclAllGather(tuples): %s", nccl_api()->GetErrorString(nr)); return 4; }
    for (int k = 0; k < r->world; ++k) {
      ring_append(r, r->recv + (size_t)k * mx * r->stride, counts[(size_t)k], s);
      r->last_payload_bytes += counts[(size_t)k] * (long long)r->stride;
    }
  }
  if (cudaStreamSynchronize(s) != cudaSuccess) { snprintf(err, errlen, "replay gather: %s", cudaGetErrorString(cudaGetLastError())); return 3; }
  r->gathered_bytes += r->last_payload_bytes;
  if (n_total) *n_total = r->total;
  return 0;
}

// copy tuples out of a host image of `count` packed tuples
static void unpack(const ReplayState* r, const Cfg& c, const unsigned char* img, int count, int8_t* boards, int8_t* to_play, float* pis, int8_t* zs,
                   int8_t* boards_hist) {
  for (int i = 0; i < count; ++i) {
    const unsigned char* t = img + (size_t)i * r->stride;
    if (pis) memcpy(pis + (size_t)i * c.A, t, (size_t)4 * c.A);
    const int8_t* b = reinterpret_cast<const int8_t*>(t + (size_t)4 * c.A);
    if (boards) memcpy(boards + (size_t)i * c.N2, b, (size_t)c.N2);
    if (boards_hist) memcpy(boards_hist + (size_t)i * 8 * c.N2, b, (size_t)8 * c.N2);
    if (to_play) to_play[i] = b[8 * c.N2];
    if (zs) zs[i] = b[8 * c.N2 + 1];
  }
}

int replay_read(ReplayState* r, const Cfg& c, int64_t first, int32_t count, int8_t* boards, int8_t* to_play, float* pis, int8_t* zs,
                cudaStream_t s, char* err, size_t errlen, int8_t* boards_hist) {
  long long oldest = r->total > r->cap ? r->total - r->cap : 0;
  if (first < oldest || first + count > r->total || count < 0) {
    snprintf(err, errlen, "replay tuples [%lld, %lld) not in the ring [%lld, %lld)", (long long)first, (long long)(first + count), oldest, r->total);
    return 5;
  }
  if (count == 0) return 0;
  std::vector<unsigned char> img((size_t)count * r->stride);
  for (long long done = 0; done < count;) {   // at most two contiguous ranges of the ring
    const long long pos = (first + done) % r->cap;
    const long long run = std::min<long long>(count - done, r->cap - pos);
    cudaMemcpyAsync(img.data() + (size_t)done * r->stride, r->ring + (size_t)pos * r->stride, (size_t)run * r->stride, cudaMemcpyDeviceToHost, s);
    done += run;
  }
  if (cudaStreamSynchronize(s) != cudaSuccess) { snprintf(err, errlen, "replay read: %s", cudaGetErrorString(cudaGetLastError())); return 3; }
  unpack(r, c, img.data(), count, boards, to_play, pis, zs, boards_hist);
  return 0;
}

// ---- get_replay_batch (src/train.jl:4-12): uniform sample without replacement, drawn ON THE DEVICE.
// Draw k of a batch is tuple oldest + perm(k), where perm is a keyed pseudo-random permutation of [0, n): a 6-round Feistel
// network over 2w >= log2(n) bits with cycle walking (values >= n are fed back until they land inside).  A permutation gives
// distinct indices by construction, every draw is independent of the others (one CTA per draw: compute the index, copy the
// tuple), and nothing O(ring) is ever built.  Restated in oracle/replay.py (sample_indices); the reference's own draw comes
// from StatsBase.sample on Julia's global RNG and is unpinned.
AGZ_HD inline uint32_t mix32(uint32_t h) {   // murmur3 finaliser
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
struct FeistelKeys { uint32_t k[6]; int w; };
static FeistelKeys feistel_keys(uint64_t seed, long long n) {
  FeistelKeys f;
  int bits = 2;
  while ((1LL << bits) < n) ++bits;
  if (bits & 1) ++bits;
  f.w = bits / 2;
  uint64_t x = seed;
  for (int i = 0; i < 6; ++i) {   // splitmix64 stream
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    f.k[i] = (uint32_t)((z ^ (z >> 31)) >> 32);
  }
  return f;
}
AGZ_HD inline long long feistel_perm(const FeistelKeys& f, long long k, long long n) {
  const uint32_t mask = (1u << f.w) - 1u;
  long long y = k;
  do {
    uint32_t L = (uint32_t)(y >> f.w), R = (uint32_t)y & mask;
    for (int i = 0; i < 6; ++i) {
      const uint32_t F = mix32(R ^ f.k[i]) & mask;
      const uint32_t t = L ^ F;
      L = R;
      R = t;
    }
    y = ((long long)L << f.w) | R;
  } while (y >= n);
  return y;
}

__global__ void __launch_bounds__(128) replay_sample_kernel(const unsigned char* __restrict__ ring, long long cap, long long oldest, long long n,
                                                            FeistelKeys f, size_t stride, unsigned char* __restrict__ out, long long* __restrict__ idx_out) {
  const int k = blockIdx.x;
  const long long idx = oldest + feistel_perm(f, k, n);
  if (threadIdx.x == 0) idx_out[k] = idx;
  const uint4* src = reinterpret_cast<const uint4*>(ring + (size_t)(idx % cap) * stride);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)k * stride);
  for (size_t i = threadIdx.x; i < stride / 16; i += blockDim.x) dst[i] = src[i];
}

// draws `batch` tuples into the device staging buffer (r->stage: [batch][stride], r->d_idx: [batch]); no host synchronisation
int replay_sample_device(ReplayState* r, int32_t batch, uint64_t seed, cudaStream_t s, const unsigned char** stage_out, const long long** idx_out,
                         char* err, size_t errlen) {
  const long long oldest = r->total > r->cap ? r->total - r->cap : 0, n = r->total - oldest;
  if (batch < 1 || batch > n) {
    snprintf(err, errlen, "cannot sample %d tuples without replacement from %lld", batch, n);
    return 5;
  }
  if (grow(&r->stage, &r->stage_cap, (size_t)batch * r->stride) || grow((unsigned char**)&r->d_idx, &r->idx_cap, (size_t)batch * sizeof(long long))) {
    snprintf(err, errlen, "replay sample staging allocation failed");
    return 3;
  }
  replay_sample_kernel<<<batch, 128, 0, s>>>(r->ring, r->cap, oldest, n, feistel_keys(seed, n), r->stride, r->stage, r->d_idx);
  if (cudaGetLastError() != cudaSuccess) { snprintf(err, errlen, "replay sample kernel launch failed"); return 3; }
  if (stage_out) *stage_out = r->stage;
  if (idx_out) *idx_out = r->d_idx;
  return 0;
}

int replay_sample(ReplayState* r, const Cfg& c, int32_t batch, uint64_t seed, int8_t* boards, int8_t* to_play, float* pis, int8_t* zs, int64_t* indices,
                  cudaStream_t s, char* err, size_t errlen, int8_t* boards_hist) {
  if (batch == 0) return 0;
  int rc = replay_sample_device(r, batch, seed, s, nullptr, nullptr, err, errlen);
  if (rc) return rc;
  std::vector<unsigned char> img((size_t)batch * r->stride);   // one transfer for the tuples, one for the indices
  std::vector<long long> idx((size_t)batch);
  cudaMemcpyAsync(img.data(), r->stage, img.size(), cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(idx.data(), r->d_idx, idx.size() * sizeof(long long), cudaMemcpyDeviceToHost, s);
  if (cudaStreamSynchronize(s) != cudaSuccess) { snprintf(err, errlen, "replay sample: %s", cudaGetErrorString(cudaGetLastError())); return 3; }
  unpack(r, c, img.data(), batch, boards, to_play, pis, zs, boards_hist);
  if (indices) for (int k = 0; k < batch; ++k) indices[k] = idx[(size_t)k];
  return 0;
}

void replay_info(const ReplayState* r, int64_t out[5]) {
  out[0] = (int64_t)r->stride; out[1] = r->cap; out[2] = r->total; out[3] = r->last_payload_bytes; out[4] = r->gathered_bytes;
}
long long replay_capacity(const ReplayState* r) { return r->cap; }
long long replay_default_capacity() { return kReplayCap; }
size_t replay_stride(const ReplayState* r) { return r->stride; }

}  // namespace agz
