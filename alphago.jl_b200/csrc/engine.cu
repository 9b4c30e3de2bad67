// engine.cu -- host side of libagz: owns device memory, launches the warp-per-game kernels of ops.cuh and the
// network kernels, and exports the C ABI of include/agz.h.  There is no CPU execution path in this library.
// (The same file is compiled with -DAGZ_EMU by tests/emu to run the tree kernels on the fiber emulator; the
// network and NCCL parts are compiled out there.)
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/agz.h"
#include "devrt.h"
#include "ops.cuh"
#if AGZ_CUDA
#include "nn.h"
#include "nn_state.h"
#include "replay.h"
#include "train.h"
#endif

using namespace agz;

static thread_local char g_err[512] = "";

struct agz_engine {
  Cfg c;
  View v;
  agz_config cfg;
  devrt::stream_t stream;
  int smem_per_warp;
  std::vector<void*> allocs;
  char err[512];
  int evaluator;
  // evaluator buffers
  float* d_dummy_pi;
  float* d_dummy_v;
  float* d_eval_pi;  // [n_games*pmax][A]
  float* d_eval_v;
  float* d_feats_f32;  // [n_games*pmax][17][N2]
  // hook scratch
  int* d_hook_result;
  float* d_hook_probs;
  agz_position *d_pos_in, *d_pos_out;
  int8_t* d_hook_legal;
  uint8_t* d_hook_libs;
  float* d_hook_feats;
  long long launches;
  int32_t *d_match_i, *d_match_j, *d_match_in;   // two-player match scratch (agz_match_*), allocated on first use
  float* d_match_f;
  uint8_t* d_match_active;
  long long* d_match_ids;
  bool train_loaded;
  bool host_stale;           // the device master copy (training) is newer than the host copy of the parameters kept in NNet
  size_t bytes_per_node;
  unsigned long long* d_trace;   // AGZ_TRACE=<records>: kernel timeline trace (simt.h), read back with agz_trace_read
  int trace_cap;
  bool started;
  unsigned long long ring_head;
  int timing;
  float phase_ms[AGZ_NKERNELS];
  long long phase_launches[AGZ_NKERNELS];
#if AGZ_CUDA
  NNet* nn;
  TrainState* train;         // fp32 master parameters + momentum on the device, created by the first agz_train_step
  cudaEvent_t ev[8];   // 0 select | 1 features | 2 stem | 3 tower | 4 heads | 5 incorporate | 6 end
  cudaEvent_t ev_step[2];
  cudaStream_t gstream[2];   // half-batch pipelining: group g's heads -> incorporate -> select -> features chain (low priority)
  cudaStream_t cstream;      // ... and the convolutions of both groups, alternating (high priority)
  cudaEvent_t ev_join[3];
  cudaEvent_t ev_feat[2];    // group g's leaf features are written: its stem may start
  cudaEvent_t ev_conv[2];    // group g's last convolution is done: its heads may start
  int pipeline;              // 1: two half batches on two streams (tree work of one hides under the network of the other)
  ReplayState* replay;
#endif
  int fuse_dummy;            // option dummy.fused_rounds
  long long replay_cap;      // option replay.capacity (0 = memory_size default of src/train.jl:38)
};

static int fail(agz_engine* e, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  char buf[512];
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  snprintf(g_err, sizeof(g_err), "%s", buf);
  if (e) snprintf(e->err, sizeof(e->err), "%s", buf);
  return code;
}

#define DCHECK(e, call)                                                                                  \
  do {                                                                                                   \
    int rc__ = (call);                                                                                   \
    if (rc__ != 0) return fail((e), AGZ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, devrt::last_error_string(rc__), __FILE__, __LINE__); \
  } while (0)

#define DISPATCH_KA(e, ...)                                      \
  switch ((e)->c.KA) {                                           \
    case 3: { constexpr int KA = 3; __VA_ARGS__; } break;        \
    case 6: { constexpr int KA = 6; __VA_ARGS__; } break;        \
    default: { constexpr int KA = 12; __VA_ARGS__; } break;      \
  }

template <class T>
static int dalloc(agz_engine* e, T** p, size_t count) {
  void* q = nullptr;
  int rc = devrt::dmalloc(&q, count * sizeof(T));
  if (rc) return rc;
  e->allocs.push_back(q);
  *p = (T*)q;
  return devrt::dmemset(q, 0, count * sizeof(T), e->stream);
}

extern "C" int32_t agz_version(void) { return 100; }

extern "C" const char* agz_last_error(agz_engine* e) { return e ? e->err : g_err; }

extern "C" int32_t agz_config_default(agz_config* cfg, int32_t board_n) { return agz_config_default_game(cfg, AGZ_GAME_GO, board_n, 0); }

extern "C" int32_t agz_config_default_game(agz_config* cfg, int32_t game, int32_t board_n, int32_t n_in_row) {
  if (!cfg || board_n < 2 || board_n > AGZ_MAX_N) return fail(nullptr, AGZ_ERR_ARG, "board_n must be in [2, %d]", AGZ_MAX_N);
  if (game != AGZ_GAME_GO && game != AGZ_GAME_GOMOKU) return fail(nullptr, AGZ_ERR_ARG, "unknown game %d", game);
  if (game == AGZ_GAME_GOMOKU && (n_in_row < 2 || n_in_row > board_n)) return fail(nullptr, AGZ_ERR_ARG, "n_in_row must be in [2, board_n]");
  memset(cfg, 0, sizeof(*cfg));
  const int N = board_n, A = N * N + (game == AGZ_GAME_GO ? 1 : 0);   // env.action_space (go.jl:12, gomoku.jl:12)
  cfg->game = game;
  cfg->n_in_row = game == AGZ_GAME_GOMOKU ? n_in_row : 0;
  cfg->board_n = N;
  cfg->planes = 17;
  cfg->filters = 256;
  cfg->tower_height = 19;                                   // neural_net.jl:13
  cfg->c_puct = 0.96;                                       // mcts.jl:11
  cfg->noise_weight = 0.25;                                 // mcts.jl:13
  cfg->noise_alpha = (double)(float)(0.03 * 361.0 / A);     // mcts.jl:22 (stored as Float32)
  cfg->max_game_length = (N * N * 7) / 5;                   // mcts.jl:21
  cfg->tau_threshold = (N * N / 12) / 2 * 2;                // mcts_play.jl:19
  cfg->parallel_readouts = 8;                               // mcts_play.jl:73
  cfg->max_parallel = 8;
  cfg->komi = 7.5f;                                         // board.jl:297
  cfg->resign_threshold = -0.9;                             // mcts_play.jl:18
  cfg->resign_disable_frac = 0.05;                          // selfplay.jl:9
  cfg->n_games = 1;
  cfg->readouts = 800;                                      // mcts_play.jl:17
  cfg->nodes_per_game = 0;
  cfg->seed = 0;
  cfg->device = 0;
  cfg->world_size = 1;
  cfg->rank = 0;
  cfg->record_ring = 0;
  cfg->evaluator = AGZ_EVAL_DUMMY;
  cfg->inject_noise = 1;
  return AGZ_OK;
}

extern "C" void agz_engine_destroy(agz_engine* e) {
  if (!e) return;
#if AGZ_CUDA
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  if (e->replay) replay_destroy(e->replay);
  if (e->train) train_destroy(e->train);
  if (e->nn) nn_destroy(e->nn);
  for (int i = 0; i < 8; ++i) cudaEventDestroy(e->ev[i]);
  for (int i = 0; i < 2; ++i) cudaEventDestroy(e->ev_step[i]);
  for (int i = 0; i < 3; ++i) cudaEventDestroy(e->ev_join[i]);
  for (int i = 0; i < 2; ++i) cudaEventDestroy(e->ev_feat[i]);
  for (int i = 0; i < 2; ++i) cudaEventDestroy(e->ev_conv[i]);
  for (int i = 0; i < 2; ++i) cudaStreamDestroy(e->gstream[i]);
  cudaStreamDestroy(e->cstream);
#endif
  for (void* p : e->allocs) devrt::dfree(p);
#if AGZ_CUDA
  cudaStreamDestroy(e->stream);
#endif
  delete e;
}

extern "C" int32_t agz_engine_create(const agz_config* cfg, agz_engine** out) {
  if (!cfg || !out) return fail(nullptr, AGZ_ERR_ARG, "null argument");
  *out = nullptr;
  const int N = cfg->board_n;
  if (N < 2 || N > AGZ_MAX_N) return fail(nullptr, AGZ_ERR_ARG, "board_n out of range");
  if (cfg->n_games < 1) return fail(nullptr, AGZ_ERR_ARG, "n_games must be >= 1");
  if (cfg->parallel_readouts < 1 || cfg->max_parallel < cfg->parallel_readouts) return fail(nullptr, AGZ_ERR_ARG, "need 1 <= parallel_readouts <= max_parallel");
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) return fail(nullptr, AGZ_ERR_ARG, "bad rank/world_size");
  if (cfg->max_game_length < 1 || cfg->max_game_length > 32000) return fail(nullptr, AGZ_ERR_ARG, "bad max_game_length");
  if (cfg->game != AGZ_GAME_GO && cfg->game != AGZ_GAME_GOMOKU) return fail(nullptr, AGZ_ERR_ARG, "unknown game %d", cfg->game);
  if (cfg->game == AGZ_GAME_GOMOKU && (cfg->n_in_row < 2 || cfg->n_in_row > N)) return fail(nullptr, AGZ_ERR_ARG, "n_in_row must be in [2, board_n]");
#if AGZ_CUDA
  {
    int ndev = 0;
    cudaError_t rc = cudaGetDeviceCount(&ndev);
    if (rc != cudaSuccess || ndev <= 0)
      return fail(nullptr, AGZ_ERR_CUDA, "no CUDA device: libagz has no CPU fallback (%s)", cudaGetErrorString(rc));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, AGZ_ERR_ARG, "device %d of %d", cfg->device, ndev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major != 10) return fail(nullptr, AGZ_ERR_CUDA, "device is sm_%d%d; libagz is built for sm_100a only", prop.major, prop.minor);
    if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(nullptr, AGZ_ERR_CUDA, "cudaSetDevice failed");
  }
#endif
  agz_engine* e = new agz_engine();
  memset(&e->c, 0, sizeof(e->c));
  memset(&e->v, 0, sizeof(e->v));
  e->cfg = *cfg;
  e->err[0] = 0;
  e->launches = 0;
  e->started = false;
  e->ring_head = 0;
  e->timing = 0;
  memset(e->phase_ms, 0, sizeof(e->phase_ms));
  memset(e->phase_launches, 0, sizeof(e->phase_launches));
#if AGZ_CUDA
  e->nn = nullptr;
  e->train = nullptr;
  e->replay = nullptr;
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete e;
    return fail(nullptr, AGZ_ERR_CUDA, "cudaStreamCreate failed");
  }
  for (int i = 0; i < 8; ++i) cudaEventCreate(&e->ev[i]);
  for (int i = 0; i < 2; ++i) cudaEventCreate(&e->ev_step[i]);
  for (int i = 0; i < 3; ++i) cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming);
  for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&e->ev_feat[i], cudaEventDisableTiming);
  for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&e->ev_conv[i], cudaEventDisableTiming);
  {
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    for (int i = 0; i < 2; ++i) cudaStreamCreateWithPriority(&e->gstream[i], cudaStreamNonBlocking, least);
    cudaStreamCreateWithPriority(&e->cstream, cudaStreamNonBlocking, greatest);
  }
  // option schedule.pipeline -- 0 (default): one batch per round on one stream.  1: two half batches, convolutions on a
  // high-priority stream and the tree / heads kernels of the other half underneath (pipelined_rounds).  On a power-capped B200 the
  // overlap does not pay: measured back to back on one box, 2551 / 2545 moves/s sequential vs 2522 / 2516 pipelined
  // (profiles/r01_schedule_ab.md).
  e->pipeline = 0;
#else
  e->stream = 0;
#endif
  Cfg& c = e->c;
  c.N = N;
  c.N2 = N * N;
  c.game = cfg->game == AGZ_GAME_GOMOKU ? GAME_GOMOKU : GAME_GO;
  c.n_in_row = cfg->n_in_row;
  c.pass = c.game == GAME_GO ? N * N : -1;
  c.A = N * N + (c.game == GAME_GO ? 1 : 0);
  int ka = (c.A + 31) / 32;
  c.KA = ka <= 3 ? 3 : (ka <= 6 ? 6 : 12);
  c.AS = c.KA * 32;
  c.KB = (c.N2 + 31) / 32;
  c.pmax = cfg->max_parallel;
  c.parallel = cfg->parallel_readouts;
  c.max_game_length = cfg->max_game_length;
  c.maxd = cfg->max_game_length + 4;
  c.tau_threshold = cfg->tau_threshold;
  c.readouts = cfg->readouts;
  c.n_games = cfg->n_games;
  c.world = cfg->world_size;
  c.rank = cfg->rank;
  c.inject_noise = cfg->inject_noise;
  c.komi = cfg->komi;
  c.c_puct = cfg->c_puct;
  c.noise_weight = cfg->noise_weight;
  c.noise_alpha = cfg->noise_alpha;
  c.resign_threshold = cfg->resign_threshold;
  c.resign_disable_frac = cfg->resign_disable_frac;
  c.seed = cfg->seed;
  c.total_games = 0;
  c.stagger_rounds = 0;
  // Node arena per game.  The reference's tree is unbounded; a game can create at most (readouts + 2*parallel - 1) nodes per
  // move, and with a sharp network nearly all of them stay alive in the re-used subtree, so only
  // max_game_length * (readouts + 2*parallel) nodes are always enough.  Default: that worst case when it fits in 40 % of the free
  // device memory, otherwise what fits, never less than 10 moves' worth (the emulation build keeps the small default).
  const int need = cfg->readouts + 2 * c.pmax + 4;
  const long long base_cap = std::max(256, 10 * need);
  long long cap = cfg->nodes_per_game > 0 ? cfg->nodes_per_game : base_cap;
  e->bytes_per_node = (size_t)c.AS * 16 + sizeof(NodeMeta) + (size_t)3 * c.KA * 4 + 8;
#if AGZ_CUDA
  if (cfg->nodes_per_game <= 0) {
    const long long worst = (long long)cfg->max_game_length * (need - 4) + need;
    size_t free_b = 0, total_b = 0;
    long long budget = base_cap;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
      budget = (long long)(0.4 * (double)free_b / (double)cfg->n_games / (double)e->bytes_per_node);
    cap = std::max(base_cap, std::min(worst, budget));
  }
#endif
  if (cap > 0x7ffffff0LL) cap = 0x7ffffff0LL;   // node ids are 32-bit
  c.cap = (int)cap;
  if (c.cap < need + 2) c.cap = need + 2;
  c.ring_cap = cfg->record_ring > 0 ? cfg->record_ring : 2 * cfg->n_games;
  e->smem_per_warp = (int)((c.KB * 32 * 7 + 15) / 16 * 16);
  e->evaluator = cfg->evaluator;

  const size_t G = c.n_games, nodes = G * c.cap, L = c.max_game_length + 2, rows = G * c.pmax;
  View& v = e->v;
  int rc = 0;
  rc |= dalloc(e, &v.N, nodes * c.AS);
  rc |= dalloc(e, &v.W, nodes * c.AS);
  rc |= dalloc(e, &v.P, nodes * c.AS);
  rc |= dalloc(e, &v.child, nodes * c.AS);
  rc |= dalloc(e, &v.meta, nodes);
  rc |= dalloc(e, &v.bits, nodes * 3 * c.KA);   // node stride 3*KA words, planes KB words apart (tree.cuh bits_of)
  rc |= dalloc(e, &v.gs, G);
  rc |= dalloc(e, &v.hist, G * 7 * 2 * c.KB);
  rc |= dalloc(e, &v.path, rows * c.maxd);
  rc |= dalloc(e, &v.leaf_node, rows);
  rc |= dalloc(e, &v.leaf_plen, rows);
  rc |= dalloc(e, &v.remap, nodes);
  rc |= dalloc(e, &v.order, nodes);
  rc |= dalloc(e, &v.rec_moves, G * L);
  rc |= dalloc(e, &v.rec_q, G * L);
  rc |= dalloc(e, &v.rec_pi, G * L * c.A);
  rc |= dalloc(e, &v.rec_vis, G * L * c.A);
  rc |= dalloc(e, &v.ring_hdr, (size_t)c.ring_cap);
  rc |= dalloc(e, &v.ring_moves, (size_t)c.ring_cap * L);
  rc |= dalloc(e, &v.ring_q, (size_t)c.ring_cap * L);
  rc |= dalloc(e, &v.ring_pi, (size_t)c.ring_cap * L * c.A);
  rc |= dalloc(e, &v.ring_vis, (size_t)c.ring_cap * L * c.A);
  rc |= dalloc(e, &v.ctr, (size_t)CTR_COUNT);
  rc |= dalloc(e, &v.noise_g, G * c.AS);
  // rcp[k] = RN(1/k): the PUCT score divides by 1 + N(child), and N(child) never exceeds the visits a game can accumulate on one
  // line (max_game_length searches of readouts + 2*parallel visits); larger divisors (test hooks) take the IEEE-division path
  {
    const long long worst = (long long)cfg->max_game_length * (cfg->readouts + 2 * c.pmax) + 64;
    const int rn = (int)std::min<long long>(std::max<long long>(worst, 4096), 1 << 20);
    double* d_rcp = nullptr;
    rc |= dalloc(e, &d_rcp, (size_t)rn);
    if (!rc) {
      std::vector<double> h((size_t)rn, 0.0);
      for (int k = 1; k < rn; ++k) { volatile double q = 1.0 / (double)k; h[(size_t)k] = q; }
      rc |= devrt::h2d(d_rcp, h.data(), h.size() * sizeof(double), e->stream);
      v.rcp = d_rcp;
      v.rcp_n = rn;
    }
  }
  rc |= dalloc(e, &e->d_dummy_pi, (size_t)c.A);
  rc |= dalloc(e, &e->d_dummy_v, (size_t)1);
  rc |= dalloc(e, &e->d_eval_pi, rows * c.A);
  rc |= dalloc(e, &e->d_eval_v, rows);
  rc |= dalloc(e, &e->d_hook_result, (size_t)8);
  rc |= dalloc(e, &e->d_hook_probs, (size_t)c.A);
  rc |= dalloc(e, &e->d_pos_in, (size_t)1);
  rc |= dalloc(e, &e->d_pos_out, (size_t)1);
  rc |= dalloc(e, &e->d_hook_legal, (size_t)AGZ_MAX_ACTIONS);
  rc |= dalloc(e, &e->d_hook_libs, (size_t)AGZ_MAX_POINTS);
  rc |= dalloc(e, &e->d_hook_feats, (size_t)17 * c.N2);
  e->d_feats_f32 = nullptr;
  e->d_trace = nullptr;
  e->trace_cap = 0;
  e->train_loaded = false;
  e->host_stale = false;
  e->d_match_i = e->d_match_j = e->d_match_in = nullptr;
  e->d_match_f = nullptr;
  e->d_match_active = nullptr;
  e->d_match_ids = nullptr;
  e->fuse_dummy = 1;
  e->replay_cap = 0;
  if (rc) {
    int code = fail(nullptr, AGZ_ERR_CUDA, "device allocation failed (n_games=%d, nodes_per_game=%d)", c.n_games, c.cap);
    agz_engine_destroy(e);
    return code;
  }
  // DummyNet default: uniform priors, value 0 (test_mcts_player.jl:14-18)
  std::vector<float> pri((size_t)c.A, (float)(1.0 / c.A));
  devrt::h2d(e->d_dummy_pi, pri.data(), pri.size() * sizeof(float), e->stream);
#if AGZ_CUDA
  {
    NNShape s{N, cfg->planes, cfg->filters, cfg->tower_height, c.A};
    char nerr[256] = "";
    if (cfg->planes != 17) {
      agz_engine_destroy(e);
      return fail(nullptr, AGZ_ERR_ARG, "planes must be 17");
    }
    e->nn = nn_create(s, (int)rows, nerr, sizeof(nerr));
    if (!e->nn) {
      int code = fail(nullptr, AGZ_ERR_CUDA, "nn_create: %s", nerr);
      agz_engine_destroy(e);
      return code;
    }
    int r2 = dalloc(e, &e->d_feats_f32, rows * 17 * c.N2);
    if (r2) {
      agz_engine_destroy(e);
      return fail(nullptr, AGZ_ERR_CUDA, "feature buffer allocation failed");
    }
  }
#endif
  int src = devrt::sync(e->stream);
  if (src) {
    agz_engine_destroy(e);
    return fail(nullptr, AGZ_ERR_CUDA, "engine init failed: %s", devrt::last_error_string(src));
  }
  *out = e;
  return AGZ_OK;
}

// ------------------------------------------------------------------------------------------- options
// Everything round 1 read from environment variables is a named integer option of the engine (include/agz.h).
extern "C" int32_t agz_set_option(agz_engine* e, const char* key, int64_t value) {
  if (!e || !key) return fail(e, AGZ_ERR_ARG, "null argument");
  if (!strcmp(key, "dummy.fused_rounds")) { e->fuse_dummy = value != 0; return AGZ_OK; }
  if (!strcmp(key, "selfplay.stagger_rounds")) {
    if (value < 0) return fail(e, AGZ_ERR_ARG, "selfplay.stagger_rounds must be >= 0");
    e->c.stagger_rounds = value;   // read by the next agz_selfplay_start
    return AGZ_OK;
  }
#if AGZ_CUDA
  if (!strcmp(key, "schedule.pipeline")) { e->pipeline = value != 0; return AGZ_OK; }
  if (!strcmp(key, "replay.capacity")) {
    if (value < 1) return fail(e, AGZ_ERR_ARG, "replay.capacity must be >= 1");
    if (e->replay) return fail(e, AGZ_ERR_ARG, "replay.capacity must be set before the replay ring exists (first agz_replay_gather / agz_nccl_init)");
    e->replay_cap = value;
    return AGZ_OK;
  }
  if (!strcmp(key, "trace.records")) {   // kernel timeline trace (agz_trace_read); the buffer is allocated once
    if (value < 0 || value > (1 << 24)) return fail(e, AGZ_ERR_ARG, "trace.records out of range");
    if (e->d_trace) return fail(e, AGZ_ERR_ARG, "trace.records is already set");
    if (value == 0) return AGZ_OK;
    cudaSetDevice(e->cfg.device);
    e->trace_cap = (int)value;
    if (dalloc(e, &e->d_trace, (size_t)2 + 4 * (size_t)e->trace_cap)) return fail(e, AGZ_ERR_CUDA, "trace buffer allocation failed");
    unsigned long long cap = (unsigned long long)e->trace_cap;
    DCHECK(e, devrt::h2d(e->d_trace + 1, &cap, sizeof(cap), e->stream));
    e->v.trace = e->d_trace;
    nn_tc_set_trace(e->nn, e->d_trace);
    return AGZ_OK;
  }
  if (!strncmp(key, "conv.", 5)) {
    const int rc = nn_tc_set_option(e->nn, key, value);
    if (rc == 1) return fail(e, AGZ_ERR_ARG, "unknown option %s", key);
    if (rc) return fail(e, AGZ_ERR_ARG, "bad value %lld for option %s", (long long)value, key);
    return AGZ_OK;
  }
#endif
  return fail(e, AGZ_ERR_ARG, "unknown option %s", key);
}

extern "C" int32_t agz_get_option(agz_engine* e, const char* key, int64_t* value) {
  if (!e || !key || !value) return fail(e, AGZ_ERR_ARG, "null argument");
  if (!strcmp(key, "dummy.fused_rounds")) { *value = e->fuse_dummy; return AGZ_OK; }
  if (!strcmp(key, "selfplay.stagger_rounds")) { *value = e->c.stagger_rounds; return AGZ_OK; }
#if AGZ_CUDA
  if (!strcmp(key, "schedule.pipeline")) { *value = e->pipeline; return AGZ_OK; }
  if (!strcmp(key, "replay.capacity")) { *value = e->replay ? replay_capacity(e->replay) : (e->replay_cap > 0 ? e->replay_cap : replay_default_capacity()); return AGZ_OK; }
  if (!strcmp(key, "trace.records")) { *value = e->trace_cap; return AGZ_OK; }
  if (!strncmp(key, "conv.", 5)) {
    long long v = 0;
    if (nn_tc_get_option(e->nn, key, &v)) return fail(e, AGZ_ERR_ARG, "unknown option %s", key);
    *value = v;
    return AGZ_OK;
  }
#endif
  return fail(e, AGZ_ERR_ARG, "unknown option %s", key);
}

// ------------------------------------------------------------------------------------------- evaluator
static void bind_evaluator(agz_engine* e) {
  if (e->evaluator == AGZ_EVAL_DUMMY) {
    e->v.eval_pi = e->d_dummy_pi;
    e->v.eval_v = e->d_dummy_v;
    e->v.pi_stride = 0;
    e->v.v_stride = 0;
  } else {
    e->v.eval_pi = e->d_eval_pi;
    e->v.eval_v = e->d_eval_v;
    e->v.pi_stride = e->c.A;
    e->v.v_stride = 1;
  }
}

extern "C" int32_t agz_set_dummy_evaluator(agz_engine* e, const float* priors, float value) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  std::vector<float> pri((size_t)e->c.A, (float)(1.0 / e->c.A));
  if (priors) memcpy(pri.data(), priors, pri.size() * sizeof(float));
  DCHECK(e, devrt::h2d(e->d_dummy_pi, pri.data(), pri.size() * sizeof(float), e->stream));
  DCHECK(e, devrt::h2d(e->d_dummy_v, &value, sizeof(float), e->stream));
  return AGZ_OK;
}

extern "C" int32_t agz_set_evaluator(agz_engine* e, int32_t evaluator) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  if (evaluator != AGZ_EVAL_DUMMY && evaluator != AGZ_EVAL_NN_TC && evaluator != AGZ_EVAL_NN_F32) return fail(e, AGZ_ERR_ARG, "unknown evaluator %d", evaluator);
#if !AGZ_CUDA
  if (evaluator != AGZ_EVAL_DUMMY) return fail(e, AGZ_ERR_ARG, "network evaluators need the CUDA build");
#endif
  e->evaluator = evaluator;
  return AGZ_OK;
}

#if AGZ_CUDA
static int sync_host(agz_engine* e);
// features + network for batch rows [row0, row0 + nrows): writes d_eval_pi / d_eval_v
static int run_network(agz_engine* e, int row0, int nrows) {
  char nerr[256] = "";
  if (!nn_ready(e->nn)) {
    if (nn_commit(e->nn, e->stream, nerr, sizeof(nerr))) return fail(e, AGZ_ERR_ARG, "network not ready: %s", nerr);
  }
  if (e->timing) cudaEventRecord(e->ev[1], e->stream);
  cudaEvent_t* nev = e->timing ? &e->ev[2] : nullptr;   // ev[2..5]: before stem, after stem, after tower, after heads
  if (e->evaluator == AGZ_EVAL_NN_F32) {
    { int rc = sync_host(e); if (rc) return rc; }   // the fp32 cross-check path builds its weights from the host copy
    float* feats = e->d_feats_f32 + (size_t)row0 * 17 * e->c.N2;
    DISPATCH_KA(e, {
      LeafFeaturesF32Op<KA> op{e->c, e->v, e->d_feats_f32};
      // the op indexes rows from 0; launch all rows up to row0+nrows and let it skip (cheap), or offset:
      (void)feats;
      DCHECK(e, devrt::launch_warps(op, row0 + nrows, e->smem_per_warp, e->stream));
    });
    e->launches += 1;
    int rc = nn_forward_f32(e->nn, e->d_feats_f32 + (size_t)row0 * 17 * e->c.N2, nrows, e->d_eval_pi + (size_t)row0 * e->c.A, e->d_eval_v + row0, e->stream, nev);
    if (rc) return fail(e, AGZ_ERR_CUDA, "nn_forward_f32: %s", cudaGetErrorString((cudaError_t)rc));
    e->launches += nn_f32_launches_per_forward(e->nn);
  } else {
    int rc = engine_tc_features(e->c, e->v, e->nn, row0, nrows, e->smem_per_warp, e->stream);
    if (rc) return fail(e, AGZ_ERR_CUDA, "tc feature kernel: %s", cudaGetErrorString((cudaError_t)rc));
    e->launches += 1;
    rc = nn_forward_tc(e->nn, row0 + nrows, e->d_eval_pi, e->d_eval_v, e->stream, nerr, sizeof(nerr), nev);
    if (rc) return fail(e, AGZ_ERR_CUDA, "nn_forward_tc: %s", nerr);
    e->launches += nn_tc_launches_per_forward(e->nn);
  }
  return AGZ_OK;
}
#endif

// ------------------------------------------------------------------------------------------- self-play
static int read_progress(agz_engine* e, agz_progress* p) {
  unsigned long long ctr[CTR_COUNT];
  DCHECK(e, devrt::d2h(ctr, e->v.ctr, sizeof(ctr), e->stream));
  std::vector<GameState> gs((size_t)e->c.n_games);
  DCHECK(e, devrt::d2h(gs.data(), e->v.gs, gs.size() * sizeof(GameState), e->stream));
  p->moves_played = (int64_t)ctr[CTR_MOVES];
  p->games_finished = (int64_t)ctr[CTR_FINISHED];
  p->games_started = (int64_t)ctr[CTR_STARTED];
  p->positions_evaluated = (int64_t)ctr[CTR_POSITIONS];
  p->readouts = (int64_t)ctr[CTR_READOUTS];
  p->path_nodes = (int64_t)ctr[CTR_PATHNODES];
  p->arena_prunes = (int32_t)ctr[CTR_PRUNES];
  p->games_live = 0;
  p->error = 0;
  for (auto& g : gs) {
    if (g.phase == PH_SEED || g.phase == PH_SEARCH || g.phase == PH_WAIT_RING) p->games_live++;   // PH_DELAY (staggered start) is not live yet
    if (g.err && !p->error) p->error = g.err;
  }
  return AGZ_OK;
}

extern "C" int32_t agz_selfplay_start(agz_engine* e, int64_t total_games) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  e->c.total_games = total_games;
  DCHECK(e, devrt::dmemset(e->v.ctr, 0, sizeof(unsigned long long) * CTR_COUNT, e->stream));
  e->ring_head = 0;
  bind_evaluator(e);
  DISPATCH_KA(e, {
    StartOp<KA> op{e->c, e->v};
    DCHECK(e, devrt::launch_warps(op, e->c.n_games, e->smem_per_warp, e->stream));
  });
  e->launches += 1;
  e->started = true;
  DCHECK(e, devrt::sync(e->stream));
  return AGZ_OK;
}

static int one_round(agz_engine* e) {
#if AGZ_CUDA
  if (e->timing) cudaEventRecord(e->ev[0], e->stream);
#endif
  if (e->c.n_games >= 2048) {
    DISPATCH_KA(e, {
      SelectOp<KA, 1> op{e->c, e->v, -1, e->c.parallel, 0};
      DCHECK(e, devrt::launch_warps(op, e->c.n_games, e->smem_per_warp, e->stream));
    });
  } else {
    DISPATCH_KA(e, {
      SelectOp<KA> op{e->c, e->v, -1, e->c.parallel, 0};
      DCHECK(e, devrt::launch_warps(op, e->c.n_games, e->smem_per_warp, e->stream));
    });
  }
  e->launches += 1;
#if AGZ_CUDA
  if (e->evaluator != AGZ_EVAL_DUMMY) {
    int rc = run_network(e, 0, e->c.n_games * e->c.pmax);
    if (rc) return rc;
  } else if (e->timing) {
    for (int i = 1; i <= 5; ++i) cudaEventRecord(e->ev[i], e->stream);
  }
#endif
  DISPATCH_KA(e, {
    IncorporateOp<KA> op{e->c, e->v, -1, 0};
    DCHECK(e, devrt::launch_warps(op, e->c.n_games, e->smem_per_warp, e->stream));
  });
  e->launches += 1;
#if AGZ_CUDA
  if (e->timing) {
    cudaEventRecord(e->ev[6], e->stream);
    cudaEventSynchronize(e->ev[6]);
    for (int i = 0; i < AGZ_NKERNELS; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]);
      e->phase_ms[i] += ms;
    }
    const bool net = e->evaluator != AGZ_EVAL_DUMMY;
    e->phase_launches[0] += 1;
    e->phase_launches[1] += net ? 1 : 0;
    e->phase_launches[2] += net ? 1 : 0;
    e->phase_launches[3] += net ? 2 * e->cfg.tower_height : 0;
    e->phase_launches[4] += net ? 1 : 0;
    e->phase_launches[5] += 1;
  }
#endif
  return AGZ_OK;
}

#if AGZ_CUDA
// Half-batch pipelining: slots [0, G/2) and [G/2, G) are two groups.  The convolutions of both groups run back to back on
// one high-priority stream (group 0's tower, group 1's tower, group 0's next tower, ...), so the tensor cores never wait;
// each group's heads -> incorporate -> select -> leaf-features chain runs on its own low-priority stream underneath the
// other group's tower (events ev_conv[g] / ev_feat[g] hand a group from one stream to the other).  The priorities matter:
// the heads kernel (27 KB of shared memory per CTA, 8 CTAs per SM) would otherwise grab the SMs the moment a tower ends and
// hold the next stem back for its whole duration (measured with the AGZ_TRACE timeline: a 135 us bubble per half round).
// Results are identical to the sequential schedule: games never interact.
static bool can_pipeline(agz_engine* e) {
  long long prec = 1;
  nn_tc_get_option(e->nn, "conv.precision", &prec);
  return prec == 1 && e->pipeline && !e->timing && e->evaluator == AGZ_EVAL_NN_TC && e->c.n_games >= 2 && e->c.n_games % 2 == 0 && nn_tc_groups(e->nn) == 2;
}

static int pipelined_rounds(agz_engine* e, int rounds) {
  char nerr[256] = "";
  if (rounds <= 0) return AGZ_OK;
  devrt::prefer_max_smem() = 1;
  if (!nn_ready(e->nn) && nn_commit(e->nn, e->stream, nerr, sizeof(nerr))) return fail(e, AGZ_ERR_ARG, "network not ready: %s", nerr);
  const int half = e->c.n_games / 2, rows = half * e->c.pmax;
  cudaEventRecord(e->ev_join[2], e->stream);
  cudaStreamWaitEvent(e->cstream, e->ev_join[2], 0);
  for (int g = 0; g < 2; ++g) cudaStreamWaitEvent(e->gstream[g], e->ev_join[2], 0);
  auto select_and_features = [&](int g) -> int {
    cudaStream_t st = e->gstream[g];
    DISPATCH_KA(e, {
      SelectOp<KA> op{e->c, e->v, -1, e->c.parallel, g * half};
      DCHECK(e, devrt::launch_warps(op, half, e->smem_per_warp, st));
    });
    int rc = engine_tc_features(e->c, e->v, e->nn, g * rows, rows, e->smem_per_warp, st);
    if (rc) return fail(e, AGZ_ERR_CUDA, "tc feature kernel: %s", cudaGetErrorString((cudaError_t)rc));
    cudaEventRecord(e->ev_feat[g], st);
    e->launches += 2;
    return AGZ_OK;
  };
  for (int g = 0; g < 2; ++g) {
    int rc = select_and_features(g);
    if (rc) return rc;
  }
  for (int r = 0; r < rounds; ++r) {
    for (int g = 0; g < 2; ++g) {
      cudaStream_t st = e->gstream[g];
      cudaStreamWaitEvent(e->cstream, e->ev_feat[g], 0);
      int rc = nn_forward_tc(e->nn, rows, e->d_eval_pi + (size_t)g * rows * e->c.A, e->d_eval_v + (size_t)g * rows, e->cstream, nerr, sizeof(nerr), nullptr, g,
                             e->ev_conv[g], st);
      if (rc) return fail(e, AGZ_ERR_CUDA, "nn_forward_tc: %s", nerr);
      DISPATCH_KA(e, {
        IncorporateOp<KA> op{e->c, e->v, -1, g * half};
        DCHECK(e, devrt::launch_warps(op, half, e->smem_per_warp, st));
      });
      e->launches += 1 + nn_tc_launches_per_forward(e->nn);
      if (r + 1 < rounds) {
        rc = select_and_features(g);
        if (rc) return rc;
      }
    }
  }
  for (int g = 0; g < 2; ++g) {
    cudaEventRecord(e->ev_join[g], e->gstream[g]);
    cudaStreamWaitEvent(e->stream, e->ev_join[g], 0);
  }
  devrt::prefer_max_smem() = 0;
  return AGZ_OK;
}
#endif

extern "C" int32_t agz_selfplay_step(agz_engine* e, int32_t rounds, agz_progress* progress) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  if (!e->started) return fail(e, AGZ_ERR_ARG, "agz_selfplay_start has not been called");
  bind_evaluator(e);
#if AGZ_CUDA
  if (progress) cudaEventRecord(e->ev_step[0], e->stream);
#endif
#if AGZ_CUDA
  if (can_pipeline(e)) {
    int rc = pipelined_rounds(e, rounds);
    if (rc) return rc;
    rounds = 0;
  }
#endif
  // DummyNet evaluator: all rounds of the call in one launch per game (ops.cuh DummyRoundsOp); per-kernel timing and the option
  // dummy.fused_rounds = 0 keep the split
  if (rounds > 0 && e->fuse_dummy && e->evaluator == AGZ_EVAL_DUMMY && !e->timing) {
    if (e->c.n_games >= 2048) {
      DISPATCH_KA(e, {
        DummyRoundsOp<KA, 1> op{e->c, e->v, rounds};
        DCHECK(e, devrt::launch_warps(op, e->c.n_games, e->smem_per_warp, e->stream));
      });
    } else {
      DISPATCH_KA(e, {
        DummyRoundsOp<KA> op{e->c, e->v, rounds};
        DCHECK(e, devrt::launch_warps(op, e->c.n_games, e->smem_per_warp, e->stream));
      });
    }
    e->launches += 1;
    rounds = 0;
  }
  for (int r = 0; r < rounds; ++r) {
    int rc = one_round(e);
    if (rc) return rc;
  }
  if (progress) {
#if AGZ_CUDA
    cudaEventRecord(e->ev_step[1], e->stream);
#endif
    DCHECK(e, devrt::sync(e->stream));
    int rc = read_progress(e, progress);
    progress->step_ms = 0.f;
#if AGZ_CUDA
    cudaEventElapsedTime(&progress->step_ms, e->ev_step[0], e->ev_step[1]);
#endif
    return rc;
  }
  return AGZ_OK;
}

extern "C" int32_t agz_selfplay_harvest(agz_engine* e, int32_t max_records, agz_game_header* headers, int16_t* moves, float* qs,
                                        float* pis, float* visits, int32_t* n_out) {
  if (!e || !n_out) return fail(e, AGZ_ERR_ARG, "null argument");
  *n_out = 0;
  DCHECK(e, devrt::sync(e->stream));
  unsigned long long ctr[CTR_COUNT];
  DCHECK(e, devrt::d2h(ctr, e->v.ctr, sizeof(ctr), e->stream));
  const unsigned long long tail = ctr[CTR_RING_TAIL];
  const size_t L = e->c.max_game_length + 2, A = e->c.A;
  int n = 0;
  // the ring is fixed-stride, so a run of finished records is at most two contiguous device ranges: copy each array
  // of a range with ONE transfer straight into the caller's buffers, then blank what lies past each game's length
  while (e->ring_head < tail && n < max_records) {
    const size_t rs = (size_t)(e->ring_head % (unsigned long long)e->c.ring_cap);
    size_t run = std::min<unsigned long long>(tail - e->ring_head, (unsigned long long)(max_records - n));
    run = std::min(run, (size_t)e->c.ring_cap - rs);
    std::vector<RingHeader> hd(run);
    DCHECK(e, devrt::d2h(hd.data(), e->v.ring_hdr + rs, run * sizeof(RingHeader), e->stream));
    if (moves) DCHECK(e, devrt::d2h(moves + n * L, e->v.ring_moves + rs * L, run * L * sizeof(int16_t), e->stream));
    if (qs) DCHECK(e, devrt::d2h(qs + n * L, e->v.ring_q + rs * L, run * L * sizeof(float), e->stream));
    if (pis) DCHECK(e, devrt::d2h(pis + n * L * A, e->v.ring_pi + rs * L * A, run * L * A * sizeof(float), e->stream));
    if (visits) DCHECK(e, devrt::d2h(visits + n * L * A, e->v.ring_vis + rs * L * A, run * L * A * sizeof(float), e->stream));
    for (size_t r = 0; r < run; ++r) {
      const size_t k = (size_t)n + r, nm = (size_t)hd[r].n_moves;
      if (headers) {
        headers[k].game_id = hd[r].game_id; headers[k].n_moves = hd[r].n_moves; headers[k].result = hd[r].result;
        headers[k].resigned = hd[r].resigned; headers[k].final_score = hd[r].final_score; headers[k].resign_threshold = hd[r].resign_threshold;
      }
      if (moves) memset(moves + k * L + nm, 0, (L - nm) * sizeof(int16_t));
      if (qs) memset(qs + k * L + nm, 0, (L - nm) * sizeof(float));
      if (pis) memset(pis + (k * L + nm) * A, 0, (L - nm) * A * sizeof(float));
      if (visits) memset(visits + (k * L + nm) * A, 0, (L - nm) * A * sizeof(float));
    }
    n += (int)run;
    e->ring_head += run;
  }
  unsigned long long h = e->ring_head;
  DCHECK(e, devrt::h2d(e->v.ctr + CTR_RING_HEAD, &h, sizeof(h), e->stream));
  *n_out = n;
  return AGZ_OK;
}

extern "C" int32_t agz_selfplay_run(agz_engine* e, int32_t total_games, agz_game_header* headers, int16_t* moves, float* qs, float* pis,
                                    float* visits) {
  if (!e || total_games < 1) return fail(e, AGZ_ERR_ARG, "bad argument");
  int rc = agz_selfplay_start(e, total_games);
  if (rc) return rc;
  // games of this rank: ids rank, rank+world, ... < total
  const int mine = (total_games - e->c.rank + e->c.world - 1) / e->c.world;
  const size_t L = e->c.max_game_length + 2, A = e->c.A;
  std::vector<agz_game_header> hd((size_t)std::max(1, e->c.ring_cap));
  std::vector<int16_t> mv; std::vector<float> q, pi, vis;
  if (moves) mv.resize(hd.size() * L);
  if (qs) q.resize(hd.size() * L);
  if (pis) pi.resize(hd.size() * L * A);
  if (visits) vis.resize(hd.size() * L * A);
  int got = 0;
  const int chunk = 8;
  long long guard = 0;
  while (got < mine) {
    agz_progress pr;
    rc = agz_selfplay_step(e, chunk, &pr);
    if (rc) return rc;
    if (pr.error) return fail(e, pr.error, "a game stopped with error %d (6 = node arena full: raise nodes_per_game)", pr.error);
    int n = 0;
    rc = agz_selfplay_harvest(e, (int)hd.size(), hd.data(), moves ? mv.data() : nullptr, qs ? q.data() : nullptr, pis ? pi.data() : nullptr,
                              visits ? vis.data() : nullptr, &n);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
      // records are returned in local game order: index = (game_id - rank) / world
      long long idx = (hd[i].game_id - e->c.rank) / e->c.world;
      if (idx < 0 || idx >= mine) continue;
      if (headers) headers[idx] = hd[i];
      if (moves) memcpy(moves + idx * L, mv.data() + i * L, L * sizeof(int16_t));
      if (qs) memcpy(qs + idx * L, q.data() + i * L, L * sizeof(float));
      if (pis) memcpy(pis + idx * L * A, pi.data() + i * L * A, L * A * sizeof(float));
      if (visits) memcpy(visits + idx * L * A, vis.data() + i * L * A, L * A * sizeof(float));
      ++got;
    }
    if (pr.games_live == 0 && got < mine && n == 0) return fail(e, AGZ_ERR_ASSERT, "self-play stalled with %d of %d games", got, mine);
    if (++guard > 100000000LL) return fail(e, AGZ_ERR_ASSERT, "self-play did not terminate");
  }
  return AGZ_OK;
}

// ------------------------------------------------------------------------------------------------ matches
// evaluate (neural_net.jl:103-158) / play (play.jl:25-77): every slot is one game seen by ONE player (this engine's network);
// the host keeps a second engine for the opponent and alternates them, exactly as the reference alternates two MCTSPlayers.
static int match_scratch(agz_engine* e) {
  if (e->d_match_i) return AGZ_OK;
  const size_t G = (size_t)e->c.n_games;
  int rc = 0;
  rc |= dalloc(e, &e->d_match_i, G);
  rc |= dalloc(e, &e->d_match_j, G);
  rc |= dalloc(e, &e->d_match_f, G);
  rc |= dalloc(e, &e->d_match_in, G);
  rc |= dalloc(e, &e->d_match_active, G);
  rc |= dalloc(e, &e->d_match_ids, G);
  if (rc) return fail(e, AGZ_ERR_CUDA, "match scratch allocation failed");
  return AGZ_OK;
}

template <class F>
static int launch_match(agz_engine* e, F fill) {
  bind_evaluator(e);
  DISPATCH_KA(e, {
    MatchOp<KA> op{e->c, e->v, 0, nullptr, nullptr, nullptr, e->d_match_i, e->d_match_j, e->d_match_f};
    fill(op);
    DCHECK(e, devrt::launch_warps(op, e->c.n_games, e->smem_per_warp, e->stream));
  });
  e->launches += 1;
  return AGZ_OK;
}

extern "C" int32_t agz_match_start(agz_engine* e, const int64_t* game_ids) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  int rc = match_scratch(e);
  if (rc) return rc;
  DCHECK(e, devrt::dmemset(e->v.ctr, 0, sizeof(unsigned long long) * CTR_COUNT, e->stream));
  if (game_ids) DCHECK(e, devrt::h2d(e->d_match_ids, game_ids, sizeof(long long) * e->c.n_games, e->stream));
  const long long* ids = game_ids ? e->d_match_ids : nullptr;
  rc = launch_match(e, [&](auto& op) { op.kind = MK_BEGIN; op.game_ids = ids; });
  if (rc) return rc;
  e->started = true;
  DCHECK(e, devrt::sync(e->stream));
  return AGZ_OK;
}

extern "C" int32_t agz_match_search(agz_engine* e, const uint8_t* active, int32_t* moves, int32_t* resigned, float* scores) {
  if (!e || !active || !moves || !resigned) return fail(e, AGZ_ERR_ARG, "null argument");
  if (!e->d_match_i) return fail(e, AGZ_ERR_ARG, "agz_match_start has not been called");
  const size_t G = (size_t)e->c.n_games;
  DCHECK(e, devrt::h2d(e->d_match_active, active, G, e->stream));
  int rc = launch_match(e, [&](auto& op) { op.kind = MK_ARM; op.active = e->d_match_active; });
  if (rc) return rc;
  // every tree_search! adds at least one visit to a searching root, so the loop ends; poll the busy counter after the
  // minimum number of rounds a search can take and then every few rounds
  const int first = (e->c.readouts + 2 * e->c.parallel - 1) / (2 * e->c.parallel);
  long long guard = 0;
  for (int chunk = first > 0 ? first : 1;; chunk = 4) {
    unsigned long long busy = 0;
    DCHECK(e, devrt::d2h(&busy, e->v.ctr + CTR_MATCH_BUSY, sizeof(busy), e->stream));
    if (busy == 0) break;
    for (int r = 0; r < chunk; ++r) {
      rc = one_round(e);
      if (rc) return rc;
    }
    if ((guard += chunk) > 100000000LL) return fail(e, AGZ_ERR_ASSERT, "match search did not terminate");
  }
  rc = launch_match(e, [&](auto& op) { op.kind = MK_PICK; op.active = e->d_match_active; });
  if (rc) return rc;
  DCHECK(e, devrt::d2h(moves, e->d_match_i, sizeof(int32_t) * G, e->stream));
  DCHECK(e, devrt::d2h(resigned, e->d_match_j, sizeof(int32_t) * G, e->stream));
  if (scores) DCHECK(e, devrt::d2h(scores, e->d_match_f, sizeof(float) * G, e->stream));
  std::vector<GameState> gs(G);
  DCHECK(e, devrt::d2h(gs.data(), e->v.gs, G * sizeof(GameState), e->stream));
  for (size_t g = 0; g < G; ++g)
    if (gs[g].err) return fail(e, gs[g].err == E_CAPACITY ? AGZ_ERR_CAPACITY : AGZ_ERR_ASSERT, "match slot %d stopped with device status %d", (int)g, gs[g].err);
  return AGZ_OK;
}

extern "C" int32_t agz_match_play(agz_engine* e, const int32_t* moves, int32_t* done, float* scores) {
  if (!e || !moves) return fail(e, AGZ_ERR_ARG, "null argument");
  if (!e->d_match_i) return fail(e, AGZ_ERR_ARG, "agz_match_start has not been called");
  const size_t G = (size_t)e->c.n_games;
  for (size_t g = 0; g < G; ++g)
    if (moves[g] >= e->c.A) return fail(e, AGZ_ERR_ARG, "move out of range");
  DCHECK(e, devrt::h2d(e->d_match_in, moves, sizeof(int32_t) * G, e->stream));
  int rc = launch_match(e, [&](auto& op) { op.kind = MK_PLAY; op.moves_in = e->d_match_in; });
  if (rc) return rc;
  std::vector<int32_t> dn(G);
  DCHECK(e, devrt::d2h(dn.data(), e->d_match_i, sizeof(int32_t) * G, e->stream));
  if (done) memcpy(done, dn.data(), sizeof(int32_t) * G);
  if (scores) DCHECK(e, devrt::d2h(scores, e->d_match_f, sizeof(float) * G, e->stream));
  std::vector<GameState> gs(G);
  DCHECK(e, devrt::d2h(gs.data(), e->v.gs, G * sizeof(GameState), e->stream));
  for (size_t g = 0; g < G; ++g)
    if (gs[g].err) return fail(e, gs[g].err == E_CAPACITY ? AGZ_ERR_CAPACITY : AGZ_ERR_ASSERT, "match slot %d stopped with device status %d", (int)g, gs[g].err);
  for (size_t g = 0; g < G; ++g)
    if (dn[g] < 0) return fail(e, AGZ_ERR_ILLEGAL_MOVE, "illegal move in slot %d", (int)g);
  return AGZ_OK;
}

// ------------------------------------------------------------------------------------------------ hooks
static int run_hook(agz_engine* e, HookParams& h, int* ival, float* fval) {
  h.result = e->d_hook_result;
  bind_evaluator(e);
  DISPATCH_KA(e, {
    HookOp<KA> op{e->c, e->v, h};
    DCHECK(e, devrt::launch_warps(op, 1, e->smem_per_warp, e->stream));
  });
  e->launches += 1;
  int res[4];
  DCHECK(e, devrt::d2h(res, e->d_hook_result, sizeof(res), e->stream));
  if (ival) *ival = res[1];
  if (fval) memcpy(fval, &res[3], sizeof(float));
  if (res[0] == E_ILLEGAL) return fail(e, AGZ_ERR_ILLEGAL_MOVE, "illegal move");
  if (res[0] == E_ASSERT) return fail(e, AGZ_ERR_ASSERT, "assertion failed on device");
  if (res[0] == E_CAPACITY) return fail(e, AGZ_ERR_CAPACITY, "node arena full");
  if (res[0]) return fail(e, AGZ_ERR_ASSERT, "device status %d", res[0]);
  return AGZ_OK;
}

static int check_slot(agz_engine* e, int slot) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  if (slot < 0 || slot >= e->c.n_games) return fail(e, AGZ_ERR_ARG, "slot %d out of range", slot);
  return AGZ_OK;
}

static int check_node(agz_engine* e, int slot, int node) {
  int rc = check_slot(e, slot);
  if (rc) return rc;
  GameState gs;
  DCHECK(e, devrt::d2h(&gs, e->v.gs + slot, sizeof(gs), e->stream));
  if (node < 0 || node >= gs.count) return fail(e, AGZ_ERR_ARG, "node %d out of range (count %d)", node, gs.count);
  return AGZ_OK;
}

static HookParams hp(int kind, int slot) {
  HookParams h;
  memset(&h, 0, sizeof(h));
  h.kind = kind;
  h.slot = slot;
  h.node = -1;
  return h;
}

extern "C" int32_t agz_tree_init(agz_engine* e, int32_t slot, const agz_position* pos, int64_t game_id) {
  int rc = check_slot(e, slot);
  if (rc) return rc;
  HookParams h = hp(HK_INIT, slot);
  h.game_id = game_id;
  if (pos) {
    if (pos->to_play != 1 && pos->to_play != -1) return fail(e, AGZ_ERR_ARG, "to_play must be +-1");
    DCHECK(e, devrt::h2d(e->d_pos_in, pos, sizeof(*pos), e->stream));
    h.pos_in = e->d_pos_in;
  }
  return run_hook(e, h, nullptr, nullptr);
}

extern "C" int32_t agz_tree_select_leaf(agz_engine* e, int32_t slot, int32_t from_node, int32_t* leaf) {
  int rc = from_node >= 0 ? check_node(e, slot, from_node) : check_slot(e, slot);
  if (rc) return rc;
  HookParams h = hp(HK_SELECT, slot);
  h.node = from_node;
  int iv = -1;
  rc = run_hook(e, h, &iv, nullptr);
  if (leaf) *leaf = iv;
  return rc;
}

static int node_hook(agz_engine* e, int kind, int slot, int node, const float* probs, float value) {
  int rc = check_node(e, slot, node);
  if (rc) return rc;
  HookParams h = hp(kind, slot);
  h.node = node;
  h.value = value;
  if (probs) {
    DCHECK(e, devrt::h2d(e->d_hook_probs, probs, sizeof(float) * e->c.A, e->stream));
    h.probs = e->d_hook_probs;
  }
  return run_hook(e, h, nullptr, nullptr);
}

extern "C" int32_t agz_tree_incorporate(agz_engine* e, int32_t slot, int32_t node, const float* probs, float value) {
  if (!probs) return fail(e, AGZ_ERR_ARG, "probs is null");
  return node_hook(e, HK_INCORPORATE, slot, node, probs, value);
}
extern "C" int32_t agz_tree_backup_value(agz_engine* e, int32_t slot, int32_t node, float value) { return node_hook(e, HK_BACKUP, slot, node, nullptr, value); }
extern "C" int32_t agz_tree_add_virtual_loss(agz_engine* e, int32_t slot, int32_t node) { return node_hook(e, HK_VLOSS_ADD, slot, node, nullptr, 0.f); }
extern "C" int32_t agz_tree_revert_virtual_loss(agz_engine* e, int32_t slot, int32_t node) { return node_hook(e, HK_VLOSS_REVERT, slot, node, nullptr, 0.f); }

extern "C" int32_t agz_tree_maybe_add_child(agz_engine* e, int32_t slot, int32_t node, int32_t fmove, int32_t* child) {
  int rc = check_node(e, slot, node);
  if (rc) return rc;
  if (fmove < 0 || fmove >= e->c.A) return fail(e, AGZ_ERR_ARG, "move out of range");
  HookParams h = hp(HK_ADD_CHILD, slot);
  h.node = node;
  h.fmove = fmove;
  int iv = -1;
  rc = run_hook(e, h, &iv, nullptr);
  if (child) *child = iv;
  return rc;
}

extern "C" int32_t agz_tree_search(agz_engine* e, int32_t slot, int32_t parallel_readouts, int32_t* n_leaves) {
  int rc = check_slot(e, slot);
  if (rc) return rc;
  if (parallel_readouts < 1 || parallel_readouts > e->c.pmax) return fail(e, AGZ_ERR_ARG, "parallel_readouts must be in [1, max_parallel=%d]", e->c.pmax);
  bind_evaluator(e);
  DISPATCH_KA(e, {
    SelectOp<KA> op{e->c, e->v, slot, parallel_readouts, 0};
    DCHECK(e, devrt::launch_warps(op, 1, e->smem_per_warp, e->stream));
  });
  e->launches += 1;
  GameState gs;
  DCHECK(e, devrt::d2h(&gs, e->v.gs + slot, sizeof(gs), e->stream));
  if (n_leaves) *n_leaves = gs.nleaf;
  if (gs.err) return fail(e, gs.err == E_CAPACITY ? AGZ_ERR_CAPACITY : AGZ_ERR_ASSERT, "select failed with device status %d", gs.err);
#if AGZ_CUDA
  if (e->evaluator != AGZ_EVAL_DUMMY && gs.nleaf > 0) {
    rc = run_network(e, slot * e->c.pmax, gs.nleaf);
    if (rc) return rc;
  }
#endif
  DISPATCH_KA(e, {
    IncorporateOp<KA> op{e->c, e->v, slot, 0};
    DCHECK(e, devrt::launch_warps(op, 1, e->smem_per_warp, e->stream));
  });
  e->launches += 1;
  DCHECK(e, devrt::d2h(&gs, e->v.gs + slot, sizeof(gs), e->stream));
  if (gs.err) return fail(e, AGZ_ERR_ASSERT, "incorporate failed with device status %d", gs.err);
  return AGZ_OK;
}

extern "C" int32_t agz_tree_inject_noise(agz_engine* e, int32_t slot) {
  int rc = check_slot(e, slot);
  if (rc) return rc;
  HookParams h = hp(HK_NOISE, slot);
  return run_hook(e, h, nullptr, nullptr);
}

extern "C" int32_t agz_tree_pick_move(agz_engine* e, int32_t slot, int32_t* fmove) {
  int rc = check_slot(e, slot);
  if (rc) return rc;
  HookParams h = hp(HK_PICK, slot);
  int iv = -1;
  rc = run_hook(e, h, &iv, nullptr);
  if (fmove) *fmove = iv;
  return rc;
}

extern "C" int32_t agz_tree_play_move(agz_engine* e, int32_t slot, int32_t fmove) {
  int rc = check_slot(e, slot);
  if (rc) return rc;
  if (fmove < 0 || fmove >= e->c.A) return fail(e, AGZ_ERR_ARG, "move out of range");
  HookParams h = hp(HK_PLAY, slot);
  h.fmove = fmove;
  return run_hook(e, h, nullptr, nullptr);
}

extern "C" int32_t agz_tree_should_resign(agz_engine* e, int32_t slot, double threshold, int32_t* yes) {
  int rc = check_slot(e, slot);
  if (rc) return rc;
  HookParams h = hp(HK_RESIGN, slot);
  h.thr = threshold;
  int iv = 0;
  rc = run_hook(e, h, &iv, nullptr);
  if (yes) *yes = iv;
  return rc;
}

extern "C" int32_t agz_tree_root(agz_engine* e, int32_t slot, int32_t* root, int32_t* node_count) {
  int rc = check_slot(e, slot);
  if (rc) return rc;
  GameState gs;
  DCHECK(e, devrt::d2h(&gs, e->v.gs + slot, sizeof(gs), e->stream));
  if (root) *root = gs.root;
  if (node_count) *node_count = gs.count;
  return AGZ_OK;
}

extern "C" int32_t agz_tree_pending_vlosses(agz_engine* e, int32_t slot, int32_t* pending) {
  int rc = check_slot(e, slot);
  if (rc) return rc;
  GameState gs;
  DCHECK(e, devrt::d2h(&gs, e->v.gs + slot, sizeof(gs), e->stream));
  if (pending) *pending = gs.vloss_balance;
  return AGZ_OK;
}

extern "C" int32_t agz_tree_read_node(agz_engine* e, int32_t slot, int32_t node, agz_node_view* out) {
  int rc = check_node(e, slot, node);
  if (rc) return rc;
  if (!out) return fail(e, AGZ_ERR_ARG, "null out");
  memset(out, 0, sizeof(*out));
  const Cfg& c = e->c;
  const size_t gi = (size_t)slot * c.cap + node, r = gi * c.AS;
  NodeMeta m;
  GameState gs;
  DCHECK(e, devrt::d2h(&m, e->v.meta + gi, sizeof(m), e->stream));
  DCHECK(e, devrt::d2h(&gs, e->v.gs + slot, sizeof(gs), e->stream));
  DCHECK(e, devrt::d2h(out->child_N, e->v.N + r, sizeof(float) * c.A, e->stream));
  DCHECK(e, devrt::d2h(out->child_W, e->v.W + r, sizeof(float) * c.A, e->stream));
  DCHECK(e, devrt::d2h(out->child_prior, e->v.P + r, sizeof(float) * c.A, e->stream));
  DCHECK(e, devrt::d2h(out->children, e->v.child + r, sizeof(int32_t) * c.A, e->stream));
  std::vector<uint32_t> bits((size_t)3 * c.KB);
  DCHECK(e, devrt::d2h(bits.data(), e->v.bits + gi * 3 * c.KA, bits.size() * sizeof(uint32_t), e->stream));
  out->parent = m.parent; out->fmove = m.fmove; out->to_play = m.to_play; out->n = m.n; out->ko = m.ko;
  out->is_expanded = (m.flags & F_EXPANDED) ? 1 : 0;
  if (!out->is_expanded) {   // child_W / child_prior of a node are first written by incorporate_results! (tree.cuh init_rows)
    memset(out->child_W, 0, sizeof(out->child_W));
    memset(out->child_prior, 0, sizeof(out->child_prior));
  }
  out->done = (m.flags & F_DONE) ? 1 : 0;
  out->last_move_pass = (m.flags & F_LASTPASS) ? 1 : 0;
  for (int p = 0; p < c.N2; ++p) {
    int b = (bits[p >> 5] >> (p & 31)) & 1, w = (bits[c.KB + (p >> 5)] >> (p & 31)) & 1;
    out->board[p] = (int8_t)(b - w);
    out->legal[p] = (int8_t)((bits[2 * c.KB + (p >> 5)] >> (p & 31)) & 1);
  }
  if (c.pass >= 0) out->legal[c.pass] = 1;
  if (m.parent < 0) { out->N = gs.root_N; out->W = gs.root_W; }
  else {
    const size_t pr = ((size_t)slot * c.cap + m.parent) * c.AS + m.fmove;
    DCHECK(e, devrt::d2h(&out->N, e->v.N + pr, sizeof(float), e->stream));
    DCHECK(e, devrt::d2h(&out->W, e->v.W + pr, sizeof(float), e->stream));
  }
  // child_action_score (mcts.jl:86-92) with the reference's float widths; host doubles, no contraction
  volatile float one_plus = 1.0f + out->N;
  volatile double cu = c.c_puct * (double)sqrtf(one_plus);
  for (int a = 0; a < c.A; ++a) {
    volatile float den = 1.0f + out->child_N[a];
    volatile float q = out->child_W[a] / den;
    volatile float qt = q * (float)m.to_play;
    volatile double u1 = cu * (double)out->child_prior[a];
    volatile double u = u1 / (double)den;
    out->action_score[a] = (double)qt + u;
  }
  return AGZ_OK;
}

extern "C" int32_t agz_tree_set_stats(agz_engine* e, int32_t slot, int32_t node, const float* self_N, const float* child_N, const int32_t* n_override) {
  int rc = check_node(e, slot, node);
  if (rc) return rc;
  const Cfg& c = e->c;
  const size_t gi = (size_t)slot * c.cap + node;
  NodeMeta m;
  DCHECK(e, devrt::d2h(&m, e->v.meta + gi, sizeof(m), e->stream));
  if (self_N) {
    if (m.parent < 0) {
      GameState gs;
      DCHECK(e, devrt::d2h(&gs, e->v.gs + slot, sizeof(gs), e->stream));
      gs.root_N = *self_N;
      DCHECK(e, devrt::h2d(e->v.gs + slot, &gs, sizeof(gs), e->stream));
    } else {
      DCHECK(e, devrt::h2d(e->v.N + ((size_t)slot * c.cap + m.parent) * c.AS + m.fmove, self_N, sizeof(float), e->stream));
    }
  }
  if (child_N) DCHECK(e, devrt::h2d(e->v.N + gi * c.AS, child_N, sizeof(float) * c.A, e->stream));
  if (n_override) {
    m.n = (int16_t)*n_override;
    DCHECK(e, devrt::h2d(e->v.meta + gi, &m, sizeof(m), e->stream));
  }
  return AGZ_OK;
}

extern "C" int32_t agz_tree_read_record(agz_engine* e, int32_t slot, int32_t* n_moves, int16_t* moves, float* qs, float* pis) {
  int rc = check_slot(e, slot);
  if (rc) return rc;
  GameState gs;
  DCHECK(e, devrt::d2h(&gs, e->v.gs + slot, sizeof(gs), e->stream));
  const size_t L = e->c.max_game_length + 2, A = e->c.A, nm = (size_t)gs.n_moves;
  if (n_moves) *n_moves = gs.n_moves;
  if (nm && moves) DCHECK(e, devrt::d2h(moves, e->v.rec_moves + slot * L, nm * sizeof(int16_t), e->stream));
  if (nm && qs) DCHECK(e, devrt::d2h(qs, e->v.rec_q + slot * L, nm * sizeof(float), e->stream));
  if (nm && pis) DCHECK(e, devrt::d2h(pis, e->v.rec_pi + slot * L * A, nm * A * sizeof(float), e->stream));
  return AGZ_OK;
}

extern "C" int32_t agz_tree_node_features(agz_engine* e, int32_t slot, int32_t node, float* out) {
  int rc = check_node(e, slot, node);
  if (rc) return rc;
  HookParams h = hp(HK_FEATURES, slot);
  h.node = node;
  h.f_out = e->d_hook_feats;
  rc = run_hook(e, h, nullptr, nullptr);
  if (rc) return rc;
  DCHECK(e, devrt::d2h(out, e->d_hook_feats, sizeof(float) * 17 * e->c.N2, e->stream));
  return AGZ_OK;
}

// ---- position hooks
static int pos_hook(agz_engine* e, int kind, const agz_position* in, int fmove, int* ival, float* fval) {
  if (!e || !in) return fail(e, AGZ_ERR_ARG, "null argument");
  if (in->to_play != 1 && in->to_play != -1) return fail(e, AGZ_ERR_ARG, "to_play must be +-1");
  DCHECK(e, devrt::h2d(e->d_pos_in, in, sizeof(*in), e->stream));
  HookParams h = hp(kind, 0);
  h.fmove = fmove;
  h.pos_in = e->d_pos_in;
  h.pos_out = e->d_pos_out;
  h.legal_out = e->d_hook_legal;
  h.libs_out = e->d_hook_libs;
  return run_hook(e, h, ival, fval);
}

extern "C" int32_t agz_pos_play_move(agz_engine* e, const agz_position* in, int32_t fmove, agz_position* out) {
  if (!e || !out) return fail(e, AGZ_ERR_ARG, "null argument");
  if (fmove < 0 || fmove >= e->c.A) return fail(e, AGZ_ERR_ARG, "move out of range");
  // note: the reference's `@assert !new_pos.done` (board.jl:462) is vacuous (deepcopy resets `done`, board.jl:304),
  // and test_go.jl:490-491 plays on after two passes, so a done position is accepted here too.
  DCHECK(e, devrt::dmemset(e->d_pos_out, 0, sizeof(agz_position), e->stream));
  int rc = pos_hook(e, HK_POS_PLAY, in, fmove, nullptr, nullptr);
  if (rc) return rc;
  DCHECK(e, devrt::d2h(out, e->d_pos_out, sizeof(*out), e->stream));
  return AGZ_OK;
}

extern "C" int32_t agz_pos_legal_moves(agz_engine* e, const agz_position* in, int8_t* legal) {
  int rc = pos_hook(e, HK_POS_LEGAL, in, 0, nullptr, nullptr);
  if (rc) return rc;
  DCHECK(e, devrt::d2h(legal, e->d_hook_legal, (size_t)e->c.A, e->stream));
  return AGZ_OK;
}

extern "C" int32_t agz_pos_score(agz_engine* e, const agz_position* in, float* score) {
  float f = 0.f;
  int rc = pos_hook(e, HK_POS_SCORE, in, 0, nullptr, &f);
  if (rc) return rc;
  if (score) *score = f;
  return AGZ_OK;
}

extern "C" int32_t agz_pos_liberties(agz_engine* e, const agz_position* in, uint8_t* liberty_cache) {
  int rc = pos_hook(e, HK_POS_LIBS, in, 0, nullptr, nullptr);
  if (rc) return rc;
  DCHECK(e, devrt::d2h(liberty_cache, e->d_hook_libs, (size_t)e->c.N2, e->stream));
  return AGZ_OK;
}

// ------------------------------------------------------------------------------------------- introspection
extern "C" int32_t agz_engine_info(agz_engine* e, int64_t out[4]) {
  if (!e || !out) return fail(e, AGZ_ERR_ARG, "null argument");
  out[0] = e->c.cap;                 // node arena capacity per game
  out[1] = (int64_t)e->bytes_per_node;
  out[2] = e->c.n_games;
  out[3] = e->c.ring_cap;
  return AGZ_OK;
}

extern "C" int32_t agz_selfplay_stats(agz_engine* e, int64_t out[4]) {
  if (!e || !out) return fail(e, AGZ_ERR_ARG, "null argument");
  unsigned long long ctr[CTR_COUNT];
  DCHECK(e, devrt::d2h(ctr, e->v.ctr, sizeof(ctr), e->stream));
  out[0] = (int64_t)ctr[CTR_DUP_LEAVES];
  out[1] = (int64_t)ctr[CTR_PRUNES];
  out[2] = (int64_t)ctr[CTR_POSITIONS];
  out[3] = (int64_t)ctr[CTR_READOUTS];
  return AGZ_OK;
}

extern "C" int32_t agz_kernel_launches(agz_engine* e, int64_t* n) {
  if (!e || !n) return fail(e, AGZ_ERR_ARG, "null argument");
  *n = e->launches;
  return AGZ_OK;
}

extern "C" int32_t agz_set_timing(agz_engine* e, int32_t enabled) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  e->timing = enabled;
  return AGZ_OK;
}

extern "C" int32_t agz_phase_times(agz_engine* e, float ms[AGZ_NKERNELS], int64_t launches[AGZ_NKERNELS], int32_t reset) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  for (int i = 0; i < AGZ_NKERNELS; ++i) {
    if (ms) ms[i] = e->phase_ms[i];
    if (launches) launches[i] = e->phase_launches[i];
  }
  if (reset) {
    memset(e->phase_ms, 0, sizeof(e->phase_ms));
    memset(e->phase_launches, 0, sizeof(e->phase_launches));
  }
  return AGZ_OK;
}

// ------------------------------------------------------------------------------------------- network ABI
extern "C" int32_t agz_net_flops(agz_engine* e, double* per_position, double* per_tower_conv_position) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  const double N2 = (double)e->c.N2, C = e->cfg.filters, A = (double)e->c.A;
  const double conv = 2.0 * 9 * C * C * N2;
  if (per_tower_conv_position) *per_tower_conv_position = conv;
  if (per_position) *per_position = 2.0 * 9 * e->cfg.planes * C * N2 + e->cfg.tower_height * 2.0 * conv + 2.0 * (3 * C * N2) + 2.0 * (256 * N2 + 256) + 2.0 * (A * 2 * N2);
  return AGZ_OK;
}
#if AGZ_CUDA
// After a training step the fp32 master parameters live on the device and the inference paths were refreshed there (train_publish);
// the host copy is brought up to date only when somebody reads or partially overwrites it.
static int sync_host(agz_engine* e) {
  if (!e->host_stale || !e->train) return AGZ_OK;
  char terr[256] = "";
  if (train_store(e->train, e->nn, e->stream, terr, sizeof(terr))) return fail(e, AGZ_ERR_CUDA, "%s", terr);
  e->host_stale = false;
  return AGZ_OK;
}

extern "C" size_t agz_net_param_count(agz_engine* e, int32_t chain) { return e ? nn_param_count(e->nn, chain) : 0; }
extern "C" size_t agz_net_bn_count(agz_engine* e, int32_t chain) { return e ? nn_bn_count(e->nn, chain) : 0; }

extern "C" int32_t agz_net_set_params(agz_engine* e, int32_t chain, const float* flat, size_t n) {
  if (!e || !flat) return fail(e, AGZ_ERR_ARG, "null argument");
  if (chain < 0 || chain > 2) return fail(e, AGZ_ERR_ARG, "chain must be 0..2");
  if (n != nn_param_count(e->nn, chain)) return fail(e, AGZ_ERR_ARG, "chain %d expects %zu parameters, got %zu", chain, nn_param_count(e->nn, chain), n);
  cudaSetDevice(e->cfg.device);
  { int rc = sync_host(e); if (rc) return rc; }   // the other chains' host copies must be current before the network is re-committed
  e->train_loaded = false;   // the device master copy of the training path is stale now
  return nn_set_params(e->nn, chain, flat, n) ? fail(e, AGZ_ERR_ARG, "nn_set_params failed") : AGZ_OK;
}

// ---- training step: _train / losses (neural_net.jl:75-101), Momentum (train.jl:54)
static int train_prepare(agz_engine* e, int32_t B) {
  if (B < 1 || B > 4096) return fail(e, AGZ_ERR_ARG, "batch must be in [1, 4096]");
  cudaSetDevice(e->cfg.device);
  char terr[256] = "";
  if (e->train && train_max_batch(e->train) < B)   // the momentum lives in the training state: it is not re-created for a larger batch
    return fail(e, AGZ_ERR_ARG, "batch %d exceeds the training state's batch %d (the first training step fixes it)", B, train_max_batch(e->train));
  if (!e->train) {
    e->train = train_create(e->nn, B < 32 ? 32 : B, terr, sizeof(terr));
    if (!e->train) return fail(e, AGZ_ERR_CUDA, "%s", terr);
    e->train_loaded = false;
  }
  if (!e->train_loaded) {
    if (train_load(e->train, e->nn, e->stream, terr, sizeof(terr))) return fail(e, AGZ_ERR_ARG, "%s", terr);
    e->train_loaded = true;
  }
  return AGZ_OK;
}

// hand the updated parameters to the inference paths on the device (fold + fp16 reorder); the host copy is synchronised lazily
static int train_finish(agz_engine* e) {
  char terr[256] = "";
  e->launches += 60 + 40 * (long long)e->cfg.tower_height;
  if (train_publish(e->train, e->nn, e->stream, terr, sizeof(terr))) return fail(e, AGZ_ERR_CUDA, "%s", terr);
  e->launches += 3 + 2 * (1 + 2 * (long long)e->cfg.tower_height);
  e->host_stale = true;
  return AGZ_OK;
}

static int train_allreduce_cb(void* ctx, float* buf, size_t n, cudaStream_t st) { return replay_allreduce_sum((ReplayState*)ctx, buf, n, st); }

extern "C" int32_t agz_train_step(agz_engine* e, const int8_t* boards_hist, const int8_t* to_play, const float* pis, const int8_t* zs, int32_t B,
                                  float lr, float momentum, float* loss_out) {
  if (!e || !boards_hist || !to_play || !pis || !zs) return fail(e, AGZ_ERR_ARG, "null argument");
  int rc = train_prepare(e, B);
  if (rc) return rc;
  char terr[256] = "";
  float* d_feats = nullptr;
  if (cudaMalloc((void**)&d_feats, (size_t)B * 17 * e->c.N2 * sizeof(float)) != cudaSuccess) return fail(e, AGZ_ERR_CUDA, "feature buffer allocation failed");
  rc = engine_host_features(e->c, boards_hist, to_play, B, nullptr, d_feats, e->stream);
  if (rc) { cudaFree(d_feats); return fail(e, AGZ_ERR_CUDA, "feature kernel: %s", cudaGetErrorString((cudaError_t)rc)); }
  std::vector<float> z((size_t)B);
  for (int b = 0; b < B; ++b) z[b] = (float)zs[b];
  // with an initialised NCCL communicator (agz_nccl_init) the step is data parallel: every rank calls it with its own minibatch
  const int world = replay_world(e->replay);
  rc = train_step(e->train, d_feats, pis, z.data(), B, lr, momentum, loss_out, e->stream, terr, sizeof(terr), world,
                  world > 1 ? train_allreduce_cb : (train_allreduce_fn) nullptr, e->replay);
  cudaFree(d_feats);
  if (rc) return fail(e, AGZ_ERR_CUDA, "%s", terr);
  return train_finish(e);
}

// get_replay_batch + _train (src/train.jl:66-70) without leaving the device: the draw (replay.cu: keyed permutation + gather kernel),
// the feature planes, the step and the hand-over of the new parameters to the self-play path; only the loss travels to the host
extern "C" int32_t agz_train_step_from_replay(agz_engine* e, int32_t batch, uint64_t seed, float lr, float momentum, float* loss_out) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  if (!e->replay) return fail(e, AGZ_ERR_ARG, "no replay ring (call agz_replay_gather first)");
  int rc = train_prepare(e, batch);
  if (rc) return rc;
  char terr[256] = "";
  const int world = replay_world(e->replay);
  // data parallel: the rings of all ranks hold the same tuples, so every rank draws with its own key
  const uint64_t rseed = seed ^ ((uint64_t)e->c.rank * 0x9E3779B97F4A7C15ull);
  const unsigned char* stage = nullptr;
  rc = replay_sample_device(e->replay, batch, world > 1 ? rseed : seed, e->stream, &stage, nullptr, terr, sizeof(terr));
  if (rc) return fail(e, rc, "%s", terr);
  e->launches += 2;
  rc = train_step_from_tuples(e->train, stage, replay_stride(e->replay), batch, lr, momentum, loss_out, e->stream, terr, sizeof(terr), world,
                              world > 1 ? train_allreduce_cb : (train_allreduce_fn) nullptr, e->replay);
  if (rc) return fail(e, AGZ_ERR_CUDA, "%s", terr);
  return train_finish(e);
}

extern "C" int32_t agz_train_read_grads(agz_engine* e, int32_t chain, float* grads, size_t n) {
  if (!e || !grads || !e->train) return fail(e, AGZ_ERR_ARG, "no training step has run");
  cudaSetDevice(e->cfg.device);
  if (train_read_grads(e->train, chain, grads, n, e->stream)) return fail(e, AGZ_ERR_ARG, "bad chain / size");
  return AGZ_OK;
}

extern "C" int32_t agz_net_get_params(agz_engine* e, int32_t chain, float* flat, size_t n) {
  if (!e || !flat) return fail(e, AGZ_ERR_ARG, "null argument");
  if (chain < 0 || chain > 2) return fail(e, AGZ_ERR_ARG, "chain must be 0..2");
  cudaSetDevice(e->cfg.device);
  { int rc = sync_host(e); if (rc) return rc; }
  if (n != nn_param_count(e->nn, chain) || e->nn->hparams[chain].size() != n) return fail(e, AGZ_ERR_ARG, "chain %d has %zu parameters set, asked for %zu", chain, e->nn->hparams[chain].size(), n);
  memcpy(flat, e->nn->hparams[chain].data(), n * sizeof(float));
  return AGZ_OK;
}

extern "C" int32_t agz_net_get_bn_stats(agz_engine* e, int32_t chain, float* mu, float* sigma, size_t n_each, int32_t* bn_mode) {
  if (!e || !mu || !sigma) return fail(e, AGZ_ERR_ARG, "null argument");
  if (chain < 0 || chain > 2) return fail(e, AGZ_ERR_ARG, "chain must be 0..2");
  if (n_each != nn_bn_count(e->nn, chain)) return fail(e, AGZ_ERR_ARG, "chain %d has %zu BatchNorm channels, got %zu", chain, nn_bn_count(e->nn, chain), n_each);
  cudaSetDevice(e->cfg.device);
  { int rc = sync_host(e); if (rc) return rc; }
  memcpy(mu, e->nn->hmu[chain].data(), n_each * sizeof(float));
  memcpy(sigma, e->nn->hsigma[chain].data(), n_each * sizeof(float));
  if (bn_mode) *bn_mode = e->nn->bn_mode[chain];
  return AGZ_OK;
}

extern "C" int32_t agz_net_set_bn_stats(agz_engine* e, int32_t chain, const float* mu, const float* sigma, size_t n_each, int32_t bn_mode) {
  if (!e || !mu || !sigma) return fail(e, AGZ_ERR_ARG, "null argument");
  if (chain < 0 || chain > 2) return fail(e, AGZ_ERR_ARG, "chain must be 0..2");
  if (n_each != nn_bn_count(e->nn, chain)) return fail(e, AGZ_ERR_ARG, "chain %d has %zu BatchNorm channels, got %zu", chain, nn_bn_count(e->nn, chain), n_each);
  if (bn_mode != AGZ_BN_VAR_EPS && bn_mode != AGZ_BN_STD) return fail(e, AGZ_ERR_ARG, "bad bn_mode");
  cudaSetDevice(e->cfg.device);
  { int rc = sync_host(e); if (rc) return rc; }
  e->train_loaded = false;   // the device copy of the running statistics kept by the training path is stale now
  return nn_set_bn(e->nn, chain, mu, sigma, n_each, bn_mode) ? fail(e, AGZ_ERR_ARG, "nn_set_bn failed") : AGZ_OK;
}

extern "C" int32_t agz_features(agz_engine* e, const int8_t* boards_hist, const int8_t* to_play, int32_t B, float* out) {
  if (!e || !boards_hist || !to_play || !out || B < 1) return fail(e, AGZ_ERR_ARG, "bad argument");
  cudaSetDevice(e->cfg.device);
  int rc = engine_host_features(e->c, boards_hist, to_play, B, out, nullptr, e->stream);
  if (rc) return fail(e, AGZ_ERR_CUDA, "feature kernel: %s", cudaGetErrorString((cudaError_t)rc));
  e->launches += 1;
  return AGZ_OK;
}

extern "C" int32_t agz_net_forward(agz_engine* e, int32_t evaluator, const int8_t* boards_hist, const int8_t* to_play, int32_t B, float* pi, float* v) {
  if (!e || !boards_hist || !to_play || !pi || !v || B < 1) return fail(e, AGZ_ERR_ARG, "bad argument");
  if (evaluator != AGZ_EVAL_NN_TC && evaluator != AGZ_EVAL_NN_F32) return fail(e, AGZ_ERR_ARG, "evaluator must be a network");
  cudaSetDevice(e->cfg.device);
  char nerr[256] = "";
  if (!nn_ready(e->nn) && nn_commit(e->nn, e->stream, nerr, sizeof(nerr))) return fail(e, AGZ_ERR_ARG, "network not ready: %s", nerr);
  const int maxb = e->c.n_games * e->c.pmax;
  const size_t A = e->c.A, N2 = e->c.N2;
  for (int b0 = 0; b0 < B; b0 += maxb) {
    const int nb = std::min(maxb, B - b0);
    int rc;
    if (evaluator == AGZ_EVAL_NN_F32) {
      rc = sync_host(e);
      if (rc) return rc;
      rc = engine_host_features(e->c, boards_hist + (size_t)b0 * 8 * N2, to_play + b0, nb, nullptr, e->d_feats_f32, e->stream);
      if (rc) return fail(e, AGZ_ERR_CUDA, "feature kernel: %s", cudaGetErrorString((cudaError_t)rc));
      rc = nn_forward_f32(e->nn, e->d_feats_f32, nb, e->d_eval_pi, e->d_eval_v, e->stream);
      if (rc) return fail(e, AGZ_ERR_CUDA, "nn_forward_f32: %s", cudaGetErrorString((cudaError_t)rc));
      e->launches += 1 + nn_f32_launches_per_forward(e->nn);
    } else {
      rc = engine_host_features_tc(e->c, e->nn, boards_hist + (size_t)b0 * 8 * N2, to_play + b0, nb, e->stream);
      if (rc) return fail(e, AGZ_ERR_CUDA, "tc feature kernel: %s", cudaGetErrorString((cudaError_t)rc));
      rc = nn_forward_tc(e->nn, nb, e->d_eval_pi, e->d_eval_v, e->stream, nerr, sizeof(nerr));
      if (rc) return fail(e, AGZ_ERR_CUDA, "nn_forward_tc: %s", nerr);
      e->launches += 1 + nn_tc_launches_per_forward(e->nn);
    }
    DCHECK(e, devrt::d2h(pi + (size_t)b0 * A, e->d_eval_pi, sizeof(float) * A * nb, e->stream));
    DCHECK(e, devrt::d2h(v + b0, e->d_eval_v, sizeof(float) * nb, e->stream));
  }
  return AGZ_OK;
}

// Self-test of the reciprocal-table quotients of select_leaf (tree.cuh) against the IEEE divisions they replace: every thread draws
// divisors d in [1, rcp_n) (half of them below 2048) and numerators (fp32 w of every magnitude, fp64 x = cu * p as the score builds
// them, and fp64 x with random exponents) from a counter-based stream and counts disagreements.
__global__ void division_selftest_kernel(const double* __restrict__ rcp, int rcp_n, unsigned long long n, uint64_t seed, unsigned long long* bad) {
  unsigned long long bad32 = 0, bad64 = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const U4 a = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0x51u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const U4 b = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0x52u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    int d = (a.x & 1u) ? 1 + (int)((a.x >> 1) % 2047u) : 1 + (int)((a.x >> 1) % (uint32_t)(rcp_n - 1));
    const float den = (float)d;
    const double dd = (double)den, r = rcp[d];
    // fp32 numerator: random sign / mantissa, exponent in [-100, 20]
    const float w = __uint_as_float((a.y & 0x807fffffu) | ((27u + a.z % 121u) << 23));
    const float q_fast = (float)__dmul_rn((double)w, r), q_ref = __fdiv_rn(w, den);
    if (__float_as_uint(q_fast) != __float_as_uint(q_ref)) ++bad32;
    // fp64 numerators: c_puct * sqrt(1 + N) * p, and a raw random double
    const float pr = __uint_as_float((b.x & 0x007fffffu) | ((90u + b.y % 38u) << 23));   // p in [2^-37, 2)
    const double cu = __dmul_rn(0.96, (double)__fsqrt_rn(1.0f + (float)(b.z % 100000u)));
    double xs[2];
    xs[0] = __dmul_rn(cu, (double)pr);
    xs[1] = __longlong_as_double((long long)((((unsigned long long)(a.w & 0xfffffu)) << 32 | b.w) | ((unsigned long long)(900u + b.z % 200u) << 52)));
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const double x = xs[t], q0 = __dmul_rn(x, r);
      const double u_fast = __fma_rn(__fma_rn(-q0, dd, x), r, q0), u_ref = __ddiv_rn(x, dd);
      if (__double_as_longlong(u_fast) != __double_as_longlong(u_ref)) ++bad64;
    }
  }
  if (bad32) atomicAdd(bad, bad32);
  if (bad64) atomicAdd(bad + 1, bad64);
}

extern "C" int32_t agz_selftest_division(agz_engine* e, uint64_t n_samples, uint64_t seed, uint64_t mismatches[2]) {
  if (!e || !mismatches) return fail(e, AGZ_ERR_ARG, "null argument");
  cudaSetDevice(e->cfg.device);
  unsigned long long* d_bad = nullptr;
  if (cudaMalloc((void**)&d_bad, 2 * sizeof(unsigned long long)) != cudaSuccess) return fail(e, AGZ_ERR_CUDA, "allocation failed");
  cudaMemsetAsync(d_bad, 0, 2 * sizeof(unsigned long long), e->stream);
  division_selftest_kernel<<<148 * 8, 256, 0, e->stream>>>(e->v.rcp, e->v.rcp_n, n_samples, seed, d_bad);
  e->launches += 1;
  unsigned long long h[2] = {0, 0};
  int rc = devrt::d2h(h, d_bad, sizeof(h), e->stream);
  cudaFree(d_bad);
  if (rc) return fail(e, AGZ_ERR_CUDA, "division self-test: %s", devrt::last_error_string(rc));
  mismatches[0] = h[0];
  mismatches[1] = h[1];
  return AGZ_OK;
}

// Test hook behind the network parity tests: one forward of at most one internal batch with the intermediate values exposed.
extern "C" int32_t agz_net_forward_debug(agz_engine* e, int32_t evaluator, const int8_t* boards_hist, const int8_t* to_play, int32_t B, int32_t n_blocks,
                                         float* pi, float* v, float* logits, float* v_pre, float* trunk) {
  if (!e || !boards_hist || !to_play || B < 1) return fail(e, AGZ_ERR_ARG, "bad argument");
  if (evaluator != AGZ_EVAL_NN_TC && evaluator != AGZ_EVAL_NN_F32) return fail(e, AGZ_ERR_ARG, "evaluator must be a network");
  if (B > e->c.n_games * e->c.pmax) return fail(e, AGZ_ERR_ARG, "the debug forward takes at most n_games * max_parallel = %d positions", e->c.n_games * e->c.pmax);
  cudaSetDevice(e->cfg.device);
  char nerr[256] = "";
  int rc = sync_host(e);
  if (rc) return rc;
  if (!nn_ready(e->nn) && nn_commit(e->nn, e->stream, nerr, sizeof(nerr))) return fail(e, AGZ_ERR_ARG, "network not ready: %s", nerr);
  const size_t A = e->c.A, N2 = e->c.N2, C = e->cfg.filters;
  const bool full = n_blocks < 0 || n_blocks >= e->cfg.tower_height;
  NNDebug dbg{n_blocks, nullptr, nullptr};
  if (trunk && cudaMalloc((void**)&dbg.trunk, (size_t)B * C * N2 * sizeof(float)) != cudaSuccess) return fail(e, AGZ_ERR_CUDA, "debug buffer allocation failed");
  if ((logits || v_pre) && cudaMalloc((void**)&dbg.raw, (size_t)B * (A + 1) * sizeof(float)) != cudaSuccess) { cudaFree(dbg.trunk); return fail(e, AGZ_ERR_CUDA, "debug buffer allocation failed"); }
  if (evaluator == AGZ_EVAL_NN_F32) {
    rc = engine_host_features(e->c, boards_hist, to_play, B, nullptr, e->d_feats_f32, e->stream);
    if (!rc) rc = nn_forward_f32(e->nn, e->d_feats_f32, B, e->d_eval_pi, e->d_eval_v, e->stream, nullptr, &dbg);
  } else {
    rc = engine_host_features_tc(e->c, e->nn, boards_hist, to_play, B, e->stream);
    if (!rc) rc = nn_forward_tc(e->nn, B, e->d_eval_pi, e->d_eval_v, e->stream, nerr, sizeof(nerr), nullptr, -1, nullptr, nullptr, &dbg);
  }
  e->launches += 3 + 2 * (long long)e->cfg.tower_height;
  if (!rc) rc = devrt::sync(e->stream);
  if (!rc && full && pi) rc = devrt::d2h(pi, e->d_eval_pi, sizeof(float) * A * B, e->stream);
  if (!rc && full && v) rc = devrt::d2h(v, e->d_eval_v, sizeof(float) * B, e->stream);
  if (!rc && trunk) rc = devrt::d2h(trunk, dbg.trunk, sizeof(float) * B * C * N2, e->stream);
  if (!rc && full && dbg.raw) {
    std::vector<float> raw((size_t)B * (A + 1));
    rc = devrt::d2h(raw.data(), dbg.raw, raw.size() * sizeof(float), e->stream);
    for (int b = 0; b < B && !rc; ++b) {
      if (logits) memcpy(logits + (size_t)b * A, raw.data() + (size_t)b * (A + 1), A * sizeof(float));
      if (v_pre) v_pre[b] = raw[(size_t)b * (A + 1) + A];
    }
  }
  cudaFree(dbg.trunk);
  cudaFree(dbg.raw);
  if (rc) return fail(e, AGZ_ERR_CUDA, "debug forward failed: %s", nerr[0] ? nerr : devrt::last_error_string(rc));
  return AGZ_OK;
}

extern "C" int32_t agz_trace_read(agz_engine* e, uint64_t* out, int32_t max_records, int32_t* n_out, int32_t reset) {
  if (!e || !n_out) return fail(e, AGZ_ERR_ARG, "null argument");
  *n_out = 0;
  if (!e->d_trace) return AGZ_OK;
  cudaSetDevice(e->cfg.device);
  DCHECK(e, (int)cudaDeviceSynchronize());
  unsigned long long n = 0;
  DCHECK(e, devrt::d2h(&n, e->d_trace, sizeof(n), e->stream));
  if (n > (unsigned long long)e->trace_cap) n = (unsigned long long)e->trace_cap;
  if (n > (unsigned long long)max_records) n = (unsigned long long)(max_records < 0 ? 0 : max_records);
  if (out && n) DCHECK(e, devrt::d2h(out, e->d_trace + 2, (size_t)n * 4 * sizeof(unsigned long long), e->stream));
  *n_out = (int32_t)n;
  if (reset) DCHECK(e, devrt::dmemset(e->d_trace, 0, sizeof(unsigned long long), e->stream));
  return AGZ_OK;
}

extern "C" int32_t agz_nccl_unique_id(uint8_t id_out[128]) { return replay_unique_id(id_out) ? fail(nullptr, AGZ_ERR_NCCL, "ncclGetUniqueId failed") : AGZ_OK; }

extern "C" int32_t agz_nccl_init(agz_engine* e, const uint8_t id[128]) {
  if (!e || !id) return fail(e, AGZ_ERR_ARG, "null argument");
  cudaSetDevice(e->cfg.device);
  char rerr[256] = "";
  if (!e->replay) e->replay = replay_create(e->c, e->replay_cap, rerr, sizeof(rerr));
  if (!e->replay) return fail(e, AGZ_ERR_CUDA, "replay_create: %s", rerr);
  if (replay_nccl_init(e->replay, id, e->c.world, e->c.rank, rerr, sizeof(rerr))) return fail(e, AGZ_ERR_NCCL, "%s", rerr);
  return AGZ_OK;
}

extern "C" int32_t agz_replay_gather(agz_engine* e, int64_t* n_tuples_total) {
  if (!e) return fail(nullptr, AGZ_ERR_ARG, "null engine");
  cudaSetDevice(e->cfg.device);
  char rerr[256] = "";
  if (!e->replay) e->replay = replay_create(e->c, e->replay_cap, rerr, sizeof(rerr));
  if (!e->replay) return fail(e, AGZ_ERR_CUDA, "replay_create: %s", rerr);
  long long nl = 0;
  int rc = replay_gather(e->replay, e->c, e->v, e->smem_per_warp, e->stream, n_tuples_total, &nl, rerr, sizeof(rerr));
  e->launches += nl;
  if (rc) return fail(e, rc, "%s", rerr);
  return AGZ_OK;
}

extern "C" int32_t agz_replay_info(agz_engine* e, int64_t out[5]) {
  if (!e || !out) return fail(e, AGZ_ERR_ARG, "null argument");
  if (!e->replay) return fail(e, AGZ_ERR_ARG, "no replay ring (call agz_replay_gather first)");
  replay_info(e->replay, out);
  return AGZ_OK;
}

extern "C" int32_t agz_replay_read(agz_engine* e, int64_t first, int32_t count, int8_t* boards, int8_t* to_play, float* pis, int8_t* zs) {
  if (!e || !e->replay) return fail(e, AGZ_ERR_ARG, "no replay ring (call agz_replay_gather first)");
  cudaSetDevice(e->cfg.device);
  char rerr[256] = "";
  int rc = replay_read(e->replay, e->c, first, count, boards, to_play, pis, zs, e->stream, rerr, sizeof(rerr));
  if (rc) return fail(e, rc, "%s", rerr);
  return AGZ_OK;
}

extern "C" int32_t agz_replay_sample(agz_engine* e, int32_t batch, uint64_t seed, int8_t* boards, int8_t* to_play, float* pis, int8_t* zs, int64_t* indices) {
  if (!e || !e->replay) return fail(e, AGZ_ERR_ARG, "no replay ring (call agz_replay_gather first)");
  cudaSetDevice(e->cfg.device);
  char rerr[256] = "";
  int rc = replay_sample(e->replay, e->c, batch, seed, boards, to_play, pis, zs, indices, e->stream, rerr, sizeof(rerr));
  if (rc) return fail(e, rc, "%s", rerr);
  return AGZ_OK;
}

extern "C" int32_t agz_replay_sample_hist(agz_engine* e, int32_t batch, uint64_t seed, int8_t* boards_hist, int8_t* to_play, float* pis, int8_t* zs, int64_t* indices) {
  if (!e || !e->replay) return fail(e, AGZ_ERR_ARG, "no replay ring (call agz_replay_gather first)");
  cudaSetDevice(e->cfg.device);
  char rerr[256] = "";
  int rc = replay_sample(e->replay, e->c, batch, seed, nullptr, to_play, pis, zs, indices, e->stream, rerr, sizeof(rerr), boards_hist);
  if (rc) return fail(e, rc, "%s", rerr);
  return AGZ_OK;
}
#else
extern "C" int32_t agz_replay_sample_hist(agz_engine* e, int32_t, uint64_t, int8_t*, int8_t*, float*, int8_t*, int64_t*) { return fail(e, AGZ_ERR_NCCL, "not in the emulation build"); }
extern "C" int32_t agz_replay_info(agz_engine* e, int64_t*) { return fail(e, AGZ_ERR_NCCL, "not in the emulation build"); }
extern "C" size_t agz_net_param_count(agz_engine*, int32_t) { return 0; }
extern "C" size_t agz_net_bn_count(agz_engine*, int32_t) { return 0; }
extern "C" int32_t agz_net_set_params(agz_engine* e, int32_t, const float*, size_t) { return fail(e, AGZ_ERR_CUDA, "no network in the emulation build"); }
extern "C" int32_t agz_net_set_bn_stats(agz_engine* e, int32_t, const float*, const float*, size_t, int32_t) { return fail(e, AGZ_ERR_CUDA, "no network in the emulation build"); }
extern "C" int32_t agz_features(agz_engine* e, const int8_t*, const int8_t*, int32_t, float*) { return fail(e, AGZ_ERR_CUDA, "not in the emulation build"); }
extern "C" int32_t agz_train_step(agz_engine* e, const int8_t*, const int8_t*, const float*, const int8_t*, int32_t, float, float, float*) { return fail(e, AGZ_ERR_CUDA, "no network in the emulation build"); }
extern "C" int32_t agz_train_step_from_replay(agz_engine* e, int32_t, uint64_t, float, float, float*) { return fail(e, AGZ_ERR_CUDA, "no network in the emulation build"); }
extern "C" int32_t agz_train_read_grads(agz_engine* e, int32_t, float*, size_t) { return fail(e, AGZ_ERR_CUDA, "no network in the emulation build"); }
extern "C" int32_t agz_net_get_params(agz_engine* e, int32_t, float*, size_t) { return fail(e, AGZ_ERR_CUDA, "no network in the emulation build"); }
extern "C" int32_t agz_net_get_bn_stats(agz_engine* e, int32_t, float*, float*, size_t, int32_t*) { return fail(e, AGZ_ERR_CUDA, "no network in the emulation build"); }
extern "C" int32_t agz_net_forward(agz_engine* e, int32_t, const int8_t*, const int8_t*, int32_t, float*, float*) { return fail(e, AGZ_ERR_CUDA, "no network in the emulation build"); }
extern "C" int32_t agz_trace_read(agz_engine*, uint64_t*, int32_t, int32_t* n_out, int32_t) { if (n_out) *n_out = 0; return AGZ_OK; }
extern "C" int32_t agz_selftest_division(agz_engine* e, uint64_t, uint64_t, uint64_t*) { return fail(e, AGZ_ERR_CUDA, "not in the emulation build"); }
extern "C" int32_t agz_net_forward_debug(agz_engine* e, int32_t, const int8_t*, const int8_t*, int32_t, int32_t, float*, float*, float*, float*, float*) { return fail(e, AGZ_ERR_CUDA, "no network in the emulation build"); }
extern "C" int32_t agz_nccl_unique_id(uint8_t*) { return AGZ_ERR_NCCL; }
extern "C" int32_t agz_nccl_init(agz_engine* e, const uint8_t*) { return fail(e, AGZ_ERR_NCCL, "not in the emulation build"); }
extern "C" int32_t agz_replay_gather(agz_engine* e, int64_t*) { return fail(e, AGZ_ERR_NCCL, "not in the emulation build"); }
extern "C" int32_t agz_replay_read(agz_engine* e, int64_t, int32_t, int8_t*, int8_t*, float*, int8_t*) { return fail(e, AGZ_ERR_NCCL, "not in the emulation build"); }
extern "C" int32_t agz_replay_sample(agz_engine* e, int32_t, uint64_t, int8_t*, int8_t*, float*, int8_t*, int64_t*) { return fail(e, AGZ_ERR_NCCL, "not in the emulation build"); }
#endif
