/*
 * agz.h -- C ABI of libagz: a B200-native AlphaGo-Zero self-play engine that sits behind
 * tejank10/AlphaGo.jl's surface (GoEnv, NeuralNet, MCTSPlayer, selfplay).
 *
 * The reference has no FFI seam of its own: its boundary is the Julia-level surface
 * (SURVEY.md section 8b).  Each entry point below names the reference function it replaces
 * (paths relative to the reference repo); the Julia `ccall` stubs a maintainer would add are in
 * INTEGRATION.md.  Conventions:
 *   - plain C types only; the caller owns every host buffer, the library owns all device memory;
 *   - every function returns an int32 status (AGZ_OK = 0); agz_last_error() describes the last failure;
 *   - one engine per GPU, driven by one host thread; calls are blocking; streams/graphs are private;
 *   - indices are 0-based: flat move f = N*col + row (the reference's 1-based fmove minus 1), pass = N*N (Go; Gomoku has no pass
 *     and N*N actions);
 *   - arrays that the reference holds column-major (W x H x C x B, A x B) keep that memory order.
 *   - there is NO CPU fallback: agz_engine_create fails with AGZ_ERR_CUDA when no sm_100 device is present.
 */
#ifndef AGZ_H
#define AGZ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGZ_OK 0
#define AGZ_ERR_ILLEGAL_MOVE 1 /* IllegalMove            (src/AlphaGo.jl:8; thrown board.jl:265,470)  */
#define AGZ_ERR_ASSERT 2       /* AssertionError         (@assert sites: mcts.jl:190,196; mcts_play.jl:68,127) */
#define AGZ_ERR_CUDA 3
#define AGZ_ERR_NCCL 4
#define AGZ_ERR_ARG 5
#define AGZ_ERR_CAPACITY 6     /* a game's node arena is full (raise nodes_per_game) */

#define AGZ_MAX_N 19
#define AGZ_MAX_POINTS 361
#define AGZ_MAX_ACTIONS 362
#define AGZ_HIST 7             /* board deltas kept per position (board.jl:505-506) */

/* games behind the reference's Position interface (src/game/env.jl): Go(n) = GoEnv(n) (src/game/go/go.jl:1-26) and
 * GomokuEnv(board_size, connect_row) (src/game/gomoku/gomoku.jl:1-19: action_space = N^2, no pass, every empty point legal, the
 * game ends on n_in_row stones in a line or a full board; score = the winner's colour, src/game/gomoku/board.jl:93-193) */
#define AGZ_GAME_GO 0
#define AGZ_GAME_GOMOKU 1

/* evaluator kinds */
#define AGZ_EVAL_DUMMY 0       /* fixed priors + value: DummyNet of test/test_mcts_player.jl:10-32; BASELINE config 5 */
#define AGZ_EVAL_NN_TC 1       /* residual tower on tcgen05 tensor cores (fp16 operands, fp32 accumulate) */
#define AGZ_EVAL_NN_F32 2      /* same network in fp32 SIMT kernels (on-device cross-check, slow) */

/* BatchNorm statistics conventions */
#define AGZ_BN_VAR_EPS 0       /* Flux 0.10.4: gamma*(x-mu)/sqrt(sigma2+1e-5)+beta */
#define AGZ_BN_STD 1           /* shipped models/agz_*.bson: gamma*(x-mu)/sigma+beta (moving std) */

#define AGZ_CHAIN_BASE 0       /* nn.base_net  (src/neural_net.jl:19-21) */
#define AGZ_CHAIN_VALUE 1      /* nn.value     (src/neural_net.jl:23-26) */
#define AGZ_CHAIN_POLICY 2     /* nn.policy    (src/neural_net.jl:28-30) */

typedef struct agz_engine agz_engine;

/* Everything the reference spreads over GoEnv (go.jl:10-25), MCTSRules (mcts.jl:20-24), the mcts.jl
 * globals (:11,:13), MCTSPlayer kwargs (mcts_play.jl:17-19), tree_search! (:73) and selfplay (selfplay.jl:9). */
typedef struct agz_config {
  int32_t board_n;            /* GoEnv.N */
  int32_t planes;             /* 17 */
  int32_t filters;            /* 256 */
  int32_t tower_height;       /* NeuralNet(env; tower_height) */
  double c_puct;              /* 0.96 */
  double noise_weight;        /* 0.25 */
  double noise_alpha;         /* Float32(0.03*361/A) widened to double */
  int32_t max_game_length;    /* N^2*7/5 */
  int32_t tau_threshold;      /* (N^2/12)/2*2, or -1 in two_player_mode */
  int32_t parallel_readouts;  /* 8 */
  int32_t max_parallel;       /* scratch is sized for this many leaves per tree_search! (>= parallel_readouts) */
  float komi;                 /* 7.5 */
  double resign_threshold;    /* -0.9 */
  double resign_disable_frac; /* 0.05 */
  int32_t n_games;            /* concurrent games (slots) on this GPU */
  int32_t readouts;           /* num_readouts per move */
  int32_t nodes_per_game;     /* node arena capacity per slot (0 = default, see agz_engine_info) */
  uint64_t seed;              /* RNG spec: oracle/rng.py */
  int32_t device;             /* CUDA device ordinal */
  int32_t world_size;         /* game sharding: global slot s lives on rank s % world_size */
  int32_t rank;
  int32_t record_ring;        /* finished-game records kept on device until harvested (0 = 2*n_games) */
  int32_t evaluator;          /* AGZ_EVAL_* */
  int32_t inject_noise;       /* 1 (selfplay.jl:23); 0 for play/evaluate-style search */
  int32_t game;               /* AGZ_GAME_GO | AGZ_GAME_GOMOKU (GoEnv / GomokuEnv) */
  int32_t n_in_row;           /* GomokuEnv.n_in_row (5); ignored for Go */
} agz_config;

/* GoPosition (board.jl:271-306) as a POD.  board[f] with f = N*col+row, values -1 W / 0 / +1 B.
 * hist[k] is the board k+1 moves ago (what features.jl:7-12 reconstructs from board_deltas); n_hist <= 7. */
typedef struct agz_position {
  int8_t board[AGZ_MAX_POINTS];
  int8_t hist[AGZ_HIST][AGZ_MAX_POINTS];
  int32_t n_hist;
  int32_t n;                  /* moves played so far */
  int32_t to_play;            /* +1 / -1 */
  int32_t ko;                 /* flat point or -1 */
  int32_t last_move_pass;     /* recent[end].move == nothing */
  int32_t done;               /* two consecutive passes */
  int32_t caps[2];            /* captures by B, W */
  float komi;
} agz_position;

/* One MCTSNode (mcts.jl:41-82) read back for inspection. */
typedef struct agz_node_view {
  int32_t parent;             /* node id or -1 */
  int32_t fmove;              /* move that led here, -1 for the root of the arena */
  int32_t to_play, n, ko, is_expanded, done, last_move_pass;
  float N, W;                 /* this node's own visit count / value sum (held in its parent's arrays) */
  float child_N[AGZ_MAX_ACTIONS], child_W[AGZ_MAX_ACTIONS], child_prior[AGZ_MAX_ACTIONS];
  int32_t children[AGZ_MAX_ACTIONS]; /* node id or -1 */
  int8_t legal[AGZ_MAX_ACTIONS];     /* all_legal_moves (board.jl:393-424) */
  int8_t board[AGZ_MAX_POINTS];
  double action_score[AGZ_MAX_ACTIONS]; /* child_action_score (mcts.jl:86-87), Float64 like the reference */
} agz_node_view;

/* Header of a finished game (what selfplay returns in the MCTSPlayer: searches_pi, qs, result, result_string). */
typedef struct agz_game_header {
  int64_t game_id;
  int32_t n_moves;
  int32_t result;             /* +1 B / -1 W / 0 */
  int32_t resigned;           /* result_string is "B+R"/"W+R" when set, else from final_score */
  float final_score;          /* score(position) (board.jl:511-533) when not resigned */
  double resign_threshold;
} agz_game_header;

typedef struct agz_progress {
  int64_t moves_played;       /* cumulative since agz_selfplay_start, this rank */
  int64_t games_finished;
  int64_t games_started;
  int64_t positions_evaluated;/* NN batch rows evaluated */
  int64_t readouts;           /* select_leaf calls */
  int64_t path_nodes;         /* sum of path lengths (for the tree-traffic roofline) */
  int32_t games_live;
  int32_t error;              /* first per-game error seen (AGZ_ERR_*) or 0 */
  float step_ms;              /* device time of this agz_selfplay_step call (CUDA events on the engine's stream) */
  int32_t arena_prunes;       /* cumulative: times a game's node arena had to forget its least-visited nodes to fit the next search */
} agz_progress;

/* ---- lifecycle ----------------------------------------------------------------------------- */
int32_t agz_config_default(agz_config* cfg, int32_t board_n);  /* GoEnv(N), MCTSRules(env), MCTSPlayer defaults */
/* the same for either game: GoEnv(N) or GomokuEnv(N, n_in_row) (noise_alpha = Float32(0.03*361/action_space), mcts.jl:22) */
int32_t agz_config_default_game(agz_config* cfg, int32_t game, int32_t board_n, int32_t n_in_row);
int32_t agz_engine_create(const agz_config* cfg, agz_engine** out);
void agz_engine_destroy(agz_engine* e);
const char* agz_last_error(agz_engine* e);                      /* valid until the next call on e (e may be NULL) */
int32_t agz_version(void);

/* Named integer options: every knob that is not part of the reference's own configuration surface (agz_config).  Nothing in the
 * library reads environment variables.  Keys (default):
 *   selfplay.stagger_rounds (0)  read by the next agz_selfplay_start: slot g starts its first game after g*R/n_games rounds, so a
 *                                throughput run reaches the steady state of src/train.jl:56-61 (games at every ply, some finishing
 *                                in every step) after one game length instead of moving through the plies in lock step
 *   dummy.fused_rounds (1)       DummyNet evaluator: all rounds of an agz_selfplay_step call in one launch per game
 *   schedule.pipeline (0)        two half batches on separate streams (tree kernels of one under the network of the other)
 *   replay.capacity (500000)     memory_size of src/train.jl:38; must be set before the replay ring exists
 *   trace.records (0)            kernel timeline trace capacity (agz_trace_read); can be set once
 *   conv.precision (1)           arithmetic of the tensor-core network: 1 = fp16 operands, fp32 accumulation (11 significant bits per
 *                                operand; within 1e-3 of fp32 on the shipped / random-init networks, ~5e-3 on sharp trained-like ones at
 *                                tower_height 19, profiles/r02_nn_error_vs_depth.jsonl); 2 = split precision, every activation and weight
 *                                a pair of fp16 numbers (hi + lo, three MMA passes): fp32-like accuracy at a third of the throughput
 *   conv.fuse_heads (1), conv.stages (6), conv.l2_prefetch (0), conv.pdl (1), conv.max_pairs (0 = all), conv.res_tma (1)
 *                                experiment knobs of the tensor-core convolution (alphago.jl_b200/csrc/nn_tc.cu)
 * Unknown keys and bad values return AGZ_ERR_ARG. */
int32_t agz_set_option(agz_engine* e, const char* key, int64_t value);
int32_t agz_get_option(agz_engine* e, const char* key, int64_t* value);

/* ---- network: NeuralNet (src/neural_net.jl:13-33) -------------------------------------------- */
/* `flat` = the chain's Flux `params` list concatenated, each tensor column-major, in the order
 * save_model writes them (src/train.jl:27-33): Conv(W,b), BatchNorm(beta,gamma), per ResidualBlock
 * W1,b1,W2,b2,beta1,gamma1,beta2,gamma2 (src/resnet.jl:3-5), Dense(W,b). */
int32_t agz_net_set_params(agz_engine* e, int32_t chain, const float* flat, size_t n);
/* BatchNorm running statistics of the chain, in layer order: all mu then all sigma, each n/2 long. */
int32_t agz_net_set_bn_stats(agz_engine* e, int32_t chain, const float* mu, const float* sigma, size_t n_each, int32_t bn_mode);
size_t agz_net_param_count(agz_engine* e, int32_t chain);
size_t agz_net_bn_count(agz_engine* e, int32_t chain);
/* (nn::NeuralNet)(positions) (src/neural_net.jl:57-68): boards_hist is B x 8 x N*N int8 (board k moves ago,
 * flat order), to_play is B int8.  pi is A x B column-major (pi[a + A*b]), v is B.  `evaluator` picks the path. */
int32_t agz_net_forward(agz_engine* e, int32_t evaluator, const int8_t* boards_hist, const int8_t* to_play, int32_t B, float* pi, float* v);
/* Test hook of the network parity tests: the same forward for at most n_games * max_parallel positions with its intermediate values.
 * n_blocks < 0 or >= tower_height: the whole network -- pi (A x B), v (B), logits (A x B, before softmax), v_pre (B, before tanh) and
 * trunk (the tower output, 256 x N*N per position in the reference's W x H x C x B order).  0 <= n_blocks < tower_height: only `trunk`
 * after the stem and n_blocks residual blocks is produced (bisecting a deviation to a block).  Any output pointer may be NULL. */
int32_t agz_net_forward_debug(agz_engine* e, int32_t evaluator, const int8_t* boards_hist, const int8_t* to_play, int32_t B, int32_t n_blocks,
                              float* pi, float* v, float* logits, float* v_pre, float* trunk);
/* get_feats (src/features.jl:24-26): out is N x N x 17 x B column-major (row fastest), values in {-1,0,1}. */
int32_t agz_features(agz_engine* e, const int8_t* boards_hist, const int8_t* to_play, int32_t B, float* out);

/* One optimisation step on a minibatch (the body of _train, src/neural_net.jl:91-97, with the optimiser of src/train.jl:54):
 * train-mode forward (BatchNorm on batch statistics, running statistics moved with momentum 0.1), loss = 0.01 * crossentropy(p, pi)
 * + 0.01 * mse(z, v) + 1e-4 * sum(theta^2) (:75-83), back-propagation, Momentum(lr, momentum): v = momentum * v - lr * grad;
 * theta += v.  Inputs as agz_net_forward / agz_replay_sample: boards_hist B x 8 x N*N, to_play B, pis B x A, zs B.  The updated
 * parameters are what every later forward / self-play call of this engine uses (handed over on the device: BatchNorm fold and fp16
 * weight conversion run there; agz_net_get_params synchronises the host copy on demand); *loss_out = the loss before the update.
 * After agz_nccl_init (world_size > 1) the step is data parallel -- an extension, the reference trains in one process: every rank
 * calls it with its own minibatch; gradients, loss and running statistics are averaged with ncclAllReduce before the update. */
int32_t agz_train_step(agz_engine* e, const int8_t* boards_hist, const int8_t* to_play, const float* pis, const int8_t* zs, int32_t B,
                       float lr, float momentum, float* loss_out);
/* The same step with get_replay_batch (src/train.jl:4-12) in front of it, entirely on the device: `batch` distinct tuples are drawn
 * from the replay ring (the draw of agz_replay_sample for this seed; with world_size > 1 every rank's draw is keyed by its rank),
 * the feature planes are built there and the updated parameters are folded / converted for the self-play path there as well --
 * nothing but the loss crosses the bus.  The first training step of an engine fixes the largest batch (>= 32). */
int32_t agz_train_step_from_replay(agz_engine* e, int32_t batch, uint64_t seed, float lr, float momentum, float* loss_out);
/* gradients of the data loss of the last agz_train_step, chain's Flux params order (test hook) */
int32_t agz_train_read_grads(agz_engine* e, int32_t chain, float* grads, size_t n);
/* current parameters / BatchNorm running statistics of a chain (what save_model writes, src/train.jl:14-35) */
int32_t agz_net_get_params(agz_engine* e, int32_t chain, float* flat, size_t n);
int32_t agz_net_get_bn_stats(agz_engine* e, int32_t chain, float* mu, float* sigma, size_t n_each, int32_t* bn_mode);

/* DummyNet (test/test_mcts_player.jl:10-32): priors NULL = uniform 1/A. */
int32_t agz_set_dummy_evaluator(agz_engine* e, const float* priors, float value);
int32_t agz_set_evaluator(agz_engine* e, int32_t evaluator);

/* ---- self-play: selfplay (src/selfplay.jl:1-45) over all slots -------------------------------- */
/* Global slot s (0 <= s < n_games*world_size) plays game ids s, s+S, s+2S, ... (S = n_games*world_size)
 * while id < total_games; total_games < 0 means play forever (throughput runs). */
int32_t agz_selfplay_start(agz_engine* e, int64_t total_games);
/* Run `rounds` batched tree_search! rounds (select 8 leaves per live game -> features -> network ->
 * incorporate -> per-game move logic).  Does not synchronise with the host unless `progress` is non-NULL. */
int32_t agz_selfplay_step(agz_engine* e, int32_t rounds, agz_progress* progress);
/* Copy up to max_records finished games (oldest first) to the host.  Per record r: headers[r],
 * moves[r*L + t] (flat move of ply t), qs[r*L + t], pis[(r*L + t)*A + a], visits[(r*L + t)*A + a],
 * with L = max_game_length.  Any of moves/qs/pis/visits may be NULL.  Returns the count in *n_out. */
int32_t agz_selfplay_harvest(agz_engine* e, int32_t max_records, agz_game_header* headers, int16_t* moves, float* qs, float* pis, float* visits, int32_t* n_out);
/* Convenience = start(total_games) + step until all finished + harvest (records sorted by game id). */
int32_t agz_selfplay_run(agz_engine* e, int32_t total_games, agz_game_header* headers, int16_t* moves, float* qs, float* pis, float* visits);
/* extract_data / replay buffer append (src/mcts_play.jl:126-139, src/train.jl:58-61): pack the finished games that
 * have not been gathered yet into (board planes, pi, z) tuples and ncclAllGather them over all ranks into the
 * device-resident replay ring; returns the number of tuples now in the ring.  World size 1 = local pack only. */
int32_t agz_replay_gather(agz_engine* e, int64_t* n_tuples_total);
/* Read tuples [first, first+count) of the replay ring: boards count x N*N int8 (position before the move),
 * to_play count, pis count x A, zs count (result from Black's view, board.jl:574). */
int32_t agz_replay_read(agz_engine* e, int64_t first, int32_t count, int8_t* boards, int8_t* to_play, float* pis, int8_t* zs);
/* get_replay_batch (src/train.jl:4-12): `batch` distinct tuples drawn uniformly without replacement from the ring
 * (deterministic in `seed`); same output arrays as agz_replay_read. */
int32_t agz_replay_sample(agz_engine* e, int32_t batch, uint64_t seed, int8_t* boards, int8_t* to_play, float* pis, int8_t* zs, int64_t* indices);
/* Same draw as agz_replay_sample, returning what the network trains on: boards_hist batch x 8 x N*N = the position before the
 * move and the 7 positions before it (oldest repeated), i.e. what get_feats rebuilds from board_deltas (features.jl:7-14). */
int32_t agz_replay_sample_hist(agz_engine* e, int32_t batch, uint64_t seed, int8_t* boards_hist, int8_t* to_play, float* pis, int8_t* zs, int64_t* indices);
/* Replay ring bookkeeping: out[0] = bytes per packed tuple, out[1] = capacity in tuples, out[2] = tuples ever appended,
 * out[3] = tuple bytes appended by the last agz_replay_gather (the payload of all ranks), out[4] = bytes appended since creation. */
int32_t agz_replay_info(agz_engine* e, int64_t out[5]);
/* NCCL bootstrap for world_size > 1: rank 0 fills a 128-byte id, every rank passes the same bytes. */
int32_t agz_nccl_unique_id(uint8_t id_out[128]);
int32_t agz_nccl_init(agz_engine* e, const uint8_t id[128]);

/* ---- two-player matches: evaluate (src/neural_net.jl:103-158) and play (src/play.jl:25-77) over all slots at once --------- */
/* One engine = one player (its network, two_player_mode: create it with tau_threshold = -1 and inject_noise = 0); slot s is
 * that player's tree of game s.  The host keeps one engine per player and alternates them like the reference alternates its
 * two MCTSPlayers.  agz_match_start = initialize_game! on every slot (game_ids key the RNG; NULL = slot index). */
int32_t agz_match_start(agz_engine* e, const int64_t* game_ids);
/* For every slot with active[s] != 0: `while N(root) < current + readouts: tree_search!` (neural_net.jl:121-126), then
 * should_resign -> resigned[s] (scores[s] = score of the root position, what evaluate reads at :150), else pick_move ->
 * moves[s].  Inactive slots get moves[s] = -1.  scores may be NULL. */
int32_t agz_match_search(agz_engine* e, const uint8_t* active, int32_t* moves, int32_t* resigned, float* scores);
/* play_move!(player, move) (mcts_play.jl:26-50) on every slot with moves[s] >= 0; done[s] = is_done(root) (mcts.jl:230-231)
 * with scores[s] = score(root position).  An illegal move leaves its tree unchanged and returns AGZ_ERR_ILLEGAL_MOVE. */
int32_t agz_match_play(agz_engine* e, const int32_t* moves, int32_t* done, float* scores);

/* ---- single-tree hooks: the reference's unit-test surface, one tree per slot ------------------ */
int32_t agz_tree_init(agz_engine* e, int32_t slot, const agz_position* pos, int64_t game_id); /* initialize_game! (mcts_play.jl:110-118) */
int32_t agz_tree_select_leaf(agz_engine* e, int32_t slot, int32_t from_node, int32_t* leaf);  /* select_leaf (mcts.jl:108-138) */
int32_t agz_tree_incorporate(agz_engine* e, int32_t slot, int32_t node, const float* probs, float value); /* incorporate_results! (mcts.jl:188-213), up_to = root */
int32_t agz_tree_backup_value(agz_engine* e, int32_t slot, int32_t node, float value);        /* backup_value! (mcts.jl:215-225) */
int32_t agz_tree_add_virtual_loss(agz_engine* e, int32_t slot, int32_t node);                 /* mcts.jl:149-163 */
int32_t agz_tree_revert_virtual_loss(agz_engine* e, int32_t slot, int32_t node);              /* mcts.jl:165-171 */
int32_t agz_tree_maybe_add_child(agz_engine* e, int32_t slot, int32_t node, int32_t fmove, int32_t* child); /* mcts.jl:140-147 */
int32_t agz_tree_search(agz_engine* e, int32_t slot, int32_t parallel_readouts, int32_t* n_leaves); /* tree_search! (mcts_play.jl:73-98) */
int32_t agz_tree_inject_noise(agz_engine* e, int32_t slot);                                   /* inject_noise! (mcts.jl:233-239) */
int32_t agz_tree_pick_move(agz_engine* e, int32_t slot, int32_t* fmove);                      /* pick_move (mcts_play.jl:52-71) */
int32_t agz_tree_play_move(agz_engine* e, int32_t slot, int32_t fmove);                       /* play_move!(player, c) (mcts_play.jl:26-50) */
int32_t agz_tree_should_resign(agz_engine* e, int32_t slot, double threshold, int32_t* yes);  /* mcts_play.jl:124 */
int32_t agz_tree_root(agz_engine* e, int32_t slot, int32_t* root, int32_t* node_count);
int32_t agz_tree_read_node(agz_engine* e, int32_t slot, int32_t node, agz_node_view* out);
int32_t agz_tree_set_stats(agz_engine* e, int32_t slot, int32_t node, const float* self_N, const float* child_N, const int32_t* n_override); /* set_N! / child_N writes used by the reference tests */
int32_t agz_tree_pending_vlosses(agz_engine* e, int32_t slot, int32_t* pending);              /* assertNoPendingVirtualLosses (test_utils.jl:76-85) */
int32_t agz_tree_read_record(agz_engine* e, int32_t slot, int32_t* n_moves, int16_t* moves, float* qs, float* pis); /* searches_pi, qs */
int32_t agz_tree_node_features(agz_engine* e, int32_t slot, int32_t node, float* out);        /* get_feats(node.position): N x N x 17 */

/* ---- position hooks: src/game/go/board.jl on the device ---------------------------------------- */
int32_t agz_pos_play_move(agz_engine* e, const agz_position* in, int32_t fmove, agz_position* out); /* play_move! (board.jl:451-509); AGZ_ERR_ILLEGAL_MOVE */
int32_t agz_pos_legal_moves(agz_engine* e, const agz_position* in, int8_t* legal);             /* all_legal_moves (board.jl:393-424), A entries */
int32_t agz_pos_score(agz_engine* e, const agz_position* in, float* score);                    /* score (board.jl:511-533) */
int32_t agz_pos_liberties(agz_engine* e, const agz_position* in, uint8_t* liberty_cache);      /* LibertyTracker.liberty_cache (board.jl:99-164), N*N entries */

/* ---- introspection for bench/roofline ---------------------------------------------------------- */
/* Cumulative since agz_selfplay_start: out[0] = leaves that duplicated an earlier leaf of their own tree_search! round and were
 * reverted (revert_visits!, mcts.jl:197-200) although a network row was spent on them, out[1] = arena-pressure prunes,
 * out[2] = network rows evaluated, out[3] = select_leaf calls. */
int32_t agz_selfplay_stats(agz_engine* e, int64_t out[4]);
int32_t agz_kernel_launches(agz_engine* e, int64_t* n);  /* kernels of this library launched since create */
/* out[0] = node arena capacity per game (nodes_per_game, or the default chosen at creation: the worst case max_game_length *
 * (readouts + 2 * parallel) when it fits in 40 % of the free device memory), out[1] = bytes per node, out[2] = n_games,
 * out[3] = finished-record ring capacity */
int32_t agz_engine_info(agz_engine* e, int64_t out[4]);
/* Per-kernel device time (ms, CUDA events on the engine's stream, accumulated while timing is enabled) and launch
 * counts since the last call with reset != 0: [0] select (tree descent + expansion), [1] leaf features,
 * [2] stem conv, [3] tower 3x3 conv (tcgen05), [4] heads, [5] incorporate + move logic. */
#define AGZ_NKERNELS 6
int32_t agz_phase_times(agz_engine* e, float ms[AGZ_NKERNELS], int64_t launches[AGZ_NKERNELS], int32_t reset);
int32_t agz_set_timing(agz_engine* e, int32_t enabled);
/* Kernel timeline trace (debug aid; enabled by agz_set_option(e, "trace.records", n)): every
 * kernel's first and last CTA append {tag | block << 8 | grid << 32, start ns, end ns, SM id} (%globaltimer).  Tags:
 * 1 select, 2 incorporate, 3 leaf features, 4 stem conv, 5 tower conv, 6 tower conv with shortcut, 7 heads, 9 other. */
int32_t agz_trace_read(agz_engine* e, uint64_t* out, int32_t max_records, int32_t* n_out, int32_t reset);
/* Self-test of the arithmetic shortcut in select_leaf's PUCT score (the reference divides by 1 + N(child), mcts.jl:89-92; the
 * kernel multiplies by a table of correctly rounded reciprocals and corrects, alphago.jl_b200/csrc/tree.cuh): n_samples random
 * (numerator, divisor) pairs on the device, mismatches[0] = fp32 quotients, mismatches[1] = fp64 quotients that differ from the
 * IEEE divisions in any bit.  Both must be 0. */
int32_t agz_selftest_division(agz_engine* e, uint64_t n_samples, uint64_t seed, uint64_t mismatches[2]);
/* FLOPs (2*MAC) of one position through the whole network / through one tower 3x3 convolution. */
int32_t agz_net_flops(agz_engine* e, double* per_position, double* per_tower_conv_position);

#ifdef __cplusplus
}
#endif
#endif /* AGZ_H */
