// ops.cuh -- the warp-per-game kernels of the engine, written as functors: op(warp_index, smem).
// k_warps<Op> (engine.cu) runs one functor instance per warp; the CPU test build runs the same functors on
// fibers (tests/emu).  Reference functions each op stands for are named at the op.
#pragma once
#include "../../include/agz.h"
#include "tree.cuh"

namespace agz {

// selfplay.jl:11-14 for every slot: global slot s = rank + world*g plays games s, s+S, s+2S, ...
template <int KA>
struct StartOp {
  Cfg c;
  View v;
  AGZ_DEV void operator()(int g, char* smem) const {
    Warp<KA> w(c, v, g, smem);
    long long id = (long long)c.rank + (long long)c.world * g;
    const long long delay = c.stagger_rounds > 0 ? (long long)g * c.stagger_rounds / c.n_games : 0;
    if (!(c.total_games < 0 || id < c.total_games)) w.st.phase = PH_IDLE;
    else if (delay > 0) {
      w.st.game_id = id;
      w.st.delay = (int32_t)delay;
      w.st.err = 0;
      w.st.nleaf = 0;
      w.st.phase = PH_DELAY;
    } else w.start_game(id);
    w.store_state();
  }
};

// tree_search! part 1 (mcts_play.jl:74-87) / the seed selection of selfplay.jl:18
// OCC = 1: compiled for high occupancy (72 registers, 7 CTAs/SM on 9x9) -- used when there are thousands of trees per GPU
// (MCTS-only config: +17 % readouts/s); OCC = 0: more registers, no spills -- better when a few warps per SM suffice.
template <int KA, int OCC = 0>
struct SelectOp {
  Cfg c;
  View v;
  int only_slot;  // >= 0: hook mode, run this slot with `parallel` leaves regardless of phase
  int parallel;
  int slot0;      // first slot of this launch (half-batch pipelining)
  AGZ_DEV void operator()(int wi, char* smem) const {
    const int g = only_slot >= 0 ? only_slot : slot0 + wi;
    Warp<KA> w(c, v, g, smem);
    w.prefetch = OCC == 0;   // few trees per SM: the descent is a latency chain
    if (only_slot >= 0) {
      w.search_select(parallel, false);
    } else if (w.st.phase == PH_SEED) {
      w.search_select(1, true);
    } else if (w.st.phase == PH_SEARCH || w.st.phase == PH_MATCH_SEARCH) {
      w.search_select(c.parallel, false);
    } else {
      w.st.nleaf = 0;
      w.st.seed_round = 0;
    }
    w.store_hot();
  }
};

// tree_search! part 2 (mcts_play.jl:88-96) + the per-move logic of selfplay.jl:22-43
template <int KA>
struct IncorporateOp {
  Cfg c;
  View v;
  int only_slot;
  int slot0;
  AGZ_DEV void operator()(int wi, char* smem) const {
    const int g = only_slot >= 0 ? only_slot : slot0 + wi;
    Warp<KA> w(c, v, g, smem);
    w.search_incorporate();
    if (only_slot < 0 && w.after_round_due()) {
      w.store_hot();
      simt::sync();
      after_round_cold<KA>(&c, &v, g, smem);
      w.st = v.gs[g];
    }
    w.st.seed_round = 0;
    w.store_hot();
  }
};

// Two-player matches (evaluate, neural_net.jl:103-158; play, play.jl:25-77) over all slots at once: one slot = one game of
// one player's tree; the host alternates the two players' engines.
enum { MK_BEGIN = 1, MK_ARM, MK_PICK, MK_PLAY };

template <int KA>
struct MatchOp {
  Cfg c;
  View v;
  int kind;
  const long long* game_ids;    // BEGIN: game id of every slot (RNG key) or nullptr = slot index
  const unsigned char* active;  // ARM / PICK: slots whose player is to move
  const int* moves_in;          // PLAY: flat move per slot, < 0 = leave the slot alone
  int* out_i;                   // PICK: move (or -1); PLAY: is_done
  int* out_j;                   // PICK: should_resign
  float* out_f;                 // PICK (resigned slots) / PLAY (finished slots): score(root position)
  AGZ_DEV void operator()(int g, char* smem) const {
    Warp<KA> w(c, v, g, smem);
    const int lane = simt::lane();
    switch (kind) {
      case MK_BEGIN: {  // initialize_game!(player) (mcts_play.jl:110-118)
        const long long id = game_ids ? game_ids[g] : (long long)g;
        w.st.game_id = id;
        w.st.game_id_lo = (uint32_t)id;
        w.st.resign_thr = c.resign_threshold;
        w.st.hist_len = 0;
        w.st.target_N = 0.f;
        w.pos.b = 0;
        w.pos.w = 0;
        w.init_root_from_scratch(0, -1, 1, 0);
        w.st.phase = PH_MATCH_WAIT;
        break;
      }
      case MK_ARM:  // current_readouts = N(root); search until N(root) >= current_readouts + readouts
        if (active[g] && w.st.phase == PH_MATCH_WAIT && !w.st.err) {
          w.st.target_N = simt::fadd(w.st.root_N, (float)c.readouts);
          w.st.phase = PH_MATCH_SEARCH;
          if (lane == 0) simt::atomic_add(&v.ctr[CTR_MATCH_BUSY], 1ULL);
        }
        break;
      case MK_PICK: {  // should_resign (mcts_play.jl:124), else pick_move (mcts_play.jl:52-71)
        int mv = -1, resign = 0;
        float sc = 0.f;
        if (active[g] && w.st.phase == PH_MATCH_WAIT && !w.st.err) {
          const NodeMeta rm = w.load_meta(w.st.root);
          const float q = simt::fdiv(w.st.root_W, simt::fadd(1.0f, w.st.root_N));
          const float qp = simt::fmul(q, (float)rm.to_play);
          if ((double)qp < w.st.resign_thr) {
            resign = 1;
            const uint32_t* lb = w.bits_of(w.st.root);
            sc = game_score(c, w.B, bits_load(w.B, lb, lb + c.KB));
          } else {
            mv = w.pick_move();
          }
        }
        if (lane == 0) { out_i[g] = mv; out_j[g] = resign; out_f[g] = sc; }
        break;
      }
      case MK_PLAY: {  // play_move!(player, move) (mcts_play.jl:26-50), then is_done (mcts.jl:230-231) / score
        int done = 0;
        float sc = 0.f;
        const int mv = moves_in[g];
        if (mv >= 0 && w.st.phase == PH_MATCH_WAIT && !w.st.err) {
          const int prc = w.play_move(mv, true);
          if (prc == E_ILLEGAL) {  // IllegalMove is caught by play_move!(player, c): the tree is unchanged (mcts_play.jl:41-47)
            done = -1;
            w.st.err = 0;
          } else if (prc == E_OK) {
            const NodeMeta nm = w.load_meta(w.st.root);
            if (w.terminal(nm)) {
              done = 1;
              const uint32_t* lb = w.bits_of(w.st.root);
              sc = game_score(c, w.B, bits_load(w.B, lb, lb + c.KB));
            }
          }
        }
        if (lane == 0) { out_i[g] = done; out_f[g] = sc; }
        break;
      }
      default: break;
    }
    w.store_state();
  }
};

// With the DummyNet evaluator (fixed priors and value: test/test_mcts_player.jl:10-32, BASELINE config C5) nothing runs between
// the two halves of tree_search!, and games never interact: one warp plays `rounds` whole rounds of its game in one launch
// (select -> incorporate -> move logic), so there is no per-round launch, no per-round tail and no state round trip.
template <int KA, int OCC = 0>
struct DummyRoundsOp {
  Cfg c;
  View v;
  int rounds;
  AGZ_DEV void operator()(int g, char* smem) const {
    Warp<KA> w(c, v, g, smem);
    for (int r = 0; r < rounds; ++r) {
      if (w.st.phase == PH_SEED) w.search_select(1, true);
      else if (w.st.phase == PH_SEARCH || w.st.phase == PH_MATCH_SEARCH) w.search_select(c.parallel, false);
      else { w.st.nleaf = 0; w.st.seed_round = 0; }
      simt::sync();  // the leaf records written by lane 0 are read by every lane below
      w.search_incorporate();
      if (w.after_round_due()) {
        w.store_hot();
        simt::sync();
        after_round_cold<KA>(&c, &v, g, smem);
        w.st = v.gs[g];
      }
      w.st.seed_round = 0;
      if (w.st.phase == PH_IDLE) break;
    }
    w.store_hot();
  }
};

// get_feats for every leaf collected this round, reference layout: out[b][17][N2] float (N x N x 17 x B)
template <int KA>
struct LeafFeaturesF32Op {
  Cfg c;
  View v;
  float* out;
  AGZ_DEV void operator()(int b, char* smem) const {
    const int g = b / c.pmax, k = b % c.pmax;
    Warp<KA> w(c, v, g, smem);
    float* o = out + (size_t)b * 17 * c.N2;
    const int lane = simt::lane();
    if (k >= w.st.nleaf) {
      for (int i = lane; i < 17 * c.N2; i += 32) o[i] = 0.f;
      return;
    }
    const int node = v.leaf_node[b];
    const int N2 = c.N2;
    int tp = w.gather_features(node, [&](int hb, int p, uint32_t mine, uint32_t theirs) {
      o[(2 * hb) * N2 + p] = (float)mine;
      o[(2 * hb + 1) * N2 + p] = (float)theirs;
    });
    for (int p = lane; p < N2; p += 32) o[16 * N2 + p] = (float)tp;
  }
};

}  // namespace agz
#if AGZ_CUDA
namespace devrt {
template <class Op> struct MinBlocks;
// 7 CTAs = 28 warps per SM (72 registers): 8192 trees fill 148 SMs in 1.98 waves; measured on C5 against 8 / 6 / 5 CTAs per SM:
// 0.207 ms vs 0.218 / 0.237 / 0.219 ms per round (round 1); again in round 2 with the leaner kernel, same box: 7 -> 0.1805,
// 8 (64 registers) -> 0.1822, 9 (56 registers) -> 0.1967 ms per round
template <> struct MinBlocks<agz::SelectOp<3, 1>> { static const int v = 7; };
template <> struct MinBlocks<agz::SelectOp<6, 1>> { static const int v = 6; };
template <> struct MinBlocks<agz::IncorporateOp<3>> { static const int v = 7; };
template <class Op> struct TraceTag;
template <int KA, int OCC> struct TraceTag<agz::SelectOp<KA, OCC>> { static const int v = 1; };
template <int KA> struct TraceTag<agz::IncorporateOp<KA>> { static const int v = 2; };
template <> struct MinBlocks<agz::DummyRoundsOp<3, 1>> { static const int v = 7; };
template <> struct MinBlocks<agz::DummyRoundsOp<6, 1>> { static const int v = 6; };
}  // namespace devrt
#endif
namespace agz {

// ------------------------------------------------------------------------------------------------ hooks
enum {
  HK_INIT = 1, HK_SELECT, HK_INCORPORATE, HK_BACKUP, HK_VLOSS_ADD, HK_VLOSS_REVERT, HK_ADD_CHILD, HK_NOISE, HK_PICK,
  HK_PLAY, HK_FEATURES, HK_RESIGN, HK_POS_PLAY, HK_POS_LEGAL, HK_POS_SCORE, HK_POS_LIBS
};

struct HookParams {
  int kind, slot, node, fmove;
  float value;
  double thr;
  long long game_id;
  const float* probs;
  const agz_position* pos_in;
  agz_position* pos_out;
  int8_t* legal_out;
  uint8_t* libs_out;
  float* f_out;
  int* result;  // [0] status, [1] int result, [2] float bits
};

template <int KA>
struct HookOp {
  Cfg c;
  View v;
  HookParams h;

  AGZ_DEV void finish(Warp<KA>& w, int status, int ival, float fval) const {
    simt::sync();
    if (simt::lane() == 0) {
      h.result[0] = status;
      h.result[1] = ival;
      h.result[2] = 0;
      reinterpret_cast<float*>(h.result)[3] = fval;
    }
    (void)w;
  }

  AGZ_DEV void operator()(int, char* smem) const {
    Warp<KA> w(c, v, h.slot, smem);
    const int lane = simt::lane();
    w.st.err = 0;
    switch (h.kind) {
      case HK_INIT: {  // initialize_game!(player, pos) (mcts_play.jl:110-118)
        const agz_position* p = h.pos_in;
        w.st.game_id = h.game_id;
        w.st.game_id_lo = (uint32_t)h.game_id;
        w.st.resign_thr = c.resign_threshold;
        w.st.phase = PH_MANUAL;
        w.st.target_N = 0.f;
        if (p == nullptr) {
          w.st.hist_len = 0;
          w.pos.b = 0;
          w.pos.w = 0;
          w.init_root_from_scratch(0, -1, 1, 0);
        } else {
          uint32_t* hh = v.hist + (size_t)h.slot * (7 * 2 * c.KB);
          int nh = p->n_hist < 0 ? 0 : (p->n_hist > 7 ? 7 : p->n_hist);
          for (int r = 0; r < nh; ++r) {
            for (int k = 0; k < c.KB; ++k) {
              int pt = k * 32 + lane;
              int x = pt < c.N2 ? p->hist[r][pt] : 0;
              unsigned b = simt::ballot(x == 1), wv = simt::ballot(x == -1);
              if (lane == 0) {
                hh[(size_t)r * 2 * c.KB + k] = b;
                hh[(size_t)r * 2 * c.KB + c.KB + k] = wv;
              }
            }
          }
          w.st.hist_len = nh;
          w.pos = bits_from_bytes(w.B, p->board);
          int flags = (p->last_move_pass ? F_LASTPASS : 0) | (p->done ? F_DONE : 0);
          w.init_root_from_scratch(p->n, p->ko, p->to_play, flags);
        }
        finish(w, 0, 0, 0.f);
        break;
      }
      case HK_SELECT: {
        PathEnt* path = w.path_of(0);
        int plen = 0;
        int leaf = w.select_leaf(h.node < 0 ? w.st.root : h.node, path, plen);
        finish(w, w.st.err, leaf, 0.f);
        break;
      }
      case HK_INCORPORATE:
      case HK_BACKUP:
      case HK_VLOSS_ADD:
      case HK_VLOSS_REVERT: {
        PathEnt* path = w.path_of(0);
        int plen = w.build_path(h.node, path);
        if (!w.st.err) {
          if (h.kind == HK_INCORPORATE) w.incorporate(h.node, path, plen, h.probs, h.value);
          else if (h.kind == HK_BACKUP) w.apply_path(path, plen, OP_BACKUP, h.value);
          else w.apply_path(path, plen, h.kind == HK_VLOSS_ADD ? OP_VLOSS_ADD : OP_VLOSS_REVERT, 0.f);
        }
        finish(w, w.st.err, 0, 0.f);
        break;
      }
      case HK_ADD_CHILD: {  // maybe_add_child! (mcts.jl:140-147)
        const size_t r = w.row(h.node);
        int child = v.child[r + h.fmove];
        if (child < 0) {
          NodeMeta m = w.load_meta(h.node);
          child = w.create_child(h.node, m, h.fmove, true);
          if (child >= 0 && lane == 0) v.child[r + h.fmove] = child;
        }
        finish(w, w.st.err, child, 0.f);
        break;
      }
      case HK_NOISE:
        w.inject_noise();
        finish(w, 0, 0, 0.f);
        break;
      case HK_PICK: {
        int mv = w.pick_move();
        finish(w, w.st.err, mv, 0.f);
        break;
      }
      case HK_PLAY: {
        int rc = w.play_move(h.fmove, true);
        finish(w, rc, w.st.root, 0.f);
        break;
      }
      case HK_RESIGN: {  // should_resign (mcts_play.jl:124)
        NodeMeta rm = w.load_meta(w.st.root);
        float q = simt::fdiv(w.st.root_W, simt::fadd(1.0f, w.st.root_N));
        float qp = simt::fmul(q, (float)rm.to_play);
        finish(w, 0, (double)qp < h.thr ? 1 : 0, qp);
        break;
      }
      case HK_FEATURES: {
        float* o = h.f_out;
        const int N2 = c.N2;
        int tp = w.gather_features(h.node, [&](int hb, int p, uint32_t mine, uint32_t theirs) {
          o[(2 * hb) * N2 + p] = (float)mine;
          o[(2 * hb + 1) * N2 + p] = (float)theirs;
        });
        for (int p = lane; p < N2; p += 32) o[16 * N2 + p] = (float)tp;
        finish(w, 0, 0, 0.f);
        return;  // no state change
      }
      case HK_POS_PLAY: {  // play_move!(pos, c) (board.jl:451-509) on a caller-supplied position
        const agz_position* p = h.pos_in;
        agz_position* o = h.pos_out;
        w.pos = bits_from_bytes(w.B, p->board);
        int status = 0, ko = -1, ncap = 0;
        const int mv = h.fmove;
        bool ended = false;
        if (mv != c.pass) {
          if (mv == p->ko) status = E_ILLEGAL;
          else if (game_play(c, w.B, w.pos, mv, p->to_play, true, ko, ncap, ended)) status = E_ILLEGAL;
        }
        if (status == 0) {
          bits_to_bytes(w.B, w.pos, o->board);
          for (int k = 0; k < c.KB; ++k) {
            int pt = k * 32 + lane;
            if (pt < c.N2) {
              o->hist[0][pt] = p->board[pt];
              for (int r = 1; r < 7; ++r) o->hist[r][pt] = p->hist[r - 1][pt];
            }
          }
          if (lane == 0) {
            o->n_hist = p->n_hist < 7 ? p->n_hist + 1 : 7;
            o->n = p->n + 1;
            o->to_play = -p->to_play;
            o->ko = mv == c.pass ? -1 : ko;
            o->last_move_pass = mv == c.pass;
            o->done = ((mv == c.pass && p->last_move_pass) || ended) ? 1 : 0;
            o->caps[0] = p->caps[0] + (p->to_play == 1 ? ncap : 0);
            o->caps[1] = p->caps[1] + (p->to_play == 1 ? 0 : ncap);
            o->komi = p->komi;
          }
        }
        finish(w, status, ncap, 0.f);
        return;
      }
      case HK_POS_LEGAL: {  // all_legal_moves (board.jl:393-424)
        const agz_position* p = h.pos_in;
        w.pos = bits_from_bytes(w.B, p->board);
        const uint32_t legal = game_legal(c, w.B, w.pos, p->to_play, p->ko);
        if (lane < c.N)
          for (int i = 0; i < c.N; ++i) h.legal_out[c.N * lane + i] = (int8_t)((legal >> i) & 1u);
        if (lane == 0 && c.pass >= 0) h.legal_out[c.pass] = 1;
        finish(w, 0, 0, 0.f);
        return;
      }
      case HK_POS_SCORE: {
        const agz_position* p = h.pos_in;
        Cfg cc = c;
        cc.komi = p->komi;
        float sc = game_score(cc, w.B, bits_from_bytes(w.B, p->board));
        finish(w, 0, 0, sc);
        return;
      }
      case HK_POS_LIBS: {  // liberty_cache: the shared-memory label / liberty-count code of go_rules.cuh
        const agz_position* p = h.pos_in;
        Board OB;
        OB.N = c.N; OB.N2 = c.N2; OB.KB = c.KB;
        board_init_masks(OB);
        RulesScratch rs = rules_scratch_at(w.smem, c.KB);
        rules_load_bytes(OB, rs, p->board);
        rules_label(OB, rs, 0);
        rules_count_liberties(OB, rs);
        for (int k = 0; k < c.KB; ++k) {
          int pt = k * 32 + lane;
          if (pt < c.N2) h.libs_out[pt] = (uint8_t)(rs.bd[pt] != 0 ? rs.cnt[rs.lab[pt]] : 0);
        }
        finish(w, 0, 0, 0.f);
        return;
      }
      default:
        finish(w, E_ASSERT, 0, 0.f);
        return;
    }
    w.store_state();
  }
};

}  // namespace agz
