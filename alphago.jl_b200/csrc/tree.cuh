// tree.cuh -- the MCTS search tree in HBM and the warp-per-tree device functions that walk it.
//
// Replaces (reference paths): src/mcts.jl  MCTSNode :41-82, select_leaf :108-138, maybe_add_child! :140-147,
// add/revert_virtual_loss! :149-171, revert_visits! :173-186, incorporate_results! :188-213,
// backup_value! :215-225, is_done :230-231, inject_noise! :233-239, children_as_pi :241-252;
// src/mcts_play.jl  tree_search! :73-98, pick_move :52-71, play_move! :26-50, should_resign :124;
// src/selfplay.jl :1-45 (per-game state machine).
//
// Layout (struct of arrays, one arena of `cap` nodes per game slot, node id local to the slot):
//   N, W, P      float [slot][node][AS]   AS = 32*KA >= A, rows are 128-byte aligned -> coalesced warp loads
//   child        int32 [slot][node][AS]   child node id or -1
//   meta         NodeMeta [slot][node]    16 bytes
//   bits         uint32 [slot][node][3*KB] black plane, white plane, legal-move mask
// A node's own N / W live in its parent's row (as in the reference, mcts.jl:94-102); the root's live in the
// per-game state.  One warp owns one game: the 8 selections of a tree_search! round are serial inside the
// warp (they are order dependent through virtual loss), parallelism comes from thousands of games.
#pragma once
#include "go_bits.cuh"
#include "go_rules.cuh"
#include "rng.cuh"
#include "simt.h"

namespace agz {

enum { PH_IDLE = 0, PH_SEED = 1, PH_SEARCH = 2, PH_WAIT_RING = 3, PH_MANUAL = 4,
       PH_MATCH_WAIT = 5,      // two-player match slot (evaluate / play): waiting for the host to arm a search or play a move
       PH_MATCH_SEARCH = 6,    // ... searching until N(root) >= target_N
       PH_DELAY = 7 };         // staggered start (option selfplay.stagger_rounds): the slot's first game begins after `delay` rounds
enum { F_EXPANDED = 1, F_DONE = 2, F_LASTPASS = 4 };
enum { E_OK = 0, E_ILLEGAL = 1, E_ASSERT = 2, E_CAPACITY = 6 };
enum { OP_VLOSS_ADD = 0, OP_VLOSS_REVERT = 1, OP_BACKUP = 2, OP_REVERT_VISITS = 3 };
enum { GAME_GO = 0, GAME_GOMOKU = 1 };  // the two games behind the reference's Position interface (src/game/env.jl)
static const unsigned long long SLOT_ROOT = ~0ULL;

struct alignas(16) NodeMeta {  // 16 bytes, moved as one 128-bit word
  int32_t parent;
  int16_t fmove;  // move that led here (-1 for an arena root)
  int16_t n;      // position.n
  int16_t ko;     // flat point or -1
  int8_t to_play;
  uint8_t flags;
  int32_t pad;
};

struct alignas(16) PathEnt {  // 16 bytes; where this path node's own N / W live
  unsigned long long slot;  // index into N / W (64-bit: arenas of all games can exceed 2^32 entries), or SLOT_ROOT
  int32_t node;
  int32_t to_play;
};

struct alignas(16) GameState {
  // written by every search round: the first 48 bytes (Warp::store_hot), so that the fields below do not have to stay live in
  // registers through the search kernels
  float root_N, root_W;
  int32_t count, err;
  uint32_t sel_ctr;
  int32_t nleaf, seed_round;
  int32_t vloss_balance;  // path entries with a virtual loss outstanding
  int32_t tiny_values;    // an evaluator value with 0 < |v| < 2^-70 was incorporated: the PUCT quotients take the IEEE divisions
  int32_t hot_pad[3];
  // read by the search, written by the per-move logic
  int32_t phase, root;
  float target_N;
  uint32_t game_id_lo;
  // per-move logic only
  int32_t n_moves;       // plies recorded for the current game
  uint32_t noise_ctr;
  int32_t hist_len;
  int32_t result, resigned;
  float final_score;
  int32_t delay;          // PH_DELAY: rounds left before the first game of this slot starts
  int32_t pad;
  double resign_thr;
  int64_t game_id;
};

struct Cfg {
  int N, N2, A, KA, AS, KB;
  int game, n_in_row;  // GAME_GO (A = N^2 + 1) | GAME_GOMOKU (A = N^2: no pass; n_in_row stones win, gomoku.jl:1-19)
  int pass;            // the flat pass move N^2, or -1 when the game has none
  int cap, maxd, pmax;
  int max_game_length, tau_threshold, readouts, parallel;
  int n_games, world, rank, inject_noise;
  int ring_cap;
  float komi;
  double c_puct, noise_weight, noise_alpha, resign_threshold, resign_disable_frac;
  uint64_t seed;
  long long total_games;
  long long stagger_rounds;  // 0 = every slot starts at once; R = slot g starts after g*R/n_games rounds (throughput runs: slots spread over all plies)
};

enum { CTR_MOVES = 0, CTR_FINISHED, CTR_STARTED, CTR_POSITIONS, CTR_READOUTS, CTR_PATHNODES, CTR_RING_TAIL, CTR_RING_HEAD,
       CTR_MATCH_BUSY,  // match slots still searching
       CTR_PRUNES,      // arena-pressure prunes (forget_leaves)
       CTR_DUP_LEAVES,  // leaves of a round that were duplicates of an earlier leaf of the same round (revert_visits!)
       CTR_COUNT };

struct RingHeader {  // mirrors agz_game_header
  int64_t game_id;
  int32_t n_moves, result, resigned;
  float final_score;
  double resign_threshold;
};

struct View {
  float *N, *W, *P;
  int32_t* child;
  NodeMeta* meta;
  uint32_t* bits;
  GameState* gs;
  uint32_t* hist;  // [slot][7][2*KB]
  PathEnt* path;   // [slot][pmax][maxd]
  int32_t *leaf_node, *leaf_plen;  // [slot][pmax]
  int32_t* remap;  // [slot][cap] old id -> new id (compaction)
  int32_t* order;  // [slot][cap] new id -> old id
  const float* eval_pi;  // row b at eval_pi + b*pi_stride
  const float* eval_v;   // eval_v[b*v_stride]
  long long pi_stride;
  int v_stride;
  int16_t* rec_moves;  // [slot][L]
  float *rec_q, *rec_pi, *rec_vis;  // [slot][L], [slot][L][A], [slot][L][A]
  RingHeader* ring_hdr;
  int16_t* ring_moves;
  float *ring_q, *ring_pi, *ring_vis;
  unsigned long long* ctr;
  unsigned long long* trace;  // kernel timeline trace buffer (simt.h) or nullptr
  double* noise_g;            // [slot][AS] scratch of inject_noise!: the gamma variates of the Dirichlet draw
  const double* rcp;          // rcp[k] = RN(1/k), k in [1, rcp_n): the divisor 1 + N(child) of the PUCT score is a small integer
  int rcp_n;
};

// ---- the Position interface (src/game/env.jl): play_move!, all_legal_moves, score dispatched on the game ------------------------
// score: Go = Tromp-Taylor area score (board.jl:511-533); Gomoku = the winner's colour (gomoku board.jl:171), 0 for a full board
AGZ_DEV float game_score(const Cfg& c, const BitsCtx& B, const Lines& L) {
  if (c.game == GAME_GOMOKU) return bits_k_in_row(B, L.b, c.n_in_row) ? 1.f : (bits_k_in_row(B, L.w, c.n_in_row) ? -1.f : 0.f);
  return bits_score(B, L, c.komi);
}
// the legal points of this lane's line (a pass, where the game has one, is always legal and not part of the set)
AGZ_DEV uint32_t game_legal(const Cfg& c, const BitsCtx& B, const Lines& L, int to_play, int ko) {
  if (c.game == GAME_GOMOKU) return B.full & ~(L.b | L.w);  // gomoku board.jl:95
  return bits_legal(B, L, to_play, ko);
}
// Play `color` at flat point mv (not a pass).  Returns 0, or 1 when check_legal is set and the move is illegal (L unchanged).
// done_out: the move ended the game (Gomoku: n_in_row made, or the board is full; Go ends by two passes, never here).
AGZ_DEV int game_play(const Cfg& c, const BitsCtx& B, Lines& L, int mv, int color, bool check_legal, int& ko_out, int& ncap_out, bool& done_out) {
  done_out = false;
  if (c.game == GAME_GOMOKU) {  // gomoku board.jl:137-169
    ko_out = -1;
    ncap_out = 0;
    const int cj = mv / B.N, ci = mv - cj * B.N;
    const uint32_t cbit = B.lane == cj ? (1u << ci) : 0u;
    if (check_legal && simt::any(((L.b | L.w) & cbit) != 0u)) return 1;
    if (color == 1) L.b |= cbit; else L.w |= cbit;
    const bool five = bits_k_in_row(B, color == 1 ? L.b : L.w, c.n_in_row);
    done_out = five || !simt::any((B.full & ~(L.b | L.w)) != 0u);
    return 0;
  }
  return bits_play(B, L, mv, color, check_legal, ko_out, ncap_out);
}

// position of the n-th (0-based) set bit of m (n < popc(m)): five popcount halvings instead of clearing n bits one by one
AGZ_DEV int nth_set_bit(unsigned m, int n) {
  int pos = 0;
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const int cnt = simt::popc((m >> pos) & ((1u << w) - 1u));
    if (n >= cnt) { n -= cnt; pos += w; }
  }
  return pos;
}

template <int KA>
struct Warp {
  const Cfg& c;
  const View& v;
  const int g;
  const int lane;
  BitsCtx B;
  Lines pos;     // the position being worked on: this lane's board line of black / white stones (go_bits.cuh)
  char* smem;    // per-warp scratch of the shared-memory rules code (go_rules.cuh; liberty-cache hook only)
  GameState st;  // register copy (warp-uniform); written back by store_state()
  unsigned nbase;   // first node of this game's arena: all arenas together hold < 2^31 nodes (a node is > 400 bytes), so node indices are
                    // 32-bit and only the final products are widened (one IMAD.WIDE instead of 64-bit adds and multiplies)
  bool prefetch;  // latency mode (few trees per SM): pull the most-visited child's rows into L2 while this level is scored

  AGZ_DEV Warp(const Cfg& c_, const View& v_, int g_, char* smem_) : c(c_), v(v_), g(g_), lane(simt::lane()), smem(smem_) {
    B = bits_ctx(c.N, c.KB);
    pos.b = 0;
    pos.w = 0;
    st = v.gs[g];
    nbase = (unsigned)g * (unsigned)c.cap;
    prefetch = false;
  }
  AGZ_DEV void store_state() {
    simt::sync();
    if (lane == 0) v.gs[g] = st;
  }
  // what a search round changes (the leading fields of GameState)
  AGZ_DEV void store_hot() {
    simt::sync();
    if (lane == 0) {
      GameState* d = v.gs + g;
      d->root_N = st.root_N; d->root_W = st.root_W; d->count = st.count; d->err = st.err;
      d->sel_ctr = st.sel_ctr; d->nleaf = st.nleaf; d->seed_round = st.seed_round; d->vloss_balance = st.vloss_balance;
      d->tiny_values = st.tiny_values;
    }
  }
  // AS = 32*KA and the bit planes' node stride 3*KA are compile-time constants here (the planes themselves sit KB words apart)
  AGZ_DEV size_t row(int node) const { return (size_t)(nbase + (unsigned)node) * (unsigned)(KA * 32); }
  AGZ_DEV uint32_t* bits_of(int node) const { return v.bits + (size_t)(nbase + (unsigned)node) * (unsigned)(3 * KA); }
  // element `idx` of a per-lane register array, as a select chain (a dynamically indexed array would be spilled to local memory)
  template <class T>
  AGZ_DEV static T pick(const T (&a)[KA], int idx) {
    T x = a[0];
#pragma unroll
    for (int k = 1; k < KA; ++k) x = idx == k ? a[k] : x;
    return x;
  }
  AGZ_DEV NodeMeta load_meta(int node) const {
#if AGZ_CUDA
    NodeMeta m;   // one 16-byte load instead of a load per field
    *reinterpret_cast<uint4*>(&m) = *reinterpret_cast<const uint4*>(v.meta + (nbase + (unsigned)node));
    return m;
#else
    return v.meta[nbase + node];
#endif
  }
  AGZ_DEV bool terminal(const NodeMeta& m) const { return (m.flags & F_DONE) || m.n >= c.max_game_length; }
  AGZ_DEV PathEnt* path_of(int k) const { return v.path + (size_t)((unsigned)g * (unsigned)c.pmax + (unsigned)k) * (unsigned)c.maxd; }
  AGZ_DEV void count(int which, unsigned long long n) {
    if (lane == 0) simt::atomic_add(&v.ctr[which], n);
  }

  // ---- node creation -----------------------------------------------------------------------
  // A new node starts with child_N = 0 and no children.  Its child_W / child_prior rows are NOT written here: nothing reads them
  // before incorporate_results! fills them (only expanded nodes are scored, mcts.jl:116-118), which saves two of the six row
  // writes of an expansion; agz_tree_read_node reports them as zeros for an unexpanded node, like the reference's fresh arrays.
  AGZ_DEV void init_rows(int node) {
    size_t r = row(node);
#pragma unroll
    for (int k = 0; k < KA; ++k) {
      int a = k * 32 + lane;
      v.N[r + a] = 0.f;
      v.child[r + a] = -1;
    }
  }

  // Write a node whose position is `pos` (terminal nodes carry no legal-move mask: `skip_legal`).
  AGZ_DEV NodeMeta write_node(int node, int parent, int fmove, int n, int ko, int to_play, int flags, bool skip_legal) {
    uint32_t bw[KA], ww[KA], lw[KA];
    uint32_t legal = 0;
    if (!skip_legal) legal = game_legal(c, B, pos, to_play, ko);
    bits_pack3<KA>(B, pos.b, pos.w, legal, bw, ww, lw);
    uint32_t* bp = bits_of(node);
#pragma unroll
    for (int k = 0; k < KA; ++k) {
      if (k < c.KB && lane == k) {
        bp[k] = bw[k];
        bp[c.KB + k] = ww[k];
        bp[2 * c.KB + k] = lw[k];
      }
    }
    NodeMeta m;
    m.parent = parent; m.fmove = (int16_t)fmove; m.n = (int16_t)n; m.ko = (int16_t)ko;
    m.to_play = (int8_t)to_play; m.flags = (uint8_t)flags; m.pad = 0;
    if (lane == 0) v.meta[nbase + node] = m;
    init_rows(node);
    return m;
  }

  // maybe_add_child! when the child is missing: play `move` from `parent` (mcts.jl:140-147 -> board.jl:451-509).
  // Returns the new node id, or -1 with st.err set (E_CAPACITY, or E_ILLEGAL when check_legal).
  AGZ_DEV int create_child(int parent, const NodeMeta& pm, int move, bool check_legal, NodeMeta* child_meta = nullptr) {
    if (st.count >= c.cap) { st.err = E_CAPACITY; return -1; }
    const uint32_t* pb = bits_of(parent);
    if (check_legal && move != c.pass) {
      uint32_t lw = pb[2 * c.KB + (move >> 5)];
      if (!((lw >> (move & 31)) & 1u)) { st.err = E_ILLEGAL; return -1; }
    }
    int idx = st.count++;
    pos = bits_load(B, pb, pb + c.KB);
    int color = pm.to_play;
    int n = pm.n + 1;
    int ko = -1, flags = 0, ncap = 0;
    bool term;
    if (move == c.pass) {  // pass_move! (board.jl:426-440)
      flags = F_LASTPASS | ((pm.flags & F_LASTPASS) ? F_DONE : 0);
    } else {
      bool ended;
      game_play(c, B, pos, move, color, false, ko, ncap, ended);
      if (ended) flags = F_DONE;
    }
    term = (flags & F_DONE) || n >= c.max_game_length;
    const NodeMeta cm = write_node(idx, parent, move, n, ko, -color, flags, term);
    if (child_meta) *child_meta = cm;
    simt::sync();
    return idx;
  }

  // ---- path updates (virtual loss, backup, visit reverts) -------------------------------------
  // Entries are distinct nodes, so the read-modify-writes of one path go out in parallel, one per lane.
  AGZ_DEV void apply_path(const PathEnt* path, int plen, int op, float value) {
    simt::sync();
    for (int d0 = 0; d0 < plen; d0 += 32) {
      int d = d0 + lane;
      if (d < plen) {
        PathEnt e = path[d];
        if (e.slot != SLOT_ROOT) {
          if (op == OP_REVERT_VISITS) v.N[e.slot] = simt::fsub(v.N[e.slot], 1.0f);
          else {
            float add = op == OP_BACKUP ? value : (op == OP_VLOSS_ADD ? (float)e.to_play : (float)(-e.to_play));
            v.W[e.slot] = simt::fadd(v.W[e.slot], add);
          }
        }
      }
    }
    PathEnt e0 = path[0];
    if (e0.slot == SLOT_ROOT) {
      if (op == OP_REVERT_VISITS) st.root_N = simt::fsub(st.root_N, 1.0f);
      else {
        float add = op == OP_BACKUP ? value : (op == OP_VLOSS_ADD ? (float)e0.to_play : (float)(-e0.to_play));
        st.root_W = simt::fadd(st.root_W, add);
      }
    }
    if (op == OP_VLOSS_ADD) st.vloss_balance += plen;
    if (op == OP_VLOSS_REVERT) st.vloss_balance -= plen;
    simt::sync();
  }

  // apply_path for the path the last select_leaf left in the lanes' registers (plen <= 32)
  AGZ_DEV void apply_path_regs(int plen, int op, float value) {
    if (lane < plen) {
      const float add = op == OP_BACKUP ? value : (op == OP_VLOSS_ADD ? (float)reg_tp : (float)(-reg_tp));
      if (reg_slot != SLOT_ROOT) v.W[reg_slot] = simt::fadd(v.W[reg_slot], add);
    }
    const unsigned long long s0 = simt::shfl((unsigned)(reg_slot == SLOT_ROOT ? 1u : 0u), 0);
    if (s0) {
      const int tp0 = simt::shfl(reg_tp, 0);
      const float add = op == OP_BACKUP ? value : (op == OP_VLOSS_ADD ? (float)tp0 : (float)(-tp0));
      st.root_W = simt::fadd(st.root_W, add);
    }
    if (op == OP_VLOSS_ADD) st.vloss_balance += plen;
    if (op == OP_VLOSS_REVERT) st.vloss_balance -= plen;
    simt::sync();
  }

  // Rebuild the path root..node by walking parent pointers (used by the single-node hooks only).
  AGZ_DEV int build_path(int node, PathEnt* path) {
    int len = 0;
    for (int x = node; x >= 0; x = load_meta(x).parent) ++len;
    if (len > c.maxd) { st.err = E_ASSERT; return 0; }
    int d = len - 1;
    for (int x = node; x >= 0;) {
      NodeMeta m = load_meta(x);
      if (lane == 0) {
        PathEnt e;
        e.slot = m.parent < 0 ? SLOT_ROOT : (unsigned long long)(row(m.parent) + m.fmove);
        e.node = x; e.to_play = m.to_play;
        path[d] = e;
      }
      --d;
      x = m.parent;
    }
    simt::sync();
    return len;
  }

  // ---- select_leaf (mcts.jl:108-138) ------------------------------------------------------------
  // Lane d of the warp also keeps path entry d (d < 32) in its registers (reg_slot / reg_tp), so that the virtual loss or the
  // terminal backup that follows the descent does not have to read the path back from memory.
  unsigned long long reg_slot;
  int reg_tp;

  AGZ_DEV int select_leaf(int from, PathEnt* path, int& plen, NodeMeta* leaf_meta = nullptr) {
    uint32_t sel_idx = st.sel_ctr++;
    int move_no = -1;   // position.n of the root (RNG key of the tie-break draw): known for free when the descent starts at the root
    int cur = from;
    int depth = 0;
    unsigned long long slot;
    float cur_N;
    {
      NodeMeta fm = load_meta(cur);
      if (fm.parent < 0) {
        slot = SLOT_ROOT;
        st.root_N = simt::fadd(st.root_N, 1.0f);
        cur_N = st.root_N;
      } else {
        slot = (unsigned long long)(row(fm.parent) + fm.fmove);
        cur_N = simt::fadd(v.N[slot], 1.0f);
        simt::sync();
        if (lane == 0) v.N[slot] = cur_N;
      }
    }
    for (;;) {
      // Everything this level needs depends only on `cur`: the node's meta word, its four statistic rows and its legal-move
      // words are requested together, before the expanded flag is known (an unexpanded leaf's rows exist and are zero), so a
      // level costs one dependent memory round trip instead of two -- the descent is a latency chain, one warp per tree.
      const size_t r = row(cur);
      NodeMeta m = load_meta(cur);
      if (depth == 0 && cur == st.root) move_no = m.n;
      if (leaf_meta) *leaf_meta = m;
      float n[KA], w[KA], p[KA];
      int ch[KA];
      uint32_t lwv[KA];
      const uint32_t* lw = bits_of(cur) + 2 * c.KB;
#pragma unroll
      for (int k = 0; k < KA; ++k) {
        int a = k * 32 + lane;
        n[k] = v.N[r + a];
        w[k] = v.W[r + a];
        p[k] = v.P[r + a];
        ch[k] = v.child[r + a];
        lwv[k] = lw[k];  // warp-uniform address; words past KB lie in the node's never-written (zero) tail: the stride is 3*KA words
      }
      if (lane == 0) {
        PathEnt e;
        e.slot = slot; e.node = cur; e.to_play = m.to_play;
        path[depth] = e;
      }
      if (lane == depth) { reg_slot = slot; reg_tp = m.to_play; }
      if (!(m.flags & F_EXPANDED)) break;
      if (depth + 1 >= c.maxd) { st.err = E_ASSERT; break; }
      if (prefetch) {   // the most-visited child is where a sharp policy goes next: its rows travel while this level is scored
        float nm = -1.f;
        int cm = -1;
#pragma unroll
        for (int k = 0; k < KA; ++k)
          if (ch[k] >= 0 && n[k] > nm) { nm = n[k]; cm = ch[k]; }
        const unsigned key = cm >= 0 ? simt::fbits(nm) + 1u : 0u;
        const unsigned best_key = simt::reduce_max(key);
        if (best_key) {
          const unsigned who = simt::ballot(key == best_key);
          const int pc = simt::shfl(cm, simt::ffs(who) - 1);
          const size_t pr = row(pc);
          const int rl = (KA * 32 * 4 + 127) / 128;   // 128-byte lines per statistics row
          const char* q = nullptr;
          if (lane < rl) q = (const char*)(v.N + pr) + 128 * lane;
          else if (lane < 2 * rl) q = (const char*)(v.W + pr) + 128 * (lane - rl);
          else if (lane < 3 * rl) q = (const char*)(v.P + pr) + 128 * (lane - 2 * rl);
          else if (lane < 4 * rl) q = (const char*)(v.child + pr) + 128 * (lane - 3 * rl);
          else if (lane == 4 * rl) q = (const char*)(v.meta + nbase + pc);
          else if (lane == 4 * rl + 1) q = (const char*)bits_of(pc);
          if (q) simt::prefetch_l2(q);
        }
      }
      const int pass = c.pass;   // -1 when the game has no pass: no action index equals it and F_LASTPASS is never set
      int best;
      // HACK of the reference: after a pass, look at the double pass first (mcts.jl:119-126)
      bool pass_first = false;
      if (m.flags & F_LASTPASS) pass_first = simt::shfl(pick(n, pass >> 5), pass & 31) == 0.f;
      if (pass_first) {
        best = pass;
      } else {
        // score = Float64(Float32(W/(1+N)) * to_play) + ((c_puct * Float64(sqrt_f32(1+N_parent))) * Float64(P)) / Float64(1+N)
        // The divisor d = 1 + N(child) is a small integer, so both correctly-rounded quotients are computed from r = RN(1/d)
        // (table, or one exact division when d is past the table) instead of two IEEE divisions per child:
        //   Float32 w/d = RN32(Float64(w) * r)             -- |w*r - w/d| <= 2^-52 relative, a quotient by d < 2^24 is >= 2^-49
        //                                                      relative away from every fp32 rounding boundary (never on one);
        //   Float64 x/d = fma(fma(-q0, d, x), r, q0), q0 = RN(x*r)   -- Markstein's correction; the remainder is exact (d small) and
        //                                                      a quotient by d < 2^24 is >= 2^-78 relative away from a midpoint.
        const double cu = simt::dmul(c.c_puct, (double)simt::fsqrt(simt::fadd(1.0f, cur_N)));
        const float tp = (float)m.to_play;
        float den[KA];
        double rc[KA];
        // Take the IEEE divisions when a divisor could lie past the table -- N(child) <= N(node) in a search, only the test hooks
        // (PH_MANUAL trees, agz_tree_set_stats) can break that -- or when a Float32 quotient could be subnormal (0 < |w| < 2^-100).
        // The latter is decided per game, not per child: every W is a sum of +-1 (virtual losses, game results) and evaluator
        // values; as long as every non-zero value seen had |v| >= 2^-70 all of them are multiples of 2^-93, and so is every
        // rounded partial sum (exact below 2^-69, a multiple of a larger ulp above), hence no non-zero W lies below 2^-93.
        // search_incorporate raises st.tiny_values on the first smaller value and the game takes the IEEE path from then on.
        const bool odd = st.phase == PH_MANUAL || st.tiny_values != 0 || !(simt::fadd(cur_N, 2.0f) < (float)v.rcp_n);
#pragma unroll
        for (int k = 0; k < KA; ++k) den[k] = simt::fadd(1.0f, n[k]);
        double s[KA];
        double mx = -1.0e300;
        if (!odd) {
#pragma unroll
          for (int k = 0; k < KA; ++k) rc[k] = v.rcp[(int)den[k]];
#pragma unroll
          for (int k = 0; k < KA; ++k) {
            const bool legal = (((lwv[k] >> lane) & 1u) | (unsigned)(k * 32 + lane == pass)) != 0u;   // mask bits past N^2 are 0
            const double dd = (double)den[k];
            const float q = simt::fmul((float)simt::dmul((double)w[k], rc[k]), tp);
            const double x = simt::dmul(cu, (double)p[k]);
            const double q0 = simt::dmul(x, rc[k]);
            const double u = simt::dfma(simt::dfma(-q0, dd, x), rc[k], q0);
            s[k] = legal ? simt::dadd((double)q, u) : -1.0e300;
            mx = s[k] > mx ? s[k] : mx;
          }
        } else {
#pragma unroll
          for (int k = 0; k < KA; ++k) {
            const bool legal = (((lwv[k] >> lane) & 1u) | (unsigned)(k * 32 + lane == pass)) != 0u;
            float q = simt::fmul(simt::fdiv(w[k], den[k]), tp);
            double u = simt::ddiv(simt::dmul(cu, (double)p[k]), (double)den[k]);
            s[k] = legal ? simt::dadd((double)q, u) : -1.0e300;
            mx = s[k] > mx ? s[k] : mx;
          }
        }
        // warp maximum of the doubles through two 32-bit REDUX.MAX on an order-preserving integer key (scores are never -0.0:
        // u >= +0 and (-0) + (+0) = +0) instead of five shuffle rounds
        {
          const long long b = simt::dbits(mx);
          const unsigned long long key = (unsigned long long)b ^ (b < 0 ? ~0ULL : 0x8000000000000000ULL);
          const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
          const unsigned mhi = simt::reduce_max(hi);
          const unsigned mlo = simt::reduce_max(hi == mhi ? lo : 0u);
          const unsigned long long mkey = ((unsigned long long)mhi << 32) | mlo;
          mx = simt::bitsd((long long)(mkey ^ ((mkey >> 63) ? 0x8000000000000000ULL : ~0ULL)));
        }
        unsigned tm[KA];
        int total = 0;
#pragma unroll
        for (int k = 0; k < KA; ++k) {
          tm[k] = simt::ballot(s[k] == mx);
          total += simt::popc(tm[k]);
        }
        int pick = 0;   // mulhi(r, 1) = 0: the draw only matters when several children tie (warp-uniform branch)
        if (total > 1) {
          if (move_no < 0) move_no = load_meta(st.root).n;
          U4 rr = rng_draw(c.seed, st.game_id_lo, SITE_SELECT, (uint32_t)move_no, sel_idx, (uint32_t)depth);
          pick = (int)simt::mulhi(rr.x, (uint32_t)total);
        }
        // the pick-th tied child in action order: every lane ranks its own candidates (ties before it in its word + the words before);
        // exactly one (lane, word) has rank `pick`, found with one warp maximum
        {
          const unsigned lt = (1u << lane) - 1u;
          int before = 0;
          unsigned mine = 0u;   // action + 1 of this lane's candidate with rank `pick`, or 0
#pragma unroll
          for (int k = 0; k < KA; ++k) {
            const bool it = ((tm[k] >> lane) & 1u) != 0u && before + simt::popc(tm[k] & lt) == pick;
            mine = it ? (unsigned)(k * 32 + lane + 1) : mine;
            before += simt::popc(tm[k]);
          }
          const unsigned hit = simt::reduce_max(mine);
          best = hit ? (int)hit - 1 : pass;
        }
      }
      const int ok = best >> 5, ol = best & 31;
      const float n_old = simt::shfl(pick(n, ok), ol);
      int child = simt::shfl(pick(ch, ok), ol);
      const float n_new = simt::fadd(n_old, 1.0f);
      if (child < 0) {
        // A child created here is the leaf of this readout (not expanded, mcts.jl:116-118): everything the next level would load
        // back from memory is known, so the descent ends without that round trip.
        NodeMeta cm;
        child = create_child(cur, m, best, false, &cm);
        if (child < 0) break;
        if (lane == ol) {
          v.child[r + best] = child;
          v.N[r + best] = n_new;  // N(child) += 1 (mcts.jl:113-114)
        }
        ++depth;
        if (lane == 0) {
          PathEnt e;
          e.slot = (unsigned long long)(r + best); e.node = child; e.to_play = cm.to_play;
          path[depth] = e;
        }
        if (lane == depth) { reg_slot = (unsigned long long)(r + best); reg_tp = cm.to_play; }
        if (leaf_meta) *leaf_meta = cm;
        cur = child;
        break;
      }
      if (lane == ol) v.N[r + best] = n_new;  // N(child) += 1 (mcts.jl:113-114)
      slot = (unsigned long long)(r + best);
      cur_N = n_new;
      cur = child;
      ++depth;
    }
    plen = depth + 1;
    simt::sync();
    return cur;
  }

  // ---- tree_search! first half: collect leaves (mcts_play.jl:74-87) ------------------------------
  AGZ_DEV void search_select(int parallel, bool seed_mode) {
    int nleaf = 0, attempts = 0;
    const int want = seed_mode ? 1 : parallel;
    unsigned long long n_readouts = 0, n_pathnodes = 0;   // one atomic per counter per warp, not per readout
    while (nleaf < want && attempts < 2 * want) {
      ++attempts;
      PathEnt* path = path_of(nleaf);
      int plen = 0;
      NodeMeta lm;
      int leaf = select_leaf(st.root, path, plen, &lm);   // lm = the leaf's meta word as the last level of the descent loaded it
      if (st.err) break;
      n_readouts += 1;
      n_pathnodes += (unsigned long long)plen;
      if (terminal(lm)) {  // game over: back up the true result, do not evaluate (mcts_play.jl:80-84)
        const uint32_t* lb = bits_of(leaf);
        float sc = game_score(c, B, bits_load(B, lb, lb + c.KB));
        float value = sc > 0.f ? 1.f : (sc < 0.f ? -1.f : 0.f);
        if (plen <= 32) apply_path_regs(plen, OP_BACKUP, value);
        else apply_path(path, plen, OP_BACKUP, value);
        continue;
      }
      if (!seed_mode) {
        if (plen <= 32) apply_path_regs(plen, OP_VLOSS_ADD, 0.f);
        else apply_path(path, plen, OP_VLOSS_ADD, 0.f);
      }
      if (lane == 0) {
        v.leaf_node[(size_t)g * c.pmax + nleaf] = leaf;
        v.leaf_plen[(size_t)g * c.pmax + nleaf] = plen;
      }
      ++nleaf;
    }
    st.nleaf = nleaf;
    st.seed_round = seed_mode ? 1 : 0;
    if (n_readouts) {
      count(CTR_READOUTS, n_readouts);
      count(CTR_PATHNODES, n_pathnodes);
    }
    if (nleaf) count(CTR_POSITIONS, (unsigned long long)nleaf);
  }

  // incorporate_results! for one node given its path (mcts.jl:188-213)
  AGZ_DEV void incorporate(int leaf, const PathEnt* path, int plen, const float* probs, float value) {
    NodeMeta lm = load_meta(leaf);
    simt::sync();  // every lane has read the flags before lane 0 rewrites them below
    if (lm.flags & F_DONE) { st.err = E_ASSERT; return; }  // @assert !position.done (mcts.jl:196)
    if (lm.flags & F_EXPANDED) {
      apply_path(path, plen, OP_REVERT_VISITS, 0.f);
      return;
    }
    if (lane == 0) v.meta[nbase + leaf].flags = (uint8_t)(lm.flags | F_EXPANDED);
    const size_t r = row(leaf);
#pragma unroll
    for (int k = 0; k < KA; ++k) {
      int a = k * 32 + lane;
      bool in = a < c.A;
      v.P[r + a] = in ? probs[a] : 0.f;
      v.W[r + a] = in ? value : 0.f;  // children start from the parent's value (mcts.jl:211)
    }
    apply_path(path, plen, OP_BACKUP, value);
  }

  // ---- tree_search! second half (mcts_play.jl:88-96) ---------------------------------------------
  AGZ_DEV void search_incorporate() {
    const bool seed_mode = st.seed_round != 0;
    const int nleaf = st.nleaf;
    int n_dup = 0;
    // Lane k fetches everything leaf k needs up front (node id, path length, value, meta word): the leaves are processed in
    // order, but their inputs do not depend on each other except for "expanded by an earlier leaf of this batch", which is the
    // same node id appearing earlier in the batch.
    for (int k0 = 0; k0 < nleaf && !st.err; k0 += 32) {   // 32 leaves per pass (parallel_readouts can exceed the warp width)
      int my_leaf = -1, my_plen = 0, my_flags = 0;
      float my_value = 0.f;
      if (k0 + lane < nleaf) {
        const size_t b = (size_t)g * c.pmax + k0 + lane;
        my_leaf = v.leaf_node[b];
        my_plen = v.leaf_plen[b];
        my_value = v.eval_v[b * v.v_stride];
        my_flags = load_meta(my_leaf).flags;
      }
      simt::sync();  // every lane has read its flags before any lane rewrites them below
      const int kn = nleaf - k0 < 32 ? nleaf - k0 : 32;
      PathEnt pre;
      pre.slot = SLOT_ROOT; pre.node = 0; pre.to_play = 0;
      if (lane < simt::shfl(my_plen, 0)) pre = path_of(k0)[lane];
      for (int kk = 0; kk < kn; ++kk) {
        const int k = k0 + kk;
        const PathEnt* path = path_of(k);
        const int leaf = simt::shfl(my_leaf, kk), plen = simt::shfl(my_plen, kk), flags = simt::shfl(my_flags, kk);
        PathEnt nxt = pre;   // the next leaf's path entries travel while this leaf is finished
        {
          const int plen_next = simt::shfl(my_plen, (kk + 1) & 31);
          if (kk + 1 < kn && lane < plen_next) nxt = path_of(k + 1)[lane];
        }
        const float value = simt::shfl(my_value, kk);
        const size_t b = (size_t)g * c.pmax + k;
        // revert_virtual_loss!, then incorporate_results! (mcts_play.jl:92-95, mcts.jl:188-213)
        if (flags & F_DONE) { st.err = E_ASSERT; break; }  // @assert !position.done (mcts.jl:196)
        const unsigned same = simt::ballot(lane < kk && my_leaf == leaf);
        const bool dup = (flags & F_EXPANDED) != 0 || same != 0u;   // already expanded (:197-200): revert_visits!
        n_dup += dup ? 1 : 0;
        if (!dup) {
          if (((simt::fbits(value) << 1) - 1u) < ((57u << 24) - 1u)) st.tiny_values = 1;   // 0 < |value| < 2^-70 (select_leaf)
          if (lane == 0) v.meta[nbase + leaf].flags = (uint8_t)(flags | F_EXPANDED);
          const float* probs = v.eval_pi + b * v.pi_stride + lane;
          const size_t r = row(leaf) + lane;
          float* Pr = v.P + r;
          float* Wr = v.W + r;
          const int left = c.A - lane;   // lane's entries q*32 with q*32 < left are actions
#pragma unroll
          for (int q = 0; q < KA; ++q) {
            const bool in = q * 32 < left;
            Pr[q * 32] = in ? probs[q * 32] : 0.f;
            Wr[q * 32] = in ? value : 0.f;  // children start from the parent's value (mcts.jl:211)
          }
        }
        finish_path(path, plen, !seed_mode, dup, value, pre);
        pre = nxt;
      }
      simt::sync();  // the flag writes of this pass are visible to the next pass's loads
    }
    if (n_dup) count(CTR_DUP_LEAVES, (unsigned long long)n_dup);
    st.nleaf = 0;
  }

  // revert_virtual_loss! followed by backup_value!(value) (fresh leaf) or by revert_visits! (duplicate): one read-modify-write per
  // path entry instead of two; every W slot sees the same two fp32 additions in the same order as the two separate passes.
  // `pre` = path[lane] (lane < plen), loaded by the caller while the previous leaf was being finished: the entry loads of successive
  // leaves are independent, only their read-modify-writes (shared ancestors) are ordered.
  AGZ_DEV void finish_path(const PathEnt* path, int plen, bool had_vloss, bool dup, float value, const PathEnt& pre) {
    simt::sync();
    for (int d0 = 0; d0 < plen; d0 += 32) {
      int d = d0 + lane;
      if (d < plen) {
        PathEnt e = pre;
        if (d0) e = path[d];
        if (e.slot != SLOT_ROOT) {
          if (had_vloss || !dup) {
            float w = v.W[e.slot];
            if (had_vloss) w = simt::fadd(w, (float)(-e.to_play));
            if (!dup) w = simt::fadd(w, value);
            v.W[e.slot] = w;
          }
          if (dup) v.N[e.slot] = simt::fsub(v.N[e.slot], 1.0f);
        }
      }
    }
    const int root0 = simt::shfl((int)(pre.slot == SLOT_ROOT ? 1 : 0), 0), tp0 = simt::shfl(pre.to_play, 0);   // path[0]
    if (root0) {
      if (had_vloss) st.root_W = simt::fadd(st.root_W, (float)(-tp0));
      if (!dup) st.root_W = simt::fadd(st.root_W, value);
      if (dup) st.root_N = simt::fsub(st.root_N, 1.0f);
    }
    if (had_vloss) st.vloss_balance -= plen;
    simt::sync();
  }

  // ---- inject_noise! (mcts.jl:233-239) -------------------------------------------------------------
  AGZ_DEV void inject_noise() {
    const size_t r = row(st.root);
    const uint32_t move_no = (uint32_t)load_meta(st.root).n;
    const uint32_t call = st.noise_ctr++;
    // The A gamma variates are rejection samples whose draws are keyed by (action, attempt), not by lane, so the lanes share them
    // as a work list: a lane that gets its variate accepted takes the next unassigned action.  A fixed three-actions-per-lane
    // split runs 3 x (the slowest of 32 rejection loops) attempts; this runs ~A x 1.35 / 32.  This is most of the per-move tail
    // of the incorporate kernel (the warps of games that do not move have long finished).
    double* gs = v.noise_g + (size_t)g * (KA * 32);
    {
      const double b = simt::dadd(1.0, simt::ddiv(c.noise_alpha, 2.718281828459045));
      const unsigned lt = (1u << lane) - 1u;
      int a = lane, next = 32;
      uint32_t t = 0;
      for (;;) {
        const bool have = a < c.A;
        if (!simt::any(have)) break;
        bool fin = false;
        if (have) {
          double x = 0.0;
          fin = gamma_small_attempt(c.noise_alpha, b, c.seed, st.game_id_lo, move_no, (uint32_t)a, call, t, x) || t == 63;
          ++t;
          if (fin) gs[a] = x;
        }
        const unsigned done = simt::ballot(fin);
        if (fin) {
          a = next + simt::popc(done & lt);
          t = 0;
        }
        next += simt::popc(done);
      }
      simt::sync();
    }
    double gm[KA];
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < KA; ++k) {
      int a = k * 32 + lane;
      gm[k] = 0.0;
      if (a < c.A) {
        gm[k] = gs[a];
        acc = simt::dadd(acc, gm[k]);
      }
    }
    const double total = butterfly_sum(acc);
    const double keep = simt::dsub(1.0, c.noise_weight);
#pragma unroll
    for (int k = 0; k < KA; ++k) {
      int a = k * 32 + lane;
      if (a < c.A) {
        double mixed = simt::dadd(simt::dmul((double)v.P[r + a], keep), simt::dmul(simt::ddiv(gm[k], total), c.noise_weight));
        v.P[r + a] = (float)mixed;
      }
    }
    simt::sync();
  }

  // ---- pick_move (mcts_play.jl:52-71): returns the flat move, or -1 with st.err = E_ASSERT -------------
  AGZ_DEV int pick_move() {
    const size_t r = row(st.root);
    const NodeMeta m = load_meta(st.root);
    float n[KA];
#pragma unroll
    for (int k = 0; k < KA; ++k) {
      int a = k * 32 + lane;
      n[k] = a < c.A ? v.N[r + a] : -1.f;
    }
    if (m.n >= c.tau_threshold) {
      float mx = -1.f;
#pragma unroll
      for (int k = 0; k < KA; ++k) mx = n[k] > mx ? n[k] : mx;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        float o = simt::shfl_xor(mx, off);
        mx = o > mx ? o : mx;
      }
      unsigned tm[KA];
      int total = 0;
#pragma unroll
      for (int k = 0; k < KA; ++k) {
        tm[k] = simt::ballot(n[k] == mx);
        total += simt::popc(tm[k]);
      }
      U4 rr = rng_draw(c.seed, st.game_id_lo, SITE_PICK_MAX, (uint32_t)m.n, 0, 0);
      int pick = (int)simt::mulhi(rr.x, (uint32_t)total);
      int best = -1;
#pragma unroll
      for (int k = 0; k < KA; ++k) {
        int cntk = simt::popc(tm[k]);
        if (best < 0) {
          if (pick < cntk) {
            best = k * 32 + nth_set_bit(tm[k], pick);
          } else {
            pick -= cntk;
          }
        }
      }
      return best;
    }
    // soft pick: cdf = cumsum(child_N) / cdf[end-1]; first index with cdf >= rand()
    float cum[KA];
    float carry = 0.f;
#pragma unroll
    for (int k = 0; k < KA; ++k) {
      float x = n[k] > 0.f ? n[k] : 0.f;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        float o = simt::shfl(x, lane - off);
        if (lane >= off) x = simt::fadd(x, o);
      }
      cum[k] = simt::fadd(x, carry);
      carry = simt::shfl(cum[k], 31);
    }
    const int last = c.A - 2;  // cdf[end-1]: pass is excluded from the normaliser (in a game without pass: the last board point, as the reference does)
    float denom = 0.f;
#pragma unroll
    for (int k = 0; k < KA; ++k)
      if (k == (last >> 5)) denom = cum[k];
    denom = simt::shfl(denom, last & 31);
    if (denom == 0.f) { st.err = E_ASSERT; return -1; }
    U4 rr = rng_draw(c.seed, st.game_id_lo, SITE_PICK_SOFT, (uint32_t)m.n, 0, 0);
    const double sel = u53(rr.x, rr.y);
    int best = -1;
#pragma unroll
    for (int k = 0; k < KA; ++k) {
      int a = k * 32 + lane;
      bool ge = a < c.A && (double)simt::fdiv(cum[k], denom) >= sel;
      unsigned mk = simt::ballot(ge);
      if (best < 0 && mk) best = k * 32 + simt::ffs(mk) - 1;
    }
    if (best < 0) { st.err = E_ASSERT; return -1; }
    float nb = 0.f;
#pragma unroll
    for (int k = 0; k < KA; ++k)
      if (k == (best >> 5)) nb = n[k];
    nb = simt::shfl(nb, best & 31);
    if (nb == 0.f) { st.err = E_ASSERT; return -1; }  // @assert child_N[fcoord] != 0 (mcts_play.jl:68)
    return best;
  }

  // ---- arena compaction: keep the subtree of the root, slide it to the front (stable) -------------------
  // Pass 1 marks live nodes (a parent always has a smaller id than its child) and builds remap (old -> new) and order
  // (new -> old).  Pass 2 walks only the live nodes, U at a time: all loads of a group are issued before its stores, so the
  // dependent-load chains of U nodes overlap (new id <= old id and ascending order make the in-place move safe).
  AGZ_DEV void compact() {
    int32_t* remap = v.remap + (size_t)g * c.cap;
    int32_t* order = v.order + (size_t)g * c.cap;
    const int count = st.count;
    int base = 0;
    for (int c0 = 0; c0 < count; c0 += 32) {
      int i = c0 + lane;
      int parent = -2;
      if (i < count) parent = load_meta(i).parent;
      unsigned mask = 0;
      for (;;) {  // resolve same-chunk parents by iteration
        bool live = false;
        if (i < count) {
          if (i == st.root) live = true;
          else if (parent >= c0) live = (mask >> (parent - c0)) & 1u;
          else if (parent >= 0) live = remap[parent] >= 0;
        }
        unsigned nm = simt::ballot(live);
        if (nm == mask) break;
        mask = nm;
      }
      if (i < count) {
        const bool live = (mask >> lane) & 1u;
        const int ni = base + simt::popc(mask & ((1u << lane) - 1u));
        remap[i] = live ? ni : -1;
        if (live) order[ni] = i;
      }
      base += simt::popc(mask);
      simt::sync();
    }
    const int nlive = base;
    constexpr int U = KA <= 3 ? 4 : (KA <= 6 ? 2 : 1);
    for (int n0 = 0; n0 < nlive; n0 += U) {
      float nn[U][KA], ww[U][KA], pp[U][KA];
      int cc[U][KA];
      NodeMeta mm[U];
      uint32_t bb[U][(3 * KA + 31) / 32 + 3];
      int src[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        src[u] = n0 + u < nlive ? order[n0 + u] : -1;
        if (src[u] >= 0) {
          const size_t ro = row(src[u]);
#pragma unroll
          for (int k = 0; k < KA; ++k) {
            int a = k * 32 + lane;
            nn[u][k] = v.N[ro + a]; ww[u][k] = v.W[ro + a]; pp[u][k] = v.P[ro + a]; cc[u][k] = v.child[ro + a];
          }
          mm[u] = load_meta(src[u]);
          const uint32_t* bo = bits_of(src[u]);
#pragma unroll
          for (int q = 0; q < (3 * KA + 31) / 32 + 3; ++q) {
            int k = q * 32 + lane;
            bb[u][q] = k < 3 * c.KB ? bo[k] : 0u;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (src[u] >= 0) {
#pragma unroll
          for (int k = 0; k < KA; ++k)
            if (cc[u][k] >= 0) cc[u][k] = remap[cc[u][k]];
        }
      simt::sync();  // every lane has loaded its part of the U sources before any destination is written
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (src[u] >= 0) {
          const int ni = n0 + u;
          const size_t rn = row(ni);
#pragma unroll
          for (int k = 0; k < KA; ++k) {
            int a = k * 32 + lane;
            v.N[rn + a] = nn[u][k]; v.W[rn + a] = ww[u][k]; v.P[rn + a] = pp[u][k]; v.child[rn + a] = cc[u][k];
          }
          uint32_t* bn = bits_of(ni);
#pragma unroll
          for (int q = 0; q < (3 * KA + 31) / 32 + 3; ++q) {
            int k = q * 32 + lane;
            if (k < 3 * c.KB) bn[k] = bb[u][q];
          }
          if (lane == 0) {
            NodeMeta m = mm[u];
            m.parent = m.parent >= 0 ? remap[m.parent] : -1;
            v.meta[nbase + ni] = m;
          }
        }
      simt::sync();
    }
    st.root = remap[st.root];
    st.count = nlive;
    simt::sync();
  }

  // ---- play_move!(player, c) (mcts_play.jl:26-50): record pi / q, re-root on the chosen child -----------
  AGZ_DEV int play_move(int mv, bool record) {
    const int root = st.root;
    const NodeMeta rm = load_meta(root);
    const size_t r = row(root);
    int child = v.child[r + mv];
    if (child < 0) {
      child = create_child(root, rm, mv, true);
      if (child < 0) return st.err;
      if (lane == 0) v.child[r + mv] = child;
      simt::sync();
    }
    const int t = st.n_moves;
    if (record && t < c.max_game_length + 2) {
      const size_t rb = ((size_t)g * (c.max_game_length + 2) + t);
      if (c.tau_threshold >= 0) {  // searches_pi is only kept outside two_player_mode
        float n[KA];
#pragma unroll
        for (int k = 0; k < KA; ++k) {
          int a = k * 32 + lane;
          n[k] = a < c.A ? v.N[r + a] : 0.f;
        }
        if (rm.n <= c.tau_threshold) {  // squash: child_N .^ 0.98 in Float64 (mcts.jl:247-251)
          double x[KA], acc = 0.0;
#pragma unroll
          for (int k = 0; k < KA; ++k) {
            int a = k * 32 + lane;
            x[k] = 0.0;
            if (a < c.A) { x[k] = det_pow((double)n[k], 0.98); acc = simt::dadd(acc, x[k]); }
          }
          double total = butterfly_sum(acc);
#pragma unroll
          for (int k = 0; k < KA; ++k) {
            int a = k * 32 + lane;
            if (a < c.A) v.rec_pi[rb * c.A + a] = (float)simt::ddiv(x[k], total);
          }
        } else {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < KA; ++k) acc = simt::fadd(acc, n[k]);
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) acc = simt::fadd(acc, simt::shfl_xor(acc, off));
#pragma unroll
          for (int k = 0; k < KA; ++k) {
            int a = k * 32 + lane;
            if (a < c.A) v.rec_pi[rb * c.A + a] = simt::fdiv(n[k], acc);
          }
        }
#pragma unroll
        for (int k = 0; k < KA; ++k) {
          int a = k * 32 + lane;
          if (a < c.A) v.rec_vis[rb * c.A + a] = n[k];
        }
      }
      if (lane == 0) {
        v.rec_q[rb] = simt::fdiv(st.root_W, simt::fadd(1.0f, st.root_N));  // Q(root) (mcts_play.jl:38)
        v.rec_moves[rb] = (int16_t)mv;
      }
      st.n_moves = t + 1;
    }
    // the child keeps its own statistics; its siblings are dropped (mcts_play.jl:40,48)
    st.root_N = v.N[r + mv];
    st.root_W = v.W[r + mv];
    uint32_t* h = v.hist + (size_t)g * (7 * 2 * c.KB);
    const uint32_t* rbits = bits_of(root);
    simt::sync();
    // shift the history ring by one board through registers: 7*2*KB <= 168 words -> at most 6 per lane
    {
      const int W2 = 2 * c.KB, total = 7 * W2;
      uint32_t tmp[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        int k = q * 32 + lane;
        tmp[q] = 0;
        if (k < total) tmp[q] = k < W2 ? rbits[k] : h[k - W2];
      }
      simt::sync();
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        int k = q * 32 + lane;
        if (k < total) h[k] = tmp[q];
      }
    }
    st.hist_len = st.hist_len < 7 ? st.hist_len + 1 : 7;
    if (lane == 0) v.meta[nbase + child].parent = -1;
    st.root = child;
    st.sel_ctr = 0;
    st.noise_ctr = 0;
    simt::sync();
    const int need = c.readouts + 2 * c.pmax + 4;
    if (st.count + need > c.cap) {
      compact();
      // The reference's tree is unbounded; here the kept subtree plus one move's growth must fit the game's arena.  Under that
      // pressure (a very sharp network on a small arena) the least-visited nodes are forgotten, N <= 1 first, then 2, 4, ...: their
      // statistics stay in the parent's rows (the PUCT scores do not change), only their own expansion is redone when they are
      // visited again.  Counted in agz_progress.arena_prunes; a game stops with AGZ_ERR_CAPACITY only if even that does not help.
      for (float thr = 1.f; st.count + need > c.cap && thr < 1.0e9f; thr *= 2.f) {
        forget_leaves(thr);
        compact();
        count(CTR_PRUNES, 1);
      }
      if (st.count + need > c.cap) { st.err = E_CAPACITY; return st.err; }
    }
    return E_OK;
  }

  // detach every node (other than the root) whose own visit count is <= thr; compact() then drops it and its descendants
  AGZ_DEV void forget_leaves(float thr) {
    const int count = st.count;
    for (int c0 = 0; c0 < count; c0 += 32) {
      const int i = c0 + lane;
      bool drop = false;
      if (i < count && i != st.root) {
        const NodeMeta m = load_meta(i);
        if (m.parent >= 0) drop = v.N[row(m.parent) + m.fmove] <= thr;
      }
      simt::sync();
      if (drop) v.meta[nbase + i].parent = -2;
    }
    simt::sync();
  }

  // ---- get_feats(node.position) (features.jl:3-26): the last 8 boards come from the node, its ancestors in
  // the tree and, past the root, the per-game ring of earlier root boards; the oldest one is repeated.
  // emit(plane_pair h, point p, mine, theirs) is called for every on-board point of every history board.
  template <class Emit>
  AGZ_DEV int gather_features(int node, Emit emit) {
    const NodeMeta m = load_meta(node);
    const int tp = m.to_play;
    int cur = node;
    const uint32_t* src = bits_of(cur);
    const uint32_t* h = v.hist + (size_t)g * (7 * 2 * c.KB);
    bool in_tree = true;
    int ri = 0;
    for (int hb = 0; hb < 8; ++hb) {
      for (int k = 0; k < c.KB; ++k) {
        int p = k * 32 + lane;
        uint32_t b = (src[k] >> lane) & 1u, w = (src[c.KB + k] >> lane) & 1u;
        if (p < c.N2) emit(hb, p, tp == 1 ? b : w, tp == 1 ? w : b);
      }
      if (in_tree) {
        int pm = load_meta(cur).parent;
        if (pm >= 0) { cur = pm; src = bits_of(cur); continue; }
        in_tree = false;
      }
      if (ri < st.hist_len) { src = h + (size_t)ri * 2 * c.KB; ++ri; }
    }
    return tp;
  }

  // ---- new game in this slot (initialize_game! + selfplay.jl:9); the root position is `pos` ----------------
  AGZ_DEV void init_root_from_scratch(int n, int ko, int to_play, int flags) {
    st.root = 0;
    st.count = 1;
    st.root_N = 0.f;
    st.root_W = 0.f;
    st.sel_ctr = 0;
    st.noise_ctr = 0;
    st.n_moves = 0;
    st.nleaf = 0;
    st.vloss_balance = 0;
    st.tiny_values = 0;
    st.err = 0;
    st.result = 0;
    st.resigned = 0;
    bool term = (flags & F_DONE) || n >= c.max_game_length;
    write_node(0, -1, -1, n, ko, to_play, flags, term);
    simt::sync();
  }

  AGZ_DEV void start_game(long long game_id) {
    st.game_id = game_id;
    st.game_id_lo = (uint32_t)game_id;
    st.hist_len = 0;
    pos.b = 0;
    pos.w = 0;
    init_root_from_scratch(0, -1, 1, 0);
    U4 rr = rng_draw(c.seed, st.game_id_lo, SITE_RESIGN, 0, 0, 0);
    st.resign_thr = u53(rr.x, rr.y) < c.resign_disable_frac ? -1.0 : c.resign_threshold;
    st.phase = PH_SEED;
    count(CTR_STARTED, 1);
  }

  // ---- game end: publish the record into the finished ring, then refill the slot --------------------
  AGZ_DEV bool publish() {
    unsigned long long t = 0;
    int ok = 0;
    if (lane == 0) {
#if AGZ_CUDA
      for (;;) {
        t = *((volatile unsigned long long*)&v.ctr[CTR_RING_TAIL]);
        unsigned long long h = *((volatile unsigned long long*)&v.ctr[CTR_RING_HEAD]);
        if (t - h >= (unsigned long long)c.ring_cap) break;
        if (atomicCAS(&v.ctr[CTR_RING_TAIL], t, t + 1) == t) { ok = 1; break; }
      }
#else
      t = v.ctr[CTR_RING_TAIL];
      if (t - v.ctr[CTR_RING_HEAD] < (unsigned long long)c.ring_cap) { v.ctr[CTR_RING_TAIL] = t + 1; ok = 1; }
#endif
    }
    ok = simt::shfl(ok, 0);
    if (!ok) return false;
    const int rslot = simt::shfl((int)(t % (unsigned long long)c.ring_cap), 0);
    const int L = c.max_game_length + 2;
    const size_t src = (size_t)g * L, dst = (size_t)rslot * L;
    const int nm = st.n_moves;
    for (int i = lane; i < nm; i += 32) {
      v.ring_moves[dst + i] = v.rec_moves[src + i];
      v.ring_q[dst + i] = v.rec_q[src + i];
    }
    const size_t tot = (size_t)nm * c.A;
    for (size_t i = lane; i < tot; i += 32) {
      v.ring_pi[dst * c.A + i] = v.rec_pi[src * c.A + i];
      v.ring_vis[dst * c.A + i] = v.rec_vis[src * c.A + i];
    }
    if (lane == 0) {
      RingHeader hd;
      hd.game_id = st.game_id; hd.n_moves = nm; hd.result = st.result; hd.resigned = st.resigned;
      hd.final_score = st.final_score; hd.resign_threshold = st.resign_thr;
      v.ring_hdr[rslot] = hd;
    }
    count(CTR_FINISHED, 1);
    return true;
  }

  AGZ_DEV void finish_or_wait() {
    if (!publish()) { st.phase = PH_WAIT_RING; return; }
    long long next = st.game_id + (long long)c.n_games * c.world;
    if (c.total_games < 0 || next < c.total_games) start_game(next);
    else st.phase = PH_IDLE;
  }

  // ---- selfplay.jl:22-43, evaluated once per round after the leaves have been incorporated -----------
  // Nearly always nothing is due (the search has not reached its visit target yet): the kernels test that inline and call the
  // per-move logic out of line (after_round_cold below), which keeps pick_move / play_move! / compaction / noise / game end --
  // nine tenths of the code -- out of the instruction stream of the search loop.
  AGZ_DEV bool after_round_due() const {
    const bool searching = (st.phase == PH_SEARCH || st.phase == PH_MATCH_SEARCH) && st.root_N < st.target_N;
    const bool idle = st.phase == PH_IDLE || st.phase == PH_MANUAL || st.phase == PH_MATCH_WAIT || (st.phase == PH_SEED && !st.seed_round);
    return !(st.err == 0 && (searching || idle));
  }

  AGZ_DEV void after_round() {
    if (st.err) {
      if (st.phase == PH_MATCH_SEARCH && lane == 0) simt::atomic_add(&v.ctr[CTR_MATCH_BUSY], ~0ULL);
      st.phase = PH_IDLE;
      return;
    }
    if (st.phase == PH_WAIT_RING) { finish_or_wait(); return; }
    if (st.phase == PH_DELAY) {
      if (--st.delay <= 0) start_game(st.game_id);
      return;
    }
    if (st.phase == PH_SEED) {
      if (st.seed_round) {
        st.phase = PH_SEARCH;
        if (c.inject_noise) inject_noise();
        st.target_N = simt::fadd(st.root_N, (float)c.readouts);
      }
      return;
    }
    if (st.phase == PH_MATCH_SEARCH) {  // evaluate / play: `while N(root) < current + readouts` (neural_net.jl:124-126); the host picks
      if (!(st.root_N < st.target_N)) {
        st.phase = PH_MATCH_WAIT;
        if (lane == 0) simt::atomic_add(&v.ctr[CTR_MATCH_BUSY], ~0ULL);
      }
      return;
    }
    if (st.phase != PH_SEARCH) return;
    if (st.root_N < st.target_N) return;  // while N(root) < current_readouts + readouts (selfplay.jl:27)
    const NodeMeta rm = load_meta(st.root);
    const float q = simt::fdiv(st.root_W, simt::fadd(1.0f, st.root_N));
    const float qp = simt::fmul(q, (float)rm.to_play);
    if ((double)qp < st.resign_thr) {  // should_resign (mcts_play.jl:124)
      st.result = -rm.to_play;
      st.resigned = 1;
      st.final_score = 0.f;
      finish_or_wait();
      return;
    }
    int mv = pick_move();
    if (mv < 0) { st.phase = PH_IDLE; return; }
    if (play_move(mv, true) != E_OK) { st.phase = PH_IDLE; return; }
    count(CTR_MOVES, 1);
    const NodeMeta nm = load_meta(st.root);
    if (terminal(nm)) {  // is_done(root) (selfplay.jl:39-42)
      const uint32_t* lb = bits_of(st.root);
      float sc = game_score(c, B, bits_load(B, lb, lb + c.KB));
      st.final_score = sc;
      st.result = sc > 0.f ? 1 : (sc < 0.f ? -1 : 0);
      st.resigned = 0;
      finish_or_wait();
      return;
    }
    if (c.inject_noise) inject_noise();
    st.target_N = simt::fadd(st.root_N, (float)c.readouts);
  }
};

// the per-move logic of one game, out of line: the caller has stored the game's state and reloads it afterwards
template <int KA>
AGZ_COLD void after_round_cold(const Cfg* c, const View* v, int g, char* smem) {
  Warp<KA> w(*c, *v, g, smem);
  w.after_round();
  w.store_state();
  simt::sync();
}

}  // namespace agz
