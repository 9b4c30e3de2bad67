// tree_duo.cuh -- two search trees per warp: each half of a warp (16 lanes) owns one game.
//
// Same reference functions and the same tree in HBM as tree.cuh (src/mcts.jl select_leaf :108-138, maybe_add_child! :140-147,
// add_virtual_loss! :149-163, incorporate_results! :188-213, backup_value! :215-225; src/mcts_play.jl tree_search! :73-98), for
// boards whose lines fit 16 lanes (N <= 9 here: KA = 3).  Why: with one warp per tree the tree kernels are issue bound, and ~85 % of the
// ~1600 warp instructions of a readout are warp-uniform bookkeeping (addresses, loop control, the Go rules on 9 of 32 lanes, the
// tie-break draw) rather than per-child arithmetic.  Two trees in one warp share every one of those instructions: lane l of a half
// scores children l, l+16, ..., l+80 (6 per lane instead of 3), the rules code holds board line j of tree h in lane 16h + j, and
// the per-warp reductions become one segmented step.
//
// The two halves run in LOCK STEP (simt.h G<16>): control flow is warp-uniform wherever a collective is reached, and what differs
// between the halves -- how deep a descent goes, whether its leaf is terminal, whether a child has to be created, how many leaves a
// round still needs -- is a per-half predicate (`on`, `alive`, ...) on the memory operations and state updates.  Each tree sees exactly
// the operations, in exactly the order, of the one-warp-per-tree code, so results are bit-identical (same tests, both modes).
// The per-move logic (pick_move, play_move!, inject_noise!, compaction, game end: once per `readouts` visits) is not duplicated:
// the warp runs tree.cuh's Warp<KA>::after_round for one half's game at a time with all 32 lanes.
#pragma once
#include "tree.cuh"

namespace agz {

template <int KA>
struct Duo {
  typedef simt::G<16> S;
  static constexpr int KV = 2 * KA;  // statistics per lane: action a = 16*k + lane, k < KV (rows are AS = 32*KA floats)
  const Cfg& c;
  const View& v;
  int g;       // this half's game slot (a valid slot even when `valid` is false, so that addresses stay in bounds)
  bool valid;  // false for the upper half of the last warp when the number of slots is odd: never `on`, never stored
  const int lane;
  BitsCtxT<16> B;
  Lines pos;
  GameState st;  // register copy, uniform within the half
  size_t nbase;
  unsigned long long reg_slot;  // lane d keeps path entry d of the last descent (d < 16)
  int reg_tp;

  AGZ_DEV Duo(const Cfg& c_, const View& v_, int g0, int g_end) : c(c_), v(v_), lane(S::lane()) {
    g = g0 + S::half();
    valid = g < g_end;
    if (!valid) g = g0;
    B = bits_ctx_w<16>(c.N, c.KB);
    pos.b = 0;
    pos.w = 0;
    st = v.gs[g];
    nbase = (size_t)g * c.cap;
    reg_slot = SLOT_ROOT;
    reg_tp = 1;
  }
  AGZ_DEV void store_state() {
    S::sync();
    if (lane == 0 && valid) v.gs[g] = st;
  }
  AGZ_DEV size_t row(int node) const { return (nbase + node) * (size_t)(KA * 32); }
  AGZ_DEV uint32_t* bits_of(int node) const { return v.bits + (nbase + node) * (size_t)(3 * KA); }
  template <class T>
  AGZ_DEV static T pick(const T (&a)[KV], int idx) {
    T x = a[0];
#pragma unroll
    for (int k = 1; k < KV; ++k) x = idx == k ? a[k] : x;
    return x;
  }
  AGZ_DEV NodeMeta load_meta(int node) const {
#if AGZ_CUDA
    NodeMeta m;
    *reinterpret_cast<uint4*>(&m) = *reinterpret_cast<const uint4*>(v.meta + nbase + node);
    return m;
#else
    return v.meta[nbase + node];
#endif
  }
  AGZ_DEV bool terminal(const NodeMeta& m) const { return (m.flags & F_DONE) || m.n >= c.max_game_length; }
  AGZ_DEV PathEnt* path_of(int k) const { return v.path + ((size_t)g * c.pmax + k) * c.maxd; }
  AGZ_DEV void count(int which, unsigned long long n) {
    if (lane == 0 && valid && n) simt::atomic_add(&v.ctr[which], n);
  }

  // ---- node creation (tree.cuh write_node / create_child; check_legal is never needed on a descent) --------------------------
  AGZ_DEV NodeMeta write_node(bool on, int node, int parent, int fmove, int n, int ko, int to_play, int flags, bool skip_legal) {
    uint32_t bw[KA], ww[KA], lw[KA];
    bits_pack<KA>(B, pos.b, bw);
    bits_pack<KA>(B, pos.w, ww);
    uint32_t legal = 0;
    if (S::any_warp(on && !skip_legal)) {
      const uint32_t l2 = bits_legal(B, pos, to_play, ko);
      legal = skip_legal ? 0u : l2;
    }
    bits_pack<KA>(B, legal, lw);
    NodeMeta m;
    m.parent = parent; m.fmove = (int16_t)fmove; m.n = (int16_t)n; m.ko = (int16_t)ko;
    m.to_play = (int8_t)to_play; m.flags = (uint8_t)flags; m.pad = 0;
    if (on) {
      uint32_t* bp = bits_of(node);
#pragma unroll
      for (int k = 0; k < KA; ++k) {
        if (k < c.KB && lane == k) {
          bp[k] = bw[k];
          bp[c.KB + k] = ww[k];
          bp[2 * c.KB + k] = lw[k];
        }
      }
      if (lane == 0) v.meta[nbase + node] = m;
      const size_t r = row(node);
#pragma unroll
      for (int k = 0; k < KV; ++k) {  // child_N = 0, no children; W / P rows are first written by incorporate_results!
        const int a = k * 16 + lane;
        v.N[r + a] = 0.f;
        v.child[r + a] = -1;
      }
    }
    return m;
  }

  // maybe_add_child! for a missing child (mcts.jl:140-147 -> board.jl:451-509).  Returns the new node id, or -1 with st.err set.
  AGZ_DEV int create_child(bool on, int parent, const NodeMeta& pm, int move, NodeMeta& cm) {
    if (on && st.count >= c.cap) {
      st.err = E_CAPACITY;
      on = false;
    }
    const int idx = on ? st.count : -1;
    if (on) st.count++;
    const uint32_t* pb = bits_of(parent);
    pos = bits_load(B, pb, pb + c.KB);
    const int color = pm.to_play;
    const int n = pm.n + 1;
    const bool is_pass = move == c.N2;
    int ko = -1, flags = 0;
    if (S::any_warp(on && !is_pass)) {
      Lines L2 = pos;
      int ko2 = -1, nc2 = 0;
      bits_play(B, L2, is_pass ? 0 : move, color, false, ko2, nc2);
      if (!is_pass) { pos = L2; ko = ko2; }
    }
    if (is_pass) flags = F_LASTPASS | ((pm.flags & F_LASTPASS) ? F_DONE : 0);  // pass_move! (board.jl:426-440)
    const bool term = (flags & F_DONE) || n >= c.max_game_length;
    cm = write_node(on, on ? idx : 0, parent, move, n, ko, -color, flags, term);
    S::sync();
    return idx;
  }

  // ---- select_leaf from the root (mcts.jl:108-138) ----------------------------------------------------------------------
  // `on` = this half performs a descent.  Returns the leaf; plen / lm = path length and the leaf's meta word.
  AGZ_DEV int select_leaf(bool on, PathEnt* path, int& plen, NodeMeta& lm) {
    const uint32_t sel_idx = st.sel_ctr;
    int cur = st.root, depth = 0, move_no = 0;
    unsigned long long slot = SLOT_ROOT;
    float cur_N = 0.f;
    if (on) {
      st.sel_ctr++;
      st.root_N = simt::fadd(st.root_N, 1.0f);  // N(root) += 1 (mcts.jl:113-114); the root's own N / W live in the game state
      cur_N = st.root_N;
    }
    bool alive = on;
    int c_move = -1, c_parent = st.root;  // a descent that ends on a missing child creates it after the loop, both halves together
    NodeMeta c_pm = lm;
    size_t c_row = 0;
    float c_nnew = 0.f;
    const int pass = c.N2;
    while (S::any_warp(alive)) {
      // one dependent memory round trip per level: meta word, the four statistic rows and the legal-move words together
      const size_t r = row(cur);
      const NodeMeta m = load_meta(cur);
      float n[KV], w[KV], p[KV];
      int ch[KV];
      uint32_t lwv[KA];
      const uint32_t* lw = bits_of(cur) + 2 * c.KB;
#pragma unroll
      for (int k = 0; k < KV; ++k) {
        const int a = k * 16 + lane;
        n[k] = 0.f; w[k] = 0.f; p[k] = 0.f; ch[k] = -1;
        if (alive) {
          n[k] = v.N[r + a];
          w[k] = v.W[r + a];
          p[k] = v.P[r + a];
          ch[k] = v.child[r + a];
        }
      }
#pragma unroll
      for (int k = 0; k < KA; ++k) lwv[k] = lw[k];  // words past KB lie in the node's never-written (zero) tail
      if (depth == 0) move_no = m.n;                // position.n of the root: RNG key of the tie-break draw
      if (alive) {
        lm = m;
        if (lane == 0) {
          PathEnt e;
          e.slot = slot; e.node = cur; e.to_play = m.to_play;
          path[depth] = e;
        }
        if (lane == depth) { reg_slot = slot; reg_tp = m.to_play; }
        if (!(m.flags & F_EXPANDED)) alive = false;
        else if (depth + 1 >= c.maxd) { st.err = E_ASSERT; alive = false; }
      }
      if (!S::any_warp(alive)) break;
      // HACK of the reference: after a pass, look at the double pass first (mcts.jl:119-126)
      const float n_pass = S::shfl(pick(n, pass >> 4), pass & 15);
      const bool pass_first = (m.flags & F_LASTPASS) && n_pass == 0.f;
      // score = Float64(Float32(W/(1+N)) * to_play) + ((c_puct * Float64(sqrt_f32(1+N_parent))) * Float64(P)) / Float64(1+N); the
      // two quotients by the small integer 1 + N(child) come from the reciprocal table (tree.cuh select_leaf)
      const double cu = simt::dmul(c.c_puct, (double)simt::fsqrt(simt::fadd(1.0f, cur_N)));
      const float tp = (float)m.to_play;
      float den[KV];
      bool odd = st.phase == PH_MANUAL || !(simt::fadd(cur_N, 2.0f) < (float)v.rcp_n);
      unsigned tiny = 0u;
#pragma unroll
      for (int k = 0; k < KV; ++k) {
        den[k] = simt::fadd(1.0f, n[k]);
        tiny |= ((simt::fbits(w[k]) << 1) - 1u) < ((27u << 24) - 1u) ? 1u : 0u;  // exponent field below 27 and not +-0
      }
      odd = odd || S::any(tiny != 0u);
      double s[KV];
      double mx = -1.0e300;
      if (!odd) {
        double rc[KV];
#pragma unroll
        for (int k = 0; k < KV; ++k) rc[k] = v.rcp[(int)den[k]];
#pragma unroll
        for (int k = 0; k < KV; ++k) {
          const bool legal = (((lwv[k >> 1] >> ((k & 1) * 16 + lane)) & 1u) | (unsigned)(k * 16 + lane == pass)) != 0u;  // mask bits past N^2 are 0
          const double dd = (double)den[k];
          const float q = simt::fmul((float)simt::dmul((double)w[k], rc[k]), tp);
          const double x = simt::dmul(cu, (double)p[k]);
          const double q0 = simt::dmul(x, rc[k]);
          const double u = simt::dfma(simt::dfma(-q0, dd, x), rc[k], q0);
          s[k] = legal ? simt::dadd((double)q, u) : -1.0e300;
          mx = s[k] > mx ? s[k] : mx;
        }
      } else {  // divisors past the table or subnormal quotients (test hooks only): IEEE divisions
#pragma unroll
        for (int k = 0; k < KV; ++k) {
          const bool legal = (((lwv[k >> 1] >> ((k & 1) * 16 + lane)) & 1u) | (unsigned)(k * 16 + lane == pass)) != 0u;
          const float q = simt::fmul(simt::fdiv(w[k], den[k]), tp);
          const double u = simt::ddiv(simt::dmul(cu, (double)p[k]), (double)den[k]);
          s[k] = legal ? simt::dadd((double)q, u) : -1.0e300;
          mx = s[k] > mx ? s[k] : mx;
        }
      }
      {  // maximum over the half through REDUX.MAX on an order-preserving integer key (scores are never -0.0)
        const long long b = simt::dbits(mx);
        const unsigned long long key = (unsigned long long)b ^ (b < 0 ? ~0ULL : 0x8000000000000000ULL);
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mhi = S::reduce_max(hi);
        const unsigned mlo = S::reduce_max(hi == mhi ? lo : 0u);
        const unsigned long long mkey = ((unsigned long long)mhi << 32) | mlo;
        mx = simt::bitsd((long long)(mkey ^ ((mkey >> 63) ? 0x8000000000000000ULL : ~0ULL)));
      }
      // exact ties: bit (a & 31) of word (a >> 5), the same order as the one-warp-per-tree code
      unsigned tm[KA];
      int total = 0;
#pragma unroll
      for (int k = 0; k < KA; ++k) {
        tm[k] = S::ballot(s[2 * k] == mx) | (S::ballot(s[2 * k + 1] == mx) << 16);
        total += simt::popc(tm[k]);
      }
      int best = pass;
      if (alive && !pass_first) {  // per-half work without collectives: the draw only matters when several children tie
        int pk = 0;
        if (total > 1) {
          U4 rr = rng_draw(c.seed, st.game_id_lo, SITE_SELECT, (uint32_t)move_no, sel_idx, (uint32_t)depth);
          pk = (int)simt::mulhi(rr.x, (uint32_t)total);
        }
        bool found = false;
#pragma unroll
        for (int k = 0; k < KA; ++k) {
          const int cntk = simt::popc(tm[k]);
          if (!found) {
            if (pk < cntk) {
              unsigned mk = tm[k];
              for (int t = 0; t < pk; ++t) mk &= mk - 1;  // drop the `pk` lowest set bits
              best = k * 32 + simt::ffs(mk) - 1;
              found = true;
            } else {
              pk -= cntk;
            }
          }
        }
      }
      const int ok = best >> 4, ol = best & 15;
      const float n_old = S::shfl(pick(n, ok), ol);
      const int child = S::shfl(pick(ch, ok), ol);
      if (alive) {
        const float n_new = simt::fadd(n_old, 1.0f);
        if (child < 0) {  // the missing child is this readout's leaf (created below; not expanded, so the descent ends on it)
          c_move = best; c_parent = cur; c_pm = m; c_row = r; c_nnew = n_new;
          alive = false;
        } else {
          if (lane == ol) v.N[r + best] = n_new;  // N(child) += 1 (mcts.jl:113-114)
          slot = (unsigned long long)(r + best);
          cur_N = n_new;
          cur = child;
          ++depth;
        }
      }
    }
    const bool mk = c_move >= 0;
    if (S::any_warp(mk)) {
      NodeMeta cm;
      const int child = create_child(mk, c_parent, c_pm, mk ? c_move : 0, cm);
      if (mk && child >= 0) {
        if (lane == (c_move & 15)) {
          v.child[c_row + c_move] = child;
          v.N[c_row + c_move] = c_nnew;  // N(child) += 1 (mcts.jl:113-114)
        }
        ++depth;
        if (lane == 0) {
          PathEnt e;
          e.slot = (unsigned long long)(c_row + c_move); e.node = child; e.to_play = cm.to_play;
          path[depth] = e;
        }
        if (lane == depth) { reg_slot = (unsigned long long)(c_row + c_move); reg_tp = cm.to_play; }
        lm = cm;
        cur = child;
      }
    }
    plen = depth + 1;
    S::sync();
    return cur;
  }

  // virtual loss (mcts.jl:149-163) or a terminal leaf's backup (mcts.jl:215-225) along the path of the descent that just ended
  AGZ_DEV void apply_descent(bool on, const PathEnt* path, int plen, bool backup, float value) {
    const int tp0 = S::shfl(reg_tp, 0);  // path[0] is the root
    S::sync();
    if (on) {
      if (plen <= 16) {
        if (lane < plen && reg_slot != SLOT_ROOT) v.W[reg_slot] = simt::fadd(v.W[reg_slot], backup ? value : (float)reg_tp);
      } else {
        for (int d = lane; d < plen; d += 16) {
          const PathEnt e = path[d];
          if (e.slot != SLOT_ROOT) v.W[e.slot] = simt::fadd(v.W[e.slot], backup ? value : (float)e.to_play);
        }
      }
      st.root_W = simt::fadd(st.root_W, backup ? value : (float)tp0);
      if (!backup) st.vloss_balance += plen;
    }
    S::sync();
  }

  // ---- tree_search! first half: collect leaves (mcts_play.jl:74-87) ------------------------------------------------------
  AGZ_DEV void search_select() {
    const bool seed_mode = st.phase == PH_SEED;
    int want = 0;
    if (valid && !st.err) want = seed_mode ? 1 : ((st.phase == PH_SEARCH || st.phase == PH_MATCH_SEARCH) ? c.parallel : 0);
    int nleaf = 0, attempts = 0;
    unsigned long long n_readouts = 0, n_pathnodes = 0;
    NodeMeta lm = load_meta(st.root);
    for (;;) {
      bool on = nleaf < want && attempts < 2 * want && !st.err;
      if (!S::any_warp(on)) break;
      if (on) ++attempts;
      PathEnt* path = path_of(nleaf < c.pmax ? nleaf : 0);
      int plen = 1;
      const int leaf = select_leaf(on, path, plen, lm);
      if (st.err) on = false;
      if (on) {
        n_readouts += 1;
        n_pathnodes += (unsigned long long)plen;
      }
      const bool term = on && terminal(lm);  // game over: back up the true result, do not evaluate (mcts_play.jl:80-84)
      float value = 0.f;
      if (S::any_warp(term)) {
        const uint32_t* lb = bits_of(leaf);
        const float sc = bits_score(B, bits_load(B, lb, lb + c.KB), c.komi);
        value = sc > 0.f ? 1.f : (sc < 0.f ? -1.f : 0.f);
      }
      apply_descent(on && (term || !seed_mode), path, plen, term, value);
      if (on && !term) {
        if (lane == 0) {
          v.leaf_node[(size_t)g * c.pmax + nleaf] = leaf;
          v.leaf_plen[(size_t)g * c.pmax + nleaf] = plen;
        }
        ++nleaf;
      }
    }
    st.nleaf = nleaf;
    st.seed_round = (seed_mode && want) ? 1 : 0;
    count(CTR_READOUTS, n_readouts);
    count(CTR_PATHNODES, n_pathnodes);
    count(CTR_POSITIONS, (unsigned long long)nleaf);
  }

  // revert_virtual_loss! + backup_value!(value) (fresh leaf) or + revert_visits! (duplicate), one read-modify-write per path entry
  // (tree.cuh finish_path: the same two fp32 additions in the same order)
  AGZ_DEV void finish_path(bool on, const PathEnt* path, int plen, bool had_vloss, bool dup, float value) {
    S::sync();
    if (on) {
      for (int d = lane; d < plen; d += 16) {
        const PathEnt e = path[d];
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
      const PathEnt e0 = path[0];
      if (e0.slot == SLOT_ROOT) {
        if (had_vloss) st.root_W = simt::fadd(st.root_W, (float)(-e0.to_play));
        if (!dup) st.root_W = simt::fadd(st.root_W, value);
        if (dup) st.root_N = simt::fsub(st.root_N, 1.0f);
      }
      if (had_vloss) st.vloss_balance -= plen;
    }
    S::sync();
  }

  // ---- tree_search! second half (mcts_play.jl:88-96, mcts.jl:188-213) -----------------------------------------------------
  AGZ_DEV void search_incorporate() {
    const bool seed_mode = st.seed_round != 0;
    const int nleaf = valid ? st.nleaf : 0;
    int n_dup = 0;
    bool stop = st.err != 0;
    for (int k0 = 0; S::any_warp(k0 < nleaf && !stop); k0 += 16) {
      int my_leaf = -1, my_plen = 0, my_flags = 0;
      float my_value = 0.f;
      if (!stop && k0 + lane < nleaf) {
        const size_t b = (size_t)g * c.pmax + k0 + lane;
        my_leaf = v.leaf_node[b];
        my_plen = v.leaf_plen[b];
        my_value = v.eval_v[b * v.v_stride];
        my_flags = load_meta(my_leaf).flags;
      }
      S::sync();  // every lane has read its flags before any lane rewrites them below
      const int left = nleaf - k0;
      const int kn = stop ? 0 : (left < 16 ? left : 16);
      for (int kk = 0; S::any_warp(kk < kn && !stop); ++kk) {
        const int k = k0 + kk;
        const int leaf = S::shfl(my_leaf, kk), plen = S::shfl(my_plen, kk), flags = S::shfl(my_flags, kk);
        const float value = S::shfl(my_value, kk);
        const unsigned same = S::ballot(lane < kk && my_leaf == leaf);
        bool go = kk < kn && !stop;
        if (go && (flags & F_DONE)) {  // @assert !position.done (mcts.jl:196)
          st.err = E_ASSERT;
          stop = true;
          go = false;
        }
        const bool dup = (flags & F_EXPANDED) != 0 || same != 0u;  // already expanded (:197-200): revert_visits!
        if (go) {
          n_dup += dup ? 1 : 0;
          if (!dup) {
            if (lane == 0) v.meta[nbase + leaf].flags = (uint8_t)(flags | F_EXPANDED);
            const size_t b = (size_t)g * c.pmax + k;
            const float* probs = v.eval_pi + b * v.pi_stride;
            const size_t r = row(leaf);
#pragma unroll
            for (int q = 0; q < KV; ++q) {
              const int a = q * 16 + lane;
              const bool in = a < c.A;
              v.P[r + a] = in ? probs[a] : 0.f;
              v.W[r + a] = in ? value : 0.f;  // children start from the parent's value (mcts.jl:211)
            }
          }
        }
        finish_path(go, path_of(go ? k : 0), plen, !seed_mode, dup, value);
      }
      S::sync();  // the flag writes of this pass are visible to the next pass's loads
    }
    count(CTR_DUP_LEAVES, (unsigned long long)n_dup);
    st.nleaf = 0;
  }

  // ---- selfplay.jl:22-43 after a round.  Nearly always nothing is due (the search has not reached its visit target); when something
  // is, the whole warp runs the one-warp-per-tree code for that game.
  AGZ_DEV void after_round(char* smem) {
    const bool searching = (st.phase == PH_SEARCH || st.phase == PH_MATCH_SEARCH) && st.root_N < st.target_N;
    const bool idle = st.phase == PH_IDLE || st.phase == PH_MANUAL || st.phase == PH_MATCH_WAIT || (st.phase == PH_SEED && !st.seed_round);
    const bool due = valid && !(st.err == 0 && (searching || idle));
    if (S::any_warp(due)) {
      store_state();
      simt::sync();
      for (int h = 0; h < 2; ++h) {
        if (simt::shfl((int)due, 16 * h)) {
          const int gh = simt::shfl(g, 16 * h);
          Warp<KA> w(c, v, gh, smem);
          w.after_round();
          w.store_state();
        }
      }
      simt::sync();
      st = v.gs[g];
    }
    st.seed_round = 0;
  }
};

}  // namespace agz
