// go_bits.cuh -- Go rules on bitboard lines held in registers, one warp per position (the hot-path rules code).
//
// Replaces (reference paths): src/game/go/board.jl  play_move! :451-509, pass_move! :426-440, add_stone!/captures
// :205-269, is_koish :47-56, all_legal_moves :393-424, is_move_suicidal :354-374, score :511-533.
//
// Design: the flat point index is p = N*j + i, so the N consecutive bits [N*j, N*j+N) of a stored bitplane are board
// line j.  Lane j of the warp holds line j of the black and of the white plane in two registers (lanes >= N hold 0):
// the four neighbours of every point of a set are two shifts inside the lane and two shuffles to lanes j-1 / j+1, and a
// group is a bitwise flood fill (`flood`: ~10 instructions per step for the whole board, no shared memory, no
// atomics).  The reference's incremental liberty tracker is not needed at all:
//   * captures: after the stone is placed, the opponent stones that are not reached by a flood from the opponent stones
//     that touch an empty point are exactly the captured groups (every group had a liberty before the move);
//   * legal moves: an empty point with an empty neighbour is legal; the few candidates without one are legal iff they
//     touch a friendly group that has another liberty, or an enemy group whose only liberty they are.  Friendly / enemy
//     groups that own a liberty which is not a candidate are found with one multi-source flood each; the rare groups whose
//     liberties are all candidates are visited one by one;
//   * score: empty points reached from black stones through empty points, same for white.
// (go_rules.cuh keeps a shared-memory label / liberty-count pass for the liberty-cache test hook only.)
#pragma once
#include "simt.h"

namespace agz {

struct BitsCtx {
  int N, KB, lane;
  uint32_t full;  // all on-board bits of this lane's line, 0 for lanes >= N
  float inv_n;    // 1 / N: flat point -> line without an integer division (bits_point)
};

struct Lines {
  uint32_t b, w;  // this lane's line of black / white stones
};

AGZ_DEV BitsCtx bits_ctx(int N, int KB) {
  BitsCtx B;
  B.N = N;
  B.KB = KB;
  B.lane = simt::lane();
  B.full = B.lane < N ? ((1u << N) - 1u) : 0u;
  B.inv_n = 1.0f / (float)N;
  return B;
}

// line j and position i of flat point c = N*j + i (c < 2^20): (c + 0.5) / N is at least 0.5 / N away from an integer, so the rounded
// float product truncates to the exact quotient
AGZ_DEV void bits_point(const BitsCtx& B, int c, int& cj, int& ci) {
  cj = (int)(((float)c + 0.5f) * B.inv_n);
  ci = c - cj * B.N;
}

// line `lane` of a stored bitplane (KB words, flat bit p = N*j + i)
AGZ_DEV uint32_t bits_line(const BitsCtx& B, const uint32_t* plane) {
  uint32_t r = 0;
  if (B.lane < B.N) {
    const int bit = B.N * B.lane, w0 = bit >> 5, s = bit & 31;
    const uint32_t lo = plane[w0];
    const uint32_t hi = (w0 + 1 < B.KB) ? plane[w0 + 1] : 0u;
    const unsigned long long both = ((unsigned long long)hi << 32) | (unsigned long long)lo;
    r = (uint32_t)(both >> s) & B.full;
  }
  return r;
}

AGZ_DEV Lines bits_load(const BitsCtx& B, const uint32_t* black, const uint32_t* white) {
  Lines L;
  L.b = bits_line(B, black);
  L.w = bits_line(B, white);
  return L;
}

// flat words of a line set; every lane receives all KB words (KW >= KB keeps them in registers)
template <int KW>
AGZ_DEV void bits_pack(const BitsCtx& B, uint32_t line, uint32_t (&out)[KW]) {
  const int bit = B.N * B.lane, w0 = bit >> 5, s = bit & 31;
  const uint32_t lo = line << s;
  const uint32_t hi = s ? (line >> (32 - s)) : 0u;
#pragma unroll
  for (int k = 0; k < KW; ++k) {
    out[k] = 0;
    if (k < B.KB) out[k] = simt::reduce_or((w0 == k ? lo : 0u) | (w0 + 1 == k ? hi : 0u));
  }
}

// three planes at once (a new node's black / white / legal planes): the per-word lane predicates are shared
template <int KW>
AGZ_DEV void bits_pack3(const BitsCtx& B, uint32_t l0, uint32_t l1, uint32_t l2, uint32_t (&o0)[KW], uint32_t (&o1)[KW], uint32_t (&o2)[KW]) {
  const int bit = B.N * B.lane, w0 = bit >> 5, s = bit & 31;
  const uint32_t lo0 = l0 << s, lo1 = l1 << s, lo2 = l2 << s;
  const uint32_t hi0 = s ? (l0 >> (32 - s)) : 0u, hi1 = s ? (l1 >> (32 - s)) : 0u, hi2 = s ? (l2 >> (32 - s)) : 0u;
#pragma unroll
  for (int k = 0; k < KW; ++k) {
    o0[k] = 0; o1[k] = 0; o2[k] = 0;
    if (k < B.KB) {
      const bool a = w0 == k, b = w0 + 1 == k;
      o0[k] = simt::reduce_or((a ? lo0 : 0u) | (b ? hi0 : 0u));
      o1[k] = simt::reduce_or((a ? lo1 : 0u) | (b ? hi1 : 0u));
      o2[k] = simt::reduce_or((a ? lo2 : 0u) | (b ? hi2 : 0u));
    }
  }
}

// board bytes (-1 W / 0 / +1 B, flat order) <-> lines; used by the position hooks only
AGZ_DEV Lines bits_from_bytes(const BitsCtx& B, const int8_t* board) {
  Lines L;
  L.b = 0;
  L.w = 0;
  if (B.lane < B.N) {
    for (int i = 0; i < B.N; ++i) {
      const int v = board[B.N * B.lane + i];
      L.b |= (v == 1 ? 1u : 0u) << i;
      L.w |= (v == -1 ? 1u : 0u) << i;
    }
  }
  return L;
}

AGZ_DEV void bits_to_bytes(const BitsCtx& B, const Lines& L, int8_t* board) {
  if (B.lane < B.N) {
    for (int i = 0; i < B.N; ++i) board[B.N * B.lane + i] = (int8_t)((int)((L.b >> i) & 1u) - (int)((L.w >> i) & 1u));
  }
}

// the points adjacent to a point of x (lanes >= N hold 0, so lane -1 = lane 31 and lane N read as empty lines)
AGZ_DEV uint32_t bits_nbr4(const BitsCtx& B, uint32_t x) {
  const uint32_t l = simt::shfl(x, B.lane - 1);
  const uint32_t r = simt::shfl(x, B.lane + 1);
  return ((x << 1) | (x >> 1) | l | r) & B.full;
}

// everything connected to `seed` through points of `mask`
AGZ_DEV uint32_t bits_flood(const BitsCtx& B, uint32_t seed, uint32_t mask) {
  uint32_t g = seed & mask;
  for (;;) {
    const uint32_t n1 = (g | bits_nbr4(B, g)) & mask;
    const uint32_t n2 = (n1 | bits_nbr4(B, n1)) & mask;
    const bool changed = n2 != g;
    g = n2;
    if (!simt::any(changed)) break;
  }
  return g;
}

// two independent floods in one loop (their shuffles overlap)
AGZ_DEV void bits_flood2(const BitsCtx& B, uint32_t& g0, uint32_t mask0, uint32_t& g1, uint32_t mask1) {
  g0 &= mask0;
  g1 &= mask1;
  for (;;) {
    const uint32_t a1 = (g0 | bits_nbr4(B, g0)) & mask0;
    const uint32_t c1 = (g1 | bits_nbr4(B, g1)) & mask1;
    const uint32_t a2 = (a1 | bits_nbr4(B, a1)) & mask0;
    const uint32_t c2 = (c1 | bits_nbr4(B, c1)) & mask1;
    const bool changed = (a2 != g0) || (c2 != g1);
    g0 = a2;
    g1 = c2;
    if (!simt::any(changed)) break;
  }
}

AGZ_DEV int bits_count(uint32_t x) { return simt::reduce_add(simt::popc(x)); }

// the lowest point of a non-empty set as a one-bit set (warp-uniform choice)
AGZ_DEV uint32_t bits_lowest(const BitsCtx& B, uint32_t x, unsigned nonzero_lanes) {
  const int dl = simt::ffs(nonzero_lanes) - 1;
  const uint32_t dv = simt::shfl(x, dl);
  return B.lane == dl ? (dv & (0u - dv)) : 0u;
}

// Play `color` at flat point c.  Returns 0, or 1 when check_legal is set and the point is occupied or the move is
// suicide (L is then unchanged).  ko_out = the point the opponent may not retake, or -1 (board.jl:473,487-491);
// ncap_out = stones captured.
AGZ_DEV int bits_play(const BitsCtx& B, Lines& L, int c, int color, bool check_legal, int& ko_out, int& ncap_out) {
  ko_out = -1;
  ncap_out = 0;
  int cj, ci;
  bits_point(B, c, cj, ci);
  const uint32_t cbit = B.lane == cj ? (1u << ci) : 0u;
  uint32_t mine = color == 1 ? L.b : L.w, opp = color == 1 ? L.w : L.b;
  if (check_legal && simt::any(((mine | opp) & cbit) != 0u)) return 1;
  const uint32_t nb = bits_nbr4(B, cbit);
  const bool koish = !simt::any((nb & ~opp) != 0u);  // is_koish: every neighbour holds the opponent's colour
  mine |= cbit;
  uint32_t empty = B.full & ~(mine | opp);
  int ncap = 0, cap_point = -1;
  if (simt::any((nb & opp) != 0u)) {
    const uint32_t alive = bits_flood(B, opp & bits_nbr4(B, empty), opp);
    const uint32_t dead = opp & ~alive;
    const unsigned dm = simt::ballot(dead != 0u);
    if (dm) {
      ncap = bits_count(dead);
      const int dl = simt::ffs(dm) - 1;
      const uint32_t dv = simt::shfl(dead, dl);
      cap_point = dl * B.N + simt::ffs(dv) - 1;
      opp &= ~dead;
      empty |= dead;
    }
  }
  if (check_legal) {  // suicide (board.jl:264-266): the played stone's group has no liberty after the captures
    const uint32_t grp = bits_flood(B, cbit, mine);
    if (!simt::any((bits_nbr4(B, grp) & empty) != 0u)) return 1;
  }
  L.b = color == 1 ? mine : opp;
  L.w = color == 1 ? opp : mine;
  ncap_out = ncap;
  if (ncap == 1 && koish) ko_out = cap_point;
  return 0;
}

// all_legal_moves for `to_play`: the legal points of this lane's line (pass is always legal and not part of the set)
AGZ_DEV uint32_t bits_legal(const BitsCtx& B, const Lines& L, int to_play, int ko) {
  const uint32_t S = to_play == 1 ? L.b : L.w, O = to_play == 1 ? L.w : L.b;
  const uint32_t E = B.full & ~(S | O);
  const uint32_t open = E & bits_nbr4(B, E);  // an empty neighbour: never suicide
  const uint32_t cand = E & ~open;
  uint32_t legal = open;
  if (simt::any(cand != 0u)) {
    // groups owning a liberty that is not a candidate: friendly ones make every adjacent candidate legal (they keep
    // that liberty), enemy ones cannot be captured by a candidate
    const uint32_t on = bits_nbr4(B, open);
    uint32_t f1 = S & on, o1 = O & on;
    bits_flood2(B, f1, S, o1, O);
    legal |= cand & bits_nbr4(B, f1);
    uint32_t srem = S & ~f1, orem = O & ~o1;  // groups whose liberties are all candidates (rare)
    for (;;) {
      const unsigned m = simt::ballot(srem != 0u);
      if (!m) break;
      const uint32_t g = bits_flood(B, bits_lowest(B, srem, m), S);
      const uint32_t libs = bits_nbr4(B, g) & E;
      if (bits_count(libs) >= 2) legal |= libs;  // a friendly group with a liberty besides the point
      srem &= ~g;
    }
    for (;;) {
      const unsigned m = simt::ballot(orem != 0u);
      if (!m) break;
      const uint32_t g = bits_flood(B, bits_lowest(B, orem, m), O);
      const uint32_t libs = bits_nbr4(B, g) & E;
      if (bits_count(libs) == 1) legal |= libs;  // an enemy group in atari: the point captures it
      orem &= ~g;
    }
  }
  if (ko >= 0) {
    const int kj = ko / B.N, ki = ko - kj * B.N;
    if (B.lane == kj) legal &= ~(1u << ki);
  }
  return legal;
}

// Tromp-Taylor area score from Black's view: Float32(#B - #W) - komi  (board.jl:511-533)
AGZ_DEV float bits_score(const BitsCtx& B, const Lines& L, float komi) {
  const uint32_t E = B.full & ~(L.b | L.w);
  uint32_t rb = E & bits_nbr4(B, L.b), rw = E & bits_nbr4(B, L.w);
  bits_flood2(B, rb, E, rw, E);
  const int nb = bits_count(L.b | (rb & ~rw));
  const int nw = bits_count(L.w | (rw & ~rb));
  return simt::fsub((float)(nb - nw), komi);
}

// Gomoku (src/game/gomoku/board.jl:97-129 has_game_ended): does the stone set x hold k in a row -- along a line, across lines, or
// on either diagonal?  Lane j holds line j (lanes >= N hold 0), so the three cross-line directions are k-1 shuffles.
AGZ_DEV bool bits_k_in_row(const BitsCtx& B, uint32_t x, int k) {
  uint32_t h = x, v = x, d1 = x, d2 = x;
  for (int t = 1; t < k; ++t) {
    const uint32_t up = simt::shfl(x, B.lane + t);
    h &= x >> t;
    v &= up;
    d1 &= up >> t;
    d2 &= up << t;
  }
  return simt::any((h | v | d1 | d2) != 0u);
}

}  // namespace agz
