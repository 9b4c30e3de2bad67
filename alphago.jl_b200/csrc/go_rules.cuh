// go_rules.cuh -- Go rules for one position, executed cooperatively by one warp.
//
// Replaces (reference paths): src/game/go/board.jl  play_move! :451-509, pass_move! :426-440,
// add_stone!/captures :205-269, is_koish :47-56, all_legal_moves :393-424, is_move_suicidal :354-374,
// score :511-533, LibertyTracker.liberty_cache :99-164.
//
// Design: instead of the reference's incremental Dict{Int,Group}-of-Sets liberty tracker (deep-copied
// for every new tree node), a position is two bitplanes in HBM.  When a child position is needed the
// warp expands the planes into a per-warp shared-memory board (1 byte / point), labels connected
// groups by min-label propagation with pointer jumping (lane l owns points l, l+32, ...), counts each
// group's distinct liberties with shared-memory atomics, applies captures / ko / suicide rules and packs
// the child's planes and legal-move mask back with warp ballots.  Results are identical to the
// liberty tracker's (pinned by the reference's test_go.jl cases through the C ABI).
#pragma once
#include "simt.h"

namespace agz {

struct RulesScratch {
  int8_t* bd;    // [NP] -1 W / 0 / +1 B
  int16_t* lab;  // [NP] group (or empty-region) label = smallest flat index in it, -1 outside the set
  int* cnt;      // [NP] per-label liberty count / border-colour flags
};

struct Board {
  int N, N2, KB;  // KB = ceil(N2 / 32) words per bitplane
  // per-lane neighbour masks of the points this lane owns (point k*32 + lane -> bits [4k, 4k+4):
  // 1 = has row-1, 2 = has row+1, 4 = has column-1, 8 = has column+1); saves a div/mod per point per sweep
  unsigned long long nbm;
};

AGZ_DEV void board_init_masks(Board& B) {
  const int lane = simt::lane();
  unsigned long long m = 0;
  for (int k = 0; k < B.KB; ++k) {
    const int p = k * 32 + lane;
    if (p < B.N2) {
      const int i = p % B.N, j = p / B.N;
      unsigned long long b = (i > 0 ? 1u : 0u) | (i < B.N - 1 ? 2u : 0u) | (j > 0 ? 4u : 0u) | (j < B.N - 1 ? 8u : 0u);
      m |= b << (4 * k);
    }
  }
  B.nbm = m;
}

AGZ_DEV size_t rules_scratch_bytes(int KB) { return (size_t)KB * 32 * (1 + 2 + 4); }

AGZ_DEV RulesScratch rules_scratch_at(char* smem, int KB) {
  RulesScratch s;
  int NP = KB * 32;
  s.cnt = reinterpret_cast<int*>(smem);
  s.lab = reinterpret_cast<int16_t*>(smem + (size_t)NP * 4);
  s.bd = reinterpret_cast<int8_t*>(smem + (size_t)NP * 6);
  return s;
}

// neighbours of flat point p = N*j + i: up/down move i (row), left/right move j (column)
#define AGZ_FOR_NEIGHBORS(B, p, q, BODY)            \
  {                                                 \
    int i__ = (p) % (B).N, j__ = (p) / (B).N;       \
    int q;                                          \
    if (i__ > 0) { q = (p)-1; BODY }                \
    if (i__ < (B).N - 1) { q = (p) + 1; BODY }      \
    if (j__ > 0) { q = (p) - (B).N; BODY }          \
    if (j__ < (B).N - 1) { q = (p) + (B).N; BODY }  \
  }

// same, for a point owned by this lane: k-th point of the lane, neighbour mask from B.nbm
#define AGZ_FOR_OWN_NEIGHBORS(B, k, p, q, BODY)               \
  {                                                           \
    const unsigned m__ = (unsigned)((B).nbm >> (4 * (k))) & 15u; \
    int q;                                                    \
    if (m__ & 1u) { q = (p)-1; BODY }                         \
    if (m__ & 2u) { q = (p) + 1; BODY }                       \
    if (m__ & 4u) { q = (p) - (B).N; BODY }                   \
    if (m__ & 8u) { q = (p) + (B).N; BODY }                   \
  }

AGZ_DEV void rules_load(const Board& B, RulesScratch& s, const uint32_t* black, const uint32_t* white) {
  const int lane = simt::lane();
  for (int k = 0; k < B.KB; ++k) {
    uint32_t b = black[k], w = white[k];  // warp-uniform addresses: one broadcast load each
    s.bd[k * 32 + lane] = (int8_t)((int)((b >> lane) & 1u) - (int)((w >> lane) & 1u));
  }
  simt::sync();
}

AGZ_DEV void rules_load_bytes(const Board& B, RulesScratch& s, const int8_t* board) {
  const int lane = simt::lane();
  for (int k = 0; k < B.KB; ++k) {
    int p = k * 32 + lane;
    s.bd[p] = p < B.N2 ? board[p] : (int8_t)0;
  }
  simt::sync();
}

// pack the scratch board into bitplanes; every lane receives all words (KW >= B.KB keeps them in registers)
template <int KW>
AGZ_DEV void rules_pack(const Board& B, const RulesScratch& s, uint32_t (&black)[KW], uint32_t (&white)[KW]) {
  const int lane = simt::lane();
#pragma unroll
  for (int k = 0; k < KW; ++k) {
    black[k] = 0;
    white[k] = 0;
    if (k < B.KB) {
      int v = s.bd[k * 32 + lane];
      black[k] = simt::ballot(v == 1);
      white[k] = simt::ballot(v == -1);
    }
  }
}

// mode 0: label stones (4-connected, same colour); mode 1: label empty regions
AGZ_DEV void rules_label(const Board& B, RulesScratch& s, int mode) {
  const int lane = simt::lane();
  for (int k = 0; k < B.KB; ++k) {
    int p = k * 32 + lane;
    int v = s.bd[p];
    bool in = p < B.N2 && (mode == 0 ? v != 0 : v == 0);
    s.lab[p] = (int16_t)(in ? p : -1);
  }
  simt::sync();
  for (;;) {
    bool changed = false;
    for (int k = 0; k < B.KB; ++k) {
      int p = k * 32 + lane;
      int m0 = s.lab[p];
      if (m0 >= 0) {
        int v = s.bd[p];
        int m = m0;
        AGZ_FOR_OWN_NEIGHBORS(B, k, p, q, {
          int lq = s.lab[q];
          if (lq >= 0 && s.bd[q] == v && lq < m) m = lq;
        })
        int mm = s.lab[m];  // pointer jumping
        if (mm >= 0 && mm < m) m = mm;
        if (m < m0) {
          s.lab[p] = (int16_t)m;
          changed = true;
        }
      }
    }
    simt::sync();
    if (!simt::any(changed)) break;
  }
}

// cnt[label] = number of distinct empty points adjacent to the group (labels from rules_label(mode 0))
AGZ_DEV void rules_count_liberties(const Board& B, RulesScratch& s) {
  const int lane = simt::lane();
  for (int k = 0; k < B.KB; ++k) s.cnt[k * 32 + lane] = 0;
  simt::sync();
  for (int k = 0; k < B.KB; ++k) {
    int p = k * 32 + lane;
    if (p < B.N2 && s.bd[p] == 0) {
      int r[4];
      int nr = 0;
      AGZ_FOR_OWN_NEIGHBORS(B, k, p, q, {
        if (s.bd[q] != 0) {
          int l = s.lab[q];
          bool dup = false;
          for (int t = 0; t < nr; ++t) dup = dup || (r[t] == l);
          if (!dup) r[nr++] = l;
        }
      })
      for (int t = 0; t < nr; ++t) simt::atomic_add(&s.cnt[r[t]], 1);
    }
  }
  simt::sync();
}

// Play `color` at point c on the scratch board (labels + liberty counts are left valid for the result).
// Returns 0, or 1 when check_legal is set and the move is on a stone or suicidal (scratch is then garbage).
// ko_out = the point the opponent may not retake, or -1 (board.jl:473,487-491); ncap_out = stones captured.
AGZ_DEV int rules_play(const Board& B, RulesScratch& s, int c, int color, bool check_legal, int& ko_out, int& ncap_out) {
  const int lane = simt::lane();
  ko_out = -1;
  ncap_out = 0;
  if (check_legal && s.bd[c] != 0) return 1;
  // is_koish on the board before the move: every neighbour of c holds the opponent's colour
  bool koish = true;
  AGZ_FOR_NEIGHBORS(B, c, q, { koish = koish && (s.bd[q] == -color); })
  simt::sync();
  if (lane == 0) s.bd[c] = (int8_t)color;
  simt::sync();
  rules_label(B, s, 0);
  rules_count_liberties(B, s);
  // opponent groups next to c that are left without liberties are captured
  int cr[4];
  int ncr = 0;
  AGZ_FOR_NEIGHBORS(B, c, q, {
    if (s.bd[q] == -color) {
      int l = s.lab[q];
      if (s.cnt[l] == 0) {
        bool dup = false;
        for (int t = 0; t < ncr; ++t) dup = dup || (cr[t] == l);
        if (!dup) cr[ncr++] = l;
      }
    }
  })
  int ncap = 0, cap_point = -1;
  if (ncr > 0) {  // warp-uniform
    simt::sync();
    for (int k = 0; k < B.KB; ++k) {
      int p = k * 32 + lane;
      bool dead = false;
      if (s.bd[p] == -color) {
        int l = s.lab[p];
        for (int t = 0; t < ncr; ++t) dead = dead || (l == cr[t]);
      }
      unsigned m = simt::ballot(dead);
      if (dead) {
        s.bd[p] = 0;
        s.lab[p] = -1;
      }
      if (m) {
        ncap += simt::popc(m);
        cap_point = k * 32 + simt::ffs(m) - 1;
      }
    }
    simt::sync();
    rules_count_liberties(B, s);
  }
  if (check_legal && s.cnt[s.lab[c]] == 0) return 1;  // suicide (board.jl:264-266)
  ncap_out = ncap;
  if (ncap == 1 && koish) ko_out = cap_point;
  return 0;
}

// all_legal_moves for the player `to_play` on the scratch board (labels + liberty counts valid).
// legal[k] receives the bit mask of points k*32..k*32+31; pass is always legal and not part of the mask.
template <int KW>
AGZ_DEV void rules_legal_mask(const Board& B, const RulesScratch& s, int to_play, int ko, uint32_t (&legal)[KW]) {
  const int lane = simt::lane();
#pragma unroll
  for (int k = 0; k < KW; ++k) {
    legal[k] = 0;
    if (k >= B.KB) continue;
    int p = k * 32 + lane;
    bool ok = false;
    if (p < B.N2 && s.bd[p] == 0 && p != ko) {
      AGZ_FOR_OWN_NEIGHBORS(B, k, p, q, {
        int v = s.bd[q];
        if (v == 0) {
          ok = true;
        } else {
          int libs = s.cnt[s.lab[q]];
          if (v == to_play ? libs >= 2 : libs == 1) ok = true;
        }
      })
    }
    legal[k] = simt::ballot(ok);
  }
}

// Tromp-Taylor area score from Black's view: Float32(#B - #W) - komi  (board.jl:511-533)
AGZ_DEV float rules_score(const Board& B, RulesScratch& s, float komi) {
  const int lane = simt::lane();
  rules_label(B, s, 1);
  for (int k = 0; k < B.KB; ++k) s.cnt[k * 32 + lane] = 0;
  simt::sync();
  for (int k = 0; k < B.KB; ++k) {
    int p = k * 32 + lane;
    if (p < B.N2 && s.bd[p] == 0) {
      int f = 0;
      AGZ_FOR_OWN_NEIGHBORS(B, k, p, q, {
        int v = s.bd[q];
        if (v == 1) f |= 1;
        if (v == -1) f |= 2;
      })
      if (f) simt::atomic_or(&s.cnt[s.lab[p]], f);
    }
  }
  simt::sync();
  int nb = 0, nw = 0;
  for (int k = 0; k < B.KB; ++k) {
    int p = k * 32 + lane;
    bool b = false, w = false;
    if (p < B.N2) {
      int v = s.bd[p];
      if (v == 1) b = true;
      else if (v == -1) w = true;
      else {
        int f = s.cnt[s.lab[p]];
        b = (f == 1);
        w = (f == 2);
      }
    }
    nb += simt::popc(simt::ballot(b));
    nw += simt::popc(simt::ballot(w));
  }
  return simt::fsub((float)(nb - nw), komi);
}

}  // namespace agz
