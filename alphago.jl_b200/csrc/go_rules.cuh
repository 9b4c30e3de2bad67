// go_rules.cuh -- connected-group labels and liberty counts for one position in per-warp shared memory.
//
// Stands for LibertyTracker.liberty_cache (src/game/go/board.jl:99-164), which the reference's tests read
// (test/test_go.jl:74-139) and which only the liberty-cache hook of the C ABI (agz_pos_liberties) needs.  The hot path does
// not use this file: play / capture / ko / legality / score run on bitboard lines in registers (go_bits.cuh).
//
// Design: 1 byte per point; groups are labelled by min-label propagation with pointer jumping in Jacobi sweeps (lane l owns
// points l, l+32, ...; reads of a sweep are separated from its writes by a warp barrier), liberties are counted per label
// with shared-memory atomics.
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

AGZ_DEV void rules_load_bytes(const Board& B, RulesScratch& s, const int8_t* board) {
  const int lane = simt::lane();
  for (int k = 0; k < B.KB; ++k) {
    int p = k * 32 + lane;
    s.bd[p] = p < B.N2 ? board[p] : (int8_t)0;
  }
  simt::sync();
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
  for (;;) {  // Jacobi sweeps: every lane reads the labels of the previous sweep, then all lanes write (no intra-warp races)
    bool changed = false;
    int nl[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      nl[k] = -1;
      if (k < B.KB) {
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
            nl[k] = m;
            changed = true;
          }
        }
      }
    }
    simt::sync();
#pragma unroll
    for (int k = 0; k < 12; ++k)
      if (k < B.KB && nl[k] >= 0) s.lab[k * 32 + lane] = (int16_t)nl[k];
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

}  // namespace agz
