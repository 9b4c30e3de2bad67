// rng.cuh -- counter-based RNG and deterministic fp64 math of the engine.
// The spec (sites, counters, samplers) is written down once in oracle/rng.py; this is its device twin.
// Replaces the reference's five draws from Julia's global RNG: src/selfplay.jl:9, src/mcts.jl:133,235,
// src/mcts_play.jl:61,66.
#pragma once
#include "simt.h"

namespace agz {

enum { SITE_RESIGN = 1, SITE_SELECT = 2, SITE_NOISE = 3, SITE_PICK_MAX = 4, SITE_PICK_SOFT = 5 };

struct U4 {
  uint32_t x, y, z, w;
};

AGZ_DEV U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = simt::mulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = simt::mulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

AGZ_DEV U4 rng_draw(uint64_t seed, uint32_t game_id, int site, uint32_t move_no, uint32_t i, uint32_t j) {
  return philox4x32_10(game_id, ((uint32_t)site << 28) | (move_no & 0x0FFFFFFFu), i, j, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// [0,1) with 53 random bits
AGZ_DEV double u53(uint32_t a, uint32_t b) {
  uint64_t k = ((uint64_t)(a >> 5) << 26) + (uint64_t)(b >> 6);
  return (double)k * 1.1102230246251565e-16;  // 2^-53, exact
}
// (0,1): (2k+1) * 2^-53 with 52 random bits
AGZ_DEV double u52c(uint32_t a, uint32_t b) {
  uint64_t k = ((uint64_t)(a >> 6) << 26) + (uint64_t)(b >> 6);
  return (double)(2 * k + 1) * 1.1102230246251565e-16;
}

// ---- deterministic log / exp: only + - * / floor, fixed order, no FMA (see oracle/rng.py) ----------
AGZ_DEV double det_log(double x) {
  const double C[12] = {1.0, 1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0,
                        1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0, 1.0 / 23.0};
  long long b = simt::dbits(x);
  int e = (int)((b >> 52) & 0x7ff) - 1022;                      // frexp: x = m * 2^e, m in [0.5, 1)
  double m = simt::bitsd((b & 0x800fffffffffffffLL) | 0x3fe0000000000000LL);
  if (m < 0.7071067811865476) { m = simt::dmul(m, 2.0); e -= 1; }
  double s = simt::ddiv(simt::dsub(m, 1.0), simt::dadd(m, 1.0));
  double z = simt::dmul(s, s);
  double p = C[11];
#pragma unroll
  for (int k = 10; k >= 0; --k) p = simt::dadd(simt::dmul(p, z), C[k]);
  double lm = simt::dmul(simt::dmul(2.0, s), p);
  return simt::dadd(simt::dmul((double)e, 0.6931471805599453), lm);
}

AGZ_DEV double det_exp(double x) {
  const double C[14] = {1.0, 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0,
                        1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0,
                        1.0 / 6227020800.0};
  if (x < -700.0) return 0.0;
  if (x > 700.0) return simt::bitsd(0x7ff0000000000000LL);
  double fk = simt::dfloor(simt::dadd(simt::dmul(x, 1.4426950408889634), 0.5));
  int k = (int)fk;
  double r = simt::dsub(simt::dsub(x, simt::dmul(fk, 6.93147180369123816490e-01)), simt::dmul(fk, 1.90821492927058770002e-10));
  double p = C[13];
#pragma unroll
  for (int n = 12; n >= 0; --n) p = simt::dadd(simt::dmul(p, r), C[n]);
  if (k < -1000) return 0.0;
  return simt::dmul(p, simt::bitsd((long long)(k + 1023) << 52));
}

AGZ_DEV double det_pow(double x, double y) {
  if (x == 0.0) return 0.0;
  return det_exp(simt::dmul(y, det_log(x)));
}

// Gamma(alpha, 1), 0 < alpha < 1: Ahrens & Dieter (1974) GS.  Counter = (action, (noise_call << 16) | attempt).
// One attempt: x = the candidate, returns whether it is accepted (b = 1 + alpha / e).
AGZ_DEV bool gamma_small_attempt(double alpha, double b, uint64_t seed, uint32_t game_id, uint32_t move_no, uint32_t a, uint32_t noise_call,
                                 uint32_t t, double& x) {
  U4 r = rng_draw(seed, game_id, SITE_NOISE, move_no, a, (noise_call << 16) | t);
  double u1 = u52c(r.x, r.y), u2 = u52c(r.z, r.w);
  double p = simt::dmul(b, u1);
  if (p <= 1.0) {
    x = det_exp(simt::ddiv(det_log(p), alpha));
    return u2 <= det_exp(-x);
  }
  x = -det_log(simt::ddiv(simt::dsub(b, p), alpha));
  return u2 <= det_exp(simt::dmul(simt::dsub(alpha, 1.0), det_log(x)));
}
// the sampler: at most 64 attempts (the last candidate is returned if none was accepted)
AGZ_DEV double gamma_small(double alpha, uint64_t seed, uint32_t game_id, uint32_t move_no, uint32_t a, uint32_t noise_call) {
  const double b = simt::dadd(1.0, simt::ddiv(alpha, 2.718281828459045));
  double x = 0.0;
  for (uint32_t t = 0; t < 64; ++t)
    if (gamma_small_attempt(alpha, b, seed, game_id, move_no, a, noise_call, t, x)) return x;
  return x;
}

// sum over the warp in the spec's fixed order (lane-local ascending, then xor butterfly 16,8,4,2,1)
AGZ_DEV double butterfly_sum(double lane_acc) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) lane_acc = simt::dadd(lane_acc, simt::shfl_xor(lane_acc, off));
  return lane_acc;
}

}  // namespace agz
