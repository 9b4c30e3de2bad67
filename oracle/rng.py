"""Counter-based RNG + deterministic fp64 math shared by the oracle and the engine.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws from Julia's global RNG at five sites
  #1 src/selfplay.jl:9        resign disabled in 5 % of games
  #2 src/mcts.jl:133          uniform choice among exactly-tied best children
  #3 src/mcts.jl:235          Dirichlet noise on the root prior
  #4 src/mcts_play.jl:61      uniform choice among most-visited children
  #5 src/mcts_play.jl:66      rand() for the soft pick
Julia's stream cannot be reproduced, so the spec below replaces it.  Every draw
is a pure function of (seed, game_id, site, move_no, i, j): results do not
depend on batching, warp scheduling or on how games are sharded over GPUs.

  key     = (seed & 0xffffffff, seed >> 32)
  counter = (game_id, (site << 28) | move_no, i, j)        all u32
  (r0, r1, r2, r3) = philox4x32_10(counter, key)

  site 1: i = j = 0.                 u = u53(r0, r1);   disabled iff u < 0.05
  site 2: i = select_leaf call index within the current root (reset on
          initialize_game!/play_move!), j = depth of the node being scored
          (root = 0).                pick = mulhi(r0, k) among k ties, ascending move
  site 3: i = action index, j = (noise_call_index << 16) | attempt.
          (U1, U2) = (u52c(r0, r1), u52c(r2, r3))  Ahrens-Dieter GS gamma sampler
  site 4: i = j = 0.                 pick = mulhi(r0, k)
  site 5: i = j = 0.                 u = u53(r0, r1)

  u53(a, b)  = ((a >> 5) * 2^26 + (b >> 6)) * 2^-53           in [0, 1)
  u52c(a, b) = (2 * ((a >> 6) * 2^26 + (b >> 6)) + 1) * 2^-53  in (0, 1)

All fp64 transcendental math goes through det_log / det_exp below, which use
only correctly-rounded IEEE operations (+ - * /, no FMA) in a fixed order, so
the CUDA side (with __dadd_rn/__dmul_rn/__ddiv_rn) reproduces them bit for bit.
"""
import math

M32 = 0xFFFFFFFF
PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85

SITE_RESIGN = 1
SITE_SELECT = 2
SITE_NOISE = 3
SITE_PICK_MAX = 4
SITE_PICK_SOFT = 5


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> 32, p0 & M32
        hi1, lo1 = p1 >> 32, p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0 = (k0 + PHILOX_W0) & M32
        k1 = (k1 + PHILOX_W1) & M32
    return c0, c1, c2, c3


def draw(seed, game_id, site, move_no, i=0, j=0):
    key = (seed & M32, (seed >> 32) & M32)
    ctr = (game_id & M32, ((site << 28) | (move_no & 0x0FFFFFFF)) & M32, i & M32, j & M32)
    return philox4x32_10(ctr, key)


def u53(a, b):
    return float(((a >> 5) << 26) + (b >> 6)) * (2.0 ** -53)


def u52c(a, b):
    return float(2 * (((a >> 6) << 26) + (b >> 6)) + 1) * (2.0 ** -53)


def mulhi(r, k):
    return (r * k) >> 32


# ---------------------------------------------------------------- det math
LN2_HI = 6.93147180369123816490e-01   # fdlibm split of ln 2 (hi has 21 trailing zero bits)
LN2_LO = 1.90821492927058770002e-10
LN2 = 0.6931471805599453
LOG2E = 1.4426950408889634
SQRT_HALF = 0.7071067811865476

# 1/(2k+1), k = 0..11
_LOG_COEF = [1.0 / (2 * k + 1) for k in range(12)]
# 1/n!, n = 0..13
_EXP_COEF = [1.0 / math.factorial(n) for n in range(14)]


def det_log(x):
    """Natural log of a positive normal double using only + - * / in fixed order."""
    m, e = math.frexp(x)            # x = m * 2^e, m in [0.5, 1): exact
    if m < SQRT_HALF:
        m = m * 2.0                 # exact
        e -= 1
    s = (m - 1.0) / (m + 1.0)
    z = s * s
    p = _LOG_COEF[11]
    for k in range(10, -1, -1):
        p = p * z + _LOG_COEF[k]
    lm = (2.0 * s) * p
    return float(e) * LN2 + lm


def det_exp(x):
    """exp(x) using only + - * / floor in fixed order.  < 2^-1000 flushes to 0."""
    if x < -700.0:
        return 0.0
    if x > 700.0:
        return math.inf
    k = math.floor(x * LOG2E + 0.5)
    fk = float(k)
    r = (x - fk * LN2_HI) - fk * LN2_LO
    p = _EXP_COEF[13]
    for n in range(12, -1, -1):
        p = p * r + _EXP_COEF[n]
    if k < -1000:
        return 0.0
    return p * math.ldexp(1.0, k)   # exact power-of-two scaling (normal range)


def det_pow(x, y):
    """x^y for x > 0 (x == 0 -> 0 for y > 0)."""
    if x == 0.0:
        return 0.0
    return det_exp(y * det_log(x))


E_CONST = 2.718281828459045
MAX_GAMMA_ATTEMPTS = 64


def gamma_small(alpha, seed, game_id, move_no, a, noise_call):
    """Gamma(alpha, 1) for 0 < alpha < 1: Ahrens & Dieter (1974) algorithm GS."""
    b = 1.0 + alpha / E_CONST
    x = 0.0
    for t in range(MAX_GAMMA_ATTEMPTS):
        r0, r1, r2, r3 = draw(seed, game_id, SITE_NOISE, move_no, a, (noise_call << 16) | t)
        u1 = u52c(r0, r1)
        u2 = u52c(r2, r3)
        p = b * u1
        if p <= 1.0:
            x = det_exp(det_log(p) / alpha)
            if u2 <= det_exp(-x):
                return x
        else:
            x = -det_log((b - p) / alpha)
            if u2 <= det_exp((alpha - 1.0) * det_log(x)):
                return x
    return x


def butterfly_sum32(vals):
    """Sum of len(vals) doubles in the engine's warp order: lane l first adds its own
    entries l, l+32, l+64, ... ascending, then a 5-step xor butterfly (16, 8, 4, 2, 1)."""
    lanes = [0.0] * 32
    for l in range(32):
        acc = 0.0
        for idx in range(l, len(vals), 32):
            acc = acc + vals[idx]
        lanes[l] = acc
    for off in (16, 8, 4, 2, 1):
        lanes = [lanes[l] + lanes[l ^ off] for l in range(32)]
    return lanes[0]


def dirichlet(alpha, n, seed, game_id, move_no, noise_call):
    g = [gamma_small(alpha, seed, game_id, move_no, a, noise_call) for a in range(n)]
    s = butterfly_sum32(g)
    return [x / s for x in g]
