"""1:1 port of the reference's test/test_mcts.jl and test/test_features.jl onto the CPU oracle."""
import numpy as np
import pytest

from oracle import go, features
from oracle import mcts as M
from oracle.go import BLACK, WHITE, GoPosition, PlayerMove
from refboards import load_board, ALMOST_DONE_BOARD, EMPTY_ROW9 as EMPTY_ROW

env = go.GoEnv(9)
A = env.action_space
f32 = np.float32
kgs = lambda s: go.from_kgs(s, env)


def send_two_return_one():                               # test_mcts.jl:34-43
    return GoPosition(env, board=load_board(ALMOST_DONE_BOARD, env), n=75, komi=0.5, caps=(0, 0), ko=None,
                      recent=[PlayerMove(BLACK, (0, 1)), PlayerMove(WHITE, (0, 8)), PlayerMove(BLACK, (1, 0))],
                      to_play=WHITE)


def test_action_flipping():                              # :45-59
    rs = np.random.RandomState(1)
    probs = (0.02 * np.ones(A) + rs.rand(A) * 0.001).astype(f32)
    black_root = M.MCTSNode(GoPosition(env))
    white_root = M.MCTSNode(GoPosition(env, to_play=WHITE))
    M.incorporate_results(M.select_leaf(black_root), probs, 0, black_root)
    M.incorporate_results(M.select_leaf(white_root), probs, 0, white_root)
    bl = M.select_leaf(black_root)
    wl = M.select_leaf(white_root)
    assert bl.fmove == wl.fmove
    # (the reference compares the two score vectors; after the selections above they are equal)
    assert (M.child_action_score(black_root) == M.child_action_score(white_root)).all()


def test_select_leaf():                                  # :61-70
    flattened = go.to_flat(kgs("D9"), env)
    probs = (0.02 * np.ones(A)).astype(f32)
    probs[flattened] = 0.4
    root = M.MCTSNode(send_two_return_one())
    M.incorporate_results(M.select_leaf(root), probs, 0, root)
    assert root.position.to_play == WHITE
    assert M.select_leaf(root) is root.children[flattened]


def test_backup_incorporate_results():                   # :72-114
    probs = (0.02 * np.ones(A)).astype(f32)
    root = M.MCTSNode(send_two_return_one())
    M.incorporate_results(M.select_leaf(root), probs, 0, root)
    leaf = M.select_leaf(root)
    M.incorporate_results(leaf, probs, -1, root)
    assert root.N == 2
    assert root.Q == pytest.approx(-1 / 3, rel=1e-6)
    assert root.child_N[leaf.fmove] == 1
    assert leaf.N == 1
    assert M.child_Q(root)[leaf.fmove] == -0.5
    assert leaf.Q == pytest.approx(-0.5)
    assert root.position.to_play == WHITE
    leaf2 = M.select_leaf(root)
    M.incorporate_results(leaf2, probs, -0.2, root)
    assert root.N == 3
    assert root.Q == pytest.approx(-0.3, rel=1e-6)
    assert leaf.N == 2
    assert leaf2.N == 1
    assert leaf2.parent is leaf      # the reference's stated assumption (root -> leaf -> leaf2)
    assert leaf.Q == pytest.approx(M.child_Q(root)[leaf.fmove])
    assert leaf.Q == pytest.approx(-0.4, rel=1e-6)
    assert M.child_Q(leaf)[leaf2.fmove] == pytest.approx(-0.6, rel=1e-6)
    assert leaf2.Q == pytest.approx(-0.6, rel=1e-6)


def test_do_not_explore_past_finish():                   # :116-127
    probs = (0.02 * np.ones(A)).astype(f32)
    root = M.MCTSNode(GoPosition(env))
    M.incorporate_results(M.select_leaf(root), probs, 0, root)
    first_pass = M.maybe_add_child(root, go.to_flat(None, env))
    M.incorporate_results(first_pass, probs, 0, root)
    second_pass = M.maybe_add_child(first_pass, go.to_flat(None, env))
    with pytest.raises(AssertionError):
        M.incorporate_results(second_pass, probs, 0, root)
    node_to_explore = M.select_leaf(second_pass)
    assert node_to_explore is second_pass


def test_add_child():                                    # :129-135 (reference fmove 17 -> 0-based 16)
    root = M.MCTSNode(GoPosition(env))
    child = M.maybe_add_child(root, 16)
    assert 16 in root.children
    assert child.parent is root
    assert child.fmove == 16


def test_add_child_idempotency():                        # :137-144
    root = M.MCTSNode(GoPosition(env))
    child = M.maybe_add_child(root, 16)
    current = dict(root.children)
    child2 = M.maybe_add_child(root, 16)
    assert child is child2
    assert current == root.children


def test_never_select_illegal_moves():                   # :146-167 (reference index 2 -> 0-based 1)
    probs = (0.02 * np.ones(A)).astype(f32)
    probs[1] = 0.99
    root = M.MCTSNode(send_two_return_one())
    M.incorporate_results(root, probs, 0, root)
    root.N = 10000
    root.child_N[go.all_legal_moves(root.position).astype(bool)] = 10000
    leaf = M.select_leaf(root)
    assert leaf.fmove != 1
    for _ in range(10):
        M.inject_noise(root)
        leaf = M.select_leaf(root)
        assert leaf.fmove != 1


def test_dont_pick_unexpanded_child():                   # :169-183 (reference index 18 -> 17)
    probs = (0.02 * np.ones(A)).astype(f32)
    probs[17] = 0.999
    root = M.MCTSNode(GoPosition(env))
    M.incorporate_results(root, probs, 0, root)
    leaf1 = M.select_leaf(root)
    assert leaf1.fmove == 17
    M.add_virtual_loss(leaf1, root)
    leaf2 = M.select_leaf(root)
    assert leaf1 is leaf2


# ------------------------------------------------------------ test_features.jl
def test_stone_features():                               # test_features.jl:39-79
    pos = GoPosition(env)
    for c in ((0, 0), (0, 1), (0, 2), (0, 3), (1, 1)):
        go.play_move(pos, c, mutate=True)
    f = features.stone_features(pos)
    assert pos.to_play == WHITE
    assert f.shape == (9, 9, 16)
    lb = lambda two_rows: load_board(two_rows + EMPTY_ROW * 7, env)
    assert (f[:, :, 0] == lb("...X.....\n.........\n")).all()
    assert (f[:, :, 1] == lb("X.X......\n.X.......\n")).all()
    assert (f[:, :, 2] == lb(".X.X.....\n.........\n")).all()
    assert (f[:, :, 3] == lb("X.X......\n.........\n")).all()
    assert (f[:, :, 4] == lb(".X.......\n.........\n")).all()
    assert (f[:, :, 5] == lb("X.X......\n.........\n")).all()
    for i in range(10, 16):
        assert (f[:, :, i] == 0).all()
    full = features.get_feats(pos)
    assert full.shape == (9, 9, 17) and (full[:, :, 16] == -1).all()


def test_replay_sample_spec_is_a_permutation():
    """oracle/replay.py: the keyed Feistel permutation behind agz_replay_sample visits every index exactly once, for ring sizes
    around powers of two and for the reference's memory_size, and different seeds give different draws."""
    from oracle import replay as orp
    for n in (1, 2, 3, 4, 5, 16, 17, 255, 256, 257, 1000):
        idx = orp.sample_indices(n, 10 ** 9, n, seed=n)
        assert sorted(idx) == list(range(n)), n
    a = orp.sample_indices(700000, 500000, 4096, seed=1)
    b = orp.sample_indices(700000, 500000, 4096, seed=2)
    assert len(set(a)) == 4096 and min(a) >= 200000 and max(a) < 700000 and a != b
    # roughly uniform: mean of 4096 draws from [200000, 700000) within 3 sigma of the centre
    mean = sum(a) / len(a)
    assert abs(mean - 450000) < 3 * (500000 / 12 ** 0.5) / 64


def test_det_pow_agrees_with_libm():
    """children_as_pi uses the deterministic pow of oracle/rng.py (shared with the engine) instead of libm's: it must agree with
    math.pow to a few ulp over the range of visit counts (oracle/__init__.py, pinning status)."""
    import math
    from oracle import rng
    worst = 0.0
    for n in list(range(1, 2000)) + [5000, 12345, 65536, 10 ** 6]:
        a, b = rng.det_pow(float(n), 0.98), math.pow(float(n), 0.98)
        worst = max(worst, abs(a - b) / b)
    assert worst < 1e-14, worst


def test_reciprocal_quotients_equal_true_division():
    """The engine's PUCT score (tree.cuh select_leaf) replaces w / d (Float32) and x / d (Float64), d = 1 + N(child) a small
    integer, by products with r = RN(1/d): RN32(w * r) and Markstein's fma(fma(-q0, d, x), r, q0) with q0 = RN(x * r).  Exact
    rational arithmetic check of both against correctly rounded division (the device twin, 10^8 samples, is a gpu test)."""
    from fractions import Fraction as F
    import struct
    rs = np.random.RandomState(3)

    def rn64(fr):
        return fr.numerator / fr.denominator            # int / int true division is correctly rounded in CPython

    def rn32(fr):
        # round an exact rational to fp32: via fp64 is NOT allowed here (double rounding), so search the two neighbours
        x = np.float32(fr.numerator / fr.denominator)
        cands = [x, np.nextafter(x, np.float32(np.inf)), np.nextafter(x, np.float32(-np.inf))]
        best = min(cands, key=lambda c: (abs(F(float(c)) - fr), int(struct.unpack("<I", struct.pack("<f", c))[0]) & 1))
        return best

    divisors = list(range(1, 600)) + [int(d) for d in rs.randint(600, 1 << 20, 400)]
    for d in divisors:
        r = rn64(F(1, d))
        for _ in range(6):
            w = np.float32(rs.standard_normal() * 10.0 ** rs.randint(-6, 3))
            q_fast = np.float32(float(w) * r)           # fp64 product, then one rounding to fp32
            assert q_fast == rn32(F(float(w)) / d), (d, w)
            x = float(np.float64(0.96) * np.float64(np.sqrt(np.float32(1 + rs.randint(0, 5000)))) * np.float64(np.float32(rs.rand() ** 4)))
            q0 = x * r
            rem = rn64(F(x) - F(q0) * d)                # fma: exact product and sum, one rounding
            u = rn64(F(q0) + F(rem) * F(r))
            assert u == rn64(F(x) / d), (d, x)
