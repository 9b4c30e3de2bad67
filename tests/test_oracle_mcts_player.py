"""1:1 port of the reference's test/test_mcts_player.jl onto the CPU oracle."""
import numpy as np
import pytest

from oracle import go
from oracle import mcts as M
from oracle import mcts_play as P
from oracle.go import BLACK, WHITE, GoPosition, PlayerMove
from refboards import load_board, ALMOST_DONE_BOARD, TT_FTW_BOARD, assert_no_pending_vlosses

env = go.GoEnv(9)
A = env.action_space
f32 = np.float32
kgs = lambda s: go.from_kgs(s, env)


class DummyNet:                                          # test_mcts_player.jl:10-32
    def __init__(self, env, fake_priors=None, fake_value=0):
        self.p = (np.ones(env.action_space) / env.action_space if fake_priors is None else np.asarray(fake_priors)).astype(f32)
        self.v = f32(fake_value)

    def __call__(self, positions):
        if positions is None or len(positions) == 0:
            raise ValueError("No positions passed!")
        n = len(positions)
        return np.repeat(self.p[:, None], n, axis=1), np.repeat(self.v, n)


def send_two_return_one():                               # :48-57
    return GoPosition(env, board=load_board(ALMOST_DONE_BOARD, env), n=70, komi=2.5, caps=(1, 4), ko=None,
                      recent=[PlayerMove(BLACK, (0, 1)), PlayerMove(WHITE, (0, 8))], to_play=BLACK)


def initialize_basic_player():                           # :59-66
    player = P.MCTSPlayer(env, DummyNet(env))
    P.initialize_game(player)
    first = M.select_leaf(player.root)
    p, v = player.network([player.root.position])
    M.incorporate_results(first, p[:, 0], v[0], player.root)
    return player


def initialize_almost_done_player():                     # :68-77
    probs = np.ones(A) * 0.001
    probs[2:5] = 0.2
    probs[-1] = 0.2
    player = P.MCTSPlayer(env, DummyNet(env, fake_priors=probs))
    P.initialize_game(player, send_two_return_one())
    return player


def test_inject_noise():                                 # :93-109
    player = initialize_basic_player()
    s = player.root.child_prior.sum()
    assert s == pytest.approx(1, rel=1e-5)
    u = M.child_U(player.root)
    assert (u == u[0]).all()
    M.inject_noise(player.root)
    assert player.root.child_prior.sum() == pytest.approx(s, rel=1e-5)
    assert player.root.child_prior.max() > 3 / A


def test_pick_moves():                                   # :111-137
    player = initialize_basic_player()
    root = player.root
    root.child_N[go.to_flat((2, 0), env)] = 10
    root.child_N[go.to_flat((1, 0), env)] = 5
    root.child_N[go.to_flat((3, 0), env)] = 1
    root.position.n = A
    assert root.position.n > player.tau_threshold
    assert P.pick_move(player) == (2, 0)
    root.position.n = 3
    assert root.position.n <= player.tau_threshold
    # soft pick (left TODO in the reference): must land on one of the visited moves
    assert P.pick_move(player) in ((2, 0), (1, 0), (3, 0))


def test_dont_pass_if_losing():                          # :139-165
    player = initialize_almost_done_player()
    assert go.score(player.root.position) == -0.5
    for _ in range(20):
        P.tree_search(player)
    flattened = go.to_flat(kgs("D9"), env)
    assert int(np.argmax(player.root.child_N)) == flattened
    assert player.root.children[flattened].Q > 0
    assert player.root.N >= 20
    assert M.child_Q(player.root)[-1] < 0
    assert_no_pending_vlosses(player.root)


def test_parallel_tree_search():                         # :167-192
    player = initialize_almost_done_player()
    assert go.score(player.root.position) == -0.5
    P.tree_search(player, 1)
    for _ in range(6):
        P.tree_search(player, 10)
    flattened = go.to_flat(kgs("D9"), env)
    best = np.flatnonzero(player.root.child_N == player.root.child_N.max())
    assert flattened in best
    assert player.root.children[flattened].Q > 0
    assert player.root.N >= 20
    assert_no_pending_vlosses(player.root)


def test_ridiculously_parallel_tree_search():            # :194-202
    player = initialize_almost_done_player()
    for _ in range(10):
        P.tree_search(player, 50)
    assert_no_pending_vlosses(player.root)


def test_long_game_tree_search():                        # :204-225
    rules = M.MCTSRules(env)
    player = P.MCTSPlayer(env, DummyNet(env))
    endgame = GoPosition(env, board=load_board(TT_FTW_BOARD, env), n=rules.max_game_length - 2, komi=2.5, ko=None,
                         recent=[PlayerMove(BLACK, (0, 1)), PlayerMove(WHITE, (0, 8))], to_play=BLACK)
    P.initialize_game(player, endgame)
    for _ in range(10):
        P.tree_search(player, 8)
    assert_no_pending_vlosses(player.root)
    assert player.root.Q > 0


def test_cold_start_parallel_tree_search():              # :227-240
    player = P.MCTSPlayer(env, DummyNet(env, fake_value=0.17))
    P.initialize_game(player)
    assert player.root.N == 0
    assert not player.root.is_expanded
    P.tree_search(player, 4)
    assert_no_pending_vlosses(player.root)
    assert player.root.N == 1          # the reference forgot @test on this line (:237); it does hold
    assert player.root.Q == pytest.approx(0.085, rel=1e-6)


def test_tree_search_failsafe():                         # :242-252
    probs = np.ones(A) * 0.001
    probs[-1] = 1
    player = P.MCTSPlayer(env, DummyNet(env, fake_priors=probs))
    P.initialize_game(player, go.pass_move(GoPosition(env)))
    P.tree_search(player, 1)
    assert_no_pending_vlosses(player.root)


def test_only_check_game_end_once():                     # :254-283
    pos = go.pass_move(go.play_move(go.play_move(go.play_move(GoPosition(env), (3, 3)), (3, 4)), (4, 3)))
    player = P.MCTSPlayer(env, DummyNet(env))
    P.initialize_game(player, pos)
    for _ in range(15):
        P.tree_search(player)
    pass_move = env.N * env.N
    assert player.root.children[pass_move].N == 1
    assert player.root.child_N[pass_move] == 1
    P.tree_search(player)
    assert player.root.child_N[pass_move] == 1


def test_extract_data_normal_end():                      # :285-301
    player = P.MCTSPlayer(env, DummyNet(env))
    P.initialize_game(player)
    P.tree_search(player)
    P.play_move(player, None)
    P.tree_search(player)
    P.play_move(player, None)
    assert M.is_done(player.root)
    P.set_result(player, go.result(player.root.position), False)
    positions, pis, results = P.extract_data(player)
    assert len(positions) == len(pis) == len(results) == 2
    assert results[0] == WHITE
    assert player.result_string == "W+%.1f" % player.root.position.komi


def test_extract_data_resign_end():                      # :303-321
    player = P.MCTSPlayer(env, DummyNet(env))
    P.initialize_game(player)
    P.tree_search(player)
    P.play_move(player, (0, 0))
    P.tree_search(player)
    P.play_move(player, None)
    P.tree_search(player)
    assert go.result(player.root.position) == BLACK
    P.set_result(player, WHITE, True)
    positions, pis, results = P.extract_data(player)
    assert results[0] == WHITE
    assert player.result_string == "W+R"
