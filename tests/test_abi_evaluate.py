"""`evaluate` (src/neural_net.jl:103-158) through the batched match entry points (agz_match_start / _search / _play):
every game's moves, result, result string and the win counter must equal the oracle's sequential two-player loop."""
import numpy as np
import pytest

from backends import BACKENDS, agz, lib_for
from oracle import evaluate as oev
from oracle import go as ogo

f32 = np.float32


class DummyNet:                                          # test_mcts_player.jl:10-32
    def __init__(self, A, fake_priors=None, fake_value=0):
        self.fake_priors = (np.ones(A) / A if fake_priors is None else np.asarray(fake_priors)).astype(f32)
        self.fake_value = f32(fake_value)

    def __call__(self, positions):
        n = len(positions)
        return np.repeat(self.fake_priors[:, None], n, axis=1), np.repeat(self.fake_value, n)


def _nets(A, seed):
    rs = np.random.RandomState(seed)
    pb = rs.rand(A).astype(f32) + f32(0.05)
    pw = rs.rand(A).astype(f32) + f32(0.05)
    return DummyNet(A, pb / pb.sum(), 0.05), DummyNet(A, pw / pw.sum(), -0.1)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("N,games,ro", [(5, 6, 24), (9, 3, 16)])
def test_evaluate_matches_oracle(backend, N, games, ro):
    env = agz.GoEnv(N, lib_path=lib_for(backend))
    oenv = ogo.GoEnv(N)
    black, white = _nets(N * N + 1, 7 + N)
    want = []
    ok_o, won_o = oev.evaluate(oenv, black, white, num_games=games, ro=ro, seed=3, details=want)
    got = []
    ok = agz.evaluate(env, black, white, num_games=games, ro=ro, seed=3, details=got)
    assert len(got) == games
    for g, (w, h) in enumerate(zip(want, got)):
        assert h.moves == w["moves"], "game %d: moves differ" % g
        assert h.result == w["result"] and h.result_string == w["result_string"], "game %d: result differs" % g
        assert h.black_won == w["black_won"]
    assert ok == ok_o and sum(h.black_won for h in got) == won_o


@pytest.mark.parametrize("backend", BACKENDS)
def test_evaluate_forced_resignation(backend):
    """A hopeless value estimate makes the side to move resign at once (neural_net.jl:129-133); the win counter still reads
    the area score of the final (empty) position: komi 7.5 -> White."""
    N = 5
    env = agz.GoEnv(N, lib_path=lib_for(backend))
    black = DummyNet(N * N + 1, fake_value=-0.99)
    white = DummyNet(N * N + 1, fake_value=-0.99)
    got, want = [], []
    ok = agz.evaluate(env, black, white, num_games=2, ro=16, details=got)
    ok_o, _ = oev.evaluate(ogo.GoEnv(N), black, white, num_games=2, ro=16, details=want)
    for w, h in zip(want, got):
        assert h.moves == w["moves"] and h.result_string == w["result_string"] == "W+R" and h.result == -1
        assert not h.black_won
    assert ok == ok_o == False   # noqa: E712


@pytest.mark.parametrize("backend", BACKENDS)
def test_match_play_illegal_move_leaves_tree_unchanged(backend):
    eng = agz.Engine(5, lib_path=lib_for(backend), n_games=2, readouts=8, tau_threshold=-1, inject_noise=0)
    eng.match_start()
    done, _ = eng.match_play(np.array([12, 12], np.int32))
    assert not done.any()
    with pytest.raises(agz.IllegalMove):                 # occupied point (play_move!(player, c) catches IllegalMove, mcts_play.jl:41)
        eng.match_play(np.array([12, -1], np.int32))
    root, count = eng.tree_root(0)
    assert eng.tree_read_node(0, root).n == 1
    done, _ = eng.match_play(np.array([25, 25], np.int32))   # pass
    done, sc = eng.match_play(np.array([25, 25], np.int32))  # second pass ends the game
    assert done.all() and (sc == 17.5).all()   # one black stone owns the whole 5x5 board: 25 - 7.5
    eng.close()


@pytest.mark.gpu
def test_evaluate_with_networks_on_gpu():
    """Two random-init residual nets (tensor-core evaluator) play 8 concurrent gating games; the same nets through the fp32
    SIMT evaluator must produce games that replay legally, and identical nets with identical seeds give identical games."""
    lib_for("cuda")
    env = agz.GoEnv(9)
    a = agz.NeuralNet(env, tower_height=1, seed=1)
    b = agz.NeuralNet(env, tower_height=1, seed=2)
    g1, g2 = [], []
    agz.evaluate(env, a, b, num_games=8, ro=32, seed=5, details=g1)
    agz.evaluate(env, a, b, num_games=8, ro=32, seed=5, details=g2)
    oenv = ogo.GoEnv(9)
    for x, y in zip(g1, g2):
        assert x.moves == y.moves and x.result_string == y.result_string
        pos = ogo.GoPosition(oenv)
        for mv in x.moves:                                # every move is legal under the oracle's rules
            pos = ogo.play_move(pos, ogo.from_flat(mv, oenv))
        assert x.result != 0 or x.result_string == "DRAW"


@pytest.mark.parametrize("backend", BACKENDS)
def test_play_scripted_game(backend):
    """play(env, nn; ...) (src/play.jl:25-77) with a scripted human: illegal and unparsable inputs are asked again, the engine answers
    with searched moves, two passes end the game and the result string is the area score."""
    env = agz.GoEnv(5, lib_path=lib_for(backend))
    net = DummyNet(26, fake_value=0.0)
    script = iter(["C3", "zz", "C3", "pass", "pass", "pass", "pass", "pass", "pass", "pass", "pass", "pass", "pass", "pass", "pass"] + ["pass"] * 40)
    said = []
    az = agz.play(env, net, num_readouts=16, mode=0, input_fn=lambda prompt: next(script), print_fn=said.append)
    assert az.is_done() and az.result_string
    assert any(line == "Try again." for line in said)            # "zz" does not parse
    assert said[-1].startswith(("You Win! ", "AlphaZero wins! "))
    moves = [mv for _, mv in az.recent]
    assert moves[0] == (2, 2) and len(moves) >= 3 and moves.count((2, 2)) == 1   # the second "C3" was rejected by play_move!
