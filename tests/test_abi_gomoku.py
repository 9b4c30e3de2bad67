"""Gomoku (src/game/gomoku/) through the C ABI: the second game behind the reference's Position interface runs on the same tree
kernels (action space N^2, no pass; rules = tree.cuh game_play / game_legal / game_score).  Position hooks, whole self-play
games and the player surface against the oracle (oracle/gomoku.py), on the emulator and on the GPU; the network with an N^2-wide
policy head and network-driven self-play on the GPU."""
import numpy as np
import pytest

from backends import BACKENDS, agz, lib_for
from oracle import gomoku as ogm
from oracle import net as onet
from oracle import selfplay as osp
from test_abi_mcts import DummyNet
from test_abi_nn import EngineBackedNet, hist_stack, push_oracle_net

f32 = np.float32
B, W = 1, -1


@pytest.fixture(scope="module", params=BACKENDS)
def lib(request):
    return lib_for(request.param)


def test_env_surface(lib):
    env = agz.GomokuEnv(lib_path=lib)
    assert (env.N, env.n_in_row, env.action_space, env.planes, env.max_action_space) == (15, 5, 225, 8, 361)
    eng = agz.Engine(9, lib_path=lib, game=agz.GAME_GOMOKU, n_in_row=4)
    assert eng.A == 81 and eng.cfg.game == agz.GAME_GOMOKU and eng.cfg.n_in_row == 4
    assert eng.cfg.noise_alpha == float(f32(0.03 * 361 / 81))                  # mcts.jl:22 with action_space = N^2
    eng.close()
    with pytest.raises(agz.AgzError):
        agz.Engine(9, lib_path=lib, game=agz.GAME_GOMOKU, n_in_row=10)
    assert isinstance(agz.api.Position(env), agz.GomokuPosition) and isinstance(agz.api.Position(agz.Go(9)), agz.GoPosition)


@pytest.mark.parametrize("N,k", [(7, 4), (9, 5), (15, 5), (19, 5), (5, 5)])
def test_random_playouts_match_oracle(lib, N, k):
    """play_move! / all_legal_moves / has_game_ended (gomoku board.jl:93-169) at every ply of random games."""
    env, oenv = agz.GomokuEnv(N, k, lib_path=lib), ogm.GomokuEnv(N, k)
    rs = np.random.RandomState(N * 10 + k)
    ended = {1: 0, -1: 0, 0: 0}
    for game in range(4 if N <= 9 else 2):
        pos, opos = agz.GomokuPosition(env), ogm.GomokuPosition(oenv)
        while not opos.done:
            legal = agz.all_legal_moves(pos)
            assert legal.shape == (N * N,) and np.array_equal(legal, ogm.all_legal_moves(opos))
            mv = ogm.from_flat(int(rs.choice(np.flatnonzero(legal))), oenv)
            pos, opos = agz.play_move(pos, mv), ogm.play_move(opos, mv)
            assert np.array_equal(pos.board, opos.board) and pos.to_play == opos.to_play and pos.n == opos.n
            assert pos.done == opos.done and pos.winner == opos.winner
            if rs.rand() < 0.2:
                occupied = np.flatnonzero(legal == 0)
                if len(occupied) and not opos.done:
                    with pytest.raises(agz.IllegalMove):
                        agz.play_move(pos, ogm.from_flat(int(occupied[0]), oenv))
        assert agz.result(pos) == ogm.result(opos) and agz.result_string(pos) == ogm.result_string(opos)
        ended[ogm.result(opos)] += 1
        with pytest.raises(AssertionError):
            agz.play_move(pos, (0, 0))
    assert sum(ended.values()) > 0


def test_lines_in_all_directions(lib):
    env = agz.GomokuEnv(9, 5, lib_path=lib)
    for cells in ([(3, c) for c in range(1, 6)], [(r, 8) for r in range(4, 9)], [(4 + t, 4 + t) for t in range(5)], [(t, 8 - t) for t in range(5)],
                  [(4 + t, 4 - t) for t in range(5)]):
        for color in (B, W):
            b = np.zeros((9, 9), np.int8)
            for c in cells:
                b[c] = color
            p = agz.GomokuPosition(env, board=b)
            assert p.done and p.winner == color
            b[cells[2]] = -color
            p = agz.GomokuPosition(env, board=b)
            assert not p.done and p.winner == 0
    full = np.array([[1 if ((i // 2 + j) % 2 == 0) else -1 for j in range(4)] for i in range(4)], np.int8)
    e4 = agz.GomokuEnv(4, 4, lib_path=lib)
    p = agz.GomokuPosition(e4, board=full)
    assert p.done and p.winner == 0 and agz.result_string(p) == "DRAW"


def check_selfplay(lib, N, k, readouts, seed, n_games=2, priors_seed=None, value=0.0, **kw):
    env, oenv = agz.GomokuEnv(N, k, lib_path=lib), ogm.GomokuEnv(N, k)
    A = N * N
    pri = None if priors_seed is None else np.random.RandomState(priors_seed).dirichlet(np.ones(A) * 0.5).astype(f32)
    net = DummyNet(A, fake_priors=pri, fake_value=value) if pri is not None else DummyNet(A, fake_value=value)
    recs = agz.selfplay(env, net, readouts, seed=seed, n_games=n_games, **kw)
    out = []
    for gid, r in enumerate(recs):
        op = osp.selfplay(oenv, net, readouts, seed=seed, game_id=gid)
        assert list(r.record.moves) == [ogm.to_flat(m.move, oenv) for m in op.root.position.recent], (seed, gid)
        assert r.result == op.result and r.result_string == op.result_string
        if r.n_moves:
            assert np.array_equal(np.array(op.searches_N), r.record.visits)
            assert np.array_equal(np.array(op.searches_pi, dtype=f32), r.record.searches_pi)
        assert np.array_equal(np.array(op.qs, dtype=f32), r.record.qs)
        out.append(r.result_string)
    return out


def test_selfplay_matches_oracle(lib):
    """Whole games bit-exact: moves, visit counts, pi, q, result -- wins of both colours, full-board draws, resignations."""
    seen = set()
    seen.update(check_selfplay(lib, 7, 4, 16, 1))
    seen.update(check_selfplay(lib, 9, 5, 16, 2, priors_seed=3, value=0.1))
    seen.update(check_selfplay(lib, 5, 5, 8, 3, n_games=3, priors_seed=5))
    seen.update(check_selfplay(lib, 6, 4, 16, 7, n_games=3, priors_seed=1, value=-0.95))       # around the resign threshold
    seen.update(check_selfplay(lib, 7, 4, 16, 5, n_games=5, concurrent=2, options={"dummy.fused_rounds": 0}))
    assert "DRAW" in seen and ("B" in seen or "W" in seen)


def test_selfplay_15x15(lib):
    check_selfplay(lib, 15, 5, 8, 4, n_games=2, priors_seed=6, value=-0.2)


def test_player_surface(lib):
    """MCTSPlayer over a Gomoku env: tree_search!, pick_move, play_move!, is_done, set_result!, extract_data."""
    env = agz.GomokuEnv(7, 4, lib_path=lib)
    pl = agz.MCTSPlayer(env, DummyNet(49), num_readouts=16, seed=2)
    pl.initialize_game()
    root = pl.root
    assert root.child_N.shape == (49,) and root.legal_moves().all() and root.position.to_play == B
    while not pl.is_done():
        for _ in range(3):
            pl.tree_search()
        mv = pl.pick_move()
        assert mv is not None and pl.play_move(mv)
    assert not pl.play_move(None)                                      # IllegalMove is caught by play_move!(player, c)
    pl.set_result(agz.result(pl.root.position), False)
    assert pl.result_string in ("B", "W", "DRAW")
    positions, pis, results = pl.extract_data()
    assert len(positions) == len(pis) == len(results) == pl.root.position.n
    assert all(isinstance(p, agz.GomokuPosition) for p in positions) and positions[1].board.any()
    assert all(r == pl.result for r in results)


@pytest.mark.gpu
@pytest.mark.parametrize("evaluator,tol", [pytest.param(agz.EVAL_NN_F32, 2e-5, id="f32"), pytest.param(agz.EVAL_NN_TC, 1e-3, id="tcgen05")])
@pytest.mark.parametrize("N,T", [(9, 2), (15, 1)])
def test_network_with_n2_policy_head(evaluator, tol, N, T):
    """NeuralNet(env::GomokuEnv): Dense(2N^2, env.action_space) with action_space = N^2 (neural_net.jl:30)."""
    oenv = ogm.GomokuEnv(N, 5)
    rs = np.random.RandomState(N)
    poss = []
    while len(poss) < 12:
        p = ogm.GomokuPosition(oenv)
        for _ in range(rs.randint(0, 40)):
            p = ogm.play_move(p, ogm.from_flat(int(rs.choice(np.flatnonzero(ogm.all_legal_moves(p)))), oenv))
            if p.done:
                break
        if not p.done:
            poss.append(p)
    nn = onet.NeuralNet(N, T, seed=3, action_space=N * N)
    nn.randomize_bn(seed=1)
    pi_ref, v_ref = nn(poss)
    assert pi_ref.shape == (N * N, 12)
    eng = agz.Engine(N, lib_path=lib_for("cuda"), n_games=8, tower_height=T, game=agz.GAME_GOMOKU, n_in_row=5)
    push_oracle_net(eng, nn)
    pi, v = eng.net_forward(evaluator, np.stack([hist_stack(p) for p in poss]), np.array([p.to_play for p in poss], np.int8))
    assert pi.shape == (12, N * N)
    assert np.abs(pi - pi_ref.T).max() <= tol and np.abs(v - v_ref).max() <= tol
    eng.close()


@pytest.mark.gpu
def test_nn_selfplay_matches_oracle_tree():
    N, T, R = 9, 1, 16
    nn = onet.NeuralNet(N, T, seed=5, action_space=N * N)
    nn.randomize_bn(seed=2)
    kw = dict(lib_path=lib_for("cuda"), tower_height=T, game=agz.GAME_GOMOKU, n_in_row=5)
    helper = agz.Engine(N, n_games=2, **kw)
    push_oracle_net(helper, nn)
    eng = agz.Engine(N, n_games=4, readouts=R, seed=9, **kw)
    push_oracle_net(eng, nn)
    eng.set_evaluator(agz.EVAL_NN_TC)
    recs = eng.selfplay_run(4)
    oenv = ogm.GomokuEnv(N, 5)
    onn = EngineBackedNet(helper, agz.EVAL_NN_TC, T)
    for gid, r in enumerate(recs):
        op = osp.selfplay(oenv, onn, R, seed=9, game_id=gid)
        assert list(r.moves) == [ogm.to_flat(m.move, oenv) for m in op.root.position.recent], gid
        assert np.array_equal(np.array(op.searches_N), r.visits)
        assert r.result == op.result and r.result_string == op.result_string
    for e in (helper, eng):
        e.close()


@pytest.mark.gpu
def test_replay_tuples_and_training_step():
    """extract_data / replay_position (gomoku board.jl:195-217) on the device: the ring's tuples are the positions before each move
    replayed with the Gomoku rules, pi is N^2 wide, z = the game's result; one optimisation step runs on them (train.jl:66-70)."""
    N = 7
    eng = agz.Engine(N, lib_path=lib_for("cuda"), n_games=4, readouts=16, seed=21, tower_height=1, game=agz.GAME_GOMOKU, n_in_row=4)
    eng.set_dummy_evaluator(None, 0.0)
    eng.selfplay_start(6)
    for _ in range(400):
        pr = eng.selfplay_step(16)
        if pr.games_finished == 6:
            break
    assert pr.games_finished == 6 and pr.error == 0
    total = eng.replay_gather()
    recs = eng.selfplay_harvest(16)
    assert total == sum(r.n_moves for r in recs) and total > 0
    boards, tp, pis, zs = eng.replay_read(0, total)
    assert pis.shape == (total, N * N)
    oenv = ogm.GomokuEnv(N, 4)
    expected = []
    for r in recs:
        pos = ogm.GomokuPosition(oenv)
        for t, m in enumerate(r.moves):
            expected.append((pos.board.flatten(order="F").copy(), pos.to_play, r.searches_pi[t], r.result))
            pos = ogm.play_move(pos, ogm.from_flat(int(m), oenv))
        assert pos.done or r.resigned
    used = set()
    for k in range(total):
        hit = next((i for i, (b, p, pi, z) in enumerate(expected)
                    if i not in used and p == tp[k] and z == zs[k] and np.array_equal(b, boards[k]) and np.array_equal(pi, pis[k])), None)
        assert hit is not None, k
        used.add(hit)
    nn = agz.NeuralNet(agz.GomokuEnv(N, 4, lib_path=lib_for("cuda")), tower_height=1, seed=1)
    nn.push(eng)
    l0 = eng.train_step_from_replay(16, seed=1)
    l1 = eng.train_step_from_replay(16, seed=1)
    assert np.isfinite(l0) and np.isfinite(l1) and l1 < l0          # the same batch again: the loss went down
    eng.close()
