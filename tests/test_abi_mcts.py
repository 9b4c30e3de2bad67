"""The reference's test/test_mcts.jl and test/test_mcts_player.jl driven through the C ABI (device tree code),
plus bit-exact self-play parity against the oracle (visit counts, chosen moves, pi, q, result)."""
import numpy as np
import pytest

from backends import BACKENDS, agz, lib_for
from oracle import go as ogo
from oracle import mcts as OM
from oracle import mcts_play as OP
from oracle import selfplay as osp
from oracle import replay as orp
from refboards import load_board, ALMOST_DONE_BOARD, TT_FTW_BOARD

BLACK, WHITE = 1, -1
f32 = np.float32


class DummyNet:                                          # test_mcts_player.jl:10-32
    def __init__(self, A=82, fake_priors=None, fake_value=0):
        self.fake_priors = (np.ones(A) / A if fake_priors is None else np.asarray(fake_priors)).astype(f32)
        self.fake_value = f32(fake_value)

    def __call__(self, positions):
        n = len(positions)
        return np.repeat(self.fake_priors[:, None], n, axis=1), np.repeat(self.fake_value, n)


@pytest.fixture(scope="module", params=BACKENDS)
def env(request):
    return agz.GoEnv(9, lib_path=lib_for(request.param))


A = 82


def root_of(env, pos=None, net=None, **kw):
    pl = agz.MCTSPlayer(env, net or DummyNet(), **kw)
    pl.initialize_game(pos)
    return pl, pl.root


def send_two_return_one(env, n=75, komi=0.5, caps=(0, 0), three=True):
    recent = [(BLACK, (0, 1)), (WHITE, (0, 8))] + ([(BLACK, (1, 0))] if three else [])
    return agz.GoPosition(env, board=load_board(ALMOST_DONE_BOARD, env), n=n, komi=komi, caps=caps, recent=recent,
                          to_play=WHITE if three else BLACK)


def test_action_flipping(env):                           # test_mcts.jl:45-59
    rs = np.random.RandomState(1)
    probs = (0.02 * np.ones(A) + rs.rand(A) * 0.001).astype(f32)
    _, broot = root_of(env)
    _, wroot = root_of(env, agz.GoPosition(env, to_play=WHITE))
    agz.incorporate_results(agz.select_leaf(broot), probs, 0)
    agz.incorporate_results(agz.select_leaf(wroot), probs, 0)
    bl, wl = agz.select_leaf(broot), agz.select_leaf(wroot)
    assert bl.fmove == wl.fmove
    assert (broot.child_action_score() == wroot.child_action_score()).all()


def test_select_leaf(env):                               # :61-70
    flattened = agz.to_flat(agz.from_kgs("D9", env), env)
    probs = (0.02 * np.ones(A)).astype(f32)
    probs[flattened] = 0.4
    _, root = root_of(env, send_two_return_one(env))
    agz.incorporate_results(agz.select_leaf(root), probs, 0)
    assert root.position.to_play == WHITE
    assert agz.select_leaf(root) == root.children[flattened]


def test_backup_incorporate_results(env):                # :72-114
    probs = (0.02 * np.ones(A)).astype(f32)
    _, root = root_of(env, send_two_return_one(env))
    agz.incorporate_results(agz.select_leaf(root), probs, 0)
    leaf = agz.select_leaf(root)
    agz.incorporate_results(leaf, probs, -1)
    assert root.N == 2
    assert root.Q == pytest.approx(-1 / 3, rel=1e-6)
    assert root.child_N[leaf.fmove] == 1 and leaf.N == 1
    assert root.child_Q()[leaf.fmove] == -0.5
    assert leaf.Q == pytest.approx(-0.5)
    leaf2 = agz.select_leaf(root)
    agz.incorporate_results(leaf2, probs, -0.2)
    assert root.N == 3 and root.Q == pytest.approx(-0.3, rel=1e-6)
    assert leaf.N == 2 and leaf2.N == 1 and leaf2.parent == leaf
    assert leaf.Q == pytest.approx(-0.4, rel=1e-6)
    assert leaf.child_Q()[leaf2.fmove] == pytest.approx(-0.6, rel=1e-6)
    assert leaf2.Q == pytest.approx(-0.6, rel=1e-6)


def test_do_not_explore_past_finish(env):                # :116-127
    probs = (0.02 * np.ones(A)).astype(f32)
    _, root = root_of(env)
    agz.incorporate_results(agz.select_leaf(root), probs, 0)
    first_pass = agz.maybe_add_child(root, 81)
    agz.incorporate_results(first_pass, probs, 0)
    second_pass = agz.maybe_add_child(first_pass, 81)
    with pytest.raises(AssertionError):
        agz.incorporate_results(second_pass, probs, 0)
    assert agz.select_leaf(second_pass) == second_pass


def test_add_child_and_idempotency(env):                 # :129-144
    _, root = root_of(env)
    child = agz.maybe_add_child(root, 16)
    assert 16 in root.children and child.parent == root and child.fmove == 16
    before = root.children
    assert agz.maybe_add_child(root, 16) == child
    assert before == root.children


def test_never_select_illegal_moves(env):                # :146-167
    probs = (0.02 * np.ones(A)).astype(f32)
    probs[1] = 0.99
    pl, root = root_of(env, send_two_return_one(env))
    agz.incorporate_results(root, probs, 0)
    legal = root.legal_moves().astype(bool)
    assert not legal[1]
    cn = root.child_N
    cn[legal] = 10000
    pl.engine.tree_set_stats(0, root.id, self_N=10000, child_N=cn)
    assert agz.select_leaf(root).fmove != 1
    for _ in range(10):
        agz.inject_noise(root)
        assert agz.select_leaf(root).fmove != 1


def test_dont_pick_unexpanded_child(env):                # :169-183
    probs = (0.02 * np.ones(A)).astype(f32)
    probs[17] = 0.999
    _, root = root_of(env)
    agz.incorporate_results(root, probs, 0)
    leaf1 = agz.select_leaf(root)
    assert leaf1.fmove == 17
    agz.add_virtual_loss(leaf1)
    assert agz.select_leaf(root) == leaf1


def test_node_features(env):                             # test_features.jl:48-79 via tree nodes
    pl, root = root_of(env)
    for c in ((0, 0), (0, 1), (0, 2), (0, 3), (1, 1)):
        assert pl.play_move(c)
    f = pl.engine.tree_node_features(0, pl.root.id).reshape(17, 9, 9)      # [c, j, i]
    lb = lambda two_rows: load_board(two_rows + "." * 9 + "\n" + ("." * 9 + "\n") * 6, env).T
    assert (f[0] == lb("...X.....\n.........\n")).all()
    assert (f[1] == lb("X.X......\n.X.......\n")).all()
    assert (f[2] == lb(".X.X.....\n.........\n")).all()
    assert (f[3] == lb("X.X......\n.........\n")).all()
    assert (f[4] == lb(".X.......\n.........\n")).all()
    assert (f[5] == lb("X.X......\n.........\n")).all()
    assert (f[10:16] == 0).all() and (f[16] == -1).all()


# ------------------------------------------------------------ test_mcts_player.jl
def basic_player(env):                                   # :59-66
    pl, root = root_of(env)
    first = agz.select_leaf(root)
    agz.incorporate_results(first, DummyNet().fake_priors, 0)
    return pl


def almost_done_player(env):                             # :68-77
    probs = np.ones(A) * 0.001
    probs[2:5] = 0.2
    probs[-1] = 0.2
    pl, _ = root_of(env, send_two_return_one(env, n=70, komi=2.5, caps=(1, 4), three=False), DummyNet(fake_priors=probs))
    return pl


def test_inject_noise(env):                              # :93-109
    pl = basic_player(env)
    s = pl.root.child_prior.sum()
    assert s == pytest.approx(1, rel=1e-5)
    agz.inject_noise(pl.root)
    assert pl.root.child_prior.sum() == pytest.approx(s, rel=1e-5)
    assert pl.root.child_prior.max() > 3 / A


def test_pick_moves(env):                                # :111-137
    pl = basic_player(env)
    cn = pl.root.child_N
    cn[agz.to_flat((2, 0), env)] = 10
    cn[agz.to_flat((1, 0), env)] = 5
    cn[agz.to_flat((3, 0), env)] = 1
    pl.engine.tree_set_stats(0, pl.root.id, child_N=cn, n_override=A)
    assert pl.pick_move() == (2, 0)
    pl.engine.tree_set_stats(0, pl.root.id, n_override=3)
    assert pl.pick_move() in ((2, 0), (1, 0), (3, 0))


def test_dont_pass_if_losing(env):                       # :139-165
    pl = almost_done_player(env)
    assert agz.score(pl.root.position) == -0.5
    for _ in range(20):
        pl.tree_search()
    flattened = agz.to_flat(agz.from_kgs("D9", env), env)
    root = pl.root
    assert int(np.argmax(root.child_N)) == flattened
    assert root.children[flattened].Q > 0
    assert root.N >= 20
    assert root.child_Q()[-1] < 0
    assert pl.engine.tree_pending_vlosses(0) == 0


def test_parallel_tree_search(env):                      # :167-202
    pl = almost_done_player(env)
    pl.tree_search(1)
    for _ in range(6):
        pl.tree_search(10)
    flattened = agz.to_flat(agz.from_kgs("D9", env), env)
    cn = pl.root.child_N
    assert flattened in np.flatnonzero(cn == cn.max())
    assert pl.root.children[flattened].Q > 0 and pl.root.N >= 20
    assert pl.engine.tree_pending_vlosses(0) == 0
    pl2 = almost_done_player(env)
    for _ in range(10):
        pl2.tree_search(50)
    assert pl2.engine.tree_pending_vlosses(0) == 0


def test_long_game_tree_search(env):                     # :204-225
    maxlen = 81 * 7 // 5
    endgame = agz.GoPosition(env, board=load_board(TT_FTW_BOARD, env), n=maxlen - 2, komi=2.5,
                             recent=[(BLACK, (0, 1)), (WHITE, (0, 8))], to_play=BLACK)
    pl, _ = root_of(env, endgame)
    for _ in range(10):
        pl.tree_search(8)
    assert pl.engine.tree_pending_vlosses(0) == 0
    assert pl.root.Q > 0


def test_cold_start_parallel_tree_search(env):           # :227-240
    pl, root = root_of(env, None, DummyNet(fake_value=0.17))
    assert root.N == 0 and not root.is_expanded
    pl.tree_search(4)
    assert pl.engine.tree_pending_vlosses(0) == 0
    assert pl.root.N == 1
    assert pl.root.Q == pytest.approx(0.085, rel=1e-6)


def test_tree_search_failsafe(env):                      # :242-252
    probs = np.ones(A) * 0.001
    probs[-1] = 1
    pl, _ = root_of(env, agz.api.pass_move(agz.GoPosition(env)), DummyNet(fake_priors=probs))
    pl.tree_search(1)
    assert pl.engine.tree_pending_vlosses(0) == 0


def test_only_check_game_end_once(env):                  # :254-283
    pos = agz.GoPosition(env)
    for c in ((3, 3), (3, 4), (4, 3), None):
        pos = agz.play_move(pos, c)
    pl, _ = root_of(env, pos)
    for _ in range(15):
        pl.tree_search()
    assert pl.root.children[81].N == 1 and pl.root.child_N[81] == 1
    pl.tree_search()
    assert pl.root.child_N[81] == 1


def test_extract_data(env):                              # :285-321
    pl, _ = root_of(env)
    pl.tree_search()
    pl.play_move(None)
    pl.tree_search()
    pl.play_move(None)
    assert pl.is_done()
    pl.set_result(agz.result(pl.root.position), False)
    positions, pis, results = pl.extract_data()
    assert len(positions) == len(pis) == len(results) == 2
    assert results[0] == WHITE and pl.result_string == "W+7.5"
    pl, _ = root_of(env)
    pl.tree_search()
    pl.play_move((0, 0))
    pl.tree_search()
    pl.play_move(None)
    pl.tree_search()
    assert agz.result(pl.root.position) == BLACK
    pl.set_result(WHITE, True)
    positions, pis, results = pl.extract_data()
    assert results[0] == WHITE and pl.result_string == "W+R"


def test_generic_callable_network(env):
    """`network` may be any callable positions -> (A x B, B): same tree as the on-device dummy evaluator."""
    rs = np.random.RandomState(3)
    pri = rs.dirichlet(np.ones(A)).astype(f32)

    class HostNet:
        def __call__(self, positions):
            n = len(positions)
            return np.repeat(pri[:, None], n, axis=1), np.repeat(f32(0.1), n)

    p1, _ = root_of(env, None, HostNet(), seed=5)
    p2, _ = root_of(env, None, DummyNet(fake_priors=pri, fake_value=0.1), seed=5)
    for _ in range(6):
        p1.tree_search()
        p2.tree_search()
    assert (p1.root.child_N == p2.root.child_N).all() and (p1.root.child_W == p2.root.child_W).all()


# ------------------------------------------------------------ self-play parity with the oracle
def check_selfplay_parity(env, N, readouts, seeds, priors_seed=None, value=0.0, n_games=1, **kw):
    oenv = ogo.GoEnv(N)
    A_ = N * N + 1
    if priors_seed is None:
        net = DummyNet(A_, fake_value=value)
    else:
        pri = np.random.RandomState(priors_seed).dirichlet(np.ones(A_) * 0.5).astype(f32)
        net = DummyNet(A_, fake_priors=pri, fake_value=value)
    OM.MAX_GAME_LENGTH_OVERRIDE = kw.get("max_game_length")
    for seed in seeds:
        recs = agz.selfplay(env, net, readouts, seed=seed, n_games=n_games, **kw)
        recs = [recs] if n_games == 1 else recs
        for gid, r in enumerate(recs):
            op = osp.selfplay(oenv, net, readouts, seed=seed, game_id=gid)
            om = [ogo.to_flat(m.move, oenv) for m in op.root.position.recent]
            assert list(r.record.moves) == om, (seed, gid)
            assert r.result == op.result and r.result_string == op.result_string
            assert len(op.searches_N) == r.n_moves
            if r.n_moves:                                    # a game resigned before its first move has empty records
                assert np.array_equal(np.array(op.searches_N), r.record.visits), (seed, gid)
                assert np.array_equal(np.array(op.searches_pi, dtype=f32), r.record.searches_pi), (seed, gid)
            assert np.array_equal(np.array(op.qs, dtype=f32), r.record.qs), (seed, gid)
    OM.MAX_GAME_LENGTH_OVERRIDE = None


def test_selfplay_matches_oracle(env):
    check_selfplay_parity(env, 9, 24, seeds=[0, 1])
    check_selfplay_parity(env, 9, 16, seeds=[2], priors_seed=7, value=-0.2)


def test_selfplay_matches_oracle_with_separate_kernels(env):
    """With the DummyNet evaluator agz_selfplay_step plays all rounds of a call in one launch per game; the option
    dummy.fused_rounds = 0 runs the select / incorporate kernels of the network path instead.  Both must reproduce the oracle."""
    check_selfplay_parity(env, 9, 24, seeds=[0], options={"dummy.fused_rounds": 0})
    check_selfplay_parity(env, 9, 16, seeds=[3], priors_seed=4, value=0.1, n_games=3, options={"dummy.fused_rounds": 0})


def test_staggered_start_changes_timing_not_games(env):
    """Option selfplay.stagger_rounds delays the first game of slot g by g*R/n_games rounds (steady-state throughput runs); the
    games themselves are keyed by game id and stay bit-exact."""
    check_selfplay_parity(env, 9, 16, seeds=[5], n_games=6, concurrent=3, options={"selfplay.stagger_rounds": 40})
    check_selfplay_parity(env, 9, 16, seeds=[5], n_games=4, options={"selfplay.stagger_rounds": 7, "dummy.fused_rounds": 0})


def test_options_surface(env):
    eng = env.util_engine()
    assert eng.get_option("dummy.fused_rounds") == 1 and eng.get_option("selfplay.stagger_rounds") == 0
    eng.set_option("selfplay.stagger_rounds", 12)
    assert eng.get_option("selfplay.stagger_rounds") == 12
    eng.set_option("selfplay.stagger_rounds", 0)
    with pytest.raises(agz.AgzError):
        eng.set_option("no.such.option", 1)
    with pytest.raises(agz.AgzError):
        eng.set_option("selfplay.stagger_rounds", -1)


def test_selfplay_parity_sweep():
    """Board sizes, readouts, priors and value levels around the resign threshold (games that resign at once, late, or never):
    every game bit-exact against the oracle."""
    for N in (5, 7):
        e = agz.GoEnv(N, lib_path=lib_for("emu"))
        for seed in range(2):
            for ro, ps, val, ng in ((8, None, 0.0, 3), (16, seed + 1, -0.5, 2), (24, seed + 7, 0.95, 2), (12, seed + 3, -0.97, 3),
                                    (16, seed + 11, 0.88, 2), (8, seed + 5, -0.89, 2)):
                check_selfplay_parity(e, N, ro, seeds=[seed], priors_seed=ps, value=val, n_games=ng)


def test_tiny_evaluator_values_take_the_exact_division_path(env):
    """The reciprocal-table quotients are exact only while W / (1 + N) stays a normal Float32; an evaluator value below 2^-70 flags
    the game (GameState.tiny_values) and its PUCT scores use IEEE divisions from then on.  1e-35 / k reaches the subnormals."""
    check_selfplay_parity(env, 9, 16, seeds=[4], priors_seed=9, value=1e-35, n_games=2)
    check_selfplay_parity(env, 9, 16, seeds=[4], priors_seed=9, value=-3e-41, n_games=2)


def test_selfplay_concurrent_games_match_oracle(env):
    check_selfplay_parity(env, 9, 16, seeds=[3], n_games=3)


def test_selfplay_with_tiny_arena_compacts_every_move(env):
    """An arena barely larger than one move's worth of nodes forces the in-place compaction at (almost) every re-root."""
    check_selfplay_parity(env, 9, 16, seeds=[5], priors_seed=2, value=0.05, n_games=2, nodes_per_game=96)


def test_sharp_network_keeps_the_whole_subtree(env):
    """A one-hot-like prior sends every readout down one line, so the re-used subtree keeps almost every node the game ever
    created (the reference's tree is unbounded).  With the worst-case arena (max_game_length * (readouts + 2 * parallel)) the games
    equal the oracle's.  An arena of barely one move's worth does not stop the game any more: under that pressure the least-visited
    nodes are forgotten (agz_progress.arena_prunes counts it) and the game plays on to a legal end."""
    pri = np.full(A, 1e-4, f32)
    pri[40] = 1.0
    pri /= pri.sum()
    oenv = ogo.GoEnv(9)
    net = DummyNet(A, fake_priors=pri, fake_value=0.3)
    OM.MAX_GAME_LENGTH_OVERRIDE = 30
    try:
        recs = agz.selfplay(env, net, 16, seed=6, n_games=2, max_game_length=30, nodes_per_game=30 * 36 + 40)
        for gid, r in enumerate(recs):
            op = osp.selfplay(oenv, net, 16, seed=6, game_id=gid)
            assert list(r.record.moves) == [ogo.to_flat(m.move, oenv) for m in op.root.position.recent]
            assert np.array_equal(np.array(op.searches_N), r.record.visits)
        eng = agz.Engine(9, lib_path=env.lib_path, n_games=2, readouts=16, seed=6, max_game_length=30, nodes_per_game=40)
        eng.set_dummy_evaluator(pri, 0.3)
        eng.selfplay_start(2)
        for _ in range(200):
            pr = eng.selfplay_step(8)
            if pr.games_finished == 2:
                break
        assert pr.games_finished == 2 and pr.error == 0 and pr.arena_prunes > 0
        assert eng.selfplay_stats()["arena_prunes"] == pr.arena_prunes
        for r in eng.selfplay_harvest(4):
            pos = ogo.GoPosition(oenv)
            for m in r.moves:                                   # every recorded move is legal under the oracle's rules
                pos = ogo.play_move(pos, ogo.from_flat(int(m), oenv))
            assert r.n_moves == 30 or r.resigned or pos.done
        eng.close()
    finally:
        OM.MAX_GAME_LENGTH_OVERRIDE = None


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("N", [5, 13])
def test_selfplay_matches_oracle_other_board_sizes(backend, N):
    """5x5 (one bit-plane word, A = 26 < 32) and 13x13 (A = 170: the 6-word kernel variant) against the oracle."""
    e = agz.GoEnv(N, lib_path=lib_for(backend))
    check_selfplay_parity(e, N, 16, seeds=[8], priors_seed=3, value=0.05, n_games=2, **({"max_game_length": 40} if N == 13 else {}))


@pytest.mark.gpu
def test_selfplay_matches_oracle_bulk():
    """C2-like readouts on a handful of games, and many concurrent small games (slot refill, ring, compaction)."""
    env = agz.GoEnv(9, lib_path=lib_for("cuda"))
    check_selfplay_parity(env, 9, 400, seeds=[0], n_games=2)
    check_selfplay_parity(env, 9, 32, seeds=[11], priors_seed=5, value=0.1, n_games=24, concurrent=8)
    env19 = agz.GoEnv(19, lib_path=lib_for("cuda"))
    check_selfplay_parity(env19, 19, 16, seeds=[4], n_games=2, max_game_length=60)


@pytest.mark.gpu
def test_19x19_sharp_policy_long_game():
    """19x19, 800 readouts, a sharp (one-hot-like) prior that sends nearly every readout down one line, so the
    re-used subtree keeps ~800 nodes per move.  (i) With the default arena (worst case max_game_length * (readouts + 2 * parallel))
    60 moves are played without error or pruning and the first moves equal the oracle's bit for bit; (ii) with an arena of one move
    worth (+ 150 nodes) the game still plays its 60 moves: the least-visited nodes are forgotten (arena_prunes > 0), every move is legal."""
    N, R, L = 19, 800, 60
    A_ = N * N + 1
    pri = np.full(A_, 1e-5, f32)
    pri[3 * N + 3] = 1.0
    pri[15 * N + 15] = 0.5
    pri /= pri.sum()
    net = DummyNet(A_, fake_priors=pri, fake_value=0.2)
    oenv = ogo.GoEnv(N)
    lib = lib_for("cuda")
    eng = agz.Engine(N, lib_path=lib, n_games=2, readouts=R, seed=3, max_game_length=L)
    eng.set_dummy_evaluator(pri, 0.2)
    eng.selfplay_start(2)
    for _ in range(400):
        pr = eng.selfplay_step(100)
        if pr.games_finished == 2:
            break
    assert pr.games_finished == 2 and pr.error == 0 and pr.arena_prunes == 0
    recs = sorted(eng.selfplay_harvest(4), key=lambda r: r.game_id)
    eng.close()
    assert all(r.n_moves == L or r.resigned for r in recs)
    OM.MAX_GAME_LENGTH_OVERRIDE = L
    try:
        class Stop(Exception):
            pass

        moves_checked = 3
        seen = []

        def on_move(player, move):
            seen.append(player)
            if len(player.searches_pi) >= moves_checked:
                raise Stop()

        try:
            osp.selfplay(oenv, net, R, seed=3, game_id=0, on_move=on_move)
        except Stop:
            pass
        op = seen[-1]
        r = recs[0]
        k = min(moves_checked, r.n_moves)
        assert [ogo.to_flat(m.move, oenv) for m in op.root.position.recent][:k] == list(r.moves[:k])
        assert np.array_equal(np.array(op.searches_N[:k]), r.visits[:k])
        assert np.array_equal(np.array(op.searches_pi[:k], dtype=f32), r.searches_pi[:k])
    finally:
        OM.MAX_GAME_LENGTH_OVERRIDE = None
    small = agz.Engine(N, lib_path=lib, n_games=2, readouts=R, seed=3, max_game_length=L, nodes_per_game=R + 20 + 150)
    small.set_dummy_evaluator(pri, 0.2)
    small.selfplay_start(2)
    for _ in range(400):
        pr = small.selfplay_step(100)
        if pr.games_finished == 2:
            break
    assert pr.games_finished == 2 and pr.error == 0 and pr.arena_prunes > 0
    for r in small.selfplay_harvest(4):
        pos = ogo.GoPosition(oenv)
        for m in r.moves:
            pos = ogo.play_move(pos, ogo.from_flat(int(m), oenv))      # raises IllegalMove on an illegal move
        assert r.n_moves == L or r.resigned or pos.done
    small.close()


# ------------------------------------------------------------ BASELINE full size (C2): size-independent properties
@pytest.mark.gpu
def test_c2_full_size_properties():
    """1024 concurrent 9x9 games, 400 readouts, tower_height 6 on the tcgen05 path (BASELINE config C2), 6 moves deep:
    (i) two runs with the same seed are bit-identical, (ii) sharding the same global games over 2 ranks gives the same
    records (RNG keyed by global game id), (iii) per-move invariants: pi sums to 1, visit rows sum to N(root)-1 >= readouts,
    (iv) the recorded moves replayed through the ORACLE rules reproduce the engine's root boards."""
    lib = lib_for("cuda")
    env = agz.GoEnv(9, lib_path=lib)
    nn = agz.NeuralNet(env, tower_height=6, seed=0)
    G, R, steps = 1024, 400, 6

    def run(n_games, world, rank):
        eng = agz.Engine(9, lib_path=lib, n_games=n_games, readouts=R, tower_height=6, seed=123, world_size=world, rank=rank,
                         evaluator=agz.EVAL_NN_TC)
        nn.push(eng)
        eng.selfplay_start(-1)
        pr = eng.selfplay_step(50 * steps + 2)
        assert pr.error == 0
        return eng, pr

    e1, p1 = run(G, 1, 0)
    e2, p2 = run(G, 1, 0)
    e3, p3 = run(G // 2, 2, 1)                      # rank 1 of 2: global games 1, 3, 5, ...
    assert p1.moves_played == p2.moves_played and p1.readouts == p2.readouts
    oenv = ogo.GoEnv(9)
    for slot in list(range(0, G, 37)) + [1, 3, 5]:
        n1, mv1, q1, pi1 = e1.tree_read_record(slot)
        n2, mv2, q2, pi2 = e2.tree_read_record(slot)
        assert n1 == n2 >= steps - 1 and np.array_equal(mv1, mv2) and np.array_equal(pi1, pi2) and np.array_equal(q1, q2)
        assert np.allclose(pi1.sum(axis=1), 1.0, atol=1e-5)
        assert e1.tree_pending_vlosses(slot) == 0
        root, count = e1.tree_root(slot)
        view = e1.tree_read_node(slot, root)
        pos = ogo.GoPosition(oenv)
        for m in mv1:
            pos = ogo.play_move(pos, ogo.from_flat(int(m), oenv))     # raises IllegalMove if the engine played an illegal move
        assert np.array_equal(np.array(view.board[:81], np.int8), pos.board.flatten(order="F"))
        assert view.n == pos.n and view.to_play == pos.to_play
        if slot % 2 == 1:                               # the same global game on the 2-rank sharding
            n3, mv3, q3, pi3 = e3.tree_read_record(slot // 2)
            assert n3 == n1 and np.array_equal(mv3, mv1) and np.array_equal(pi3, pi1)
    for e in (e1, e2, e3):
        e.close()


@pytest.mark.gpu
def test_c5_full_size_matches_oracle():
    """BASELINE config C5 at full size (9x9, 8192 trees x 1600 readouts, uniform prior, noise on): the first moves of sampled games
    equal the oracle's bit for bit (visit counts, pi, q, chosen moves), whatever slot of the 8192 the game runs in."""
    lib = lib_for("cuda")
    G, R, moves = 8192, 1600, 2
    eng = agz.Engine(9, lib_path=lib, n_games=G, readouts=R, seed=9, nodes_per_game=3600)
    eng.set_dummy_evaluator(None, 0.0)
    eng.selfplay_start(-1)
    pr = eng.selfplay_step(moves * (R // 8) + 2)
    assert pr.error == 0 and pr.moves_played >= moves * G
    oenv = ogo.GoEnv(9)
    net = DummyNet(82)

    class Stop(Exception):
        pass

    for slot in (0, 4097, 8191):
        n, mv, q, pi = eng.tree_read_record(slot)
        assert n >= moves and eng.tree_pending_vlosses(slot) == 0
        seen = []

        def on_move(player, move):
            seen.append(player)
            if len(player.searches_pi) >= moves:
                raise Stop()

        try:
            osp.selfplay(oenv, net, R, seed=9, game_id=slot, on_move=on_move)
        except Stop:
            pass
        op = seen[-1]
        assert [ogo.to_flat(m.move, oenv) for m in op.root.position.recent][:moves] == list(mv[:moves]), slot
        assert np.array_equal(np.array(op.searches_pi[:moves], dtype=f32), pi[:moves]), slot
        assert np.array_equal(np.array(op.qs[:moves], dtype=f32), q[:moves]), slot
    eng.close()


@pytest.mark.gpu
def test_c3_share_full_size_properties():
    """One GPU's share of BASELINE config C3 (19x19, 512 games, 800 readouts, tower_height 19): one full move of every game on the
    tensor-core network -- no device error, no pending virtual losses, pi sums to 1 with exactly the move's visits behind it, and
    the recorded move is legal under the oracle's rules."""
    lib = lib_for("cuda")
    env = agz.GoEnv(19, lib_path=lib)
    nn = agz.NeuralNet(env, tower_height=19, seed=0)
    eng = agz.Engine(19, lib_path=lib, n_games=512, readouts=800, tower_height=19, seed=4, evaluator=agz.EVAL_NN_TC)
    nn.push(eng)
    eng.selfplay_start(-1)
    pr = eng.selfplay_step(104)
    assert pr.error == 0 and pr.moves_played >= 512
    oenv = ogo.GoEnv(19)
    for slot in range(0, 512, 73):
        n, mv, q, pi = eng.tree_read_record(slot)
        assert n >= 1 and eng.tree_pending_vlosses(slot) == 0
        assert abs(float(pi[0].sum()) - 1.0) < 1e-5 and -1.0 <= q[0] <= 1.0
        ogo.play_move(ogo.GoPosition(oenv), ogo.from_flat(int(mv[0]), oenv))
    eng.close()


# ------------------------------------------------------------ replay ring (extract_data / replay_position / get_replay_batch)
@pytest.mark.gpu
def test_replay_gather_read_sample():
    """agz_replay_gather packs (position before the move, pi, z) for every ply of every finished game -- the same tuples the
    oracle's extract_data / replay_position produce -- and agz_replay_sample draws distinct tuples from the ring."""
    eng = agz.Engine(9, lib_path=lib_for("cuda"), n_games=4, readouts=16, seed=77)
    eng.set_dummy_evaluator(None, 0.0)
    eng.selfplay_start(6)
    for _ in range(400):
        pr = eng.selfplay_step(16)
        if pr.games_finished == 6:
            break
    assert pr.games_finished == 6 and pr.error == 0
    total = eng.replay_gather()
    recs = sorted(eng.selfplay_harvest(16), key=lambda r: r.game_id)
    assert total == sum(r.n_moves for r in recs) and total > 0
    boards, tp, pis, zs = eng.replay_read(0, total)
    # tuples are appended game by game in ring order; rebuild the expected stream from the harvested records + oracle rules
    oenv = ogo.GoEnv(9)
    expected = {}
    for r in recs:
        pos = ogo.GoPosition(oenv)
        for t, m in enumerate(r.moves):
            expected[(r.game_id, t)] = (pos.board.flatten(order="F").copy(), pos.to_play, r.searches_pi[t], r.result)
            pos = ogo.play_move(pos, ogo.from_flat(int(m), oenv))
    # match every ring tuple to exactly one expected tuple
    used = set()
    for k in range(total):
        hit = None
        for key, (b, p, pi, z) in expected.items():
            if key not in used and p == tp[k] and z == zs[k] and np.array_equal(b, boards[k]) and np.array_equal(pi, pis[k]):
                hit = key
                break
        assert hit is not None, k
        used.add(hit)
    assert len(used) == total
    sb, stp, spi, sz, idx = eng.replay_sample(8, seed=3)
    assert len(set(idx.tolist())) == 8 and idx.min() >= 0 and idx.max() < total
    info = eng.replay_info()
    assert info["total"] == total and info["capacity"] == 500000 and info["last_gather_bytes"] == total * info["tuple_bytes"]
    assert idx.tolist() == orp.sample_indices(total, info["capacity"], 8, seed=3)       # the device draw = the oracle's spec
    full = eng.replay_sample(total, seed=9)[4]
    assert sorted(full.tolist()) == list(range(total)) and full.tolist() == orp.sample_indices(total, 500000, total, seed=9)
    for j, i in enumerate(idx):
        assert np.array_equal(sb[j], boards[i]) and np.array_equal(spi[j], pis[i]) and sz[j] == zs[i] and stp[j] == tp[i]
    # the history variant returns the same draw with the 8 boards get_feats needs: board k = the position k plies earlier in the
    # same game (empty before the first move)
    hb, htp, hpi, hz, hidx = eng.replay_sample_hist(8, seed=3)
    assert np.array_equal(hidx, idx) and np.array_equal(hb[:, 0], sb) and np.array_equal(hpi, spi)
    by_board = {}
    for r in recs:
        pos, hist = ogo.GoPosition(oenv), []
        for t, m in enumerate(r.moves):
            cur = pos.board.flatten(order="F").copy()
            stack = [cur] + hist[::-1][:7]
            while len(stack) < 8:
                stack.append(stack[-1] if t > 7 else np.zeros_like(cur))
            by_board[(r.game_id, t)] = np.stack(stack[:8])
            hist.append(cur)
            pos = ogo.play_move(pos, ogo.from_flat(int(m), oenv))
    for j in range(8):
        assert any(np.array_equal(hb[j], want) for want in by_board.values()), j
    with pytest.raises(agz.AgzError):
        eng.replay_sample(total + 1)


@pytest.mark.gpu
def test_replay_ring_trims_oldest():
    """memory_size semantics (src/train.jl:52,63-65): once more tuples than the ring holds have been appended, the oldest are gone,
    the newest `cap` are readable in order, and sampling draws only from them."""
    eng = agz.Engine(9, lib_path=lib_for("cuda"), n_games=4, readouts=8, seed=31, options={"replay.capacity": 150})
    assert eng.get_option("replay.capacity") == 150
    eng.set_dummy_evaluator(None, 0.0)
    eng.selfplay_start(8)
    seen, total = [], 0
    for _ in range(2000):
        pr = eng.selfplay_step(16)
        before = total
        total = eng.replay_gather()
        recs = sorted(eng.selfplay_harvest(16), key=lambda r: r.game_id)
        if total > before:
            seen.append((before, total))
        if pr.games_finished == 8:
            break
    assert total > 300                                     # the ring (150) has wrapped at least once
    with pytest.raises(agz.AgzError):
        eng.replay_read(0, 1)                              # trimmed
    with pytest.raises(agz.AgzError):
        eng.replay_read(total - 150 - 1, 1)
    boards, tp, pis, zs = eng.replay_read(total - 150, 150)
    assert np.allclose(pis.sum(axis=1), 1.0, atol=1e-5) and set(np.unique(zs)) <= {-1, 0, 1} and set(np.unique(tp)) <= {-1, 1}
    sb, stp, spi, sz, idx = eng.replay_sample(100, seed=1)
    assert idx.tolist() == orp.sample_indices(total, 150, 100, seed=1)
    assert idx.min() >= total - 150 and idx.max() < total and len(set(idx.tolist())) == 100
    for j in range(0, 100, 9):
        k = int(idx[j] - (total - 150))
        assert np.array_equal(sb[j], boards[k]) and np.array_equal(spi[j], pis[k])
    with pytest.raises(agz.AgzError):
        eng.replay_sample(151)
    eng.close()


@pytest.mark.gpu
def test_reciprocal_quotients_equal_ieee_divisions_on_device():
    """select_leaf's score divides by 1 + N(child) through a table of correctly rounded reciprocals (fp32: one fp64 product; fp64:
    Markstein's fused correction).  3 x 10^8 random (numerator, divisor) pairs against __fdiv_rn / __ddiv_rn: no bit may differ."""
    eng = agz.Engine(9, lib_path=lib_for("cuda"), n_games=1, readouts=1600)
    bad32, bad64 = eng.selftest_division(100_000_000, seed=1)
    assert (bad32, bad64) == (0, 0)
    eng.close()
    eng = agz.Engine(19, lib_path=lib_for("cuda"), n_games=1, readouts=800)   # 412 000-entry table
    assert eng.selftest_division(50_000_000, seed=2) == (0, 0)
    eng.close()
