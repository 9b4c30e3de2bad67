"""The reference's test/test_go.jl cases driven through the C ABI position hooks (device Go rules), plus
randomised agreement with the oracle's liberty tracker."""
import numpy as np
import pytest

from backends import BACKENDS, agz, lib_for
from oracle import go as ogo
from refboards import load_board, pc_set, EMPTY_ROW9 as EMPTY_ROW
from test_oracle_go import SUICIDE_BOARD, LEGAL_BOARD, SCORE_BOARD_1, SCORE_BOARD_2

BLACK, WHITE = 1, -1


@pytest.fixture(scope="module", params=BACKENDS)
def env(request):
    return agz.GoEnv(9, lib_path=lib_for(request.param))


def kgs(s, env):
    return agz.from_kgs(s, env)


def pos_of(env, rows, **kw):
    return agz.GoPosition(env, board=load_board(rows, env), **kw)


def test_liberty_cache_and_merge(env):                   # test_go.jl:74-139
    p = pos_of(env, "X........" + EMPTY_ROW * 8)
    assert agz.api.liberties(p)[kgs("A9", env)] == 2
    p2 = agz.play_move(p, kgs("B9", env))
    lib = agz.api.liberties(p2)
    assert lib[kgs("A9", env)] == 3 and lib[kgs("B9", env)] == 3
    pw = agz.play_move(agz.GoPosition(env, board=p.board, to_play=WHITE), kgs("B9", env))
    lib = agz.api.liberties(pw)
    assert lib[kgs("A9", env)] == 1 and lib[kgs("B9", env)] == 2
    m = pos_of(env, ".X.......\nX.X......\n.X.......\n" + EMPTY_ROW * 6)
    lib = agz.api.liberties(agz.play_move(m, kgs("B8", env)))
    for s in pc_set("B9 A8 B8 C8 B7", env):
        assert lib[s] == 6


def test_captures(env):                                  # test_go.jl:141-226
    p = pos_of(env, ".X.......\nXO.......\n.X.......\n" + EMPTY_ROW * 6)
    q = agz.play_move(p, kgs("C8", env))
    assert q.board[kgs("B8", env)] == 0 and q.caps == (1, 0)
    p = pos_of(env, ".XX......\nXOO......\n.XX......\n" + EMPTY_ROW * 6)
    q = agz.play_move(p, kgs("D8", env))
    assert q.board[kgs("B8", env)] == 0 and q.board[kgs("C8", env)] == 0 and q.caps == (2, 0)
    lib = agz.api.liberties(q)
    for s, n in (("B9", 4), ("C9", 4), ("A8", 3), ("D8", 4), ("B7", 6), ("C7", 6), ("B8", 0), ("C8", 0)):
        assert lib[kgs(s, env)] == n, s
    p = pos_of(env, ".OX......\nOXX......\nXX.......\n" + EMPTY_ROW * 6)
    q = agz.play_move(p, kgs("A9", env))
    assert q.caps == (2, 0) and q.board[kgs("B9", env)] == 0 and q.board[kgs("A8", env)] == 0
    lib = agz.api.liberties(q)
    assert lib[kgs("A9", env)] == 2
    for s in pc_set("C9 B8 C8 A7 B7", env):
        assert lib[s] == 7


def test_same_group_neighboring_twice(env):              # test_go.jl:228-262
    p = pos_of(env, "XX.......\nX........\n" + EMPTY_ROW * 7)
    lib = agz.api.liberties(agz.play_move(p, kgs("B8", env)))
    for s in pc_set("A9 B9 A8 B8", env):
        assert lib[s] == 4
    pw = agz.GoPosition(env, board=p.board, to_play=WHITE)
    q = agz.play_move(pw, kgs("B8", env))
    lib = agz.api.liberties(q)
    assert lib[kgs("A9", env)] == 2 and lib[kgs("B8", env)] == 2 and q.caps == (0, 0)


def test_passing_and_game_over(env):                     # test_go.jl:264-285, 509-516
    tb = load_board(".X.....OO\nX........\n" + EMPTY_ROW * 7, env)
    start = agz.GoPosition(env, board=tb, n=0, komi=6.5, caps=(1, 2), ko=kgs("A1", env), to_play=BLACK)
    p = agz.api.pass_move(start)
    assert (p.board == tb).all() and p.n == 1 and p.caps == (1, 2) and p.ko is None and p.to_play == WHITE
    assert p.recent == [(BLACK, None)]
    root = agz.GoPosition(env)
    first = agz.play_move(root, None)
    assert not first.done
    assert agz.play_move(first, None).done


def test_is_move_suicidal_and_legal(env):                # test_go.jl:310-378
    pos = pos_of(env, SUICIDE_BOARD, to_play=BLACK)
    legal = agz.all_legal_moves(pos)
    for mv in pc_set("E9 H5", env):
        assert not legal[agz.to_flat(mv, env)]
    for mv in pc_set("B5 J1 A9", env):
        assert legal[agz.to_flat(mv, env)]
    board = load_board(LEGAL_BOARD, env)
    oenv = ogo.GoEnv(9)
    for b, tp in ((board, BLACK), (-board, WHITE)):
        pos = agz.GoPosition(env, board=b, to_play=tp)
        legal = agz.all_legal_moves(pos)
        for mv in pc_set("A9 E9 J9", env):
            assert not legal[agz.to_flat(mv, env)]
        for mv in pc_set("A4 G1 J1 H7", env):
            assert legal[agz.to_flat(mv, env)]
        assert (legal == ogo.all_legal_moves(ogo.GoPosition(oenv, board=b.copy(), to_play=tp))).all()
        assert legal[-1] == 1


def test_move_and_capture(env):                          # test_go.jl:380-456
    tb = load_board(".X.....OO\nX........\n" + EMPTY_ROW * 7, env)
    start = agz.GoPosition(env, board=tb, n=0, komi=6.5, caps=(1, 2), to_play=BLACK)
    a1 = agz.play_move(start, kgs("C9", env))
    assert (a1.board == load_board(".XX....OO\nX........\n" + EMPTY_ROW * 7, env)).all()
    assert a1.n == 1 and a1.caps == (1, 2) and a1.ko is None and a1.to_play == WHITE and a1.recent == [(BLACK, kgs("C9", env))]
    a2 = agz.play_move(a1, kgs("J8", env))
    assert (a2.board == load_board(".XX....OO\nX.......O\n" + EMPTY_ROW * 7, env)).all()
    assert a2.n == 2 and a2.to_play == BLACK
    sb = load_board(EMPTY_ROW * 5 + "XXXX.....\nXOOX.....\nO.OX.....\nOOXX.....\n", env)
    q = agz.play_move(agz.GoPosition(env, board=sb, komi=6.5, caps=(1, 2), to_play=BLACK), kgs("B2", env))
    assert (q.board == load_board(EMPTY_ROW * 5 + "XXXX.....\nX..X.....\n.X.X.....\n..XX.....\n", env)).all()
    assert q.caps == (7, 2) and q.ko is None and q.n == 1


def test_ko_move(env):                                   # test_go.jl:458-507
    sb = load_board(".OX......\nOX.......\n" + EMPTY_ROW * 7, env)
    start = agz.GoPosition(env, board=sb, komi=6.5, caps=(1, 2), to_play=BLACK)
    a = agz.play_move(start, kgs("A9", env))
    assert (a.board == load_board("X.X......\nOX.......\n" + EMPTY_ROW * 7, env)).all()
    assert a.caps == (2, 2) and a.ko == kgs("B9", env) and a.to_play == WHITE
    with pytest.raises(agz.IllegalMove):
        agz.play_move(a, kgs("B9", env))
    retake = agz.play_move(agz.api.pass_move(agz.api.pass_move(a)), kgs("B9", env))
    assert (retake.board == sb).all() and retake.n == 4 and retake.caps == (2, 3) and retake.ko == kgs("A9", env)
    with pytest.raises(agz.IllegalMove):                  # occupied point
        agz.play_move(start, kgs("B9", env))


def test_scoring(env):                                   # test_go.jl:518-564
    assert agz.score(pos_of(env, SCORE_BOARD_1, n=54, komi=6.5, to_play=BLACK)) == 1.5
    assert agz.score(pos_of(env, SCORE_BOARD_2, n=55, komi=6.5, to_play=WHITE)) == 2.5
    assert agz.result(pos_of(env, SCORE_BOARD_2, n=55, komi=6.5)) == 1


def test_features_hook(env):                             # test_features.jl:48-79 (position-level get_feats)
    pos = agz.GoPosition(env)
    for c in ((0, 0), (0, 1), (0, 2), (0, 3), (1, 1)):
        pos = agz.play_move(pos, c)
    lb = lambda two_rows: load_board(two_rows + EMPTY_ROW * 7, env)
    if env.lib_path is not None:
        pytest.skip("agz_features (host-supplied positions) is a CUDA-library entry point")
    f = agz.get_feats(pos)
    assert pos.to_play == WHITE and f.shape == (9, 9, 17)
    assert (f[:, :, 0] == lb("...X.....\n.........\n")).all()
    assert (f[:, :, 1] == lb("X.X......\n.X.......\n")).all()
    assert (f[:, :, 2] == lb(".X.X.....\n.........\n")).all()
    assert (f[:, :, 3] == lb("X.X......\n.........\n")).all()
    assert (f[:, :, 4] == lb(".X.......\n.........\n")).all()
    assert (f[:, :, 5] == lb("X.X......\n.........\n")).all()
    assert (f[:, :, 10:16] == 0).all() and (f[:, :, 16] == -1).all()


@pytest.mark.parametrize("N", [5, 9, 13, 19])   # 5: one bit-plane word, 13: the 6-word kernel variant, 19: 12 words
def test_random_playouts_match_oracle(request, N):
    """Random legal playouts: board, ko, captures, legal mask, liberties and score equal the oracle's at every ply."""
    backend = "emu"
    e = agz.GoEnv(N, lib_path=lib_for(backend))
    _random_playouts(e, N, games=2 if N <= 9 else 1, plies={5: 60, 9: 120, 13: 120, 19: 150}[N])


@pytest.mark.gpu
@pytest.mark.parametrize("N", [5, 9, 13, 19])
def test_random_playouts_match_oracle_cuda(N):
    e = agz.GoEnv(N, lib_path=lib_for("cuda"))
    _random_playouts(e, N, games=3, plies={5: 80, 9: 200, 13: 300, 19: 400}[N])


def _random_playouts(env, N, games, plies):
    oenv = ogo.GoEnv(N)
    rs = np.random.RandomState(N)
    for g in range(games):
        op = ogo.GoPosition(oenv)
        gp = agz.GoPosition(env)
        for t in range(plies):
            ol = ogo.all_legal_moves(op)
            gl = agz.all_legal_moves(gp)
            assert (ol == gl).all(), (g, t)
            cand = np.flatnonzero(ol[:-1])
            if len(cand) == 0 or rs.rand() < 0.03:
                mv = None
            else:
                mv = ogo.from_flat(int(rs.choice(cand)), oenv)
            op = ogo.play_move(op, mv)
            gp = agz.play_move(gp, mv)
            assert (op.board == gp.board).all(), (g, t)
            assert op.ko == gp.ko and op.caps == gp.caps and op.n == gp.n and op.to_play == gp.to_play
            if t % 10 == 0:
                assert (op.lib_tracker.liberty_cache == agz.api.liberties(gp)).all()
                assert float(ogo.score(op)) == agz.score(gp)
            if op.done:
                assert gp.done
                break


@pytest.mark.parametrize("N,games,plies", [(5, 12, 250), (7, 6, 250)])
def test_small_board_capture_stress(N, games, plies):
    """Long random games on small boards: thousands of captured stones, ko positions and enclosed (suicide-candidate) points, the
    regime where the legal-mask code visits groups one by one.  Legal mask, board, ko, captures at every ply; score every 7."""
    e = agz.GoEnv(N, lib_path=lib_for("emu"))
    oenv = ogo.GoEnv(N)
    rs = np.random.RandomState(100 + N)
    captured = kos = enclosed = 0
    for g in range(games):
        op, gp = ogo.GoPosition(oenv), agz.GoPosition(e)
        for t in range(plies):
            ol, gl = ogo.all_legal_moves(op), agz.all_legal_moves(gp)
            assert (ol == gl).all(), (g, t)
            enclosed += int(((op.board.flatten(order="F") == 0) & (ol[:-1] == 0)).sum())
            cand = np.flatnonzero(ol[:-1])
            mv = None if (len(cand) == 0 or rs.rand() < 0.02) else ogo.from_flat(int(rs.choice(cand)), oenv)
            c0 = sum(op.caps)
            op, gp = ogo.play_move(op, mv), agz.play_move(gp, mv)
            captured += sum(op.caps) - c0
            kos += int(op.ko is not None)
            assert (op.board == gp.board).all() and op.ko == gp.ko and op.caps == gp.caps, (g, t)
            if t % 7 == 0:
                assert float(ogo.score(op)) == agz.score(gp)
            if op.done:
                break
    assert captured > 300 and enclosed > 300 and (N != 5 or kos > 0)
