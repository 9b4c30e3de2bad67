"""Gomoku restatement (oracle/gomoku.py) against the behaviour of src/game/gomoku/board.jl.  The reference ships no tests for
this game, so the cases are hand-checked against its code: has_game_ended :97-129 (four directions, the scan order, full board),
play_move! :137-169, all_legal_moves :95, result / result_string :173-193, replay_position :195-217, coords.jl:6-12."""
import numpy as np
import pytest

from oracle import game as G
from oracle import gomoku as gm
from oracle import mcts as M
from oracle import selfplay as osp

B, W, E = gm.BLACK, gm.WHITE, gm.EMPTY


def board_from(rows):
    m = {"X": B, "O": W, ".": E}
    return np.array([[m[ch] for ch in r] for r in rows], dtype=np.int8)


def test_env_and_coords():
    env = gm.GomokuEnv()
    assert (env.N, env.n_in_row, env.action_space, env.planes, env.max_action_space) == (15, 5, 225, 8, 361)   # gomoku.jl:9-17
    env = gm.GomokuEnv(9, 4)
    assert gm.to_flat((2, 3), env) == 9 * 3 + 2 and gm.from_flat(29, env) == (2, 3)
    assert gm.to_flat(None, env) == 81 and gm.from_flat(81, env) is None       # coords.jl:6,11 (not an action of this game)


@pytest.mark.parametrize("cells,winner", [
    ([(3, c) for c in range(1, 6)], B),                      # a row
    ([(r, 4) for r in range(2, 7)], B),                      # a column
    ([(1 + k, 2 + k) for k in range(5)], B),                 # diagonal down-right
    ([(1 + k, 7 - k) for k in range(5)], B),                 # diagonal down-left
    ([(0, c) for c in range(0, 6)], B),                      # an overline still holds a window of five
])
def test_five_in_a_row_ends_the_game(cells, winner):
    env = gm.GomokuEnv(9, 5)
    for color in (B, W):
        b = env.empty_board()
        for c in cells:
            b[c] = color
        assert gm.has_game_ended(b, env) == (True, winner * color)
        pos = gm.GomokuPosition(env, board=b)
        assert pos.done and pos.winner == winner * color and gm.result(pos) == winner * color
        assert gm.result_string(pos) == ("B" if winner * color > 0 else "W")
    b = env.empty_board()
    for c in cells[:4]:
        b[c] = B
    assert gm.has_game_ended(b, env) == (False, E)           # four are not enough
    b[cells[4]] = W
    assert gm.has_game_ended(b, env) == (False, E)           # mixed colours


def test_full_board_is_a_draw():
    env = gm.GomokuEnv(4, 4)
    b = board_from(["XXOO", "OOXX", "XXOO", "OOXX"])
    assert gm.has_game_ended(b, env) == (True, E)
    pos = gm.GomokuPosition(env, board=b)
    assert pos.done and gm.result(pos) == 0 and gm.result_string(pos) == "DRAW"
    b2 = b.copy()
    b2[3, 3] = E
    assert gm.has_game_ended(b2, env) == (False, E)


def test_play_move_and_legal_moves():
    env = gm.GomokuEnv(5, 3)
    pos = gm.GomokuPosition(env)
    assert gm.all_legal_moves(pos).shape == (25,) and gm.all_legal_moves(pos).all()
    p1 = gm.play_move(pos, (1, 2))
    assert pos.board.sum() == 0 and pos.n == 0                # play_move! copies unless mutate
    assert p1.board[1, 2] == B and p1.to_play == W and p1.n == 1 and not p1.done
    assert p1.recent[-1].color == B and p1.recent[-1].move == (1, 2)
    assert p1.board_deltas.shape == (1, 5, 5) and p1.board_deltas[0, 1, 2] == B
    legal = gm.all_legal_moves(p1)
    assert legal.sum() == 24 and legal[gm.to_flat((1, 2), env)] == 0          # column-major vec (board.jl:95)
    with pytest.raises(gm.IllegalMove):
        gm.play_move(p1, (1, 2))
    p = p1
    for c in [(0, 0), (2, 2), (0, 1), (3, 2)]:               # B completes (1,2) (2,2) (3,2)
        p = gm.play_move(p, c)
    assert p.done and p.winner == B
    with pytest.raises(AssertionError):
        gm.play_move(p, (4, 4))                               # @assert !new_pos.done (board.jl:144)
    q = gm.GomokuPosition(gm.GomokuEnv(5, 5))
    for k in range(9):
        q = gm.play_move(q, (k // 5, k % 5))
    assert q.board_deltas.shape[0] == 7 and q.n == 9 and not q.done           # 7 deltas kept (board.jl:163-164)


def test_replay_position():
    env = gm.GomokuEnv(5, 4)
    p = gm.GomokuPosition(env)
    moves = [(0, 0), (1, 1), (0, 1), (2, 2), (0, 2)]
    for c in moves:
        p = gm.play_move(p, c)
    ctx = gm.replay_position(p, 1)
    assert [x.next_move for x in ctx] == moves and all(x.result == 1 for x in ctx)
    assert ctx[0].position.board.sum() == 0 and ctx[3].position.n == 3
    p.n += 1
    with pytest.raises(AssertionError):
        gm.replay_position(p, 1)


def test_search_has_no_pass_hack_and_ends_on_a_win():
    """mcts.jl:121 guards the pass hack with typeof(env) == GoEnv; terminal leaves are backed up with result(pos)."""
    env = gm.GomokuEnv(5, 3)

    class Net:
        def __call__(self, positions):
            n = len(positions)
            return np.full((25, n), 1 / 25, np.float32), np.zeros(n, np.float32)

    pl = osp.selfplay(env, Net(), 16, seed=3, game_id=0)
    assert pl.root.position.done and pl.result == gm.result(pl.root.position) and pl.result_string in ("B", "W", "DRAW")
    assert all(pm.move is not None for pm in pl.root.position.recent)
    assert all(p.shape == (25,) for p in pl.searches_pi)
    assert M.MCTSRules(env).dirichlet_noise_alpha == np.float32(0.03 * 361 / 25)
    assert isinstance(G.Position(env), gm.GomokuPosition)
