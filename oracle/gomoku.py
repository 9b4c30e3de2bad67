"""Gomoku rules: restatement of src/game/gomoku/{gomoku,board,coords}.jl of the reference.  TEST INFRASTRUCTURE.

The second game behind the reference's `Position` interface (src/game/env.jl): action_space = N^2 (no pass), every empty
point is legal, the game ends when `n_in_row` stones of one colour line up (row, column or either diagonal) or the board is
full (draw).  0-based indices as in oracle/go.py: board[i, j], flat move f = N*j + i.  The reference ships no tests for this
game; tests/test_oracle_gomoku.py pins the restatement on hand-checked positions.
"""
import numpy as np

from .go import BLACK, EMPTY, WHITE, IllegalMove, PlayerMove, PositionWithContext, get_first_n  # noqa: F401  (shared definitions)


class GomokuEnv:                                         # gomoku.jl:1-19
    def __init__(self, board_size=15, connect_row=5, planes=17):
        N = board_size
        self.N = N
        self.n_in_row = connect_row
        self.action_space = N * N
        assert planes % 2 == 1
        self.planes = (planes - 1) // 2
        self.max_action_space = 361

    def empty_board(self):
        return np.zeros((self.N, self.N), dtype=np.int8)


def to_flat(coord, env):                                 # coords.jl:6-7 (nothing -> N^2, which is not an action of this game)
    return env.N * env.N if coord is None else env.N * coord[1] + coord[0]


def from_flat(f, env):                                   # coords.jl:10-12
    if f == env.N * env.N:
        return None
    j, i = divmod(f, env.N)
    return (i, j)


def has_game_ended(board, env):                          # board.jl:97-129: (done, winner); scan order h, w as the reference
    k, dim = env.n_in_row, env.N
    for h in range(dim):
        for w in range(dim):
            if board[h, w] == EMPTY:
                continue
            if w <= dim - k and len({int(board[h, i]) for i in range(w, w + k)}) == 1:
                return True, int(board[h, w])
            if h <= dim - k and len({int(board[i, w]) for i in range(h, h + k)}) == 1:
                return True, int(board[h, w])
            if w <= dim - k and h <= dim - k and len({int(board[h + i, w + i]) for i in range(k)}) == 1:
                return True, int(board[h, w])
            if w >= k - 1 and h <= dim - k and len({int(board[h + i, w - i]) for i in range(k)}) == 1:
                return True, int(board[h, w])
    if (board == EMPTY).sum() == 0:
        return True, EMPTY
    return False, EMPTY


class GomokuPosition:                                    # board.jl:25-58
    def __init__(self, env, board=None, n=0, recent=None, board_deltas=None, to_play=BLACK):
        self.env = env
        self.board = board if board is not None else env.empty_board()
        self.n = n
        self.recent = recent if recent is not None else []
        self.board_deltas = board_deltas if board_deltas is not None else np.zeros((0, env.N, env.N), dtype=np.int8)
        self.to_play = to_play
        self.done, self.winner = has_game_ended(self.board, env)

    def copy(self):                                      # board.jl:60-65
        return GomokuPosition(self.env, board=self.board.copy(), n=self.n, recent=list(self.recent),
                              board_deltas=self.board_deltas, to_play=self.to_play)


def is_move_legal(pos, move):                            # board.jl:93
    return pos.board[move] == EMPTY


def all_legal_moves(pos):                                # board.jl:95 (vec = column-major)
    return (pos.board == EMPTY).flatten(order="F").astype(np.int8)


def flip_playerturn(pos, mutate=False):                  # board.jl:131-135
    new_pos = pos if mutate else pos.copy()
    new_pos.to_play *= -1
    return new_pos


def play_move(pos, c, color=None, mutate=False):         # board.jl:137-169
    if color is None:
        color = pos.to_play
    new_pos = pos if mutate else pos.copy()
    assert not new_pos.done
    if c is None or not is_move_legal(pos, c):
        raise IllegalMove()
    new_pos.board[c] = color
    N = pos.env.N
    delta = np.zeros((N, N), dtype=np.int8)
    delta[c] = color
    new_pos.n += 1
    new_pos.recent.append(PlayerMove(color, c))
    new_pos.board_deltas = np.concatenate(
        [delta.reshape(1, N, N), get_first_n(new_pos.board_deltas, new_pos.env.planes - 2)], axis=0)
    new_pos.to_play *= -1
    new_pos.done, new_pos.winner = has_game_ended(new_pos.board, new_pos.env)
    return new_pos


def score(pos):                                          # board.jl:171
    return pos.winner


def result(pos):                                         # board.jl:173-182
    points = score(pos)
    return 1 if points > 0 else (-1 if points < 0 else 0)


def result_string(pos):                                  # board.jl:184-193
    points = score(pos)
    return "B" if points > 0 else ("W" if points < 0 else "DRAW")


def replay_position(pos, result):                        # board.jl:195-217
    if pos.n != len(pos.recent):
        raise AssertionError("GomokuPosition history is incomplete")
    out = []
    dummy = GomokuPosition(pos.env)
    for pm in pos.recent:
        out.append(PositionWithContext(dummy, pm.move, result))
        dummy = play_move(dummy, pm.move, color=pm.color)
    return out
