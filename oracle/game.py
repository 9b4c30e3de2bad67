"""The reference's `Position` interface (src/game/env.jl, src/AlphaGo.jl:10): the search code calls these generic functions and
Julia dispatches on the position type -- GoPosition (oracle/go.py) or GomokuPosition (oracle/gomoku.py).  TEST INFRASTRUCTURE."""
from . import go, gomoku


def _mod(env):
    return gomoku if isinstance(env, gomoku.GomokuEnv) else go


def is_go(env):                                          # `typeof(env) == GoEnv` (mcts.jl:121)
    return isinstance(env, go.GoEnv)


def Position(env):                                       # env.jl:1,4
    return gomoku.GomokuPosition(env) if isinstance(env, gomoku.GomokuEnv) else go.GoPosition(env)


def play_move(pos, c, color=None):
    return _mod(pos.env).play_move(pos, c, color=color)


def all_legal_moves(pos):
    return _mod(pos.env).all_legal_moves(pos)


def result(pos):
    return _mod(pos.env).result(pos)


def result_string(pos):
    return _mod(pos.env).result_string(pos)


def replay_position(pos, result):
    return _mod(pos.env).replay_position(pos, result)


def to_flat(coord, env):
    return _mod(env).to_flat(coord, env)


def from_flat(f, env):
    return _mod(env).from_flat(f, env)
