"""Gating match between two networks: restatement of `evaluate` (src/neural_net.jl:103-158).  TEST INFRASTRUCTURE.

Two MCTSPlayers in two_player_mode (tau_threshold = -1: pick_move is always arg-max, no Dirichlet noise, no searches_pi)
alternate; each keeps its own tree and both play every move.  Parity unpinned for the random tie-breaks (oracle/rng.py):
the black player draws with (seed_black, game id), the white player with (seed_white, game id).

Reference quirk kept on purpose: the win counter reads `result(black.root.position)` (:150), i.e. the area score of the
final position, even when the game ended by resignation.
"""
from . import game, go
from . import mcts as M
from . import mcts_play as P


def play_match_game(env, black_net, white_net, ro, seed_black, seed_white, game_id, resign_threshold=-0.9):
    black = P.MCTSPlayer(env, black_net, num_readouts=ro, two_player_mode=True, resign_threshold=resign_threshold,
                         seed=seed_black, game_id=game_id)
    white = P.MCTSPlayer(env, white_net, num_readouts=ro, two_player_mode=True, resign_threshold=resign_threshold,
                         seed=seed_white, game_id=game_id)
    P.initialize_game(black)
    P.initialize_game(white)
    num_move = 0
    moves = []
    while True:
        active, inactive = (white, black) if num_move % 2 == 1 else (black, white)
        current_readouts = active.root.N
        while active.root.N < current_readouts + active.num_readouts:      # :124-126
            P.tree_search(active)
        if P.should_resign(active):                                          # :129-133
            winner = -active.root.position.to_play
            P.set_result(active, winner, True)
            P.set_result(inactive, winner, True)
            break
        move = P.pick_move(active)                                           # :135-138
        P.play_move(active, move)
        P.play_move(inactive, move)
        moves.append(game.to_flat(move, env))
        num_move += 1
        if P.is_done(active):                                                # :140-146
            winner = game.result(active.root.position)
            P.set_result(active, winner, False)
            P.set_result(inactive, winner, False)
            break
    return black, white, moves


def evaluate(env, black_net, white_net, num_games=400, ro=800, seed=0, resign_threshold=-0.9, details=None):
    games_won = 0
    for i in range(num_games):
        black, white, moves = play_match_game(env, black_net, white_net, ro, seed, seed + 1, i, resign_threshold)
        won = game.result(black.root.position) == go.BLACK                     # :150
        games_won += int(won)
        if details is not None:
            details.append({"moves": moves, "result": black.result, "result_string": black.result_string, "black_won": won})
    return games_won / num_games >= 0.55, games_won
