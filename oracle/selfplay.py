"""Self-play driver: restatement of src/selfplay.jl.  TEST INFRASTRUCTURE."""
import numpy as np
from . import game, go, rng
from . import mcts as M
from . import mcts_play as P


def resign_threshold_for(seed, game_id, base=-0.9, disable_frac=0.05):
    """selfplay.jl:9 with the evident intent `rand() < 0.05 ? -1.0 : -0.9` (the committed line is a typo)."""
    r = rng.draw(seed, game_id, rng.SITE_RESIGN, 0)
    return -1.0 if rng.u53(r[0], r[1]) < disable_frac else base


def selfplay(env, nn, num_ro=800, seed=0, game_id=0, resign_threshold=-0.9, resign_disable_frac=0.05,
             on_move=None):                              # selfplay.jl:1-45
    thr = resign_threshold_for(seed, game_id, resign_threshold, resign_disable_frac)
    player = P.MCTSPlayer(env, nn, num_readouts=num_ro, resign_threshold=thr, seed=seed, game_id=game_id)
    readouts = player.num_readouts
    P.initialize_game(player)
    first_node = M.select_leaf(player.root)
    prob, val = nn([first_node.position])
    M.incorporate_results(first_node, np.asarray(prob[:, 0]), val[0], first_node)
    while True:
        M.inject_noise(player.root)
        current_readouts = player.root.N
        while player.root.N < current_readouts + readouts:
            P.tree_search(player)
        if P.should_resign(player):
            P.set_result(player, -player.root.position.to_play, True)
            break
        move = P.pick_move(player)
        P.play_move(player, move)
        if on_move is not None:
            on_move(player, move)
        if M.is_done(player.root):
            P.set_result(player, game.result(player.root.position), False)
            break
    return player
