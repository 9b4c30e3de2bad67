"""Search player: restatement of src/mcts_play.jl.  TEST INFRASTRUCTURE."""
import numpy as np
from . import game, go, rng
from . import mcts as M

f32 = np.float32


class MCTSPlayer:                                        # mcts_play.jl:3-24
    def __init__(self, env, network, num_readouts=800, two_player_mode=False, resign_threshold=-0.9,
                 seed=0, game_id=0):
        self.env = env
        self.network = network
        self.num_readouts = num_readouts
        self.two_player_mode = two_player_mode
        self.tau_threshold = -1 if two_player_mode else (env.N * env.N // 12) // 2 * 2
        self.qs = []
        self.searches_pi = []
        self.searches_N = []      # raw visit counts per move (oracle extra, for parity checks)
        self.result = 0
        self.result_string = ""
        self.root = None
        self.resign_threshold = resign_threshold
        self.position = None
        self.rng = M.RngCtx(seed, game_id)


def play_move(player, c):                                # mcts_play.jl:26-50
    root = player.root
    if not player.two_player_mode:
        player.searches_pi.append(M.children_as_pi(root, root.position.n <= player.tau_threshold).astype(f32))
        player.searches_N.append(root.child_N.copy())
    player.qs.append(f32(root.Q))
    try:
        player.root = M.maybe_add_child(root, game.to_flat(c, root.position.env))
    except go.IllegalMove:
        if not player.two_player_mode:
            player.searches_pi.pop()
            player.searches_N.pop()
        player.qs.pop()
        return False
    player.position = player.root.position
    player.root.parent.children = {}                     # siblings dropped; the chosen subtree is kept
    player.rng.reset_root()
    return True


def pick_move(player):                                   # mcts_play.jl:52-71
    root = player.root
    ctx = player.rng
    if root.position.n >= player.tau_threshold:
        max_val = root.child_N.max()
        possible = np.flatnonzero(root.child_N == max_val)
        r0 = rng.draw(ctx.seed, ctx.game_id, rng.SITE_PICK_MAX, root.position.n)[0]
        fcoord = int(possible[rng.mulhi(r0, len(possible))])
    else:
        cdf = np.cumsum(root.child_N, dtype=f32)
        cdf = (cdf / cdf[-2]).astype(f32)                # prevents passing via softpick
        r = rng.draw(ctx.seed, ctx.game_id, rng.SITE_PICK_SOFT, root.position.n)
        selection = rng.u53(r[0], r[1])
        fcoord = int(np.searchsorted(cdf.astype(np.float64), selection, side="left"))
        assert root.child_N[fcoord] != 0
    return game.from_flat(fcoord, root.position.env)


def tree_search(player, parallel_readouts=8):            # mcts_play.jl:73-98
    leaves = []
    failsafe = 0
    while len(leaves) < parallel_readouts and failsafe < 2 * parallel_readouts:
        failsafe += 1
        leaf = M.select_leaf(player.root)
        if M.is_done(leaf):
            value = game.result(leaf.position)
            M.backup_value(leaf, value, player.root)
            continue
        M.add_virtual_loss(leaf, player.root)
        leaves.append(leaf)
    if leaves:
        move_probs, values = player.network([leaf.position for leaf in leaves])
        for k, leaf in enumerate(leaves):
            M.revert_virtual_loss(leaf, player.root)
            M.incorporate_results(leaf, np.asarray(move_probs[:, k]), values[k], player.root)
    return leaves


def set_result(player, winner, was_resign):              # mcts_play.jl:100-108
    player.result = winner
    if was_resign:
        s = "B+R" if winner == go.BLACK else "W+R"
    else:
        s = game.result_string(player.root.position)
    player.result_string = s


def initialize_game(player, pos=None):                   # mcts_play.jl:110-118
    if pos is None:
        pos = game.Position(player.env)
    player.rng.reset_root()
    player.root = M.MCTSNode(pos, rng_ctx=player.rng)
    player.result = 0
    player.searches_pi = []
    player.searches_N = []
    player.qs = []


def is_done(player):                                     # mcts_play.jl:120
    return player.result != 0 or M.is_done(player.root)


def should_resign(player):                               # mcts_play.jl:124  (Float32 < Float64)
    return float(player.root.Q_perspective) < player.resign_threshold


def extract_data(player):                                # mcts_play.jl:126-139
    assert len(player.searches_pi) == player.root.position.n
    positions, results = [], []
    pis = [p.copy() for p in player.searches_pi]
    for pwc in game.replay_position(player.root.position, player.result):
        positions.append(pwc.position)
        results.append(pwc.result)
    return positions, pis, results


def suggest_move(player):                                # mcts_play.jl:144-151
    current = player.root.N
    while player.root.N < current + player.num_readouts:
        tree_search(player)
    return pick_move(player)
