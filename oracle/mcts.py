"""Search tree: restatement of src/mcts.jl:1-252.  TEST INFRASTRUCTURE.

Float widths follow the reference exactly: child_N / child_W / priors are
Float32; Q is Float32; U and the score comparison are Float64 because `c_puct`
is a Float64 global (mcts.jl:11, 86-92).
"""
import numpy as np
from . import game, go, rng

c_puct = 0.96                                            # mcts.jl:11
dirichlet_noise_weight = 0.25                            # mcts.jl:13
f32 = np.float32


MAX_GAME_LENGTH_OVERRIDE = None   # tests shorten 19x19 games; None = the reference's N^2*7/5


class MCTSRules:                                         # mcts.jl:15-25
    def __init__(self, env):
        self.max_game_length = MAX_GAME_LENGTH_OVERRIDE or (env.N ** 2 * 7) // 5
        self.dirichlet_noise_alpha = f32(0.03 * env.max_action_space / env.action_space)


class RngCtx:
    """Explicit replacement of Julia's global RNG state (oracle/rng.py spec)."""

    def __init__(self, seed=0, game_id=0):
        self.seed, self.game_id = seed, game_id
        self.sel_ctr = 0      # select_leaf calls since the root last changed
        self.noise_ctr = 0    # inject_noise! calls since the root last changed

    def reset_root(self):
        self.sel_ctr = 0
        self.noise_ctr = 0


class DummyNode:                                         # mcts.jl:27-39
    def __init__(self):
        self.parent = None
        self.child_N = {None: f32(0)}
        self.child_W = {None: f32(0)}


class MCTSNode:                                          # mcts.jl:41-82
    def __init__(self, position, fmove=None, parent=None, rng_ctx=None):
        if parent is None:
            parent = DummyNode()
            self.rng = rng_ctx if rng_ctx is not None else RngCtx()
        else:
            self.rng = parent.rng
        A = position.env.action_space
        self.parent = parent
        self.fmove = fmove
        self.position = position
        self.is_expanded = False
        self.losses_applied = 0
        self.child_N = np.zeros(A, dtype=f32)
        self.child_W = np.zeros(A, dtype=f32)
        self.original_prior = np.zeros(A, dtype=f32)
        self.child_prior = np.zeros(A, dtype=f32)
        self.children = {}
        self.mcts_rules = MCTSRules(position.env)

    # N / W live in the parent's arrays (mcts.jl:94-102)
    @property
    def N(self):
        return self.parent.child_N[self.fmove]

    @N.setter
    def N(self, v):
        self.parent.child_N[self.fmove] = f32(v)

    @property
    def W(self):
        return self.parent.child_W[self.fmove]

    @W.setter
    def W(self, v):
        self.parent.child_W[self.fmove] = f32(v)

    @property
    def Q(self):                                         # mcts.jl:94
        return f32(self.W / f32(f32(1) + self.N))

    @property
    def Q_perspective(self):                             # mcts.jl:105
        return f32(self.Q * f32(self.position.to_play))


def legal_moves(x):                                      # mcts.jl:84
    return game.all_legal_moves(x.position)


def child_Q(x):                                          # mcts.jl:89 (Float32)
    return (x.child_W / (f32(1) + x.child_N)).astype(f32)


def child_U(x):                                          # mcts.jl:91-92 (Float64)
    s = np.sqrt(f32(f32(1) + x.N), dtype=f32)            # sqrt in Float32
    return ((c_puct * np.float64(s)) * x.child_prior.astype(np.float64)) / (f32(1) + x.child_N).astype(np.float64)


def child_action_score(x):                               # mcts.jl:86-87
    return (child_Q(x) * f32(x.position.to_play)).astype(np.float64) + child_U(x)


def select_leaf(root):                                   # mcts.jl:108-138
    current = root
    N2 = root.position.env.N ** 2
    pass_move = N2
    ctx = root.rng
    sel_idx = ctx.sel_ctr
    ctx.sel_ctr += 1
    depth = 0
    move_no = root.position.n
    while True:
        current.N = current.N + f32(1)
        if not current.is_expanded:
            break
        pos = current.position
        if game.is_go(pos.env) and len(pos.recent) != 0 and pos.recent[-1].move is None and current.child_N[pass_move] == 0:   # Go only (mcts.jl:121)
            current = maybe_add_child(current, pass_move)
            depth += 1
            continue
        cas = child_action_score(current)
        legal = legal_moves(current).astype(bool)
        max_score = cas[legal].max()
        possible = np.flatnonzero(legal & (cas == max_score))
        r0 = rng.draw(ctx.seed, ctx.game_id, rng.SITE_SELECT, move_no, sel_idx, depth)[0]
        best = int(possible[rng.mulhi(r0, len(possible))])
        current = maybe_add_child(current, best)
        depth += 1
    return current


def maybe_add_child(node, fcoord):                       # mcts.jl:140-147
    if fcoord not in node.children:
        new_pos = game.play_move(node.position, game.from_flat(fcoord, node.position.env))
        node.children[fcoord] = MCTSNode(new_pos, fcoord, node)
    return node.children[fcoord]


def add_virtual_loss(node, up_to):                       # mcts.jl:149-163
    while True:
        node.losses_applied += 1
        node.W = node.W + f32(node.position.to_play)
        if node.parent is None or node is up_to:
            return
        node = node.parent


def revert_virtual_loss(node, up_to):                    # mcts.jl:165-171
    while True:
        node.losses_applied -= 1
        node.W = node.W + f32(-node.position.to_play)
        if node.parent is None or node is up_to:
            return
        node = node.parent


def revert_visits(node, up_to):                          # mcts.jl:173-186
    while True:
        node.N = node.N - f32(1)
        if node.parent is None or node is up_to:
            return
        node = node.parent


def backup_value(node, value, up_to):                    # mcts.jl:215-225
    value = f32(value)
    while True:
        node.W = node.W + value
        if node.parent is None or node is up_to:
            return
        node = node.parent


def incorporate_results(node, move_probs, value, up_to):  # mcts.jl:188-213
    assert move_probs.shape == (node.position.env.action_space,)
    assert not node.position.done
    if node.is_expanded:
        revert_visits(node, up_to)
        return
    node.is_expanded = True
    node.child_prior[:] = move_probs
    node.original_prior[:] = node.child_prior
    node.child_W[:] = f32(value)
    backup_value(node, value, up_to)


def is_done(node):                                       # mcts.jl:230-231
    return node.position.done or node.position.n >= node.mcts_rules.max_game_length


def inject_noise(node):                                  # mcts.jl:233-239
    A = node.position.env.action_space
    ctx = node.rng
    alpha = float(node.mcts_rules.dirichlet_noise_alpha)
    dirch = np.array(rng.dirichlet(alpha, A, ctx.seed, ctx.game_id, node.position.n, ctx.noise_ctr), dtype=np.float64)
    ctx.noise_ctr += 1
    mixed = node.child_prior.astype(np.float64) * (1 - dirichlet_noise_weight) + dirch * dirichlet_noise_weight
    node.child_prior[:] = mixed.astype(f32)


def children_as_pi(node, squash=False):                  # mcts.jl:241-252
    probs = node.child_N
    if squash:
        # Float32 .^ Float64 -> Float64; pow restated with det_pow so the engine can match bit for bit
        p64 = np.array([rng.det_pow(float(x), 0.98) for x in probs], dtype=np.float64)
        return p64 / rng.butterfly_sum32(list(p64))
    return probs / probs.sum(dtype=f32)
