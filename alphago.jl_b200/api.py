"""Host-side mirror of AlphaGo.jl's surface for the self-play path, over libagz.

Names, argument meaning and error behaviour follow the reference (0-based indices instead of Julia's 1-based):
  GoEnv (src/game/go/go.jl:1-26), GoPosition + play_move!/pass_move!/all_legal_moves/score/result
  (src/game/go/board.jl), GomokuEnv / GomokuPosition (src/game/gomoku/gomoku.jl, board.jl), Position (src/game/env.jl), NeuralNet (src/neural_net.jl:13-73), MCTSPlayer / MCTSNode and their functions
  (src/mcts.jl, src/mcts_play.jl), selfplay (src/selfplay.jl:1-45).
Everything that computes runs in CUDA kernels behind the C ABI; this file only marshals.
"""
import ctypes as C

import numpy as np

from . import binding as B
from .binding import IllegalMove

WHITE, EMPTY, BLACK = -1, 0, 1
_KGS_COLUMNS = "ABCDEFGHJKLMNOPQRST"


# ------------------------------------------------------------------ GoEnv / coords (go.jl, coords.jl)
class GoEnv:
    def __init__(self, board_size=19, planes=17, lib_path=None, device=0):
        assert planes % 2 == 1
        self.N = board_size
        self.action_space = board_size * board_size + 1
        self.planes = (planes - 1) // 2
        self.max_action_space = 361
        self.lib_path = lib_path
        self.device = device
        self._util = None
        self._fwd = {}

    game_kwargs = property(lambda s: {})                  # agz_config fields that select the game (AGZ_GAME_GO is the default)

    def forward_engine(self, tower_height):
        """A small engine whose network has `tower_height` blocks, for NeuralNet.__call__ (one per tower height)."""
        if tower_height not in self._fwd:
            self._fwd[tower_height] = B.Engine(self.N, lib_path=self.lib_path, **self.game_kwargs, n_games=8, readouts=8, device=self.device, tower_height=tower_height)
        return self._fwd[tower_height]

    def util_engine(self):
        """A one-slot engine used for position-level calls (rules run on the device)."""
        if self._util is None:
            self._util = B.Engine(self.N, lib_path=self.lib_path, **self.game_kwargs, n_games=1, readouts=8, device=self.device)
        return self._util


def to_flat(coord, env):                                  # coords.jl:5-7
    return env.N * env.N if coord is None else env.N * coord[1] + coord[0]


def from_flat(f, env):                                    # coords.jl:10-12
    if f == env.N * env.N:
        return None
    j, i = divmod(int(f), env.N)
    return (i, j)


def from_kgs(s, env):                                     # coords.jl:25-34
    if s == "pass":
        return None
    s = s.upper()
    return (env.N - int(s[1:]), _KGS_COLUMNS.index(s[0]))


def to_kgs(coord, env):                                   # coords.jl:37
    return "pass" if coord is None else "%s%d" % (_KGS_COLUMNS[coord[1]], env.N - coord[0])


class GomokuEnv(GoEnv):
    """GomokuEnv(board_size = 15, connect_row = 5, planes = 17) (src/game/gomoku/gomoku.jl:1-19): N^2 actions, no pass."""

    def __init__(self, board_size=15, connect_row=5, planes=17, lib_path=None, device=0):
        super().__init__(board_size, planes, lib_path, device)
        self.n_in_row = connect_row
        self.action_space = board_size * board_size

    game_kwargs = property(lambda s: {"game": B.GAME_GOMOKU, "n_in_row": s.n_in_row})


def Go(n):                                                # src/game/env.jl:2
    return GoEnv(n)


def Position(env, **kw):                                  # src/game/env.jl:1,4
    return GomokuPosition(env, **kw) if isinstance(env, GomokuEnv) else GoPosition(env, **kw)


# ------------------------------------------------------------------ GoPosition (board.jl:271-306)
class GoPosition:
    def __init__(self, env, board=None, n=0, komi=7.5, caps=(0, 0), ko=None, recent=None, to_play=BLACK, history=None):
        self.env = env
        self.board = np.zeros((env.N, env.N), np.int8) if board is None else np.array(board, dtype=np.int8)
        self.n = n
        self.komi = float(np.float32(komi))
        self.caps = tuple(caps)
        self.ko = ko
        self.recent = list(recent) if recent is not None else []   # [(color, move)], move None = pass
        self.to_play = to_play
        self.done = False
        self.history = list(history) if history is not None else []  # boards 1.. moves ago (<= 7), [i, j] arrays

    def to_c(self):
        env = self.env
        p = B.Position()
        flat = self.board.flatten(order="F")
        C.memmove(p.board, flat.ctypes.data, flat.size)
        for k, hb in enumerate(self.history[:7]):
            hf = np.asarray(hb, dtype=np.int8).flatten(order="F")
            C.memmove(p.hist[k], hf.ctypes.data, hf.size)
        p.n_hist = min(len(self.history), 7)
        p.n = self.n
        p.to_play = self.to_play
        p.ko = -1 if self.ko is None else to_flat(self.ko, env)
        p.last_move_pass = 1 if (self.recent and self.recent[-1][1] is None) else 0
        p.done = 1 if self.done else 0
        p.caps[0], p.caps[1] = self.caps
        p.komi = self.komi
        return p

    @staticmethod
    def from_c(env, p, recent):
        N2 = env.N * env.N
        board = np.frombuffer(p.board, dtype=np.int8, count=N2).reshape(env.N, env.N, order="F").copy()
        hist = [np.frombuffer(p.hist[k], dtype=np.int8, count=N2).reshape(env.N, env.N, order="F").copy() for k in range(p.n_hist)]
        pos = GoPosition(env, board=board, n=p.n, komi=p.komi, caps=(p.caps[0], p.caps[1]),
                         ko=None if p.ko < 0 else from_flat(p.ko, env), recent=recent, to_play=p.to_play, history=hist)
        pos.done = bool(p.done)
        return pos


class GomokuPosition(GoPosition):
    """GomokuPosition(env; board, n, recent, to_play) (src/game/gomoku/board.jl:25-58): `done` / `winner` come from
    has_game_ended on construction (:97-129), evaluated by the device rules."""

    def __init__(self, env, board=None, n=0, recent=None, to_play=BLACK, history=None, **_):
        super().__init__(env, board=board, n=n, komi=0.0, recent=recent, to_play=to_play, history=history)
        if board is None:
            self.winner, self.done = 0, False
        else:
            self.winner = int(env.util_engine().pos_score(self.to_c()))
            self.done = self.winner != 0 or not (self.board == EMPTY).any()


def play_move(pos_or_player, c, color=None):
    """play_move!(pos, c) (board.jl:451-509; gomoku board.jl:137-169) or play_move!(player, c) (mcts_play.jl:26-50)."""
    if isinstance(pos_or_player, MCTSPlayer):
        return pos_or_player.play_move(c)
    pos = pos_or_player
    env = pos.env
    if isinstance(pos, GomokuPosition):
        assert not pos.done                                   # @assert !new_pos.done (gomoku board.jl:144)
        if c is None:
            raise IllegalMove()
    cp = pos.to_c()
    if color is not None:
        cp.to_play = color
    out = env.util_engine().pos_play_move(cp, to_flat(c, env))
    if isinstance(pos, GomokuPosition):
        N2 = env.N * env.N
        board = np.frombuffer(out.board, dtype=np.int8, count=N2).reshape(env.N, env.N, order="F").copy()
        hist = [np.frombuffer(out.hist[k], dtype=np.int8, count=N2).reshape(env.N, env.N, order="F").copy() for k in range(out.n_hist)]
        return GomokuPosition(env, board=board, n=out.n, recent=pos.recent + [(cp.to_play, c)], to_play=out.to_play, history=hist)
    return GoPosition.from_c(env, out, pos.recent + [(cp.to_play, c)])


def pass_move(pos):                                       # board.jl:426-440
    return play_move(pos, None)


def all_legal_moves(pos):                                 # board.jl:393-424
    return pos.env.util_engine().pos_legal_moves(pos.to_c())


def is_move_legal(pos, move):                             # board.jl:376-391
    return bool(all_legal_moves(pos)[to_flat(move, pos.env)])


def score(pos):                                           # board.jl:511-533
    return pos.env.util_engine().pos_score(pos.to_c())


def result(pos):                                          # board.jl:535-544
    s = score(pos)
    return 1 if s > 0 else (-1 if s < 0 else 0)


def result_string(pos):                                   # board.jl:546-555; gomoku board.jl:184-193
    s = score(pos)
    if isinstance(pos, GomokuPosition):
        return "B" if s > 0 else ("W" if s < 0 else "DRAW")
    return "B+%.1f" % s if s > 0 else ("W+%.1f" % abs(s) if s < 0 else "DRAW")


def liberties(pos):
    """LibertyTracker.liberty_cache (board.jl:99-164) as an (N, N) array."""
    env = pos.env
    return env.util_engine().pos_liberties(pos.to_c()).reshape(env.N, env.N, order="F")


def _hist_stack(pos):
    """8 boards (current first), flat order, oldest repeated (features.jl:7-14)."""
    boards = [pos.board] + list(pos.history[:7])
    while len(boards) < 8:
        boards.append(boards[-1])
    return np.stack([np.asarray(b, np.int8).flatten(order="F") for b in boards])


def get_feats(pos):
    """get_feats (features.jl:24-26): (N, N, 17) array [i, j, c]."""
    env = pos.env
    out = env.util_engine().features(_hist_stack(pos)[None], np.array([pos.to_play], np.int8))
    return np.transpose(out[0].reshape(17, env.N, env.N), (2, 1, 0)).copy()


# ------------------------------------------------------------------ NeuralNet (neural_net.jl:7-73)
def _glorot_uniform(rs, *dims):
    if len(dims) == 2:
        fan_out, fan_in = dims
    else:
        rf = int(np.prod(dims[:-2]))
        fan_in, fan_out = dims[-2] * rf, dims[-1] * rf
    return ((rs.random_sample(dims) - 0.5) * np.sqrt(24.0 / (fan_in + fan_out))).astype(np.float32)


class NeuralNet:
    """NeuralNet(env; tower_height = 19).  Parameters are kept as the three Flux `params` lists that save_model
    writes (src/train.jl:27-33) and pushed to every engine that evaluates with this net."""

    def __init__(self, env, tower_height=19, seed=0, evaluator=B.EVAL_NN_TC):
        self.env, self.tower_height, self.evaluator = env, tower_height, evaluator
        N, Cf = env.N, 256
        rs = np.random.RandomState(seed)
        z, o = (lambda n: np.zeros(n, np.float32)), (lambda n: np.ones(n, np.float32))
        base = [_glorot_uniform(rs, 3, 3, 2 * env.planes + 1, Cf), z(Cf), z(Cf), o(Cf)]
        for _ in range(tower_height):
            base += [_glorot_uniform(rs, 3, 3, Cf, Cf), z(Cf), _glorot_uniform(rs, 3, 3, Cf, Cf), z(Cf), z(Cf), o(Cf), z(Cf), o(Cf)]
        value = [_glorot_uniform(rs, 1, 1, Cf, 1), z(1), z(1), o(1), _glorot_uniform(rs, 256, N * N), z(256), _glorot_uniform(rs, 1, 256), z(1)]
        policy = [_glorot_uniform(rs, 1, 1, Cf, 2), z(2), z(2), o(2), _glorot_uniform(rs, env.action_space, 2 * N * N), z(env.action_space)]
        self.params = [base, value, policy]
        nbn = [1 + 2 * tower_height, 1, 1]
        self.bn_mu = [np.zeros(n * (Cf if k == 0 else (1 if k == 1 else 2)), np.float32) for k, n in enumerate(nbn)]
        self.bn_sigma = [np.ones_like(m) for m in self.bn_mu]
        self.bn_mode = B.BN_VAR_EPS
        self._version = 0

    def load_params(self, base=None, value=None, policy=None, bn_mu=None, bn_sigma=None, bn_mode=None):
        for k, lst in enumerate((base, value, policy)):
            if lst is not None:
                self.params[k] = [np.asarray(a, np.float32) for a in lst]
        if bn_mu is not None:
            self.bn_mu = [np.asarray(a, np.float32).ravel() for a in bn_mu]
        if bn_sigma is not None:
            self.bn_sigma = [np.asarray(a, np.float32).ravel() for a in bn_sigma]
        if bn_mode is not None:
            self.bn_mode = bn_mode
        self._version += 1

    def push(self, engine):
        if getattr(engine, "_nn_token", None) == (id(self), self._version):
            return
        for k in range(3):
            flat = np.concatenate([np.asarray(a, np.float32).flatten(order="F") for a in self.params[k]])
            engine.net_set_params(k, flat)
            engine.net_set_bn_stats(k, self.bn_mu[k], self.bn_sigma[k], self.bn_mode)
        engine._nn_token = (id(self), self._version)

    def __call__(self, positions):
        """(nn::NeuralNet)(positions) -> (pi: A x B, v: B) (neural_net.jl:57-68); a single Position gives (pi, v)."""
        single = isinstance(positions, GoPosition)   # (GomokuPosition is a GoPosition here)
        plist = [positions] if single else list(positions)
        eng = self.env.forward_engine(self.tower_height)   # the engine's network shape is fixed at creation (tower_height)
        self.push(eng)
        bh = np.stack([_hist_stack(p) for p in plist])
        tp = np.array([p.to_play for p in plist], np.int8)
        pi, v = eng.net_forward(self.evaluator, bh, tp)
        return (pi[0], v[0]) if single else (pi.T.copy(), v)


# ------------------------------------------------------------------ MCTSNode / MCTSPlayer (mcts.jl, mcts_play.jl)
class MCTSNode:
    """Read-only view of one node of a device tree."""

    def __init__(self, player, node_id):
        self.player, self.id = player, node_id

    def _view(self):
        return self.player.engine.tree_read_node(0, self.id)

    def __eq__(self, o):
        return isinstance(o, MCTSNode) and o.player is self.player and o.id == self.id

    def __hash__(self):
        return hash((id(self.player), self.id))

    A = property(lambda s: s.player.env.action_space)
    child_N = property(lambda s: np.array(s._view().child_N[:s.A], np.float32))
    child_W = property(lambda s: np.array(s._view().child_W[:s.A], np.float32))
    child_prior = property(lambda s: np.array(s._view().child_prior[:s.A], np.float32))
    is_expanded = property(lambda s: bool(s._view().is_expanded))
    N = property(lambda s: np.float32(s._view().N))
    W = property(lambda s: np.float32(s._view().W))
    Q = property(lambda s: np.float32(np.float32(s._view().W) / np.float32(np.float32(1) + np.float32(s._view().N))))

    @property
    def fmove(self):
        f = self._view().fmove
        return None if f < 0 else f

    @property
    def parent(self):
        p = self._view().parent
        return None if p < 0 else MCTSNode(self.player, p)

    @property
    def children(self):
        ch = self._view().children
        return {a: MCTSNode(self.player, ch[a]) for a in range(self.A) if ch[a] >= 0}

    @property
    def position(self):
        v, env = self._view(), self.player.env
        board = np.array(v.board[:env.N * env.N], np.int8).reshape(env.N, env.N, order="F")
        if isinstance(env, GomokuEnv):
            return GomokuPosition(env, board=board, n=v.n, to_play=v.to_play)
        pos = GoPosition(env, board=board, n=v.n, komi=self.player.engine.cfg.komi, ko=None if v.ko < 0 else from_flat(v.ko, env),
                         recent=[(-v.to_play, None)] if v.last_move_pass else [], to_play=v.to_play)
        pos.done = bool(v.done)
        return pos

    def child_action_score(self):                          # mcts.jl:86-87
        return np.array(self._view().action_score[:self.A], np.float64)

    def child_Q(self):                                     # mcts.jl:89
        return (self.child_W / (np.float32(1) + self.child_N)).astype(np.float32)

    def legal_moves(self):                                 # mcts.jl:84
        return np.array(self._view().legal[:self.A], np.int8)


class MCTSPlayer:
    """MCTSPlayer(env, network; num_readouts = 800, two_player_mode = false, resign_threshold = -0.9) (mcts_play.jl:3-24).
    `network` is a NeuralNet, an object with `fake_priors` / `fake_value` (the reference tests' DummyNet) or any
    callable positions -> (A x B priors, B values)."""

    def __init__(self, env, network, num_readouts=800, two_player_mode=False, resign_threshold=-0.9, seed=0, game_id=0,
                 max_parallel=64, nodes_per_game=0):
        self.env, self.network = env, network
        self.num_readouts, self.two_player_mode = num_readouts, two_player_mode
        self.tau_threshold = -1 if two_player_mode else (env.N * env.N // 12) // 2 * 2
        self.resign_threshold = resign_threshold
        self.seed, self.game_id = seed, game_id
        self.result, self.result_string = 0, ""
        self.recent = []
        self.engine = B.Engine(env.N, lib_path=env.lib_path, **env.game_kwargs, n_games=1, readouts=num_readouts, tau_threshold=self.tau_threshold,
                               resign_threshold=resign_threshold, seed=seed, max_parallel=max_parallel, device=env.device,
                               nodes_per_game=nodes_per_game, tower_height=getattr(network, "tower_height", 1))
        self._bind_network()
        self.komi = 7.5

    def _bind_network(self):
        net = self.network
        if isinstance(net, NeuralNet):
            net.push(self.engine)
            self.engine.set_evaluator(net.evaluator)
            self._mode = "nn"
        elif hasattr(net, "fake_priors"):
            self.engine.set_dummy_evaluator(net.fake_priors, float(net.fake_value))
            self.engine.set_evaluator(B.EVAL_DUMMY)
            self._mode = "dummy"
        else:
            self._mode = "callable"

    root = property(lambda s: MCTSNode(s, s.engine.tree_root(0)[0]))
    position = property(lambda s: s.root.position)

    @property
    def searches_pi(self):
        return list(self.engine.tree_read_record(0)[3]) if not self.two_player_mode else []

    @property
    def qs(self):
        return list(self.engine.tree_read_record(0)[2])

    def initialize_game(self, pos=None):                   # mcts_play.jl:110-118
        self.result, self.result_string = 0, ""
        if pos is None:
            self.recent, self.komi = [], 7.5
            self.engine.cfg.komi = 7.5
            self.engine.tree_init(0, None, self.game_id)
        else:
            if not isinstance(pos, GomokuPosition) and abs(pos.komi - self.engine.cfg.komi) > 0:
                # komi lives in the engine config: rebuild the handle for a non-default komi
                self.engine.close()
                self.engine = B.Engine(self.env.N, lib_path=self.env.lib_path, **self.env.game_kwargs, n_games=1, readouts=self.num_readouts,
                                       tau_threshold=self.tau_threshold, resign_threshold=self.resign_threshold, seed=self.seed,
                                       max_parallel=64, komi=pos.komi, device=self.env.device,
                                       tower_height=getattr(self.network, "tower_height", 1))
                self._bind_network()
            self.recent, self.komi = list(pos.recent), pos.komi
            self.engine.tree_init(0, pos.to_c(), self.game_id)

    def tree_search(self, parallel_readouts=8):            # mcts_play.jl:73-98
        if self._mode != "callable":
            return self.engine.tree_search(0, parallel_readouts)
        # generic host callable: the same steps with the network call made from the host
        leaves, failsafe = [], 0
        while len(leaves) < parallel_readouts and failsafe < 2 * parallel_readouts:
            failsafe += 1
            leaf = MCTSNode(self, self.engine.tree_select_leaf(0))
            v = leaf._view()
            if v.done or v.n >= self.engine.cfg.max_game_length:
                self.engine.tree_backup_value(0, leaf.id, float(result(leaf.position)))
                continue
            self.engine.tree_add_virtual_loss(0, leaf.id)
            leaves.append(leaf)
        if leaves:
            probs, values = self.network([l.position for l in leaves])
            for k, leaf in enumerate(leaves):
                self.engine.tree_revert_virtual_loss(0, leaf.id)
                self.engine.tree_incorporate(0, leaf.id, np.asarray(probs)[:, k], float(values[k]))
        return len(leaves)

    def pick_move(self):                                   # mcts_play.jl:52-71
        return from_flat(self.engine.tree_pick_move(0), self.env)

    def play_move(self, c):                                # mcts_play.jl:26-50
        to_play = self.root._view().to_play
        try:
            if c is None and isinstance(self.env, GomokuEnv):   # no pass in this game; the reference's `catch` swallows the error
                raise IllegalMove()
            self.engine.tree_play_move(0, to_flat(c, self.env))
        except IllegalMove:
            print("Illegal move")
            return False
        self.recent.append((to_play, c))
        return True

    def should_resign(self):                               # mcts_play.jl:124
        return self.engine.tree_should_resign(0, self.resign_threshold)

    def is_done(self):                                     # mcts_play.jl:120
        v = self.root._view()
        return self.result != 0 or bool(v.done) or v.n >= self.engine.cfg.max_game_length

    def set_result(self, winner, was_resign):              # mcts_play.jl:100-108
        self.result = winner
        self.result_string = ("B+R" if winner == BLACK else "W+R") if was_resign else result_string(self.root.position)

    def extract_data(self):                                # mcts_play.jl:126-139 -> board.jl:557-578
        pis = self.searches_pi
        assert len(pis) == self.root._view().n
        positions, results = [], []
        pos = Position(self.env, komi=self.komi)
        for color, mv in self.recent:
            positions.append(pos)
            results.append(self.result)
            pos = play_move(pos, mv, color=color)
        return positions, pis, results

    def suggest_move(self):                                # mcts_play.jl:144-151
        cur = self.root.N
        while self.root.N < cur + self.num_readouts:
            self.tree_search()
        return self.pick_move()


# ------------------------------------------------------------------ evaluate (neural_net.jl:103-158)
class MatchGame:
    """What evaluate leaves in its two players per game: the moves, `result` / `result_string` (set_result!) and whether the
    final position scores as a Black win (what the win counter reads, neural_net.jl:150)."""

    def __init__(self):
        self.moves, self.result, self.result_string, self.black_won = [], 0, "", False


def _bind(eng, net):
    if isinstance(net, NeuralNet):
        net.push(eng)
        eng.set_evaluator(net.evaluator)
    else:
        eng.set_dummy_evaluator(getattr(net, "fake_priors", None), float(getattr(net, "fake_value", 0.0)))
        eng.set_evaluator(B.EVAL_DUMMY)


def _score_string(s, env=None):
    if isinstance(env, GomokuEnv):                          # result_string of src/game/gomoku/board.jl:184-193
        return "B" if s > 0 else ("W" if s < 0 else "DRAW")
    return "B+%.1f" % s if s > 0 else ("W+%.1f" % abs(s) if s < 0 else "DRAW")


def evaluate(env, black_net, white_net, num_games=400, ro=800, verbose=False, seed=0, resign_threshold=-0.9, details=None,
             **engine_overrides):
    """evaluate(env, black_net, white_net; num_games, ro) -> Bool: all `num_games` gating games run concurrently, one engine
    per player (two_player_mode: tau_threshold = -1, no noise), alternating exactly like the reference's two MCTSPlayers.
    `details`, if a list, receives one MatchGame per game."""
    G = num_games
    engines = []
    try:
        for k, net in enumerate((black_net, white_net)):
            eng = B.Engine(env.N, lib_path=env.lib_path, **env.game_kwargs, n_games=G, readouts=ro, tau_threshold=-1, inject_noise=0,
                           resign_threshold=resign_threshold, seed=seed + k, device=env.device,
                           tower_height=getattr(net, "tower_height", 1), **engine_overrides)
            engines.append(eng)
            _bind(eng, net)
            eng.match_start()
        games = [MatchGame() for _ in range(G)]
        alive = np.ones(G, bool)
        black_score = np.zeros(G, np.float32)
        num_move = 0
        while alive.any():
            active, inactive = (engines[1], engines[0]) if num_move % 2 == 1 else (engines[0], engines[1])
            to_play = WHITE if num_move % 2 == 1 else BLACK
            moves, resigned, rscore = active.match_search(alive)
            for g in np.flatnonzero(alive & resigned):              # forced resignation (:129-133)
                games[g].result = -to_play
                games[g].result_string = "B+R" if -to_play == BLACK else "W+R"
                black_score[g] = rscore[g]
                alive[g] = False
            mv = np.where(alive, moves, -1).astype(np.int32)
            done, sc = active.match_play(mv)
            inactive.match_play(mv)
            for g in np.flatnonzero(alive):
                games[g].moves.append(int(mv[g]))
            for g in np.flatnonzero(alive & done):                  # is_done(active) (:140-146)
                games[g].result = 1 if sc[g] > 0 else (-1 if sc[g] < 0 else 0)
                games[g].result_string = _score_string(float(sc[g]), env)
                black_score[g] = sc[g]
                alive[g] = False
            num_move += 1
        for g in range(G):
            games[g].black_won = bool(black_score[g] > 0)           # result(black.root.position) == BLACK (:150)
        games_won = int(sum(gm.black_won for gm in games))
    finally:
        for eng in engines:
            eng.close()
    if details is not None:
        details.extend(games)
    if verbose:
        print("Won %d / %d. Win rate: %s. " % (games_won, G, games_won / G), end="")
    return games_won / G >= 0.55


# ------------------------------------------------------------------ play (play.jl:25-77)
def play(env, nn=None, tower_height=19, num_readouts=800, mode=0, input_fn=input, print_fn=print, seed=0):
    """play(env, nn; tower_height, num_readouts, mode): a game against the engine on the terminal (mode 0: the human starts with
    Black, otherwise with White).  Same loop as the reference: the human types KGS coordinates ("D4", "pass"), the engine searches
    `num_readouts` more readouts from its current root and plays pick_move; both kinds of move go through play_move!(player, c),
    which rejects illegal ones.  `input_fn` / `print_fn` make it scriptable.  Returns the MCTSPlayer (result set)."""
    assert 0 <= tower_height <= 19
    if nn is None:
        nn = NeuralNet(env, tower_height=tower_height, seed=seed)
    az = MCTSPlayer(env, nn, num_readouts=num_readouts, two_player_mode=True, seed=seed)
    az.initialize_game()
    num_moves = 0
    mode = 0 if mode == 0 else 1
    while not az.is_done():
        print_fn(_board_string(az.position))
        if num_moves % 2 == mode:
            text = input_fn("Your turn: ")
            try:
                move = from_kgs(text.strip(), env)
            except Exception:
                print_fn("Try again.")
                continue
        else:
            move = az.suggest_move()
            print_fn("AlphaZero's turn: " + to_kgs(move, env))
        if az.play_move(move):
            num_moves += 1
    print_fn(_board_string(az.position))
    winner = result(az.position)
    az.set_result(winner, False)
    human = BLACK if mode == 0 else WHITE
    print_fn(("You Win! " if winner == human else "AlphaZero wins! ") + az.result_string)
    return az


def _board_string(pos):
    sym = {BLACK: "X", WHITE: "O", EMPTY: "."}
    N = pos.env.N
    rows = ["%2d %s" % (N - i, " ".join(sym[int(pos.board[i, j])] for j in range(N))) for i in range(N)]
    return "\n".join(rows + ["   " + " ".join(_KGS_COLUMNS[:N])])


# ------------------------------------------------------------------ train (train.jl:38-92)
def train(env, num_games=25000, memory_size=500000, batch_size=32, epochs=1, ckp_freq=1000, readouts=800, tower_height=19,
          model=None, start_training_after=50000, concurrent=None, seed=0, model_dir=None, verbose=True, lr=0.02, momentum=0.9,
          **engine_overrides):
    """train(env; num_games, memory_size, batch_size, epochs, ckp_freq, readouts, tower_height, model, start_training_after) -> NeuralNet.
    The reference plays one game, appends its (position, pi, z) tuples to the buffers, and -- once `start_training_after` tuples
    are there -- takes `epochs` optimisation steps on one uniform batch per finished game (train.jl:56-73).  Here `concurrent` games
    run at once on the GPU with the current network; every harvested game triggers the same per-game step, so the ratio of
    optimisation steps to games is the reference's.  The replay ring is the engine's (500 000 deep, trim-oldest)."""
    from . import weights_io
    cur_nn = model if model is not None else NeuralNet(env, tower_height=tower_height, seed=seed)
    conc = concurrent or min(num_games, 1024)
    eng = B.Engine(env.N, lib_path=env.lib_path, **env.game_kwargs, n_games=conc, readouts=readouts, seed=seed, device=env.device,
                   tower_height=cur_nn.tower_height, evaluator=cur_nn.evaluator, options={"replay.capacity": memory_size}, **engine_overrides)
    losses = []
    try:
        cur_nn.push(eng)
        eng.selfplay_start(num_games)
        done, last_ckp, stalls = 0, 0, 0
        while done < num_games:
            pr = eng.selfplay_step(8)
            if pr.error:
                raise B.AgzError(pr.error, "a game stopped on the device")
            n_tuples = eng.replay_gather()
            recs = eng.selfplay_harvest(conc)
            stalls = 0 if (recs or pr.games_live) else stalls + 1
            if stalls > 4:
                break
            for r in recs:
                done += 1
                if n_tuples >= start_training_after and n_tuples >= batch_size:
                    loss = 0.0
                    for ep in range(epochs):     # get_replay_batch + _train (train.jl:66-70), all on the device
                        loss += eng.train_step_from_replay(batch_size, seed=seed * 1000003 + done * 131 + ep, lr=lr, momentum=momentum)
                    losses.append(loss / epochs)
                    if verbose:
                        print("Episode %d over. Loss: %s. Winner: %s. Moves: %d." % (done, losses[-1], r.result_string, r.n_moves))
                if model_dir is not None and done // ckp_freq > last_ckp:
                    last_ckp = done // ckp_freq
                    weights_io.save_model(cur_nn, model_dir, engine=eng)
                    if verbose:
                        print("Model saved. ", end="")
        weights_io.pull_from_engine(cur_nn, eng)
    finally:
        eng.close()
    cur_nn.train_losses = losses
    return cur_nn


# free-function spellings used by the reference and its tests
def initialize_game(player, pos=None): return player.initialize_game(pos)
def tree_search(player, parallel_readouts=8): return player.tree_search(parallel_readouts)
def pick_move(player): return player.pick_move()
def should_resign(player): return player.should_resign()
def is_done(x): return x.is_done()
def set_result(player, winner, was_resign): return player.set_result(winner, was_resign)
def extract_data(player): return player.extract_data()
def select_leaf(node): return MCTSNode(node.player, node.player.engine.tree_select_leaf(0, node.id))
def incorporate_results(node, move_probs, value, up_to=None): return node.player.engine.tree_incorporate(0, node.id, move_probs, float(value))
def maybe_add_child(node, fcoord): return MCTSNode(node.player, node.player.engine.tree_maybe_add_child(0, node.id, fcoord))
def add_virtual_loss(node, up_to=None): return node.player.engine.tree_add_virtual_loss(0, node.id)
def revert_virtual_loss(node, up_to=None): return node.player.engine.tree_revert_virtual_loss(0, node.id)
def inject_noise(node): return node.player.engine.tree_inject_noise(0)


# ------------------------------------------------------------------ selfplay (selfplay.jl:1-45)
class SelfPlayResult:
    """The fields callers read from the MCTSPlayer that `selfplay` returns (train.jl:57-58,73)."""

    def __init__(self, env, rec):
        self.env, self.record = env, rec
        self.searches_pi, self.qs = list(rec.searches_pi), list(rec.qs)
        self.result, self.result_string = rec.result, rec.result_string
        self.moves = [from_flat(int(m), env) for m in rec.moves]
        self.n_moves = rec.n_moves

    def extract_data(self):
        positions, results = [], []
        pos, color = Position(self.env), BLACK
        for mv in self.moves:
            positions.append(pos)
            results.append(self.result)
            pos = play_move(pos, mv)
        return positions, self.searches_pi, results


def selfplay(env, nn, num_ro=800, seed=0, n_games=1, concurrent=None, options=None, **engine_overrides):
    """selfplay(env, nn, num_ro) -> player; with n_games > 1 plays that many games concurrently on the GPU and
    returns a list (game ids 0..n_games-1).  `nn` is a NeuralNet or a DummyNet-like object.  `options`: agz_set_option keys."""
    conc = concurrent or n_games
    eng = B.Engine(env.N, lib_path=env.lib_path, **env.game_kwargs, n_games=conc, readouts=num_ro, seed=seed, device=env.device,
                   tower_height=getattr(nn, "tower_height", 1), options=options, **engine_overrides)
    try:
        if isinstance(nn, NeuralNet):
            nn.push(eng)
            eng.set_evaluator(nn.evaluator)
        else:
            eng.set_dummy_evaluator(getattr(nn, "fake_priors", None), float(getattr(nn, "fake_value", 0.0)))
            eng.set_evaluator(B.EVAL_DUMMY)
        recs = eng.selfplay_run(n_games)
    finally:
        eng.close()
    out = [SelfPlayResult(env, r) for r in recs]
    return out[0] if n_games == 1 else out
