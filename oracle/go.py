"""Go rules: restatement of src/game/go/{go,board,coords}.jl of the reference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  0-based indices throughout:
board[i, j], i = row from the top, j = column; flat move f = N*j + i, pass = N*N.
"""
import numpy as np

WHITE, EMPTY, BLACK, FILL, KO, UNKNOWN = range(-1, 5)   # board.jl:12
MISSING_GROUP_ID = -1                                    # board.jl:15


class IllegalMove(Exception):                            # AlphaGo.jl:8
    pass


class GoEnv:
    """go.jl:1-26."""

    def __init__(self, board_size=19, planes=17):
        N = board_size
        self.N = N
        self.action_space = N * N + 1
        assert planes % 2 == 1
        self.planes = (planes - 1) // 2
        self.max_action_space = 361
        ok = lambda c: 0 <= c[0] < N and 0 <= c[1] < N
        self.NEIGHBORS = {}
        self.DIAGONALS = {}
        for x in range(N):
            for y in range(N):
                self.NEIGHBORS[(x, y)] = [c for c in ((x + 1, y), (x - 1, y), (x, y + 1), (x, y - 1)) if ok(c)]
                self.DIAGONALS[(x, y)] = [c for c in ((x + 1, y + 1), (x + 1, y - 1), (x - 1, y + 1), (x - 1, y - 1)) if ok(c)]

    def empty_board(self):
        return np.zeros((self.N, self.N), dtype=np.int8)


# ------------------------------------------------------------------ coords.jl
_KGS_COLUMNS = "ABCDEFGHJKLMNOPQRST"
_SGF_COLUMNS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def to_flat(coord, env):                                 # coords.jl:5-7
    return env.N * env.N if coord is None else env.N * coord[1] + coord[0]


def from_flat(f, env):                                   # coords.jl:10-12
    if f == env.N * env.N:
        return None
    j, i = divmod(f, env.N)
    return (i, j)


def from_sgf(s):                                         # coords.jl:14-20
    if s is None or s == "":
        return None
    return (_SGF_COLUMNS.index(s[1]), _SGF_COLUMNS.index(s[0]))


def to_sgf(coord):                                       # coords.jl:23
    return "" if coord is None else _SGF_COLUMNS[coord[1]] + _SGF_COLUMNS[coord[0]]


def from_kgs(s, env):                                    # coords.jl:25-34
    if s == "pass":
        return None
    s = s.upper()
    col = _KGS_COLUMNS.index(s[0])
    row_from_bottom = int(s[1:])
    return (env.N - row_from_bottom, col)


def to_kgs(coord, env):                                  # coords.jl:37
    return "pass" if coord is None else "%s%d" % (_KGS_COLUMNS[coord[1]], env.N - coord[0])


def coordify(x, N):                                      # board.jl:84 (0-based flat -> (i, j))
    return (x % N, x // N)


# ------------------------------------------------------------------ board.jl
class PlayerMove:                                        # board.jl:17-20
    __slots__ = ("color", "move")

    def __init__(self, color, move):
        self.color, self.move = color, move

    def __eq__(self, o):
        return self.color == o.color and self.move == o.move

    def __repr__(self):
        return "PlayerMove(%r, %r)" % (self.color, self.move)


def find_reached(board, c, env):                         # board.jl:28-45
    color = board[c]
    chain = {c}
    reached = set()
    frontier = [c]
    while frontier:
        current = frontier.pop()
        chain.add(current)
        for n in env.NEIGHBORS[current]:
            if board[n] == color and n not in chain:
                frontier.append(n)
            elif board[n] != color:
                reached.add(n)
    return chain, reached


def is_koish(board, c, env):                             # board.jl:47-56
    if board[c] != EMPTY:
        return None
    neighs = {int(board[n]) for n in env.NEIGHBORS[c]}
    if len(neighs) == 1 and EMPTY not in neighs:
        return next(iter(neighs))
    return None


def is_eyeish(board, c, env):                            # board.jl:58-81
    color = is_koish(board, c, env)
    if color is None:
        return None
    diag_faults = 0
    diag = env.DIAGONALS[c]
    if len(diag) < 4:
        diag_faults += 1
    for d in diag:
        if board[d] not in (color, EMPTY):
            diag_faults += 1
    return None if diag_faults > 1 else color


class Group:                                             # board.jl:88-97
    __slots__ = ("id", "stones", "liberties", "color")

    def __init__(self, id, stones, liberties, color):
        self.id, self.stones, self.liberties, self.color = id, stones, liberties, color

    def __eq__(self, o):
        return self.stones == o.stones and self.liberties == o.liberties and self.color == o.color


class LibertyTracker:                                    # board.jl:99-116
    def __init__(self, env, group_index=None, groups=None, liberty_cache=None, max_group_id=1):
        N = env.N
        self.group_index = group_index if group_index is not None else -np.ones((N, N), dtype=np.int16)
        self.groups = groups if groups is not None else {}
        self.liberty_cache = liberty_cache if liberty_cache is not None else np.zeros((N, N), dtype=np.uint8)
        self.max_group_id = max_group_id

    def copy(self, env):                                 # board.jl:118-130
        groups = {g.id: Group(g.id, set(g.stones), set(g.liberties), g.color) for g in self.groups.values()}
        return LibertyTracker(env, self.group_index.copy(), groups, self.liberty_cache.copy(), self.max_group_id)

    @staticmethod
    def from_board(board, env):                          # board.jl:132-164
        board = board.copy()
        curr_group_id = 0
        lt = LibertyTracker(env)
        for color in (WHITE, BLACK):
            while (board == color).any():
                curr_group_id += 1
                flat = board.flatten(order="F")
                coord = coordify(int(np.argmax(flat == color)), env.N)
                chain, reached = find_reached(board, coord, env)
                liberties = {r for r in reached if board[r] == EMPTY}
                lt.groups[curr_group_id] = Group(curr_group_id, chain, liberties, color)
                for s in chain:
                    lt.group_index[s] = curr_group_id
                    board[s] = FILL
        lt.max_group_id = curr_group_id
        for g in lt.groups.values():
            for s in g.stones:
                lt.liberty_cache[s] = len(g.liberties)
        return lt

    def _merge_from_played(self, color, played, libs, other_group_ids):   # board.jl:166-190
        stones = {played}
        liberties = set(libs)
        for gid in other_group_ids:
            other = self.groups.pop(gid)
            stones |= other.stones
            liberties |= other.liberties
        if other_group_ids:
            liberties -= {played}
        assert not (stones & liberties)
        self.max_group_id += 1
        result = Group(self.max_group_id, stones, liberties, color)
        self.groups[result.id] = result
        for s in result.stones:
            self.group_index[s] = result.id
            self.liberty_cache[s] = len(result.liberties)
        return result

    def _update_liberties(self, group_id, add=(), remove=()):             # board.jl:192-203
        g = self.groups[group_id]
        new_libs = (g.liberties | set(add)) - set(remove)
        self.groups[group_id] = Group(group_id, g.stones, new_libs, g.color)
        for s in g.stones:
            self.liberty_cache[s] = len(new_libs)

    def _capture_group(self, group_id):                                   # board.jl:205-213
        dead = self.groups.pop(group_id)
        for s in dead.stones:
            self.group_index[s] = MISSING_GROUP_ID
            self.liberty_cache[s] = 0
        return dead.stones

    def _handle_captures(self, captured_stones, env):                     # board.jl:216-225
        for s in captured_stones:
            for n in env.NEIGHBORS[s]:
                gid = int(self.group_index[n])
                if gid != MISSING_GROUP_ID:
                    self._update_liberties(gid, add={s})

    def add_stone(self, color, c, env):                                   # board.jl:227-269
        assert self.group_index[c] == MISSING_GROUP_ID
        captured_stones = set()
        opp_ids, friendly_ids, empty_neighbors = set(), set(), set()
        for n in env.NEIGHBORS[c]:
            gid = int(self.group_index[n])
            if gid != MISSING_GROUP_ID:
                if self.groups[gid].color == color:
                    friendly_ids.add(gid)
                else:
                    opp_ids.add(gid)
            else:
                empty_neighbors.add(n)
        new_group = self._merge_from_played(color, c, empty_neighbors, friendly_ids)
        for gid in opp_ids:
            if len(self.groups[gid].liberties) == 1:
                captured_stones |= self._capture_group(gid)
            else:
                self._update_liberties(gid, remove={c})
        self._handle_captures(captured_stones, env)
        if len(self.groups[new_group.id].liberties) == 0:                 # suicide is illegal
            raise IllegalMove()
        return captured_stones


class GoPosition:                                        # board.jl:271-306
    def __init__(self, env, board=None, n=0, komi=7.5, caps=(0, 0), lib_tracker=None, ko=None,
                 recent=None, board_deltas=None, to_play=BLACK):
        self.env = env
        self.board = board if board is not None else env.empty_board()
        self.n = n
        self.komi = np.float32(komi)
        self.caps = caps
        self.lib_tracker = lib_tracker if lib_tracker is not None else LibertyTracker.from_board(self.board, env)
        self.ko = ko
        self.recent = recent if recent is not None else []
        # deltas stacked on axis 0 here (reference: axis 3), newest first
        self.board_deltas = board_deltas if board_deltas is not None else np.zeros((0, env.N, env.N), dtype=np.int8)
        self.to_play = to_play
        self.done = False

    def copy(self):                                      # board.jl:308-315 (note: `done` is reset)
        return GoPosition(self.env, board=self.board.copy(), n=self.n, komi=self.komi, caps=self.caps,
                          lib_tracker=self.lib_tracker.copy(self.env), ko=self.ko,
                          recent=list(self.recent), board_deltas=self.board_deltas, to_play=self.to_play)


def is_move_suicidal(pos, move):                         # board.jl:354-374
    potential_libs = set()
    for n in pos.env.NEIGHBORS[move]:
        gid = int(pos.lib_tracker.group_index[n])
        if gid == MISSING_GROUP_ID:
            return False
        g = pos.lib_tracker.groups[gid]
        if g.color == pos.to_play:
            potential_libs |= g.liberties
        elif len(g.liberties) == 1:
            return False
    potential_libs -= {move}
    return not potential_libs


def is_move_legal(pos, move):                            # board.jl:376-391
    if move is None:
        return True
    if pos.board[move] != EMPTY:
        return False
    if move == pos.ko:
        return False
    if is_move_suicidal(pos, move):
        return False
    return True


def all_legal_moves(pos):                                # board.jl:393-424
    N = pos.env.N
    legal = np.ones((N, N), dtype=np.int8)
    legal[pos.board != EMPTY] = 0
    adjacent = np.ones((N + 2, N + 2), dtype=np.int8)
    adjacent[1:-1, 1:-1] = np.abs(pos.board)
    num_adj = adjacent[:-2, 1:-1] + adjacent[1:-1, :-2] + adjacent[2:, 1:-1] + adjacent[1:-1, 2:]
    surrounded = (pos.board == EMPTY) & (num_adj == 4)
    for c in np.flatnonzero(surrounded.flatten(order="F")):
        coord = coordify(int(c), N)
        if is_move_suicidal(pos, coord):
            legal[coord] = 0
    if pos.ko is not None:
        legal[pos.ko] = 0
    return np.concatenate([legal.flatten(order="F"), np.ones(1, dtype=np.int8)])


def get_first_n(x, n):                                   # board.jl:86
    return x if x.shape[0] < n else x[:n]


def pass_move(pos, mutate=False):                        # board.jl:426-440
    new_pos = pos if mutate else pos.copy()
    new_pos.n += 1
    N = pos.env.N
    new_pos.recent.append(PlayerMove(new_pos.to_play, None))
    new_pos.board_deltas = np.concatenate(
        [np.zeros((1, N, N), dtype=np.int8), get_first_n(new_pos.board_deltas, new_pos.env.planes - 2)], axis=0)
    new_pos.to_play *= -1
    new_pos.ko = None
    if len(new_pos.recent) > 1 and new_pos.recent[-2].move is None:
        new_pos.done = True
    return new_pos


def flip_playerturn(pos, mutate=False):                  # board.jl:442-447
    new_pos = pos if mutate else pos.copy()
    new_pos.ko = None
    new_pos.to_play *= -1
    return new_pos


def play_move(pos, c, color=None, mutate=False):         # board.jl:451-509
    if color is None:
        color = pos.to_play
    new_pos = pos if mutate else pos.copy()
    assert not new_pos.done
    if c is None:
        return pass_move(new_pos, mutate=mutate)
    if not is_move_legal(pos, c):
        raise IllegalMove()
    potential_ko = is_koish(new_pos.board, c, pos.env)
    new_pos.board[c] = color
    captured = new_pos.lib_tracker.add_stone(color, c, pos.env)
    for s in captured:
        new_pos.board[s] = EMPTY
    opp_color = -color
    N = pos.env.N
    delta = np.zeros((N, N), dtype=np.int8)
    delta[c] = color
    for s in captured:
        delta[s] = color
    new_ko = next(iter(captured)) if (len(captured) == 1 and potential_ko == opp_color) else None
    if new_pos.to_play == BLACK:
        new_caps = (new_pos.caps[0] + len(captured), new_pos.caps[1])
    else:
        new_caps = (new_pos.caps[0], new_pos.caps[1] + len(captured))
    new_pos.n += 1
    new_pos.caps = new_caps
    new_pos.ko = new_ko
    new_pos.recent.append(PlayerMove(color, c))
    new_pos.board_deltas = np.concatenate(
        [delta.reshape(1, N, N), get_first_n(new_pos.board_deltas, new_pos.env.planes - 2)], axis=0)
    new_pos.to_play *= -1
    return new_pos


def score(pos):                                          # board.jl:511-533
    wb = pos.board.copy()
    N = pos.env.N
    while (wb == EMPTY).any():
        flat = wb.flatten(order="F")
        c = coordify(int(np.argmax(flat == EMPTY)), N)
        territory, borders = find_reached(wb, c, pos.env)
        border_colors = {int(wb[b]) for b in borders}
        x_border = BLACK in border_colors
        o_border = WHITE in border_colors
        if x_border and not o_border:
            tcolor = BLACK
        elif o_border and not x_border:
            tcolor = WHITE
        else:
            tcolor = UNKNOWN
        for s in territory:
            wb[s] = tcolor
    # Int - Float32 komi -> Float32 in the reference
    return np.float32(int((wb == BLACK).sum()) - int((wb == WHITE).sum())) - pos.komi


def result(pos):                                         # board.jl:535-544
    points = score(pos)
    return 1 if points > 0 else (-1 if points < 0 else 0)


def result_string(pos):                                  # board.jl:546-555
    points = score(pos)
    if points > 0:
        return "B+%.1f" % points
    if points < 0:
        return "W+%.1f" % abs(points)
    return "DRAW"


class PositionWithContext:                               # AlphaGo.jl:12-16
    def __init__(self, position, next_move, result):
        self.position, self.next_move, self.result = position, next_move, result


def replay_position(pos, result):                        # board.jl:557-578
    if pos.n != len(pos.recent):
        raise AssertionError("GoPosition history is incomplete")
    out = []
    dummy = GoPosition(pos.env, komi=pos.komi)
    for pm in pos.recent:
        out.append(PositionWithContext(dummy, pm.move, result))
        dummy = play_move(dummy, pm.move, color=pm.color)
    return out
