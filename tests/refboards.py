"""Fixtures shared by the oracle tests and the C-ABI (GPU) tests: ports of the
reference's test helpers (test/test_utils.jl) and board fixtures."""
import re

import numpy as np

from oracle import go


def load_board(s, env):
    """test/test_utils.jl:1-18: row = text line, col = char; X=+1, O=-1."""
    rev = {"X": 1, "O": -1, ".": 0, "#": 2}
    s = re.sub(r"[^XO\.#]+", "", s)
    assert len(s) == env.N ** 2
    b = np.array([rev[ch] for ch in s], dtype=np.int8).reshape(env.N, env.N)  # [row, col]
    return b


def pc_set(string, env):
    return {go.from_kgs(t, env) for t in string.split()}


EMPTY_ROW9 = "." * 9 + "\n"

ALMOST_DONE_BOARD = """
.XO.XO.OO
X.XXOOOO.
XXXXXOOOO
XXXXXOOOO
.XXXXOOO.
XXXXXOOOO
.XXXXOOO.
XXXXXOOOO
XXXXOOOOO
"""

TT_FTW_BOARD = """
.XXOOOOOO
X.XOO...O
.XXOO...O
X.XOO...O
.XXOO..OO
X.XOOOOOO
.XXOOOOOO
X.XXXXXXX
XXXXXXXXX
"""


def lib_tracker_canon(lt):
    """Group-id-invariant view of a liberty tracker (test_utils.jl:27-60)."""
    mapping = {}
    gi = lt.group_index.flatten(order="F")
    for g in gi:
        g = int(g)
        if g != go.MISSING_GROUP_ID and g not in mapping:
            mapping[g] = len(mapping)
    remapped = [mapping.get(int(g), go.MISSING_GROUP_ID) for g in gi]
    groups = {mapping.get(gid, 0): (frozenset(g.stones), frozenset(g.liberties), g.color) for gid, g in lt.groups.items()}
    return remapped, groups, lt.liberty_cache.copy()


def assert_equal_positions(p1, p2):
    """test_utils.jl:62-74."""
    assert (p1.board == p2.board).all()
    a, b = lib_tracker_canon(p1.lib_tracker), lib_tracker_canon(p2.lib_tracker)
    assert a[0] == b[0]
    assert a[1] == b[1]
    assert (a[2] == b[2]).all()
    assert p1.n == p2.n
    assert p1.caps == p2.caps
    assert p1.ko == p2.ko
    r = min(len(p1.recent), len(p2.recent))
    if r > 0:
        assert p1.recent[-r:] == p2.recent[-r:]
    assert p1.to_play == p2.to_play


def assert_no_pending_vlosses(root):
    """test_utils.jl:76-85."""
    queue = [root]
    while queue:
        cur = queue.pop()
        assert cur.losses_applied == 0
        queue.extend(cur.children.values())
