"""1:1 port of the reference's test/test_go.jl onto the CPU oracle (pins oracle/go.py)."""
import numpy as np
import pytest

from oracle import go
from oracle.go import BLACK, WHITE, EMPTY, MISSING_GROUP_ID, GoPosition, PlayerMove, LibertyTracker
from refboards import load_board, pc_set, EMPTY_ROW9 as EMPTY_ROW, assert_equal_positions

env = go.GoEnv(9)
kgs = lambda s: go.from_kgs(s, env)
TEST_BOARD = load_board(".X.....OO\nX........\n" + EMPTY_ROW * 7, env)


def test_load_board():                                   # test_go.jl:19-22
    assert (env.empty_board() == np.zeros((9, 9))).all()
    assert (env.empty_board() == load_board(". \n" * 81, env)).all()


def test_parsing():                                      # :24-30 (0-based here)
    assert kgs("A9") == (0, 0)
    assert go.from_sgf("aa") == (0, 0)
    assert kgs("A3") == (6, 0)
    assert go.from_sgf("ac") == (2, 0)
    assert kgs("D4") == go.from_sgf("df")


def test_neighbors():                                    # :32-40
    assert len(env.NEIGHBORS[kgs("A1")]) == 2
    assert len(env.NEIGHBORS[kgs("A2")]) == 3


def test_is_koish():                                     # :42-47
    assert go.is_koish(TEST_BOARD, kgs("A9"), env) == BLACK
    assert go.is_koish(TEST_BOARD, kgs("B8"), env) is None
    assert go.is_koish(TEST_BOARD, kgs("B9"), env) is None
    assert go.is_koish(TEST_BOARD, kgs("E5"), env) is None


def test_is_eyeish():                                    # :49-73
    board = load_board("""
        .XX...XXX
        X.X...X.X
        XX.....X.
        ........X
        XXXX.....
        OOOX....O
        X.OXX.OO.
        .XO.X.O.O
        XXO.X.OO.
    """, env)
    for be in pc_set("A2 A9 B8 J7 H8", env):
        assert go.is_eyeish(board, be, env) == BLACK, be
    for we in pc_set("H2 J1 J3", env):
        assert go.is_eyeish(board, we, env) == WHITE, we
    for ne in pc_set("B3 E5", env):
        assert go.is_eyeish(board, ne, env) is None, ne


def test_lib_tracker_init():                             # :74-85
    board = load_board("X........" + EMPTY_ROW * 8, env)
    lt = LibertyTracker.from_board(board, env)
    assert len(lt.groups) == 1
    assert lt.group_index[kgs("A9")] != MISSING_GROUP_ID
    assert lt.liberty_cache[kgs("A9")] == 2
    g = lt.groups[int(lt.group_index[kgs("A9")])]
    assert g.stones == pc_set("A9", env)
    assert g.liberties == pc_set("B9 A8", env)
    assert g.color == BLACK


def test_place_stone():                                  # :87-99
    board = load_board("X........" + EMPTY_ROW * 8, env)
    lt = LibertyTracker.from_board(board, env)
    lt.add_stone(BLACK, kgs("B9"), env)
    assert len(lt.groups) == 1
    assert lt.group_index[kgs("A9")] != MISSING_GROUP_ID
    assert lt.liberty_cache[kgs("A9")] == 3
    assert lt.liberty_cache[kgs("B9")] == 3
    g = lt.groups[int(lt.group_index[kgs("A9")])]
    assert g.stones == pc_set("A9 B9", env)
    assert g.liberties == pc_set("C9 A8 B8", env)
    assert g.color == BLACK


def test_place_stone_opposite_color():                   # :101-118
    board = load_board("X........" + EMPTY_ROW * 8, env)
    lt = LibertyTracker.from_board(board, env)
    lt.add_stone(WHITE, kgs("B9"), env)
    assert len(lt.groups) == 2
    assert lt.group_index[kgs("A9")] != MISSING_GROUP_ID
    assert lt.group_index[kgs("B9")] != MISSING_GROUP_ID
    assert lt.liberty_cache[kgs("A9")] == 1
    assert lt.liberty_cache[kgs("B9")] == 2
    bg = lt.groups[int(lt.group_index[kgs("A9")])]
    wg = lt.groups[int(lt.group_index[kgs("B9")])]
    assert bg.stones == pc_set("A9", env) and bg.liberties == pc_set("A8", env) and bg.color == BLACK
    assert wg.stones == pc_set("B9", env) and wg.liberties == pc_set("C9 B8", env) and wg.color == WHITE


def test_merge_multiple_groups():                        # :120-139
    board = load_board(".X.......\nX.X......\n.X.......\n" + EMPTY_ROW * 6, env)
    lt = LibertyTracker.from_board(board, env)
    lt.add_stone(BLACK, kgs("B8"), env)
    assert len(lt.groups) == 1
    assert lt.group_index[kgs("B8")] != MISSING_GROUP_ID
    g = lt.groups[int(lt.group_index[kgs("B8")])]
    assert g.stones == pc_set("B9 A8 B8 C8 B7", env)
    assert g.liberties == pc_set("A9 C9 D8 A7 C7 B6", env)
    assert g.color == BLACK
    for s in g.stones:
        assert lt.liberty_cache[s] == 6, s


def test_capture_stone():                                # :141-152
    board = load_board(".X.......\nXO.......\n.X.......\n" + EMPTY_ROW * 6, env)
    lt = LibertyTracker.from_board(board, env)
    captured = lt.add_stone(BLACK, kgs("C8"), env)
    assert len(lt.groups) == 4
    assert lt.group_index[kgs("B8")] == MISSING_GROUP_ID
    assert captured == pc_set("B8", env)


def test_capture_many():                                 # :154-198
    board = load_board(".XX......\nXOO......\n.XX......\n" + EMPTY_ROW * 6, env)
    lt = LibertyTracker.from_board(board, env)
    captured = lt.add_stone(BLACK, kgs("D8"), env)
    assert len(lt.groups) == 4
    assert lt.group_index[kgs("B8")] == MISSING_GROUP_ID
    assert captured == pc_set("B8 C8", env)
    grp = lambda s: lt.groups[int(lt.group_index[kgs(s)])]
    left, right, top, bottom = grp("A8"), grp("D8"), grp("B9"), grp("B7")
    assert left.stones == pc_set("A8", env) and left.liberties == pc_set("A9 B8 A7", env)
    assert right.stones == pc_set("D8", env) and right.liberties == pc_set("D9 C8 E8 D7", env)
    assert top.stones == pc_set("B9 C9", env) and top.liberties == pc_set("A9 D9 B8 C8", env)
    assert bottom.stones == pc_set("B7 C7", env) and bottom.liberties == pc_set("B8 C8 A7 D7 B6 C6", env)
    for g, n in ((top, 4), (left, 3), (right, 4), (bottom, 6)):
        for s in g.stones:
            assert lt.liberty_cache[s] == n, s
    for s in captured:
        assert lt.liberty_cache[s] == 0, s


def test_capture_multiple_groups():                      # :200-226
    board = load_board(".OX......\nOXX......\nXX.......\n" + EMPTY_ROW * 6, env)
    lt = LibertyTracker.from_board(board, env)
    captured = lt.add_stone(BLACK, kgs("A9"), env)
    assert len(lt.groups) == 2
    assert captured == pc_set("B9 A8", env)
    corner = lt.groups[int(lt.group_index[kgs("A9")])]
    assert corner.stones == pc_set("A9", env) and corner.liberties == pc_set("B9 A8", env)
    sur = lt.groups[int(lt.group_index[kgs("C9")])]
    assert sur.stones == pc_set("C9 B8 C8 A7 B7", env)
    assert sur.liberties == pc_set("B9 D9 A8 D8 C7 A6 B6", env)
    for s in corner.stones:
        assert lt.liberty_cache[s] == 2
    for s in sur.stones:
        assert lt.liberty_cache[s] == 7


def test_same_friendly_group_neighboring_twice():        # :228-242
    board = load_board("XX.......\nX........\n" + EMPTY_ROW * 7, env)
    lt = LibertyTracker.from_board(board, env)
    captured = lt.add_stone(BLACK, kgs("B8"), env)
    assert len(lt.groups) == 1
    g = lt.groups[int(lt.group_index[kgs("A9")])]
    assert g.stones == pc_set("A9 B9 A8 B8", env)
    assert g.liberties == pc_set("C9 C8 A7 B7", env)
    assert captured == set()


def test_same_opponent_group_neighboring_twice():        # :244-262
    board = load_board("XX.......\nX........\n" + EMPTY_ROW * 7, env)
    lt = LibertyTracker.from_board(board, env)
    captured = lt.add_stone(WHITE, kgs("B8"), env)
    assert len(lt.groups) == 2
    bg = lt.groups[int(lt.group_index[kgs("A9")])]
    assert bg.stones == pc_set("A9 B9 A8", env) and bg.liberties == pc_set("C9 A7", env)
    wg = lt.groups[int(lt.group_index[kgs("B8")])]
    assert wg.stones == pc_set("B8", env) and wg.liberties == pc_set("C8 B7", env)
    assert captured == set()


def test_passing():                                      # :264-285
    start = GoPosition(env, board=TEST_BOARD.copy(), n=0, komi=6.5, caps=(1, 2), ko=kgs("A1"), recent=[], to_play=BLACK)
    expected = GoPosition(env, board=TEST_BOARD.copy(), n=1, komi=6.5, caps=(1, 2), ko=None,
                          recent=[PlayerMove(BLACK, None)], to_play=WHITE)
    assert_equal_positions(go.pass_move(start), expected)


def test_flipturn():                                     # :287-308
    start = GoPosition(env, board=TEST_BOARD.copy(), n=0, komi=6.5, caps=(1, 2), ko=kgs("A1"), recent=[], to_play=BLACK)
    expected = GoPosition(env, board=TEST_BOARD.copy(), n=0, komi=6.5, caps=(1, 2), ko=None, recent=[], to_play=WHITE)
    assert_equal_positions(go.flip_playerturn(start), expected)


SUICIDE_BOARD = """
    ...O.O...
    ....O....
    XO.....O.
    OXO...OXO
    O.XO.OX.O
    OXO...OOX
    XO.......
    ......XXO
    .....XOO.
"""


def test_is_move_suicidal():                             # :310-336
    pos = GoPosition(env, board=load_board(SUICIDE_BOARD, env), to_play=BLACK)
    for mv in pc_set("E9 H5", env):
        assert pos.board[mv] == EMPTY
        assert go.is_move_suicidal(pos, mv), mv
    for mv in pc_set("B5 J1 A9", env):
        assert pos.board[mv] == EMPTY
        assert not go.is_move_suicidal(pos, mv), mv


LEGAL_BOARD = """
    .O.O.XOX.
    O..OOOOOX
    ......O.O
    OO.....OX
    XO.....X.
    .O.......
    OX.....OO
    XX...OOOX
    .....O.X.
"""


def test_legal_moves():                                  # :338-378
    board = load_board(LEGAL_BOARD, env)
    for b, tp in ((board, BLACK), (-board, WHITE)):
        pos = GoPosition(env, board=b.copy(), to_play=tp)
        for mv in pc_set("A9 E9 J9", env):
            assert not go.is_move_legal(pos, mv)
        for mv in pc_set("A4 G1 J1 H7", env):
            assert go.is_move_legal(pos, mv)
        bulk = go.all_legal_moves(pos)
        for i, bl in enumerate(bulk):
            assert go.is_move_legal(pos, go.from_flat(i, env)) == bool(bl)


def test_move():                                         # :380-421
    start = GoPosition(env, board=TEST_BOARD.copy(), n=0, komi=6.5, caps=(1, 2), ko=None, recent=[], to_play=BLACK)
    eb = load_board(".XX....OO\nX........\n" + EMPTY_ROW * 7, env)
    expected = GoPosition(env, board=eb, n=1, komi=6.5, caps=(1, 2), ko=None,
                          recent=[PlayerMove(BLACK, kgs("C9"))], to_play=WHITE)
    actual = go.play_move(start, kgs("C9"))
    assert_equal_positions(actual, expected)
    eb2 = load_board(".XX....OO\nX.......O\n" + EMPTY_ROW * 7, env)
    expected2 = GoPosition(env, board=eb2, n=2, komi=6.5, caps=(1, 2), ko=None,
                           recent=[PlayerMove(BLACK, kgs("C9")), PlayerMove(WHITE, kgs("J8"))], to_play=BLACK)
    actual2 = go.play_move(actual, kgs("J8"))
    assert_equal_positions(actual2, expected2)


def test_move_with_capture():                            # :423-456
    sb = load_board(EMPTY_ROW * 5 + "XXXX.....\nXOOX.....\nO.OX.....\nOOXX.....\n", env)
    start = GoPosition(env, board=sb, n=0, komi=6.5, caps=(1, 2), ko=None, recent=[], to_play=BLACK)
    eb = load_board(EMPTY_ROW * 5 + "XXXX.....\nX..X.....\n.X.X.....\n..XX.....\n", env)
    expected = GoPosition(env, board=eb, n=1, komi=6.5, caps=(7, 2), ko=None,
                          recent=[PlayerMove(BLACK, kgs("B2"))], to_play=WHITE)
    assert_equal_positions(go.play_move(start, kgs("B2")), expected)


def test_ko_move():                                      # :458-507
    sb = load_board(".OX......\nOX.......\n" + EMPTY_ROW * 7, env)
    start = GoPosition(env, board=sb.copy(), n=0, komi=6.5, caps=(1, 2), ko=None, recent=[], to_play=BLACK)
    eb = load_board("X.X......\nOX.......\n" + EMPTY_ROW * 7, env)
    expected = GoPosition(env, board=eb, n=1, komi=6.5, caps=(2, 2), ko=kgs("B9"),
                          recent=[PlayerMove(BLACK, kgs("A9"))], to_play=WHITE)
    actual = go.play_move(start, kgs("A9"))
    assert_equal_positions(actual, expected)
    with pytest.raises(go.IllegalMove):
        go.play_move(actual, kgs("B9"))
    pass_twice = go.pass_move(go.pass_move(actual))
    retake = go.play_move(pass_twice, kgs("B9"))
    expected = GoPosition(env, board=sb.copy(), n=4, komi=6.5, caps=(2, 3), ko=kgs("A9"),
                          recent=[PlayerMove(BLACK, kgs("A9")), PlayerMove(WHITE, None), PlayerMove(BLACK, None),
                                  PlayerMove(WHITE, kgs("B9"))], to_play=BLACK)
    assert_equal_positions(retake, expected)


def test_is_game_over():                                 # :509-516
    root = GoPosition(env)
    assert not root.done
    first = go.play_move(root, None)
    assert not first.done
    second = go.play_move(first, None)
    assert second.done


SCORE_BOARD_1 = """
    .XX......
    OOXX.....
    OOOX...X.
    OXX......
    OOXXXXXX.
    OOOXOXOXX
    .O.OOXOOX
    .O.O.OOXX
    ......OOO
"""
SCORE_BOARD_2 = SCORE_BOARD_1.replace(".XX......", "XXX......", 1)


def test_scoring():                                      # :518-564
    p1 = GoPosition(env, board=load_board(SCORE_BOARD_1, env), n=54, komi=6.5, caps=(2, 5), to_play=BLACK)
    assert go.score(p1) == 1.5
    p2 = GoPosition(env, board=load_board(SCORE_BOARD_2, env), n=55, komi=6.5, caps=(2, 5), to_play=WHITE)
    assert go.score(p2) == 2.5
