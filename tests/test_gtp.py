"""GTP handlers (src/gtp_engine.jl:12-89) over the engine: boardsize / clear_board / komi / play / genmove / final_score /
showboard / undo through the line protocol, with the reference tests' DummyNet as the network."""
import io

import numpy as np
import pytest

from backends import BACKENDS, agz, lib_for
from test_abi_mcts import DummyNet


@pytest.fixture(scope="module", params=BACKENDS)
def env(request):
    return agz.GoEnv(9, lib_path=lib_for(request.param))


def serve(env, text, **kw):
    G = agz.gtp
    out = io.StringIO()
    h = G.run(env, DummyNet(82), num_readouts=16, stdin=io.StringIO(text), stdout=out, **kw)
    return h, [r for r in out.getvalue().split("\n\n") if r != ""]


def test_protocol_and_play(env):
    h, r = serve(env, "protocol_version\n1 name\nboardsize 9\nboardsize 19\nclear_board\nkomi 5.5\nplay black D4\nplay white D4\n"
                      "play white E5\nshowboard\n7 genmove black\nundo\nfrobnicate\nknown_command genmove\nquit\nname\n")
    assert r[0] == "= 2" and r[1] == "=1 AlphaGo.jl on B200" and r[2] == "= "
    assert r[3].startswith("? unsupported board size")                      # gtp_engine.jl:22-24
    assert r[4] == "= " and r[5] == "= " and r[6] == "= "
    assert r[7].startswith("? illegal move")                                # D4 is taken
    assert r[8] == "= "
    assert "X" in r[9] and "O" in r[9]
    assert r[10].startswith("=7 ") and r[10][3:] not in ("", "resign")
    assert r[11].startswith("? Not Implemented") and r[12].startswith("? unknown command") and r[13] == "= true"
    assert len(r) == 15                                                     # nothing after quit
    pos = h._pos
    assert pos.komi == 5.5 and pos.board[agz.from_kgs("D4", env)] == 1 and pos.board[agz.from_kgs("E5", env)] == -1
    assert pos.n == 3 and pos.to_play == -1                                 # B D4, W E5, B genmove
    mv = agz.from_kgs(r[10][3:], env)
    assert mv is None or pos.board[mv] == 1
    root = h._player.root.position
    assert np.array_equal(root.board, pos.board) and root.to_play == pos.to_play


def test_out_of_turn_and_courtesy_pass(env):
    # two black moves in a row: the second flips the player to move (gtp_engine.jl:72-77)
    h, r = serve(env, "play black C3\nplay black G7\nplay white pass\ngenmove black\nfinal_score\n", courtesy_pass=True)
    assert r[:3] == ["= ", "= ", "= "]
    assert r[3] == "= pass"                                                 # courtesy pass after the opponent's pass (:48-55)
    assert h._pos.board[agz.from_kgs("C3", env)] == 1 and h._pos.board[agz.from_kgs("G7", env)] == 1
    assert h._player.is_done() and r[4].startswith("= B+")                  # two black stones, komi 6.5: Black owns the board


def test_genmove_plays_a_whole_game(env):
    h, r = serve(env, "".join("genmove %s\n" % ("b" if k % 2 == 0 else "w") for k in range(40)) + "final_score\n")
    moves = [x[2:] for x in r[:-1]]
    assert all(m == "resign" or m == "pass" or agz.from_kgs(m, env) is not None for m in moves)
    assert r[-1][2:3] in ("B", "W", "D")
