"""GTP front end over the engine: the command handlers of src/gtp_engine.jl (GTPHandler :12-80, KGSHandler :82-89) on top of
MCTSPlayer (api.py), plus the line protocol around them (the reference file defines the handlers only, and does not parse as
committed: unbalanced parentheses :8,:62, Python leftovers :41-42; the evident intent is restated).

Handlers, as in the reference: boardsize (must equal the env's N), clear_board (initialize_game! on an empty position with the
handler's komi), komi, play [color] move (out-of-turn colours flip the player to move, flip_playerturn! board.jl:442-447),
genmove [color] (courtesy pass, suggest_move, "resign" when should_resign, set_result! when the game ends), undo (not
implemented), final_score (result_string), showboard.  Everything that computes runs on the GPU behind the C ABI.
"""
import sys

from . import api
from .api import BLACK, WHITE


def translate_gtp_color(gtp_color):                      # gtp_engine.jl:4-8
    c = gtp_color.lower()
    if c in ("b", "black"):
        return BLACK
    if c in ("w", "white"):
        return WHITE
    raise ValueError("invalid color %s" % gtp_color)


class GTPHandler:                                        # gtp_engine.jl:12-20
    def __init__(self, player, courtesy_pass=False):
        self._komi = 6.5
        self._player = player
        self._courtesy_pass = courtesy_pass
        self._pos = None                                  # host-side mirror of the root position (board history for re-rooting)
        self.cmd_clear_board()

    env = property(lambda s: s._player.env)

    def _reroot(self, pos):
        self._pos = pos
        self._player.initialize_game(pos)

    def cmd_boardsize(self, n):                           # :22-24
        if int(n) != self.env.N:
            raise AssertionError("unsupported board size: %s" % n)

    def cmd_clear_board(self):                            # :26-27
        self._reroot(api.GoPosition(self.env, komi=self._komi))

    def cmd_komi(self, komi):                             # :29-31
        self._komi = float(komi)
        p = self._pos
        self._reroot(api.GoPosition(self.env, board=p.board, n=p.n, komi=self._komi, caps=p.caps, ko=p.ko, recent=p.recent,
                                    to_play=p.to_play, history=p.history))

    def _accomodate_out_of_turn(self, color):             # :72-77
        if translate_gtp_color(color) != self._pos.to_play:
            p = self._pos                                 # flip_playerturn!(pos; mutate = true): ko cleared, to_play flipped
            self._reroot(api.GoPosition(self.env, board=p.board, n=p.n, komi=p.komi, caps=p.caps, ko=None, recent=p.recent,
                                        to_play=-p.to_play, history=p.history))

    def _advance(self, move):
        if not self._player.play_move(move):
            return False
        self._pos = api.play_move(self._pos, move)
        return True

    def cmd_play(self, arg0, arg1=None):                  # :33-41
        if arg1 is None:
            move = arg0
        else:
            self._accomodate_out_of_turn(arg0)
            move = arg1
        return self._advance(api.from_kgs(move.lower() if move.lower() == "pass" else move, self.env))

    def cmd_genmove(self, color=None):                    # :43-66
        if color is not None:
            self._accomodate_out_of_turn(color)
        if self._courtesy_pass and self._pos.recent and self._pos.recent[-1][1] is None:
            self._advance(None)
            return "pass"
        move = self._player.suggest_move()
        if self._player.should_resign():
            self._player.set_result(-self._pos.to_play, True)
            return "resign"
        self._advance(move)
        if self._player.is_done():
            self._player.set_result(api.result(self._player.root.position), False)
        return api.to_kgs(move, self.env)

    def cmd_undo(self):                                   # :68
        raise NotImplementedError("Not Implemented")

    def cmd_final_score(self):                            # :70
        if not self._player.result_string:
            self._player.set_result(api.result(self._player.root.position), False)
        return self._player.result_string

    def cmd_showboard(self):                              # KGSHandler (:82-89)
        return "\n\n" + api._board_string(self._pos) + "\n\n"


KNOWN = ["protocol_version", "name", "version", "known_command", "list_commands", "quit", "boardsize", "clear_board", "komi", "play",
         "genmove", "undo", "final_score", "showboard"]


def handle_line(handler, line):
    """One GTP command line -> (response text, quit?).  Responses follow GTP 2: "=[id] result" or "?[id] message"."""
    line = line.split("#")[0].strip()
    if not line:
        return None, False
    parts = line.split()
    cid = ""
    if parts[0].isdigit():
        cid, parts = parts[0], parts[1:]
    cmd, args = parts[0].lower(), parts[1:]
    ok = lambda s="": ("=%s %s\n\n" % (cid, s), False)
    err = lambda s: ("?%s %s\n\n" % (cid, s), False)
    try:
        if cmd == "protocol_version":
            return ok("2")
        if cmd == "name":
            return ok("AlphaGo.jl on B200")
        if cmd == "version":
            return ok("0.2")
        if cmd == "known_command":
            return ok("true" if args and args[0] in KNOWN else "false")
        if cmd == "list_commands":
            return ok("\n".join(KNOWN))
        if cmd == "quit":
            return ("=%s \n\n" % cid, True)
        if cmd == "boardsize":
            handler.cmd_boardsize(int(args[0]))
            return ok()
        if cmd == "clear_board":
            handler.cmd_clear_board()
            return ok()
        if cmd == "komi":
            handler.cmd_komi(float(args[0]))
            return ok()
        if cmd == "play":
            return ok() if handler.cmd_play(*args[:2]) else err("illegal move")
        if cmd == "genmove":
            return ok(handler.cmd_genmove(*args[:1]))
        if cmd == "final_score":
            return ok(handler.cmd_final_score())
        if cmd == "showboard":
            return ok("\n" + handler.cmd_showboard().strip("\n"))   # a GTP response must not contain an empty line
        if cmd == "undo":
            handler.cmd_undo()
        return err("unknown command")
    except NotImplementedError as ex:
        return err(str(ex))
    except (AssertionError, ValueError, IndexError) as ex:
        return err(str(ex) or "syntax error")


def run(env, nn=None, tower_height=19, num_readouts=800, courtesy_pass=False, stdin=None, stdout=None, seed=0):
    """Serve GTP on stdin / stdout with an MCTSPlayer in two_player_mode (what play.jl:25-34 builds for a game against a human)."""
    stdin, stdout = stdin or sys.stdin, stdout or sys.stdout
    if nn is None:
        nn = api.NeuralNet(env, tower_height=tower_height, seed=seed)
    player = api.MCTSPlayer(env, nn, num_readouts=num_readouts, two_player_mode=True, seed=seed)
    handler = GTPHandler(player, courtesy_pass)
    for line in stdin:
        out, stop = handle_line(handler, line)
        if out is not None:
            stdout.write(out)
            stdout.flush()
        if stop:
            break
    return handler
