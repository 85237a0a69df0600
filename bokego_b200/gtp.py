"""Go Text Protocol front-end over the batched tree search.

Answers the same commands with the same texts as the reference's engine shell (/root/reference/bokego/gtp.py:16-330 --
command set gtp.py:36-43, reply format `=id text` / `?id text` gtp.py:327-330), including its observable quirks: the move
history survives `clear_board`, `loadsgf` replays the whole file whatever move number is given, a colour moving twice in a row
inserts a pass that is not listed in the history, only one `undo` is allowed.  tests/test_gtp.py replays a scripted session
recorded from the unmodified reference (tests/golden/gtp_transcript.json) and compares the replies byte for byte.

Differences, all on the search side: positions are bokego_b200.go.Game objects and the search is bokego_b200.mcts.MCTS
(device-resident position pool, leaves evaluated in batches); it is created lazily, so every command that needs no search
works without a GPU.  Like the reference (whose N / V dicts keep every node, gtp.py:332-337, mcts.py:110-131) the tree is kept
from move to move: the engine's own move re-roots with `choose`, an opponent move with `advance`, and only a move that is not in
the tree (a pass, a position set by clear_board / undo / loadsgf) starts a new one; the resign test runs on the statistics the
root already has, before the search (gtp.py:344-356).  `pondering` is accepted and stored but there is no background search: a
rollout batch here is one blocking device round trip, so nothing is searched while waiting for input.  `analyze` (a streaming
Sabaki extension) and `clear_cache` are accepted but answer `?`.
"""
import copy
import os
import sys
from timeit import default_timer

from . import go

COLORS = ("black", "b", "w", "white")
COMMANDS = ("name", "boardsize", "clear_board", "komi", "play", "genmove", "reg_genmove", "final_score", "quit", "version",
            "showboard", "clear_cache", "last_move", "move_history", "undo", "help", "known_command", "protocol_version",
            "list_commands", "set_fixed_handicap", "printsgf", "loadsgf", "analyze", "pondering")


def _after(game, mv):
    """copy of `game` with `mv` played (Go_MCTS.make_move, mcts.py:340-346)"""
    g = copy.deepcopy(game)
    g.play_move(mv)
    return g


class GTP:
    """policy_net / value_net: nets of bokego_b200.nnet (or PackedNet blobs).  kwargs: time_lim (seconds per move, default 0 =
    use n_rollouts), n_rollouts (default 1600), leaf_batch (32), expand_thresh (100), exploration_weight (4.0), device."""

    colors, commands = COLORS, COMMANDS

    def __init__(self, policy_net=None, value_net=None, root=None, **kwargs):
        self.policy_net, self.value_net = policy_net, value_net
        self.time_lim = kwargs.pop("time_lim", 0)
        self.n_rollouts = kwargs.pop("n_rollouts", 1600)
        self.pondering = kwargs.pop("pondering", False)
        self.search_kwargs = kwargs
        self.root = root if root is not None else go.Game()
        self.tree = None
        self.running = False
        self._move_history = []
        self._last_root = None
        self._undid = False

    # ---- engine side ---------------------------------------------------------------------------------------------------
    MAX_TREE_NODES = 1 << 20        # a kept tree is dropped beyond this size (the node pool only grows)

    def _set_root(self, game):
        self.root = game
        self.tree = None

    def input_move(self, sq_c):
        node = _after(self.root, sq_c)
        self._last_root = self.root
        tree = self.tree
        self.root = node
        # keep the subtree of the move when the tree has it (gtp.py:332-337 -> set_root keeps the statistics)
        if tree is not None and (tree.n > self.MAX_TREE_NODES or not tree.advance(sq_c)):
            tree = None
        self.tree = tree
        self._move_history.append(sq_c)
        self._undid = False

    def search(self):
        """run the configured amount of search from the current position and return the tree"""
        from . import mcts
        if self.tree is None:
            self.tree = mcts.MCTS(self.root, self.policy_net, self.value_net, **self.search_kwargs)
        if self.time_lim:
            t0 = default_timer()
            while default_timer() < t0 + self.time_lim:
                self.tree.rollout(max(1, self.tree.leaf_batch))
        else:
            self.tree.rollout(self.n_rollouts)
        return self.tree

    def genmove(self, resign=None):
        """move for the player to move (gtp.py:344-366); go.RESIGN when the position is lost (winrate < 0.1 after move 50)"""
        if resign is not None:
            give_up = resign
        else:                                            # `surrender` (gtp.py:339-342) on the statistics from before the search
            give_up = self.tree is not None and self.tree.winrate() < 0.1 and self.root.turn > 50
        if give_up:
            self.running = False
            return go.RESIGN
        mv = self.search().best_move()
        self.input_move(mv)
        return mv

    # ---- protocol side -------------------------------------------------------------------------------------------------
    def send(self, cmd):
        """one GTP command -> the reply string (None once the engine has quit, like the reference)"""
        if not self.running or not cmd:
            return None
        words = cmd.lower().split()
        cmd_id = ""
        if words[0].isdigit():
            cmd_id, words = words[0], words[1:]
        name, args = words[0], words[1:]
        ok, out = False, ""
        handler = getattr(self, "_cmd_" + name, None) if name in COMMANDS else None
        if name not in COMMANDS:
            out = f"unknown command '{name}'"
        elif handler is not None:
            ok, out = handler(args, self.root.turn)
        return f"{'=' if ok else '?'}{cmd_id} {out}\n\n"

    def _cmd_protocol_version(self, a, turn): return True, "2"
    def _cmd_version(self, a, turn): return True, "0.3"
    def _cmd_name(self, a, turn): return True, "boke"
    def _cmd_help(self, a, turn): return True, "\n".join(COMMANDS)
    _cmd_list_commands = _cmd_help

    def _cmd_known_command(self, a, turn):
        return (True, "true" if a[0] in COMMANDS else "false") if len(a) == 1 else (False, "")

    def _cmd_boardsize(self, a, turn):
        return (True, "") if a == ["9"] else (False, "boke only plays on 9x9 board")

    def _cmd_clear_board(self, a, turn):
        self._set_root(go.Game())
        return True, ""

    def _cmd_komi(self, a, turn):
        if not a:
            return False, "usage: komi <num-komi>"
        try:
            self.root.komi = float(a[0])
        except ValueError:
            return False, "invalid komi value"
        return True, ""

    def _cmd_play(self, a, turn):
        if len(a) < 2 or a[0] not in COLORS:
            return False, "usage: play <color> <vertex>"
        if a[1] == "resign":
            self.running = False
            return True, ""
        try:
            mv = go.squash(a[1])
        except Exception:  # noqa: BLE001  (any malformed vertex)
            return False, "invalid coordinate"
        if (0 if "b" in a[0] else 1) != turn % 2:
            new = _after(self.root, go.PASS)          # the same colour again: the other side passes (not listed in the history)
            if not new.is_legal(mv):
                return False, "illegal move"
            self._last_root = self.root
            self._set_root(_after(new, mv))
            self._move_history.append(mv)
            self._undid = False
            return True, ""
        try:
            self.input_move(mv)
        except go.IllegalMove:
            return False, "illegal move"
        return True, ""

    def _cmd_showboard(self, a, turn): return True, "\n" + str(self.root)

    def _cmd_genmove(self, a, turn, regular=False):
        if len(a) != 1 or a[0] not in COLORS:
            return False, f"usage: {'reg_genmove' if regular else 'genmove'} <color>"
        if (0 if "b" in a[0] else 1) != turn % 2:
            self.input_move(go.PASS)
            self._undid = True
        mv = self.genmove(False if regular else None)
        if mv == go.RESIGN:
            return True, "resign"
        return True, go.unsquash(mv)

    def _cmd_reg_genmove(self, a, turn): return self._cmd_genmove(a, turn, regular=True)

    def _cmd_undo(self, a, turn):
        if self._undid or self._last_root is None:
            return False, "cannot undo"
        self._set_root(self._last_root)
        self._move_history.pop()
        self._last_root, self._undid = None, True
        return True, ""

    def _cmd_last_move(self, a, turn):
        mv = self.root.last_move
        if mv is None:
            return False, "no previous move known"
        return True, ("black " if turn % 2 == 1 else "white ") + go.unsquash(mv)

    def _cmd_quit(self, a, turn):
        self.running = False
        return True, ""

    def _cmd_clear_cache(self, a, turn):
        self.tree = None
        self._undid = True
        return False, ""

    def _cmd_final_score(self, a, turn):
        score = self.root.score()
        if abs(score) < 1e-4:
            return True, "0"
        return True, f"B+{score}" if score > 0 else f"W+{-score}"

    def _cmd_move_history(self, a, turn): return True, "\n".join(go.unsquash(self._move_history))

    def _cmd_set_fixed_handicap(self, a, turn):
        if len(a) != 1 or not a[0].isnumeric():
            return False, "usage: set_fixed_handicap <num-handicaps>"
        if self.root.board != go.EMPTY_BOARD:
            return False, "board is not empty"
        if not 1 < int(a[0]) <= 5:
            return False, "invalid number of handicaps"
        stones = go.FLOWERS9[:int(a[0])]
        self._set_root(go.Game(board=go.bulk_place_stones(go.BLACK, go.EMPTY_BOARD, stones), turn=1))
        return True, " ".join(go.unsquash(list(stones)))

    def _cmd_printsgf(self, a, turn):
        path = a[0] if len(a) == 1 else os.path.join(os.getcwd(), "bokego.sgf")
        return True, go.write_sgf(self._move_history, path, komi=self.root.komi)

    def _cmd_loadsgf(self, a, turn):
        if len(a) != 2 or not a[1].isnumeric():
            return False, "usage: loadsgf <path-to-sgf> <move-number>"
        try:
            for mv in go.get_moves(a[0]):
                self.input_move(mv)
        except IOError as e:
            return False, str(e)
        except go.IllegalMove:
            return False, "illegal move in sgf"
        return True, "black" if (int(a[1]) - 1) % 2 == 0 else "white"

    def _cmd_analyze(self, a, turn): return False, "analyze is not supported"

    def _cmd_pondering(self, a, turn):
        if len(a) != 1 or a[0] not in ("on", "off"):
            return False, "usage: pondering <on/off>"
        self.pondering = a[0] == "on"
        return True, ""

    def start(self, stream_in=None, stream_out=None):
        """blocking main loop over lines of stream_in (default stdin)"""
        stream_in, stream_out = stream_in or sys.stdin, stream_out or sys.stdout
        self.running = True
        for line in stream_in:
            out = self.send(line.strip())
            if out is not None:
                stream_out.write(out)
                stream_out.flush()
            if not self.running:
                break
