"""ctypes binding of libagz.so (include/agz.h).  The host language of the reference is Julia (absent from this
image); the same entry points are what its `ccall`s would bind.  No compute happens here."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libagz.so")

AGZ_OK, ERR_ILLEGAL_MOVE, ERR_ASSERT, ERR_CUDA, ERR_NCCL, ERR_ARG, ERR_CAPACITY = 0, 1, 2, 3, 4, 5, 6
EVAL_DUMMY, EVAL_NN_TC, EVAL_NN_F32 = 0, 1, 2
GAME_GO, GAME_GOMOKU = 0, 1
BN_VAR_EPS, BN_STD = 0, 1
CHAIN_BASE, CHAIN_VALUE, CHAIN_POLICY = 0, 1, 2
MAX_POINTS, MAX_ACTIONS, HIST = 361, 362, 7


class IllegalMove(Exception):
    """src/AlphaGo.jl:8"""


class AgzError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libagz error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("board_n", C.c_int32), ("planes", C.c_int32), ("filters", C.c_int32), ("tower_height", C.c_int32),
                ("c_puct", C.c_double), ("noise_weight", C.c_double), ("noise_alpha", C.c_double),
                ("max_game_length", C.c_int32), ("tau_threshold", C.c_int32), ("parallel_readouts", C.c_int32),
                ("max_parallel", C.c_int32), ("komi", C.c_float), ("resign_threshold", C.c_double),
                ("resign_disable_frac", C.c_double), ("n_games", C.c_int32), ("readouts", C.c_int32),
                ("nodes_per_game", C.c_int32), ("seed", C.c_uint64), ("device", C.c_int32), ("world_size", C.c_int32),
                ("rank", C.c_int32), ("record_ring", C.c_int32), ("evaluator", C.c_int32), ("inject_noise", C.c_int32),
                ("game", C.c_int32), ("n_in_row", C.c_int32)]


class Position(C.Structure):
    _fields_ = [("board", C.c_int8 * MAX_POINTS), ("hist", (C.c_int8 * MAX_POINTS) * HIST), ("n_hist", C.c_int32),
                ("n", C.c_int32), ("to_play", C.c_int32), ("ko", C.c_int32), ("last_move_pass", C.c_int32),
                ("done", C.c_int32), ("caps", C.c_int32 * 2), ("komi", C.c_float)]


class NodeView(C.Structure):
    _fields_ = [("parent", C.c_int32), ("fmove", C.c_int32), ("to_play", C.c_int32), ("n", C.c_int32), ("ko", C.c_int32),
                ("is_expanded", C.c_int32), ("done", C.c_int32), ("last_move_pass", C.c_int32), ("N", C.c_float),
                ("W", C.c_float), ("child_N", C.c_float * MAX_ACTIONS), ("child_W", C.c_float * MAX_ACTIONS),
                ("child_prior", C.c_float * MAX_ACTIONS), ("children", C.c_int32 * MAX_ACTIONS),
                ("legal", C.c_int8 * MAX_ACTIONS), ("board", C.c_int8 * MAX_POINTS),
                ("action_score", C.c_double * MAX_ACTIONS)]


class GameHeader(C.Structure):
    _fields_ = [("game_id", C.c_int64), ("n_moves", C.c_int32), ("result", C.c_int32), ("resigned", C.c_int32),
                ("final_score", C.c_float), ("resign_threshold", C.c_double)]


class Progress(C.Structure):
    _fields_ = [("moves_played", C.c_int64), ("games_finished", C.c_int64), ("games_started", C.c_int64),
                ("positions_evaluated", C.c_int64), ("readouts", C.c_int64), ("path_nodes", C.c_int64),
                ("games_live", C.c_int32), ("error", C.c_int32), ("step_ms", C.c_float), ("arena_prunes", C.c_int32)]


# every symbol include/agz.h declares (the CPU test-suite checks the built library exports all of them)
SYMBOLS = [
    "agz_config_default", "agz_config_default_game", "agz_engine_create", "agz_engine_destroy", "agz_last_error", "agz_version",
    "agz_net_set_params", "agz_net_set_bn_stats", "agz_net_param_count", "agz_net_bn_count", "agz_net_forward",
    "agz_features", "agz_set_dummy_evaluator", "agz_set_evaluator", "agz_selfplay_start", "agz_selfplay_step",
    "agz_selfplay_harvest", "agz_selfplay_run", "agz_replay_gather", "agz_replay_read", "agz_replay_sample", "agz_nccl_unique_id",
    "agz_nccl_init", "agz_tree_init", "agz_tree_select_leaf", "agz_tree_incorporate", "agz_tree_backup_value",
    "agz_tree_add_virtual_loss", "agz_tree_revert_virtual_loss", "agz_tree_maybe_add_child", "agz_tree_search",
    "agz_tree_inject_noise", "agz_tree_pick_move", "agz_tree_play_move", "agz_tree_should_resign", "agz_tree_root",
    "agz_tree_read_node", "agz_tree_set_stats", "agz_tree_pending_vlosses", "agz_tree_read_record",
    "agz_tree_node_features", "agz_pos_play_move", "agz_pos_legal_moves", "agz_pos_score", "agz_pos_liberties",
    "agz_kernel_launches", "agz_phase_times", "agz_set_timing", "agz_net_flops", "agz_trace_read", "agz_engine_info",
    "agz_match_start", "agz_match_search", "agz_match_play",
    "agz_replay_sample_hist", "agz_train_step", "agz_train_read_grads", "agz_net_get_params", "agz_net_get_bn_stats",
    "agz_set_option", "agz_get_option", "agz_replay_info", "agz_selftest_division", "agz_train_step_from_replay", "agz_net_forward_debug", "agz_selfplay_stats",
]
KERNEL_NAMES = ["select", "features", "stem_conv", "tower_conv", "heads", "incorporate"]

_libs = {}


def load_library(path=None):
    """Load libagz.so.  Fails loudly when it has not been built: there is no Python/CPU fallback."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                          "The engine has no CPU fallback." % path)
    lib = C.CDLL(path)
    lib.agz_last_error.restype = C.c_char_p
    lib.agz_last_error.argtypes = [C.c_void_p]
    lib.agz_net_param_count.restype = C.c_size_t
    lib.agz_net_bn_count.restype = C.c_size_t
    lib.agz_net_param_count.argtypes = [C.c_void_p, C.c_int32]
    lib.agz_net_bn_count.argtypes = [C.c_void_p, C.c_int32]
    lib.agz_engine_destroy.restype = None
    lib.agz_engine_destroy.argtypes = [C.c_void_p]
    _libs[path] = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Engine:
    """One agz_engine handle (one GPU)."""

    def __init__(self, board_n=9, lib_path=None, options=None, **overrides):
        """`overrides` are agz_config fields; `options` is a dict of named integer options (agz_set_option)."""
        self.lib = load_library(lib_path)
        self.cfg = Config()
        self._h = C.c_void_p()
        game, n_in_row = overrides.pop("game", GAME_GO), overrides.pop("n_in_row", 5)
        self._check(self.lib.agz_config_default_game(C.byref(self.cfg), game, board_n, n_in_row), None)   # GoEnv(N) / GomokuEnv(N, n_in_row)
        for k, v in overrides.items():
            if not hasattr(self.cfg, k):
                raise TypeError("unknown config field %r" % k)
            setattr(self.cfg, k, v)
        if "max_parallel" not in overrides:
            self.cfg.max_parallel = max(self.cfg.max_parallel, self.cfg.parallel_readouts)
        h = C.c_void_p()
        self._check(self.lib.agz_engine_create(C.byref(self.cfg), C.byref(h)), None)
        self._h = h
        self.N = self.cfg.board_n
        self.N2 = self.N * self.N
        self.A = self.N2 + (1 if self.cfg.game == GAME_GO else 0)   # env.action_space
        self.L = self.cfg.max_game_length + 2
        for k, v in (options or {}).items():
            self.set_option(k, v)

    def close(self):
        if self._h:
            self.lib.agz_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, h="self"):
        if rc == AGZ_OK:
            return
        msg = self.lib.agz_last_error(self._h if h == "self" else None)
        msg = msg.decode() if msg else ""
        if rc == ERR_ILLEGAL_MOVE:
            raise IllegalMove(msg)
        if rc == ERR_ASSERT:
            raise AssertionError(msg)
        raise AgzError(rc, msg)

    # ---- options (include/agz.h: agz_set_option)
    def set_option(self, key, value):
        self._check(self.lib.agz_set_option(self._h, key.encode(), C.c_int64(int(value))))

    def get_option(self, key):
        v = C.c_int64()
        self._check(self.lib.agz_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    # ---- network
    def net_param_count(self, chain):
        return self.lib.agz_net_param_count(self._h, chain)

    def net_bn_count(self, chain):
        return self.lib.agz_net_bn_count(self._h, chain)

    def net_set_params(self, chain, flat):
        flat = _f32(flat)
        self._check(self.lib.agz_net_set_params(self._h, C.c_int32(chain), _ptr(flat, C.c_float), C.c_size_t(flat.size)))

    def net_set_bn_stats(self, chain, mu, sigma, mode):
        mu, sigma = _f32(mu), _f32(sigma)
        self._check(self.lib.agz_net_set_bn_stats(self._h, C.c_int32(chain), _ptr(mu, C.c_float), _ptr(sigma, C.c_float),
                                                  C.c_size_t(mu.size), C.c_int32(mode)))

    def train_step(self, boards_hist, to_play, pis, zs, lr=0.02, momentum=0.9):
        bh = np.ascontiguousarray(boards_hist, dtype=np.int8)
        tp = np.ascontiguousarray(to_play, dtype=np.int8)
        pi = _f32(pis)
        z = np.ascontiguousarray(zs, dtype=np.int8)
        B = tp.shape[0]
        assert bh.shape == (B, 8, self.N2) and pi.shape == (B, self.A) and z.shape == (B,)
        loss = C.c_float()
        self._check(self.lib.agz_train_step(self._h, _ptr(bh, C.c_int8), _ptr(tp, C.c_int8), _ptr(pi, C.c_float), _ptr(z, C.c_int8),
                                            C.c_int32(B), C.c_float(lr), C.c_float(momentum), C.byref(loss)))
        return loss.value

    def train_step_from_replay(self, batch, seed=0, lr=0.02, momentum=0.9):
        loss = C.c_float()
        self._check(self.lib.agz_train_step_from_replay(self._h, C.c_int32(batch), C.c_uint64(seed), C.c_float(lr), C.c_float(momentum), C.byref(loss)))
        return loss.value

    def train_read_grads(self, chain):
        g = np.zeros(self.net_param_count(chain), np.float32)
        self._check(self.lib.agz_train_read_grads(self._h, C.c_int32(chain), _ptr(g, C.c_float), C.c_size_t(g.size)))
        return g

    def net_get_params(self, chain):
        p = np.zeros(self.net_param_count(chain), np.float32)
        self._check(self.lib.agz_net_get_params(self._h, C.c_int32(chain), _ptr(p, C.c_float), C.c_size_t(p.size)))
        return p

    def net_get_bn_stats(self, chain):
        n = self.net_bn_count(chain)
        mu, sg = np.zeros(n, np.float32), np.zeros(n, np.float32)
        mode = C.c_int32()
        self._check(self.lib.agz_net_get_bn_stats(self._h, C.c_int32(chain), _ptr(mu, C.c_float), _ptr(sg, C.c_float), C.c_size_t(n), C.byref(mode)))
        return mu, sg, mode.value

    def net_forward(self, evaluator, boards_hist, to_play):
        bh = np.ascontiguousarray(boards_hist, dtype=np.int8)
        tp = np.ascontiguousarray(to_play, dtype=np.int8)
        B = tp.shape[0]
        assert bh.shape == (B, 8, self.N2)
        pi = np.empty((B, self.A), np.float32)
        v = np.empty(B, np.float32)
        self._check(self.lib.agz_net_forward(self._h, C.c_int32(evaluator), _ptr(bh, C.c_int8), _ptr(tp, C.c_int8),
                                             C.c_int32(B), _ptr(pi, C.c_float), _ptr(v, C.c_float)))
        return pi, v

    def net_forward_debug(self, evaluator, boards_hist, to_play, n_blocks=-1, want_trunk=False):
        """Test hook: dict with pi, v, logits, v_pre (whole network only) and, if asked, the trunk [B][256][N2] after n_blocks blocks."""
        bh = np.ascontiguousarray(boards_hist, dtype=np.int8)
        tp = np.ascontiguousarray(to_play, dtype=np.int8)
        B = tp.shape[0]
        full = n_blocks < 0 or n_blocks >= self.cfg.tower_height
        out = {}
        if full:
            out = {"pi": np.empty((B, self.A), np.float32), "v": np.empty(B, np.float32), "logits": np.empty((B, self.A), np.float32),
                   "v_pre": np.empty(B, np.float32)}
        if want_trunk or not full:
            out["trunk"] = np.empty((B, 256, self.N2), np.float32)
        g = lambda k: _ptr(out.get(k), C.c_float)
        self._check(self.lib.agz_net_forward_debug(self._h, C.c_int32(evaluator), _ptr(bh, C.c_int8), _ptr(tp, C.c_int8), C.c_int32(B), C.c_int32(n_blocks),
                                                   g("pi"), g("v"), g("logits"), g("v_pre"), g("trunk")))
        return out

    def features(self, boards_hist, to_play):
        bh = np.ascontiguousarray(boards_hist, dtype=np.int8)
        tp = np.ascontiguousarray(to_play, dtype=np.int8)
        B = tp.shape[0]
        out = np.empty((B, 17, self.N2), np.float32)
        self._check(self.lib.agz_features(self._h, _ptr(bh, C.c_int8), _ptr(tp, C.c_int8), C.c_int32(B), _ptr(out, C.c_float)))
        return out

    def set_dummy_evaluator(self, priors=None, value=0.0):
        p = _f32(priors) if priors is not None else None
        self._check(self.lib.agz_set_dummy_evaluator(self._h, _ptr(p, C.c_float), C.c_float(value)))

    def set_evaluator(self, kind):
        self._check(self.lib.agz_set_evaluator(self._h, C.c_int32(kind)))

    # ---- self-play
    def selfplay_start(self, total_games=-1):
        self._check(self.lib.agz_selfplay_start(self._h, C.c_int64(total_games)))

    def selfplay_step(self, rounds, want_progress=True):
        pr = Progress()
        self._check(self.lib.agz_selfplay_step(self._h, C.c_int32(rounds), C.byref(pr) if want_progress else None))
        return pr if want_progress else None

    def _record_buffers(self, n):
        return ((GameHeader * n)(), np.zeros((n, self.L), np.int16), np.zeros((n, self.L), np.float32),
                np.zeros((n, self.L, self.A), np.float32), np.zeros((n, self.L, self.A), np.float32))

    def selfplay_harvest(self, max_records):
        hd, mv, q, pi, vis = self._record_buffers(max_records)
        n = C.c_int32()
        self._check(self.lib.agz_selfplay_harvest(self._h, C.c_int32(max_records), hd, _ptr(mv, C.c_int16), _ptr(q, C.c_float),
                                                  _ptr(pi, C.c_float), _ptr(vis, C.c_float), C.byref(n)))
        return [GameRecord(hd[i], mv[i], q[i], pi[i], vis[i], self.cfg.game) for i in range(n.value)]

    def selfplay_harvest_discard(self, max_records=1 << 20):
        """Release finished-game records without copying them to the host."""
        n = C.c_int32()
        self._check(self.lib.agz_selfplay_harvest(self._h, C.c_int32(max_records), None, None, None, None, None, C.byref(n)))
        return n.value

    def selfplay_run(self, total_games):
        mine = (total_games - self.cfg.rank + self.cfg.world_size - 1) // self.cfg.world_size
        hd, mv, q, pi, vis = self._record_buffers(max(mine, 1))
        self._check(self.lib.agz_selfplay_run(self._h, C.c_int32(total_games), hd, _ptr(mv, C.c_int16), _ptr(q, C.c_float),
                                              _ptr(pi, C.c_float), _ptr(vis, C.c_float)))
        return [GameRecord(hd[i], mv[i], q[i], pi[i], vis[i], self.cfg.game) for i in range(mine)]

    def replay_gather(self):
        n = C.c_int64()
        self._check(self.lib.agz_replay_gather(self._h, C.byref(n)))
        return n.value

    def replay_info(self):
        out = (C.c_int64 * 5)()
        self._check(self.lib.agz_replay_info(self._h, out))
        return {"tuple_bytes": out[0], "capacity": out[1], "total": out[2], "last_gather_bytes": out[3], "gathered_bytes": out[4]}

    def replay_read(self, first, count):
        boards = np.zeros((count, self.N2), np.int8)
        tp = np.zeros(count, np.int8)
        pis = np.zeros((count, self.A), np.float32)
        zs = np.zeros(count, np.int8)
        self._check(self.lib.agz_replay_read(self._h, C.c_int64(first), C.c_int32(count), _ptr(boards, C.c_int8), _ptr(tp, C.c_int8),
                                             _ptr(pis, C.c_float), _ptr(zs, C.c_int8)))
        return boards, tp, pis, zs

    def replay_sample(self, batch, seed=0):
        """get_replay_batch (src/train.jl:4-12): uniform, without replacement."""
        boards = np.zeros((batch, self.N2), np.int8)
        tp = np.zeros(batch, np.int8)
        pis = np.zeros((batch, self.A), np.float32)
        zs = np.zeros(batch, np.int8)
        idx = np.zeros(batch, np.int64)
        self._check(self.lib.agz_replay_sample(self._h, C.c_int32(batch), C.c_uint64(seed), _ptr(boards, C.c_int8), _ptr(tp, C.c_int8),
                                               _ptr(pis, C.c_float), _ptr(zs, C.c_int8), _ptr(idx, C.c_int64)))
        return boards, tp, pis, zs, idx

    def replay_sample_hist(self, batch, seed=0):
        """get_replay_batch with the 8-board history the network trains on: (boards_hist, to_play, pis, zs, indices)."""
        bh = np.zeros((batch, 8, self.N2), np.int8)
        tp = np.zeros(batch, np.int8)
        pis = np.zeros((batch, self.A), np.float32)
        zs = np.zeros(batch, np.int8)
        idx = np.zeros(batch, np.int64)
        self._check(self.lib.agz_replay_sample_hist(self._h, C.c_int32(batch), C.c_uint64(seed), _ptr(bh, C.c_int8), _ptr(tp, C.c_int8),
                                                    _ptr(pis, C.c_float), _ptr(zs, C.c_int8), _ptr(idx, C.c_int64)))
        return bh, tp, pis, zs, idx

    def nccl_unique_id(self):
        buf = (C.c_uint8 * 128)()
        self._check(self.lib.agz_nccl_unique_id(buf), None)
        return bytes(buf)

    def nccl_init(self, id_bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(id_bytes)
        self._check(self.lib.agz_nccl_init(self._h, buf))

    # ---- tree hooks
    def tree_init(self, slot, pos=None, game_id=0):
        self._check(self.lib.agz_tree_init(self._h, C.c_int32(slot), C.byref(pos) if pos is not None else None, C.c_int64(game_id)))

    def tree_select_leaf(self, slot, from_node=-1):
        leaf = C.c_int32()
        self._check(self.lib.agz_tree_select_leaf(self._h, C.c_int32(slot), C.c_int32(from_node), C.byref(leaf)))
        return leaf.value

    def tree_incorporate(self, slot, node, probs, value):
        p = _f32(probs)
        assert p.shape == (self.A,)
        self._check(self.lib.agz_tree_incorporate(self._h, C.c_int32(slot), C.c_int32(node), _ptr(p, C.c_float), C.c_float(value)))

    def tree_backup_value(self, slot, node, value):
        self._check(self.lib.agz_tree_backup_value(self._h, C.c_int32(slot), C.c_int32(node), C.c_float(value)))

    def tree_add_virtual_loss(self, slot, node):
        self._check(self.lib.agz_tree_add_virtual_loss(self._h, C.c_int32(slot), C.c_int32(node)))

    def tree_revert_virtual_loss(self, slot, node):
        self._check(self.lib.agz_tree_revert_virtual_loss(self._h, C.c_int32(slot), C.c_int32(node)))

    def tree_maybe_add_child(self, slot, node, fmove):
        ch = C.c_int32()
        self._check(self.lib.agz_tree_maybe_add_child(self._h, C.c_int32(slot), C.c_int32(node), C.c_int32(fmove), C.byref(ch)))
        return ch.value

    def tree_search(self, slot, parallel_readouts=8):
        n = C.c_int32()
        self._check(self.lib.agz_tree_search(self._h, C.c_int32(slot), C.c_int32(parallel_readouts), C.byref(n)))
        return n.value

    def tree_inject_noise(self, slot):
        self._check(self.lib.agz_tree_inject_noise(self._h, C.c_int32(slot)))

    def tree_pick_move(self, slot):
        mv = C.c_int32()
        self._check(self.lib.agz_tree_pick_move(self._h, C.c_int32(slot), C.byref(mv)))
        return mv.value

    def tree_play_move(self, slot, fmove):
        self._check(self.lib.agz_tree_play_move(self._h, C.c_int32(slot), C.c_int32(fmove)))

    def tree_should_resign(self, slot, threshold):
        y = C.c_int32()
        self._check(self.lib.agz_tree_should_resign(self._h, C.c_int32(slot), C.c_double(threshold), C.byref(y)))
        return bool(y.value)

    def tree_root(self, slot):
        r, n = C.c_int32(), C.c_int32()
        self._check(self.lib.agz_tree_root(self._h, C.c_int32(slot), C.byref(r), C.byref(n)))
        return r.value, n.value

    def tree_read_node(self, slot, node):
        nv = NodeView()
        self._check(self.lib.agz_tree_read_node(self._h, C.c_int32(slot), C.c_int32(node), C.byref(nv)))
        return nv

    def tree_set_stats(self, slot, node, self_N=None, child_N=None, n_override=None):
        sn = C.c_float(self_N) if self_N is not None else None
        cn = _f32(child_N) if child_N is not None else None
        no = C.c_int32(n_override) if n_override is not None else None
        self._check(self.lib.agz_tree_set_stats(self._h, C.c_int32(slot), C.c_int32(node), C.byref(sn) if sn is not None else None,
                                                _ptr(cn, C.c_float), C.byref(no) if no is not None else None))

    def tree_pending_vlosses(self, slot):
        p = C.c_int32()
        self._check(self.lib.agz_tree_pending_vlosses(self._h, C.c_int32(slot), C.byref(p)))
        return p.value

    def tree_read_record(self, slot):
        n = C.c_int32()
        mv = np.zeros(self.L, np.int16)
        q = np.zeros(self.L, np.float32)
        pi = np.zeros((self.L, self.A), np.float32)
        self._check(self.lib.agz_tree_read_record(self._h, C.c_int32(slot), C.byref(n), _ptr(mv, C.c_int16), _ptr(q, C.c_float), _ptr(pi, C.c_float)))
        return n.value, mv[:n.value], q[:n.value], pi[:n.value]

    def tree_node_features(self, slot, node):
        out = np.zeros((17, self.N2), np.float32)
        self._check(self.lib.agz_tree_node_features(self._h, C.c_int32(slot), C.c_int32(node), _ptr(out, C.c_float)))
        return out

    # ---- position hooks
    def pos_play_move(self, pos, fmove):
        out = Position()
        self._check(self.lib.agz_pos_play_move(self._h, C.byref(pos), C.c_int32(fmove), C.byref(out)))
        return out

    def pos_legal_moves(self, pos):
        legal = np.zeros(self.A, np.int8)
        self._check(self.lib.agz_pos_legal_moves(self._h, C.byref(pos), _ptr(legal, C.c_int8)))
        return legal

    def pos_score(self, pos):
        s = C.c_float()
        self._check(self.lib.agz_pos_score(self._h, C.byref(pos), C.byref(s)))
        return s.value

    def pos_liberties(self, pos):
        out = np.zeros(self.N2, np.uint8)
        self._check(self.lib.agz_pos_liberties(self._h, C.byref(pos), _ptr(out, C.c_uint8)))
        return out

    # ---- introspection
    def info(self):
        out = (C.c_int64 * 4)()
        self._check(self.lib.agz_engine_info(self._h, out))
        return {"nodes_per_game": out[0], "bytes_per_node": out[1], "n_games": out[2], "record_ring": out[3]}

    def selfplay_stats(self):
        out = (C.c_int64 * 4)()
        self._check(self.lib.agz_selfplay_stats(self._h, out))
        return {"duplicate_leaves": out[0], "arena_prunes": out[1], "positions_evaluated": out[2], "readouts": out[3]}

    def kernel_launches(self):
        n = C.c_int64()
        self._check(self.lib.agz_kernel_launches(self._h, C.byref(n)))
        return n.value

    def set_timing(self, on):
        self._check(self.lib.agz_set_timing(self._h, C.c_int32(1 if on else 0)))

    def phase_times(self, reset=True):
        ms = (C.c_float * 6)()
        ln = (C.c_int64 * 6)()
        self._check(self.lib.agz_phase_times(self._h, ms, ln, C.c_int32(1 if reset else 0)))
        return list(ms), list(ln)


    # ---- two-player matches (evaluate / play)
    def match_start(self, game_ids=None):
        ids = None if game_ids is None else np.ascontiguousarray(game_ids, np.int64)
        self._check(self.lib.agz_match_start(self._h, _ptr(ids, C.c_int64)))

    def match_search(self, active):
        G = self.cfg.n_games
        act = np.ascontiguousarray(active, np.uint8)
        assert act.size == G
        moves, res, sc = np.zeros(G, np.int32), np.zeros(G, np.int32), np.zeros(G, np.float32)
        self._check(self.lib.agz_match_search(self._h, _ptr(act, C.c_uint8), _ptr(moves, C.c_int32), _ptr(res, C.c_int32), _ptr(sc, C.c_float)))
        return moves, res.astype(bool), sc

    def match_play(self, moves):
        G = self.cfg.n_games
        mv = np.ascontiguousarray(moves, np.int32)
        assert mv.size == G
        done, sc = np.zeros(G, np.int32), np.zeros(G, np.float32)
        self._check(self.lib.agz_match_play(self._h, _ptr(mv, C.c_int32), _ptr(done, C.c_int32), _ptr(sc, C.c_float)))
        return done.astype(bool), sc

    def trace_read(self, max_records=1 << 16, reset=True):
        """Kernel timeline trace (set_option("trace.records", n)): array of (tag, block, grid, start_ns, end_ns, sm)."""
        buf = np.zeros((max_records, 4), np.uint64)
        n = C.c_int32()
        self._check(self.lib.agz_trace_read(self._h, _ptr(buf, C.c_uint64), C.c_int32(max_records), C.byref(n), C.c_int32(1 if reset else 0)))
        b = buf[:n.value]
        w0 = b[:, 0]
        return np.stack([w0 & np.uint64(0xFF), (w0 >> np.uint64(8)) & np.uint64(0xFFFFFF), w0 >> np.uint64(32), b[:, 1], b[:, 2], b[:, 3]], axis=1).astype(np.int64)

    def selftest_division(self, n_samples, seed=0):
        out = (C.c_uint64 * 2)()
        self._check(self.lib.agz_selftest_division(self._h, C.c_uint64(n_samples), C.c_uint64(seed), out))
        return int(out[0]), int(out[1])

    def net_flops(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.agz_net_flops(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


class GameRecord:
    """What `selfplay` leaves in the returned MCTSPlayer (searches_pi, qs, result, result_string) + the move list."""

    def __init__(self, hd, moves, qs, pis, visits, game=GAME_GO):
        n = hd.n_moves
        self.game = game
        self.game_id, self.n_moves, self.result, self.resigned = hd.game_id, n, hd.result, bool(hd.resigned)
        self.final_score, self.resign_threshold = hd.final_score, hd.resign_threshold
        self.moves = moves[:n].copy()
        self.qs = qs[:n].copy()
        self.searches_pi = pis[:n].copy()
        self.visits = visits[:n].copy()

    @property
    def result_string(self):          # set_result! (mcts_play.jl:100-108) / result_string (board.jl:546-555)
        if self.resigned:
            return "B+R" if self.result == 1 else "W+R"
        if self.game == GAME_GOMOKU:  # result_string (src/game/gomoku/board.jl:184-193); final_score = the winner's colour
            return "B" if self.final_score > 0 else ("W" if self.final_score < 0 else "DRAW")
        if self.final_score > 0:
            return "B+%.1f" % self.final_score
        if self.final_score < 0:
            return "W+%.1f" % abs(self.final_score)
        return "DRAW"
