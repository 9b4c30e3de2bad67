"""alphago.jl_b200 -- host-side mirror of AlphaGo.jl's self-play surface over libagz.so (hand-written sm_100a
CUDA behind the C ABI of include/agz.h).  Importing this package never falls back to a CPU implementation:
creating an Engine without libagz.so / without a B200 raises."""
from .binding import (Engine, Config, Position, NodeView, GameRecord, IllegalMove, AgzError, load_library, LIB_PATH,
                      SYMBOLS, EVAL_DUMMY, EVAL_NN_TC, EVAL_NN_F32, BN_VAR_EPS, BN_STD, CHAIN_BASE, CHAIN_VALUE,
                      CHAIN_POLICY, GAME_GO, GAME_GOMOKU)
from . import api
from . import gtp
from .api import (GoEnv, GoPosition, GomokuEnv, GomokuPosition, Go, result_string, NeuralNet, MCTSPlayer, MCTSNode, selfplay, evaluate, train, play, MatchGame, initialize_game, tree_search, pick_move,
                  play_move, should_resign, is_done, set_result, extract_data, select_leaf, incorporate_results,
                  maybe_add_child, add_virtual_loss, revert_virtual_loss, inject_noise, to_flat, from_flat, from_kgs,
                  to_kgs, all_legal_moves, score, result, get_feats, BLACK, WHITE, EMPTY)
