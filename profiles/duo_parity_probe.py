import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_abi_mcts as T
from backends import lib_for, agz
env = agz.GoEnv(9, lib_path=lib_for("cuda"))
T.check_selfplay_parity(env, 9, 24, seeds=[0], options={"tree.duo": 1}); print("a")
T.check_selfplay_parity(env, 9, 16, seeds=[2], priors_seed=7, value=-0.2, n_games=3, options={"tree.duo": 1}); print("b")
T.check_selfplay_parity(env, 9, 24, seeds=[0], n_games=2, options={"tree.duo": 1, "dummy.fused_rounds": 0}); print("c")
T.check_selfplay_parity(env, 9, 16, seeds=[5], priors_seed=2, value=0.05, n_games=2, nodes_per_game=96, options={"tree.duo": 1}); print("d")
T.check_selfplay_parity(env, 9, 400, seeds=[0], n_games=2, options={"tree.duo": 1}); print("e")
T.check_selfplay_parity(env, 9, 32, seeds=[11], priors_seed=5, value=0.1, n_games=24, concurrent=7, options={"tree.duo": 1}); print("f")
