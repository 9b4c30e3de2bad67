"""C2 with a sharp policy (the dense policy weights of the random-init network scaled up): deep, narrow trees as with a trained
network -- long descent paths, large re-used subtrees.  Reports the per-kernel split, mean path length and the device error flag."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import pkg  # noqa: E402

agz = pkg.load()
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
env = agz.GoEnv(9)
nn = agz.NeuralNet(env, tower_height=6, seed=0)
nn.params[2][4] = (nn.params[2][4] * scale).astype(np.float32)      # Dense(2N^2 -> A) of the policy head
eng = agz.Engine(9, n_games=1024, readouts=400, tower_height=6, seed=0, evaluator=agz.EVAL_NN_TC)
nn.push(eng)
eng.selfplay_start(-1)
pr0 = eng.selfplay_step(50 * 3)
t0 = time.perf_counter()
pr = eng.selfplay_step(50 * steps)
dt = time.perf_counter() - t0
eng.set_timing(True)
eng.phase_times(reset=True)
eng.selfplay_step(50)
kms, kln = eng.phase_times(reset=True)
print(json.dumps({"config": "C2 with a sharp policy (policy dense weights x %g), %d move-steps" % (scale, steps),
                  "moves_per_s": (pr.moves_played - pr0.moves_played) / dt, "mean_path_len": (pr.path_nodes - pr0.path_nodes) / max(1, pr.readouts - pr0.readouts),
                  "leaf_fill": (pr.positions_evaluated - pr0.positions_evaluated) / (50.0 * steps * 8192), "games_finished": int(pr.games_finished), "error": int(pr.error),
                  "kernel_ms_per_round": {n: round(kms[i] / max(1, kln[0]), 4) for i, n in enumerate(agz.binding.KERNEL_NAMES)}}), flush=True)
eng.close()
