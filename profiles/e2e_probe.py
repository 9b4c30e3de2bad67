"""Where the end-to-end step spends its time besides the 50 rounds: times each call of bench.py's e2e loop (parameter upload,
self-play step, replay gather, harvest) on the C2 workload after a short burn-in."""
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
env = agz.GoEnv(9)
nn = agz.NeuralNet(env, tower_height=6, seed=0)
eng = agz.Engine(9, n_games=1024, readouts=400, tower_height=6, seed=0, evaluator=agz.EVAL_NN_TC, options={"selfplay.stagger_rounds": 112 * 50})
nn.push(eng)
flat = [np.concatenate([np.asarray(a, np.float32).flatten(order="F") for a in nn.params[k]]) for k in range(3)]
eng.selfplay_start(-1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    eng.selfplay_step(50)
    eng.replay_gather()
    eng.selfplay_harvest_discard()
acc = {"set_params": 0.0, "step_wall": 0.0, "step_dev": 0.0, "gather": 0.0, "harvest": 0.0}
K = 10
for _ in range(K):
    t = time.perf_counter()
    for k in range(3):
        eng.net_set_params(k, flat[k])
    acc["set_params"] += time.perf_counter() - t
    t = time.perf_counter()
    pr = eng.selfplay_step(50)
    acc["step_wall"] += time.perf_counter() - t
    acc["step_dev"] += pr.step_ms / 1e3
    t = time.perf_counter()
    eng.replay_gather()
    acc["gather"] += time.perf_counter() - t
    t = time.perf_counter()
    recs = eng.selfplay_harvest(4096)
    acc["harvest"] += time.perf_counter() - t
print(json.dumps({k: round(1e3 * v / K, 3) for k, v in acc.items()}), "ms per step; finished per step:", len(recs))
eng.close()
