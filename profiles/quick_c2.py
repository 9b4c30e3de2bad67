"""Quick C2 throughput probe (device-resident moves/s only) -- used to compare engine knobs (env vars) run against run.
   python profiles/quick_c2.py [steps] [games] [warmup_steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

agz = pkg.load()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
games = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3
env = agz.GoEnv(9)
nn = agz.NeuralNet(env, tower_height=6, seed=0)
eng = agz.Engine(9, n_games=games, readouts=400, tower_height=6, seed=0, evaluator=agz.EVAL_NN_TC)
nn.push(eng)
eng.selfplay_start(-1)
for _ in range(warm):
    pr = eng.selfplay_step(50)
m0 = pr.moves_played
t0 = time.perf_counter()
dev = 0.0
for _ in range(steps):
    pr = eng.selfplay_step(50)
    dev += pr.step_ms
wall = time.perf_counter() - t0
eng.set_timing(True)
eng.phase_times(reset=True)
eng.selfplay_step(50)
kms, kln = eng.phase_times(reset=True)
eng.set_timing(False)
knobs = {k: v for k, v in os.environ.items() if k.startswith("AGZ_")}
print(json.dumps({"knobs": knobs, "moves_per_s": (pr.moves_played - m0) / wall, "ms_per_round_wall": 1e3 * wall / (50 * steps),
                  "ms_per_round_device": dev / (50 * steps), "error": pr.error,
                  "kernel_ms_per_round": {n: round(kms[i] / max(1, kln[0]), 4) for i, n in enumerate(agz.binding.KERNEL_NAMES)}}), flush=True)
eng.close()
