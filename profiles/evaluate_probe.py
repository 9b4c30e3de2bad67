"""Throughput of the batched gating match (agz_match_*): 256 concurrent 9x9 games between two random-init T=6 networks."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

agz = pkg.load()
G, RO = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 200
env = agz.GoEnv(9)
a = agz.NeuralNet(env, tower_height=6, seed=1)
b = agz.NeuralNet(env, tower_height=6, seed=2)
games = []
t0 = time.perf_counter()
ok = agz.evaluate(env, a, b, num_games=G, ro=RO, details=games)
dt = time.perf_counter() - t0
moves = sum(len(g.moves) for g in games)
print(json.dumps({"config": "evaluate: %d concurrent 9x9 games, T=6, %d readouts per move" % (G, RO), "seconds": dt, "moves": moves,
                  "moves_per_s": moves / dt, "games_per_s": G / dt, "black_wins": sum(g.black_won for g in games), "passes_gate": bool(ok),
                  "resigned": sum(g.result_string.endswith("+R") for g in games)}), flush=True)
