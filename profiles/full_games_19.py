"""19x19 games played to the end (up to 505 moves) on the tensor-core network: invariants only -- no device error, every finished
game's moves replay legally through the oracle's rules and end in the recorded result."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import pkg  # noqa: E402
from oracle import go as ogo  # noqa: E402

agz = pkg.load()
env = agz.GoEnv(19)
nn = agz.NeuralNet(env, tower_height=2, seed=3)
eng = agz.Engine(19, n_games=64, readouts=32, tower_height=2, seed=2, evaluator=agz.EVAL_NN_TC)
nn.push(eng)
eng.selfplay_start(64)
t0 = time.perf_counter()
recs = []
for it in range(100000):
    pr = eng.selfplay_step(16)
    eng.replay_gather()
    recs += eng.selfplay_harvest(64)
    if pr.error or pr.games_finished == 64:
        break
dt = time.perf_counter() - t0
oenv = ogo.GoEnv(19)
checked = 0
for r in recs[:6]:
    pos = ogo.GoPosition(oenv)
    for m in r.moves:
        pos = ogo.play_move(pos, ogo.from_flat(int(m), oenv))      # raises on an illegal move
    if not r.resigned:
        assert ogo.result(pos) == r.result, (r.game_id, ogo.score(pos), r.final_score)
    checked += 1
print(json.dumps({"config": "19x19 full games: 64 games, 32 readouts, T=2", "seconds": dt, "error": int(pr.error), "finished": len(recs),
                  "moves": int(pr.moves_played), "mean_length": float(np.mean([r.n_moves for r in recs])), "max_length": int(max(r.n_moves for r in recs)),
                  "resigned": int(sum(r.resigned for r in recs)), "replayed_ok": checked, "arena": eng.info()}), flush=True)
eng.close()
