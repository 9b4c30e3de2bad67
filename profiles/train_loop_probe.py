"""The whole RL loop at moderate scale: train(env; ...) = concurrent self-play with the current network -> replay gather -> one
optimisation step per finished game -> the updated parameters drive the next searches.  Prints the loss curve and timings."""
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
games = int(sys.argv[1]) if len(sys.argv) > 1 else 768
t0 = time.perf_counter()
nn = agz.train(env, num_games=games, batch_size=32, readouts=64, tower_height=2, start_training_after=4000, concurrent=256, seed=1,
               verbose=False)
dt = time.perf_counter() - t0
L = np.array(nn.train_losses)
k = max(1, len(L) // 5)
print(json.dumps({"config": "train loop: 9x9, T=2, 64 readouts, 256 concurrent games, %d games, batch 32" % games, "seconds": dt,
                  "train_steps": int(len(L)), "loss_first_fifth": float(L[:k].mean()) if len(L) else None,
                  "loss_last_fifth": float(L[-k:].mean()) if len(L) else None, "finite": bool(np.all(np.isfinite(L)))}), flush=True)
