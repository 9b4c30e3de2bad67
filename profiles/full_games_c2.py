"""C2 played to the end: 1024 concurrent 9x9 games, 400 readouts, tower_height 6, every slot refilled when its game ends.
Reports throughput over the whole run (early game, late game with terminal leaves / resignations / compaction, refills),
finished-game statistics and the device error flag.  python profiles/full_games_c2.py [steps]"""
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
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
env = agz.GoEnv(9)
nn = agz.NeuralNet(env, tower_height=6, seed=0)
eng = agz.Engine(9, n_games=1024, readouts=400, tower_height=6, seed=0, evaluator=agz.EVAL_NN_TC)
nn.push(eng)
eng.selfplay_start(-1)
pr = eng.selfplay_step(50)
m0, t0 = pr.moves_played, time.perf_counter()
lengths, results, resigned, per_step = [], [], 0, []
tuples = 0
for s in range(steps):
    m_prev, t_prev = pr.moves_played, time.perf_counter()
    pr = eng.selfplay_step(50)
    tuples = eng.replay_gather()
    recs = eng.selfplay_harvest(4096)
    per_step.append((pr.moves_played - m_prev) / (time.perf_counter() - t_prev))
    for r in recs:
        lengths.append(r.n_moves); results.append(r.result); resigned += int(r.resigned)
    if pr.error:
        break
wall = time.perf_counter() - t0
print(json.dumps({"config": "C2 full games: 9x9, 1024 slots, 400 readouts, T=6, %d move-steps" % steps, "moves_per_s_whole_run": (pr.moves_played - m0) / wall,
                  "moves_per_s_first_10_steps": float(np.mean(per_step[:10])), "moves_per_s_last_10_steps": float(np.mean(per_step[-10:])),
                  "min_step_moves_per_s": float(np.min(per_step)), "games_finished": len(lengths), "mean_game_length": float(np.mean(lengths)) if lengths else None,
                  "max_game_length": int(np.max(lengths)) if lengths else None, "resigned_frac": resigned / max(1, len(lengths)),
                  "black_win_frac": float(np.mean(np.array(results) == 1)) if results else None, "replay_tuples": int(tuples),
                  "positions_evaluated": int(pr.positions_evaluated), "readouts": int(pr.readouts), "device_error": int(pr.error)}), flush=True)
eng.close()
