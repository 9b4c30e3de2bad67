"""Gomoku self-play throughput on the tensor-core network: 15x15 (the reference's GomokuEnv default), 1024 concurrent games, 400
readouts per move, tower_height 6 -- the C2 shape on the second game.  python profiles/gomoku_probe.py [board n_in_row steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

agz = pkg.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 15
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
env = agz.GomokuEnv(N, K)
nn = agz.NeuralNet(env, tower_height=6, seed=0)
eng = agz.Engine(N, n_games=1024, readouts=400, tower_height=6, seed=0, evaluator=agz.EVAL_NN_TC, game=agz.GAME_GOMOKU, n_in_row=K,
                 options={"selfplay.stagger_rounds": 30 * 50})
nn.push(eng)
eng.selfplay_start(-1)
for _ in range(34):                       # staggered start + burn-in: slots spread over the plies of a game
    eng.selfplay_step(50)
    eng.replay_gather()
    eng.selfplay_harvest_discard()
pr0 = eng.selfplay_step(50)
ms, pr = 0.0, pr0
for _ in range(steps):
    pr = eng.selfplay_step(50)
    ms += pr.step_ms
    eng.replay_gather()
    eng.selfplay_harvest_discard()
flops, _ = eng.net_flops()
print(json.dumps({"config": "Gomoku %dx%d (%d in a row), 1024 games, 400 readouts, T=6" % (N, N, K), "moves_per_s": (pr.moves_played - pr0.moves_played) / (ms * 1e-3),
                  "ms_per_round": ms / (steps * 50), "games_finished": int(pr.games_finished), "network_tflops": flops * 8192 * 50 * steps / (ms * 1e-3) / 1e12,
                  "device_error": int(pr.error)}), flush=True)
eng.close()
