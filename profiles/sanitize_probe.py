"""Small run of every tree-side kernel for compute-sanitizer (memcheck): self-play with the DummyNet evaluator on 9x9 and 19x19,
arena compaction, replay packing, the match entry points.  compute-sanitizer --tool memcheck python profiles/sanitize_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import pkg  # noqa: E402

agz = pkg.load()
for N, ro, cap in ((9, 24, 0), (19, 16, 0), (9, 16, 60)):
    eng = agz.Engine(N, n_games=4, readouts=ro, seed=1, nodes_per_game=cap)
    eng.set_dummy_evaluator(None, 0.0)
    eng.selfplay_start(4)
    recs, n = [], 0
    for _ in range(4000):
        pr = eng.selfplay_step(16)
        n = eng.replay_gather()              # packs the finished games before they are harvested
        recs += eng.selfplay_harvest(8)
        if pr.games_finished == 4 or pr.error:
            break
    assert pr.error == 0 and n == sum(r.n_moves for r in recs)
    eng.replay_sample_hist(4, seed=1)
    print(N, ro, cap, "games", len(recs), "moves", sum(r.n_moves for r in recs), "tuples", n, flush=True)
    eng.close()
# Gomoku (the second game of the Position interface): whole games, replay packing
eng = agz.Engine(9, n_games=4, readouts=16, seed=2, game=agz.GAME_GOMOKU, n_in_row=5)
eng.set_dummy_evaluator(None, 0.0)
eng.selfplay_start(4)
for _ in range(4000):
    pr = eng.selfplay_step(16)
    n = eng.replay_gather()
    eng.selfplay_harvest(8)
    if pr.games_finished == 4 or pr.error:
        break
assert pr.error == 0 and pr.games_finished == 4
print("gomoku ok", n, flush=True)
eng.close()
eng = agz.Engine(9, n_games=3, readouts=16, tau_threshold=-1, inject_noise=0)
eng.set_dummy_evaluator(None, 0.0)
eng.match_start()
alive = np.ones(3, bool)
for _ in range(12):
    mv, res, _ = eng.match_search(alive)
    eng.match_play(np.where(alive & ~res, mv, -1).astype(np.int32))
print("match ok", flush=True)
eng.close()

if len(sys.argv) > 1 and sys.argv[1] == "nn":   # the network kernels too (tcgen05 / TMA): small tower, 16 positions
    env = agz.GoEnv(9)
    nn = agz.NeuralNet(env, tower_height=1, seed=0)
    eng = agz.Engine(9, n_games=2, readouts=16, tower_height=1, evaluator=agz.EVAL_NN_TC)
    nn.push(eng)
    eng.selfplay_start(2)
    for _ in range(20):
        pr = eng.selfplay_step(4)
    rs = np.random.RandomState(0)
    bh = rs.randint(-1, 2, size=(8, 8, 81)).astype(np.int8)
    loss = eng.train_step(bh, np.ones(8, np.int8), rs.dirichlet(np.ones(82), size=8).astype(np.float32), np.ones(8, np.int8))
    for _ in range(200):                       # finish the games, then the device-resident replay -> train path and split precision
        pr = eng.selfplay_step(8)
        if pr.games_finished >= 2:
            break
    eng.replay_gather()
    loss2 = eng.train_step_from_replay(8, seed=1)
    eng.set_option("conv.precision", 2)
    eng.net_forward_debug(agz.EVAL_NN_TC, bh, np.ones(8, np.int8), want_trunk=True)
    print("nn ok", pr.moves_played, pr.error, loss, loss2, flush=True)
    eng.close()
