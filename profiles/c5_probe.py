"""C5 probe (MCTS-only 9x9, 8192 trees x 1600 readouts, DummyNet): prints the roofline_tree leg of bench.py; with `ncu` as argv[1]
runs few short launches instead (for `ncu -k regex:k_warps`): 2 warm-up launches of 210 rounds, then launches of 8 rounds."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import pkg  # noqa: E402

agz = pkg.load()
if len(sys.argv) > 1 and sys.argv[1] == "ncu":
    eng = agz.Engine(9, n_games=8192, readouts=1600, tower_height=1, seed=0, evaluator=agz.EVAL_DUMMY, nodes_per_game=3600)
    eng.selfplay_start(-1)
    eng.selfplay_step(210)
    eng.selfplay_step(210)
    for _ in range(4):
        pr0 = eng.selfplay_step(8)
    pr1 = eng.selfplay_step(8)
    print(json.dumps({"readouts_per_8_round_launch": pr1.readouts - pr0.readouts, "path_nodes": pr1.path_nodes - pr0.path_nodes}))
    eng.close()
else:
    print(json.dumps(bench.leg_c5(agz, 0, bench.peaks()[0])))
