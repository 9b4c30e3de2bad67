"""Kernel timeline of a few C2 rounds from the engine's own %globaltimer trace (no nsys in the image).
   python profiles/trace_c2.py [rounds] [pipeline 0|1] [max_pairs]  -> one line per kernel launch, time relative to the first."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import pkg  # noqa: E402

agz = pkg.load()
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
env = agz.GoEnv(9)
nn = agz.NeuralNet(env, tower_height=6, seed=0)
pipeline = int(sys.argv[2]) if len(sys.argv) > 2 else 0
max_pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 0
eng = agz.Engine(9, n_games=1024, readouts=400, tower_height=6, seed=0, evaluator=agz.EVAL_NN_TC,
                 options={"trace.records": 20000, "schedule.pipeline": pipeline, "conv.max_pairs": max_pairs})
nn.push(eng)
eng.selfplay_start(-1)
eng.selfplay_step(100)
eng.trace_read()
eng.selfplay_step(rounds)
tr = eng.trace_read()
names = {1: "select", 2: "incorp", 3: "feats", 4: "stem", 5: "conv", 6: "conv+res", 7: "heads", 9: "other"}
# merge the first/last CTA records of a launch: same tag + grid, starts within the launch
tr = tr[np.argsort(tr[:, 3])]
t00 = tr[0, 3]
launches = []
for tag, blk, grid, t0, t1, sm in tr:
    for L in reversed(launches[-40:]):
        if L["tag"] == tag and L["grid"] == grid and blk not in L["blks"] and len(L["blks"]) < 2 and abs(t0 - L["t0"]) < 2_000_000:
            L["blks"].append(blk); L["t0"] = min(L["t0"], t0); L["t1"] = max(L["t1"], t1)
            break
    else:
        launches.append({"tag": tag, "grid": grid, "blks": [blk], "t0": t0, "t1": t1})
print("pipeline=%d pairs=%s" % (pipeline, max_pairs or "auto"))
for L in launches:
    print("%-9s grid %5d  start %9.1f us  end %9.1f us  dur %7.1f us" % (names.get(L["tag"], "?"), L["grid"], (L["t0"] - t00) / 1e3, (L["t1"] - t00) / 1e3, (L["t1"] - L["t0"]) / 1e3))
eng.close()
