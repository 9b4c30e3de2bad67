"""A/B of the tree kernels on C5 (MCTS-only 9x9, 8192 trees x 1600 readouts): one warp per tree vs two trees per warp
(option tree.duo), optionally with library variants built for other occupancies (libagz_duo<N>.so = AGZ_DUO_CTAS N)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import pkg  # noqa: E402

agz = pkg.load()
hbm = bench.peaks()[0]
runs = [("solo", {"tree.duo": 0}, None), ("duo5", {"tree.duo": 1}, None)]
for n in (4, 6):
    p = os.path.join(ROOT, "alphago.jl_b200", "libagz_duo%d.so" % n)
    if os.path.exists(p):
        runs.append(("duo%d" % n, {"tree.duo": 1}, p))
runs += runs[:2]
for name, opt, lib in runs:
    r = bench.leg_c5(agz, 0, hbm, options=opt, lib_path=lib)
    print(json.dumps({"run": name, "ms_per_round": r["ms_per_round"], "moves_per_s": r["moves_per_s"], "frac": r["frac"], "mean_path_nodes": r["mean_path_nodes"],
                      "readouts_per_s": r["readouts_per_s"], "error": r["error"]}), flush=True)
