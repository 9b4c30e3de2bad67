"""A/B of library builds on the C5 leg (MCTS-only 9x9, 8192 trees x 1600 readouts): python profiles/variant_probe.py [lib.so ...]
Each library (default: the product build) is timed twice, interleaved, on the same GPU."""
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
libs = [None] + [os.path.join(ROOT, p) for p in sys.argv[1:]]
for rep in range(2):
    for lib in libs:
        r = bench.leg_c5(agz, 0, hbm, lib_path=lib)
        print(json.dumps({"lib": os.path.basename(lib) if lib else "libagz.so", "ms_per_round": r["ms_per_round"], "moves_per_s": r["moves_per_s"],
                          "frac": r["frac"], "error": r["error"]}), flush=True)
