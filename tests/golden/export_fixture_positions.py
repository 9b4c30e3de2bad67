"""agz_shipped_9x9.npz -> agz_shipped_9x9_positions.json: the fixture positions in a form tests/golden/gen_golden.jl reads without
extra Julia packages (8 history boards as flat column-major lists, current first, and to_play)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
g = np.load(os.path.join(HERE, "agz_shipped_9x9.npz"))
out = {"boards_hist": g["boards_hist"].astype(int).tolist(), "to_play": g["to_play"].astype(int).tolist()}
json.dump(out, open(os.path.join(HERE, "agz_shipped_9x9_positions.json"), "w"))
print("wrote agz_shipped_9x9_positions.json:", len(out["to_play"]), "positions")
