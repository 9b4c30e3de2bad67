"""Error of the tcgen05 (fp16 operands / fp32 accumulate) network against the fp32 oracle as a function of depth, for Glorot-init and
"sharp" (trained-like: biases, BatchNorm statistics, policy gain) networks: per-block trunk deviation, centred logits, value before
tanh, pi, v.  One JSON line per network.   python profiles/nn_error_vs_depth.py [N=9] [T=19] [positions=8]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import nn_parity  # noqa: E402
import pkg  # noqa: E402
from oracle import net as onet  # noqa: E402
from test_abi_nn import push_oracle_net, random_positions  # noqa: E402

agz = pkg.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 9
T = int(sys.argv[2]) if len(sys.argv) > 2 else 19
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
poss = random_positions(N, B, 7, max_plies=70 if N == 9 else 250)
for kind in ("glorot", "sharp"):
    nn = onet.NeuralNet(N, T, seed=21)
    if kind == "sharp":      # trained-like: biases, policy gain, BatchNorm statistics fitted to the activations (+- 10 %)
        nn.sharpen(seed=6)
        nn.calibrate_bn(nn.feats_to_torch(random_positions(N, 16, 99, max_plies=100)), seed=5)
    else:
        nn.randomize_bn(seed=5)
    ref = nn.forward_debug(nn.feats_to_torch(poss))
    eng = agz.Engine(N, n_games=max(1, (B + 7) // 8), tower_height=T)
    push_oracle_net(eng, nn)
    bh, tp = nn_parity.engine_inputs(poss)
    out = {"network": kind, "N": N, "tower_height": T, "positions": B}
    for name, ev in (("tcgen05_f16", agz.EVAL_NN_TC), ("tcgen05_split_f16x2", agz.EVAL_NN_TC), ("simt_f32", agz.EVAL_NN_F32)):
        eng.set_option("conv.precision", 2 if "split" in name else 1)
        rep = nn_parity.report(eng.net_forward_debug(ev, bh, tp), ref)
        rep["trunk_rel_rms_by_blocks"] = nn_parity.trunk_errors(eng, ev, poss, ref, [0, 1, 2, 4, 6, 10, 14, T] if T >= 14 else list(range(T + 1)))
        out[name] = rep
    print(json.dumps(out), flush=True)
    eng.close()
