"""Times agz_train_step (one 32-position minibatch, fp32) on the two network sizes of BASELINE.json."""
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
for N, T in ((9, 6), (19, 19)):
    env = agz.GoEnv(N)
    nn = agz.NeuralNet(env, tower_height=T, seed=0)
    eng = agz.Engine(N, n_games=8, readouts=8, tower_height=T, evaluator=agz.EVAL_NN_TC)
    nn.push(eng)
    rs = np.random.RandomState(0)
    B, A = 32, N * N + 1
    bh = rs.randint(-1, 2, size=(B, 8, N * N)).astype(np.int8)
    tp = rs.choice([-1, 1], size=B).astype(np.int8)
    pis = rs.dirichlet(np.full(A, 0.3), size=B).astype(np.float32)
    zs = rs.choice([-1, 1], size=B).astype(np.int8)
    losses = [eng.train_step(bh, tp, pis, zs) for _ in range(2)]
    t0 = time.perf_counter()
    K = 5
    for _ in range(K):
        losses.append(eng.train_step(bh, tp, pis, zs))
    dt = (time.perf_counter() - t0) / K
    flops, _ = eng.net_flops()
    print(json.dumps({"config": "train step %dx%d T=%d B=%d fp32" % (N, N, T, B), "ms_per_step_incl_host_roundtrip": 1e3 * dt,
                      "approx_TFLOPs": 3 * flops * B / dt / 1e12, "losses": [round(float(x), 5) for x in losses]}), flush=True)
    eng.close()
