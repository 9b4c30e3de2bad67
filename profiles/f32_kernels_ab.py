"""Old vs new fp32 SIMT kernels (forward / data-gradient convolution, weight gradient): outputs must be bit-identical; timing.
python profiles/f32_kernels_ab.py <old lib> -- compares the product library with another build."""
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
old = os.path.join(ROOT, sys.argv[1])
for N, T in ((9, 6), (19, 2), (9, 0)):
    out = {}
    for name, lib in (("new", None), ("old", old)):
        env = agz.GoEnv(N, lib_path=lib)
        nn = agz.NeuralNet(env, tower_height=T, seed=0)
        eng = agz.Engine(N, lib_path=lib, n_games=8, readouts=8, tower_height=T, evaluator=agz.EVAL_NN_TC)
        nn.push(eng)
        rs = np.random.RandomState(0)
        B, A = 32, N * N + 1
        bh = rs.randint(-1, 2, size=(B, 8, N * N)).astype(np.int8)
        tp = rs.choice([-1, 1], size=B).astype(np.int8)
        pis = rs.dirichlet(np.full(A, 0.3), size=B).astype(np.float32)
        zs = rs.choice([-1, 1], size=B).astype(np.int8)
        pi, v = eng.net_forward(agz.EVAL_NN_F32, bh[:20], tp[:20])
        losses = [eng.train_step(bh, tp, pis, zs) for _ in range(2)]
        grads = [eng.train_read_grads(k) for k in range(3)]
        t0 = time.perf_counter()
        K = 5 if N < 19 or T < 19 else 3
        for _ in range(K):
            losses.append(eng.train_step(bh, tp, pis, zs))
        dt = (time.perf_counter() - t0) / K
        params = [eng.net_get_params(k) for k in range(3)]
        out[name] = (pi, v, losses, grads, params, dt)
        eng.close()
    n, o = out["new"], out["old"]
    same = (np.array_equal(n[0], o[0]) and np.array_equal(n[1], o[1]) and n[2] == o[2] and all(np.array_equal(a, b) for a, b in zip(n[3], o[3]))
            and all(np.array_equal(a, b) for a, b in zip(n[4], o[4])))
    d = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))
    print(json.dumps({"config": "%dx%d T=%d B=32" % (N, N, T), "bit_identical": bool(same), "pi_diff": d(n[0], o[0]), "v_diff": d(n[1], o[1]),
                      "loss_diff": d(n[2], o[2]), "grad_diff": [d(a, b) for a, b in zip(n[3], o[3])], "grad_max": [float(np.max(np.abs(b))) for b in o[3]], "ms_per_step_new": 1e3 * n[5], "ms_per_step_old": 1e3 * o[5]}), flush=True)
