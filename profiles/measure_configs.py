"""Measures the other BASELINE.json configs on one B200 (C3 per-GPU share, C4 network-only, C5 MCTS-only) and prints one
JSON line each.  Same engine / kernels as bench.py; used to fill the results table in DESIGN.md."""
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
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}


def run(name, N, games, readouts, tower, evaluator, warm_rounds, rounds, nodes_per_game=0):
    env = agz.GoEnv(N)
    eng = agz.Engine(N, n_games=games, readouts=readouts, tower_height=tower, seed=0, evaluator=evaluator, nodes_per_game=nodes_per_game)
    if evaluator != agz.EVAL_DUMMY:
        agz.NeuralNet(env, tower_height=tower, seed=0).push(eng)
    eng.selfplay_start(-1)
    pr0 = eng.selfplay_step(warm_rounds)
    # plain run first (no per-kernel events): with the DummyNet evaluator all rounds of a call are one launch per game
    t0 = time.perf_counter()
    prp = eng.selfplay_step(rounds)
    plain_wall = time.perf_counter() - t0
    plain = {"ms_per_round_plain": 1e3 * plain_wall / rounds, "moves_per_s_plain": (prp.moves_played - pr0.moves_played) / plain_wall,
             "readouts_per_s_plain": (prp.readouts - pr0.readouts) / plain_wall}
    eng.set_timing(True)
    eng.phase_times(reset=True)
    pr1 = eng.selfplay_step(1)
    t0 = time.perf_counter()
    pr = eng.selfplay_step(rounds)
    wall = time.perf_counter() - t0
    kms, kln = eng.phase_times(reset=True)
    A = N * N + 1
    readouts_done = pr.readouts - pr1.readouts
    pathnodes = pr.path_nodes - pr1.path_nodes
    d = pathnodes / max(1, readouts_done)
    tree_bytes = readouts_done * ((d - 1) * 13 * A + 16 * A + 8 * N * N + 16 * d)      # SURVEY 8d formula
    tree_ms = kms[0] + kms[5]
    out = {"config": name, "board": N, "games": games, "readouts": readouts, "tower_height": tower, "rounds": rounds,
           "ms_per_round": 1e3 * wall / rounds, "step_ms_device": pr.step_ms / rounds,
           "moves_per_s": (pr.moves_played - pr1.moves_played) / wall, "readouts_per_s": readouts_done / wall, "mean_path_len": d,
           "kernel_ms_per_round": {n: kms[i] / max(1, kln[0]) for i, n in enumerate(agz.binding.KERNEL_NAMES)},
           "tree_GBps_algorithmic": tree_bytes / (tree_ms * 1e-3) / 1e9 if tree_ms > 0 else None,
           "tree_frac_of_hbm": tree_bytes / (tree_ms * 1e-3) / 1e9 / PEAKS["hbm_gbs"] if tree_ms > 0 else None, "error": pr.error}
    out.update(plain)
    out["tree_frac_of_hbm_plain"] = (tree_bytes / readouts_done) * plain["readouts_per_s_plain"] / 1e9 / PEAKS["hbm_gbs"] if readouts_done else None
    if evaluator != agz.EVAL_DUMMY:
        fpos, fconv = eng.net_flops()
        rows = games * 8
        net_ms = (kms[2] + kms[3] + kms[4]) / max(1, kln[0])
        conv_ms = kms[3] / max(1, kln[3])
        out.update({"positions_per_s_network": rows / (net_ms * 1e-3), "network_ms_per_batch": net_ms, "network_TFLOPs": fpos * rows / (net_ms * 1e-3) / 1e12,
                    "tower_conv_TFLOPs": fconv * rows / (conv_ms * 1e-3) / 1e12,
                    "tower_conv_frac_of_sustained_peak": fconv * rows / (conv_ms * 1e-3) / 1e12 / PEAKS["bf16_tflops_sustained"]})
    print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["C5", "C4", "C3"]
    if "C5" in which:   # MCTS-only: 9x9, uniform prior / value 0 (DummyNet), 8192 trees x 1600 readouts
        run("C5 MCTS-only 9x9 8192 trees x 1600 readouts", 9, 8192, 1600, 1, agz.EVAL_DUMMY, 210, 200, nodes_per_game=3600)
    if "C4" in which:   # network-only: 19x19, batch 8192, tower_height 19 (1024 games x 8 leaves), positions from live self-play
        run("C4 NN-only 19x19 batch 8192 T=19", 19, 1024, 800, 19, agz.EVAL_NN_TC, 3, 6)
    if "C3" in which:   # one GPU's share of C3: 19x19, 512 games, 800 readouts, T=19
        run("C3 per-GPU share: 19x19 512 games 800 readouts T=19", 19, 512, 800, 19, agz.EVAL_NN_TC, 3, 10)
