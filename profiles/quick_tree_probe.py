"""Ad-hoc GPU probe used during development: timing of the tree kernels with the dummy evaluator."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pkg
agz = pkg.load()

def run(N, G, R, rounds, ev=0, T=1):
    e = agz.Engine(N, n_games=G, readouts=R, tower_height=T, nodes_per_game=0)
    e.selfplay_start(-1)
    e.selfplay_step(10)
    e.set_timing(True)
    t = time.time()
    pr0 = e.selfplay_step(1)
    t = time.time()
    pr = e.selfplay_step(rounds)
    dt = time.time() - t
    ms, ln = e.phase_times()
    print(f"N={N} G={G} R={R}: {rounds} rounds in {dt*1e3:.1f} ms; moves={pr.moves_played-pr0.moves_played} readouts={pr.readouts-pr0.readouts} "
          f"pathnodes={pr.path_nodes-pr0.path_nodes} phases(ms)={[round(x,2) for x in ms]} err={pr.error} live={pr.games_live}", flush=True)
    e.close()

if __name__ == "__main__":
    run(9, 1024, 400, 100)
    run(9, 8192, 1600, 100)
    run(19, 512, 800, 50)
