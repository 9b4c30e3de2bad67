import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import pkg
agz = pkg.load()
e = agz.Engine(9, n_games=8192, readouts=1600, tower_height=1, seed=0, evaluator=agz.EVAL_DUMMY, nodes_per_game=3600)
e.selfplay_start(-1)
pr = e.selfplay_step(230)
print(pr.readouts, pr.error)
