"""CPU oracle for the AlphaGo.jl self-play hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is a CPU restatement (Python + numpy,
torch-CPU fp32 for the network) of the reference algorithm
(tejank10/AlphaGo.jl @ 8a2651c) and exists only so that the CUDA engine can be
checked against it.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the
product package (`alphago.jl_b200/`) never does and has no CPU fallback.

Pinning status
--------------
* Go rules, MCTS, player and feature code: PINNED by a 1:1 port of the
  reference's own unit tests (`test/test_go.jl`, `test/test_mcts.jl`,
  `test/test_mcts_player.jl`, `test/test_features.jl`) in
  `tests/test_oracle_*.py`.
* Gomoku (`gomoku.py`, dispatched by `game.py` like the reference's `Position` interface): the reference has no test for
  this game; pinned only by hand-checked cases derived from src/game/gomoku/board.jl (`tests/test_oracle_gomoku.py`).
* Random draws (tie-breaks, Dirichlet noise, soft-pick): the reference draws
  from Julia's global RNG (MersenneTwister + Distributions.jl) which cannot be
  reproduced here (no Julia).  The oracle and the engine share ONE explicit
  counter-based spec instead (`oracle/rng.py`); draws happen at exactly the
  reference's five call sites.  "parity unpinned" for the random stream itself.
* pi = children_as_pi (src/mcts.jl:241-252): the reference computes child_N .^ 0.98 with libm `pow` and sums with Julia's
  pairwise `sum`; neither is bit-reproducible across platforms, so the oracle restates them in the engine's deterministic form
  (`rng.det_pow`: exp/log built from +,-,*,/ only; `rng.butterfly_sum32`: the 32-lane xor-butterfly order).  pi "bit-exact" is
  therefore engine-vs-this-spec at the last ulp; against libm it agrees to ~1e-15 relative (tests/test_oracle_mcts.py checks
  det_pow against math.pow).  The same holds for the Dirichlet sampler's log/exp.
* Replay sampling (src/train.jl:4-12, StatsBase.sample on the global RNG): unpinned by the reference; shared explicit spec in
  `oracle/replay.py` (keyed Feistel permutation of the ring's index range).
* Network arithmetic: lives in Flux 0.10.4 / NNlib 0.6.6 (not vendored in the
  reference, Manifest.toml:193-197,298-302) and no reference test constructs a
  NeuralNet => "parity unpinned" for NN outputs; anchored on the layer
  definitions (src/neural_net.jl:13-33, src/resnet.jl:2-32) and the shipped
  models/weights/agz_*.bson tensors.

Index convention: everything here is 0-based.  A board point is (i, j) = (row
from the top, column); the flat move is f = N*j + i (Julia: N*(j-1)+i, 1-based,
src/game/go/coords.jl:6-7), pass = N*N.
"""
