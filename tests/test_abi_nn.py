"""Network + feature parity through the C ABI (GPU): agz_features / agz_net_forward against the oracle
(oracle/net.py, torch-CPU fp32) on the shipped agz weights (golden fixture) and random-init towers, and
NN-driven self-play checked move by move against the oracle tree fed with the same network outputs."""
import os

import numpy as np
import pytest

from backends import agz, lib_for
from oracle import go as ogo
from oracle import net as onet
from oracle import selfplay as osp
from oracle import features as ofeat

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agz_shipped_9x9.npz")
TOL_F32 = 2e-5      # fp32 SIMT path vs torch fp32
TOL_TC = 1e-3       # north_star: policy/value within 1e-3 of the fp32 reference


def test_golden_fixture_matches_oracle_feature_code():
    """CPU: the committed fixture is self-consistent (features <-> boards_hist) -- guards the generator."""
    g = np.load(GOLD)
    bh, tp, feats = g["boards_hist"], g["to_play"], g["feats"]
    for b in range(bh.shape[0]):
        for k in range(8):
            board = bh[b, k].reshape(9, 9, order="F")
            assert (feats[b, :, :, 2 * k] == (board == tp[b])).all()
            assert (feats[b, :, :, 2 * k + 1] == (board == -tp[b])).all()
        assert (feats[b, :, :, 16] == tp[b]).all()
    assert np.allclose(g["pi"].sum(axis=1), 1, atol=1e-5)


def hist_stack(pos):
    f = ofeat.stone_features(pos)
    return np.stack([((f[:, :, 2 * k] - f[:, :, 2 * k + 1]) * pos.to_play).astype(np.int8).flatten(order="F") for k in range(8)])


def push_oracle_net(eng, nn):
    flat = lambda lst: np.concatenate([np.asarray(a, np.float32).flatten(order="F") for a in lst])
    eng.net_set_params(0, flat(nn.base_params()))
    eng.net_set_params(1, flat(nn.value_params()))
    eng.net_set_params(2, flat(nn.policy_params()))
    bns = nn.base_bns()
    eng.net_set_bn_stats(0, np.concatenate([b.mu for b in bns]), np.concatenate([b.sigma for b in bns]), bns[0].mode)
    eng.net_set_bn_stats(1, nn.v_bn.mu, nn.v_bn.sigma, nn.v_bn.mode)
    eng.net_set_bn_stats(2, nn.p_bn.mu, nn.p_bn.sigma, nn.p_bn.mode)


def random_positions(N, count, seed, max_plies=70):
    env = ogo.GoEnv(N)
    rs = np.random.RandomState(seed)
    out = []
    while len(out) < count:
        pos = ogo.GoPosition(env)
        for t in range(rs.randint(0, max_plies)):
            legal = np.flatnonzero(ogo.all_legal_moves(pos)[:-1])
            mv = None if len(legal) == 0 or rs.rand() < 0.05 else ogo.from_flat(int(rs.choice(legal)), env)
            pos = ogo.play_move(pos, mv)
            if pos.done:
                break
        if not pos.done:
            out.append(pos)
    return out


EVALS = [pytest.param(agz.EVAL_NN_F32, TOL_F32, id="f32"), pytest.param(agz.EVAL_NN_TC, TOL_TC, id="tcgen05")]


@pytest.mark.gpu
def test_features_match_golden_and_oracle():
    g = np.load(GOLD)
    eng = agz.Engine(9, lib_path=lib_for("cuda"), n_games=4, tower_height=0)
    out = eng.features(g["boards_hist"], g["to_play"])                      # [B][17][N2]
    ref = np.transpose(g["feats"], (0, 3, 2, 1)).reshape(-1, 17, 81)        # [b][c][j][i] -> p = 9*j + i
    assert np.array_equal(out, ref.astype(np.float32))
    poss = random_positions(19, 6, 3, max_plies=200)
    eng19 = agz.Engine(19, lib_path=lib_for("cuda"), n_games=2, tower_height=0)
    out = eng19.features(np.stack([hist_stack(p) for p in poss]), np.array([p.to_play for p in poss], np.int8))
    ref = np.stack([np.transpose(ofeat.get_feats(p), (2, 1, 0)).reshape(17, 361) for p in poss])
    assert np.array_equal(out, ref.astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("evaluator,tol", EVALS)
def test_shipped_agz_net_matches_golden(evaluator, tol):
    g = np.load(GOLD)
    eng = agz.Engine(9, lib_path=lib_for("cuda"), n_games=4, tower_height=0)
    eng.net_set_params(0, g["base"]); eng.net_set_params(1, g["value"]); eng.net_set_params(2, g["policy"])
    eng.net_set_bn_stats(0, g["bn_mu_base"], g["bn_sigma_base"], agz.BN_STD)
    eng.net_set_bn_stats(1, g["bn_mu_value"], g["bn_sigma_value"], agz.BN_STD)
    eng.net_set_bn_stats(2, g["bn_mu_policy"], g["bn_sigma_policy"], agz.BN_STD)
    pi, v = eng.net_forward(evaluator, g["boards_hist"], g["to_play"])
    assert np.abs(pi - g["pi"]).max() <= tol, np.abs(pi - g["pi"]).max()
    assert np.abs(v - g["v"]).max() <= tol, np.abs(v - g["v"]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("evaluator,tol", EVALS)
@pytest.mark.parametrize("N,T,B", [(9, 1, 40), (9, 6, 70), (19, 2, 9), (13, 1, 5), (5, 2, 3)])
def test_random_towers_match_oracle(evaluator, tol, N, T, B):
    nn = onet.NeuralNet(N, T, seed=10 + T)
    nn.randomize_bn(seed=T)
    poss = random_positions(N, B, 100 + N + T, max_plies=70 if N == 9 else 250)
    pi_ref, v_ref = nn(poss)
    eng = agz.Engine(N, lib_path=lib_for("cuda"), n_games=8, tower_height=T)     # 64 rows per internal batch -> B=70 spans 2
    push_oracle_net(eng, nn)
    pi, v = eng.net_forward(evaluator, np.stack([hist_stack(p) for p in poss]), np.array([p.to_play for p in poss], np.int8))
    assert np.abs(pi - pi_ref.T).max() <= tol, np.abs(pi - pi_ref.T).max()
    assert np.abs(v - v_ref).max() <= tol, np.abs(v - v_ref).max()
    # batch invariance: a position's outputs must not depend on its batch neighbours (hard part 7)
    idx = np.arange(B)[::-1].copy()
    pi2, v2 = eng.net_forward(evaluator, np.stack([hist_stack(poss[i]) for i in idx]), np.array([poss[i].to_play for i in idx], np.int8))
    assert np.array_equal(pi2, pi[idx]) and np.array_equal(v2, v[idx])


class EngineBackedNet:
    """Oracle-side network whose numbers come from the engine's own forward pass, so the oracle tree and the
    engine tree see bit-identical (pi, v): isolates tree/feature/batching parity from NN rounding."""

    def __init__(self, eng, evaluator, tower_height):
        self.eng, self.evaluator, self.tower_height = eng, evaluator, tower_height

    def __call__(self, positions):
        pi, v = self.eng.net_forward(self.evaluator, np.stack([hist_stack(p) for p in positions]),
                                     np.array([p.to_play for p in positions], np.int8))
        return pi.T.copy(), v


@pytest.mark.gpu
@pytest.mark.parametrize("evaluator", [pytest.param(agz.EVAL_NN_F32, id="f32"), pytest.param(agz.EVAL_NN_TC, id="tcgen05")])
def test_nn_selfplay_matches_oracle_tree(evaluator):
    N, T, R = 9, 1, 24
    nn = onet.NeuralNet(N, T, seed=5)
    nn.randomize_bn(seed=2)
    helper = agz.Engine(N, lib_path=lib_for("cuda"), n_games=2, tower_height=T)
    push_oracle_net(helper, nn)
    # 4 slots: the tcgen05 evaluator then runs the two-stream half-batch pipeline, the fp32 one the sequential schedule
    eng = agz.Engine(N, lib_path=lib_for("cuda"), n_games=4, readouts=R, seed=9, tower_height=T)
    push_oracle_net(eng, nn)
    eng.set_evaluator(evaluator)
    recs = eng.selfplay_run(6)
    oenv = ogo.GoEnv(N)
    onn = EngineBackedNet(helper, evaluator, T)
    for gid, r in enumerate(recs):
        op = osp.selfplay(oenv, onn, R, seed=9, game_id=gid)
        om = [ogo.to_flat(m.move, oenv) for m in op.root.position.recent]
        assert list(r.moves) == om, gid
        assert np.array_equal(np.array(op.searches_N), r.visits)
        assert r.result == op.result and r.result_string == op.result_string


# ------------------------------------------------------------------------------------------------ depth / sharpness (round 2)
import nn_parity  # noqa: E402


def _trained_like(N, T, seed):
    """A network with the statistics of a trained one: biases, a policy head with tens of logit units of spread (pi_max up to ~1),
    BatchNorm running statistics fitted to the activations of a calibration batch (+-10 %)."""
    nn = onet.NeuralNet(N, T, seed=seed)
    nn.sharpen(seed=seed + 1)
    nn.calibrate_bn(nn.feats_to_torch(random_positions(N, 16, 99, max_plies=100)), seed=seed + 2)
    return nn


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,B", [(9, 19, 8), (19, 19, 4)])
def test_north_star_depth_random_init_within_1e3(N, T, B):
    """tower_height 19 (39 fp16 conv layers), the network BASELINE.json benchmarks: Flux-default random init.  pi and v within 1e-3 of
    the fp32 oracle, and -- a far sharper discriminator on a near-uniform policy -- the centred logits and the value before tanh."""
    nn = onet.NeuralNet(N, T, seed=0)
    poss = random_positions(N, B, 11, max_plies=70 if N == 9 else 250)
    ref = nn.forward_debug(nn.feats_to_torch(poss))
    eng = agz.Engine(N, lib_path=lib_for("cuda"), n_games=1, tower_height=T)
    push_oracle_net(eng, nn)
    bh, tp = nn_parity.engine_inputs(poss)
    rep = nn_parity.report(eng.net_forward_debug(agz.EVAL_NN_TC, bh, tp), ref)
    assert rep["pi_abs"] <= TOL_TC and rep["v_abs"] <= TOL_TC, rep
    assert rep["logit_rel"] <= 1e-2 and rep["vpre_rel"] <= 1e-2, rep
    rep32 = nn_parity.report(eng.net_forward_debug(agz.EVAL_NN_F32, bh, tp), ref)
    assert rep32["pi_abs"] <= TOL_F32 and rep32["v_abs"] <= 5e-5 and rep32["logit_rel"] <= 5e-5, rep32
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,B", [(9, 6, 8), (9, 19, 8), (19, 19, 4)])
def test_trained_like_networks_need_split_precision_for_1e3(N, T, B):
    """Sharp policies expose operand rounding: with fp16 operands (11 significant bits) the error grows ~sqrt(depth) and passes 1e-3
    (measured 3e-3 ... 8e-3, profiles/r02_nn_error_vs_depth.jsonl) -- kept under a regression bound here -- while split precision
    (option conv.precision = 2: hi + lo fp16 pairs) meets the north-star 1e-3 on pi and v at every depth and board size."""
    nn = _trained_like(N, T, 21)
    poss = random_positions(N, B, 7, max_plies=70 if N == 9 else 250)
    ref = nn.forward_debug(nn.feats_to_torch(poss))
    eng = agz.Engine(N, lib_path=lib_for("cuda"), n_games=1, tower_height=T)
    push_oracle_net(eng, nn)
    bh, tp = nn_parity.engine_inputs(poss)
    fast = nn_parity.report(eng.net_forward_debug(agz.EVAL_NN_TC, bh, tp), ref)
    assert fast["pi_max"] > 0.5 and fast["logit_span"] > 10, fast            # the test network really is sharp
    assert fast["pi_abs"] <= 2.5e-2 and fast["v_abs"] <= 2.5e-2 and fast["logit_rel"] <= 1e-2, fast
    eng.set_option("conv.precision", 2)
    out = eng.net_forward_debug(agz.EVAL_NN_TC, bh, tp)
    fine = nn_parity.report(out, ref)
    assert fine["pi_abs"] <= TOL_TC and fine["v_abs"] <= TOL_TC and fine["logit_rel"] <= 1e-3 and fine["vpre_rel"] <= 2e-3, fine
    pi, v = eng.net_forward(agz.EVAL_NN_TC, bh, tp)                            # the plain entry point runs the same arithmetic
    assert np.array_equal(pi, out["pi"]) and np.array_equal(v, out["v"])
    eng.set_option("conv.precision", 1)
    again = nn_parity.report(eng.net_forward_debug(agz.EVAL_NN_TC, bh, tp), ref)
    assert again["pi_abs"] == fast["pi_abs"]                                   # switching back restores the fp16 path bit for bit
    eng.close()


@pytest.mark.gpu
def test_injected_tap_error_is_caught_and_bisected():
    """Sensitivity of the parity checks: the engine is given a network whose block-4 second convolution has two taps exchanged (the
    kind of mistake a wrong kernel flip or im2col offset makes) while the oracle keeps the right one.  The end-to-end metrics must
    fail their bounds, and the per-block bisect helper must name block 4 (trunk after 5 blocks) as the first one off."""
    N, T, B, bad = 9, 8, 8, 4
    nn = _trained_like(N, T, 31)
    poss = random_positions(N, B, 5)
    ref = nn.forward_debug(nn.feats_to_torch(poss))
    broken = _trained_like(N, T, 31)
    W = broken.blocks[bad]["W2"].copy()
    W[0, 0], W[2, 2] = broken.blocks[bad]["W2"][2, 2].copy(), broken.blocks[bad]["W2"][0, 0].copy()
    broken.blocks[bad]["W2"] = W
    eng = agz.Engine(N, lib_path=lib_for("cuda"), n_games=1, tower_height=T)
    bh, tp = nn_parity.engine_inputs(poss)
    for precision, bound in ((2, lambda nb: 6e-4), (1, lambda nb: 4e-3)):
        eng.set_option("conv.precision", precision)
        push_oracle_net(eng, nn)
        ok = nn_parity.report(eng.net_forward_debug(agz.EVAL_NN_TC, bh, tp), ref)
        assert nn_parity.first_bad_block(eng, agz.EVAL_NN_TC, poss, ref, T, bound) is None, precision
        push_oracle_net(eng, broken)
        rep = nn_parity.report(eng.net_forward_debug(agz.EVAL_NN_TC, bh, tp), ref)
        assert rep["pi_abs"] > 10 * max(ok["pi_abs"], 1e-4) and rep["logit_rel"] > 10 * ok["logit_rel"], (precision, ok, rep)
        assert nn_parity.first_bad_block(eng, agz.EVAL_NN_TC, poss, ref, T, bound) == bad + 1, precision
    eng.close()


@pytest.mark.gpu
def test_reference_julia_outputs():
    """Pins the network against the reference itself -- once somebody has run tests/golden/gen_golden.jl (needs Julia + Flux +
    BSON.jl, absent from this image) and committed its output.  Until then: skipped, and NN parity stays "unpinned"."""
    import json
    path = os.path.join(os.path.dirname(GOLD), "agz_shipped_9x9_julia.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/agz_shipped_9x9_julia.json has not been generated (no Julia in this image)")
    ref = json.load(open(path))
    g = np.load(GOLD)
    pi_j, v_j = np.array(ref["pi"], np.float32), np.array(ref["v"], np.float32)
    assert np.abs(g["pi"] - pi_j).max() <= 1e-5 and np.abs(g["v"] - v_j).max() <= 1e-5, "the fp32 oracle differs from the reference's Flux network"
    eng = agz.Engine(9, lib_path=lib_for("cuda"), n_games=4, tower_height=0)
    eng.net_set_params(0, g["base"]); eng.net_set_params(1, g["value"]); eng.net_set_params(2, g["policy"])
    for k, name in enumerate(("base", "value", "policy")):
        eng.net_set_bn_stats(k, g["bn_mu_" + name], g["bn_sigma_" + name], agz.BN_STD)
    pi, v = eng.net_forward(agz.EVAL_NN_TC, g["boards_hist"], g["to_play"])
    assert np.abs(pi - pi_j).max() <= TOL_TC and np.abs(v - v_j).max() <= TOL_TC
