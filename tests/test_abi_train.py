"""agz_train_step against the oracle restatement of _train / losses / Momentum (oracle/train.py, torch autograd fp32):
loss, gradients, updated parameters, running statistics, and the network the engine then plays with."""
import numpy as np
import pytest

from backends import agz, lib_for
from oracle import go as ogo
from oracle import net as onet
from oracle import train as otrain

pytestmark = pytest.mark.gpu


def _batch(N, B, seed):
    """Positions from random legal playouts (boards + 7-board history), visit-count-like policy targets, results."""
    rs = np.random.RandomState(seed)
    env = ogo.GoEnv(N)
    positions = []
    while len(positions) < B:
        pos = ogo.GoPosition(env)
        for _ in range(rs.randint(1, 3 * N)):
            legal = np.flatnonzero(ogo.all_legal_moves(pos)[:-1])
            if len(legal) == 0:
                break
            pos = ogo.play_move(pos, ogo.from_flat(int(rs.choice(legal)), env))
        positions.append(pos)
    A = N * N + 1
    pis = rs.dirichlet(np.full(A, 0.3), size=B).astype(np.float32)
    zs = rs.choice([-1, 1], size=B).astype(np.int8)
    return positions, pis, zs


def _hist(positions, N):
    from oracle import features as ofe
    out = np.zeros((len(positions), 8, N * N), np.int8)
    for b, p in enumerate(positions):
        boards = ofe.history_boards(p) if hasattr(ofe, "history_boards") else None
        if boards is None:
            f = ofe.get_feats(p)                                  # (i, j, 17): planes 2k = mine, 2k+1 = theirs
            for k in range(8):
                mine, theirs = f[:, :, 2 * k], f[:, :, 2 * k + 1]
                out[b, k] = ((mine - theirs) * p.to_play).astype(np.int8).flatten(order="F")
        else:
            for k in range(8):
                out[b, k] = np.asarray(boards[k], np.int8).flatten(order="F")
    return out, np.array([p.to_play for p in positions], np.int8)


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


@pytest.mark.parametrize("N,T,B", [(5, 1, 6), (9, 2, 8)])
def test_train_step_matches_oracle(N, T, B):
    lib_for("cuda")
    onn = onet.NeuralNet(N, T, seed=3)
    onn.randomize_bn(seed=4)
    eng = agz.Engine(N, n_games=8, readouts=8, tower_height=T, evaluator=agz.EVAL_NN_TC)
    flat = otrain.flat_params(onn)
    for k in range(3):
        eng.net_set_params(k, flat[k])
    bns = [onn.base_bns(), [onn.v_bn], [onn.p_bn]]
    for k in range(3):
        eng.net_set_bn_stats(k, np.concatenate([b.mu for b in bns[k]]), np.concatenate([b.sigma for b in bns[k]]), agz.BN_VAR_EPS)
    tr = otrain.Trainer(onn)
    for step in range(2):
        positions, pis, zs = _batch(N, B, 10 + step)
        bh, tp = _hist(positions, N)
        loss_o = tr.step(positions, pis, zs, lr=0.02, rho=0.9)
        loss_e = eng.train_step(bh, tp, pis, zs, lr=0.02, momentum=0.9)
        assert abs(loss_e - loss_o) <= 2e-4 * abs(loss_o), (step, loss_e, loss_o)
        for k in range(3):
            assert _rel(eng.train_read_grads(k), tr.last_grads[k]) < 2e-3, "chain %d gradients (step %d)" % (k, step)
        want = otrain.flat_params(onn)
        for k in range(3):
            got = eng.net_get_params(k)
            assert np.max(np.abs(got - want[k])) < 2e-5, "chain %d parameters after step %d" % (k, step)
        for k in range(3):
            mu, var, mode = eng.net_get_bn_stats(k)
            assert mode == agz.BN_VAR_EPS
            assert np.allclose(mu, np.concatenate([b.mu for b in bns[k]]), atol=2e-5, rtol=1e-4)
            assert np.allclose(var, np.concatenate([b.sigma for b in bns[k]]), atol=2e-5, rtol=1e-4)
    # the engine now evaluates (test mode, tensor-core path) with the trained parameters and the moved running statistics
    positions, _, _ = _batch(N, 4, 99)
    bh, tp = _hist(positions, N)
    pi_e, v_e = eng.net_forward(agz.EVAL_NN_TC, bh, tp)
    pi_o, v_o = onn(positions)
    assert np.max(np.abs(pi_e.T - pi_o)) < 1e-3 and np.max(np.abs(v_e - v_o)) < 1e-3
    eng.close()


def test_training_reduces_the_loss_on_a_fixed_batch():
    lib_for("cuda")
    N, T, B = 9, 1, 32
    env = agz.GoEnv(N)
    nn = agz.NeuralNet(env, tower_height=T, seed=0)
    eng = agz.Engine(N, n_games=8, readouts=8, tower_height=T, evaluator=agz.EVAL_NN_TC)
    nn.push(eng)
    positions, pis, zs = _batch(N, B, 5)
    bh, tp = _hist(positions, N)
    losses = [eng.train_step(bh, tp, pis, zs) for _ in range(12)]
    assert losses[-1] < losses[0] and all(np.isfinite(losses))
    eng.close()


@pytest.mark.parametrize("game", ["go", "gomoku"])
def test_train_loop_selfplay_replay_train(tmp_path, game):
    """train(env; ...) (src/train.jl:38-92): self-play fills the replay ring, every finished game triggers one optimisation step on
    a uniform batch once `start_training_after` tuples are there, checkpoints go through save_model.  Both games of the reference's
    GameEnv (GoEnv, GomokuEnv) go through the same loop."""
    lib_for("cuda")
    from alphago_jl_b200 import weights_io
    env = agz.GoEnv(5) if game == "go" else agz.GomokuEnv(6, 4)
    nn0 = agz.NeuralNet(env, tower_height=1, seed=2)
    before = [np.concatenate([a.flatten(order="F") for a in lst]) for lst in nn0.params]
    nn = agz.train(env, num_games=24, batch_size=16, readouts=16, tower_height=1, model=nn0, start_training_after=40, concurrent=8,
                   ckp_freq=8, model_dir=str(tmp_path), verbose=False)
    assert len(nn.train_losses) >= 8 and all(np.isfinite(nn.train_losses))
    after = [np.concatenate([a.flatten(order="F") for a in lst]) for lst in nn.params]
    assert all(np.max(np.abs(a - b)) > 0 for a, b in zip(before, after))
    assert nn.bn_mode == agz.BN_VAR_EPS and np.max(np.abs(nn.bn_mu[0])) > 0
    saved = weights_io.load_saved_model(str(tmp_path), agz.NeuralNet(env, tower_height=1, seed=9))
    assert saved.params[0][0].shape == nn.params[0][0].shape
    # the trained network plays: a short gating match against the initial one runs to completion
    games = []
    agz.evaluate(env, nn, agz.NeuralNet(env, tower_height=1, seed=2), num_games=4, ro=16, details=games)
    assert len(games) == 4 and all(g.result_string for g in games)


def _ring_engine(seed_net):
    """A small engine whose replay ring holds a few finished 5x5 games (identical for identical arguments: games are deterministic)."""
    env = agz.GoEnv(5)
    nn = agz.NeuralNet(env, tower_height=1, seed=seed_net)
    eng = agz.Engine(5, n_games=8, readouts=16, tower_height=1, evaluator=agz.EVAL_NN_TC, seed=4)
    nn.push(eng)
    eng.selfplay_start(8)
    for _ in range(400):
        pr = eng.selfplay_step(8)
        if pr.games_finished == 8:
            break
    assert pr.games_finished == 8 and pr.error == 0
    total = eng.replay_gather()
    assert total >= 32
    return env, nn, eng, total


def test_train_step_from_replay_equals_sample_then_step():
    """agz_train_step_from_replay (draw, feature planes, step and hand-over on the device) = agz_replay_sample_hist +
    agz_train_step through host buffers: same loss, bit-identical parameters and running statistics after two steps.  (The host-path
    engine has no ring of its own: the order in which concurrently finishing games enter a ring is not deterministic, so it trains
    on the batches read back from the device-path engine's ring.)"""
    lib_for("cuda")
    env, nn, dev, total = _ring_engine(7)
    host = agz.Engine(5, n_games=8, readouts=16, tower_height=1, evaluator=agz.EVAL_NN_TC, seed=4)
    nn.push(host)
    for step in range(2):
        bh, tp, pis, zs, idx = dev.replay_sample_hist(16, seed=100 + step)      # the draw the device step is about to make
        loss_d = dev.train_step_from_replay(16, seed=100 + step)
        loss_h = host.train_step(bh, tp, pis, zs)
        assert abs(loss_d - loss_h) <= 2e-6 * abs(loss_h), (step, loss_d, loss_h)   # the loss terms are summed with float atomics
    for k in range(3):
        assert np.array_equal(dev.net_get_params(k), host.net_get_params(k)), k
        assert all(np.array_equal(a, b) for a, b in zip(dev.net_get_bn_stats(k)[:2], host.net_get_bn_stats(k)[:2])), k
    dev.close()
    host.close()


def test_parameters_published_on_device_equal_a_host_reload():
    """After a training step the self-play / forward path uses parameters folded and converted ON THE DEVICE (train_publish); a fresh
    engine loaded with the same parameters through the host (agz_net_set_params: host fold + upload) must give the same outputs."""
    lib_for("cuda")
    env, nn, eng, _ = _ring_engine(3)
    for step in range(3):
        eng.train_step_from_replay(32, seed=step)
    positions, _, _ = _batch(5, 16, 1)
    bh, tp = _hist(positions, 5)
    pi_a, v_a = eng.net_forward(agz.EVAL_NN_TC, bh, tp)           # published weights
    fresh = agz.Engine(5, n_games=8, readouts=16, tower_height=1, evaluator=agz.EVAL_NN_TC)
    for k in range(3):
        fresh.net_set_params(k, eng.net_get_params(k))
        mu, var, mode = eng.net_get_bn_stats(k)
        fresh.net_set_bn_stats(k, mu, var, mode)
    pi_b, v_b = fresh.net_forward(agz.EVAL_NN_TC, bh, tp)
    assert np.array_equal(pi_a, pi_b) and np.array_equal(v_a, v_b)
    # ... and the fp32 cross-check path (built from the lazily synchronised host copy) agrees within the usual tolerance
    pi_f, v_f = eng.net_forward(agz.EVAL_NN_F32, bh, tp)
    assert np.max(np.abs(pi_f - pi_a)) < 1e-3 and np.max(np.abs(v_f - v_a)) < 1e-3
    # self-play continues on the updated network without a reload
    eng.selfplay_start(4)
    for _ in range(300):
        pr = eng.selfplay_step(8)
        if pr.games_finished == 4:
            break
    assert pr.games_finished == 4 and pr.error == 0
    eng.close()
    fresh.close()
