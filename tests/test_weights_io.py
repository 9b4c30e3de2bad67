"""the product-side BSON reader loads the reference's shipped agz weights; checked against the committed
golden fixture (whose parameters were extracted by the oracle's independent reader)."""
import os

import numpy as np
import pytest

import pkg

agz = pkg.load()
REF = "/root/reference/models"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agz_shipped_9x9.npz")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_shipped_weights_load_and_match_golden():
    from alphago_jl_b200 import weights_io
    env = agz.GoEnv(9, lib_path="unused")
    nn = agz.NeuralNet(env, tower_height=0)
    weights_io.load_reference_model(REF, nn)
    g = np.load(GOLD)
    flat = lambda lst: np.concatenate([np.asarray(a, np.float32).flatten(order="F") for a in lst])
    for k, name in enumerate(("base", "value", "policy")):
        assert np.array_equal(flat(nn.params[k]), g[name]), name
        assert np.array_equal(nn.bn_mu[k], g["bn_mu_" + name].ravel())
        assert np.array_equal(nn.bn_sigma[k], g["bn_sigma_" + name].ravel())
    assert nn.bn_mode == agz.BN_STD


def test_save_model_round_trip(tmp_path):
    """save_model (src/train.jl:14-35) writes BSON.jl array documents that both readers (product and oracle) parse back bit for bit."""
    from alphago_jl_b200 import weights_io
    from oracle import bson as obson
    env = agz.GoEnv(5, lib_path="unused")
    nn = agz.NeuralNet(env, tower_height=1, seed=7)
    rs = np.random.RandomState(1)
    nn.bn_mu = [rs.randn(*m.shape).astype(np.float32) for m in nn.bn_mu]
    nn.bn_sigma = [(0.5 + rs.rand(*m.shape)).astype(np.float32) for m in nn.bn_sigma]
    weights_io.save_model(nn, str(tmp_path))
    back = weights_io.load_saved_model(str(tmp_path), agz.NeuralNet(env, tower_height=1, seed=8))
    for k in range(3):
        assert len(back.params[k]) == len(nn.params[k])
        for a, b in zip(nn.params[k], back.params[k]):
            assert a.shape == b.shape and np.array_equal(a, b)
        assert np.array_equal(back.bn_mu[k], nn.bn_mu[k]) and np.array_equal(back.bn_sigma[k], nn.bn_sigma[k])
    # the oracle's independent reader sees the same tensors in the same order
    arrs = obson.find_arrays(obson.load(os.path.join(str(tmp_path), "weights", "agz_base.bson")))
    assert len(arrs) == len(nn.params[0]) and all(np.array_equal(a, b) for a, b in zip(arrs, nn.params[0]))


def test_save_model_converts_a_moving_std_to_a_variance(tmp_path):
    """A network loaded from the shipped files keeps its BatchNorm statistics as a moving standard deviation (BN_STD); save_model
    writes (mean, variance) and load_saved_model reads them back as BN_VAR_EPS, so the statistics survive the round trip."""
    from alphago_jl_b200 import weights_io
    env = agz.GoEnv(5, lib_path="unused")
    nn = agz.NeuralNet(env, tower_height=1, seed=3)
    rs = np.random.RandomState(2)
    nn.bn_sigma = [(0.2 + rs.rand(*m.shape)).astype(np.float32) for m in nn.bn_sigma]
    nn.bn_mode = agz.BN_STD
    weights_io.save_model(nn, str(tmp_path))
    back = weights_io.load_saved_model(str(tmp_path), agz.NeuralNet(env, tower_height=1, seed=4))
    assert back.bn_mode == agz.BN_VAR_EPS
    for k in range(3):
        assert np.array_equal(back.bn_sigma[k], nn.bn_sigma[k] ** 2)
