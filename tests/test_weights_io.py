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
