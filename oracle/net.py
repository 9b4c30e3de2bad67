"""Policy/value network: restatement of src/neural_net.jl:7-33,51-73 and src/resnet.jl in torch-CPU fp32.

TEST INFRASTRUCTURE.  The arithmetic of the reference lives in Flux 0.10.4 / NNlib 0.6.6
(Manifest.toml:193-197,298-302; not vendored) -- "parity unpinned": semantics restated here are
  * Conv = true convolution (kernel flipped) + bias, weight (k1, k2, Cin, Cout), data (W, H, C, B) column-major;
  * BatchNorm test mode = gamma*(x-mu)/sqrt(sigma2+1e-5)+beta (BN_VAR_EPS), or, for the shipped
    models/agz_*.bson (older Flux), gamma*(x-mu)/sigma with a stored moving std (BN_STD);
  * Dense = W*x+b with W (out, in); softmax over the action dimension;
  * param order = Flux `params`: Conv(W,b), BatchNorm(beta,gamma), ResidualBlock(W1,b1,W2,b2,beta1,gamma1,beta2,gamma2),
    Dense(W,b) -- confirmed by the shipped weight lists (models/weights/agz_*.bson).

Layout mapping: a Julia array X[i, j, c, b] (i = board row, j = board column) is held as a torch tensor
x[b, c, j, i]; flattening x[b] row-major then equals Julia's column-major reshape (i + N*j + N^2*c).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import features

BN_VAR_EPS = 0
BN_STD = 1


def flux_conv_to_torch(W):
    """Flux (k_i, k_j, Cin, Cout) true-convolution weight -> torch cross-correlation weight (Cout, Cin, k_j, k_i)."""
    W = np.asarray(W, dtype=np.float32)
    Wf = W[::-1, ::-1, :, :]
    return torch.from_numpy(np.transpose(Wf, (3, 2, 1, 0)).copy())


class BN:
    def __init__(self, C):
        self.beta = np.zeros(C, np.float32)
        self.gamma = np.ones(C, np.float32)
        self.mu = np.zeros(C, np.float32)
        self.sigma = np.ones(C, np.float32)     # variance (BN_VAR_EPS) or std (BN_STD)
        self.mode = BN_VAR_EPS

    def __call__(self, x):
        t = lambda a: torch.from_numpy(a).view(1, -1, 1, 1)
        if self.mode == BN_VAR_EPS:
            return t(self.gamma) * (x - t(self.mu)) / torch.sqrt(t(self.sigma) + np.float32(1e-5)) + t(self.beta)
        return t(self.gamma) * (x - t(self.mu)) / t(self.sigma) + t(self.beta)

    def fold(self):
        """(scale, shift) with y = scale*x + shift."""
        den = np.sqrt(self.sigma + np.float32(1e-5)) if self.mode == BN_VAR_EPS else self.sigma
        scale = (self.gamma / den).astype(np.float32)
        return scale, (self.beta - self.mu * scale).astype(np.float32)


def glorot_uniform(rs, *dims):
    """Flux.glorot_uniform: (rand(dims) - 0.5) * sqrt(24 / sum(nfan(dims)))."""
    if len(dims) == 2:
        fan_out, fan_in = dims          # Dense weight is (out, in)
    else:
        rf = int(np.prod(dims[:-2]))
        fan_in, fan_out = dims[-2] * rf, dims[-1] * rf
    return ((rs.random_sample(dims) - 0.5) * np.sqrt(24.0 / (fan_in + fan_out))).astype(np.float32)


class NeuralNet:
    """neural_net.jl:7-33.  Parameters are numpy arrays in Flux shapes."""

    def __init__(self, N=19, tower_height=19, filters=256, planes=17, seed=0, action_space=None):
        rs = np.random.RandomState(seed)
        self.N, self.T, self.C, self.planes = N, tower_height, filters, planes
        A = N * N + 1 if action_space is None else action_space   # env.action_space (neural_net.jl:30): N^2 for Gomoku
        C = filters
        self.stem_W = glorot_uniform(rs, 3, 3, planes, C); self.stem_b = np.zeros(C, np.float32); self.stem_bn = BN(C)
        self.blocks = []
        for _ in range(tower_height):
            blk = dict(W1=glorot_uniform(rs, 3, 3, C, C), b1=np.zeros(C, np.float32),
                       W2=glorot_uniform(rs, 3, 3, C, C), b2=np.zeros(C, np.float32), bn1=BN(C), bn2=BN(C))
            self.blocks.append(blk)
        self.v_W = glorot_uniform(rs, 1, 1, C, 1); self.v_b = np.zeros(1, np.float32); self.v_bn = BN(1)
        self.v_D1W = glorot_uniform(rs, 256, N * N); self.v_D1b = np.zeros(256, np.float32)
        self.v_D2W = glorot_uniform(rs, 1, 256); self.v_D2b = np.zeros(1, np.float32)
        self.p_W = glorot_uniform(rs, 1, 1, C, 2); self.p_b = np.zeros(2, np.float32); self.p_bn = BN(2)
        self.p_DW = glorot_uniform(rs, A, 2 * N * N); self.p_Db = np.zeros(A, np.float32)

    def randomize_bn(self, seed=1):
        """Give every BatchNorm non-trivial statistics (for stronger parity tests)."""
        rs = np.random.RandomState(seed)
        for bn in self.all_bns():
            C = bn.gamma.shape[0]
            bn.gamma = (1 + 0.2 * rs.randn(C)).astype(np.float32)
            bn.beta = (0.1 * rs.randn(C)).astype(np.float32)
            bn.mu = (0.1 * rs.randn(C)).astype(np.float32)
            bn.sigma = (0.5 + rs.random_sample(C)).astype(np.float32)

    def sharpen(self, seed=1, policy_gain=6.0, bias_scale=0.05):
        """Trained-like statistics for parity tests: non-zero conv / dense biases and a policy head whose logits spread over tens of
        units (a Glorot-init net gives a near-uniform policy, under which wrong arithmetic can hide)."""
        rs = np.random.RandomState(seed)
        self.stem_b = (bias_scale * rs.randn(self.C)).astype(np.float32)
        for b in self.blocks:
            b["b1"] = (bias_scale * rs.randn(self.C)).astype(np.float32)
            b["b2"] = (bias_scale * rs.randn(self.C)).astype(np.float32)
        self.v_b = (bias_scale * rs.randn(1)).astype(np.float32)
        self.p_b = (bias_scale * rs.randn(2)).astype(np.float32)
        self.v_D1b = (bias_scale * rs.randn(256)).astype(np.float32)
        self.v_D2b = (bias_scale * rs.randn(1)).astype(np.float32)
        self.p_Db = (bias_scale * rs.randn(self.p_Db.shape[0])).astype(np.float32)
        self.p_DW = (self.p_DW * np.float32(policy_gain)).astype(np.float32)
        self.v_D2W = (self.v_D2W * np.float32(2.0)).astype(np.float32)

    def calibrate_bn(self, x, seed=1, jitter=0.1):
        """Set every BatchNorm's running statistics to the statistics of its own input over the batch x (what training's moving
        averages converge to), times a (1 +- jitter) perturbation, with random gamma / beta: activations stay O(1) at every depth, as
        in a trained network, instead of growing block by block."""
        rs = np.random.RandomState(seed)
        t = torch.from_numpy

        def fit(bn, z):
            C = z.shape[1]
            mu = z.mean(dim=(0, 2, 3)).numpy()
            var = z.var(dim=(0, 2, 3), unbiased=False).numpy()
            bn.mode = BN_VAR_EPS
            bn.mu = (mu * (1 + jitter * rs.randn(C))).astype(np.float32)
            bn.sigma = (np.maximum(var, 1e-4) * (1 + jitter * rs.rand(C))).astype(np.float32)
            bn.gamma = (1 + 0.2 * rs.randn(C)).astype(np.float32)
            bn.beta = (0.2 * rs.randn(C)).astype(np.float32)
            return bn(z)

        with torch.no_grad():
            h = F.relu(fit(self.stem_bn, F.conv2d(x, flux_conv_to_torch(self.stem_W), t(self.stem_b), padding=1)))
            for b in self.blocks:
                y = F.relu(fit(b["bn1"], F.conv2d(h, flux_conv_to_torch(b["W1"]), t(b["b1"]), padding=1)))
                y = fit(b["bn2"], F.conv2d(y, flux_conv_to_torch(b["W2"]), t(b["b2"]), padding=1))
                h = F.relu(y + h)
            fit(self.v_bn, F.conv2d(h, flux_conv_to_torch(self.v_W), t(self.v_b)))
            fit(self.p_bn, F.conv2d(h, flux_conv_to_torch(self.p_W), t(self.p_b)))

    def all_bns(self):
        out = [self.stem_bn]
        for b in self.blocks:
            out += [b["bn1"], b["bn2"]]
        return out + [self.v_bn, self.p_bn]

    # -- Flux `params` order, per chain (what save_model writes, train.jl:27-33)
    def base_params(self):
        out = [self.stem_W, self.stem_b, self.stem_bn.beta, self.stem_bn.gamma]
        for b in self.blocks:
            out += [b["W1"], b["b1"], b["W2"], b["b2"], b["bn1"].beta, b["bn1"].gamma, b["bn2"].beta, b["bn2"].gamma]
        return out

    def value_params(self):
        return [self.v_W, self.v_b, self.v_bn.beta, self.v_bn.gamma, self.v_D1W, self.v_D1b, self.v_D2W, self.v_D2b]

    def policy_params(self):
        return [self.p_W, self.p_b, self.p_bn.beta, self.p_bn.gamma, self.p_DW, self.p_Db]

    def base_bns(self):
        out = [self.stem_bn]
        for b in self.blocks:
            out += [b["bn1"], b["bn2"]]
        return out

    def load_flux_lists(self, base, value, policy):
        it = iter(base)
        self.stem_W, self.stem_b, self.stem_bn.beta, self.stem_bn.gamma = next(it), next(it), next(it), next(it)
        for b in self.blocks:
            b["W1"], b["b1"], b["W2"], b["b2"] = next(it), next(it), next(it), next(it)
            b["bn1"].beta, b["bn1"].gamma, b["bn2"].beta, b["bn2"].gamma = next(it), next(it), next(it), next(it)
        (self.v_W, self.v_b, self.v_bn.beta, self.v_bn.gamma, self.v_D1W, self.v_D1b, self.v_D2W, self.v_D2b) = value
        (self.p_W, self.p_b, self.p_bn.beta, self.p_bn.gamma, self.p_DW, self.p_Db) = policy

    # -- forward ------------------------------------------------------------
    def forward_feats(self, x):
        """x: torch (B, 17, N, N) laid out [b, c, j, i].  Returns (pi (A, B), v (B,)) as numpy float32."""
        with torch.no_grad():
            t = torch.from_numpy
            h = F.conv2d(x, flux_conv_to_torch(self.stem_W), t(self.stem_b), padding=1)
            h = F.relu(self.stem_bn(h))                                   # neural_net.jl:19-20
            for b in self.blocks:                                         # resnet.jl:26-32
                y = F.relu(b["bn1"](F.conv2d(h, flux_conv_to_torch(b["W1"]), t(b["b1"]), padding=1)))
                y = b["bn2"](F.conv2d(y, flux_conv_to_torch(b["W2"]), t(b["b2"]), padding=1))
                h = F.relu(y + h)
            B = x.shape[0]
            v = F.relu(self.v_bn(F.conv2d(h, flux_conv_to_torch(self.v_W), t(self.v_b))))   # :23-24
            v = v.reshape(B, -1)
            v = F.relu(v @ t(self.v_D1W).T + t(self.v_D1b))               # :25
            v = torch.tanh(v @ t(self.v_D2W).T + t(self.v_D2b))           # :26
            p = F.relu(self.p_bn(F.conv2d(h, flux_conv_to_torch(self.p_W), t(self.p_b))))   # :28-29
            p = p.reshape(B, -1)
            p = torch.softmax(p @ t(self.p_DW).T + t(self.p_Db), dim=1)   # :30
            return p.numpy().T.copy(), v.numpy()[:, 0].copy()

    def forward_debug(self, x):
        """forward_feats with the intermediate values: trunk after the stem and after every block (list of (B, C, N*N) arrays in
        the reference's W x H x C order, p = N*j + i), logits (B, A) before softmax, v_pre (B,) before tanh, pi (B, A), v (B,)."""
        with torch.no_grad():
            t = torch.from_numpy
            B = x.shape[0]
            flat = lambda h: h.reshape(B, self.C, -1).numpy().copy()     # [b, c, j, i] row-major = p = N*j + i
            h = F.conv2d(x, flux_conv_to_torch(self.stem_W), t(self.stem_b), padding=1)
            h = F.relu(self.stem_bn(h))
            trunks = [flat(h)]
            for b in self.blocks:
                y = F.relu(b["bn1"](F.conv2d(h, flux_conv_to_torch(b["W1"]), t(b["b1"]), padding=1)))
                y = b["bn2"](F.conv2d(y, flux_conv_to_torch(b["W2"]), t(b["b2"]), padding=1))
                h = F.relu(y + h)
                trunks.append(flat(h))
            v = F.relu(self.v_bn(F.conv2d(h, flux_conv_to_torch(self.v_W), t(self.v_b)))).reshape(B, -1)
            v = F.relu(v @ t(self.v_D1W).T + t(self.v_D1b))
            v_pre = (v @ t(self.v_D2W).T + t(self.v_D2b))[:, 0]
            p = F.relu(self.p_bn(F.conv2d(h, flux_conv_to_torch(self.p_W), t(self.p_b)))).reshape(B, -1)
            logits = p @ t(self.p_DW).T + t(self.p_Db)
            return {"trunks": trunks, "logits": logits.numpy().copy(), "v_pre": v_pre.numpy().copy(),
                    "pi": torch.softmax(logits, dim=1).numpy().copy(), "v": torch.tanh(v_pre).numpy().copy()}

    @staticmethod
    def feats_to_torch(positions):
        fs = np.stack([features.get_feats(p) for p in positions])        # (B, i, j, c)
        return torch.from_numpy(np.ascontiguousarray(np.transpose(fs, (0, 3, 2, 1))).astype(np.float32))

    def __call__(self, positions):                                        # neural_net.jl:57-68
        return self.forward_feats(self.feats_to_torch(positions))


def load_shipped_agz(models_dir):
    """The shipped 9x9, tower_height = 0 net (models/weights/agz_*.bson + BN stats from models/agz_*.bson)."""
    from . import bson
    nn = NeuralNet(N=9, tower_height=0)
    lists = []
    for name in ("base", "value", "policy"):
        doc = bson.load("%s/weights/agz_%s.bson" % (models_dir, name))
        lists.append(bson.find_arrays(doc))
    nn.load_flux_lists(*lists)
    for name, bn in (("base", nn.stem_bn), ("value", nn.v_bn), ("policy", nn.p_bn)):
        arrs = bson.find_arrays(bson.load("%s/agz_%s.bson" % (models_dir, name)))
        bn.mu, bn.sigma, bn.mode = arrs[0], arrs[1], BN_STD
    return nn
