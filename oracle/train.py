"""One optimisation step: restatement of `_train`'s minibatch body (src/neural_net.jl:91-97), the losses (:75-83) and
`Momentum(2f-2)` (src/train.jl:54) with torch autograd on the CPU, fp32.  TEST INFRASTRUCTURE.

"Parity unpinned": the reference's `_train` does not run as committed (NamedTuple signature vs tuple call site, undefined
`loss_avg`) and the layer / loss / optimiser semantics live in Flux 0.10.4, which is not vendored.  Restated here:
  * loss = 0.01 * crossentropy(p, pi) + 0.01 * mse(z, v) + 1e-4 * sum(theta .^ 2) over *all* params(nn) (weights, biases,
    BatchNorm beta / gamma), crossentropy = -sum(pi .* log.(p)) / B, mse = sum((z .- v) .^ 2) / B;
  * train-mode BatchNorm: batch mean, biased batch variance, eps = 1e-5; running mean / variance moved with momentum 0.1,
    the variance unbiased by m / (m - 1);
  * Momentum(eta, rho = 0.9): v = rho * v - eta * grad; theta += v.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import net as onet

BN_EPS, BN_MOMENTUM = 1e-5, 0.1
W_POLICY, W_VALUE, W_REG = 0.01, 0.01, 1e-4


def _conv_w(W):
    """Flux (k_i, k_j, Cin, Cout) true-convolution weight (torch tensor) -> cross-correlation weight (Cout, Cin, k_j, k_i)."""
    return torch.flip(W, dims=[0, 1]).permute(3, 2, 1, 0)


def _bn_train(x, bn, gamma, beta, new_stats):
    mu = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    m = x.shape[0] * x.shape[2] * x.shape[3]
    new_stats.append((bn, mu.detach().numpy().copy(), (var.detach() * (m / max(1, m - 1))).numpy().copy()))
    return gamma.view(1, -1, 1, 1) * (x - mu.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS) + beta.view(1, -1, 1, 1)


class Trainer:
    """Holds the momentum buffers for one oracle NeuralNet (parameters are updated in place, Flux shapes)."""

    def __init__(self, nn):
        self.nn = nn
        self.vel = None

    def _leaves(self):
        nn = self.nn
        lists = [nn.base_params(), nn.value_params(), nn.policy_params()]
        return [[torch.tensor(np.asarray(a, np.float32), requires_grad=True) for a in lst] for lst in lists]

    def grads(self, positions_or_feats, pis, zs):
        """Train-mode forward + back-propagation of the data loss.  Returns (loss incl. regulariser, leaves, gradients, BatchNorm
        batch statistics); nothing is updated."""
        nn = self.nn
        x = positions_or_feats if isinstance(positions_or_feats, torch.Tensor) else onet.NeuralNet.feats_to_torch(positions_or_feats)
        pi = torch.from_numpy(np.asarray(pis, np.float32))          # (B, A)
        z = torch.from_numpy(np.asarray(zs, np.float32))            # (B,)
        B = x.shape[0]
        base, value, policy = self._leaves()
        stats = []
        it = iter(base)
        W, b, beta, gamma = next(it), next(it), next(it), next(it)
        h = F.relu(_bn_train(F.conv2d(x, _conv_w(W), b, padding=1), nn.stem_bn, gamma, beta, stats))
        for blk in nn.blocks:                                        # resnet.jl:26-32
            W1, b1, W2, b2, be1, ga1, be2, ga2 = (next(it) for _ in range(8))
            y = F.relu(_bn_train(F.conv2d(h, _conv_w(W1), b1, padding=1), blk["bn1"], ga1, be1, stats))
            y = _bn_train(F.conv2d(y, _conv_w(W2), b2, padding=1), blk["bn2"], ga2, be2, stats)
            h = F.relu(y + h)
        vW, vb, vbe, vga, D1W, D1b, D2W, D2b = value
        v = F.relu(_bn_train(F.conv2d(h, _conv_w(vW), vb), nn.v_bn, vga, vbe, stats)).reshape(B, -1)
        v = F.relu(v @ D1W.T + D1b)
        v = torch.tanh(v @ D2W.T + D2b)[:, 0]
        pW, pb, pbe, pga, DW, Db = policy
        p = F.relu(_bn_train(F.conv2d(h, _conv_w(pW), pb), nn.p_bn, pga, pbe, stats)).reshape(B, -1)
        logp = torch.log_softmax(p @ DW.T + Db, dim=1)
        loss_pi = W_POLICY * (-(pi * logp).sum() / B)                # loss_pi (neural_net.jl:75)
        loss_v = W_VALUE * ((z - v) ** 2).sum() / B                  # loss_value (:77)
        data_loss = loss_pi + loss_v
        leaves = base + value + policy
        data_grads = torch.autograd.grad(data_loss, leaves, retain_graph=False)
        reg = W_REG * sum((t.detach() ** 2).sum() for t in leaves)   # loss_reg (:80-83)
        self._split = (len(base), len(value))
        return float(data_loss.detach() + reg), leaves, list(data_grads), stats

    def apply(self, leaves, data_grads, stats_list, lr=0.02, rho=0.9):
        """Momentum (train.jl:54) with the regulariser's gradient, then the running statistics.  `stats_list` holds the batch
        statistics of every contributing minibatch (one for the reference's single-process step; several = their mean)."""
        nn = self.nn
        if self.vel is None:
            self.vel = [torch.zeros_like(t) for t in leaves]
        new = []
        for t, g, vel in zip(leaves, data_grads, self.vel):
            g_total = g + 2.0 * W_REG * t.detach()
            vel.mul_(rho).sub_(lr * g_total)
            new.append((t.detach() + vel).numpy().astype(np.float32))
        nb, nv = self._split
        nn.load_flux_lists(new[:nb], new[nb:nb + nv], new[nb + nv:])
        k = len(stats_list)
        for entries in zip(*stats_list):                             # the same BatchNorm layer in every minibatch
            bn = entries[0][0]
            cur_var = bn.sigma ** 2 if bn.mode == onet.BN_STD else bn.sigma
            mu = sum(((1 - BN_MOMENTUM) * bn.mu + BN_MOMENTUM * e[1]) for e in entries) / k
            var = sum(((1 - BN_MOMENTUM) * cur_var + BN_MOMENTUM * e[2]) for e in entries) / k
            bn.mu, bn.sigma, bn.mode = mu.astype(np.float32), var.astype(np.float32), onet.BN_VAR_EPS
        lists = [data_grads[:nb], data_grads[nb:nb + nv], data_grads[nb + nv:]]
        self.last_grads = [np.concatenate([g.numpy().flatten(order="F") for g in lst]).astype(np.float32) for lst in lists]

    def step(self, positions_or_feats, pis, zs, lr=0.02, rho=0.9):
        loss, leaves, grads, stats = self.grads(positions_or_feats, pis, zs)
        self.apply(leaves, grads, [stats], lr, rho)
        return loss

    def step_data_parallel(self, batches, lr=0.02, rho=0.9):
        """The engine's multi-GPU extension: one minibatch per rank, gradients / loss / moved running statistics averaged."""
        outs = [self.grads(*b) for b in batches]
        k = len(outs)
        grads = [sum(o[2][i] for o in outs) / k for i in range(len(outs[0][2]))]
        self.apply(outs[0][1], grads, [o[3] for o in outs], lr, rho)
        return sum(o[0] for o in outs) / k


def flat_params(nn):
    """The three Flux `params` lists flattened column-major (the engine's agz_net_set_params / agz_net_get_params layout)."""
    return [np.concatenate([np.asarray(a, np.float32).flatten(order="F") for a in lst])
            for lst in (nn.base_params(), nn.value_params(), nn.policy_params())]
