"""Generates tests/golden/*.npz from the reference's shipped weights (models/**/agz_*.bson) and the oracle.

Run in the build container (needs /root/reference); the GPU box only reads the committed .npz files.
  agz_shipped_9x9.npz : the shipped 9x9 / tower_height = 0 net (Flux param lists + BN stats read from the BSON
                        files), 24 positions from seeded random play, and the oracle's (pi, v) for them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import go, net, features  # noqa: E402


def hist_stack(pos):
    f = features.stone_features(pos)                      # (N, N, 16): planes 2k / 2k+1 = mine / theirs, board k moves ago
    boards = [(f[:, :, 2 * k] - f[:, :, 2 * k + 1]) * pos.to_play for k in range(8)]
    return np.stack([b.astype(np.int8).flatten(order="F") for b in boards])


def random_positions(N, count, seed):
    env = go.GoEnv(N)
    rs = np.random.RandomState(seed)
    out = []
    while len(out) < count:
        pos = go.GoPosition(env)
        for t in range(rs.randint(0, 70)):
            legal = np.flatnonzero(go.all_legal_moves(pos)[:-1])
            mv = None if len(legal) == 0 or rs.rand() < 0.05 else go.from_flat(int(rs.choice(legal)), env)
            pos = go.play_move(pos, mv)
            if pos.done:
                break
        if not pos.done:
            out.append(pos)
    return out


def main():
    nn = net.load_shipped_agz("/root/reference/models")
    poss = random_positions(9, 24, 0)
    pi, v = nn(poss)
    flat = lambda lst: np.concatenate([np.asarray(a, np.float32).flatten(order="F") for a in lst])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "agz_shipped_9x9.npz"),
                        base=flat(nn.base_params()), value=flat(nn.value_params()), policy=flat(nn.policy_params()),
                        bn_mu_base=nn.stem_bn.mu, bn_sigma_base=nn.stem_bn.sigma, bn_mu_value=nn.v_bn.mu, bn_sigma_value=nn.v_bn.sigma,
                        bn_mu_policy=nn.p_bn.mu, bn_sigma_policy=nn.p_bn.sigma,
                        boards_hist=np.stack([hist_stack(p) for p in poss]), to_play=np.array([p.to_play for p in poss], np.int8),
                        feats=np.stack([features.get_feats(p) for p in poss]).astype(np.int8), pi=pi.T.copy(), v=v)
    print("wrote agz_shipped_9x9.npz", pi.shape, v[:4])


if __name__ == "__main__":
    main()
