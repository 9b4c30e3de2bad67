"""17-plane input features: restatement of src/features.jl.  TEST INFRASTRUCTURE."""
import numpy as np


def stone_features(pos):                                 # features.jl:3-20
    """Returns (N, N, 2*planes) float64; [:, :, 2k] = (board_k == to_play), [:, :, 2k+1] = (board_k == -to_play)."""
    N, planes = pos.env.N, pos.env.planes
    num_deltas_avail = pos.board_deltas.shape[0]
    cumulative = np.cumsum(pos.board_deltas.astype(np.int16), axis=0)
    last_eight = np.repeat(pos.board.astype(np.int16)[None], planes, axis=0)
    last_eight[1:num_deltas_avail + 1] = last_eight[1:num_deltas_avail + 1] - cumulative
    last_eight[num_deltas_avail + 1:] = last_eight[num_deltas_avail]
    feats = np.zeros((N, N, 2 * planes), dtype=np.float64)
    feats[:, :, 0::2] = np.transpose(last_eight == pos.to_play, (1, 2, 0))
    feats[:, :, 1::2] = np.transpose(last_eight == -pos.to_play, (1, 2, 0))
    return feats


def color_to_play_feature(pos):                          # features.jl:22  (+1 / -1, not 0/1)
    return pos.to_play * np.ones((pos.env.N, pos.env.N, 1), dtype=np.float64)


def get_feats(pos):                                      # features.jl:26
    return np.concatenate([stone_features(pos), color_to_play_feature(pos)], axis=2)
