"""Helpers of the network parity tests: compare an engine forward (agz_net_forward_debug) with the fp32 oracle at the level of
logits / pre-tanh value / per-block trunk activations, and bisect a deviation to the first residual block that exceeds its bound."""
import numpy as np

from oracle import features as ofeat


def hist_stack(pos):
    f = ofeat.stone_features(pos)
    return np.stack([((f[:, :, 2 * k] - f[:, :, 2 * k + 1]) * pos.to_play).astype(np.int8).flatten(order="F") for k in range(8)])


def engine_inputs(positions):
    return np.stack([hist_stack(p) for p in positions]), np.array([p.to_play for p in positions], np.int8)


def rel_rms(a, b):
    return float(np.sqrt(np.sum((a.astype(np.float64) - b) ** 2) / max(1e-30, np.sum(b.astype(np.float64) ** 2))))


def report(eng_out, ref):
    """Deviation metrics of a whole-network forward.  Logits are compared after removing each position's mean (softmax ignores it)."""
    lg_e = eng_out["logits"] - eng_out["logits"].mean(axis=1, keepdims=True)
    lg_r = ref["logits"] - ref["logits"].mean(axis=1, keepdims=True)
    return {"pi_abs": float(np.max(np.abs(eng_out["pi"] - ref["pi"]))),
            "v_abs": float(np.max(np.abs(eng_out["v"] - ref["v"]))),
            "logit_rel": float(np.max(np.abs(lg_e - lg_r)) / max(1.0, float(np.max(np.abs(lg_r))))),
            "vpre_rel": float(np.max(np.abs(eng_out["v_pre"] - ref["v_pre"]) / np.maximum(1.0, np.abs(ref["v_pre"])))),
            "logit_span": float(np.max(lg_r) - np.min(lg_r)), "pi_max": float(np.max(ref["pi"]))}


def trunk_errors(eng, evaluator, positions, ref, blocks):
    """Relative RMS deviation of the trunk after the stem (block 0) and after each listed number of blocks."""
    bh, tp = engine_inputs(positions)
    return {nb: rel_rms(eng.net_forward_debug(evaluator, bh, tp, n_blocks=nb, want_trunk=True)["trunk"], ref["trunks"][nb]) for nb in blocks}


def first_bad_block(eng, evaluator, positions, ref, tower_height, bound):
    """Binary search for the first block whose trunk deviates by more than bound(n_blocks) (None if all are inside)."""
    bh, tp = engine_inputs(positions)
    err = lambda nb: rel_rms(eng.net_forward_debug(evaluator, bh, tp, n_blocks=nb, want_trunk=True)["trunk"], ref["trunks"][nb])
    if err(tower_height) <= bound(tower_height):
        return None
    lo, hi = -1, tower_height           # invariant: hi is bad, lo is good (or -1 = before the stem)
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if err(mid) > bound(mid):
            hi = mid
        else:
            lo = mid
    return hi
