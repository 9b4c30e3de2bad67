"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  get_replay_batch (src/train.jl:4-12): a uniform draw of `batch` distinct
tuples from the replay buffer.  The reference draws with StatsBase.sample on Julia's global RNG (unpinned); the engine and this
restatement share one explicit spec instead: draw k is tuple oldest + perm(k), perm = a keyed pseudo-random permutation of [0, n)
(6-round Feistel network over 2w >= log2(n) bits, murmur3-finaliser round function, splitmix64 round keys, cycle walking)."""
M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


def _mix32(h):
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def feistel_keys(seed, n):
    bits = 2
    while (1 << bits) < n:
        bits += 1
    if bits & 1:
        bits += 1
    x, keys = seed & M64, []
    for _ in range(6):
        x = (x + 0x9E3779B97F4A7C15) & M64
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        keys.append(((z ^ (z >> 31)) >> 32) & M32)
    return keys, bits // 2


def feistel_perm(keys, w, k, n):
    mask = (1 << w) - 1
    y = k
    while True:
        L, R = y >> w, y & mask
        for key in keys:
            L, R = R, L ^ (_mix32(R ^ key) & mask)
        y = (L << w) | R
        if y < n:
            return y


def sample_indices(total, capacity, batch, seed):
    """Ring indices agz_replay_sample returns: `total` tuples ever appended to a ring of `capacity`."""
    oldest = max(0, total - capacity)
    n = total - oldest
    assert 0 <= batch <= n
    keys, w = feistel_keys(seed, n)
    return [oldest + feistel_perm(keys, w, k, n) for k in range(batch)]
