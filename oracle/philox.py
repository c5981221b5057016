"""ORACLE (test infrastructure).  Philox4x32-10 (Salmon et al., SC'11) in numpy.

The reference draws resets from unseeded global Mersenne-Twister streams
(graph/util.py:132,142; environments/gym_graph/graph.py:47), so bit-parity with the reference is
only defined under injected reset streams.  The device path's own per-env counter-based RNG is
Philox4x32-10 keyed by ``seed`` with counter ``(global_env_id, reset_epoch, 0, 0)``; this file is
its independent CPU restatement, checked against the Random123 known-answer vectors in
tests/test_oracle_golden.py (test_philox_known_answers).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(counter, key, rounds=10):
    """counter: uint32 [..., 4]; key: (k0, k1) ints.  Returns uint32 [..., 4]."""
    c = np.asarray(counter, dtype=np.uint64) & MASK
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], -1).astype(np.uint32)


def reset_draws(seed, env_id, epoch):
    """The 4 uint32 a reset of global env ``env_id`` at reset-epoch ``epoch`` consumes:
    [0] task choice, [1] curriculum bucket (un-oriented 0.9/0.1 rule), [2] candidate index, [3] spare."""
    env_id = np.asarray(env_id, dtype=np.uint32)
    epoch = np.asarray(epoch, dtype=np.uint32)
    ctr = np.stack(np.broadcast_arrays(env_id, epoch, np.uint32(0), np.uint32(0)), -1)
    return philox4x32(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


def mulhi(r, n):
    """floor(r * n / 2^32): maps a uint32 draw to [0, n)."""
    return (np.asarray(r, dtype=np.uint64) * np.uint64(n) >> np.uint64(32)).astype(np.int64)


#: P(bucket = "<= optimal distance") = 0.9 as a uint32 threshold (graph/util.py:112-113)
BUCKET_THRESHOLD = 3865470566   # floor(0.9 * 2^32)
