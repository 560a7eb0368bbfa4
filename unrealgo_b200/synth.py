"""Synthetic packed leaves for benchmarks (product-side: no oracle involved).

Each position is a stone sequence of random length placed on distinct points with alternating colours; history
step k is the board with the last k stones removed (no captures: this generator only has to produce inputs of
the right shape and density for throughput runs — parity tests use real games from the oracle's board).
Layout = ug_packed_pos (include/ugeval.h)."""
import numpy as np

WORDS = 196


def synthetic_packed_positions(n, seed=20260119, max_stones=300):
    rng = np.random.default_rng(seed)
    out = np.zeros((n, WORDS), np.uint32)
    for i in range(n):
        m = int(rng.integers(0, max_stones + 1))
        pts = rng.permutation(361)[:m]
        for k in range(8):
            keep = max(m - k, 0)
            for colour in (0, 1):
                sel = pts[colour:keep:2]
                words = np.zeros(12, np.uint32)
                np.bitwise_or.at(words, sel >> 5, (np.uint32(1) << (sel & 31).astype(np.uint32)))
                out[i, colour * 96 + k * 12: colour * 96 + (k + 1) * 12] = words
        out[i, 192] = m & 1  # black moves first
    return out


def packed_to_planes(packed):
    """host-side expansion to the reference's char[n][17][19][19] (used to feed the drop-in Evaluate() entry in
    benchmarks; the GPU encoder is what production uses)."""
    packed = np.ascontiguousarray(packed, np.uint32).reshape(-1, WORDS)
    n = packed.shape[0]
    planes = np.zeros((n, 17, 361), np.int8)
    shifts = np.arange(32, dtype=np.uint32)
    for i in range(n):
        tp = int(packed[i, 192])
        for k in range(8):
            for colour in (0, 1):
                w = packed[i, colour * 96 + k * 12: colour * 96 + (k + 1) * 12]
                bits = ((w[:, None] >> shifts) & 1).reshape(-1)[:361]
                planes[i, 2 * k + (0 if colour == tp else 1)] = bits
        planes[i, 16] = 1 - tp
    return planes.reshape(n, 17, 19, 19)
