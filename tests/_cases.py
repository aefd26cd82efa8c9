"""Seeded inputs shared by the CPU (oracle / sim) and GPU parity tests."""
from __future__ import annotations

import numpy as np


def crafted(kind: int, n: int, rng: np.random.Generator) -> np.ndarray:
    """Byte patterns chosen to reach the rare coder branches (long 0xFF runs, carries into
    pending runs, single-symbol blocks, tiny alphabets)."""
    if kind == 0:
        return rng.integers(0, 256, n, dtype=np.uint8)
    if kind == 1:
        return (rng.integers(0, int(rng.integers(1, 5)), n) + int(rng.integers(0, 252))).astype(np.uint8)
    if kind == 2:
        return rng.choice(256, n, p=rng.dirichlet(np.ones(256) * 0.05)).astype(np.uint8)
    if kind == 3:
        reps = rng.integers(1, 100, n // 50 + 1)
        d = np.repeat(rng.integers(0, 256, n // 50 + 1, dtype=np.uint8), reps)[:n]
        return np.concatenate([d, np.zeros(n - d.size, np.uint8)])
    if kind == 4:
        d = np.full(n, 255, dtype=np.uint8)
        k = max(1, n // 1000)
        d[rng.integers(0, n, k)] = rng.integers(0, 256, k, dtype=np.uint8)
        return d
    if kind == 5:  # half zeros, half one other symbol: thousands of pending 0xFF bytes, then carries
        h = n // 2
        return np.concatenate([np.zeros(h, np.uint8), np.full(n - h, int(rng.integers(1, 256)), np.uint8)])
    return np.full(n, int(rng.integers(0, 256)), np.uint8)


def crafted_stream(nblocks: int, block: int, seed: int, ragged: int = 0) -> np.ndarray:
    """`nblocks` blocks, each its own crafted kind, optionally followed by a short last block."""
    rng = np.random.default_rng(seed)
    parts = [crafted(i % 7, block, rng) for i in range(nblocks)]
    if ragged:
        parts.append(crafted(int(rng.integers(0, 7)), ragged, rng))
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)
