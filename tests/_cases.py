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


def blk_cases() -> list:
    """(label, bytes) inputs for the block-sort transform (blksort.h), shared by the CPU and GPU tests:
    whole Canterbury files (full 32 KiB blocks + a raw tail), inputs shorter than a block, exact multiples,
    long repeats (many doubling rounds) and blocks with a period (equal rotations: the tie cases)."""
    from _oracle import CANTERBURY, canterbury
    rng = np.random.default_rng(0xB150)
    n = 32768
    out = [(name, np.frombuffer(canterbury(name), dtype=np.uint8)) for name in CANTERBURY]
    out += [
        ("empty", np.zeros(0, np.uint8)),
        ("short100", rng.integers(0, 256, 100, dtype=np.uint8)),
        ("one-less", rng.integers(0, 256, n - 1, dtype=np.uint8)),
        ("random-3-blocks+777", rng.integers(0, 256, 3 * n + 777, dtype=np.uint8)),
        ("bits-2-blocks", rng.integers(0, 2, 2 * n, dtype=np.uint8)),
        ("runs64", np.repeat(rng.integers(0, 3, n // 64, dtype=np.uint8), 64)),
        ("zeros-then-one", np.concatenate([np.zeros(n - 1, np.uint8), np.ones(1, np.uint8)])),
        ("two-halves-differ-last", np.concatenate([np.tile(rng.integers(0, 4, n // 2, dtype=np.uint8), 2)[:-1],
                                                   np.full(1, 9, np.uint8)])),
        ("crafted-7", crafted_stream(7, n, 0xB151, ragged=5000)),
    ]
    return out


def blk_periodic_cases() -> list:
    """Blocks whose rotations tie (period 1 ... 16 384).  The reference takes seconds for each on a CPU."""
    rng = np.random.default_rng(0xB152)
    n = 32768
    out = [("period1", np.zeros(n, np.uint8))]
    for p in (2, 4, 64, 1024, 16384):
        base = rng.integers(0, 256 if p <= 64 else 4, p, dtype=np.uint8)
        if p > 1 and np.all(base == base[0]):
            base[0] ^= 1
        out.append((f"period{p}", np.tile(base, n // p)))
    for p, alpha in ((16384, 2), (4096, 2), (8192, 3), (256, 2)):  # few symbols: deep recursion, the heapsort fallback
        out.append((f"period{p}-alpha{alpha}", np.tile(rng.integers(0, alpha, p, dtype=np.uint8), n // p)))
    out.append(("periodic-among-others", np.concatenate([rng.integers(0, 256, n, dtype=np.uint8), np.tile(np.arange(8, dtype=np.uint8), n // 8),
                                                         rng.integers(0, 7, n + 99, dtype=np.uint8)])))
    return out


def blk_fuzz_stream(seed: int, nblocks: int = 48) -> np.ndarray:
    """Blocks of mixed structure for the block sort: iid over small and large alphabets, geometric runs, a
    chunk repeated with a few mutations (long common prefixes, many doubling rounds, the switch between full
    and short rounds at different depths), slices of text."""
    from _oracle import canterbury
    rng = np.random.default_rng(seed)
    n = 32768
    text = np.frombuffer(canterbury("lcet10.txt"), dtype=np.uint8)
    parts = []
    for b in range(nblocks):
        kind = b % 6
        if kind == 0:
            blk = rng.integers(0, int(rng.choice([2, 3, 5, 17, 64, 256])), n, dtype=np.uint8)
        elif kind == 1:
            lens = rng.geometric(1.0 / float(rng.choice([3, 20, 200])), n)
            blk = np.repeat(rng.integers(0, 4, n, dtype=np.uint8), lens)[:n]
        elif kind == 2:
            chunk = rng.integers(0, 8, int(rng.choice([100, 1000, 5000])), dtype=np.uint8)
            blk = np.tile(chunk, n // chunk.size + 1)[:n].copy()
            hits = rng.integers(0, n, int(rng.choice([1, 5, 50])))
            blk[hits] = rng.integers(8, 16, hits.size, dtype=np.uint8)
        elif kind == 3:
            at = int(rng.integers(0, text.size - n))
            blk = text[at:at + n].copy()
        elif kind == 4:
            blk = rng.choice(256, n, p=rng.dirichlet(np.ones(256) * 0.02)).astype(np.uint8)
        else:
            blk = np.zeros(n, np.uint8)
            k = int(rng.choice([1, 2, 30]))
            blk[rng.integers(0, n, k)] = 1
        parts.append(blk)
    parts.append(rng.integers(0, 256, int(rng.integers(0, n)), dtype=np.uint8))
    return np.concatenate(parts)
