"""Synthetic byte streams of BASELINE.json's configs (definitions: SURVEY.md section 8d).

All generators are counter-based (splitmix64 in counter mode), so any byte range
of a stream can be produced independently -- a rank generates only its shard and
the CPU baseline sees exactly the bytes the GPU sees.  numpy only; no torch.

  zipf      config 3  iid Zipf(s=1) over 256 symbols, seed 0x5EEDC0DE
  mixed     config 4  1 MiB segments alternating text-like / ptt5-like, seed 0xB2000004
  kennedy   config 5  iid kennedy.xls-like marginals, seed 0xB2000005
"""
from __future__ import annotations

import numpy as np

_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)

SEED_ZIPF = 0x5EEDC0DE
SEED_MIXED = 0xB2000004
SEED_KENNEDY = 0xB2000005
SEGMENT = 1 << 20


def splitmix64(seed: int, first: int, count: int) -> np.ndarray:
    """Outputs first .. first+count-1 of the splitmix64 sequence started at ``seed``."""
    with np.errstate(over="ignore"):
        idx = np.arange(first + 1, first + 1 + count, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * _GAMMA
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _lut16(weights: np.ndarray) -> np.ndarray:
    """65536-entry inverse CDF: lut[u] = smallest r with C[r] >= (u + 0.5) / 65536 * C[-1]."""
    c = np.cumsum(np.asarray(weights, dtype=np.float64))
    u = (np.arange(65536, dtype=np.float64) + 0.5) / 65536.0 * c[-1]
    return np.searchsorted(c, u, side="left").astype(np.uint8)


def _iid(lut: np.ndarray, seed: int, start: int, count: int) -> np.ndarray:
    """Bytes start .. start+count-1 of the iid stream: every 64-bit draw yields four 16-bit
    indices (little-endian order) into ``lut``."""
    first = start // 4
    last = (start + count + 3) // 4
    out = np.empty((last - first) * 4, dtype=np.uint8)
    step = 1 << 22
    for a in range(first, last, step):
        b = min(a + step, last)
        draws = splitmix64(seed, a, b - a).view(np.uint16)
        out[(a - first) * 4:(b - first) * 4] = lut[draws]
    off = start - first * 4
    return out[off:off + count]


_ZIPF_W = 1.0 / (np.arange(256, dtype=np.float64) + 1.0)


def zipf(n: int, start: int = 0, seed: int = SEED_ZIPF) -> np.ndarray:
    """Config 3: byte value == Zipf rank, p(r) proportional to 1/(r+1)."""
    return _iid(_lut16(_ZIPF_W), seed, start, n)


def _kennedy_weights() -> np.ndarray:
    w = np.zeros(256, dtype=np.float64)
    head = {0x00: 0.443, 0x03: 0.155, 0x01: 0.080, 0x09: 0.077, 0x05: 0.057, 0x02: 0.027}
    for k, v in head.items():
        w[k] = v
    rest = [s for s in range(256) if s not in head]
    z = 1.0 / (np.arange(len(rest), dtype=np.float64) + 1.0)
    z *= 0.161 / z.sum()
    w[rest] = z
    return w


def kennedy(n: int, start: int = 0, seed: int = SEED_KENNEDY) -> np.ndarray:
    """Config 5: iid with kennedy.xls-like marginals (H0 about 3.6 bit)."""
    return _iid(_lut16(_kennedy_weights()), seed, start, n)


def _text_lut() -> np.ndarray:
    # 74 printable symbols, space most frequent, Zipf(1) over the rest
    alphabet = [0x20] + [ord(c) for c in "etaoinshrdlcumwfgypbvkjxqz"] + [0x0A, ord(","), ord("."), ord("'"), ord('"')]
    alphabet += [ord(c) for c in "ETAOINSHRDLCUMWFGYPBVKJXQZ"] + [ord(c) for c in "0123456789;:!?-()"]
    alphabet = alphabet[:74]
    w = np.zeros(256, dtype=np.float64)
    w[alphabet] = 1.0 / (np.arange(len(alphabet), dtype=np.float64) + 1.0)
    return _lut16(w)


def _runs_segment(seed: int, seg: int) -> np.ndarray:
    """ptt5-like segment: mostly long zero runs, short bursts of a skewed 158-symbol alphabet."""
    budget = SEGMENT // 4 + 4096  # more runs than a segment can need
    raw = splitmix64(seed ^ (0xA5A5 << 16), seg * budget, budget)
    u = (raw >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    kind = (raw & np.uint64(0xFFFF)).astype(np.float64) / 65536.0
    # 51.7 % of runs are zero runs (mean 25) and the rest bursts (mean 4): 87 % of BYTES are 0x00, as in ptt5
    is_zero = kind < 0.517
    length = np.where(is_zero, np.floor(np.log1p(-u) / np.log(1.0 - 1.0 / 25.0)),
                      np.floor(np.log1p(-u) / np.log(1.0 - 1.0 / 4.0))).astype(np.int64) + 1
    pool = np.array([0xFF, 0x0F, 0x1F, 0x07, 0xF0, 0x3F, 0x7F, 0x03, 0xE0, 0xC0, 0x80, 0xFE, 0xFC, 0xF8, 0x01]
                    + [s for s in range(2, 160) if s not in (0x0F, 0x1F, 0x07, 0x3F, 0x7F, 0x03)][:143], dtype=np.uint8)
    pw = 1.0 / (np.arange(pool.size, dtype=np.float64) + 1.0)
    pick = np.searchsorted(np.cumsum(pw) / pw.sum(), ((raw >> np.uint64(16)) & np.uint64(0xFFFFFF)).astype(np.float64)
                           / float(1 << 24), side="left").clip(0, pool.size - 1)
    value = np.where(is_zero, np.uint8(0), pool[pick]).astype(np.uint8)
    out = np.repeat(value, length)
    if out.size < SEGMENT:  # cannot happen with the budget above; keep the stream well defined anyway
        out = np.concatenate([out, np.zeros(SEGMENT - out.size, dtype=np.uint8)])
    return out[:SEGMENT]


def mixed(n: int, start: int = 0, seed: int = SEED_MIXED) -> np.ndarray:
    """Config 4: even 1 MiB segments text-like (iid), odd segments ptt5-like (runs)."""
    lut = _text_lut()
    out = np.empty(n, dtype=np.uint8)
    pos = start
    end = start + n
    while pos < end:
        seg = pos // SEGMENT
        seg_lo = seg * SEGMENT
        take = min(end, seg_lo + SEGMENT) - pos
        if seg % 2 == 0:
            out[pos - start:pos - start + take] = _iid(lut, seed, pos, take)
        else:
            out[pos - start:pos - start + take] = _runs_segment(seed, seg)[pos - seg_lo:pos - seg_lo + take]
        pos += take
    return out


GENERATORS = {"zipf": zipf, "mixed": mixed, "kennedy": kennedy}


def entropy_bits(data: np.ndarray) -> float:
    h = np.bincount(data, minlength=256).astype(np.float64)
    p = h[h > 0] / h.sum()
    return float(-(p * np.log2(p)).sum())
