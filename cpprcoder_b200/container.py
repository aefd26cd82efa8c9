"""The B2RC container (framing documented in include/b2rc.h) on the host, in numpy.

Used by the multi-GPU path to stitch per-rank shards into one container, and by
tests to pull payloads apart.  Pure bookkeeping: no coding happens here.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

MAGIC = 0x43523242  # 'B','2','R','C'
HEADER = 32
MIN_BLOCK, MAX_BLOCK = 64, 1 << 23
ADAPTIVE_RESTART_WORDS = 131  # u32 per restart point of the adaptive coder: bytes shifted, low, range, 256 x u16 counts
ADAPTIVE_RESTART_WORDS_WIDE = 259  # ... for blocks above 65536 bytes: 256 x u32 counts


def nblocks_of(n: int, block: int) -> int:
    return (n + block - 1) // block


def block_ok(block: int) -> bool:
    return MIN_BLOCK <= block <= MAX_BLOCK and block % 64 == 0


@dataclass
class Info:
    mode: int
    block: int
    total: int
    nblocks: int
    offsets: np.ndarray  # uint64, nblocks + 1
    payload_base: int
    seg_syms: int = 0                 # restart points (static range coder, byte rANS) every so many symbols (0: none)
    restart: np.ndarray | None = None  # uint32 [nblocks][records][3]: static: bytes shifted, encoder low, range;
    #                                    byte rANS: coded bytes still ahead of the decoder, its state x, 0;
    #                                    adaptive: [..][131]: the static coder's three, then 256 u16 symbol counts

    def payload(self, buf: np.ndarray, b: int) -> np.ndarray:
        lo = self.payload_base + int(self.offsets[b])
        hi = self.payload_base + int(self.offsets[b + 1])
        return buf[lo:hi]


def restart_records(block: int, seg_syms: int) -> int:
    ok = seg_syms >= 64 and seg_syms % 64 == 0 and seg_syms < block
    return (block + seg_syms - 1) // seg_syms - 1 if ok else 0


def pack_header(mode: int, block: int, total: int, nblocks: int, seg_syms: int = 0) -> bytes:
    flags = (1 | ((seg_syms // 64) << 8)) if seg_syms else 0
    return struct.pack("<IHHIIQQ", MAGIC, 1, mode, block, flags, total, nblocks)


def parse(buf) -> Info:
    buf = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    if buf.size < HEADER + 8:
        raise ValueError("container shorter than its header")
    magic, version, mode, block, flags, total, nblocks = struct.unpack("<IHHIIQQ", buf[:HEADER].tobytes())
    if magic != MAGIC or version != 1 or mode > 3 or not block_ok(block):
        raise ValueError("bad container header")
    seg_syms = (flags >> 8) * 64
    if flags and ((flags & 0xFF) != 1 or mode not in (0, 1, 2) or restart_records(block, seg_syms) == 0):
        raise ValueError("bad container flags")
    if nblocks != nblocks_of(total, block) or HEADER + 8 * (nblocks + 1) > buf.size:
        raise ValueError("container index does not fit")
    offsets = np.frombuffer(buf[HEADER:HEADER + 8 * (nblocks + 1)].tobytes(), dtype=np.uint64)
    base = HEADER + 8 * (nblocks + 1)
    if offsets[0] != 0 or np.any(np.diff(offsets.astype(np.int64)) < 0) or base + int(offsets[-1]) > buf.size:
        raise ValueError("container offsets are not monotone / in range")
    restart = None
    if flags:
        at = base + ((int(offsets[-1]) + 3) & ~3)
        per = 3 if mode != 1 else (ADAPTIVE_RESTART_WORDS if block <= 65536 else ADAPTIVE_RESTART_WORDS_WIDE)
        words = nblocks * restart_records(block, seg_syms) * per
        if at + 4 * words > buf.size:
            raise ValueError("restart table does not fit")
        restart = np.frombuffer(buf[at:at + 4 * words].tobytes(), dtype=np.uint32).reshape(nblocks, -1, per)
    return Info(mode, block, total, nblocks, offsets, base, seg_syms if flags else 0, restart)


def build(mode: int, block: int, total: int, payloads) -> np.ndarray:
    """Container from a list of payload byte strings (one per block)."""
    nb = len(payloads)
    if nb != nblocks_of(total, block):
        raise ValueError("payload count does not match total / block")
    sizes = np.array([len(p) for p in payloads], dtype=np.uint64)
    offsets = np.zeros(nb + 1, dtype=np.uint64)
    if nb:
        offsets[1:] = np.cumsum(sizes)
    out = np.empty(HEADER + 8 * (nb + 1) + int(offsets[-1]), dtype=np.uint8)
    out[:HEADER] = np.frombuffer(pack_header(mode, block, total, nb), dtype=np.uint8)
    out[HEADER:HEADER + 8 * (nb + 1)] = offsets.view(np.uint8)
    at = HEADER + 8 * (nb + 1)
    for p in payloads:
        out[at:at + len(p)] = np.frombuffer(bytes(p), dtype=np.uint8)
        at += len(p)
    return out


def shard_range(nblocks: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block range of `rank`: [floor(r*nb/W), floor((r+1)*nb/W))."""
    return (rank * nblocks) // world, ((rank + 1) * nblocks) // world


def merge_sizes(per_rank_sizes) -> np.ndarray:
    """Global offsets (uint64, nblocks+1) from each rank's per-block payload sizes, in rank order."""
    sizes = np.concatenate([np.asarray(s, dtype=np.uint64) for s in per_rank_sizes]) if per_rank_sizes else \
        np.zeros(0, dtype=np.uint64)
    offsets = np.zeros(sizes.size + 1, dtype=np.uint64)
    if sizes.size:
        offsets[1:] = np.cumsum(sizes)
    return offsets
