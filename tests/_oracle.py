"""ctypes doors onto the parity checkers (TEST INFRASTRUCTURE).

``Oracle``  -> oracle/liboracle.so, the plain-C restatement (oracle/rc_oracle.c).
``Ref``     -> oracle/_ref/libcpprcoder_ref.so, the unmodified reference header
               behind extern "C" (oracle/ref_shim.cpp); present wherever
               ``make -C oracle`` ran with /root/reference mounted.

Both expose the same per-block calls so tests can swap one for the other.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tarfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
GOLDEN = ROOT / "tests" / "golden"

STATIC, ADAPTIVE = 0, 1
RANS_BYTE, RANS_WORD = 2, 3  # cppans::rANS::encode / encode_simd (oracle/ans_oracle.h)
RANS_HEADER = 1032
FNV_OFFSET = 1469598103934665603
FNV_PRIME = 1099511628211


def slot_bytes(n: int, mode: int = STATIC) -> int:
    if mode in (RANS_BYTE, RANS_WORD):
        return (RANS_HEADER + 32 + 2 * n + 15) & ~15
    s = n + n // 8 + 1024
    return (s + 127) & ~127


def fnv1a64(data: bytes | np.ndarray) -> int:
    lib = Oracle.get().lib
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    return int(lib.rco_fnv1a64(buf.ctypes.data_as(C.c_void_p), buf.size, 0))


class _Stats(C.Structure):
    _fields_ = [("carries", C.c_uint32), ("carries_with_run", C.c_uint32), ("max_pending_run", C.c_uint32),
                ("final_low", C.c_uint32), ("rescales", C.c_uint32)]


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


class _Coder:
    """Shared blocked helpers; subclasses bind the four C entry points."""

    name = "?"

    def encode(self, mode: int, src) -> bytes:
        raise NotImplementedError

    def decode(self, mode: int, payload, cap: int) -> bytes:
        raise NotImplementedError

    def encode_blocks(self, mode: int, src, block: int, threads: int = 1):
        """-> (list of payload bytes) for each block of ``src``."""
        src = _u8(src)
        n = src.size
        nblocks = (n + block - 1) // block
        stride = slot_bytes(block, mode)
        slots = np.empty(max(nblocks, 1) * stride, dtype=np.uint8)
        sizes = np.zeros(max(nblocks, 1), dtype=np.uint32)
        rc = self._encode_blocks(mode, src.ctypes.data_as(C.c_void_p), n, block, slots.ctypes.data_as(C.c_void_p),
                                 stride, sizes.ctypes.data_as(C.c_void_p), threads)
        if rc != 0:
            raise RuntimeError(f"{self.name}: encode_blocks failed")
        return [slots[b * stride: b * stride + int(sizes[b])].tobytes() for b in range(nblocks)]

    def decode_blocks(self, mode: int, stream, offsets, block: int, n: int, threads: int = 1) -> np.ndarray:
        stream = _u8(stream)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nblocks = offsets.size - 1
        dst = np.empty(max(n, 1), dtype=np.uint8)
        rc = self._decode_blocks(mode, stream.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), nblocks,
                                 block, dst.ctypes.data_as(C.c_void_p), n, threads)
        if rc != 0:
            raise RuntimeError(f"{self.name}: decode_blocks failed")
        return dst[:n]


class Oracle(_Coder):
    name = "oracle"
    _inst = None

    @classmethod
    def get(cls) -> "Oracle":
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        so = ORACLE_DIR / "liboracle.so"
        if not so.exists():
            subprocess.check_call(["make", "-C", str(ORACLE_DIR), "liboracle.so"], stdout=subprocess.DEVNULL)
        self.lib = lib = C.CDLL(str(so))
        lib.rco_fnv1a64.restype = C.c_uint64
        lib.rco_fnv1a64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        for fn in (lib.rco_static_encode, lib.rco_adaptive_encode):
            fn.restype = C.c_long
            fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
        for fn in (lib.rco_static_decode, lib.rco_adaptive_decode):
            fn.restype = C.c_long
            fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.rco_static_count.restype = None
        lib.rco_static_count.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.rco_encode_blocks.restype = C.c_int
        lib.rco_encode_blocks.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64,
                                          C.c_void_p, C.c_int]
        lib.rco_decode_blocks.restype = C.c_int
        lib.rco_decode_blocks.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p,
                                          C.c_uint64, C.c_int]
        self._encode_blocks = lib.rco_encode_blocks
        self._decode_blocks = lib.rco_decode_blocks
        lib.rao_encode.restype = C.c_long
        lib.rao_encode.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]
        lib.rao_decode.restype = C.c_long
        lib.rao_decode.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.rao_model.restype = None
        lib.rao_model.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]

    def encode(self, mode: int, src, with_stats: bool = False):
        src = _u8(src)
        cap = slot_bytes(src.size) + 2 * src.size
        dst = np.empty(cap, dtype=np.uint8)
        if mode in (RANS_BYTE, RANS_WORD):
            r = self.lib.rao_encode(mode, src.ctypes.data_as(C.c_void_p), src.size, dst.ctypes.data_as(C.c_void_p), cap)
            if r < 0:
                raise RuntimeError("oracle rANS encode failed")
            return dst[:r].tobytes()
        st = _Stats()
        fn = self.lib.rco_static_encode if mode == STATIC else self.lib.rco_adaptive_encode
        r = fn(src.ctypes.data_as(C.c_void_p), src.size, dst.ctypes.data_as(C.c_void_p), cap, C.byref(st))
        if r < 0:
            raise RuntimeError("oracle encode failed")
        out = dst[:r].tobytes()
        if with_stats:
            return out, {k: getattr(st, k) for k, _ in _Stats._fields_}
        return out

    def decode(self, mode: int, payload, cap: int) -> bytes:
        payload = _u8(payload)
        dst = np.empty(max(cap, 1), dtype=np.uint8)
        if mode in (RANS_BYTE, RANS_WORD):
            r = self.lib.rao_decode(mode, payload.ctypes.data_as(C.c_void_p), payload.size,
                                    dst.ctypes.data_as(C.c_void_p), cap)
        else:
            fn = self.lib.rco_static_decode if mode == STATIC else self.lib.rco_adaptive_decode
            r = fn(payload.ctypes.data_as(C.c_void_p), payload.size, dst.ctypes.data_as(C.c_void_p), cap)
        if r < 0:
            raise RuntimeError("oracle decode failed")
        return dst[:r].tobytes()

    def rans_model(self, src, bits: int) -> tuple[np.ndarray, np.ndarray]:
        """(freq[256], cum[257]) after the reference's normalize (cppans.h:138-177)."""
        src = _u8(src)
        freq = np.zeros(256, dtype=np.uint32)
        cum = np.zeros(257, dtype=np.uint32)
        self.lib.rao_model(src.ctypes.data_as(C.c_void_p), src.size, 1 << bits, freq.ctypes.data_as(C.c_void_p),
                           cum.ctypes.data_as(C.c_void_p))
        return freq, cum

    def static_count(self, src) -> tuple[np.ndarray, int]:
        src = _u8(src)
        freq = np.zeros(256, dtype=np.uint32)
        ev = C.c_uint32(0)
        self.lib.rco_static_count(src.ctypes.data_as(C.c_void_p), src.size, freq.ctypes.data_as(C.c_void_p), C.byref(ev))
        return freq, int(ev.value)


class Ref(_Coder):
    """The unmodified reference (oracle/_ref). ``Ref.available()`` is False where it was never built."""

    name = "reference"
    _inst = None
    SO = ORACLE_DIR / "_ref" / "libcpprcoder_ref.so"

    @classmethod
    def available(cls) -> bool:
        return cls.SO.exists()

    @classmethod
    def get(cls) -> "Ref":
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        self.lib = lib = C.CDLL(str(self.SO))
        lib.ref_encode.restype = C.c_long
        lib.ref_encode.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]
        lib.ref_decode.restype = C.c_long
        lib.ref_decode.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.ref_encode_blocks.restype = C.c_int
        lib.ref_encode_blocks.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64,
                                          C.c_void_p, C.c_int]
        lib.ref_decode_blocks.restype = C.c_int
        lib.ref_decode_blocks.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p,
                                          C.c_uint64, C.c_int]
        lib.ref_hardware_threads.restype = C.c_int
        self._encode_blocks = lib.ref_encode_blocks
        self._decode_blocks = lib.ref_decode_blocks

    def encode(self, mode: int, src) -> bytes:
        src = _u8(src)
        cap = slot_bytes(src.size) + 2 * src.size
        dst = np.empty(cap, dtype=np.uint8)
        r = self.lib.ref_encode(mode, src.ctypes.data_as(C.c_void_p), src.size, dst.ctypes.data_as(C.c_void_p), cap)
        if r < 0:
            raise RuntimeError("reference encode failed")
        return dst[:r].tobytes()

    def decode(self, mode: int, payload, cap: int) -> bytes:
        payload = _u8(payload)
        dst = np.empty(max(cap, 1), dtype=np.uint8)
        r = self.lib.ref_decode(mode, payload.ctypes.data_as(C.c_void_p), payload.size, dst.ctypes.data_as(C.c_void_p),
                                cap)
        if r < 0:
            raise RuntimeError("reference decode failed")
        return dst[:r].tobytes()

    def hardware_threads(self) -> int:
        return int(self.lib.ref_hardware_threads())


# ----------------------------------------------------------------- block sort --
BLK_BLOCK, BLK_CODED = 32768, 32770  # blksort::BlkSort::BlockSize / ::EncodedSize (blksort.h:80-83)


def blk_encode_bound(n: int) -> int:
    return (n >> 15) * BLK_CODED + (n & (BLK_BLOCK - 1))


def blk_decoded_size(n: int) -> int:
    return (n // BLK_CODED) * BLK_BLOCK + n % BLK_CODED


class BlkSort:
    """blksort::BlkSort on the CPU: ``BlkSort(Oracle.get())`` is the C restatement (oracle/blk_oracle.c),
    ``BlkSort(Ref.get())`` the unmodified blksort.h (oracle/ref_shim.cpp, ref_blk_*)."""

    def __init__(self, side):
        self.name = side.name
        lib = side.lib
        self._enc, self._dec = (lib.bso_encode, lib.bso_decode) if isinstance(side, Oracle) else (lib.ref_blk_encode,
                                                                                                  lib.ref_blk_decode)
        for fn in (self._enc, self._dec):
            fn.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        self._enc.restype = None
        self._dec.restype = C.c_int if isinstance(side, Oracle) else None
        if isinstance(side, Oracle):
            lib.bso_block_is_periodic.restype = C.c_int
            lib.bso_block_is_periodic.argtypes = [C.c_void_p]
            self._periodic = lib.bso_block_is_periodic

    def encode(self, src, threads: int = 1) -> np.ndarray:
        src = _u8(src)
        dst = np.zeros(max(blk_encode_bound(src.size), 1), dtype=np.uint8)
        self._enc(src.size, dst.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), threads)
        return dst[:blk_encode_bound(src.size)]

    def decode(self, coded, threads: int = 1) -> np.ndarray:
        coded = _u8(coded)
        dst = np.zeros(max(blk_decoded_size(coded.size), 1), dtype=np.uint8)
        rc = self._dec(coded.size, dst.ctypes.data_as(C.c_void_p), coded.ctypes.data_as(C.c_void_p), threads)
        if rc not in (None, 0):
            raise RuntimeError(f"{self.name}: block-sort decode failed")
        return dst[:blk_decoded_size(coded.size)]

    def periodic_blocks(self, src) -> list:
        """Indices of the full blocks of ``src`` that have a period (some rotations are equal)."""
        src = _u8(src)
        return [b for b in range(src.size >> 15)
                if self._periodic(src[b * BLK_BLOCK:(b + 1) * BLK_BLOCK].ctypes.data_as(C.c_void_p))]


# --------------------------------------------------------------------- corpus --
CANTERBURY = ["alice29.txt", "asyoulik.txt", "cp.html", "fields.c", "grammar.lsp", "kennedy.xls", "lcet10.txt",
              "plrabn12.txt", "ptt5", "sum", "xargs.1"]
_corpus_cache: dict[str, bytes] = {}


def canterbury(name: str) -> bytes:
    """One file of the Canterbury corpus (tests/golden/cantrbry.tar.bz2, the corpus the
    reference ships as test/cantrbry.tar.bz2 and names in test/main.cpp:1246-1258)."""
    if not _corpus_cache:
        with tarfile.open(GOLDEN / "cantrbry.tar.bz2", "r:bz2") as tf:
            for m in tf.getmembers():
                if m.isfile():
                    _corpus_cache[os.path.basename(m.name)] = tf.extractfile(m).read()
    return _corpus_cache[name]


def offsets_of(payloads) -> np.ndarray:
    off = np.zeros(len(payloads) + 1, dtype=np.uint64)
    if payloads:
        off[1:] = np.cumsum([len(p) for p in payloads], dtype=np.uint64)
    return off
