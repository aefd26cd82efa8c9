"""Python face of the C ABI: torch owns device memory and streams, libb2rc.so does the work.

Every function here ends in a kernel launch inside libb2rc.so; nothing is coded in
Python or on the CPU.  Names follow include/b2rc.h.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import B2rcError, MODE_ADAPTIVE, MODE_NAMES, MODE_RANS_BYTE, MODE_RANS_WORD, MODE_STATIC  # noqa: F401

DEFAULT_BLOCK = _lib.DEFAULT_BLOCK


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def bound(mode: int, n: int, block: int = DEFAULT_BLOCK) -> int:
    return int(_lib.load().b2rc_bound(mode, n, block))


def slot_bytes(n: int, mode: int = MODE_STATIC) -> int:
    return int(_lib.load().b2rc_slot_bytes_for(mode, n))


def nblocks(n: int, block: int) -> int:
    return int(_lib.load().b2rc_nblocks(n, block))


def blk_encode_bound(n: int) -> int:
    """BlkSort::encodeBound (blksort.h:404-409)."""
    return int(_lib.load().b2rc_blk_encode_bound(n))


def blk_decoded_size(n: int) -> int:
    return int(_lib.load().b2rc_blk_decoded_size(n))


class Context:
    """b2rc_ctx: one per host thread and device."""

    def __init__(self, device: int | None = None, devices: list[int] | None = None):
        """`devices`: b2rc_ctx_create_multi -- the host-pointer calls (encode / decode) shard the blocks over
        these devices of this process; the device-pointer calls run on the first."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("cpprcoder_b200 needs a CUDA device: there is no CPU coding path")
        h = C.c_void_p()
        if devices:
            self.device = int(devices[0])
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.lib.b2rc_ctx_create_multi(arr, len(devices), C.byref(h))
        else:
            self.device = torch.cuda.current_device() if device is None else int(device)
            rc = self.lib.b2rc_ctx_create(self.device, C.byref(h))
        if rc != _lib.OK:
            raise B2rcError(rc, "b2rc_ctx_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2rc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != _lib.OK:
            raise B2rcError(rc, what, self.lib.b2rc_last_cuda_error(self.h).decode())

    @property
    def launches(self) -> int:
        return int(self.lib.b2rc_launch_count(self.h))

    KERNELS = {"histogram": 0, "encode": 1, "scan": 2, "compact": 3, "decode": 4, "blk_forward": 5, "blk_inverse": 6,
               "ranges": 7, "seams": 8}

    @staticmethod
    def supported_modes() -> tuple:
        """Coders this build carries kernels for, in container mode numbers."""
        return (MODE_STATIC, MODE_ADAPTIVE, MODE_RANS_BYTE, MODE_RANS_WORD)

    def restart_for(self, mode: int, block: int, nblocks: int) -> int:
        """Spacing of the restart points this context writes a stream of `nblocks` blocks with (0: none)."""
        return int(self.lib.b2rc_restart_for(self.h, mode, block, nblocks))

    def force_restart(self, seg_syms: int):
        """One spacing for every stream from now on (0: automatic again) -- for containers that get stitched."""
        self._check(self.lib.b2rc_ctx_force_restart(self.h, seg_syms), "b2rc_ctx_force_restart")

    def profile(self, enable: bool = True):
        self._check(self.lib.b2rc_profile(self.h, 1 if enable else 0), "b2rc_profile")

    def kernel_ms(self) -> dict:
        """Duration of the last launch of each kernel since profile(True); missing ones are omitted."""
        out = {}
        for name, k in self.KERNELS.items():
            ms = C.c_float(-1.0)
            self._check(self.lib.b2rc_kernel_ms(self.h, k, C.byref(ms)), "b2rc_kernel_ms")
            if ms.value >= 0:
                out[name] = float(ms.value)
        return out

    # ---- whole container, device memory ------------------------------------
    def encode_device(self, mode: int, src: torch.Tensor, dst: torch.Tensor | None = None,
                      block: int = DEFAULT_BLOCK) -> tuple[torch.Tensor, int]:
        """src: uint8 CUDA tensor.  Returns (dst, bytes used)."""
        n = src.numel()
        if dst is None:
            dst = torch.empty(bound(mode, n, block), dtype=torch.uint8, device=src.device)
        out_n = C.c_uint64(0)
        rc = self.lib.b2rc_encode_device(self.h, mode, block, _ptr(src), n, _ptr(dst), dst.numel(), C.byref(out_n),
                                         _stream())
        self._check(rc, "b2rc_encode_device")
        return dst, int(out_n.value)

    def decode_device(self, src: torch.Tensor, n_src: int, dst: torch.Tensor) -> int:
        out_n = C.c_uint64(0)
        rc = self.lib.b2rc_decode_device(self.h, _ptr(src), n_src, _ptr(dst), dst.numel(), C.byref(out_n), _stream())
        self._check(rc, "b2rc_decode_device")
        return int(out_n.value)

    # ---- whole container, host memory (the call the C++ drop-in classes make) ---
    def encode(self, mode: int, src, block: int = DEFAULT_BLOCK, dst: np.ndarray | None = None) -> np.ndarray:
        src = np.ascontiguousarray(np.frombuffer(src, dtype=np.uint8) if isinstance(src, (bytes, bytearray)) else src,
                                   dtype=np.uint8)
        if dst is None:
            dst = np.empty(bound(mode, src.size, block), dtype=np.uint8)
        out_n = C.c_uint64(0)
        rc = self.lib.b2rc_encode(self.h, mode, block, src.ctypes.data_as(C.c_void_p), src.size,
                                  dst.ctypes.data_as(C.c_void_p), dst.size, C.byref(out_n))
        self._check(rc, "b2rc_encode")
        return dst[:int(out_n.value)]

    def decode(self, src, dst: np.ndarray | None = None) -> np.ndarray:
        src = np.ascontiguousarray(np.frombuffer(src, dtype=np.uint8) if isinstance(src, (bytes, bytearray)) else src,
                                   dtype=np.uint8)
        mode, block, total, nb = self.peek(src)
        if dst is None:
            dst = np.empty(max(total, 1), dtype=np.uint8)
        out_n = C.c_uint64(0)
        rc = self.lib.b2rc_decode(self.h, src.ctypes.data_as(C.c_void_p), src.size, dst.ctypes.data_as(C.c_void_p),
                                  dst.size, C.byref(out_n))
        self._check(rc, "b2rc_decode")
        return dst[:int(out_n.value)]

    def peek(self, src: np.ndarray):
        mode, block, total, nb = C.c_int(0), C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.b2rc_peek(src.ctypes.data_as(C.c_void_p), src.size, C.byref(mode), C.byref(block),
                                C.byref(total), C.byref(nb))
        self._check(rc, "b2rc_peek")
        return mode.value, block.value, total.value, nb.value

    # ---- per-kernel doors -----------------------------------------------------
    def histogram(self, src: torch.Tensor, block: int, freq16: torch.Tensor | None = None) -> torch.Tensor:
        nb = nblocks(src.numel(), block)
        if freq16 is None:
            freq16 = torch.empty((max(nb, 1), 256), dtype=torch.int16, device=src.device)
        self._check(self.lib.b2rc_k_histogram(self.h, _ptr(src), src.numel(), block, _ptr(freq16), _stream()),
                    "b2rc_k_histogram")
        return freq16

    def encode_blocks(self, mode: int, src: torch.Tensor, block: int = DEFAULT_BLOCK, freq16=None, slots=None,
                      sizes=None, err=None, restart=None, seg_syms: int = 0):
        """K1 (when needed) + K2.  Returns (slots, slot_stride, sizes, err).  `restart` (int32 tensor,
        nblocks * restart_records(block, seg_syms) * 3) receives the static coder's restart points."""
        n = src.numel()
        nb = nblocks(n, block)
        stride = slot_bytes(block, mode)
        dev = src.device
        if slots is None:
            slots = torch.empty(max(nb, 1) * stride, dtype=torch.uint8, device=dev)
        if sizes is None:
            sizes = torch.zeros(max(nb, 1), dtype=torch.int32, device=dev)
        if err is None:
            err = torch.zeros(4, dtype=torch.int32, device=dev)
        if mode == MODE_STATIC and freq16 is None:
            freq16 = self.histogram(src, block)
        self._check(self.lib.b2rc_k_encode_blocks_r(self.h, mode, block, _ptr(src), n, _ptr(freq16), _ptr(slots), stride,
                                                    _ptr(sizes), _ptr(restart), seg_syms, _ptr(err), _stream()),
                    "b2rc_k_encode_blocks_r")
        return slots, stride, sizes, err

    def scan(self, sizes: torch.Tensor, nb: int, offsets: torch.Tensor | None = None) -> torch.Tensor:
        if offsets is None:
            offsets = torch.empty(nb + 1, dtype=torch.int64, device=sizes.device)
        self._check(self.lib.b2rc_k_scan(self.h, _ptr(sizes), nb, _ptr(offsets), _stream()), "b2rc_k_scan")
        return offsets

    def compact(self, slots, stride, sizes, offsets, nb, payload: torch.Tensor, err: torch.Tensor,
                mode: int = MODE_STATIC):
        self._check(self.lib.b2rc_k_compact_for(self.h, mode, _ptr(slots), stride, _ptr(sizes), _ptr(offsets), nb,
                                                _ptr(payload), payload.numel(), _ptr(err), _stream()),
                    "b2rc_k_compact_for")

    def decode_blocks(self, mode: int, payload: torch.Tensor, payload_len: int, offsets: torch.Tensor, nb: int,
                      dst: torch.Tensor, n: int, block: int = DEFAULT_BLOCK, err: torch.Tensor | None = None,
                      restart=None, seg_syms: int = 0):
        if err is None:
            err = torch.zeros(4, dtype=torch.int32, device=dst.device)
        self._check(self.lib.b2rc_k_decode_blocks_r(self.h, mode, block, _ptr(payload), payload_len, _ptr(offsets), nb,
                                                    _ptr(dst), n, _ptr(restart), seg_syms, _ptr(err), _stream()),
                    "b2rc_k_decode_blocks_r")
        return err

    @staticmethod
    def restart_records(block: int, seg_syms: int) -> int:
        return int(_lib.load().b2rc_restart_records(block, seg_syms))

    # ---- block sort (blksort::BlkSort) ----------------
    def blk_encode_device(self, src: torch.Tensor, dst: torch.Tensor | None = None) -> torch.Tensor:
        """BlkSort::encode on device memory: every full 32 KiB block -> column + row number."""
        n = src.numel()
        need = blk_encode_bound(n)
        if dst is None:
            dst = torch.empty(max(need, 16), dtype=torch.uint8, device=src.device)
        out_n = C.c_uint64(0)
        self._check(self.lib.b2rc_blk_encode_device(self.h, _ptr(src), n, _ptr(dst), dst.numel(), C.byref(out_n),
                                                    _stream()), "b2rc_blk_encode_device")
        return dst[:int(out_n.value)]

    def blk_decode_device(self, src: torch.Tensor, n_src: int | None = None, dst: torch.Tensor | None = None) -> torch.Tensor:
        n = src.numel() if n_src is None else n_src
        need = blk_decoded_size(n)
        if dst is None:
            dst = torch.empty(max(need, 16), dtype=torch.uint8, device=src.device)
        out_n = C.c_uint64(0)
        self._check(self.lib.b2rc_blk_decode_device(self.h, _ptr(src), n, _ptr(dst), dst.numel(), C.byref(out_n),
                                                    _stream()), "b2rc_blk_decode_device")
        return dst[:int(out_n.value)]

    def blkrc_encode_device(self, mode: int, src: torch.Tensor, dst: torch.Tensor | None = None,
                            block: int = DEFAULT_BLOCK) -> tuple[torch.Tensor, int]:
        """Block sort, then the coder, one call on the device (the reference's run_zlib_blk shape, test/main.cpp:944-1002)."""
        n = src.numel()
        if dst is None:
            dst = torch.empty(int(self.lib.b2rc_blkrc_bound(mode, n, block)), dtype=torch.uint8, device=src.device)
        out_n = C.c_uint64(0)
        self._check(self.lib.b2rc_blkrc_encode_device(self.h, mode, block, _ptr(src), n, _ptr(dst), dst.numel(), C.byref(out_n),
                                                      _stream()), "b2rc_blkrc_encode_device")
        return dst, int(out_n.value)

    def blkrc_decode_device(self, src: torch.Tensor, n_src: int, dst: torch.Tensor) -> int:
        out_n = C.c_uint64(0)
        self._check(self.lib.b2rc_blkrc_decode_device(self.h, _ptr(src), n_src, _ptr(dst), dst.numel(), C.byref(out_n), _stream()),
                    "b2rc_blkrc_decode_device")
        return int(out_n.value)

    def _blk_host(self, fn, what: str, src, need: int, dst):
        src = np.ascontiguousarray(np.frombuffer(src, dtype=np.uint8) if isinstance(src, (bytes, bytearray)) else src,
                                   dtype=np.uint8)
        if dst is None:
            dst = np.empty(max(need(src.size), 1), dtype=np.uint8)
        out_n = C.c_uint64(0)
        self._check(fn(self.h, src.ctypes.data_as(C.c_void_p), src.size, dst.ctypes.data_as(C.c_void_p), dst.size,
                       C.byref(out_n)), what)
        return dst[:int(out_n.value)]

    def blk_encode(self, src, dst: np.ndarray | None = None) -> np.ndarray:
        return self._blk_host(self.lib.b2rc_blk_encode, "b2rc_blk_encode", src, blk_encode_bound, dst)

    def blk_decode(self, src, dst: np.ndarray | None = None) -> np.ndarray:
        return self._blk_host(self.lib.b2rc_blk_decode, "b2rc_blk_decode", src, blk_decoded_size, dst)

    def blk_rounds(self) -> np.ndarray:
        """Doubling rounds per block of the last forward call (bit 31: the block has a period)."""
        nb = C.c_uint64(0)
        self._check(self.lib.b2rc_blk_rounds(self.h, None, 0, C.byref(nb)), "b2rc_blk_rounds")
        out = np.zeros(int(nb.value), dtype=np.uint32)
        if out.size:
            self._check(self.lib.b2rc_blk_rounds(self.h, out.ctypes.data_as(C.c_void_p), out.size, C.byref(nb)),
                        "b2rc_blk_rounds")
        return out
