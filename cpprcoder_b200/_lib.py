"""ctypes binding of libb2rc.so (include/b2rc.h).  There is no fallback: a missing
library is an error, and every coding call needs a CUDA device."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# B2RC_LIB: another build of the same library (kernel experiments, tools/ only); default the in-tree one
SO = Path(os.environ["B2RC_LIB"]) if os.environ.get("B2RC_LIB") else PKG / "libb2rc.so"

OK, E_ARG, E_DST_SMALL, E_CORRUPT, E_CUDA, E_EXPAND, E_NOMEM, E_INTERNAL = 0, -1, -2, -3, -4, -5, -6, -7
MODE_STATIC, MODE_ADAPTIVE = 0, 1
MODE_RANS_BYTE, MODE_RANS_WORD = 2, 3  # cppans::rANS::encode / ::encode_simd
MODE_NAMES = {"static": 0, "adaptive": 1, "rans": 2, "rans-word": 3}
HEADER_BYTES = 32
DEFAULT_BLOCK = 65536
BLK_BLOCK, BLK_CODED = 32768, 32770  # blksort::BlkSort::BlockSize / ::EncodedSize

_P = C.c_void_p
_U64 = C.c_uint64
_U32 = C.c_uint32

# name -> (restype, argtypes); mirrors include/b2rc.h one to one
SIGNATURES = {
    "b2rc_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "b2rc_ctx_destroy": (None, [_P]),
    "b2rc_ctx_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "b2rc_ctx_devices": (C.c_int, [_P]),
    "b2rc_restart_for": (_U32, [_P, C.c_int, _U32, _U64]),
    "b2rc_ctx_force_restart": (C.c_int, [_P, _U32]),
    "b2rc_strerror": (C.c_char_p, [C.c_int]),
    "b2rc_last_cuda_error": (C.c_char_p, [_P]),
    "b2rc_bound": (_U64, [C.c_int, _U64, _U32]),
    "b2rc_slot_bytes": (_U64, [_U32]),
    "b2rc_slot_bytes_for": (_U64, [C.c_int, _U32]),
    "b2rc_nblocks": (_U64, [_U64, _U32]),
    "b2rc_encode": (C.c_int, [_P, C.c_int, _U32, _P, _U64, _P, _U64, C.POINTER(_U64)]),
    "b2rc_decode": (C.c_int, [_P, _P, _U64, _P, _U64, C.POINTER(_U64)]),
    "b2rc_peek": (C.c_int, [_P, _U64, C.POINTER(C.c_int), C.POINTER(_U32), C.POINTER(_U64), C.POINTER(_U64)]),
    "b2rc_check": (C.c_int, [_P, _U64, C.POINTER(_U64)]),
    "b2rc_container_bytes": (C.c_int, [_P, _U64, C.POINTER(_U64)]),
    "b2rc_encode_staged": (C.c_int, [_P, C.c_int, _U32, _P, _U64, C.POINTER(_P), C.POINTER(_U64)]),
    "b2rc_decode_staged": (C.c_int, [_P, _P, _U64, C.POINTER(_P), C.POINTER(_U64)]),
    "b2rc_host_alloc": (C.c_int, [_U64, C.POINTER(_P)]),
    "b2rc_host_free": (None, [_P]),
    "b2rc_host_copy": (None, [_P, _P, _U64]),
    "b2rc_encode_device": (C.c_int, [_P, C.c_int, _U32, _P, _U64, _P, _U64, C.POINTER(_U64), _P]),
    "b2rc_decode_device": (C.c_int, [_P, _P, _U64, _P, _U64, C.POINTER(_U64), _P]),
    "b2rc_k_histogram": (C.c_int, [_P, _P, _U64, _U32, _P, _P]),
    "b2rc_k_encode_blocks": (C.c_int, [_P, C.c_int, _U32, _P, _U64, _P, _P, _U64, _P, _P, _P]),
    "b2rc_restart_records": (_U32, [_U32, _U32]),
    "b2rc_k_encode_blocks_r": (C.c_int, [_P, C.c_int, _U32, _P, _U64, _P, _P, _U64, _P, _P, _U32, _P, _P]),
    "b2rc_k_decode_blocks_r": (C.c_int, [_P, C.c_int, _U32, _P, _U64, _P, _U64, _P, _U64, _P, _U32, _P, _P]),
    "b2rc_k_scan": (C.c_int, [_P, _P, _U64, _P, _P]),
    "b2rc_k_compact": (C.c_int, [_P, _P, _U64, _P, _P, _U64, _P, _U64, _P, _P]),
    "b2rc_k_compact_for": (C.c_int, [_P, C.c_int, _P, _U64, _P, _P, _U64, _P, _U64, _P, _P]),
    "b2rc_k_decode_blocks": (C.c_int, [_P, C.c_int, _U32, _P, _U64, _P, _U64, _P, _U64, _P, _P]),
    "b2rc_launch_count": (_U64, [_P]),
    "b2rc_profile": (C.c_int, [_P, C.c_int]),
    "b2rc_kernel_ms": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
    "b2rc_blk_encode_bound": (_U64, [_U64]),
    "b2rc_blk_decoded_size": (_U64, [_U64]),
    "b2rc_blk_encode_device": (C.c_int, [_P, _P, _U64, _P, _U64, C.POINTER(_U64), _P]),
    "b2rc_blk_decode_device": (C.c_int, [_P, _P, _U64, _P, _U64, C.POINTER(_U64), _P]),
    "b2rc_blk_encode": (C.c_int, [_P, _P, _U64, _P, _U64, C.POINTER(_U64)]),
    "b2rc_blk_decode": (C.c_int, [_P, _P, _U64, _P, _U64, C.POINTER(_U64)]),
    "b2rc_blk_rounds": (C.c_int, [_P, _P, _U64, C.POINTER(_U64)]),
    "b2rc_blkrc_bound": (_U64, [C.c_int, _U64, _U32]),
    "b2rc_blkrc_encode_device": (C.c_int, [_P, C.c_int, _U32, _P, _U64, _P, _U64, C.POINTER(_U64), _P]),
    "b2rc_blkrc_decode_device": (C.c_int, [_P, _P, _U64, _P, _U64, C.POINTER(_U64), _P]),
    "b2rc_build_arch": (C.c_char_p, []),
}

_lib = None


class B2rcError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        msg = f"{what}: {strerror(code)} ({code})"
        if detail:
            msg += f" [{detail}]"
        super().__init__(msg)


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not SO.exists():
            raise RuntimeError(f"{SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for the coder)")
        lib = C.CDLL(str(SO))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def strerror(code: int) -> str:
    return load().b2rc_strerror(code).decode()
