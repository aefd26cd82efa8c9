"""The C-ABI library must load on a CPU-only box and export every symbol include/b2rc.h
declares; calls that would code bytes must fail loudly without a device (no fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from cpprcoder_b200 import _lib, build
    build.build_native()
    return _lib.load()


def declared_functions():
    text = (ROOT / "include" / "b2rc.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2rc_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from cpprcoder_b200 import _lib
    names = declared_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b2rc.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of step with include/b2rc.h"


def test_size_helpers_need_no_device(lib):
    assert lib.b2rc_build_arch() == b"sm_100a"
    assert lib.b2rc_nblocks(0, 65536) == 0
    assert lib.b2rc_nblocks(65537, 65536) == 2
    assert lib.b2rc_slot_bytes(65536) % 128 == 0 and lib.b2rc_slot_bytes(65536) >= 65536 + 8192 + 1024
    assert lib.b2rc_bound(0, 0, 65536) == 40
    assert lib.b2rc_bound(0, 100, 63) == 0  # bad block size
    assert lib.b2rc_strerror(-3) == b"corrupt container or payload"


def test_block_sort_size_helpers_need_no_device(lib):
    # BlkSort::encodeBound (blksort.h:404-409): two bytes per full 32 KiB block; the exact inverse of it for decode
    for n, coded in [(0, 0), (1, 1), (32767, 32767), (32768, 32770), (32769, 32771), (152089, 152097), (5 << 30, (5 << 30) + 2 * 163840)]:
        assert lib.b2rc_blk_encode_bound(n) == coded
        assert lib.b2rc_blk_decoded_size(coded) == n
    # without a context nothing is coded: null context is an argument error, never a CPU fallback
    out = C.c_uint64(0)
    buf = (C.c_uint8 * 64)()
    assert lib.b2rc_blk_encode(None, buf, 64, buf, 64, C.byref(out)) == -1
    assert lib.b2rc_blk_decode_device(None, buf, 64, buf, 64, C.byref(out), None) == -1


def test_peek_validates_headers(lib):
    from cpprcoder_b200 import container
    good = container.build(1, 65536, 0, [])
    buf = (C.c_uint8 * len(good)).from_buffer_copy(good.tobytes())
    mode, block, total, nb = C.c_int(), C.c_uint32(), C.c_uint64(), C.c_uint64()
    assert lib.b2rc_peek(buf, len(good), C.byref(mode), C.byref(block), C.byref(total), C.byref(nb)) == 0
    assert (mode.value, block.value, total.value, nb.value) == (1, 65536, 0, 0)
    bad = bytearray(good.tobytes())
    bad[0] ^= 1
    buf = (C.c_uint8 * len(bad)).from_buffer_copy(bytes(bad))
    assert lib.b2rc_peek(buf, len(bad), None, None, None, None) == -3
    assert lib.b2rc_peek(buf, 10, None, None, None, None) == -3
    # nblocks that does not match total / block
    lying = bytearray(container.pack_header(0, 65536, 1 << 20, 3)) + bytes(8 * 4)
    buf = (C.c_uint8 * len(lying)).from_buffer_copy(bytes(lying))
    assert lib.b2rc_peek(buf, len(lying), None, None, None, None) == -3


def test_no_device_means_error_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    h = C.c_void_p()
    assert lib.b2rc_ctx_create(0, C.byref(h)) == -4  # B2RC_E_CUDA
    assert not h.value
    from cpprcoder_b200 import api
    with pytest.raises(RuntimeError):
        api.Context()


def test_restart_helpers_and_flags(lib):
    from cpprcoder_b200 import container
    # points per block: every seg_syms-th symbol after the first; seg must be a multiple of 64 below the block size
    assert lib.b2rc_restart_records(65536, 8192) == 7
    assert lib.b2rc_restart_records(1 << 20, 8192) == 127
    assert lib.b2rc_restart_records(65536, 65536) == 0 and lib.b2rc_restart_records(65536, 100) == 0
    assert lib.b2rc_restart_records(8192, 8192) == 0
    # the bound leaves room for the densest table a context may be set to write (1024 symbols)
    nb = 16
    plain = 32 + 8 * (nb + 1) + nb * lib.b2rc_slot_bytes(65536)
    assert lib.b2rc_bound(0, nb * 65536, 65536) >= plain + 3 + nb * 63 * 12
    # the adaptive coder's points carry the model: 131 words each, at most every 4096 symbols
    assert lib.b2rc_bound(1, nb * 65536, 65536) >= plain + 3 + nb * 15 * 524
    assert lib.b2rc_bound(3, nb * 65536, 65536) == 32 + 8 * (nb + 1) + nb * lib.b2rc_slot_bytes_for(3, 65536)  # rANS word: none
    # peek accepts the flag for the coders that have restart points, with a legal segment length
    def peek(header):
        raw = bytes(header) + bytes(8 * 17)
        buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        return lib.b2rc_peek(buf, len(raw), None, None, None, None)
    assert peek(container.pack_header(0, 65536, nb * 65536, nb, 8192)) == 0
    assert peek(container.pack_header(1, 65536, nb * 65536, nb, 8192)) == 0
    assert peek(container.pack_header(3, 65536, nb * 65536, nb, 8192)) == -3
    bad = bytearray(container.pack_header(0, 65536, nb * 65536, nb, 8192))
    bad[12] = 3                                   # unknown flag bit
    assert peek(bad) == -3
    bad = bytearray(container.pack_header(0, 65536, nb * 65536, nb))
    bad[13], bad[12] = 0x10, 1                    # segment length 65536 * ... not below the block size
    bad[14] = 0x10
    assert peek(bad) == -3
