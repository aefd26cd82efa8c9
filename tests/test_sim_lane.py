"""The lane arithmetic the kernels run (cpprcoder_b200/csrc/rc_lane.cuh), driven on the
CPU by tests/sim/sim_lane.cpp and checked against the oracle.  This is a test of the
device code's logic, not a CPU coding path: nothing in the product calls it."""
import ctypes as C

import numpy as np
import pytest

from _cases import crafted
from _oracle import ADAPTIVE, CANTERBURY, STATIC, Oracle, canterbury, slot_bytes


@pytest.fixture(scope="module")
def sim(built):
    lib = C.CDLL(str(built.build_sim()))
    lib.sim_encode.restype = C.c_long
    lib.sim_encode.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_int]
    lib.sim_decode.restype = C.c_long
    lib.sim_decode.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t]
    lib.sim_check_div.restype = C.c_uint64
    lib.sim_check_div.argtypes = [C.c_uint32] * 4
    return lib


def _enc(sim, mode, d, exact=0):
    d = np.ascontiguousarray(d, dtype=np.uint8)
    cap = slot_bytes(d.size) + 2 * d.size
    out = np.empty(cap, np.uint8)
    r = sim.sim_encode(mode, d.ctypes.data_as(C.c_void_p), d.size, out.ctypes.data_as(C.c_void_p), cap, exact)
    assert r >= 0
    return out[:r].tobytes()


def _dec(sim, mode, pay, n, lead):
    st = np.frombuffer(bytes(lead) + pay + bytes(3), dtype=np.uint8).copy()
    st[:lead] = 0xA5
    out = np.empty(max(n, 1), np.uint8)
    r = sim.sim_decode(mode, st.ctypes.data_as(C.c_void_p), len(st) - 3, lead, out.ctypes.data_as(C.c_void_p), n)
    assert r == n
    return out[:n].tobytes()


def test_magic_division_is_exact(sim):
    rng = np.random.default_rng(5)
    bad = 0
    for d in [1, 2, 3, 7, 255, 256, 257, 65535, 65536, 65537, 0x8000, 1000003, (1 << 24) - 1, 1 << 24]:
        bad += sim.sim_check_div(d, 0, 4099, 1 << 18)
        bad += sim.sim_check_div(d, 0xFFFC0000, 1, 0x3FFFF)
    for _ in range(200):
        bad += sim.sim_check_div(int(rng.integers(1, 1 << 24)), int(rng.integers(0, 1 << 32)),
                                 int(rng.integers(1, 1 << 16)), 5000)
    assert bad == 0


def test_range_chain_for_total_65536_is_exact(sim):
    """The three-candidate minimum of the range-pass kernels (k_enc_ranges, k_enc_ranges2): every t in [2^8, 2^16)
    for some 310 frequencies (the edges and random ones) -- 20 million links, all equal to the plain form."""
    sim.sim_check_range16.restype = C.c_uint64
    sim.sim_check_range16.argtypes = [C.c_uint32]
    rng = np.random.default_rng(16)
    fs = sorted({1, 2, 3, 255, 256, 257, 4095, 4096, 32767, 32768, 65534, 65535} | set(int(x) for x in rng.integers(1, 65536, 300)))
    assert sum(sim.sim_check_range16(f) for f in fs) == 0


def test_lane_coder_matches_oracle_on_crafted_blocks(sim):
    o = Oracle.get()
    rng = np.random.default_rng(7)
    for it in range(210):
        n = 65536 if it % 3 == 0 else int(rng.integers(1, 65537))
        d = crafted(it % 7, n, rng)
        for mode in (STATIC, ADAPTIVE):
            a = o.encode(mode, d)
            assert _enc(sim, mode, d) == a, (it, mode, n)
            if mode == STATIC:  # the reference-shaped path used for the low_==0xFFFFFFFF flush quirk
                assert _enc(sim, STATIC, d, 1) == a
            assert _dec(sim, mode, a, n, it % 4) == d.tobytes()


@pytest.mark.parametrize("name", ["alice29.txt", "kennedy.xls", "ptt5", "xargs.1"])
def test_lane_coder_on_canterbury_blocks(sim, name):
    o = Oracle.get()
    data = canterbury(name)
    for mode in (STATIC, ADAPTIVE):
        pays = o.encode_blocks(mode, data, 65536)
        for i, p in enumerate(pays):
            blk = data[i * 65536:(i + 1) * 65536]
            assert _enc(sim, mode, np.frombuffer(blk, np.uint8)) == p
            assert _dec(sim, mode, p, len(blk), (i + 1) % 4) == blk


# ---- byte-wise rANS lane arithmetic (cpprcoder_b200/csrc/ans_lane.cuh) --------------------
def _ans_fns(sim):
    sim.sim_ans_byte_encode.restype = C.c_long
    sim.sim_ans_byte_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t]
    sim.sim_ans_byte_decode.restype = C.c_long
    sim.sim_ans_byte_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_size_t]
    return sim


def _ans_enc(sim, o, d):
    d = np.ascontiguousarray(d, dtype=np.uint8)
    _, cum = o.rans_model(d, 14)
    cap = 2 * d.size + 2048
    out = np.empty(cap, np.uint8)
    r = sim.sim_ans_byte_encode(d.ctypes.data_as(C.c_void_p), d.size, cum.ctypes.data_as(C.c_void_p),
                                out.ctypes.data_as(C.c_void_p), cap)
    assert r >= 0
    return out[:r].tobytes()


def _ans_dec(sim, pay, n, lead):
    st = np.frombuffer(bytes(lead) + pay + bytes(11), dtype=np.uint8).copy()
    st[:lead] = 0xA5
    out = np.empty(max(n, 1), np.uint8)
    r = sim.sim_ans_byte_decode(st.ctypes.data_as(C.c_void_p), len(st) - 11, lead, out.ctypes.data_as(C.c_void_p), n)
    assert r == n, r
    return out[:n].tobytes()


def test_rans_byte_lane_matches_oracle(sim):
    from _oracle import RANS_BYTE
    _ans_fns(sim)
    o = Oracle.get()
    rng = np.random.default_rng(11)
    cases = [crafted(it % 7, 65536 if it % 3 == 0 else int(rng.integers(1, 65537)), rng) for it in range(60)]
    cases += [np.frombuffer(b, np.uint8) for b in (b"A", b"AB", b"A" * 65536, bytes(range(256)) * 5,
                                                   b"A" * 60000 + bytes(range(256)))]
    cases += [np.frombuffer(canterbury(n)[:65536], np.uint8) for n in ("alice29.txt", "kennedy.xls", "ptt5", "sum")]
    for it, d in enumerate(cases):
        want = o.encode(RANS_BYTE, d)
        assert _ans_enc(sim, o, d) == want, (it, d.size)
        assert _ans_dec(sim, want, d.size, it % 4) == d.tobytes(), (it, d.size)


# ---- restart points of the static coder (decode a block from several entry points) ----------
def test_static_restart_points(sim):
    sim.sim_encode_restart.restype = C.c_long
    sim.sim_encode_restart.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p]
    sim.sim_decode_from.restype = C.c_long
    sim.sim_decode_from.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.c_uint32, C.c_void_p]
    o = Oracle.get()
    rng = np.random.default_rng(21)
    for it in range(80):
        n = 65536 if it % 2 == 0 else int(rng.integers(300, 65537))
        d = np.ascontiguousarray(crafted(it % 7, n, rng))
        nseg = 4
        seg = 16384 if n == 65536 else max(64, (n // nseg) & ~63)
        cap = slot_bytes(n) + 2 * n
        out = np.empty(cap, np.uint8)
        rec = np.zeros(3 * (nseg - 1), np.uint32)
        r = sim.sim_encode_restart(d.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), cap, seg, nseg,
                                   rec.ctypes.data_as(C.c_void_p))
        if r == -2:
            continue
        assert r >= 0
        pay = out[:r].tobytes()
        assert pay == o.encode(STATIC, d), (it, n)        # capturing changes nothing in the payload
        lead = it % 4
        st = np.frombuffer(bytes(lead) + pay + bytes(8), dtype=np.uint8).copy()
        for j in range(1, nseg):
            k = j * seg
            m, low, rge = (int(v) for v in rec[3 * (j - 1):3 * j])
            if k >= n:
                assert m == 0xFFFFFFFF
                continue
            count = min(seg, n - k)
            got = np.empty(count, np.uint8)
            rr = sim.sim_decode_from(st.ctypes.data_as(C.c_void_p), len(st) - 8, lead, k, m, low, rge, count,
                                     got.ctypes.data_as(C.c_void_p))
            assert rr == count
            assert got.tobytes() == d[k:k + count].tobytes(), (it, n, j)


# ---- static encode as many chains per block (range-only pass, segments from low = 0, seams) ----
def test_segmented_static_encode_matches_oracle(sim):
    """The CPU model of k_enc_ranges / k_enc_seg / k_enc_seams: every payload equals the oracle's at
    every alignment, and the restart points equal those of the one-chain encoder."""
    sim.sim_encode_segmented.restype = C.c_long
    sim.sim_encode_segmented.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                         C.c_uint32, C.c_uint32, C.c_void_p]
    sim.sim_encode_restart.restype = C.c_long
    sim.sim_encode_restart.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p]
    o = Oracle.get()
    rng = np.random.default_rng(33)
    quirks = 0
    for it in range(160):
        n = 65536 if it % 2 == 0 else int(rng.integers(1, 65537))
        d = np.ascontiguousarray(crafted(it % 7, n, rng))
        P = int(rng.choice([64, 128, 256, 1024, 2048, 8192]))
        seg = P * int(rng.choice([1, 2, 4]))
        nrec = (65536 + seg - 1) // seg - 1
        lead = int(rng.integers(0, 8))
        cap = slot_bytes(n) + 2 * n + 16
        out = np.full(cap, 0x5C, np.uint8)
        rec = np.zeros(3 * max(nrec, 1), np.uint32)
        r = sim.sim_encode_segmented(d.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), cap, lead, P, seg,
                                     nrec, rec.ctypes.data_as(C.c_void_p))
        if r == -2:
            quirks += 1
            continue
        assert r >= 0, (it, n, P, r)
        want = o.encode(STATIC, d)
        assert out[lead:lead + r].tobytes() == want, (it, n, P, lead)
        assert bytes(out[:lead]) == b"\x5c" * lead and out[lead + r] == 0x5C     # nothing outside the payload
        # the same restart points as the one-chain encoder records
        nseg = nrec + 1
        out2 = np.empty(cap, np.uint8)
        rec2 = np.zeros(3 * max(nrec, 1), np.uint32)
        if nrec:
            r2 = sim.sim_encode_restart(d.ctypes.data_as(C.c_void_p), n, out2.ctypes.data_as(C.c_void_p), cap, seg, nseg,
                                        rec2.ctypes.data_as(C.c_void_p))
            assert r2 == r
            for j in range(nrec):
                a, b = rec[3 * j:3 * j + 3], rec2[3 * j:3 * j + 3]
                assert a[0] == b[0] and a[1] == b[1], (it, n, P, seg, j, a, b)
                if a[0] != 0xFFFFFFFF:  # any range with the same range / total serves
                    tot = int(np.bincount(d, minlength=256).clip(max=0x8000 if n == 65536 else None).sum())
                    assert int(a[2]) // tot == int(b[2]) // tot
    assert quirks == 0
