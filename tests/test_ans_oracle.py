"""Pins the rANS oracle (oracle/ans_oracle.c) before it may judge the CUDA path (CPU only).

  * oracle == committed golden vectors made by the unmodified cppans.h
    (tests/golden/golden_rans.json, tests/golden/make_golden_rans.py)
  * oracle == the reference itself, live, wherever oracle/_ref was built
  * the normalised model obeys what the reference's own debug asserts state
    (cppans.h:165-171): slices sum to the scale, present symbols keep a slice.
"""
import json

import numpy as np
import pytest

from _oracle import CANTERBURY, GOLDEN, RANS_BYTE, RANS_HEADER, RANS_WORD, Oracle, Ref, canterbury, fnv1a64, offsets_of
from cpprcoder_b200 import synth

VARIANTS = [(RANS_BYTE, "rans_byte"), (RANS_WORD, "rans_word")]
EDGE = {"1xA": b"A", "2xA": b"AA", "7xA": b"A" * 7, "8xA": b"A" * 8, "9xA": b"A" * 9, "65535xA": b"A" * 65535,
        "65536xA": b"A" * 65536, "65536xFF": b"\xff" * 65536, "64x00": bytes(64), "AB*32": b"AB" * 32,
        "0..255": bytes(range(256)), "255..0x2": bytes(range(255, -1, -1)) * 2,
        "rare": b"A" * 60000 + bytes(range(256)), "ABx4097": b"AB" * 4097}


@pytest.fixture(scope="module")
def oracle(built):
    return Oracle.get()


@pytest.fixture(scope="module")
def golden_rans():
    return json.loads((GOLDEN / "golden_rans.json").read_text())


@pytest.mark.parametrize("name", CANTERBURY)
def test_canterbury_whole_and_blocks(oracle, golden_rans, name):
    data = canterbury(name)
    ent = golden_rans["canterbury"][name]
    assert len(data) == ent["bytes"]
    for mode, key in VARIANTS:
        whole = oracle.encode(mode, data)
        assert len(whole) == ent["whole"][key]["size"]
        assert f"{fnv1a64(whole):016x}" == ent["whole"][key]["fnv"]
        assert oracle.decode(mode, whole, len(data)) == data
        pays = oracle.encode_blocks(mode, data, 65536, threads=2)
        assert [len(p) for p in pays] == ent["blocks64k"][key]["sizes"]
        assert f"{fnv1a64(b''.join(pays)):016x}" == ent["blocks64k"][key]["cat_fnv"]
        back = oracle.decode_blocks(mode, b"".join(pays), offsets_of(pays), 65536, len(data), threads=2)
        assert back.tobytes() == data


def test_edge_cases(oracle, golden_rans):
    for ent in golden_rans["edge"]:
        mode = RANS_BYTE if ent["mode"] == "rans_byte" else RANS_WORD
        d = EDGE[ent["label"]]
        w = oracle.encode(mode, d)
        assert len(w) == ent["size"], ent
        assert f"{fnv1a64(w):016x}" == ent["fnv"], ent
        assert w[RANS_HEADER:][-48:].hex() == ent["tail_hex"]
        assert oracle.decode(mode, w, len(d)) == d
    # one symbol owning the whole 12-bit scale makes the word coder emit 16 bits per symbol
    # (the u32 bound at cppans.h:357 wraps to zero): 8 states flushed + one word per symbol
    assert len(oracle.encode(RANS_WORD, b"A" * 65536)) == RANS_HEADER + 32 + 2 * 65536
    assert len(oracle.encode(RANS_BYTE, b"A" * 65536)) == RANS_HEADER + 4


def test_synthetic_streams(oracle, golden_rans):
    for ent in golden_rans["synthetic"]:
        mode = RANS_BYTE if ent["mode"] == "rans_byte" else RANS_WORD
        d = synth.GENERATORS[ent["gen"]](ent["n"])
        assert f"{fnv1a64(d):016x}" == ent["src_fnv"]
        pays = oracle.encode_blocks(mode, d, ent["block"], threads=4)
        assert [len(p) for p in pays] == ent["sizes"]
        assert f"{fnv1a64(b''.join(pays)):016x}" == ent["cat_fnv"]
        back = oracle.decode_blocks(mode, b"".join(pays), offsets_of(pays), ent["block"], len(d), threads=4)
        assert back.tobytes() == d.tobytes()


@pytest.mark.parametrize("bits", [12, 14])
def test_model_invariants(oracle, bits):
    rng = np.random.default_rng(bits)
    cases = [canterbury("alice29.txt")[:65536], canterbury("ptt5")[:65536], bytes(range(256)) * 3,
             b"A" * 60000 + bytes(range(256)), rng.integers(0, 256, 300, dtype=np.uint8).tobytes(),
             synth.zipf(65536).tobytes()]
    for d in cases:
        a = np.frombuffer(d, dtype=np.uint8)
        freq, cum = oracle.rans_model(a, bits)
        assert cum[0] == 0 and cum[256] == 1 << bits
        assert np.array_equal(np.diff(cum.astype(np.int64)), freq.astype(np.int64))
        present = np.bincount(a, minlength=256) > 0
        assert np.all(freq[present] >= 1)          # cppans.h:169
        assert np.all(freq[~present] == 0)         # cppans.h:167


def test_rejects_bad_payloads(oracle):
    d = canterbury("fields.c")
    for mode, _ in VARIANTS:
        w = bytearray(oracle.encode(mode, d))
        with pytest.raises(RuntimeError):
            oracle.decode(mode, bytes(w[:RANS_HEADER - 1]), len(d))     # shorter than the header
        with pytest.raises(RuntimeError):
            oracle.decode(mode, bytes(w), len(d) - 1)                   # destination too small
        bad = bytearray(w)
        bad[4 + 4 * 256 + 1] ^= 0x40                                    # cum[256] no longer the scale
        with pytest.raises(RuntimeError):
            oracle.decode(mode, bytes(bad), len(d))


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (reference tree not mounted)")
def test_oracle_equals_reference_live(oracle):
    ref = Ref.get()
    rng = np.random.default_rng(7)
    cases = [canterbury(n)[:200000] for n in ("alice29.txt", "kennedy.xls", "ptt5", "sum", "xargs.1")]
    cases += [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (1, 2, 7, 8, 9, 15, 16, 17, 100, 4097)]
    cases += [rng.integers(0, 3, n, dtype=np.uint8).tobytes() for n in (5, 64, 1000, 65536)]
    cases += [b"A" * n for n in (1, 8, 4096, 65536)]
    for d in cases:
        for mode, _ in VARIANTS:
            w = oracle.encode(mode, d)
            assert w == ref.encode(mode, d)
            assert ref.decode(mode, w, len(d)) == d
            assert oracle.decode(mode, w, len(d)) == d
