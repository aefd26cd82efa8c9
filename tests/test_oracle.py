"""The oracle must be pinned before it may judge the CUDA path (CPU only).

  * oracle port == committed golden vectors (made by the unmodified reference)
  * oracle port == the reference itself, live, wherever oracle/_ref was built
  * golden whole-file sizes reproduce the reference README's ratio tables
"""
import numpy as np
import pytest

from _cases import crafted
from _oracle import ADAPTIVE, CANTERBURY, STATIC, Oracle, Ref, canterbury, fnv1a64, offsets_of
from cpprcoder_b200 import synth

MODES = [(STATIC, "static"), (ADAPTIVE, "adaptive")]


@pytest.fixture(scope="module")
def oracle(built):
    return Oracle.get()


def test_readme_ratios_reproduced(golden):
    # README.md:20-30 / 36-46 of the reference: compressed/original to six decimals
    for name, (rs, ra) in golden["readme_ratios"].items():
        ent = golden["canterbury"][name]
        assert round(ent["whole"]["static"]["size"] / ent["bytes"], 6) == pytest.approx(rs, abs=1e-6)
        assert round(ent["whole"]["adaptive"]["size"] / ent["bytes"], 6) == pytest.approx(ra, abs=1e-6)


@pytest.mark.parametrize("name", CANTERBURY)
def test_canterbury_whole_and_blocks(oracle, golden, name):
    data = canterbury(name)
    ent = golden["canterbury"][name]
    assert len(data) == ent["bytes"]
    for mode, key in MODES:
        whole = oracle.encode(mode, data)
        assert len(whole) == ent["whole"][key]["size"]
        assert f"{fnv1a64(whole):016x}" == ent["whole"][key]["fnv"]
        assert oracle.decode(mode, whole, len(data)) == data
        pays = oracle.encode_blocks(mode, data, 65536, threads=2)
        assert [len(p) for p in pays] == ent["blocks64k"][key]["sizes"]
        assert [f"{fnv1a64(p):016x}" for p in pays] == ent["blocks64k"][key]["fnv"]
        back = oracle.decode_blocks(mode, b"".join(pays), offsets_of(pays), 65536, len(data), threads=2)
        assert back.tobytes() == data


def test_edge_cases(oracle, golden):
    inputs = {"empty": b"", "1xA": b"A", "2xA": b"AA", "65535xA": b"A" * 65535, "65536xA": b"A" * 65536,
              "65536xFF": b"\xff" * 65536, "64x00": bytes(64), "AB*32": b"AB" * 32, "0..255": bytes(range(256)),
              "255..0x2": bytes(range(255, -1, -1)) * 2}
    for ent in golden["edge"]:
        mode = STATIC if ent["mode"] == "static" else ADAPTIVE
        d = inputs[ent["label"]]
        w = oracle.encode(mode, d)
        assert len(w) == ent["size"], ent
        assert f"{fnv1a64(w):016x}" == ent["fnv"], ent
        assert w[(516 if mode == STATIC else 4):][-32:].hex() == ent["tail_hex"]
    # known short payloads, spelled out
    assert oracle.encode(ADAPTIVE, b"A").hex() == "010000000040ffffbf00"
    assert oracle.encode(ADAPTIVE, b"AA").hex() == "02000000004140beff7e"
    assert oracle.encode(ADAPTIVE, b"").hex() == "00" * 9
    assert oracle.encode(STATIC, b"A" * 65536)[-5:].hex() == "0000000000"
    # the reference's decoder emits one spurious byte for an empty adaptive stream (cpprcoder.h:909-914)
    assert oracle.decode(ADAPTIVE, bytes(9), 4) == b"\x00"


def test_synthetic_golden(oracle, golden):
    for ent in golden["synthetic"]:
        d = synth.GENERATORS[ent["gen"]](ent["n"])
        assert f"{fnv1a64(d):016x}" == ent["src_fnv"], "generator drifted"
        mode = STATIC if ent["mode"] == "static" else ADAPTIVE
        pays = oracle.encode_blocks(mode, d, ent["block"], threads=4)
        assert [len(p) for p in pays] == ent["sizes"], (ent["gen"], ent["block"], ent["mode"])
        assert f"{fnv1a64(b''.join(pays)):016x}" == ent["cat_fnv"]


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built here (reference tree absent)")
def test_port_matches_live_reference(oracle):
    ref = Ref.get()
    rng = np.random.default_rng(2024)
    seen = {"carries_with_run": 0, "max_pending_run": 0, "rescales": 0}
    for it in range(140):
        n = int(rng.integers(1, 70000)) if it % 5 else int(rng.integers(65536, 400000))
        d = crafted(it % 7, n, rng)
        for mode, _ in MODES:
            a, st = oracle.encode(mode, d, with_stats=True)
            assert a == ref.encode(mode, d), (it, mode, n)
            assert ref.decode(mode, a, n) == d.tobytes()
            assert oracle.decode(mode, a, n) == d.tobytes()
            seen["carries_with_run"] += st["carries_with_run"]
            seen["max_pending_run"] = max(seen["max_pending_run"], st["max_pending_run"])
            seen["rescales"] += st["rescales"]
    # the crafted inputs must actually reach the rare branches they are there for
    assert seen["carries_with_run"] > 100 and seen["max_pending_run"] > 1000 and seen["rescales"] > 10


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built here")
def test_adaptive_rescale_matches_reference(oracle):
    # >= 2^24 - 256 symbols in ONE stream reach AdaptiveFrequencyTable's halving (cpprcoder.h:1138-1154);
    # the reference's own test_adaptive() (test/main.cpp:1200-1238) is the only caller that gets there.
    n = (1 << 24) + 4096
    d = synth.zipf(n, seed=77)
    a, st = oracle.encode(ADAPTIVE, d, with_stats=True)
    assert st["rescales"] >= 1
    assert a == Ref.get().encode(ADAPTIVE, d)
    assert oracle.decode(ADAPTIVE, a, n) == d.tobytes()
