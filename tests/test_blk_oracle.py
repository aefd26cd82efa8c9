"""CPU checks of the block-sort oracle (oracle/blk_oracle.c).

The restatement is pinned three ways: against golden vectors the unmodified blksort.h produced
(tests/golden/golden_blk.json, made by tests/golden/make_golden_blk.py), live against that
reference where oracle/_ref exists, and by round trips.  Blocks with a period are part of it:
their row number depends on the exact swaps of the reference's multikey quicksort, which the
restatement replays."""
import json
from pathlib import Path

import numpy as np
import pytest

from _cases import blk_cases, blk_periodic_cases
from _oracle import (BLK_BLOCK, BLK_CODED, BlkSort, Oracle, Ref, blk_decoded_size, blk_encode_bound, fnv1a64)

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "golden_blk.json").read_text())
BY_LABEL = {c["label"]: c for c in GOLDEN["cases"]}


@pytest.fixture(scope="module")
def oracle(built):
    return BlkSort(Oracle.get())


def test_bounds_follow_the_reference():
    # BlkSort::encodeBound (blksort.h:404-409): two bytes per FULL block, the tail as it is
    assert blk_encode_bound(0) == 0
    assert blk_encode_bound(32767) == 32767
    assert blk_encode_bound(32768) == 32770
    assert blk_encode_bound(152089) == 152097
    assert blk_decoded_size(152097) == 152089
    lib = Oracle.get().lib
    lib.bso_encode_bound.restype = lib.bso_decode_bound.restype = np.ctypeslib.ctypes.c_uint32
    for n in (0, 1, 32767, 32768, 32769, 65536, 1 << 20, 152089):
        assert lib.bso_encode_bound(n) == blk_encode_bound(n)
        assert lib.bso_decode_bound(n) == n  # as written: an upper bound, not the decoded size (blksort.h:411-416)


@pytest.mark.parametrize("label,data", blk_cases() + blk_periodic_cases(), ids=lambda x: x if isinstance(x, str) else "")
def test_oracle_matches_golden(oracle, label, data):
    g = BY_LABEL[label]
    assert g["n"] == data.size and g["src_fnv"] == f"{fnv1a64(data):016x}", "seeded input drifted"
    coded = oracle.encode(data, threads=8)
    assert coded.size == g["coded"]
    rows = [int(coded[b * BLK_CODED + BLK_BLOCK]) | int(coded[b * BLK_CODED + BLK_BLOCK + 1]) << 8
            for b in range(data.size // BLK_BLOCK)]
    assert rows == g["rows"]
    assert f"{fnv1a64(coded):016x}" == g["coded_fnv"]
    assert np.array_equal(oracle.decode(coded, threads=4), data)


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (reference tree not mounted)")
def test_oracle_matches_reference_live(oracle):
    ref = BlkSort(Ref.get())
    rng = np.random.default_rng(77)
    for k in range(12):
        n = int(rng.integers(0, 5 * BLK_BLOCK))
        alpha = int(rng.choice([2, 4, 16, 256]))
        d = rng.integers(0, alpha, n, dtype=np.uint8)
        if k % 3 == 0 and n > 4000:  # long repeats: many passes of the quicksort, the heap fallback
            d[n // 2:n // 2 + n // 4] = d[:n // 4]
        a, b = oracle.encode(d, threads=4), ref.encode(d, threads=4)
        assert np.array_equal(a, b), f"case {k}: n={n} alpha={alpha}"
        assert np.array_equal(ref.decode(a), d) and np.array_equal(oracle.decode(b), d)


def test_decode_rejects_a_row_number_out_of_range(oracle):
    d = np.arange(BLK_BLOCK, dtype=np.uint32).astype(np.uint8)
    coded = oracle.encode(d).copy()
    coded[BLK_BLOCK + 1] |= 0x80
    with pytest.raises(RuntimeError):
        oracle.decode(coded)


def test_periodic_detector(oracle):
    for label, d in blk_periodic_cases():
        assert oracle.periodic_blocks(d), label
    assert oracle.periodic_blocks(np.random.default_rng(1).integers(0, 256, 2 * BLK_BLOCK, dtype=np.uint8)) == []
