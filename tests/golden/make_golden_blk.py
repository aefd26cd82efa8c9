#!/usr/bin/env python
"""Regenerates tests/golden/golden_blk.json from the UNMODIFIED reference block sort.

Run in the build container (where /root/reference is mounted and `make -C oracle` has
produced oracle/_ref/libcpprcoder_ref.so):

    python tests/golden/make_golden_blk.py

Every number comes out of blksort::BlkSort::encode of the reference's blksort.h (through
oracle/ref_shim.cpp, ref_blk_encode); nothing from this repository's kernels or oracle port
takes part.  The reference ships no known answers for the transform, so these vectors are
what pins the oracle where the reference cannot travel.  Inputs: tests/_cases.py
(blk_cases, blk_periodic_cases), all seeded.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

from _cases import blk_cases, blk_periodic_cases  # noqa: E402
from _oracle import BLK_BLOCK, BLK_CODED, BlkSort, Ref, fnv1a64  # noqa: E402


def main():
    ref = BlkSort(Ref.get())
    out = {"cases": []}
    for group, cases in (("plain", blk_cases()), ("periodic", blk_periodic_cases())):
        for label, d in cases:
            coded = ref.encode(d, threads=8)
            back = ref.decode(coded, threads=8)
            assert np.array_equal(back, d), label
            nb = d.size // BLK_BLOCK
            rows = [int(coded[b * BLK_CODED + BLK_BLOCK]) | int(coded[b * BLK_CODED + BLK_BLOCK + 1]) << 8 for b in range(nb)]
            out["cases"].append({"group": group, "label": label, "n": int(d.size), "src_fnv": f"{fnv1a64(d):016x}",
                                 "coded": int(coded.size), "coded_fnv": f"{fnv1a64(coded):016x}", "rows": rows})
    (HERE / "golden_blk.json").write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {len(out['cases'])} cases")


if __name__ == "__main__":
    main()
