#!/usr/bin/env python
"""Regenerates tests/golden/golden_rans.json from the UNMODIFIED reference rANS coder.

Run in the build container (where /root/reference is mounted and `make -C oracle` has
produced oracle/_ref/libcpprcoder_ref.so):

    python tests/golden/make_golden_rans.py

Every number comes out of cppans::rANS::encode / ::encode_simd of the reference's
cppans.h (through oracle/ref_shim.cpp, modes 2 and 3); nothing from this repository's
coder or oracle port takes part.  The reference ships no known answers for this coder,
so these vectors are what pins the oracle where the reference cannot travel.
"""
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

from _oracle import CANTERBURY, RANS_BYTE, RANS_WORD, Ref, canterbury, fnv1a64  # noqa: E402
from cpprcoder_b200 import synth  # noqa: E402

VARIANTS = ((RANS_BYTE, "rans_byte"), (RANS_WORD, "rans_word"))
EDGE = [("1xA", b"A"), ("2xA", b"AA"), ("7xA", b"A" * 7), ("8xA", b"A" * 8), ("9xA", b"A" * 9),
        ("65535xA", b"A" * 65535), ("65536xA", b"A" * 65536), ("65536xFF", b"\xff" * 65536), ("64x00", bytes(64)),
        ("AB*32", b"AB" * 32), ("0..255", bytes(range(256))), ("255..0x2", bytes(range(255, -1, -1)) * 2),
        ("rare", b"A" * 60000 + bytes(range(256))), ("ABx4097", b"AB" * 4097)]


def blocks_entry(ref, mode, data, block):
    pays = ref.encode_blocks(mode, data, block, threads=4)
    return {"sizes": [len(p) for p in pays], "cat_fnv": f"{fnv1a64(b''.join(pays)):016x}"}


def main():
    ref = Ref.get()
    out = {"canterbury": {}, "edge": [], "synthetic": []}
    for name in CANTERBURY:
        d = canterbury(name)
        ent = {"bytes": len(d), "whole": {}, "blocks64k": {}}
        for mode, key in VARIANTS:
            w = ref.encode(mode, d)
            ent["whole"][key] = {"size": len(w), "fnv": f"{fnv1a64(w):016x}"}
            ent["blocks64k"][key] = blocks_entry(ref, mode, d, 65536)
        out["canterbury"][name] = ent
    for label, d in EDGE:
        for mode, key in VARIANTS:
            w = ref.encode(mode, d)
            out["edge"].append({"label": label, "mode": key, "n": len(d), "size": len(w), "fnv": f"{fnv1a64(w):016x}",
                                "tail_hex": w[1032:][-48:].hex()})
    for gen, n, block in [("zipf", 1 << 20, 4096), ("zipf", (1 << 20) + 12345, 65536), ("mixed", 3 << 20, 65536),
                          ("kennedy", 2 << 20, 262144)]:
        d = synth.GENERATORS[gen](n)
        for mode, key in VARIANTS:
            e = blocks_entry(ref, mode, d, block)
            out["synthetic"].append({"gen": gen, "n": n, "block": block, "mode": key, "src_fnv": f"{fnv1a64(d):016x}",
                                     "sizes": e["sizes"], "cat_fnv": e["cat_fnv"]})
    (HERE / "golden_rans.json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", HERE / "golden_rans.json")


if __name__ == "__main__":
    main()
