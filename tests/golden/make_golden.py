#!/usr/bin/env python
"""Regenerates tests/golden/golden.json from the UNMODIFIED reference.

Run in the build container, where /root/reference is mounted and
`make -C oracle` has produced oracle/_ref/libcpprcoder_ref.so:

    python tests/golden/make_golden.py

Every number below comes out of the reference's own RangeEncoder<> /
AdaptiveRangeEncoder<> (through oracle/ref_shim.cpp); nothing from this
repository's coder or oracle port takes part.  The file pins
  * whole-file output sizes + FNV-1a-64 for the 11 Canterbury files (these
    reproduce the ratio columns of the reference's README.md:20-30, 36-46),
  * per-block payload sizes + hashes at 64 KiB blocks for the same files,
  * the edge cases of SURVEY.md section 8c as hex payloads,
  * seeded synthetic streams (cpprcoder_b200/synth.py) at several block sizes,
    including blocks > 64 KiB that exercise the order-dependent count() halving.
"""
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

import numpy as np  # noqa: E402
from _oracle import CANTERBURY, Ref, canterbury, fnv1a64  # noqa: E402
from cpprcoder_b200 import synth  # noqa: E402

README_RATIOS = {  # compressed / original, reference README.md:20-30 (static) and :36-46 (adaptive)
    "alice29.txt": (0.574532, 0.573000), "asyoulik.txt": (0.605293, 0.603400), "cp.html": (0.674836, 0.662480),
    "fields.c": (0.672646, 0.642511), "grammar.lsp": (0.718893, 0.619457), "kennedy.xls": (0.452938, 0.447426),
    "lcet10.txt": (0.585129, 0.584625), "plrabn12.txt": (0.567788, 0.567367), "ptt5": (0.157010, 0.152158),
    "sum": (0.679759, 0.670450), "xargs.1": (0.735510, 0.648924),
}


def blocks_entry(ref, mode, data, block):
    pays = ref.encode_blocks(mode, data, block, threads=4)
    return {"sizes": [len(p) for p in pays], "fnv": [f"{fnv1a64(p):016x}" for p in pays],
            "cat_fnv": f"{fnv1a64(b''.join(pays)):016x}"}


def main():
    ref = Ref.get()
    out = {"readme_ratios": README_RATIOS, "canterbury": {}, "edge": [], "synthetic": []}
    for name in CANTERBURY:
        d = canterbury(name)
        ent = {"bytes": len(d), "whole": {}, "blocks64k": {}}
        for mode, key in ((0, "static"), (1, "adaptive")):
            w = ref.encode(mode, d)
            ent["whole"][key] = {"size": len(w), "fnv": f"{fnv1a64(w):016x}"}
            ent["blocks64k"][key] = blocks_entry(ref, mode, d, 65536)
        out["canterbury"][name] = ent
    edge_inputs = [("empty", b""), ("1xA", b"A"), ("2xA", b"AA"), ("65535xA", b"A" * 65535), ("65536xA", b"A" * 65536),
                   ("65536xFF", b"\xff" * 65536), ("64x00", bytes(64)), ("AB*32", b"AB" * 32),
                   ("0..255", bytes(range(256))), ("255..0x2", bytes(range(255, -1, -1)) * 2)]
    for label, d in edge_inputs:
        for mode, key in ((0, "static"), (1, "adaptive")):
            w = ref.encode(mode, d)
            # static payloads carry a 516-byte header; keep only size+hash+coded tail for those
            out["edge"].append({"label": label, "mode": key, "n": len(d), "size": len(w), "fnv": f"{fnv1a64(w):016x}",
                                "tail_hex": w[(516 if mode == 0 else 4):][-32:].hex()})
    synth_cases = [("zipf", 1 << 20, 4096), ("zipf", 1 << 20, 65536), ("zipf", (1 << 20) + 12345, 65536),
                   ("mixed", 3 << 20, 65536), ("kennedy", 1 << 20, 16384), ("kennedy", 2 << 20, 262144),
                   ("kennedy", 2 << 20, 1048576), ("mixed", 3 << 20, 1048576)]
    for gen, n, block in synth_cases:
        d = synth.GENERATORS[gen](n)
        for mode, key in ((0, "static"), (1, "adaptive")):
            e = blocks_entry(ref, mode, d, block)
            out["synthetic"].append({"gen": gen, "n": n, "block": block, "mode": key, "src_fnv": f"{fnv1a64(d):016x}",
                                     "sizes": e["sizes"], "cat_fnv": e["cat_fnv"]})
    (HERE / "golden.json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", HERE / "golden.json")


if __name__ == "__main__":
    main()
