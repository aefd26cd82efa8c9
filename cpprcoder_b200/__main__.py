"""File-level front end: B2RC containers on disk and the reference
harness's row format.

    python -m cpprcoder_b200 encode [--adaptive | --coder static|adaptive|rans|rans-word] [--block N] IN OUT
    python -m cpprcoder_b200 decode IN OUT
    python -m cpprcoder_b200 rows   [--block N] [--blk] FILE...   # |file|ratio|enc MiB/s|dec MiB/s| per coder
    python -m cpprcoder_b200 blksort IN OUT                 # blksort::BlkSort::encode, the reference's byte format
    python -m cpprcoder_b200 unblksort IN OUT               # blksort::BlkSort::decode

`rows` prints what the reference's run_rangecoder / run_adaptive / run_ans / run_ans_simd print
(test/main.cpp:104-107, :290-294, :367-395, :457-485): ratio = original / coded bytes, speeds in
MiB/s, and fails loudly on a mismatch.
All coding happens on the GPU through libb2rc.so.
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

from . import api


def _read(path: str) -> np.ndarray:
    return np.fromfile(path, dtype=np.uint8)


def cmd_encode(a) -> int:
    ctx = api.Context()
    data = _read(a.input)
    mode = api.MODE_ADAPTIVE if a.adaptive else api.MODE_NAMES[a.coder]
    out = ctx.encode(mode, data, a.block)
    Path(a.output).write_bytes(out.tobytes())
    print(f"{a.input}: {data.size} -> {out.size} bytes ({out.size / max(data.size, 1):.6f})")
    return 0


def cmd_decode(a) -> int:
    ctx = api.Context()
    out = ctx.decode(_read(a.input))
    Path(a.output).write_bytes(out.tobytes())
    print(f"{a.input}: -> {out.size} bytes")
    return 0


def cmd_blksort(a) -> int:
    ctx = api.Context()
    data = _read(a.input)
    out = ctx.blk_decode(data) if a.cmd == "unblksort" else ctx.blk_encode(data)
    Path(a.output).write_bytes(out.tobytes())
    print(f"{a.input}: {data.size} -> {out.size} bytes")
    return 0


def _blk_rows(ctx, path, data, block) -> int:
    """run_blksort and run_zlib_blk (test/main.cpp:791-1002) with the static coder where zlib stood."""
    mib = data.size / (1024.0 * 1024.0)
    t0 = time.perf_counter()
    coded = ctx.blk_encode(data)
    t1 = time.perf_counter()
    back = ctx.blk_decode(coded)
    t2 = time.perf_counter()
    print("|%s|%f|%f|%f|" % (path, data.size / max(coded.size, 1), mib / (t1 - t0), mib / (t2 - t1)))
    t0 = time.perf_counter()
    enc = ctx.encode(api.MODE_STATIC, ctx.blk_encode(data), block)
    t1 = time.perf_counter()
    back2 = ctx.blk_decode(ctx.decode(enc))
    t2 = time.perf_counter()
    print("|%s|%f|%f|%f|" % (path, data.size / max(enc.size, 1), mib / (t1 - t0), mib / (t2 - t1)))
    return int(back.tobytes() != data.tobytes()) + int(back2.tobytes() != data.tobytes())


def cmd_rows(a) -> int:
    ctx = api.Context()
    bad = 0
    for path in a.files:
        data = _read(path)
        if a.blk:
            bad += _blk_rows(ctx, path, data, a.block)
            continue
        for mode in ctx.supported_modes():
            t0 = time.perf_counter()
            enc = ctx.encode(mode, data, a.block)
            t1 = time.perf_counter()
            dec = ctx.decode(enc)
            t2 = time.perf_counter()
            mib = data.size / (1024.0 * 1024.0)
            print("|%s|%f|%f|%f|" % (path, data.size / max(enc.size, 1), mib / (t1 - t0), mib / (t2 - t1)))
            if dec.tobytes() != data.tobytes():
                print(f"{path}: round trip MISMATCH", file=sys.stderr)
                bad += 1
    return 1 if bad else 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m cpprcoder_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    e = sub.add_parser("encode")
    e.add_argument("--adaptive", action="store_true")
    e.add_argument("--coder", default="static", choices=sorted(api.MODE_NAMES))
    e.add_argument("--block", type=int, default=api.DEFAULT_BLOCK)
    e.add_argument("input")
    e.add_argument("output")
    e.set_defaults(fn=cmd_encode)
    d = sub.add_parser("decode")
    d.add_argument("input")
    d.add_argument("output")
    d.set_defaults(fn=cmd_decode)
    r = sub.add_parser("rows")
    r.add_argument("--block", type=int, default=api.DEFAULT_BLOCK)
    r.add_argument("--blk", action="store_true", help="block sort alone, and block sort in front of the static coder")
    r.add_argument("files", nargs="+")
    r.set_defaults(fn=cmd_rows)
    for name in ("blksort", "unblksort"):
        b = sub.add_parser(name)
        b.add_argument("input")
        b.add_argument("output")
        b.set_defaults(fn=cmd_blksort)
    a = ap.parse_args(argv)
    return a.fn(a)


if __name__ == "__main__":
    sys.exit(main())
