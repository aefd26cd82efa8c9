"""Launches every kernel of libb2rc.so once or twice on a 64 MiB Zipf stream -- the command profiled for
profiles/r1_ncu_all_kernels.csv (ncu --set full over this script, one row per kernel).
    python tools/all_kernels_once.py [bytes]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cpprcoder_b200 import api, synth  # noqa: E402


def round_trip(ctx, mode, src, block):
    enc, used = ctx.encode_device(mode, src, block=block)
    dst = torch.empty_like(src)
    assert ctx.decode_device(enc, used, dst) == src.numel() and torch.equal(dst, src)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 << 20
    data = synth.zipf(n)
    src = torch.from_numpy(data).cuda()
    ctx = api.Context(0)
    for mode in ctx.supported_modes():        # containers with restart points where the coder has them
        round_trip(ctx, mode, src, 65536)
    round_trip(ctx, api.MODE_STATIC, src, 262144)    # k_hist_wide, k_enc_static<wide>
    round_trip(ctx, api.MODE_ADAPTIVE, src, 262144)  # the u32 model tables
    os.environ["B2RC_RESTART_SYMS"] = "0"            # containers without the table: the one-chain decoders
    plain = api.Context(0)
    round_trip(plain, api.MODE_STATIC, src, 65536)
    round_trip(plain, api.MODE_RANS_BYTE, src, 65536)
    plain.close()
    # block sort: forward, inverse (station walk), and one periodic block (tie replay, doubling inverse)
    part = torch.cat([src[:n // 2], torch.from_numpy(np.tile(np.arange(64, dtype=np.uint8), 512)).cuda()])
    coded = ctx.blk_encode_device(part)
    assert torch.equal(ctx.blk_decode_device(coded), part)
    print("launches:", ctx.launches)


if __name__ == "__main__":
    main()
