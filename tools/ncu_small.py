"""One launch each of the kernels that small calls and long blocks take (for an ncu capture):
k_enc_ranges2<3> and k_dec_static_seg on a 128 MiB static stream (2048 blocks: what a rank of the 8-GPU strong leg
codes), k_dec_adaptive_seg<LeaflessW> on 256 MiB of 1 MiB blocks.
    ncu --set full --clock-control none -k regex:'k_enc_ranges2|k_dec_adaptive_seg|k_dec_static_seg' -o out python tools/ncu_small.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch

from cpprcoder_b200 import api, synth


def main():
    ctx = api.Context(0)
    for mode, gen, n, block in ((0, "zipf", 128 << 20, 65536), (1, "kennedy", 256 << 20, 1 << 20)):
        data = synth.GENERATORS[gen](n)
        src = torch.from_numpy(data).cuda()
        dst = torch.empty(n, dtype=torch.uint8, device="cuda")
        enc, used = ctx.encode_device(mode, src, block=block)
        ctx.decode_device(enc, used, dst)
        torch.cuda.synchronize()
        assert torch.equal(dst, src)
        print(mode, gen, n, block, used / n)


if __name__ == "__main__":
    main()
