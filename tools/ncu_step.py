"""One warm-up step and one more of `encode_device` + `decode_device` for ncu to capture (every kernel of the step once):
    ncu --set full --clock-control none --import-source on -s <launches of one step> -c <the same> -o out python tools/ncu_step.py static|adaptive
static: k_hist k_enc_ranges k_scan k_enc_seg k_enc_seams k_put_table k_dec_static_seg (7);
adaptive: k_enc_adaptive k_scan k_compact k_put_table k_dec_adaptive_seg (5)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch

from cpprcoder_b200 import api, synth


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "static"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 30)
    mode, gen = (0, "zipf") if what == "static" else (1, "mixed")
    src = torch.from_numpy(synth.GENERATORS[gen](n)).cuda()
    ctx = api.Context(0)
    enc = torch.empty(api.bound(mode, n, 65536), dtype=torch.uint8, device="cuda")
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        _, used = ctx.encode_device(mode, src, enc, 65536)
        ctx.decode_device(enc, used, dst)
    torch.cuda.synchronize()
    assert torch.equal(dst, src)
    print(f"{what}: {ctx.launches} launches in two steps, ratio {used / n:.5f}")
    ctx.close()


if __name__ == "__main__":
    main()
