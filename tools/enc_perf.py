"""Static encode: per-kernel device times of the segmented path (range pass, segments, seams) for several
segment lengths, next to the one-chain kernel; scratch tool.
    python tools/enc_perf.py [bytes] [generator] [block] [P,P,...]"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch

from cpprcoder_b200 import api, synth


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
    gen = sys.argv[2] if len(sys.argv) > 2 else "zipf"
    block = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
    plist = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 1024, 2048, 4096, 8192]
    data = synth.GENERATORS[gen](n)
    src = torch.from_numpy(data).cuda()
    enc = torch.empty(api.bound(0, n, block), dtype=torch.uint8, device="cuda")
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    ref = None
    for P in plist:
        os.environ["B2RC_ENC_SEG_SYMS"] = str(P)
        ctx = api.Context(0)
        for _ in range(2):
            _, used = ctx.encode_device(0, src, enc, block)
        ctx.profile(True)
        best, km = 1e9, {}
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.encode_device(0, src, enc, block)
            b.record()
            torch.cuda.synchronize()
            if a.elapsed_time(b) < best:
                best, km = a.elapsed_time(b), ctx.kernel_ms()
        ctx.profile(False)
        ctx.decode_device(enc, used, dst)
        ok = torch.equal(dst, src)
        h = hash(enc[:used].cpu().numpy().tobytes())
        ref = h if ref is None else ref
        parts = " ".join(f"{k}={v:.3f}" for k, v in km.items() if k in ("histogram", "ranges", "encode", "scan", "compact", "seams"))
        print(f"block {block} P {P:5d}: encode_device {best:7.3f} ms {n / best / 1e6:7.1f} GB/s  [{parts}]  roundtrip={ok} same_bytes={h == ref}",
              flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
