"""Adaptive coder on the mixed stream: encode_device / decode_device times for several restart spacings; scratch tool.
    python tools/adapt_perf.py [bytes] [seg,seg,...]"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch

from cpprcoder_b200 import api, synth


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
    segs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 16384, 8192, 4096]
    gen = sys.argv[3] if len(sys.argv) > 3 else "mixed"
    data = synth.GENERATORS[gen](n)
    src = torch.from_numpy(data).cuda()
    enc = torch.empty(api.bound(1, n, 65536), dtype=torch.uint8, device="cuda")
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    for seg in segs:
        os.environ["B2RC_ADAPTIVE_RESTART_SYMS"] = str(seg)
        ctx = api.Context(0)
        for _ in range(2):
            _, used = ctx.encode_device(1, src, enc, 65536)
            ctx.decode_device(enc, used, dst)
        te = td = 1e9
        for _ in range(3):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            ctx.encode_device(1, src, enc, 65536)
            ev[1].record()
            ctx.decode_device(enc, used, dst)
            ev[2].record()
            torch.cuda.synchronize()
            te, td = min(te, ev[0].elapsed_time(ev[1])), min(td, ev[1].elapsed_time(ev[2]))
        print(f"{gen} seg {seg:6d}: encode {te:7.3f} ms  decode {td:7.3f} ms  ratio {used / n:.5f}  ok={torch.equal(dst, src)}", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
