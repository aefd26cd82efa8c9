"""Static decode (device resident) for several restart spacings and stream sizes; scratch tool.
    python tools/dec_perf.py [bytes,bytes,...] [seg,seg,...] [generator] [block]"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch

from cpprcoder_b200 import api, synth


def main():
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1 << 30, 1 << 27]
    segs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8192, 4096, 2048, 1024]
    gen = sys.argv[3] if len(sys.argv) > 3 else "zipf"
    block = int(sys.argv[4]) if len(sys.argv) > 4 else 65536
    data = synth.GENERATORS[gen](max(sizes))
    for n in sizes:
        src = torch.from_numpy(data[:n]).cuda()
        dst = torch.empty(n, dtype=torch.uint8, device="cuda")
        for seg in segs:
            os.environ["B2RC_RESTART_SYMS"] = str(seg)
            ctx = api.Context(0)
            enc, used = ctx.encode_device(0, src, block=block)
            best_e, best_d = 1e9, 1e9
            for _ in range(5):
                a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                a.record()
                ctx.encode_device(0, src, enc, block)
                b.record()
                ctx.decode_device(enc, used, dst)
                c.record()
                torch.cuda.synchronize()
                best_e, best_d = min(best_e, a.elapsed_time(b)), min(best_d, b.elapsed_time(c))
            ok = torch.equal(dst, src)
            print(f"{gen} n {n >> 20:5d} MiB block {block} seg {seg:5d}: encode {best_e:6.3f} ms decode {best_d:6.3f} ms "
                  f"({n / best_d / 1e6:6.1f} GB/s) ratio {used / n:.5f} ok={ok}", flush=True)
            ctx.close()
            del enc


if __name__ == "__main__":
    main()
