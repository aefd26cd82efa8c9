"""Per-kernel device timings (CUDA events) of the coder on a synthetic stream; scratch tool."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch

from cpprcoder_b200 import api, synth


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 28)
    gen = sys.argv[2] if len(sys.argv) > 2 else "zipf"
    block = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
    ctx = api.Context(0)
    t0 = time.time()
    data = synth.GENERATORS[gen](n)
    src = torch.from_numpy(data).cuda()
    print(f"generated {n} bytes of {gen} in {time.time() - t0:.1f}s, block {block}", flush=True)
    nb = api.nblocks(n, block)
    which = sys.argv[4].split(",") if len(sys.argv) > 4 else ["static", "adaptive", "rans", "rans-word"]
    for name in which:
        mode = api.MODE_NAMES[name]
        freq = None
        if mode == 0:
            freq = ctx.histogram(src, block)
            ms = timed(lambda: ctx.histogram(src, block, freq))
            print(f"{name:9s} K1 hist      {ms:8.3f} ms  {n / ms / 1e6:8.1f} GB/s")
        slots, stride, sizes, err = ctx.encode_blocks(mode, src, block, freq16=freq)
        ms = timed(lambda: ctx.encode_blocks(mode, src, block, freq16=freq, slots=slots, sizes=sizes, err=err))
        print(f"{name:9s} K2 encode    {ms:8.3f} ms  {n / ms / 1e6:8.1f} GB/s")
        if mode >= 2:
            ctx.profile(True)
            ctx.encode_blocks(mode, src, block, slots=slots, sizes=sizes, err=err)
            torch.cuda.synchronize()
            km = ctx.kernel_ms()
            ctx.profile(False)
            print(f"{name:9s}    model {km.get('histogram', -1):.3f} ms, coder {km.get('encode', -1):.3f} ms")
        offsets = ctx.scan(sizes, nb)
        ms = timed(lambda: ctx.scan(sizes, nb, offsets))
        total = int(offsets[nb].item())
        print(f"{name:9s} K4 scan      {ms:8.3f} ms  ratio {total / n:.6f}")
        payload = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: ctx.compact(slots, stride, sizes, offsets, nb, payload, err, mode))
        print(f"{name:9s} K4 compact   {ms:8.3f} ms  {2 * total / ms / 1e6:8.1f} GB/s (r+w)")
        dst = torch.empty(n, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: ctx.decode_blocks(mode, payload, total, offsets, nb, dst, n, block))
        print(f"{name:9s} K3 decode    {ms:8.3f} ms  {n / ms / 1e6:8.1f} GB/s   ok={torch.equal(dst, src)} err={int(err[0])}")
        enc, used = ctx.encode_device(mode, src, block=block)
        ms = timed(lambda: ctx.encode_device(mode, src, enc, block=block))
        print(f"{name:9s} encode_device {ms:8.3f} ms {n / ms / 1e6:8.1f} GB/s")
        ms = timed(lambda: ctx.decode_device(enc, used, dst))
        print(f"{name:9s} decode_device {ms:8.3f} ms {n / ms / 1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
