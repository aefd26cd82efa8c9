"""Block-sort timing on the device: forward / inverse GB/s for a synthetic stream, rounds histogram.
    python tools/blk_perf.py [bytes] [zipf|mixed|kennedy|text]      (text: the Canterbury text files, tiled)"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cpprcoder_b200 import api, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
    gen = sys.argv[2] if len(sys.argv) > 2 else "zipf"
    if gen == "text":
        sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
        from _oracle import canterbury
        one = np.frombuffer(b"".join(canterbury(f) for f in ("lcet10.txt", "plrabn12.txt", "alice29.txt", "asyoulik.txt")), np.uint8)
        data = np.tile(one, n // one.size + 1)[:n].copy()
    else:
        data = synth.GENERATORS[gen](n)
    ctx = api.Context(0)
    src = torch.from_numpy(data).cuda()
    coded = torch.empty(api.blk_encode_bound(n), dtype=torch.uint8, device="cuda")
    back = torch.empty(n, dtype=torch.uint8, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for it in range(3):
        ev[0].record()
        ctx.blk_encode_device(src, coded)
        ev[1].record()
        ctx.blk_decode_device(coded, dst=back)
        ev[2].record()
        torch.cuda.synchronize()
        f, i = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        print(f"{gen} {n} B: forward {f:.2f} ms = {n / f / 1e6:.1f} GB/s, inverse {i:.2f} ms = {n / i / 1e6:.1f} GB/s")
    assert torch.equal(back, src)
    r = ctx.blk_rounds()
    total, short = r & 0xFF, (r >> 8) & 0xFF
    print("doubling rounds per block (total, of which short):",
          {(int(a), int(b)): int(c) for (a, b), c in zip(*np.unique(np.stack([total, short], 1), axis=0, return_counts=True))})


if __name__ == "__main__":
    main()
