"""How long the tie replay (k_blk_ties) takes for blocks with a period, against the oracle's answer.
    python tools/blk_ties_probe.py [--hard]      (--hard adds the quadratic case: long zero runs with period 16 384)"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from _cases import blk_periodic_cases  # noqa: E402
from _oracle import BlkSort, Oracle  # noqa: E402
from cpprcoder_b200 import api  # noqa: E402


def main():
    ctx = api.Context(0)
    oracle = BlkSort(Oracle.get())
    cases = [(k, v) for k, v in blk_periodic_cases() if v.size == 32768]
    if "--hard" in sys.argv:
        half = np.zeros(16384, np.uint8)
        half[-1] = 1
        cases.append(("zeros-then-one x2", np.tile(half, 2)))
    ctx.blk_encode_device(torch.zeros(32768, dtype=torch.uint8, device="cuda"))
    for label, d in cases:
        src = torch.from_numpy(d.copy()).cuda()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = ctx.blk_encode_device(src).cpu().numpy()
        t1 = time.perf_counter()
        want = oracle.encode(d)
        t2 = time.perf_counter()
        print(f"{label:22s} gpu {1e3 * (t1 - t0):9.2f} ms   cpu oracle {1e3 * (t2 - t1):9.2f} ms   equal {np.array_equal(got, want)}", flush=True)


if __name__ == "__main__":
    main()
