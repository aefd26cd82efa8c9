"""Where the end-to-end step's time goes: b2rc_encode and b2rc_decode timed apart (pinned buffers), next to the two
copy-only phases of bench.py's ceiling (H2D input || D2H container; H2D container || D2H output); scratch tool."""
import os
import sys
import time
from pathlib import Path

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch

from cpprcoder_b200 import api, synth


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
    ctx = api.Context(0)
    data = synth.zipf(n)
    h_src = torch.from_numpy(data).pin_memory()
    h_enc = torch.empty(api.bound(0, n, 65536), dtype=torch.uint8).pin_memory()
    h_dec = torch.empty(n, dtype=torch.uint8).pin_memory()
    src, enc, dec = h_src.numpy(), h_enc.numpy(), h_dec.numpy()
    for _ in range(3):
        out = ctx.encode(0, src, 65536, dst=enc)
        ctx.decode(out, dst=dec)
    te, td = [], []
    for _ in range(7):
        t0 = time.perf_counter()
        out = ctx.encode(0, src, 65536, dst=enc)
        t1 = time.perf_counter()
        ctx.decode(out, dst=dec)
        t2 = time.perf_counter()
        te.append(t1 - t0)
        td.append(t2 - t1)
    c = out.size
    assert (dec == src).all()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(c, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def phase(up, down, both=True):
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            up[0].copy_(up[1], non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                down[0].copy_(down[1], non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        return time.perf_counter() - t0

    res = {}
    for name, up, down in (("enc", (d_a, h_src), (h_enc[:c], d_b)), ("dec", (d_b, h_enc[:c]), (h_dec, d_a))):
        phase(up, down)
        res[name + "_duplex"] = min(phase(up, down) for _ in range(5))
        res[name + "_h2d_alone"] = min(phase(up, down, False) for _ in range(5))
        res[name + "_d2h_alone"] = min(phase((down[0], down[1]), None, False) for _ in range(5))
    print(f"n {n >> 20} MiB container {c / n:.4f}")
    print(f"b2rc_encode {min(te) * 1e3:7.2f} ms (median {sorted(te)[3] * 1e3:.2f})   copies only: duplex {res['enc_duplex'] * 1e3:.2f}  "
          f"H2D alone {res['enc_h2d_alone'] * 1e3:.2f}  D2H alone {res['enc_d2h_alone'] * 1e3:.2f}")
    print(f"b2rc_decode {min(td) * 1e3:7.2f} ms (median {sorted(td)[3] * 1e3:.2f})   copies only: duplex {res['dec_duplex'] * 1e3:.2f}  "
          f"H2D alone {res['dec_h2d_alone'] * 1e3:.2f}  D2H alone {res['dec_d2h_alone'] * 1e3:.2f}")


if __name__ == "__main__":
    main()
