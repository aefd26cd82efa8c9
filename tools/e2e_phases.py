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
    # the pipeline's shape without its kernels: 16 chunks, each on a stream of its own, a chunk's copy home queued
    # behind its copy in (what b2rc_decode's streams look like to the copy engines)
    streams = [torch.cuda.Stream() for _ in range(16)]

    def chunked(up_dev, up_host, down_host, down_dev, k=16):
        t0 = time.perf_counter()
        nu, nd = up_dev.numel(), down_dev.numel()
        for i in range(k):
            a0, a1 = nu * i // k, nu * (i + 1) // k
            b0, b1 = nd * i // k, nd * (i + 1) // k
            with torch.cuda.stream(streams[i]):
                up_dev[a0:a1].copy_(up_host[a0:a1], non_blocking=True)
                down_host[b0:b1].copy_(down_dev[b0:b1], non_blocking=True)
        for st in streams:
            st.synchronize()
        return time.perf_counter() - t0

    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def two_streams(up_dev, up_host, down_host, down_dev, k=16):
        # the same chunks, but every copy in on ONE stream and every copy home on ONE other, tied by events
        t0 = time.perf_counter()
        nu, nd = up_dev.numel(), down_dev.numel()
        for i in range(k):
            a0, a1 = nu * i // k, nu * (i + 1) // k
            b0, b1 = nd * i // k, nd * (i + 1) // k
            with torch.cuda.stream(s_in):
                up_dev[a0:a1].copy_(up_host[a0:a1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev)
                down_host[b0:b1].copy_(down_dev[b0:b1], non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()
        return time.perf_counter() - t0

    res["enc_two"] = min(two_streams(d_a, h_src, h_enc[:c], d_b) for _ in range(5))
    res["dec_two"] = min(two_streams(d_b, h_enc[:c], h_dec, d_a) for _ in range(5))
    print(f"copies only, 16 chunks, ONE stream in and ONE stream home: encode shape {res['enc_two'] * 1e3:.2f} ms, "
          f"decode shape {res['dec_two'] * 1e3:.2f} ms")
    for k in (4, 8, 32, 64, 128):
        e = min(two_streams(d_a, h_src, h_enc[:c], d_b, k) for _ in range(4))
        d = min(two_streams(d_b, h_enc[:c], h_dec, d_a, k) for _ in range(4))
        print(f"copies only, {k:3d} chunks, two streams: encode shape {e * 1e3:.2f} ms, decode shape {d * 1e3:.2f} ms")
    res["enc_chunked"] = min(chunked(d_a, h_src, h_enc[:c], d_b) for _ in range(5))
    res["dec_chunked"] = min(chunked(d_b, h_enc[:c], h_dec, d_a) for _ in range(5))
    print(f"copies only, 16 chunks on 16 streams (in, then home): encode shape {res['enc_chunked'] * 1e3:.2f} ms, "
          f"decode shape {res['dec_chunked'] * 1e3:.2f} ms")
    print(f"n {n >> 20} MiB container {c / n:.4f}")
    print(f"b2rc_encode {min(te) * 1e3:7.2f} ms (median {sorted(te)[3] * 1e3:.2f})   copies only: duplex {res['enc_duplex'] * 1e3:.2f}  "
          f"H2D alone {res['enc_h2d_alone'] * 1e3:.2f}  D2H alone {res['enc_d2h_alone'] * 1e3:.2f}")
    print(f"b2rc_decode {min(td) * 1e3:7.2f} ms (median {sorted(td)[3] * 1e3:.2f})   copies only: duplex {res['dec_duplex'] * 1e3:.2f}  "
          f"H2D alone {res['dec_h2d_alone'] * 1e3:.2f}  D2H alone {res['dec_d2h_alone'] * 1e3:.2f}")


if __name__ == "__main__":
    main()
