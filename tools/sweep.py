"""BASELINE config 5: block-size sweep 4 KiB .. 1 MiB on kennedy.xls-like synthetic data.
Reports device-resident encode / decode GB/s and the compression ratio per block size and coder;
every configuration is round-trip checked, and up to 16 blocks per configuration are compared byte for
byte with the oracle (test infrastructure, used here only as the checker)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch

from _oracle import Oracle
from cpprcoder_b200 import api, container, synth


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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
    out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/block_sweep.json"
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None  # e.g. "rans": refresh those rows of an older file
    ctx = api.Context(0)
    oracle = Oracle.get()
    data = synth.kennedy(n)
    src = torch.from_numpy(data).cuda()
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    rows = []
    for block in [4096, 8192, 16384, 32768, 65536, 131072, 262144, 524288, 1048576]:
        for mode, name in ((0, "static"), (1, "adaptive"), (2, "rans"), (3, "rans-word")):
            if only and name not in only:
                continue
            enc, used = ctx.encode_device(mode, src, block=block)
            t_enc = timed(lambda: ctx.encode_device(mode, src, enc, block=block))
            t_dec = timed(lambda: ctx.decode_device(enc, used, dst))
            ok = bool(torch.equal(dst, src))
            head = enc[:used].cpu().numpy()
            info = container.parse(head)
            picks = sorted({0, info.nblocks // 2, info.nblocks - 1} | {(k * 2654435761) % info.nblocks for k in range(1, 14)})
            exact = all(bytes(info.payload(head, b)) == oracle.encode(mode, data[b * block:(b + 1) * block]) for b in picks)
            rows.append({"block": block, "coder": name, "ratio": used / n, "encode_GBps": n / t_enc / 1e6,
                         "decode_GBps": n / t_dec / 1e6, "round_trip": ok, "sampled_blocks_equal_oracle": exact,
                         "sampled_blocks": len(picks)})
            print(rows[-1], flush=True)
            del enc
    Path(out).write_text(json.dumps({"bytes": n, "generator": "synth.kennedy", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
