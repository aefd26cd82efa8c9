"""One-off: > 4 GiB through the device API (64-bit sizes and offsets), round trip + sampled blocks vs the oracle."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from _oracle import Oracle
from cpprcoder_b200 import api, container

n = (5 << 30) + 123457
torch.manual_seed(1)
# skewed bytes made on the device: min of two uniform draws
src = torch.minimum(torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda"),
                    torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda"))
ctx = api.Context(0)
o = Oracle.get()
for mode in (0, 1, 2, 3):
    enc, used = ctx.encode_device(mode, src)
    nb = (n + 65535) // 65536
    idx = enc[:32 + 8 * (nb + 1)].cpu().numpy()
    offsets = np.frombuffer(idx[32:].tobytes(), dtype=np.uint64)
    base = 32 + 8 * (nb + 1)
    assert int(offsets[-1]) + base <= used and used > (1 << 32), (used,)  # a restart table may follow the payloads
    for b in (0, 65535, 65536, nb - 2, nb - 1):
        lo, hi = base + int(offsets[b]), base + int(offsets[b + 1])
        blk = src[b * 65536:(b + 1) * 65536].cpu().numpy()
        assert enc[lo:hi].cpu().numpy().tobytes() == o.encode(mode, blk), f"block {b}"
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert ctx.decode_device(enc, used, dst) == n
    assert torch.equal(dst, src)
    print(f"mode {mode}: {n} -> {used} bytes, blocks {nb}: round trip ok, sampled blocks equal the oracle", flush=True)
    del enc, dst
