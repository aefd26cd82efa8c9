import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from _oracle import canterbury
from cpprcoder_b200 import api, container
ctx = api.Context(0)
data = np.frombuffer(canterbury("alice29.txt"), dtype=np.uint8)
for mode in (0, 1):
    enc = ctx.encode(mode, data, 65536)
    info = container.parse(enc)
    print("mode", mode, "container", enc.size, info.offsets.tolist())
    # device path
    src = torch.from_numpy(enc.copy()).cuda()
    dst = torch.zeros(data.size, dtype=torch.uint8, device="cuda")
    try:
        n = ctx.decode_device(src, enc.size, dst)
        print("  device decode ok", n, bool((dst.cpu().numpy() == data).all()))
    except Exception as e:
        print("  device decode failed", e)
    # kernel door
    nb = info.nblocks
    offs = torch.from_numpy(info.offsets.astype(np.int64)).cuda()
    pay = src[info.payload_base:]
    pay2 = pay.clone()
    for name, p in (("view", pay), ("clone", pay2)):
        dst.zero_()
        err = ctx.decode_blocks(mode, p, int(info.offsets[-1]), offs, nb, dst, data.size, 65536)
        torch.cuda.synchronize()
        print("  door", name, "err", err.cpu().tolist(), bool((dst.cpu().numpy() == data).all()))
    try:
        out = ctx.decode(enc)
        print("  host decode ok", bool((out == data).all()))
    except Exception as e:
        print("  host decode failed", e)
