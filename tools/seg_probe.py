"""Segmented static decode from restart points vs the one-chain kernel (scratch tool)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from cpprcoder_b200 import api, synth
from quick_perf import timed

n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
gen = sys.argv[2] if len(sys.argv) > 2 else "zipf"
block = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
ctx = api.Context(0)
data = synth.GENERATORS[gen](n)
src = torch.from_numpy(data).cuda()
nb = api.nblocks(n, block)
freq = ctx.histogram(src, block)
for seg in [int(v) for v in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["0", "32768", "16384", "8192"])]:
    nrec = ctx.restart_records(block, seg) if seg else 0
    restart = torch.empty(max(nb * nrec * 3, 1), dtype=torch.int32, device="cuda") if nrec else None
    slots, stride, sizes, err = ctx.encode_blocks(0, src, block, freq16=freq, restart=restart, seg_syms=seg)
    t_enc = timed(lambda: ctx.encode_blocks(0, src, block, freq16=freq, slots=slots, sizes=sizes, err=err, restart=restart, seg_syms=seg))
    offsets = ctx.scan(sizes, nb)
    total = int(offsets[nb].item())
    payload = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
    ctx.compact(slots, stride, sizes, offsets, nb, payload, err, 0)
    dst = torch.zeros(n, dtype=torch.uint8, device="cuda")
    t_dec = timed(lambda: ctx.decode_blocks(0, payload, total, offsets, nb, dst, n, block, err, restart=restart, seg_syms=seg))
    torch.cuda.synchronize()
    print(f"seg {seg:6d} ({nrec + 1} per block): encode {t_enc:7.3f} ms  decode {t_dec:7.3f} ms  {n / t_dec / 1e6:7.1f} GB/s  "
          f"ok={torch.equal(dst, src)} err={int(err[0])}", flush=True)
