"""Times the host-pointer calls (pinned buffers) separately for encode and decode; scratch tool."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from cpprcoder_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
data = synth.zipf(n)
h_src = torch.from_numpy(data).pin_memory()
h_enc = torch.empty(api.bound(0, n, 65536), dtype=torch.uint8).pin_memory()
h_dec = torch.empty(n, dtype=torch.uint8).pin_memory()
# raw PCIe for reference
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h_src, non_blocking=True)), ("D2H", lambda: h_dec.copy_(d, non_blocking=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name} 1 GiB pinned: {dt*1e3:.1f} ms  {n/dt/1e9:.1f} GB/s")
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h_src, non_blocking=True)
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
with torch.cuda.stream(s2): h_dec.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D+D2H concurrently: {dt*1e3:.1f} ms")
ctx = api.Context(0)
for rep in range(3):
    t0 = time.perf_counter(); out = ctx.encode(0, h_src.numpy(), 65536, dst=h_enc.numpy()); t1 = time.perf_counter()
    dec = ctx.decode(out, dst=h_dec.numpy()); t2 = time.perf_counter()
    print(f"chunks={os.environ.get('B2RC_PIPE_CHUNKS','16')} encode {1e3*(t1-t0):.1f} ms  decode {1e3*(t2-t1):.1f} ms  total {1e3*(t2-t0):.1f} ms  ok={bool((dec[:1000]==data[:1000]).all())}")
