import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, "/root/repo")
import torch
from cpprcoder_b200 import api, synth
n = 1 << 30
ctx = api.Context(0)
data = synth.zipf(n)
h_src = torch.from_numpy(data).pin_memory()
h_enc = torch.empty(api.bound(0, n, 65536), dtype=torch.uint8).pin_memory()
h_dec = torch.empty(n, dtype=torch.uint8).pin_memory()
src, enc, dec = h_src.numpy(), h_enc.numpy(), h_dec.numpy()
for _ in range(3):
    out = ctx.encode(0, src, 65536, dst=enc)
    ctx.decode(out, dst=dec)
os.environ["B2RC_TRACE"] = "1"
out = ctx.encode(0, src, 65536, dst=enc)
ctx.decode(out, dst=dec)
