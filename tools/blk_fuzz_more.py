"""Ten more seeds of the mixed-structure stream through the block sort, device and host calls, against the oracle."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'tests')); sys.path.insert(0, str(ROOT))
import numpy as np, torch
from _cases import blk_fuzz_stream
from _oracle import BlkSort, Oracle
from cpprcoder_b200 import api
ctx=api.Context(0); o=BlkSort(Oracle.get())
bad=0
for seed in range(3,13):
    d=blk_fuzz_stream(seed, nblocks=36)
    want=o.encode(d,threads=16)
    got=ctx.blk_encode_device(torch.from_numpy(d).cuda()).cpu().numpy()
    back=ctx.blk_decode_device(torch.from_numpy(want).cuda()).cpu().numpy()
    ok=np.array_equal(got,want) and np.array_equal(back,d)
    hb=np.array_equal(ctx.blk_decode(ctx.blk_encode(d)), d)
    print(seed, d.size, ok, hb, flush=True); bad+= (not ok) or (not hb)
print('bad',bad)
