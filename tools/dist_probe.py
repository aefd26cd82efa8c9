"""Where a multi-GPU step spends its time (scratch tool): torchrun --nproc-per-node N tools/dist_probe.py"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch, torch.distributed as dist
from cpprcoder_b200 import api, synth, dist as rcdist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_total = (1 << 30) * world; block = 65536
lo, hi, blk_lo, blk_hi = rcdist.shard_of(n_total, block, rank, world)
data = synth.zipf(hi - lo, start=lo)
src = torch.from_numpy(data).cuda()
ctx = api.Context(local)
dec = torch.empty(hi - lo, dtype=torch.uint8, device="cuda")
sizes = torch.ones(blk_hi - blk_lo, dtype=torch.int32, device="cuda")
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(6):
    dist.barrier(); torch.cuda.synchronize()
    e = [ev() for _ in range(6)]
    t0 = time.perf_counter()
    e[0].record()
    slots, stride, szs, err = ctx.encode_blocks(0, src, block)
    e[1].record()
    nb = blk_hi - blk_lo
    local_off = ctx.scan(szs, nb)
    e[2].record()
    alls = rcdist.allgather_sizes(szs[:nb], n_total, block)
    e[3].record()
    offs = rcdist.global_offsets(alls)
    tot = int((offs[blk_hi] - offs[blk_lo]).item())
    payload = torch.empty(tot + 16, dtype=torch.uint8, device="cuda")
    ctx.compact(slots, stride, szs, local_off, nb, payload, err, 0)
    e[4].record()
    ctx.decode_blocks(0, payload, tot, local_off, nb, dec, hi - lo, block)
    e[5].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if rank == 0 and it >= 3:
        names = ["hist+encode", "scan", "allgather", "offsets+alloc+compact", "decode"]
        print(f"world {world} it {it}: " + ", ".join(f"{nm} {e[i].elapsed_time(e[i+1]):.2f}" for i, nm in enumerate(names)) + f", wall {wall*1e3:.2f} ms", flush=True)
dist.destroy_process_group()
