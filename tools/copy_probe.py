"""What the box allows host <-> device, kernels aside: every rank moves 1 GiB each way between PINNED host
memory and its GPU with plain cudaMemcpyAsync in 64 MiB pieces on two streams (H2D on one, D2H on the
other, at the same time), all ranks at once.  That is the ceiling of bench.py's `e2e` at N GPUs.
    python tools/copy_probe.py                                   one GPU
    torchrun --nproc-per-node N tools/copy_probe.py              N ranks of one box
Prints one JSON line (rank 0): per-direction and both-directions GB/s, the slowest rank and the sum."""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch
import torch.distributed as dist

from bench import bind_to_gpu_numa


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bind = "--no-bind" not in sys.argv
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if bind else "not bound (--no-bind)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, piece, reps = 1 << 30, 64 << 20, 5
    h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_up.fill_(7)
    d_up = torch.empty(n, dtype=torch.uint8, device=dev)
    d_dn = torch.ones(n, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(up, down):
        sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            for at in range(0, n, piece):
                if up:
                    with torch.cuda.stream(s_up):
                        d_up[at:at + piece].copy_(h_up[at:at + piece], non_blocking=True)
                if down:
                    with torch.cuda.stream(s_dn):
                        h_dn[at:at + piece].copy_(d_dn[at:at + piece], non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()
        mine = time.perf_counter() - t0
        sync()
        t = torch.tensor([mine], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return n * reps / float(t.item()) / 1e9  # GB/s per direction, slowest rank

    run(True, True)
    res = {"h2d_alone": run(True, False), "d2h_alone": run(False, True), "both_at_once_each_way": run(True, True)}
    notes = [None] * world
    if world > 1:
        dist.all_gather_object(notes, numa)
    else:
        notes = [numa]
    if rank == 0:
        print(json.dumps({"probe": "pinned 1 GiB each way per rank, cudaMemcpyAsync in 64 MiB pieces, two streams",
                          "n_gpus": world, "GBps_per_rank_slowest": res,
                          "GBps_all_ranks": {k: v * world for k, v in res.items()},
                          "roundtrip_ceiling_GBps_all_ranks": res["both_at_once_each_way"] * world / 2,
                          "roundtrip_ceiling_note": "a round trip moves the stream in twice (input, container) and out twice "
                                                    "(container, output): uncompressed bytes / (2 x bytes / both-way rate), "
                                                    "for a ratio near 1",
                          "numa": notes}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
