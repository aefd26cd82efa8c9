"""Multi-GPU: one process per GPU, blocks sharded by contiguous range, ONE small collective.

Blocks are independent, so rank r codes blocks [floor(r*nb/W), floor((r+1)*nb/W)) of the
stream on its own GPU with no data-path communication.  The only exchange step is an
all-gather of the per-block payload sizes (4 bytes per block) so that every rank holds
the global offset index; payload bytes stay on the GPU that produced them.
Decoding needs no collective at all: the index is replicated.

torch.distributed is the plumbing (NCCL on GPUs; the same code runs over gloo on CPU
tensors, which is how tests/test_dist_gloo.py covers it without a GPU).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import container

RESTART_SYMS = 8192  # B2RC_DEFAULT_RESTART_SYMS (include/b2rc.h)


def shard_of(n_total: int, block: int, rank: int, world: int):
    """(byte_lo, byte_hi, blk_lo, blk_hi) of `rank`'s contiguous share of an n_total-byte stream."""
    nb = container.nblocks_of(n_total, block)
    blk_lo, blk_hi = container.shard_range(nb, rank, world)
    return blk_lo * block, min(blk_hi * block, n_total), blk_lo, blk_hi


def allgather_sizes(local_sizes: torch.Tensor, n_total: int, block: int, group=None) -> torch.Tensor:
    """All ranks' per-block payload sizes, in block order (int64, length nblocks).

    `local_sizes` holds this rank's blocks (int32, on the device NCCL runs on, or on CPU for gloo).
    Shards differ by at most one block, so every rank pads to the same length before the collective."""
    world = dist.get_world_size(group)
    nb = container.nblocks_of(n_total, block)
    counts = [container.shard_range(nb, r, world)[1] - container.shard_range(nb, r, world)[0] for r in range(world)]
    width = max(counts) if counts else 0
    if width == 0:
        return torch.zeros(0, dtype=torch.int64, device=local_sizes.device)
    mine = torch.zeros(width, dtype=torch.int32, device=local_sizes.device)
    mine[:local_sizes.numel()] = local_sizes.to(torch.int32)
    gathered = torch.empty(world * width, dtype=torch.int32, device=local_sizes.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    rows = gathered.view(world, width)
    return torch.cat([rows[r, :counts[r]] for r in range(world)]).to(torch.int64)


def global_offsets(all_sizes: torch.Tensor) -> torch.Tensor:
    """offsets[nblocks+1] (int64) of the stitched container from the gathered sizes."""
    off = torch.zeros(all_sizes.numel() + 1, dtype=torch.int64, device=all_sizes.device)
    if all_sizes.numel():
        off[1:] = torch.cumsum(all_sizes, 0)
    return off


@dataclass
class Shard:
    rank: int
    world: int
    mode: int
    block: int
    n_total: int
    blk_lo: int
    blk_hi: int
    payload: torch.Tensor       # this rank's payloads, back to back (device)
    payload_bytes: int
    local_offsets: torch.Tensor  # int64, blk_hi - blk_lo + 1, relative to `payload`
    offsets: torch.Tensor        # int64, global index, replicated on every rank
    err: torch.Tensor
    restart: torch.Tensor | None = None  # static coder: this rank's restart records (int32, device)
    seg_syms: int = 0

    @property
    def base(self) -> int:
        """Where this rank's payload starts in the stitched payload area."""
        return int(self.offsets[self.blk_lo].item())


def encode_shard(ctx, mode: int, src_shard: torch.Tensor, n_total: int, block: int, group=None) -> Shard:
    """Code this rank's blocks on its GPU (K1, K2, K4 through the C ABI) and all-gather the sizes."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi, blk_lo, blk_hi = shard_of(n_total, block, rank, world)
    assert src_shard.numel() == hi - lo, "src_shard must hold exactly this rank's byte range"
    nb = blk_hi - blk_lo
    # restart points (static coder): each block becomes several independent chains for the decoder
    seg = RESTART_SYMS if mode == 0 and ctx.restart_records(block, RESTART_SYMS) else 0
    restart = None
    if seg:
        restart = torch.empty(max(nb, 1) * ctx.restart_records(block, seg) * 3, dtype=torch.int32, device=src_shard.device)
    slots, stride, sizes, err = ctx.encode_blocks(mode, src_shard, block, restart=restart, seg_syms=seg)
    local_offsets = ctx.scan(sizes, nb)
    all_sizes = allgather_sizes(sizes[:nb], n_total, block, group)  # the one collective
    offsets = global_offsets(all_sizes)
    total_local = int((offsets[blk_hi] - offsets[blk_lo]).item())
    payload = torch.empty(max(total_local, 1) + 16, dtype=torch.uint8, device=src_shard.device)
    ctx.compact(slots, stride, sizes, local_offsets, nb, payload, err, mode)
    return Shard(rank, world, mode, block, n_total, blk_lo, blk_hi, payload, total_local, local_offsets, offsets, err,
                 restart, seg)


def decode_shard(ctx, shard: Shard, dst_shard: torch.Tensor) -> torch.Tensor:
    """Inverse of encode_shard on the same rank; no collective."""
    lo, hi, blk_lo, blk_hi = shard_of(shard.n_total, shard.block, shard.rank, shard.world)
    return ctx.decode_blocks(shard.mode, shard.payload, shard.payload_bytes, shard.local_offsets, blk_hi - blk_lo,
                             dst_shard, hi - lo, shard.block, restart=shard.restart, seg_syms=shard.seg_syms)


def stitch_on_host(shard: Shard, group=None) -> np.ndarray | None:
    """Convenience, not on the timed path: gather every rank's payload to rank 0 and return the
    B2RC container there (None elsewhere)."""
    world, rank = shard.world, shard.rank
    mine = shard.payload[:shard.payload_bytes].cpu().numpy().tobytes()
    recs = shard.restart.cpu().numpy().tobytes() if shard.restart is not None and shard.blk_hi > shard.blk_lo else b""
    parts = [None] * world if rank == 0 else None
    dist.gather_object((mine, recs), parts, dst=0, group=group)
    if rank != 0:
        return None
    nb = shard.offsets.numel() - 1
    head = container.pack_header(shard.mode, shard.block, shard.n_total, nb, shard.seg_syms)
    index = shard.offsets.cpu().numpy().astype(np.uint64).tobytes()
    body = b"".join(p for p, _ in parts)
    if shard.seg_syms:
        body += bytes(-len(body) % 4) + b"".join(r for _, r in parts)
    return np.frombuffer(head + index + body, dtype=np.uint8)
