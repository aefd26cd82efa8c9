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

def shard_of(n_total: int, block: int, rank: int, world: int):
    """(byte_lo, byte_hi, blk_lo, blk_hi) of `rank`'s contiguous share of an n_total-byte stream."""
    nb = container.nblocks_of(n_total, block)
    blk_lo, blk_hi = container.shard_range(nb, rank, world)
    return blk_lo * block, min(blk_hi * block, n_total), blk_lo, blk_hi


class _Gather:
    """An all-gather of payload sizes in flight; result() waits for it and puts the rows in block order."""

    def __init__(self, work, gathered, counts, width):
        self.work, self.gathered, self.counts, self.width = work, gathered, counts, width

    def result(self) -> torch.Tensor:
        if self.work is not None:
            self.work.wait()
            self.work = None
        if self.width == 0 or all(c == self.width for c in self.counts):
            return self.gathered  # equal shares: the rows are the block order already (no kernel launched here)
        rows = self.gathered.view(len(self.counts), self.width)
        return torch.cat([rows[r, :c] for r, c in enumerate(self.counts)])


def allgather_sizes_async(local_sizes: torch.Tensor, n_total: int, block: int, group=None) -> _Gather:
    """Starts the path's one collective -- every rank's per-block payload sizes (4 bytes per block) -- and
    returns at once; nothing on the data path waits for it (decoding a rank's own blocks needs only its own
    index), so it overlaps the decode that follows."""
    world = dist.get_world_size(group)
    nb = container.nblocks_of(n_total, block)
    counts = [container.shard_range(nb, r, world)[1] - container.shard_range(nb, r, world)[0] for r in range(world)]
    width = max(counts) if counts else 0
    if width == 0:
        return _Gather(None, torch.zeros(0, dtype=torch.int64, device=local_sizes.device), counts, 0)
    if local_sizes.numel() == width:
        mine = local_sizes.to(torch.int32).contiguous()
    else:
        mine = torch.zeros(width, dtype=torch.int32, device=local_sizes.device)
        mine[:local_sizes.numel()] = local_sizes.to(torch.int32)
    gathered = torch.empty(world * width, dtype=torch.int32, device=local_sizes.device)
    work = dist.all_gather_into_tensor(gathered, mine, group=group, async_op=True)
    return _Gather(work, gathered, counts, width)


def allgather_sizes(local_sizes: torch.Tensor, n_total: int, block: int, group=None) -> torch.Tensor:
    """All ranks' per-block payload sizes, in block order (int32, length nblocks).

    `local_sizes` holds this rank's blocks (int32, on the device NCCL runs on, or on CPU for gloo).
    Shards differ by at most one block, so every rank pads to the same length before the collective."""
    return allgather_sizes_async(local_sizes, n_total, block, group).result()


def global_offsets(all_sizes: torch.Tensor) -> torch.Tensor:
    """offsets[nblocks+1] (int64) of the stitched container from the gathered sizes."""
    off = torch.zeros(all_sizes.numel() + 1, dtype=torch.int64, device=all_sizes.device)
    if all_sizes.numel():
        torch.cumsum(all_sizes, 0, dtype=torch.int64, out=off[1:])
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
    container: torch.Tensor      # this rank's blocks as a B2RC container of their own (device), b2rc_encode_device's output
    used: int                    # its length
    gather: _Gather | None       # the all-gather of payload sizes, possibly still in flight (None: not started yet)
    group: object = None
    _offsets: torch.Tensor | None = None

    def start_gather(self) -> None:
        """Starts the collective if encode_shard was told to leave it (defer_gather): a caller that decodes its
        own blocks right away launches the decode FIRST and the exchange behind it, so the collective's
        kernel takes no SM away from the decoder (which fills the GPU in exactly one wave) and the host's
        launch work hides behind the decode."""
        if self.gather is None:
            local = self.local_offsets
            self.gather = allgather_sizes_async((local[1:] - local[:-1]).to(torch.int32), self.n_total, self.block, self.group)

    @property
    def offsets(self) -> torch.Tensor:
        """int64, the global index, replicated on every rank (waits for the collective the first time)."""
        if self._offsets is None:
            self.start_gather()
            self._offsets = global_offsets(self.gather.result())
        return self._offsets

    @property
    def nblocks(self) -> int:
        return self.blk_hi - self.blk_lo

    @property
    def local_offsets(self) -> torch.Tensor:
        """int64, nblocks + 1, relative to this rank's payload area (the shard container's own index)."""
        return self.container[container.HEADER:container.HEADER + 8 * (self.nblocks + 1)].view(torch.int64)

    @property
    def payload_bytes(self) -> int:
        return int(self.local_offsets[-1].item())

    @property
    def payload(self) -> torch.Tensor:
        at = container.HEADER + 8 * (self.nblocks + 1)
        return self.container[at:at + self.payload_bytes]

    @property
    def base(self) -> int:
        """Where this rank's payload starts in the stitched payload area."""
        return int(self.offsets[self.blk_lo].item())


def encode_shard(ctx, mode: int, src_shard: torch.Tensor, n_total: int, block: int, group=None,
                 dst: torch.Tensor | None = None, defer_gather: bool = False) -> Shard:
    """Code this rank's blocks on its GPU -- ONE b2rc_encode_device call, the same one a single GPU makes for
    a whole stream -- and all-gather the payload sizes (4 bytes per block), the path's only exchange.
    A rank without blocks (world > nblocks) still enters the collective.  defer_gather: the collective starts
    at Shard.start_gather() / the first use of Shard.offsets instead (EVERY rank must get there)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi, blk_lo, blk_hi = shard_of(n_total, block, rank, world)
    assert src_shard.numel() == hi - lo, "src_shard must hold exactly this rank's byte range"
    # every rank writes its restart points at the spacing a share of ceil(nblocks / world) blocks gets: the shards
    # may be stitched into one container (one spacing), and each is decoded by a GPU of its own
    force = getattr(ctx, "force_restart", None)
    if force is not None:
        nb_all = container.nblocks_of(n_total, block)
        force(ctx.restart_for(mode, block, -(-nb_all // world)))
    try:
        enc, used = ctx.encode_device(mode, src_shard, dst, block)   # raises B2rcError on any device-side error bit
    finally:
        if force is not None:
            force(0)
    shard = Shard(rank, world, mode, block, n_total, blk_lo, blk_hi, enc, used, None, group)
    if not defer_gather:
        shard.start_gather()  # the one collective; Shard.offsets waits for it
    return shard


def decode_shard(ctx, shard: Shard, dst_shard: torch.Tensor) -> int:
    """Inverse of encode_shard on the same rank; no collective.  Returns the bytes decoded."""
    return ctx.decode_device(shard.container, shard.used, dst_shard)


def stitch_on_host(shard: Shard, group=None) -> np.ndarray | None:
    """Convenience, not on the timed path: gather every rank's payload to rank 0 and return the
    B2RC container of the whole stream there (None elsewhere)."""
    world, rank = shard.world, shard.rank
    host = shard.container[:shard.used].cpu().numpy()
    info = container.parse(host)
    mine = host[info.payload_base:info.payload_base + int(info.offsets[-1])].tobytes()
    recs = info.restart.tobytes() if info.restart is not None and shard.nblocks else b""
    parts = [None] * world if rank == 0 else None
    dist.gather_object((mine, recs, info.seg_syms), parts, dst=0, group=group)
    if rank != 0:
        return None
    return stitch(shard.mode, shard.block, shard.n_total, shard.offsets.cpu().numpy(), parts)


def stitch(mode: int, block: int, n_total: int, offsets: np.ndarray, parts) -> np.ndarray:
    """The container of the whole stream from every rank's (payload bytes, restart records, segment length)."""
    nb = len(offsets) - 1
    seg = max((p[2] for p in parts), default=0)
    head = container.pack_header(mode, block, n_total, nb, seg)
    body = b"".join(p[0] for p in parts)
    if seg:
        body += bytes(-len(body) % 4) + b"".join(p[1] for p in parts)
    return np.frombuffer(head + np.asarray(offsets).astype(np.uint64).tobytes() + body, dtype=np.uint8)
