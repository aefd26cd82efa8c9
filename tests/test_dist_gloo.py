"""world_size-2 (and 3) runs of the sharding / index-exchange logic over gloo on CPU tensors.
The coder itself is CUDA-only; here the ORACLE stands in for it (test infrastructure) so the
collective, the shard arithmetic and the stitched container can be checked without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _oracle import ADAPTIVE, RANS_BYTE, RANS_WORD, STATIC, Oracle
from cpprcoder_b200 import container
from cpprcoder_b200 import dist as rcdist
from cpprcoder_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class OracleCtx:
    """Stands in for api.Context on CPU tensors: the same two calls dist.py makes, answered by the oracle
    (test infrastructure -- the product's context launches CUDA kernels and has no CPU path)."""

    def encode_device(self, mode, src, dst=None, block=65536):
        data = src.numpy()
        pays = Oracle.get().encode_blocks(mode, data, block) if data.size else []
        buf = container.build(mode, block, data.size, pays)
        return torch.from_numpy(buf.copy()), buf.size

    def decode_device(self, enc, used, dst):
        buf = enc[:used].numpy()
        info = container.parse(buf)
        if info.total:
            back = Oracle.get().decode_blocks(info.mode, buf[info.payload_base:], info.offsets, info.block, info.total)
            dst[:info.total] = torch.from_numpy(back.copy())
        return info.total


def _worker(rank, world, port, n_total, block, mode, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = synth.kennedy(n_total)
        lo, hi, blk_lo, blk_hi = rcdist.shard_of(n_total, block, rank, world)
        ctx = OracleCtx()
        # odd ranks leave the collective for later (what bench.py does: the decode of a rank's own blocks is launched
        # first); it starts at the first use of the index, and every rank gets there
        shard = rcdist.encode_shard(ctx, mode, torch.from_numpy(data[lo:hi].copy()), n_total, block,
                                    defer_gather=bool(rank & 1))
        assert (shard.gather is None) == bool(rank & 1)
        assert shard.nblocks == blk_hi - blk_lo
        # every rank must now hold the same, complete index
        gathered = [None] * world
        dist.all_gather_object(gathered, shard.offsets.tolist())
        assert all(g == gathered[0] for g in gathered)
        assert shard.payload_bytes == int(shard.local_offsets[-1])
        back = torch.zeros(max(hi - lo, 1), dtype=torch.uint8)
        assert rcdist.decode_shard(ctx, shard, back) == hi - lo
        assert back[:hi - lo].numpy().tobytes() == data[lo:hi].tobytes()
        buf = rcdist.stitch_on_host(shard)
        if rank == 0:
            np.save(os.path.join(out_dir, "stitched.npy"), buf)
        else:
            assert buf is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,block,mode", [(2, 5 * 4096 + 123, 4096, STATIC), (2, 300, 4096, ADAPTIVE),
                                                      (3, 10 * 1024, 1024, ADAPTIVE),
                                                      (4, 2 * 4096 + 5, 4096, STATIC),   # world > nblocks: a rank with no block
                                                      (2, 9 * 4096 + 77, 4096, RANS_WORD), (3, 7 * 1024 + 1, 1024, RANS_BYTE)])
def test_sharded_index_exchange_and_stitching(tmp_path, built, world, n_total, block, mode):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, block, mode, str(tmp_path)), nprocs=world, join=True)
    buf = np.load(tmp_path / "stitched.npy")
    info = container.parse(buf)
    data = synth.kennedy(n_total)
    o = Oracle.get()
    # the stitched container is exactly what a single rank would have produced
    whole = o.encode_blocks(mode, data, block)
    assert [bytes(info.payload(buf, b)) for b in range(info.nblocks)] == whole
    back = o.decode_blocks(mode, buf[info.payload_base:], info.offsets, block, n_total)
    assert back.tobytes() == data.tobytes()


def test_shard_of_covers_the_stream():
    for n_total, block in [(0, 4096), (1, 4096), (4096 * 7 + 5, 4096), (1 << 20, 65536)]:
        for world in (1, 2, 4, 8):
            spans = [rcdist.shard_of(n_total, block, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _blk_worker(rank, world, port, n_total, out_dir):
    """Block sort over ranks: shards at 32 KiB block boundaries, no exchange on the data path (sizes are a
    function of n); rank 0 gathers the pieces only to check them."""
    from _oracle import BLK_BLOCK, BlkSort
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = synth.kennedy(n_total)
        lo, hi, _, _ = rcdist.shard_of(n_total, BLK_BLOCK, rank, world)
        piece = BlkSort(Oracle.get()).encode(data[lo:hi]) if hi > lo else np.zeros(0, np.uint8)
        parts = [None] * world if rank == 0 else None
        dist.gather_object(piece.tobytes(), parts, dst=0)
        if rank == 0:
            np.save(os.path.join(out_dir, "blk.npy"), np.frombuffer(b"".join(parts), dtype=np.uint8))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 5 * 32768 + 999), (3, 4 * 32768), (2, 1000)])
def test_block_sort_shards_concatenate(tmp_path, built, world, n_total):
    from _oracle import BlkSort
    port = _free_port()
    mp.spawn(_blk_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "blk.npy")
    want = BlkSort(Oracle.get()).encode(synth.kennedy(n_total))
    assert np.array_equal(got, want), "shards cut at block boundaries must concatenate to the single-rank output"
