import numpy as np
import pytest

from cpprcoder_b200 import container


def test_round_trip_and_validation():
    pays = [b"abc", b"", b"defgh"]
    with pytest.raises(ValueError):
        container.build(0, 64, 100, pays)  # 100 bytes at block 64 is 2 blocks, not 3
    c = container.build(0, 64, 130, pays)
    info = container.parse(c)
    assert (info.mode, info.block, info.total, info.nblocks) == (0, 64, 130, 3)
    assert [bytes(info.payload(c, b)) for b in range(3)] == pays
    bad = c.copy()
    bad[32 + 8] = 9  # offsets[1] > offsets[2]
    with pytest.raises(ValueError):
        container.parse(bad)
    with pytest.raises(ValueError):
        container.parse(c[:40])


def test_shard_ranges_tile_the_blocks():
    for nb in (0, 1, 7, 49, 16384):
        for world in (1, 2, 3, 4, 8):
            spans = [container.shard_range(nb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_merge_sizes():
    off = container.merge_sizes([[3, 4], [], [5]])
    assert off.tolist() == [0, 3, 7, 12]
    assert container.merge_sizes([]).tolist() == [0]


def test_restart_table_framing():
    # flags = 1 | (seg / 64) << 8; the table sits behind the payloads at the next 4-byte boundary
    block, seg, total = 16384, 4096, 16384 * 2 + 5
    assert container.restart_records(block, seg) == 3
    assert container.restart_records(block, block) == 0 and container.restart_records(block, 100) == 0
    pays = [b"x" * 7, b"yz", b"q" * 10]
    plain = container.build(0, block, total, pays)
    assert container.parse(plain).restart is None and container.parse(plain).seg_syms == 0
    nb = 3
    table = np.arange(nb * 3 * 3, dtype=np.uint32)
    body = plain[32 + 8 * (nb + 1):].tobytes()
    buf = np.frombuffer(container.pack_header(0, block, total, nb, seg) + plain[32:32 + 8 * (nb + 1)].tobytes() + body
                        + bytes(-len(body) % 4) + table.tobytes(), dtype=np.uint8)
    info = container.parse(buf)
    assert info.seg_syms == seg and info.restart.shape == (nb, 3, 3)
    assert (info.restart.reshape(-1) == table).all()
    assert [bytes(info.payload(buf, b)) for b in range(nb)] == pays
    with pytest.raises(ValueError):
        container.parse(buf[:-4])                      # table cut short
    bad = buf.copy()
    bad[6] = 1                                         # adaptive mode with a restart table
    with pytest.raises(ValueError):
        container.parse(bad)


def test_adaptive_restart_points_carry_the_model():
    # the adaptive coder's points: 3 words + 256 counts -- u16 pairs (131 words) for blocks of at most 65536 bytes,
    # u32 (259 words) above
    for block, seg, per in ((65536, 21888, 131), (131072, 43712, 259), (1 << 20, 43712, 259)):
        nb = 2
        total = block + 17
        nrec = container.restart_records(block, seg)
        assert nrec == -(-block // seg) - 1
        pays = [b"abcde", b"fg"]
        plain = container.build(1, block, total, pays)
        body = plain[32 + 8 * (nb + 1):].tobytes()
        table = np.arange(nb * nrec * per, dtype=np.uint32)
        buf = np.frombuffer(container.pack_header(1, block, total, nb, seg) + plain[32:32 + 8 * (nb + 1)].tobytes() + body
                            + bytes(-len(body) % 4) + table.tobytes(), dtype=np.uint8)
        info = container.parse(buf)
        assert info.seg_syms == seg and info.restart.shape == (nb, nrec, per)
        assert (info.restart.reshape(-1) == table).all()
        with pytest.raises(ValueError):
            container.parse(buf[:-4])
