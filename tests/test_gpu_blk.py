"""GPU parity of the block-sort transform (k_blk_fwd / k_blk_inv behind b2rc_blk_*) against the oracle
(oracle/blk_oracle.c) and the golden vectors of the unmodified reference.  Bar: bit exact."""
import json
from pathlib import Path

import numpy as np
import pytest

from _cases import blk_cases, blk_fuzz_stream, blk_periodic_cases
from _oracle import BLK_BLOCK, BLK_CODED, BlkSort, Oracle, blk_decoded_size, blk_encode_bound, canterbury, fnv1a64
from cpprcoder_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = {c["label"]: c for c in json.loads((Path(__file__).resolve().parent / "golden" / "golden_blk.json").read_text())["cases"]}


@pytest.fixture(scope="module")
def ctx(built):
    built.build_native()
    from cpprcoder_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle(built):
    return BlkSort(Oracle.get())


def _dev(a):
    import torch
    t = torch.empty(max(a.size, 16), dtype=torch.uint8, device="cuda")
    t[:a.size] = torch.from_numpy(np.array(a, dtype=np.uint8, copy=True))
    return t[:a.size]


@pytest.mark.parametrize("label,data", blk_cases() + blk_periodic_cases(), ids=lambda x: x if isinstance(x, str) else "")
def test_forward_equals_reference_and_oracle(ctx, oracle, label, data):
    coded = ctx.blk_encode_device(_dev(data)).cpu().numpy()
    g = GOLDEN[label]
    assert coded.size == g["coded"] == blk_encode_bound(data.size)
    assert f"{fnv1a64(coded):016x}" == g["coded_fnv"], "differs from the unmodified reference's output"
    assert np.array_equal(coded, oracle.encode(data, threads=8))
    flags = ctx.blk_rounds()
    assert flags.size == data.size // BLK_BLOCK
    assert [b for b in range(flags.size) if flags[b] >> 31] == oracle.periodic_blocks(data)


@pytest.mark.parametrize("label,data", blk_cases() + blk_periodic_cases(), ids=lambda x: x if isinstance(x, str) else "")
def test_inverse_of_the_reference_output(ctx, oracle, label, data):
    coded = oracle.encode(data, threads=8)
    back = ctx.blk_decode_device(_dev(coded)).cpu().numpy()
    assert back.size == data.size == blk_decoded_size(coded.size)
    assert np.array_equal(back, data)


@pytest.mark.parametrize("seed", [1, 2])
def test_fuzz_blocks_equal_the_oracle(ctx, oracle, seed):
    data = blk_fuzz_stream(seed)
    want = oracle.encode(data, threads=16)
    got = ctx.blk_encode_device(_dev(data)).cpu().numpy()
    if not np.array_equal(got, want):
        bad = int(np.flatnonzero(got != want)[0]) // BLK_CODED
        raise AssertionError(f"seed {seed}: block {bad} (kind {bad % 6}) differs from the oracle")
    assert np.array_equal(ctx.blk_decode_device(_dev(want)).cpu().numpy(), data)
    r = ctx.blk_rounds()
    assert ((r >> 8) & 0xFF).max() >= 3 and ((r & 0xFF) - ((r >> 8) & 0xFF)).max() >= 3  # both kinds of round ran


def test_many_periodic_blocks_replay_in_batches(ctx, oracle):
    """More blocks with ties than one batch of the replay takes (592): every one gets the reference's row."""
    rng = np.random.default_rng(11)
    a = np.tile(rng.integers(0, 4, 16384, dtype=np.uint8), 2)
    b = np.tile(rng.integers(0, 256, 8, dtype=np.uint8), 4096)
    ra, rb = oracle.encode(a), oracle.encode(b)
    blocks = [a] * 650 + [b] + [a] * 49
    want = np.concatenate([ra] * 650 + [rb] + [ra] * 49)
    got = ctx.blk_encode_device(_dev(np.concatenate(blocks))).cpu().numpy()
    assert np.array_equal(got, want)
    assert (ctx.blk_rounds() >> 31).all()


def test_host_pointer_calls(ctx, oracle):
    rng = np.random.default_rng(5)
    for n in (0, 1, BLK_BLOCK - 1, BLK_BLOCK, 9 * BLK_BLOCK + 4321):
        d = rng.integers(0, 6, n, dtype=np.uint8)
        coded = ctx.blk_encode(d)
        assert np.array_equal(coded, oracle.encode(d, threads=8)), n
        assert np.array_equal(ctx.blk_decode(coded), d), n


def test_any_row_of_a_tie_decodes(ctx, oracle):
    # equal rotations: whichever of them is named, the walk reads the same block
    d = np.tile(np.array([3, 1, 4, 1, 5, 9, 2, 6], np.uint8), BLK_BLOCK // 8)
    coded = oracle.encode(d).copy()
    row = int(coded[BLK_BLOCK]) | int(coded[BLK_BLOCK + 1]) << 8
    run = BLK_BLOCK // 8
    for other in ((row // run) * run, (row // run) * run + run - 1):
        coded[BLK_BLOCK], coded[BLK_BLOCK + 1] = other & 0xFF, other >> 8
        assert np.array_equal(ctx.blk_decode_device(_dev(coded)).cpu().numpy(), d)


def test_errors(ctx, oracle):
    import torch
    from cpprcoder_b200 import _lib
    d = np.random.default_rng(6).integers(0, 256, 2 * BLK_BLOCK, dtype=np.uint8)
    coded = oracle.encode(d).copy()
    coded[BLK_CODED + BLK_BLOCK + 1] |= 0x80  # row number of block 1 >= 32768
    with pytest.raises(_lib.B2rcError) as e:
        ctx.blk_decode_device(_dev(coded))
    assert e.value.code == _lib.E_CORRUPT
    with pytest.raises(_lib.B2rcError) as e:
        ctx.blk_encode_device(_dev(d), dst=torch.empty(2 * BLK_CODED - 1, dtype=torch.uint8, device="cuda"))
    assert e.value.code == _lib.E_DST_SMALL
    with pytest.raises(_lib.B2rcError) as e:
        ctx.blk_encode_device(_dev(np.concatenate([np.zeros(1, np.uint8), d]))[1:])  # misaligned source
    assert e.value.code == _lib.E_ARG


def test_large_stream_by_properties(ctx, oracle):
    """256 MiB of Zipf bytes: round trip on the device, sampled blocks equal the oracle, and the column of
    every block is a permutation of the block (same histogram)."""
    import torch
    from cpprcoder_b200 import synth
    n = 256 << 20
    data = synth.zipf(n)
    src = torch.from_numpy(data).cuda()
    coded = ctx.blk_encode_device(src)
    assert coded.numel() == blk_encode_bound(n)
    back = ctx.blk_decode_device(coded)
    assert torch.equal(back, src)
    host = coded.cpu().numpy()
    for b in (0, 1, 4095, 8190, 8191):
        want = oracle.encode(data[b * BLK_BLOCK:(b + 1) * BLK_BLOCK])
        assert np.array_equal(host[b * BLK_CODED:(b + 1) * BLK_CODED], want), b
    cols = coded.view(-1, BLK_CODED)[:, :BLK_BLOCK]
    assert torch.equal(torch.sort(cols[:64], dim=1).values, torch.sort(src.view(-1, BLK_BLOCK)[:64], dim=1).values)


def test_block_sort_then_coder_in_one_device_call(ctx):
    """b2rc_blkrc_*: BlkSort::encode and a coder over its output in one C-ABI call, nothing visiting the host
    in between (run_zlib_blk's shape, test/main.cpp:944-1002).  Equal to the two calls made one after the other,
    and its payloads are what the reference's coder makes of the reference's block-sort output."""
    import torch
    from _oracle import ADAPTIVE, STATIC, BlkSort, Oracle
    from cpprcoder_b200 import container
    from cpprcoder_b200._lib import B2rcError, E_CORRUPT
    o = Oracle.get()
    data = np.concatenate([np.frombuffer(canterbury("lcet10.txt"), dtype=np.uint8), np.tile(np.arange(16, dtype=np.uint8), 4096),
                           synth.mixed(5 * 32768 + 999)])
    src = torch.from_numpy(data).cuda()
    sorted_ref = BlkSort(o).encode(data, threads=4)
    for mode in (STATIC, ADAPTIVE):
        enc, used = ctx.blkrc_encode_device(mode, src)
        host = enc[:used].cpu().numpy()
        assert host[:4].tobytes() == b"B2BS" and int(np.frombuffer(host[8:16].tobytes(), dtype="<u8")[0]) == data.size
        inner = host[16:]
        two_calls, used2 = ctx.encode_device(mode, ctx.blk_encode_device(src))
        assert two_calls[:used2].cpu().numpy().tobytes() == inner.tobytes()
        info = container.parse(inner)
        assert info.total == sorted_ref.size
        want = o.encode_blocks(mode, sorted_ref, 65536, threads=4)
        assert [bytes(info.payload(inner, b)) for b in range(info.nblocks)] == want
        dst = torch.empty(data.size, dtype=torch.uint8, device="cuda")
        assert ctx.blkrc_decode_device(enc, used, dst) == data.size
        assert dst.cpu().numpy().tobytes() == data.tobytes()
        bad = enc.clone()
        bad[8] ^= 1  # the size in front no longer fits the container behind it
        with pytest.raises(B2rcError) as e:
            ctx.blkrc_decode_device(bad, used, dst)
        assert e.value.code == E_CORRUPT
