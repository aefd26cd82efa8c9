"""Parity tests proper: the CUDA path, called through the C ABI (libb2rc.so), against the
oracle on the same inputs -- bit exact, per block.  Run on the B200 with `-m gpu`."""
import ctypes as C

import numpy as np
import pytest

from _cases import crafted, crafted_stream
from _oracle import ADAPTIVE, CANTERBURY, RANS_BYTE, RANS_WORD, STATIC, Oracle, canterbury, fnv1a64, offsets_of
from cpprcoder_b200 import container, synth

pytestmark = pytest.mark.gpu
MODES = [(STATIC, "static"), (ADAPTIVE, "adaptive")]


@pytest.fixture(scope="module")
def ctx(built):
    import torch
    from cpprcoder_b200 import api
    built.build_native()
    assert torch.cuda.is_available()
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle(built):
    return Oracle.get()


def payloads_of(buf: np.ndarray):
    info = container.parse(buf)
    return info, [bytes(info.payload(buf, b)) for b in range(info.nblocks)]


def assert_blocks_equal(got, want, what):
    assert len(got) == len(want), what
    for b, (g, w) in enumerate(zip(got, want)):
        if g != w:
            k = next((i for i in range(min(len(g), len(w))) if g[i] != w[i]), min(len(g), len(w)))
            raise AssertionError(f"{what}: block {b} differs at byte {k} (sizes {len(g)} vs {len(w)})")


# ------------------------------------------------------------------ config 1 + 2 --
@pytest.mark.parametrize("name", CANTERBURY)
def test_canterbury_every_block_is_the_reference_payload(ctx, oracle, golden, name):
    data = np.frombuffer(canterbury(name), dtype=np.uint8)
    ent = golden["canterbury"][name]
    for mode, key in MODES:
        enc = ctx.encode(mode, data, 65536)
        info, pays = payloads_of(enc)
        assert (info.mode, info.block, info.total) == (mode, 65536, data.size)
        assert [len(p) for p in pays] == ent["blocks64k"][key]["sizes"]
        assert [f"{fnv1a64(p):016x}" for p in pays] == ent["blocks64k"][key]["fnv"]  # golden = unmodified reference
        assert_blocks_equal(pays, oracle.encode_blocks(mode, data, 65536), f"{name}/{key}")
        assert ctx.decode(enc).tobytes() == data.tobytes()


def test_alice29_single_block_equals_whole_file_reference_output(ctx, golden):
    # config 1: one block >= the file => the payload IS the reference's whole-file stream (87380 B, ratio 0.574532)
    data = np.frombuffer(canterbury("alice29.txt"), dtype=np.uint8)
    for mode, key in MODES:
        enc = ctx.encode(mode, data, 1 << 18)
        _, pays = payloads_of(enc)
        assert len(pays) == 1
        want = golden["canterbury"]["alice29.txt"]["whole"][key]
        assert len(pays[0]) == want["size"] and f"{fnv1a64(pays[0]):016x}" == want["fnv"]
        assert ctx.decode(enc).tobytes() == data.tobytes()


# ------------------------------------------------------------- crafted / ragged --
@pytest.mark.parametrize("block,nblocks,ragged", [(65536, 70, 12345), (65536, 33, 0), (4096, 200, 1), (16384, 64, 16383),
                                                  (64, 300, 7), (1024, 31, 0)])
def test_crafted_streams_match_oracle(ctx, oracle, block, nblocks, ragged):
    data = crafted_stream(nblocks, block, seed=block + nblocks, ragged=ragged)
    for mode, key in MODES:
        enc = ctx.encode(mode, data, block)
        _, pays = payloads_of(enc)
        assert_blocks_equal(pays, oracle.encode_blocks(mode, data, block, threads=4), f"crafted {block}/{key}")
        assert ctx.decode(enc).tobytes() == data.tobytes()


@pytest.mark.parametrize("block", [131072, 262144, 1048576])
def test_blocks_above_64k_follow_the_order_dependent_count(ctx, oracle, block):
    # kennedy-like data: 44 % zeros => RangeEncoder::count halves (cpprcoder.h:549-555) inside every block
    data = synth.kennedy(3 * block + 4097)
    _, events = oracle.static_count(data[:block])
    assert events >= 1 or block < 262144  # 44 % of 128 KiB is still below 0xFFFF
    parts = [data, np.concatenate([crafted(5, block, np.random.default_rng(3)), crafted(6, block // 2, np.random.default_rng(4))])]
    for d in parts:
        for mode, key in MODES:
            enc = ctx.encode(mode, d, block)
            _, pays = payloads_of(enc)
            assert_blocks_equal(pays, oracle.encode_blocks(mode, d, block, threads=4), f"wide {block}/{key}")
            assert ctx.decode(enc).tobytes() == d.tobytes()


@pytest.mark.parametrize("block", [131072, 524288])
def test_wide_histogram_is_the_order_dependent_count(ctx, oracle, block):
    # K1 for blocks above 64 KiB: segments that cannot reach 0xFFFF are added in parallel, the
    # others replayed with the reference's halving rule (cpprcoder.h:549-555)
    import torch
    rng = np.random.default_rng(block)
    parts = [np.zeros(block, np.uint8),                                   # a halving every 32 Ki bytes
             synth.kennedy(block),                                        # 44 % zeros
             np.where(rng.random(block) < 0.7, 7, rng.integers(0, 256, block)).astype(np.uint8),
             rng.integers(0, 4, block // 2 + 4097, dtype=np.uint8)]      # short last block
    data = np.concatenate(parts)
    src = torch.from_numpy(data).cuda()
    f16 = ctx.histogram(src, block).cpu().numpy().view(np.uint16)
    events = 0
    for b in range((data.size + block - 1) // block):
        want, ev = oracle.static_count(data[b * block:(b + 1) * block])
        events += ev
        assert (f16[b].astype(np.uint32) == want).all(), f"histogram of block {b}"
    assert events >= 3


def test_edges(ctx, oracle):
    for d in [b"", b"A", b"AB" * 32, b"A" * 65535, b"A" * 65536, b"A" * 65537, b"\xff" * 65536, bytes(range(256)) * 3]:
        data = np.frombuffer(d, dtype=np.uint8)
        for mode, key in MODES:
            enc = ctx.encode(mode, data, 65536)
            info, pays = payloads_of(enc)
            assert info.nblocks == (len(d) + 65535) // 65536  # never an empty block
            assert_blocks_equal(pays, oracle.encode_blocks(mode, data, 65536), f"edge {len(d)}/{key}")
            assert ctx.decode(enc).tobytes() == d


def test_synthetic_golden_vectors(ctx, golden):
    for ent in golden["synthetic"]:
        d = synth.GENERATORS[ent["gen"]](ent["n"])
        mode = STATIC if ent["mode"] == "static" else ADAPTIVE
        enc = ctx.encode(mode, d, ent["block"])
        _, pays = payloads_of(enc)
        assert [len(p) for p in pays] == ent["sizes"], (ent["gen"], ent["block"], ent["mode"])
        assert f"{fnv1a64(b''.join(pays)):016x}" == ent["cat_fnv"]


# ----------------------------------------------------------------- kernel doors --
def test_kernel_doors_step_by_step(ctx, oracle):
    import torch
    block = 65536
    data = crafted_stream(40, block, seed=99, ragged=777)
    n = data.size
    nb = (n + block - 1) // block
    src = torch.from_numpy(data).cuda()
    # K1 against RangeEncoder::count
    f16 = ctx.histogram(src, block).cpu().numpy().view(np.uint16)
    for b in range(nb):
        want, _ = oracle.static_count(data[b * block:(b + 1) * block])
        assert (f16[b].astype(np.uint32) == want).all(), f"histogram of block {b}"
    for mode, key in MODES:
        slots, stride, sizes, err = ctx.encode_blocks(mode, src, block)
        offsets = ctx.scan(sizes, nb)
        total = int(offsets[nb].item())
        want = oracle.encode_blocks(mode, data, block, threads=4)
        assert sizes[:nb].cpu().tolist() == [len(p) for p in want]
        assert offsets.cpu().tolist() == offsets_of(want).tolist()
        payload = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
        # unaligned destination: compaction must cope with any byte offset
        ctx.compact(slots, stride, sizes, offsets, nb, payload[3:], err)
        assert int(err[0].item()) == 0
        assert payload[3:3 + total].cpu().numpy().tobytes() == b"".join(want)
        dst = torch.zeros(n, dtype=torch.uint8, device="cuda")
        derr = ctx.decode_blocks(mode, payload[3:], total, offsets, nb, dst, n, block)
        assert int(derr[0].item()) == 0
        assert dst.cpu().numpy().tobytes() == data.tobytes()


# ------------------------------------------------------ full size, by properties --
@pytest.mark.parametrize("gen,mode", [("zipf", STATIC), ("mixed", ADAPTIVE)])
def test_full_size_round_trip_and_sampled_blocks(ctx, oracle, gen, mode):
    import torch
    n = (1 << 28) + 4321  # 256 MiB + ragged tail, device resident
    data = synth.GENERATORS[gen](n)
    src = torch.from_numpy(data).cuda()
    enc, used = ctx.encode_device(mode, src)
    head = enc[:used].cpu().numpy()
    info = container.parse(head)
    assert info.total == n and info.nblocks == (n + 65535) // 65536
    sizes = np.diff(info.offsets.astype(np.int64))
    hdr = 521 if mode == STATIC else 9
    assert sizes.min() >= hdr and sizes.max() <= 65536 + 65536 // 8 + 1024
    for b in list(range(0, info.nblocks, 257)) + [info.nblocks - 1]:  # sampled blocks, byte exact
        want = oracle.encode(mode, data[b * 65536:(b + 1) * 65536])
        assert bytes(info.payload(head, b)) == want, f"block {b}"
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert ctx.decode_device(enc, used, dst) == n
    assert torch.equal(dst, src)


# ----------------------------------------------------------------------- errors --
def test_error_paths(ctx):
    import torch
    from cpprcoder_b200._lib import B2rcError, E_ARG, E_CORRUPT, E_DST_SMALL
    data = synth.zipf(200000)
    enc = ctx.encode(STATIC, data, 65536).copy()
    with pytest.raises(B2rcError) as e:
        ctx.decode(enc, dst=np.empty(1000, np.uint8))
    assert e.value.code == E_DST_SMALL
    bad = enc.copy()
    bad[0] ^= 0xFF
    with pytest.raises(B2rcError) as e:
        ctx.decode(bad)
    assert e.value.code == E_CORRUPT
    bad = enc.copy()
    bad[32 + 8:32 + 16] = 0xFF  # offsets[1] far outside the container
    with pytest.raises(B2rcError) as e:
        ctx.decode(bad)
    assert e.value.code == E_CORRUPT
    bad = enc.copy()
    base = 32 + 8 * 5
    bad[base] ^= 1  # block 0's own size field no longer matches the container
    with pytest.raises(B2rcError) as e:
        ctx.decode(bad)
    assert e.value.code == E_CORRUPT
    # zeroed tail (the restart table and the last payload bytes): must neither hang nor fault; the
    # segmented decoder sees that its chains do not end on the records and says so
    bad = enc.copy()
    bad[-2000:] = 0
    with pytest.raises(B2rcError) as e:
        ctx.decode(bad)
    assert e.value.code == E_CORRUPT
    # the same damage in a container without restart points decodes to garbage of the right length:
    # the reference's stream has nothing to check it against (cpprcoder.h:494-517)
    info = container.parse(enc)
    plain = container.build(STATIC, 65536, data.size, [bytes(info.payload(enc, b)) for b in range(info.nblocks)])
    plain[-300:] = 0
    assert ctx.decode(plain).size == data.size
    with pytest.raises(B2rcError) as e:
        ctx.encode(STATIC, data, 1000)  # block size not a multiple of 64
    assert e.value.code == E_ARG
    src = torch.from_numpy(data).cuda()
    with pytest.raises(B2rcError) as e:
        ctx.encode_device(STATIC, src[1:])  # misaligned device pointer
    assert e.value.code == E_ARG
    small = torch.empty(64, dtype=torch.uint8, device="cuda")
    with pytest.raises(B2rcError) as e:
        ctx.encode_device(STATIC, src, small)
    assert e.value.code == E_DST_SMALL


# ------------------------------------------------------------- more coverage --
def test_random_shapes_match_oracle(ctx, oracle):
    """Seeded fuzz over block sizes (narrow and wide tables), lengths and byte patterns."""
    rng = np.random.default_rng(20261018)
    for it in range(14):
        block = int(rng.choice([64, 128, 1024, 4096, 16384, 65536, 131072])) if it % 3 else 64 * int(rng.integers(1, 1500))
        nblocks = int(rng.integers(1, 80 if block <= 16384 else 40))
        ragged = int(rng.integers(0, block))
        data = crafted_stream(nblocks, block, seed=1000 + it, ragged=ragged)
        for mode, key in MODES:
            enc = ctx.encode(mode, data, block)
            _, pays = payloads_of(enc)
            assert_blocks_equal(pays, oracle.encode_blocks(mode, data, block, threads=4), f"fuzz {it} block {block}/{key}")
            assert ctx.decode(enc).tobytes() == data.tobytes()


def test_side_stream_and_second_context(ctx, oracle):
    import torch
    from cpprcoder_b200 import api
    data = synth.mixed(20 * 65536 + 99, start=1 << 20)
    src = torch.from_numpy(data).cuda()
    other = api.Context(0)
    side = torch.cuda.Stream()
    try:
        with torch.cuda.stream(side):
            enc, used = other.encode_device(ADAPTIVE, src)
            dst = torch.empty(data.size, dtype=torch.uint8, device="cuda")
            assert other.decode_device(enc, used, dst) == data.size
        side.synchronize()
        assert torch.equal(dst, src)
        enc2, used2 = ctx.encode_device(ADAPTIVE, src)  # same answer from the other context on the default stream
        assert used2 == used and torch.equal(enc2[:used2], enc[:used])
        head = enc[:used].cpu().numpy()
        info = container.parse(head)
        assert bytes(info.payload(head, 7)) == oracle.encode(ADAPTIVE, data[7 * 65536:8 * 65536])
    finally:
        other.close()


def test_pipelined_host_calls_on_many_chunks(ctx, oracle):
    """Large enough for the host-pointer pipeline to cut several chunks (64 MiB each)."""
    n = (3 << 26) + 65536 * 5 + 17
    data = synth.kennedy(n)
    for mode, key in MODES:
        enc = ctx.encode(mode, data, 65536)
        info = container.parse(enc)
        assert info.nblocks == (n + 65535) // 65536
        for b in [0, 1023, 1024, 2047, 2048, info.nblocks - 1]:  # around the chunk seams
            assert bytes(info.payload(enc, b)) == oracle.encode(mode, data[b * 65536:(b + 1) * 65536]), f"block {b}"
        assert ctx.decode(enc).tobytes() == data.tobytes()


@pytest.mark.parametrize("mode", [STATIC, ADAPTIVE])
@pytest.mark.parametrize("gen,block,extra", [("zipf", 65536, 0), ("kennedy", 32768, 4097), ("mixed", 131072, 65536 + 5),
                                             ("kennedy", 1 << 20, 12345)])
def test_phased_decode_through_the_host_api(ctx, mode, gen, block, extra):
    """b2rc_decode cuts long range-coder blocks into launches of >= 16 Ki symbols per block and ships
    each stripe home while the next launch runs; the coder state (and the adaptive model) is parked
    between launches."""
    n = (2 << 26) + extra                      # two or three pipeline chunks
    data = synth.GENERATORS[gen](n)
    enc = ctx.encode(mode, data, block)
    out = ctx.decode(enc)
    assert out.size == n
    if out.tobytes() != data.tobytes():
        bad = int(np.flatnonzero(out != data)[0])
        raise AssertionError(f"first difference at byte {bad} (block {bad // block}, symbol {bad % block})")
    # and the same container through the one-launch device path
    import torch
    d_enc = torch.from_numpy(enc).cuda()
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert ctx.decode_device(d_enc, enc.size, d_out) == n
    assert d_out.cpu().numpy().tobytes() == data.tobytes()


# ---------------------------------------------------------------- restart points --
def _sim_lib(built):
    lib = C.CDLL(str(built.build_sim()))
    lib.sim_encode_restart.restype = C.c_long
    lib.sim_encode_restart.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p]
    return lib


@pytest.fixture
def spacing(ctx, request):
    """Restart points every `param` symbols (0: the context's own rule, which packs them closer for the few
    blocks of a test stream: b2rc_restart_for)."""
    ctx.force_restart(request.param)
    yield request.param
    ctx.force_restart(0)


@pytest.mark.parametrize("spacing", [8192, 0], indirect=True)
@pytest.mark.parametrize("gen,block,extra", [("zipf", 65536, 0), ("kennedy", 65536, 4097), ("mixed", 16384, 77),
                                             ("kennedy", 262144, 100000)])
def test_static_container_carries_restart_points(ctx, oracle, built, gen, block, extra, spacing):
    """The static coder's containers end in a table of restart points (one per 4096 symbols, closer together
    for a stream of few blocks): the payloads are still the reference's, the records are what the lane code
    computes on the CPU, and the decoder -- which now runs a chain per segment -- gives the input back."""
    n = 20 * block + extra
    data = synth.GENERATORS[gen](n)
    enc = ctx.encode(STATIC, data, block)
    info = container.parse(enc)
    seg = spacing or ctx.restart_for(STATIC, block, info.nblocks)
    assert spacing or seg == 1024                       # 21 blocks are far from filling the GPU
    assert info.seg_syms == seg and info.restart.shape == (info.nblocks, block // seg - 1, 3)
    want = oracle.encode_blocks(STATIC, data, block, threads=4)
    assert_blocks_equal([bytes(info.payload(enc, b)) for b in range(info.nblocks)], want, f"{gen}/{block}")
    assert ctx.decode(enc).tobytes() == data.tobytes()
    if block <= 65536:
        sim = _sim_lib(built)
        nseg = block // seg
        for b in (0, info.nblocks // 2, info.nblocks - 1):
            blk = np.ascontiguousarray(data[b * block:(b + 1) * block])
            out = np.empty(2 * blk.size + 4096, np.uint8)
            rec = np.zeros(3 * (nseg - 1), np.uint32)
            r = sim.sim_encode_restart(blk.ctypes.data_as(C.c_void_p), blk.size, out.ctypes.data_as(C.c_void_p), out.size,
                                       seg, nseg, rec.ctypes.data_as(C.c_void_p))
            assert r > 0
            got, ref = info.restart[b].reshape(-1, 3), rec.reshape(-1, 3)
            reached = ref[:, 0] != 0xFFFFFFFF
            assert (got[reached][:, :2] == ref[reached][:, :2]).all(), f"position / low of block {b}"
            # any range with the same range / total serves the decoder (a power-of-two total keeps only that)
            total = int(np.bincount(blk, minlength=256).clip(max=0x8000 if blk.size == 65536 else None).sum())
            assert (got[reached][:, 2] // total == ref[reached][:, 2] // total).all(), f"range of block {b}"
            assert (got[~reached][:, 0] == 0xFFFFFFFF).all()
    # the device API writes the same container
    import torch
    d_enc, used = ctx.encode_device(STATIC, torch.from_numpy(data).cuda(), block=block)
    assert used == enc.size and d_enc[:used].cpu().numpy().tobytes() == enc.tobytes()
    # and a container without the table (what container.build makes) decodes as before
    plain = container.build(STATIC, block, n, want)
    assert container.parse(plain).restart is None
    assert ctx.decode(plain).tobytes() == data.tobytes()


def test_restart_spacing_follows_the_number_of_blocks(ctx):
    """b2rc_restart_for: the coder's default spacing (static 4096, byte rANS 8192 symbols) for a stream that gives
    the decoder 4096 warps at that spacing, else halved until ceil(nblocks / 32) x segments reaches 4096 warps, not
    below 1024; a forced spacing wins."""
    assert ctx.restart_for(STATIC, 65536, 16384) == 4096          # the static coder's default: sixteen chains per block
    assert ctx.restart_for(STATIC, 65536, 8192) == 4096
    assert ctx.restart_for(STATIC, 65536, 4096) == 2048
    assert ctx.restart_for(STATIC, 65536, 2048) == 1024
    assert ctx.restart_for(STATIC, 65536, 1) == 1024
    assert ctx.restart_for(STATIC, 1 << 20, 1024) == 4096         # 1 GiB of 1 MiB blocks: 32 x 256 warps
    assert ctx.restart_for(STATIC, 1 << 20, 64) == 1024
    assert ctx.restart_for(RANS_BYTE, 65536, 16384) == 8192       # the byte rANS coder's default
    assert ctx.restart_for(RANS_BYTE, 65536, 2048) == 1024
    assert ctx.restart_for(STATIC, 4096, 1 << 18) == 0            # blocks no longer than the spacing: no table
    assert ctx.restart_for(STATIC, 8192, 1 << 17) == 4096
    assert ctx.restart_for(ADAPTIVE, 65536, 5) == 21888          # the adaptive coder's points carry the model: fixed
    assert ctx.restart_for(RANS_WORD, 65536, 5) == 0
    ctx.force_restart(4096)
    try:
        assert ctx.restart_for(STATIC, 65536, 5) == 4096 and ctx.restart_for(STATIC, 65536, 1 << 20) == 4096
        assert ctx.restart_for(ADAPTIVE, 65536, 5) == 21888
    finally:
        ctx.force_restart(0)
    from cpprcoder_b200._lib import B2rcError
    with pytest.raises(B2rcError):
        ctx.force_restart(100)


def test_restart_points_can_be_switched_off(built, oracle, monkeypatch):
    import torch
    from cpprcoder_b200 import api
    monkeypatch.setenv("B2RC_RESTART_SYMS", "0")
    other = api.Context(0)
    try:
        data = synth.zipf(65536 * 9 + 5)
        enc = other.encode(STATIC, data, 65536)
        info = container.parse(enc)
        assert info.restart is None and enc.size == info.payload_base + int(info.offsets[-1])
        assert other.decode(enc).tobytes() == data.tobytes()
    finally:
        other.close()
