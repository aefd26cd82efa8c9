"""rANS (cppans::rANS of the reference) on the block framework: the
CUDA path through the C ABI against the pinned oracle -- bit exact, per block.  `-m gpu`."""
import json

import numpy as np
import pytest

from _cases import crafted_stream
from _oracle import CANTERBURY, GOLDEN, RANS_BYTE, RANS_HEADER, RANS_WORD, Oracle, canterbury, fnv1a64, offsets_of
from cpprcoder_b200 import container, synth

pytestmark = pytest.mark.gpu
MODES = [(RANS_BYTE, "rans_byte"), (RANS_WORD, "rans_word")]
KEY_OF = dict(MODES)


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


@pytest.fixture(scope="module")
def golden_rans():
    return json.loads((GOLDEN / "golden_rans.json").read_text())


def payloads_of(buf):
    info = container.parse(buf)
    return info, [bytes(info.payload(buf, b)) for b in range(info.nblocks)]


def assert_blocks_equal(got, want, what):
    assert len(got) == len(want), what
    for b, (g, w) in enumerate(zip(got, want)):
        if g != w:
            k = next((i for i in range(min(len(g), len(w))) if g[i] != w[i]), min(len(g), len(w)))
            raise AssertionError(f"{what}: block {b} differs at byte {k} (sizes {len(g)} vs {len(w)})")


@pytest.mark.parametrize("name", CANTERBURY)
def test_canterbury_every_block_is_the_reference_payload(ctx, oracle, golden_rans, name):
    data = np.frombuffer(canterbury(name), dtype=np.uint8)
    ent = golden_rans["canterbury"][name]
    for mode, key in MODES:
        enc = ctx.encode(mode, data, 65536)
        info, pays = payloads_of(enc)
        assert (info.mode, info.block, info.total) == (mode, 65536, data.size)
        assert [len(p) for p in pays] == ent["blocks64k"][key]["sizes"]
        assert f"{fnv1a64(b''.join(pays)):016x}" == ent["blocks64k"][key]["cat_fnv"]  # golden = unmodified reference
        assert_blocks_equal(pays, oracle.encode_blocks(mode, data, 65536), f"{name}/{key}")
        assert ctx.decode(enc).tobytes() == data.tobytes()


def test_single_block_equals_whole_file_reference_output(ctx, golden_rans):
    for name in ("alice29.txt", "kennedy.xls", "xargs.1"):
        data = np.frombuffer(canterbury(name), dtype=np.uint8)
        block = 1 << 20 if data.size <= (1 << 20) else 1 << 21
        for mode, key in MODES:
            enc = ctx.encode(mode, data, block)
            _, pays = payloads_of(enc)
            assert len(pays) == 1
            want = golden_rans["canterbury"][name]["whole"][key]
            assert len(pays[0]) == want["size"] and f"{fnv1a64(pays[0]):016x}" == want["fnv"]
            assert ctx.decode(enc).tobytes() == data.tobytes()


@pytest.mark.parametrize("block,nblocks,ragged", [(65536, 70, 12345), (65536, 33, 0), (4096, 200, 1), (16384, 64, 16383),
                                                  (64, 300, 7), (1024, 31, 5), (262144, 5, 99999)])
def test_crafted_streams_match_oracle(ctx, oracle, block, nblocks, ragged):
    data = crafted_stream(nblocks, block, seed=block + nblocks, ragged=ragged)
    for mode, key in MODES:
        enc = ctx.encode(mode, data, block)
        _, pays = payloads_of(enc)
        assert_blocks_equal(pays, oracle.encode_blocks(mode, data, block, threads=4), f"crafted {block}/{key}")
        assert ctx.decode(enc).tobytes() == data.tobytes()


def test_edges(ctx, oracle, golden_rans):
    cases = [b"", b"A", b"AA", b"A" * 7, b"A" * 8, b"A" * 9, b"AB" * 32, b"A" * 65535, b"A" * 65536, b"A" * 65537,
             b"\xff" * 65536, bytes(range(256)) * 3, b"A" * 60000 + bytes(range(256)), b"AB" * 4097]
    for d in cases:
        data = np.frombuffer(d, dtype=np.uint8)
        for mode, key in MODES:
            enc = ctx.encode(mode, data, 65536)
            info, pays = payloads_of(enc)
            assert info.nblocks == (len(d) + 65535) // 65536
            assert_blocks_equal(pays, oracle.encode_blocks(mode, data, 65536), f"edge {len(d)}/{key}")
            assert ctx.decode(enc).tobytes() == d
    # single-block inputs of the golden file: the reference's own bytes
    from test_ans_oracle import EDGE
    for ent in golden_rans["edge"]:
        mode = RANS_BYTE if ent["mode"] == "rans_byte" else RANS_WORD
        d = np.frombuffer(EDGE[ent["label"]], dtype=np.uint8)
        _, pays = payloads_of(ctx.encode(mode, d, 65536))
        assert len(pays[0]) == ent["size"] and f"{fnv1a64(pays[0]):016x}" == ent["fnv"], (ent["label"], ent["mode"])


def test_synthetic_golden_vectors(ctx, golden_rans):
    for ent in golden_rans["synthetic"]:
        mode = RANS_BYTE if ent["mode"] == "rans_byte" else RANS_WORD
        d = synth.GENERATORS[ent["gen"]](ent["n"])
        enc = ctx.encode(mode, d, ent["block"])
        _, pays = payloads_of(enc)
        assert [len(p) for p in pays] == ent["sizes"], (ent["gen"], ent["block"])
        assert f"{fnv1a64(b''.join(pays)):016x}" == ent["cat_fnv"]
        assert ctx.decode(enc).tobytes() == d.tobytes()


def test_decodes_what_the_oracle_encoded(ctx, oracle):
    # decoder alone: containers assembled on the host from oracle payloads
    for block, n in [(65536, 65536 * 9 + 4321), (4096, 4096 * 37 + 5), (64, 64 * 50 + 63)]:
        data = synth.mixed(n)
        for mode, _ in MODES:
            pays = oracle.encode_blocks(mode, data, block, threads=4)
            buf = container.build(mode, block, data.size, pays)
            assert ctx.decode(buf).tobytes() == data.tobytes()


def test_kernel_doors_step_by_step(ctx, oracle):
    import torch
    block = 65536
    data = crafted_stream(41, block, seed=5, ragged=777)
    n = data.size
    nb = (n + block - 1) // block
    src = torch.from_numpy(data).cuda()
    for mode, key in MODES:
        slots, stride, sizes, err = ctx.encode_blocks(mode, src, block)
        # the model kernel left size + normalised cumulative counts at the head of every slot
        head = slots.view(-1, stride)[:nb, :RANS_HEADER].cpu().numpy().copy().view(np.uint32)
        for b in (0, 1, nb // 2, nb - 1):
            blk = data[b * block:(b + 1) * block]
            _, cum = oracle.rans_model(blk, 12 if mode == RANS_WORD else 14)
            assert head[b][0] == blk.size and (head[b][1:] == cum).all(), f"model of block {b}"
        offsets = ctx.scan(sizes, nb)
        total = int(offsets[nb].item())
        payload = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
        lead = 2 if mode == RANS_WORD else 3                                     # word payloads need even addresses
        ctx.compact(slots, stride, sizes, offsets, nb, payload[lead:], err, mode)
        want = oracle.encode_blocks(mode, data, block, threads=4)
        assert total == sum(len(p) for p in want)
        assert payload[lead:lead + total].cpu().numpy().tobytes() == b"".join(want)
        dst = torch.zeros(n, dtype=torch.uint8, device="cuda")
        pl = payload[lead:lead + total + 32]
        ctx.decode_blocks(mode, pl, total, offsets, nb, dst, n, block, err)
        torch.cuda.synchronize()
        assert int(err[0].item()) == 0
        assert dst.cpu().numpy().tobytes() == data.tobytes()


@pytest.mark.parametrize("mode", [RANS_BYTE, RANS_WORD])
def test_corrupt_payloads_are_rejected(ctx, oracle, mode):
    from cpprcoder_b200.api import B2rcError
    data = synth.zipf(65536 * 3 + 100)
    enc = ctx.encode(mode, data, 65536)
    info = container.parse(enc)
    base = info.payload_base + int(info.offsets[1])
    bad = enc.copy()
    bad[base + 4 + 4 * 256 + 1] ^= 0x40          # cum[256] of block 1 is no longer the scale
    with pytest.raises(B2rcError) as e:
        ctx.decode(bad)
    assert e.value.code == -3
    bad = enc.copy()
    bad[base] ^= 1                                # size field disagrees with the container
    with pytest.raises(B2rcError) as e:
        ctx.decode(bad)
    assert e.value.code == -3
    bad = enc.copy()
    bad[base + 4 + 4 * 10] ^= 0xFF                # cumulative counts not monotone
    with pytest.raises(B2rcError) as e:
        ctx.decode(bad)
    assert e.value.code == -3
    # an odd offset in the index: word payloads are 2-byte aligned; for the byte coder the
    # shifted block no longer starts with its size
    bad = enc.copy()
    off = bad[32:32 + 8 * (info.nblocks + 1)].view(np.uint64)
    off[1] += 1
    with pytest.raises(B2rcError) as e:
        ctx.decode(bad)
    assert e.value.code == -3
    # flipped coded bits decode to something, without faulting, and the next call is clean
    bad = enc.copy()
    bad[base + RANS_HEADER + 40: base + RANS_HEADER + 400] ^= 0x5A
    try:
        ctx.decode(bad)
    except B2rcError as e2:
        assert e2.code == -3
    assert ctx.decode(enc).tobytes() == data.tobytes()


@pytest.mark.parametrize("mode", [RANS_BYTE, RANS_WORD])
def test_device_api_and_large_property(ctx, oracle, mode):
    import torch
    n = (256 << 20) + 12345
    data = synth.zipf(n)
    src = torch.from_numpy(data).cuda()
    enc, used = ctx.encode_device(mode, src, block=65536)
    dec = torch.zeros(n, dtype=torch.uint8, device="cuda")
    got = ctx.decode_device(enc, used, dec)
    assert got == n and torch.equal(dec, src)
    host = enc[:used].cpu().numpy()
    info = container.parse(host)
    rng = np.random.default_rng(0)
    for b in [0, info.nblocks - 1] + [int(v) for v in rng.integers(0, info.nblocks, 6)]:
        blk = data[b * 65536:(b + 1) * 65536]
        assert bytes(info.payload(host, b)) == oracle.encode(mode, blk), f"block {b}"
    # a different shape of data through the same context
    d2 = synth.kennedy(64 << 20)
    s2 = torch.from_numpy(d2).cuda()
    e2, u2 = ctx.encode_device(mode, s2, block=16384)
    o2 = torch.zeros(d2.size, dtype=torch.uint8, device="cuda")
    g2 = ctx.decode_device(e2, u2, o2)
    assert g2 == d2.size and torch.equal(o2, s2)


@pytest.mark.parametrize("mode", [0, 1, RANS_BYTE, RANS_WORD])
def test_random_damage_never_faults(ctx, mode):
    """Bytes flipped anywhere in the payload area (the index stays valid): every decode either succeeds
    with some output or reports corruption; nothing faults, and the context keeps working."""
    from cpprcoder_b200.api import B2rcError
    data = synth.mixed((4 << 20) + 777)
    enc = ctx.encode(mode, data, 65536)
    info = container.parse(enc)
    rng = np.random.default_rng(100 + mode)
    outcomes = {"ok": 0, "corrupt": 0}
    for trial in range(12):
        bad = enc.copy()
        k = int(rng.integers(1, 200))
        at = rng.integers(info.payload_base, bad.size, k)
        bad[at] ^= rng.integers(1, 256, k).astype(np.uint8)
        if trial % 3 == 0:   # a burst inside one block's header / model
            b = int(rng.integers(0, info.nblocks))
            lo = info.payload_base + int(info.offsets[b])
            bad[lo:lo + 64] = rng.integers(0, 256, 64, dtype=np.uint8)
        try:
            out = ctx.decode(bad)
            assert out.size == data.size
            outcomes["ok"] += 1
        except B2rcError as e:
            assert e.code == -3, e
            outcomes["corrupt"] += 1
    assert outcomes["corrupt"] >= 1
    assert ctx.decode(enc).tobytes() == data.tobytes()


# ---------------------------------------------------------------- restart points --
def _decode_segment(payload: bytes, x: int, ahead: int, count: int) -> tuple[bytes, int]:
    """rANS::decode's loop (cppans.h:545-560; get :313-316, advance :321-334) from the middle of a payload:
    state `x`, `ahead` coded bytes still unread (they are the LAST `ahead` bytes of the payload)."""
    cum = np.frombuffer(payload[4:4 + 4 * 257], dtype="<u4").astype(np.int64)
    sym_of = np.repeat(np.arange(256), np.diff(cum))
    r = len(payload) - ahead
    out = bytearray()
    for _ in range(count):
        slot = x & 0x3FFF
        s = int(sym_of[slot])
        out.append(s)
        x = int(cum[s + 1] - cum[s]) * (x >> 14) + slot - int(cum[s])
        while x < (1 << 23):
            x = (x << 8) | payload[r]
            r += 1
    return bytes(out), x


@pytest.mark.parametrize("gen,block,extra", [("zipf", 65536, 0), ("kennedy", 65536, 4097), ("mixed", 16384, 77)])
def test_byte_rans_container_carries_restart_points(ctx, oracle, gen, block, extra):
    """Containers of the byte variant end in a table of restart points like the static range coder's: after
    coding everything from symbol 8192 k on, the encoder's state and how many bytes it had emitted.  The
    payloads stay the reference's; every record lets a plain CPU decoder start in the middle of the block
    (and end on the next record's state); the segmented kernel gives the input back."""
    import torch
    from cpprcoder_b200 import container
    n = 12 * block + extra
    data = synth.GENERATORS[gen](n)
    ctx.force_restart(8192)  # (12 blocks on their own would get the points every 1024 symbols: b2rc_restart_for)
    try:
        enc = ctx.encode(RANS_BYTE, data, block)
    finally:
        ctx.force_restart(0)
    info = container.parse(enc)
    assert info.seg_syms == 8192 and info.restart.shape == (info.nblocks, block // 8192 - 1, 3)
    want = oracle.encode_blocks(RANS_BYTE, data, block, threads=4)
    assert_blocks_equal([bytes(info.payload(enc, b)) for b in range(info.nblocks)], want, f"{gen}/{block}")
    for b in (0, info.nblocks - 2, info.nblocks - 1):
        blk = data[b * block:(b + 1) * block]
        recs = info.restart[b]
        for k in (0, recs.shape[0] - 1):
            at = 8192 * (k + 1)
            if at >= blk.size:
                assert recs[k][0] == 0xFFFFFFFF
                continue
            stop = min(at + 8192, blk.size)
            got, x_end = _decode_segment(want[b], int(recs[k][1]), int(recs[k][0]), stop - at)
            assert got == blk[at:stop].tobytes(), f"block {b} segment {k + 1}"
            nxt = int(recs[k + 1][1]) if k + 1 < recs.shape[0] and 8192 * (k + 2) < blk.size else 1 << 23
            assert x_end == nxt
    assert ctx.decode(enc).tobytes() == data.tobytes()
    ctx.force_restart(8192)
    try:
        d_enc, used = ctx.encode_device(RANS_BYTE, torch.from_numpy(data).cuda(), block=block)
    finally:
        ctx.force_restart(0)
    assert used == enc.size and d_enc[:used].cpu().numpy().tobytes() == enc.tobytes()
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert ctx.decode_device(d_enc, used, d_out) == n and d_out.cpu().numpy().tobytes() == data.tobytes()
    # a container without the table (what container.build makes) takes the one-chain kernel as before
    plain = container.build(RANS_BYTE, block, n, want)
    assert container.parse(plain).restart is None
    assert ctx.decode(plain).tobytes() == data.tobytes()
    # a damaged record is an error or a clean decode, never a fault: the state handed on must match
    bad = enc.copy()
    table_at = info.payload_base + ((int(info.offsets[-1]) + 3) & ~3)
    bad[table_at + 4] ^= 0x40  # x of block 0, first record
    with pytest.raises(Exception):
        ctx.decode(bad)
