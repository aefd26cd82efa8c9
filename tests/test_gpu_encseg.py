"""The static encoder as many chains per block (cpprcoder_b200/csrc/b2rc_encseg.cuh: range pass,
segments coded straight into the container, seams), through the C ABI against the oracle, and
against the one-chain kernel it replaces (same containers, byte for byte, restart table included)."""
import os

import numpy as np
import pytest

from _cases import crafted, crafted_stream
from _oracle import STATIC, Oracle
from cpprcoder_b200 import container, synth

pytestmark = pytest.mark.gpu


def make_ctx(**env):
    """A context created under the given B2RC_* environment (read once, at creation)."""
    from cpprcoder_b200 import api
    saved = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            os.environ[k] = str(v)
        return api.Context(0)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.fixture(scope="module")
def oracle(built):
    built.build_native()
    return Oracle.get()


@pytest.fixture(scope="module")
def one_chain(oracle):
    c = make_ctx(B2RC_ENC_SEG_SYMS=0)  # k_enc_static: one chain per block, staging slots, compaction
    yield c
    c.close()


def payloads(buf):
    info = container.parse(buf)
    return [bytes(info.payload(buf, b)) for b in range(info.nblocks)]


@pytest.mark.parametrize("P", [64, 512, 2048, 8192])
@pytest.mark.parametrize("block,nblocks,ragged", [(65536, 37, 4321), (65536, 32, 0), (4096, 70, 1), (16384, 33, 16383),
                                                  (64, 100, 7), (262144, 5, 70000)])
def test_segments_give_the_reference_payloads(oracle, one_chain, P, block, nblocks, ragged):
    import torch
    data = crafted_stream(nblocks, block, seed=7 * block + P, ragged=ragged)
    want = oracle.encode_blocks(STATIC, data, block, threads=4)
    ctx = make_ctx(B2RC_ENC_SEG_SYMS=P)
    # the encoder's segments are never longer than the spacing of the restart points, which for these few blocks
    # would be 1024 (b2rc_restart_for): the long segments are tested at the spacing a full GPU gets
    spacing = 8192 if P >= 2048 else 0
    ctx.force_restart(spacing)
    one_chain.force_restart(spacing)
    try:
        host = ctx.encode(STATIC, data, block)                       # chunked host pipeline
        assert payloads(host) == want
        enc, used = ctx.encode_device(STATIC, torch.from_numpy(data).cuda(), block=block)
        dev = enc[:used].cpu().numpy()
        assert dev.tobytes() == host.tobytes()
        # the same container as the one-chain kernel writes, restart points and all
        assert one_chain.encode(STATIC, data, block).tobytes() == host.tobytes()
        assert ctx.decode(host).tobytes() == data.tobytes()
    finally:
        one_chain.force_restart(0)
        ctx.close()


def test_range_pass_with_one_two_three_warps_writes_the_same_container(oracle):
    """k_enc_ranges (one warp per 32 blocks) and k_enc_ranges2<2>, <3> (a chain warp and one or two helper warps,
    the form small streams of 64 KiB blocks take): same containers byte for byte, payloads the reference's.  The
    streams hold blocks of one repeated byte (total 0x8000: that CTA falls back to the one-warp path), a ragged
    last block, a last warp with missing lanes, and restart spacings from 1024 to 8192."""
    import torch
    streams = [crafted_stream(37, 65536, seed=91, ragged=4321), synth.zipf(70 * 65536), synth.mixed(64 * 65536),
               synth.kennedy(33 * 65536 + 1), np.concatenate([synth.zipf(3 * 65536), np.zeros(65536, np.uint8)] * 9)]
    ctxs = {w: make_ctx(B2RC_RANGES_WARPS=w) for w in (1, 2, 3)}
    try:
        for k, data in enumerate(streams):
            want = oracle.encode_blocks(STATIC, data, 65536, threads=4)
            for spacing in (0, 8192):
                got = {}
                for w, ctx in ctxs.items():
                    ctx.force_restart(spacing)
                    enc, used = ctx.encode_device(STATIC, torch.from_numpy(data).cuda(), block=65536)
                    got[w] = enc[:used].cpu().numpy()
                    assert payloads(got[w]) == want, (k, spacing, w)
                    assert ctx.encode(STATIC, data, 65536).tobytes() == got[w].tobytes(), (k, spacing, w)
                assert got[1].tobytes() == got[2].tobytes() == got[3].tobytes(), (k, spacing)
                assert ctxs[3].decode(got[3]).tobytes() == data.tobytes()
    finally:
        for ctx in ctxs.values():
            ctx.close()


@pytest.mark.parametrize("mode_name", ["static", "adaptive", "rans", "rans-word"])
def test_host_pipeline_with_many_uneven_chunks(oracle, mode_name):
    """b2rc_encode / b2rc_decode cut a stream into up to 16 chunks, shorter at both ends (plan_chunks).  With the
    smallest chunk size a test may ask for, a 40 MiB stream goes through the same plan a 1 GiB stream gets: the
    container must be the device call's, byte for byte, with equal chunks, with uneven ones and with one chunk."""
    import torch
    mode = {"static": 0, "adaptive": 1, "rans": 2, "rans-word": 3}[mode_name]
    data = np.concatenate([synth.zipf(23 * (1 << 20) + 4321), synth.mixed(17 * (1 << 20))])
    plans = {"uneven": dict(B2RC_PIPE_MIN_CHUNK=1 << 20), "equal": dict(B2RC_PIPE_MIN_CHUNK=1 << 20, B2RC_PIPE_RAMP=0),
             "seven": dict(B2RC_PIPE_MIN_CHUNK=1 << 20, B2RC_PIPE_CHUNKS=7), "one": {}}
    ctxs = {k: make_ctx(**v) for k, v in plans.items()}
    try:
        for block in (65536, 4096):
            ref, used = ctxs["one"].encode_device(mode, torch.from_numpy(data).cuda(), block=block)
            ref = ref[:used].cpu().numpy()
            for name, ctx in ctxs.items():
                ctx.force_restart(ctxs["one"].restart_for(mode, block, container.nblocks_of(data.size, block)))
                enc = ctx.encode(mode, data, block)
                assert enc.tobytes() == ref.tobytes(), (name, block)
                assert ctx.decode(enc).tobytes() == data.tobytes(), (name, block)
    finally:
        for ctx in ctxs.values():
            ctx.close()


def test_every_alignment_of_the_payloads(oracle):
    """Payload offsets take every residue mod 4 (sizes are data dependent); a canary behind the
    container stays intact and the bound is respected."""
    import torch
    rng = np.random.default_rng(5)
    ctx = make_ctx(B2RC_ENC_SEG_SYMS=256)
    try:
        for it in range(6):
            block = int(rng.choice([1024, 4096, 65536]))
            data = np.concatenate([crafted(int(k) % 7, block, rng) for k in rng.integers(0, 7, 40)] +
                                  [crafted(2, int(rng.integers(1, block)), rng)])
            src = torch.from_numpy(data).cuda()
            from cpprcoder_b200 import api
            bound = api.bound(STATIC, data.size, block)
            dst = torch.full((bound + 64,), 0x5A, dtype=torch.uint8, device="cuda")
            _, used = ctx.encode_device(STATIC, src, dst[:bound], block=block)
            out = dst.cpu().numpy()
            assert (out[used:] == 0x5A).all()
            info = container.parse(out[:used])
            assert {int(o) % 4 for o in info.offsets[:-1]} == {0, 1, 2, 3}
            assert payloads(out[:used]) == oracle.encode_blocks(STATIC, data, block, threads=4)
    finally:
        ctx.close()


def test_reference_shaped_path_of_the_flush_quirk(oracle):
    """low_ == 0xFFFFFFFF at the end of a block (cpprcoder.h:439-451) happens once in 2^32 blocks;
    B2RC_FORCE_EXACT sends every block down that path, which must give the reference's bytes too."""
    data = crafted_stream(35, 65536, seed=11, ragged=999)
    ctx = make_ctx(B2RC_FORCE_EXACT=1)
    plain = make_ctx()
    try:
        enc = ctx.encode(STATIC, data, 65536)
        assert payloads(enc) == oracle.encode_blocks(STATIC, data, 65536, threads=4)
        # the restart points as well: same place, same low, a range with the same range / total
        a, b = container.parse(enc), container.parse(plain.encode(STATIC, data, 65536))
        assert np.array_equal(a.restart[:, :, :2], b.restart[:, :, :2])
        for blk in range(a.nblocks):
            n_b = min(65536, data.size - blk * 65536)
            total = 0x8000 if (n_b == 65536 and len(set(data[blk * 65536:(blk + 1) * 65536].tolist())) == 1) else n_b
            live = a.restart[blk, :, 0] != 0xFFFFFFFF
            assert np.array_equal(a.restart[blk, live, 2] // total, b.restart[blk, live, 2] // total)
        assert plain.decode(enc).tobytes() == data.tobytes()
    finally:
        ctx.close()
        plain.close()


def test_destination_too_small_is_reported_and_respected(oracle):
    import torch
    from cpprcoder_b200._lib import B2rcError, E_DST_SMALL
    data = synth.zipf(40 * 65536)
    src = torch.from_numpy(data).cuda()
    ctx = make_ctx()
    try:
        _, used = ctx.encode_device(STATIC, src)
        dst = torch.full((used + 64,), 0x5A, dtype=torch.uint8, device="cuda")
        cap = used - 100000
        with pytest.raises(B2rcError) as e:
            ctx.encode_device(STATIC, src, dst[:cap])
        assert e.value.code == E_DST_SMALL
        assert (dst[cap:].cpu().numpy() == 0x5A).all()
    finally:
        ctx.close()


def test_full_size_every_block_hash(oracle):
    """256 MiB + a ragged tail: EVERY payload against the oracle's (hashes of all 4097 blocks)."""
    import torch
    from _oracle import fnv1a64
    n = (1 << 28) + 4321
    data = synth.zipf(n)
    ctx = make_ctx()
    try:
        enc, used = ctx.encode_device(STATIC, torch.from_numpy(data).cuda())
        head = enc[:used].cpu().numpy()
        got = payloads(head)
        want = oracle.encode_blocks(STATIC, data, 65536, threads=os.cpu_count() or 4)
        assert len(got) == len(want)
        bad = [b for b in range(len(got)) if got[b] != want[b]]
        assert not bad, bad[:10]
        dst = torch.empty(n, dtype=torch.uint8, device="cuda")
        assert ctx.decode_device(enc, used, dst) == n
        assert torch.equal(dst.cpu(), torch.from_numpy(data))
    finally:
        ctx.close()


# ---------------------------------------------------------- restart records are untrusted --
def test_damaged_restart_records_are_detected(oracle):
    """The static decoder starts a chain at every restart record; a damaged record (place, low or range)
    must end in B2RC_E_CORRUPT, never in silently wrong bytes: every segment has to end where the next
    record stands, the last one with the block's coded bytes used up (k_dec_static_seg)."""
    from cpprcoder_b200._lib import B2rcError, E_CORRUPT
    data = np.concatenate([synth.zipf(20 * 65536), synth.mixed(12 * 65536 + 3000)])
    ctx = make_ctx()
    try:
        ctx.force_restart(8192)  # seven records per block, as a stream that fills the GPU gets them
        enc = ctx.encode(STATIC, data, 65536)
        info = container.parse(enc)
        assert info.restart is not None and info.restart.shape[1] == 7
        table_at = info.payload_base + ((int(info.offsets[-1]) + 3) & ~3)
        rng = np.random.default_rng(77)
        for trial in range(40):
            bad = enc.copy()
            b = int(rng.integers(0, info.nblocks - 1))        # a full block: all seven records are live
            j = int(rng.integers(0, 7))
            w = trial % 3                                      # 0: bytes shifted, 1: low, 2: range
            at = table_at + 4 * ((b * 7 + j) * 3 + w)
            word = int(np.frombuffer(bad[at:at + 4].tobytes(), dtype="<u4")[0])
            if w == 2:   # range: only range / total matters (total = 65536 for these blocks): damage the quotient
                word ^= 1 << int(rng.integers(16, 32))
            elif w == 1:
                word ^= 1 << int(rng.integers(0, 32))
            else:
                word = (word + int(rng.choice([-3, -1, 1, 2, 300]))) & 0xFFFFFFFF
            bad[at:at + 4] = np.frombuffer(np.uint32(word).tobytes(), dtype=np.uint8)
            with pytest.raises(B2rcError) as e:
                ctx.decode(bad)
            assert e.value.code == E_CORRUPT, (trial, b, j, w)
        assert ctx.decode(enc).tobytes() == data.tobytes()
        # low bits of a recorded range do not matter (any range with the same range / total serves)
        ok = enc.copy()
        at = table_at + 4 * ((3 * 7 + 2) * 3 + 2)
        ok[at] ^= 0x55
        assert ctx.decode(ok).tobytes() == data.tobytes()
    finally:
        ctx.close()


def test_last_symbol_and_trailing_absent_symbols(oracle):
    """The 16-bit cumulative table of the segmented decoder stores 65535 where the true value is 65536
    (behind the last symbol that occurs); blocks whose LAST occurring symbol is frequent, rare, 255 or 0
    exercise the rule that replaces the search result there."""
    import torch
    rng = np.random.default_rng(9)
    blocks = []
    for hi in (255, 254, 128, 7, 1, 0):
        for p_hi in (0.9, 0.5, 0.01, 1.0 / 65536):
            sym = rng.integers(0, hi + 1, 65536) if hi else np.zeros(65536, np.int64)
            d = np.where(rng.random(65536) < p_hi, hi, sym).astype(np.uint8)
            d[int(rng.integers(0, 65536))] = hi
            blocks.append(d)
    data = np.concatenate(blocks)
    ctx = make_ctx()
    try:
        enc, used = ctx.encode_device(STATIC, torch.from_numpy(data).cuda())
        assert payloads(enc[:used].cpu().numpy()) == oracle.encode_blocks(STATIC, data, 65536, threads=4)
        dst = torch.empty(data.size, dtype=torch.uint8, device="cuda")
        assert ctx.decode_device(enc, used, dst) == data.size
        assert dst.cpu().numpy().tobytes() == data.tobytes()
    finally:
        ctx.close()
