"""Restart points of the ADAPTIVE coder (cpprcoder_b200/csrc/b2rc_adaptseg.cuh): the payloads stay the
reference's, the points behind them let the decoder run several chains per block; points are untrusted."""
import os

import numpy as np
import pytest

from _cases import crafted_stream
from _oracle import ADAPTIVE, CANTERBURY, Oracle, canterbury
from cpprcoder_b200 import container, synth
from test_gpu_encseg import make_ctx, payloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle(built):
    built.build_native()
    return Oracle.get()


@pytest.mark.parametrize("seg", [0, 4096, 16384, 32768])
def test_payloads_are_the_references_and_segments_decode(oracle, seg):
    import torch
    ctx = make_ctx(B2RC_ADAPTIVE_RESTART_SYMS=seg)
    try:
        streams = [crafted_stream(37, 65536, seed=31 + seg, ragged=4321), synth.mixed(40 * 65536 + 17),
                   crafted_stream(70, 16384, seed=5, ragged=100)] + \
                  [np.frombuffer(canterbury(n), dtype=np.uint8) for n in ("kennedy.xls", "ptt5", "alice29.txt")]
        for k, data in enumerate(streams):
            block = 16384 if k == 2 else 65536
            enc = ctx.encode(ADAPTIVE, data, block)
            info = container.parse(enc)
            assert payloads(enc) == oracle.encode_blocks(ADAPTIVE, data, block, threads=4), k
            live = seg and seg < block
            assert (info.seg_syms == seg and info.restart.shape[2] == 131) if live else info.restart is None
            if live:  # the model of every point: the counts of the symbols in front of it
                for b in range(0, info.nblocks, 7):
                    blk = data[b * block:(b + 1) * block]
                    for j in range(info.restart.shape[1]):
                        at = (j + 1) * seg
                        rec = info.restart[b, j]
                        if at >= blk.size:
                            assert rec[0] == 0xFFFFFFFF
                            continue
                        counts = rec[3:].view(np.uint16)
                        assert np.array_equal(counts, np.bincount(blk[:at], minlength=256).astype(np.uint16)), (k, b, j)
            assert ctx.decode(enc).tobytes() == data.tobytes(), k
            src = torch.from_numpy(data).cuda()
            e2, used = ctx.encode_device(ADAPTIVE, src, block=block)
            assert e2[:used].cpu().numpy().tobytes() == enc.tobytes()
            dst = torch.empty(max(data.size, 16), dtype=torch.uint8, device="cuda")
            assert ctx.decode_device(e2, used, dst) == data.size
            assert dst[:data.size].cpu().numpy().tobytes() == data.tobytes()
    finally:
        ctx.close()


def test_containers_with_and_without_points_are_interchangeable(oracle):
    data = synth.mixed(33 * 65536 + 999)
    a, b = make_ctx(B2RC_ADAPTIVE_RESTART_SYMS=0), make_ctx()
    try:
        plain, pointed = a.encode(ADAPTIVE, data, 65536), b.encode(ADAPTIVE, data, 65536)
        assert payloads(plain) == payloads(pointed)
        nrec = container.restart_records(65536, container.parse(pointed).seg_syms)
        assert nrec == 2                                                            # the default: three chains per block
        assert pointed.size - plain.size == 4 * 131 * nrec * 34 + (-plain.size) % 4   # 524 B a point
        for c in (a, b):
            assert c.decode(plain).tobytes() == data.tobytes()
            assert c.decode(pointed).tobytes() == data.tobytes()
    finally:
        a.close()
        b.close()


def test_damaged_points_are_detected(oracle):
    from cpprcoder_b200._lib import B2rcError, E_CORRUPT
    data = np.concatenate([synth.mixed(12 * 65536), synth.zipf(9 * 65536 + 3000)])
    ctx = make_ctx()
    try:
        enc = ctx.encode(ADAPTIVE, data, 65536)
        info = container.parse(enc)
        table_at = info.payload_base + ((int(info.offsets[-1]) + 3) & ~3)
        rng = np.random.default_rng(78)
        nrec = info.restart.shape[1]
        for trial in range(40):
            bad = enc.copy()
            b = int(rng.integers(0, info.nblocks - 1))
            j = int(rng.integers(0, nrec))
            w = trial % 4                                      # 0: bytes shifted, 1: low, 2: range, 3: a count
            word_ix = w if w < 3 else 3 + int(rng.integers(0, 128))
            at = table_at + 4 * ((b * nrec + j) * 131 + word_ix)
            word = int(np.frombuffer(bad[at:at + 4].tobytes(), dtype="<u4")[0])
            if w == 0:
                word = (word + int(rng.choice([-3, -1, 1, 2, 300]))) & 0xFFFFFFFF
            elif w == 3:
                word = (word + int(rng.choice([1, 0x10000, 0x10001, 5]))) & 0xFFFFFFFF
            else:
                word ^= 1 << int(rng.integers(0, 32))
            bad[at:at + 4] = np.frombuffer(np.uint32(word).tobytes(), dtype=np.uint8)
            with pytest.raises(B2rcError) as e:
                ctx.decode(bad)
            assert e.value.code == E_CORRUPT, (trial, b, j, w)
        # counts moved between two symbols (the sum still fits the position): the chain goes astray and
        # does not end on the next point
        bad = enc.copy()
        at = table_at + 4 * ((2 * nrec + 1) * 131 + 3)
        pair = np.frombuffer(bad[at:at + 4].tobytes(), dtype="<u2").copy()
        if pair[0] > 0:
            pair[0] -= 1
            pair[1] += 1
            bad[at:at + 4] = pair.view(np.uint8)
            with pytest.raises(B2rcError) as e:
                ctx.decode(bad)
            assert e.value.code == E_CORRUPT
        assert ctx.decode(enc).tobytes() == data.tobytes()
    finally:
        ctx.close()


def test_two_warp_encoder_writes_the_same_container(oracle):
    """k_enc_adaptive2 (model and coder on a warp each, B2RC_ADAPTIVE_TWO_WARPS=1): measured slower than the
    one-warp kernel and not the default, but it must write the same bytes, restart points included."""
    data = np.concatenate([crafted_stream(37, 65536, seed=13, ragged=4321), synth.mixed(9 * 65536 + 5)])
    a, b = make_ctx(), make_ctx(B2RC_ADAPTIVE_TWO_WARPS=1)
    try:
        one, two = a.encode(ADAPTIVE, data, 65536), b.encode(ADAPTIVE, data, 65536)
        assert one.tobytes() == two.tobytes()
        assert payloads(two) == oracle.encode_blocks(ADAPTIVE, data, 65536, threads=4)
        assert b.decode(two).tobytes() == data.tobytes()
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("block,seg", [(131072, None), (262144, 43712), (1 << 20, 8192), (131072, 0)])
def test_wide_blocks_carry_points_with_32_bit_counts(oracle, block, seg):
    """Blocks above 65536 bytes: a symbol's count no longer fits 16 bits, the points hold 256 u32 counts (259 words)
    and the decoder walks a tree of 32-bit nodes (LeaflessW).  Same contract: the reference's payloads, the model of
    every point = the counts of the symbols in front of it, several chains per block, damage is detected."""
    import torch
    from cpprcoder_b200._lib import B2rcError, E_CORRUPT
    ctx = make_ctx(**({} if seg is None else {"B2RC_ADAPTIVE_RESTART_SYMS_WIDE": seg}))
    want_seg = 43712 if seg is None else seg
    try:
        # one block of a single symbol (its count passes 65535 inside the block), text-like and mixed blocks, a ragged end
        data = np.concatenate([synth.kennedy(2 * block), np.full(block, 0x41, np.uint8), synth.mixed(2 * block + 70001)])
        enc = ctx.encode(ADAPTIVE, data, block)
        info = container.parse(enc)
        assert payloads(enc) == oracle.encode_blocks(ADAPTIVE, data, block, threads=4)
        if not want_seg:
            assert info.restart is None
        else:
            assert info.seg_syms == want_seg and info.restart.shape == (info.nblocks, -(-block // want_seg) - 1, 259)
            for b in range(info.nblocks):
                blk = data[b * block:(b + 1) * block]
                for j in range(info.restart.shape[1]):
                    at = (j + 1) * want_seg
                    rec = info.restart[b, j]
                    if at >= blk.size:
                        assert rec[0] == 0xFFFFFFFF
                        continue
                    assert np.array_equal(rec[3:], np.bincount(blk[:at], minlength=256).astype(np.uint32)), (b, j)
            assert int(info.restart[2, -1, 3 + 0x41]) > 65535
        assert ctx.decode(enc).tobytes() == data.tobytes()
        src = torch.from_numpy(data).cuda()
        e2, used = ctx.encode_device(ADAPTIVE, src, block=block)
        assert e2[:used].cpu().numpy().tobytes() == enc.tobytes()
        dst = torch.empty(data.size, dtype=torch.uint8, device="cuda")
        assert ctx.decode_device(e2, used, dst) == data.size and dst.cpu().numpy().tobytes() == data.tobytes()
        # a container without points (what the reference-shaped path writes) decodes with either context
        plain = container.build(ADAPTIVE, block, data.size, payloads(enc))
        assert ctx.decode(plain).tobytes() == data.tobytes()
        if want_seg:
            table_at = info.payload_base + ((int(info.offsets[-1]) + 3) & ~3)
            nrec = info.restart.shape[1]
            rng = np.random.default_rng(block + want_seg)
            for trial in range(12):
                bad = enc.copy()
                b = int(rng.integers(0, 2))                       # full blocks: every point is live
                j = int(rng.integers(0, nrec))
                w = trial % 4                                      # 0: bytes shifted, 1: low, 2: range, 3: a count
                word_ix = w if w < 3 else 3 + int(rng.integers(0, 256))
                at = table_at + 4 * ((b * nrec + j) * 259 + word_ix)
                word = int(np.frombuffer(bad[at:at + 4].tobytes(), dtype="<u4")[0])
                word = (word + 1) & 0xFFFFFFFF if w in (0, 3) else word ^ (1 << int(rng.integers(0, 32)))
                bad[at:at + 4] = np.frombuffer(np.uint32(word).tobytes(), dtype=np.uint8)
                with pytest.raises(B2rcError) as e:
                    ctx.decode(bad)
                assert e.value.code == E_CORRUPT, (trial, b, j, w)
    finally:
        ctx.close()
