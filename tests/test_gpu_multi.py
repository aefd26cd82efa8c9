"""b2rc_ctx_create_multi: the host-pointer calls sharded over several devices of one process (the path a
C++ caller of the drop-in classes gets N GPUs through).  On a one-GPU box the same device is listed
twice or three times -- the sharding, the stitching and the threads are the same."""
import numpy as np
import pytest

from _cases import crafted_stream
from _oracle import ADAPTIVE, RANS_BYTE, RANS_WORD, STATIC, Oracle
from cpprcoder_b200 import container, synth

pytestmark = pytest.mark.gpu


def _devices(k):
    import torch
    have = torch.cuda.device_count()
    return [d % have for d in range(k)]


@pytest.mark.parametrize("k", [2, 3])
def test_multi_device_container_is_the_single_device_container(built, k):
    from cpprcoder_b200 import api
    built.build_native()
    one = api.Context(0)
    many = api.Context(devices=_devices(k))
    try:
        assert many.lib.b2rc_ctx_devices(many.h) == k
        for mode in (STATIC, ADAPTIVE, RANS_BYTE, RANS_WORD):
            for data, block in ((crafted_stream(41, 65536, seed=3 + mode, ragged=777), 65536),
                                (synth.mixed(23 * 16384 + 5), 16384), (synth.zipf(3 * 65536), 65536)):
                # the sharded call spaces its restart points for ONE device's share of the blocks
                nb = container.nblocks_of(data.size, block)
                one.force_restart(many.restart_for(mode, block, -(-nb // k)) if mode in (STATIC, RANS_BYTE) else 0)
                a = one.encode(mode, data, block)
                b = many.encode(mode, data, block)
                assert a.tobytes() == b.tobytes(), (mode, block, data.size)
                assert many.decode(a).tobytes() == data.tobytes()
                assert one.decode(b).tobytes() == data.tobytes()
        # payloads against the oracle once more, through the sharded path
        data = crafted_stream(50, 4096, seed=77, ragged=1)
        enc = many.encode(STATIC, data, 4096)
        info = container.parse(enc)
        want = Oracle.get().encode_blocks(STATIC, data, 4096, threads=4)
        assert [bytes(info.payload(enc, i)) for i in range(info.nblocks)] == want
    finally:
        one.close()
        many.close()


def test_multi_device_errors_come_back(built):
    from cpprcoder_b200 import api
    from cpprcoder_b200._lib import B2rcError, E_CORRUPT, E_DST_SMALL
    many = api.Context(devices=_devices(2))
    try:
        data = synth.zipf(40 * 65536)
        enc = many.encode(STATIC, data, 65536).copy()
        with pytest.raises(B2rcError) as e:
            many.encode(STATIC, data, 65536, dst=np.empty(enc.size - 1000, np.uint8))
        assert e.value.code == E_DST_SMALL
        bad = enc.copy()
        info = container.parse(enc)
        bad[info.payload_base + int(info.offsets[30])] ^= 0x55    # block 30's size field: the second device's share
        with pytest.raises(B2rcError) as e:
            many.decode(bad)
        assert e.value.code == E_CORRUPT
        assert many.decode(enc).tobytes() == data.tobytes()
    finally:
        many.close()
