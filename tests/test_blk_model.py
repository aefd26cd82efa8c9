"""A numpy model of what k_blk_fwd / k_blk_inv compute (cpprcoder_b200/csrc/b2rc_blk.cuh), checked against the
oracle on the CPU: four-byte start, doubling rounds (full, and short once at most half of the rows still share
a bucket), row number = rank of rotation 0; inverse by legs between stations + list ranking, with the doubling
fallback for permutations of several cycles.  The kernels are tested on the GPU (tests/test_gpu_blk.py); this
keeps the ALGORITHM pinned where there is none, step for step in the order the kernels take them."""
import numpy as np
import pytest

from _cases import blk_fuzz_stream, blk_periodic_cases
from _oracle import BLK_BLOCK, BLK_CODED, BlkSort, Oracle

N, M = BLK_BLOCK, BLK_BLOCK - 1


def stable_by(rows, key):
    return rows[np.argsort(key, kind="stable")]


def model_forward(s):
    """-> (column, row, rounds as a list of 'F' (full) / ('S', active rows))"""
    s = s.astype(np.int64)
    rows = np.arange(N)
    for k in (3, 2, 1, 0):  # the kernel sorts bytes 3 and 2 at once, unstably: rows tied on all four bytes
        #                     come out in another order, which nothing downstream depends on
        rows = stable_by(rows, s[(rows + k) & M])
    key = (s[rows] << 24) | (s[(rows + 1) & M] << 16) | (s[(rows + 2) & M] << 8) | s[(rows + 3) & M]
    head = np.ones(N + 1, bool)
    head[1:N] = key[1:] != key[:-1]
    rk = np.empty(N, np.int64)
    rk[rows] = np.maximum.accumulate(np.where(head[:N], np.arange(N), 0))
    sa, h, log = rows, 4, []
    while head[:N].sum() < N and h < N:
        active = N - int((head[:N] & head[1:]).sum())
        q = (sa - h) & M                                     # all rows, ordered by their second half
        if active > N // 2:
            sa = stable_by(stable_by(q, rk[q] & 0xFF), rk[stable_by(q, rk[q] & 0xFF)] >> 8)
            key = (rk[sa] << 16) | rk[(sa + h) & M]
            head = np.ones(N + 1, bool)
            head[1:N] = key[1:] != key[:-1]
            new = np.empty(N, np.int64)
            new[sa] = np.maximum.accumulate(np.where(head[:N], np.arange(N), 0))
            rk = new
            log.append("F")
        else:
            lst = q[~head[rk[q] + 1]]                        # rows whose bucket holds two or more
            assert lst.size == active
            lst = stable_by(lst, rk[lst] & 0xFF)
            lst = stable_by(lst, rk[lst] >> 8)
            kb, k2, j = rk[lst], rk[(lst + h) & M], np.arange(lst.size)
            b = np.ones(lst.size, bool)
            b[1:] = kb[1:] != kb[:-1]
            g = b.copy()
            g[1:] |= k2[1:] != k2[:-1]
            jf = np.maximum.accumulate(np.where(b, j, 0))
            jg = np.maximum.accumulate(np.where(g, j, 0))
            sa, rk, head = sa.copy(), rk.copy(), head.copy()
            sa[kb + (j - jf)] = lst
            rk[lst] = kb + (jg - jf)
            head[(kb + (j - jf))[g]] = True
            log.append(("S", active))
        h *= 2
    return s[(sa + M) & M].astype(np.uint8), int(rk[0]), log


def model_inverse(col, top):
    nxt = np.argsort(col, kind="stable")                     # counting_sort, blksort.h:379-402
    p0 = int(nxt[top])
    stations = set(range(0, N, 8)) | {p0}
    legs = {}
    for st in stations:                                      # every leg: from a station to the next one
        q, path = int(nxt[st]), [st]
        while q not in stations and len(path) < 1024:
            path.append(q)
            q = int(nxt[q])
        if q not in stations:
            legs = None
            break
        legs[st] = (path, q)
    out, by_legs = None, False
    if legs is not None:
        order, st = [], p0
        while True:                                          # the list ranking, done the slow way
            order.append(st)
            st = legs[st][1]
            if st == p0:
                break
        walk = [p for st in order for p in legs[st][0]]
        if len(walk) == N:
            out, by_legs = col[np.array(walk)], True
    if out is None:                                          # several cycles: pointer doubling
        walk = np.empty(N, np.int64)
        walk[0] = p0
        jump = nxt.copy()
        for k in range(15):
            walk[1 << k:2 << k] = jump[walk[:1 << k]]
            jump = jump[jump]
        out = col[walk]
    return out, by_legs


@pytest.fixture(scope="module")
def oracle(built):
    return BlkSort(Oracle.get())


def test_model_equals_oracle_on_mixed_blocks(oracle):
    data = blk_fuzz_stream(7, nblocks=12)
    want = oracle.encode(data, threads=8)
    kinds, legged = set(), 0
    for b in range(12):
        s = data[b * N:(b + 1) * N]
        col, row, log = model_forward(s)
        rec = want[b * BLK_CODED:(b + 1) * BLK_CODED]
        assert np.array_equal(col, rec[:N]), f"block {b}: column"
        assert row == int(rec[N]) | int(rec[N + 1]) << 8, f"block {b}: row number"
        kinds |= {x if isinstance(x, str) else x[0] for x in log}
        back, by_legs = model_inverse(rec[:N], row)
        assert np.array_equal(back, s), f"block {b}: inverse"
        legged += by_legs
    assert kinds == {"F", "S"}
    assert legged >= 9  # a `next` that dodges the stations (few distinct bytes in long runs) takes the doubling


def test_model_on_periodic_blocks(oracle):
    """Ties: the column is the oracle's, the row number is the FIRST row of rotation 0's run (the kernel's
    canonical answer before the tie replay), the inverse needs the doubling and accepts any row of the run."""
    for label, s in blk_periodic_cases():
        if s.size != N or label in ("period64", "period256-alpha2", "period4096-alpha2"):
            continue  # a few are enough; the model is slow on long ties
        col, row, log = model_forward(s)
        rec = oracle.encode(s)
        assert np.array_equal(col, rec[:N]), label
        ref_row = int(rec[N]) | int(rec[N + 1]) << 8
        back, by_legs = model_inverse(rec[:N], ref_row)
        assert np.array_equal(back, s), label
        back2, _ = model_inverse(rec[:N], row)
        assert np.array_equal(back2, s), label
        if label != "period1":
            assert not by_legs
