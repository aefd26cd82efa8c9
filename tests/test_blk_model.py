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


def model_tie_replay(s, rk):
    """What k_blk_ties does for a block whose rotations tie: the reference's multikey quicksort (mqsort,
    blksort.h:281-362) replayed range by range -- rank comparisons where the reference compares whole rotations,
    a jump to the first depth at which the smallest and the largest rotation of a range differ, no work at all for
    a range of equal rotations -- and the place rotation 0 ends up in.  `rk`: final ranks (ties share one)."""
    s = [int(x) for x in s]
    rk = [int(x) for x in rk]
    v = list(range(N))

    def byte(row, d):
        return s[(row + d) & M]

    def sift(lo, root, last, x):       # 1-based heap over v[lo ..]
        i = root
        while 2 * i <= last:
            j = 2 * i
            if j < last and rk[v[lo + j - 1]] < rk[v[lo + j]]:
                j += 1
            if not rk[x] < rk[v[lo + j - 1]]:
                break
            v[lo + i - 1] = v[lo + j - 1]
            i = j
        v[lo + i - 1] = x

    queue = [(0, N, 0, 11)]
    while queue:
        nxt = []
        for lo, size, d, level in queue:
            if level == 0:             # heapsort, blksort.h:235-279
                for k in range(size // 2, 0, -1):
                    sift(lo, k, size, v[lo + k - 1])
                last = size
                while last > 1:
                    x = v[lo + last - 1]
                    v[lo + last - 1] = v[lo]
                    last -= 1
                    sift(lo, 1, last, x)
                continue
            if size < 37:              # insertionsort, blksort.h:223-233
                for i in range(1, size):
                    x, j = v[lo + i], i - 1
                    while j >= 0 and rk[x] < rk[v[lo + j]]:
                        v[lo + j + 1] = v[lo + j]
                        j -= 1
                    v[lo + j + 1] = x
                continue
            w = v[lo:lo + size]
            a, b = min(w, key=lambda r: (rk[r], r)), max(w, key=lambda r: (rk[r], r))
            if rk[a] == rk[b]:
                continue               # equal rotations: no pass ever moves one of them
            while byte(a, d) == byte(b, d):
                d += 1                 # passes in which every row shows the same byte move nothing
            q1 = size >> 2
            x0, x1, x2 = s[w[q1]], s[w[2 * q1]], s[w[3 * q1]]      # median on byte 0, whatever the depth
            if x0 < x1:
                pick = w[2 * q1] if x1 < x2 else (w[3 * q1] if x0 < x2 else w[q1])
            else:
                pick = w[q1] if x0 < x2 else (w[3 * q1] if x1 < x2 else w[2 * q1])
            p, hi = byte(pick, d), size - 1
            i0, i1, m0, m1 = 0, hi, 0, hi
            while True:
                while i0 <= i1:
                    c = byte(w[i0], d)
                    if p < c:
                        break
                    if p == c:
                        w[i0], w[m0] = w[m0], w[i0]
                        m0 += 1
                    i0 += 1
                while i0 <= i1:
                    c = byte(w[i1], d)
                    if c < p:
                        break
                    if p == c:
                        w[i1], w[m1] = w[m1], w[i1]
                        m1 -= 1
                    i1 -= 1
                if i1 < i0:
                    break
                w[i0], w[i1] = w[i1], w[i0]
                i0 += 1
                i1 -= 1
            for i in range(min(m0, i0 - m0)):
                w[i], w[i1 - i] = w[i1 - i], w[i]
            less_n = i0 - m0
            for i in range(min(hi - m1, m1 - i1)):
                w[i0 + i], w[hi - i] = w[hi - i], w[i0 + i]
            gt_at = hi - (m1 - i1) + 1
            v[lo:lo + size] = w
            if less_n >= 2:
                nxt.append((lo, less_n, d, level - 1))
            if size - gt_at >= 2:
                nxt.append((lo + gt_at, size - gt_at, d, level - 1))
            if gt_at - less_n >= 2 and d + 1 < N:
                nxt.append((lo + less_n, gt_at - less_n, d + 1, level))
        queue = nxt
    return v.index(0)


def model_ranks(s):
    """Final ranks of all rotations (ties share the place of their run's first row), the slow sure way."""
    twice = np.concatenate([s, s])
    order = sorted(range(N), key=lambda r: twice[r:r + N].tobytes())
    rk = np.empty(N, np.int64)
    head = 0
    for place, r in enumerate(order):
        if place and twice[r:r + N].tobytes() != twice[order[place - 1]:order[place - 1] + N].tobytes():
            head = place
        rk[r] = head
    return rk


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


@pytest.mark.parametrize("label", ["period2", "period4", "period1024", "period16384", "period8192-alpha3"])
def test_tie_replay_model_names_the_reference_row(oracle, label):
    s = dict(blk_periodic_cases())[label]
    rec = oracle.encode(s)
    _, canonical, _ = model_forward(s)
    rk = model_ranks(s)
    assert canonical == rk[0]
    assert model_tie_replay(s, rk) == int(rec[N]) | int(rec[N + 1]) << 8
