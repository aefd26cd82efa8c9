// b2rc_blk.cuh -- the block-sort transform of the reference (blksort::BlkSort, blksort.h) on sm_100a.
// the pre-transform the reference's own pipelines put in front of a coder.
//
//   B1  k_blk_fwd   BlkSort::encode_internal (blksort.h:444-543): the last column of the sorted cyclic
//                   rotations of a 32 KiB block + the row of the unrotated block
//   B2  k_blk_inv   BlkSort::decode_internal (blksort.h:545-672): stable counting sort of the column,
//                   then the walk  p = next[p]  -- 32 768 dependent loads in the reference
//   B3  k_blk_ties  the row number of blocks whose rotations tie: a replay of the reference's quicksort
//
// Mapping: one 32 KiB block per CTA of 1024 threads, everything in shared memory (3 x 64 KiB of u16
// rows/ranks + 16 KiB of counters: the 227 KB of an sm_100 SM is what makes this a one-CTA problem).
//
// Forward: the reference sorts rotations with a multikey quicksort that compares byte by byte
// (O(n * common prefix): 1-2 s for one block of zeros).  Here: prefix doubling.
//   start   rows by their first four bytes: bytes 3 and 2 in one UNSTABLE counting sort on 65 536 shared
//           counters (blk_start16; the first pass of an LSD sort need not be stable), bytes 1 and 0 in two
//           stable passes (blk_pass), then ranks = bucket heads (blk_rerank);
//   rounds  each turns the order by h bytes into the order by 2h.  A full round (more than half of the rows
//           still share a bucket) walks the current order stepping every row h back -- all rows by their
//           SECOND half -- and finishes with two stable passes on the rank of the FIRST half.  A short round
//           does the same for the listed active rows only (blk_list_active, blk_pass_short, blk_place_short);
//           at 2048 active rows or fewer the (bucket, second-half rank) pairs go through a bitonic network
//           (blk_sort_tiny): any sort of the pairs will do.  Rows alone in their bucket never move again.
//   end     all ranks differ (log2 of the longest repeat: 1-2 rounds on Zipf bytes, 6-8 on text), or
//           h = 32 768: the block has a period, equal rotations have equal last bytes, the column does not
//           depend on the tie order; the row number is settled by B3.
// The stable pass is warp-chunked: warp w owns rows [1024 w, 1024 w + 1024) of the input order, counts
// them per digit (ballots find the lanes with equal digits, the first of them adds the group: no atomics)
// and remembers each row's rank within the warp, a CTA scan in (digit, warp) order gives each warp its
// first slot per digit, and every row is then stored at slot + rank.
//
// Inverse: the walk is a linked list through a permutation.  It is cut at stations (every 8th row and the
// start): threads follow the legs between stations, Wyllie's pointer jumping ranks the 4 097 legs, and
// the legs are walked once more, writing.  Permutations with several cycles (blocks with a period) and
// legs longer than 1024 steps fall back to pointer doubling over all rows (J <- J o J: after round k the
// first 2^(k+1) positions of the walk are known), which leaves cycles exactly as the reference's walk does.
#pragma once
#include "b2rc_kernels.cuh"

namespace b2rc
{
constexpr u32 BLK_N = 32768u;      // BlkSort::BlockSize, blksort.h:80
constexpr u32 BLK_M = BLK_N - 1u;
constexpr u32 BLK_CODED = 32770u;  // BlkSort::EncodedSize, blksort.h:83
constexpr u32 BLK_THREADS = 1024u;
constexpr u32 BLK_TINY = 2048u;    // rounds over this many active rows or fewer sort pairs in a bitonic network (4096 measured slower)

// forward kernel, byte offsets in dynamic shared memory
constexpr u32 BF_SA = 0u, BF_TMP = 65536u, BF_RK = 131072u, BF_CNT = 196608u, BF_FB = 212992u, BF_MISC = 217088u;
constexpr u32 BF_LB = BF_MISC + 512u;  // short rounds: two ballot arrays of 512 words
// the start (blk_start16): 65 536 packed u16 counters over sa + tmp, the block's bytes in the first half of rk,
// the rows by their third and fourth byte in what lies behind (second half of rk, cnt, fbits, misc, lb + 7 KiB)
constexpr u32 BF_START_OUT = BF_RK + 32768u, BF_START_X = BF_START_OUT + 65536u;
constexpr u32 BLK_FWD_SMEM = BF_START_X + 512u;
static_assert(BF_LB + 4096u <= BLK_FWD_SMEM && BLK_FWD_SMEM <= 232448u, "forward kernel shared memory");
// inverse kernel: two jump tables, the walk, the column; the counters of the one counting pass sit in the
// second jump table, which is not in use yet
constexpr u32 BI_JA = 0u, BI_JB = 65536u, BI_P = 131072u, BI_L = 196608u, BI_MISC = 229376u;
constexpr u32 BLK_INV_SMEM = BI_MISC + 512u;

__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Exclusive scan of the 32 x 256 per-warp digit counters in (digit, warp) order, in place: afterwards
// cnt[w][d] is the first slot of warp w's rows with digit d.  Called by all threads between two barriers
// of the caller; contains two of its own.
__device__ __forceinline__ void blk_scan_counters(u16* cnt, u32* misc)
{
    const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    // 8 counters per thread
    const u32 d = t >> 2, w0 = (t & 3u) * 8u;
    u32 v[8], sum = 0;
#pragma unroll
    for(u32 k = 0; k < 8u; ++k) {
        v[k] = cnt[(w0 + k) * 256u + d];
        sum += v[k];
    }
    u32 incl = sum;
#pragma unroll
    for(u32 o = 1; o < 32u; o <<= 1) {
        const u32 up = __shfl_up_sync(FULL, incl, o);
        if(lane >= o) {
            incl += up;
        }
    }
    if(lane == 31u) {
        misc[warp] = incl;
    }
    __syncthreads();
    if(warp == 0u) {
        const u32 x = misc[lane];
        u32 xi = x;
#pragma unroll
        for(u32 o = 1; o < 32u; o <<= 1) {
            const u32 up = __shfl_up_sync(FULL, xi, o);
            if(lane >= o) {
                xi += up;
            }
        }
        misc[32u + lane] = xi - x;
    }
    __syncthreads();
    u32 run = misc[32u + warp] + incl - sum;
#pragma unroll
    for(u32 k = 0; k < 8u; ++k) {
        cnt[(w0 + k) * 256u + d] = (u16)run;
        run += v[k];
    }
}

// One stable counting pass over the 32 768 rows of a block on a BITS-bit digit (BITS <= 8).  `elem(i)` is
// the row at place i of the input order, `digit(e)` its digit; rows land in `dst` ordered by digit, input
// order kept within a digit.  cnt: u16[32 warps][256]; misc: u32[64].  Ends with a CTA barrier.
// Warp w owns places [1024 w, 1024 w + 1024).  Counting walks them 32 at a time: the lanes that hold the
// same digit find each other with one ballot per digit bit (match.any costs ~50 cycles of the ADU pipe on
// sm_100 and was 85 % of the first version of this kernel), the first of them adds the group to the warp's
// counter, and every lane keeps (digit, rank among the warp's rows of that digit) in registers -- so the
// placing phase after the CTA scan is one counter read and one store per row, no warp traffic at all.
template <u32 BITS, class Elem, class Digit>
__device__ __forceinline__ void blk_pass(Elem elem, Digit digit, u16* dst, u16* cnt, u32* misc)
{
    const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    u32* c32 = reinterpret_cast<u32*>(cnt);
#pragma unroll
    for(u32 k = 0; k < 4u; ++k) {
        c32[t + k * BLK_THREADS] = 0u;
    }
    __syncthreads();
    u16* mine = cnt + warp * 256u;
    const u32 i0 = warp * 1024u + lane;
    const u32 lt = lanemask_lt();
    u32 dig[8];   // four 8-bit digits per word
    u32 rnk[11];  // three 10-bit ranks per word
#pragma unroll
    for(u32 k = 0; k < 8u; ++k) {
        dig[k] = 0u;
    }
#pragma unroll
    for(u32 k = 0; k < 11u; ++k) {
        rnk[k] = 0u;
    }
#pragma unroll
    for(u32 r = 0; r < 32u; ++r) {
        const u32 d = digit(elem(i0 + r * 32u));
        u32 peers = FULL;
        if(!__all_sync(FULL, d == __shfl_sync(FULL, d, 0))) {  // 32 rows of one run share their digit: no search
#pragma unroll
            for(u32 b = 0; b < BITS; ++b) {
                const u32 bit = (d >> b) & 1u;
                peers &= __ballot_sync(FULL, bit) ^ (bit - 1u);
            }
        }
        const u32 leader = (u32)__ffs((int)peers) - 1u;
        u32 before = 0;
        if(lane == leader) {
            before = mine[d];
            mine[d] = (u16)(before + __popc(peers));
        }
        before = __shfl_sync(FULL, before, leader) + __popc(peers & lt);
        dig[r >> 2] |= d << (8u * (r & 3u));
        rnk[r / 3u] |= before << (10u * (r % 3u));
        __syncwarp();
    }
    __syncthreads();
    blk_scan_counters(cnt, misc);
    __syncthreads();
#pragma unroll
    for(u32 r = 0; r < 32u; ++r) {
        const u32 d = (dig[r >> 2] >> (8u * (r & 3u))) & 0xFFu;
        const u32 at = (u32)mine[d] + ((rnk[r / 3u] >> (10u * (r % 3u))) & 0x3FFu);
        dst[at] = (u16)elem(i0 + r * 32u);
    }
    __syncthreads();
}

// New ranks after a sort: rank = place of the first row of the run of rows that do not differ (bucket
// head).  `keys(e)` gives the pair a row is compared by under the order just established, packed into one
// word; it may read rk (or whatever lives there): nothing is written to rk before every flag is known.
// A row's predecessor is in the lane below, so its keys arrive by shuffle instead of two more gathers.
// fbits: u32[1024]; misc: u32[128], entries 64.. used here.  Returns the number of distinct ranks.
template <class Keys>
__device__ __forceinline__ u32 blk_rerank(const u16* sa, u16* rk, u32* fbits, u32* misc, Keys keys)
{
    const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    const u32 base = warp * 1024u;
    u32 last = 0, heads = 0;
    u32 carry_key = warp ? keys((u32)sa[base - 1u]) : 0u;
#pragma unroll 8
    for(u32 r = 0; r < 32u; ++r) {
        const u32 i = base + r * 32u + lane;
        const u32 key = keys((u32)sa[i]);
        u32 prev = __shfl_up_sync(FULL, key, 1);
        if(lane == 0u) {
            prev = carry_key;
        }
        const bool df = (i == 0u) || key != prev;
        const u32 bits = __ballot_sync(FULL, df);
        if(lane == 0u) {
            fbits[warp * 32u + r] = bits;
        }
        if(bits) {
            last = base + r * 32u + 31u - (u32)__clz((int)bits);
        }
        heads += (u32)__popc(bits);
        carry_key = __shfl_sync(FULL, key, 31);
    }
    if(lane == 0u) {
        misc[64u + warp] = last;
        misc[96u + warp] = heads;
    }
    __syncthreads();
    u32 carry = __reduce_max_sync(FULL, lane < warp ? misc[64u + lane] : 0u);
    const u32 distinct = __reduce_add_sync(FULL, misc[96u + lane]);
    const u32 le = lanemask_lt() | (1u << lane);
#pragma unroll 8
    for(u32 r = 0; r < 32u; ++r) {
        const u32 bits = fbits[warp * 32u + r];
        const u32 i = base + r * 32u + lane;
        const u32 mine = bits & le;
        const u32 head = mine ? base + r * 32u + 31u - (u32)__clz((int)mine) : carry;
        rk[sa[i]] = (u16)head;
        if(bits) {
            carry = base + r * 32u + 31u - (u32)__clz((int)bits);
        }
    }
    __syncthreads();
    return distinct;
}

// ----------------------------------------------------------------- the start --
// The first of the passes of a least-significant-digit-first sort need not be stable (nothing is ordered
// yet), so the rows go by their third AND fourth byte in one unstable counting sort with plain shared
// atomics on 65 536 counters (two u16 per word; a count never passes 32 768, so the halves never touch) --
// where two stable passes would cost 16 ballots per row.  A warp whose 32 rows show the same two bytes (runs)
// adds once for all of them.  s8: the block; c32: 32 768 words; out: u16[32 768]; xs: u32[64].
__device__ __forceinline__ void blk_start16(const u8* s8, u32* c32, u16* out, u32* xs)
{
    const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    {
        uint4* c4 = reinterpret_cast<uint4*>(c32);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for(u32 k = 0; k < 8u; ++k) {
            c4[k * BLK_THREADS + t] = z;
        }
    }
    __syncthreads();
    auto key_of = [&](u32 i) { return ((u32)s8[(i + 2u) & BLK_M] << 8) | (u32)s8[(i + 3u) & BLK_M]; };
#pragma unroll 4
    for(u32 k = 0; k < 32u; ++k) {
        const u32 key = key_of(k * BLK_THREADS + t);
        const u32 inc = (key & 1u) ? 0x10000u : 1u;
        if(__all_sync(FULL, key == __shfl_sync(FULL, key, 0))) {
            if(lane == 0u) {
                atomicAdd(&c32[key >> 1], inc * 32u);
            }
        } else {
            atomicAdd(&c32[key >> 1], inc);
        }
    }
    __syncthreads();
    // exclusive scan in key order: warp w owns words [1024 w, 1024 w + 1024), a row of 32 words per step
    u32* mine = c32 + warp * 1024u;
    {
        u32 sum = 0;
#pragma unroll 8
        for(u32 r = 0; r < 32u; ++r) {
            const u32 v = mine[r * 32u + lane];
            sum += (v & 0xFFFFu) + (v >> 16);
        }
        sum = __reduce_add_sync(FULL, sum);
        if(lane == 0u) {
            xs[warp] = sum;
        }
    }
    __syncthreads();
    u32 run = __reduce_add_sync(FULL, lane < warp ? xs[lane] : 0u);
#pragma unroll 4
    for(u32 r = 0; r < 32u; ++r) {
        const u32 v = mine[r * 32u + lane];
        const u32 lo = v & 0xFFFFu, both = lo + (v >> 16);
        u32 incl = both;
#pragma unroll
        for(u32 o = 1; o < 32u; o <<= 1) {
            const u32 up = __shfl_up_sync(FULL, incl, o);
            if(lane >= o) {
                incl += up;
            }
        }
        const u32 first = run + incl - both;
        mine[r * 32u + lane] = first | ((first + lo) << 16);
        run += __shfl_sync(FULL, incl, 31);
    }
    __syncthreads();
#pragma unroll 4
    for(u32 k = 0; k < 32u; ++k) {
        const u32 i = k * BLK_THREADS + t;
        const u32 key = key_of(i);
        const u32 inc = (key & 1u) ? 0x10000u : 1u, sh = (key & 1u) * 16u;
        u32 at;
        if(__all_sync(FULL, key == __shfl_sync(FULL, key, 0))) {
            u32 old = 0;
            if(lane == 0u) {
                old = atomicAdd(&c32[key >> 1], inc * 32u);
            }
            at = ((__shfl_sync(FULL, old, 0) >> sh) & 0xFFFFu) + lane;
        } else {
            at = (atomicAdd(&c32[key >> 1], inc) >> sh) & 0xFFFFu;
        }
        out[at] = (u16)i;
    }
    __syncthreads();
}

// ------------------------------------------------------------ short rounds --
// Once most rows are alone in their bucket a round only has to deal with the others ("active" rows: their
// bucket holds two rows or more).  Such a round lists the active rows by their second half (the same walk
// as before, filtered), sorts the list by bucket with the same two stable passes -- over n <= 16 384
// entries instead of 32 768 -- and puts every row at  bucket head + its index within the bucket's part of
// the list;  rows that are alone never move.  The list and the first pass's output share tmp (two halves).

// Rows in buckets of two or more, from the head flags: a row is alone when its place and the next are heads.
__device__ __forceinline__ u32 blk_count_active(const u32* fbits, u32* misc)
{
    const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    const u32 w = fbits[t];
    const u32 nx = t + 1u < BLK_THREADS ? fbits[t + 1u] : 1u;  // the place behind the last row counts as a head
    u32 alone = (u32)__popc(w & ((w >> 1) | (nx << 31)));
    alone = __reduce_add_sync(FULL, alone);
    if(lane == 0u) {
        misc[warp] = alone;
    }
    __syncthreads();
    const u32 total = __reduce_add_sync(FULL, misc[lane]);
    __syncthreads();
    return BLK_N - total;
}

// The active rows in the order of their second half -> list[0 .. n).  Returns n.
__device__ __forceinline__ u32 blk_list_active(const u16* sa, const u16* rk, const u32* fbits, u32 h, u16* list, u32* misc)
{
    const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    const u32 base = warp * 1024u + lane;
    u32 mine = 0;  // bit r: my row of step r is active
#pragma unroll 8
    for(u32 r = 0; r < 32u; ++r) {
        const u32 q = ((u32)sa[base + r * 32u] - h) & BLK_M;
        const u32 nx = (u32)rk[q] + 1u;  // the place behind the head of q's bucket: a head itself when q is alone
        const u32 word = nx < BLK_N ? fbits[nx >> 5] : 0xFFFFFFFFu;
        mine |= (((word >> (nx & 31u)) & 1u) ^ 1u) << r;
    }
    u32 total = __reduce_add_sync(FULL, (u32)__popc(mine));
    if(lane == 0u) {
        misc[warp] = total;
    }
    __syncthreads();
    const u32 other = misc[lane];
    u32 at = __reduce_add_sync(FULL, lane < warp ? other : 0u);
    const u32 n = __reduce_add_sync(FULL, other);
    const u32 lt = lanemask_lt();
#pragma unroll 8
    for(u32 r = 0; r < 32u; ++r) {
        const bool act = (mine >> r) & 1u;
        const u32 bits = __ballot_sync(FULL, act);
        if(act) {
            list[at + (u32)__popc(bits & lt)] = (u16)(((u32)sa[base + r * 32u] - h) & BLK_M);
        }
        at += (u32)__popc(bits);
    }
    __syncthreads();
    return n;
}

// blk_pass over a list of n <= 16 384 rows: warp w owns entries [32 w steps, 32 (w + 1) steps), steps = ceil(n / 1024).
template <u32 BITS, class Elem, class Digit>
__device__ __forceinline__ void blk_pass_short(u32 n, Elem elem, Digit digit, u16* dst, u16* cnt, u32* misc)
{
    const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    u32* c32 = reinterpret_cast<u32*>(cnt);
#pragma unroll
    for(u32 k = 0; k < 4u; ++k) {
        c32[t + k * BLK_THREADS] = 0u;
    }
    __syncthreads();
    u16* mine = cnt + warp * 256u;
    const u32 steps = (n + 1023u) >> 10;
    const u32 j0 = warp * steps * 32u + lane;
    const u32 lt = lanemask_lt();
    u32 dig[4], rnk[6];
#pragma unroll
    for(u32 k = 0; k < 4u; ++k) {
        dig[k] = 0u;
    }
#pragma unroll
    for(u32 k = 0; k < 6u; ++k) {
        rnk[k] = 0u;
    }
#pragma unroll
    for(u32 r = 0; r < 16u; ++r) {
        if(r < steps) {
            const u32 j = j0 + r * 32u;
            const bool valid = j < n;
            const u32 d = valid ? digit(elem(j)) : 0u;
            u32 peers = __ballot_sync(FULL, valid);
#pragma unroll
            for(u32 b = 0; b < BITS; ++b) {
                const u32 bit = (d >> b) & 1u;
                peers &= __ballot_sync(FULL, bit) ^ (bit - 1u);
            }
            const u32 leader = valid ? (u32)__ffs((int)peers) - 1u : 0u;
            u32 before = 0;
            if(valid && lane == leader) {
                before = mine[d];
                mine[d] = (u16)(before + __popc(peers));
            }
            before = __shfl_sync(FULL, before, leader) + __popc(peers & lt);
            if(valid) {
                dig[r >> 2] |= d << (8u * (r & 3u));
                rnk[r / 3u] |= before << (10u * (r % 3u));
            }
            __syncwarp();
        }
    }
    __syncthreads();
    blk_scan_counters(cnt, misc);
    __syncthreads();
#pragma unroll
    for(u32 r = 0; r < 16u; ++r) {
        const u32 j = j0 + r * 32u;
        if(r < steps && j < n) {
            const u32 d = (dig[r >> 2] >> (8u * (r & 3u))) & 0xFFu;
            const u32 at = (u32)mine[d] + ((rnk[r / 3u] >> (10u * (r % 3u))) & 0x3FFu);
            dst[at] = (u16)elem(j);
        }
    }
    __syncthreads();
}

// A round over at most 2048 active rows.  The pair (bucket, rank of the second half) says all there is to say
// about a row's new place, so ANY sort of the pairs will do -- no walk in second-half order, no stability:
// the rows are picked up by place (straight from the head flags), keyed, and put through a bitonic network
// in shared memory (66 compare-exchange steps for 2048 pairs) instead of two counting passes with their
// 16 KiB of counters to clear and scan.  keys: u32[2048], rows: u16[2048]; the sorted rows land in list[0 .. n).
__device__ __forceinline__ void blk_sort_tiny(u32 n, const u16* sa, const u16* rk, const u32* fbits, u32 h, u32* keys, u16* rows,
                                              u16* list, u32* misc)
{
    const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    const u32 w = fbits[t];
    const u32 nx = t + 1u < BLK_THREADS ? fbits[t + 1u] : 1u;
    u32 act = ~(w & ((w >> 1) | (nx << 31)));  // places whose bucket holds two rows or more
    const u32 mine = (u32)__popc(act);
    u32 incl = mine;
#pragma unroll
    for(u32 o = 1; o < 32u; o <<= 1) {
        const u32 up = __shfl_up_sync(FULL, incl, o);
        if(lane >= o) {
            incl += up;
        }
    }
    if(lane == 31u) {
        misc[warp] = incl;
    }
    __syncthreads();
    u32 at = __reduce_add_sync(FULL, lane < warp ? misc[lane] : 0u) + incl - mine;
    while(act) {
        const u32 bit = (u32)__ffs((int)act) - 1u;
        act &= act - 1u;
        const u32 q = sa[t * 32u + bit];
        keys[at] = ((u32)rk[q] << 16) | (u32)rk[(q + h) & BLK_M];
        rows[at] = (u16)q;
        ++at;
    }
    u32 pow2 = 2u;
    while(pow2 < n) {
        pow2 <<= 1;
    }
    for(u32 i = n + t; i < pow2; i += BLK_THREADS) {
        keys[i] = 0xFFFFFFFFu;  // larger than any pair: ranks stay below 2^15
        rows[i] = 0;
    }
    __syncthreads();
    for(u32 k = 2u; k <= pow2; k <<= 1) {
        for(u32 j = k >> 1; j > 0u; j >>= 1) {
            for(u32 x = t; x < (pow2 >> 1); x += BLK_THREADS) {
                const u32 i = 2u * x - (x & (j - 1u));
                const u32 p = i + j;
                const u32 a = keys[i], b = keys[p];
                if((a > b) == ((i & k) == 0u)) {
                    keys[i] = b;
                    keys[p] = a;
                    const u16 ra = rows[i];
                    rows[i] = rows[p];
                    rows[p] = ra;
                }
            }
            // a thread's pair lies within its warp's 64 entries when j <= 32: only the wide steps, and the last
            // step before a wide one, need the whole CTA (15 barriers instead of 66 for 2048 pairs)
            if(j > 32u || (j == 1u && k >= 64u)) {
                __syncthreads();
            } else {
                __syncwarp();
            }
        }
    }
    __syncthreads();
    for(u32 i = t; i < n; i += BLK_THREADS) {
        list[i] = rows[i];
    }
    __syncthreads();
}

// The sorted list (by bucket, then by second half) goes home: row j of the list lands at
//   head of its bucket + (j - first j of that bucket),
// its new rank is the place of the first row of its run of equal (bucket, second half) pairs, and every
// such run that does not start a bucket sets a new head flag.  lb / lg: 512 words each (ballots of the
// list: bucket starts, run starts).  Returns the number of new heads.
__device__ __forceinline__ u32 blk_place_short(u32 n, const u16* list, u16* sa, u16* rk, u32* fbits, u32 h, u32* lb, u32* lg, u32* misc)
{
    const u32 t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    const u32 steps = (n + 1023u) >> 10;
    const u32 jw = warp * steps * 32u;
    auto key_of = [&](u32 q) { return ((u32)rk[q] << 16) | (u32)rk[(q + h) & BLK_M]; };
    u32 last_b = 0, last_g = 0, fresh = 0;
    u32 carry_key = (jw > 0u && jw - 1u < n) ? key_of((u32)list[jw - 1u]) : 0u;
#pragma unroll 4
    for(u32 r = 0; r < steps; ++r) {
        const u32 j = jw + r * 32u + lane;
        const bool valid = j < n;
        const u32 key = valid ? key_of((u32)list[j]) : 0u;
        u32 prev = __shfl_up_sync(FULL, key, 1);
        if(lane == 0u) {
            prev = carry_key;
        }
        const bool b = valid && (j == 0u || (key >> 16) != (prev >> 16));
        const bool g = valid && (b || key != prev);
        const u32 bb = __ballot_sync(FULL, b), gb = __ballot_sync(FULL, g);
        if(lane == 0u) {
            lb[warp * steps + r] = bb;
            lg[warp * steps + r] = gb;
        }
        if(bb) {
            last_b = jw + r * 32u + 31u - (u32)__clz((int)bb);
        }
        if(gb) {
            last_g = jw + r * 32u + 31u - (u32)__clz((int)gb);
        }
        fresh += (u32)__popc(gb & ~bb);
        carry_key = __shfl_sync(FULL, key, 31);
    }
    if(lane == 0u) {
        misc[64u + warp] = last_b;
        misc[96u + warp] = last_g;
        misc[warp] = fresh;
    }
    __syncthreads();
    u32 cb = __reduce_max_sync(FULL, lane < warp ? misc[64u + lane] : 0u);
    u32 cg = __reduce_max_sync(FULL, lane < warp ? misc[96u + lane] : 0u);
    const u32 total_fresh = __reduce_add_sync(FULL, misc[lane]);
    const u32 le = lanemask_lt() | (1u << lane);
    u32 out[16];  // place | new rank << 16
#pragma unroll
    for(u32 r = 0; r < 16u; ++r) {
        out[r] = 0u;
        if(r < steps) {
            const u32 bb = lb[warp * steps + r], gb = lg[warp * steps + r];
            const u32 j = jw + r * 32u + lane;
            const u32 mb = bb & le, mg = gb & le;
            const u32 jf = mb ? jw + r * 32u + 31u - (u32)__clz((int)mb) : cb;
            const u32 jg = mg ? jw + r * 32u + 31u - (u32)__clz((int)mg) : cg;
            if(j < n) {
                const u32 head = rk[list[j]];
                out[r] = (head + (j - jf)) | ((head + (jg - jf)) << 16);
            }
            if(bb) {
                cb = jw + r * 32u + 31u - (u32)__clz((int)bb);
            }
            if(gb) {
                cg = jw + r * 32u + 31u - (u32)__clz((int)gb);
            }
        }
    }
    __syncthreads();  // every rank has been read
#pragma unroll
    for(u32 r = 0; r < 16u; ++r) {
        const u32 j = jw + r * 32u + lane;
        if(r < steps && j < n) {
            const u32 q = list[j];
            const u32 place = out[r] & 0xFFFFu;
            sa[place] = (u16)q;
            rk[q] = (u16)(out[r] >> 16);
            const u32 bb = lb[warp * steps + r], gb = lg[warp * steps + r];
            if(((gb & ~bb) >> lane) & 1u) {
                atomicOr(&fbits[place >> 5], 1u << (place & 31u));
            }
        }
    }
    __syncthreads();
    return total_fresh;
}

// ------------------------------------------------------------------- B1 forward --
// src: nblocks x 32 768 bytes (16-byte aligned); dst: nblocks x 32 770 bytes (2-byte aligned).
// rounds (one u32 per block): how many doubling rounds the block took; bit 31 = the block has a period
// (rotations tie): its row number is the canonical one (ties by position) until B3 has replayed it;
// bit 30 = all rotations are equal (one repeated byte): row 0, which is what the reference reports.
// list / rk_out (both optional, used together by the tie replay B3): CTA c sorts block list[c] and leaves the
// final rank of every rotation in rk_out[c][32 768] instead of touching dst.
// `ties` (optional): ties[0] counts the blocks of this call that have a period (and are not one repeated
// byte), ties[1 ..] lists them as tie_base + block number, in no particular order -- the host reads the
// counter alone and, only when it is not zero, the list (blocks are numbered in 32 bits: 128 TiB per call).
__global__ void __launch_bounds__(BLK_THREADS, 1) k_blk_fwd(const u8* __restrict__ src, u8* __restrict__ dst, u32* __restrict__ rounds,
                                                             const u32* __restrict__ list, u16* __restrict__ rk_out,
                                                             u32* __restrict__ ties = nullptr, u32 tie_base = 0)
{
    extern __shared__ __align__(16) u8 blk_sm[];
    u16* sa = reinterpret_cast<u16*>(blk_sm + BF_SA);
    u16* tmp = reinterpret_cast<u16*>(blk_sm + BF_TMP);
    u16* rk = reinterpret_cast<u16*>(blk_sm + BF_RK);
    u16* cnt = reinterpret_cast<u16*>(blk_sm + BF_CNT);
    u32* fbits = reinterpret_cast<u32*>(blk_sm + BF_FB);
    u32* misc = reinterpret_cast<u32*>(blk_sm + BF_MISC);
    const u32 t = threadIdx.x;
    const u64 blk = list ? list[blockIdx.x] : blockIdx.x;
    const u8* s = src + blk * BLK_N;
    // the block's bytes wait where the ranks will live: they are last read by the flags of the first ranking
    u8* s8 = blk_sm + BF_RK;
    {
        const uint4* g = reinterpret_cast<const uint4*>(s);
        uint4* d4 = reinterpret_cast<uint4*>(s8);
        d4[t] = __ldg(g + t);
        d4[t + BLK_THREADS] = __ldg(g + t + BLK_THREADS);
    }
    __syncthreads();
    // rows by their first four bytes, least significant first: bytes 3 and 2 in one unstable counting sort,
    // bytes 1 and 0 in two stable passes, then one ranking (two bytes + a doubling round would be two)
    {
        u16* start_out = reinterpret_cast<u16*>(blk_sm + BF_START_OUT);
        blk_start16(s8, reinterpret_cast<u32*>(blk_sm + BF_SA), start_out, reinterpret_cast<u32*>(blk_sm + BF_START_X));
        const uint4* o4 = reinterpret_cast<const uint4*>(start_out);
        uint4* s4 = reinterpret_cast<uint4*>(sa);
#pragma unroll
        for(u32 k = 0; k < 4u; ++k) {
            s4[k * BLK_THREADS + t] = o4[k * BLK_THREADS + t];
        }
        __syncthreads();
    }
    blk_pass<8>([&](u32 i) { return (u32)sa[i]; }, [&](u32 e) { return (u32)s8[(e + 1u) & BLK_M]; }, tmp, cnt, misc);
    blk_pass<8>([&](u32 i) { return (u32)tmp[i]; }, [&](u32 e) { return (u32)s8[e]; }, sa, cnt, misc);
    u32 distinct = blk_rerank(sa, rk, fbits, misc, [&](u32 e) {
        return ((u32)s8[e] << 24) | ((u32)s8[(e + 1u) & BLK_M] << 16) | ((u32)s8[(e + 2u) & BLK_M] << 8) | (u32)s8[(e + 3u) & BLK_M];
    });
    u32 h = 4, nrounds = 0;
    u32* lb = reinterpret_cast<u32*>(blk_sm + BF_LB);
    u16* half = tmp + BLK_N / 2u;
    while(distinct < BLK_N && h < BLK_N) {
        const u32 active = blk_count_active(fbits, misc);
        if(active <= BLK_TINY) {
            blk_sort_tiny(active, sa, rk, fbits, h, reinterpret_cast<u32*>(half), half + 2u * BLK_TINY, tmp, misc);
            distinct += blk_place_short(active, tmp, sa, rk, fbits, h, lb, lb + 512u, misc);
            nrounds += 0x100u;
        } else if(active > BLK_N / 2u) {
            blk_pass<8>([&](u32 i) { return ((u32)sa[i] - h) & BLK_M; }, [&](u32 e) { return (u32)rk[e] & 0xFFu; }, tmp, cnt, misc);
            blk_pass<7>([&](u32 i) { return (u32)tmp[i]; }, [&](u32 e) { return (u32)rk[e] >> 8; }, sa, cnt, misc);
            distinct = blk_rerank(sa, rk, fbits, misc, [&](u32 e) { return ((u32)rk[e] << 16) | (u32)rk[(e + h) & BLK_M]; });
        } else {
            const u32 n = blk_list_active(sa, rk, fbits, h, tmp, misc);
            blk_pass_short<8>(n, [&](u32 j) { return (u32)tmp[j]; }, [&](u32 e) { return (u32)rk[e] & 0xFFu; }, half, cnt, misc);
            blk_pass_short<7>(n, [&](u32 j) { return (u32)half[j]; }, [&](u32 e) { return (u32)rk[e] >> 8; }, tmp, cnt, misc);
            distinct += blk_place_short(n, tmp, sa, rk, fbits, h, lb, lb + 512u, misc);
            nrounds += 0x100u;  // bits 8..15 of the rounds word: how many of the rounds were short ones
        }
        h <<= 1;
        ++nrounds;
    }
    if(rk_out) {
        const uint4* r4 = reinterpret_cast<const uint4*>(rk);
        uint4* o4 = reinterpret_cast<uint4*>(rk_out + (u64)blockIdx.x * BLK_N);
#pragma unroll
        for(u32 k = 0; k < 4u; ++k) {
            o4[k * BLK_THREADS + t] = r4[k * BLK_THREADS + t];
        }
        return;
    }
    // the column: byte in front of each row; the bytes come back into the free half of tmp
    u8* sb = blk_sm + BF_TMP;
    {
        const uint4* g = reinterpret_cast<const uint4*>(s);
        uint4* d4 = reinterpret_cast<uint4*>(sb);
        d4[t] = __ldg(g + t);
        d4[t + BLK_THREADS] = __ldg(g + t + BLK_THREADS);
    }
    __syncthreads();
    u8* o = dst + blk * BLK_CODED;
    const u32* sa2 = reinterpret_cast<const u32*>(sa);
#pragma unroll 4
    for(u32 k = 0; k < 16u; ++k) {
        const u32 pair = sa2[k * BLK_THREADS + t];
        const u32 a = sb[((pair & 0xFFFFu) + BLK_M) & BLK_M];
        const u32 b = sb[((pair >> 16) + BLK_M) & BLK_M];
        *reinterpret_cast<u16*>(o + 2u * (k * BLK_THREADS + t)) = (u16)(a | (b << 8));
    }
    if(t == 0u) {
        // rank of rotation 0 = its row when all rows differ; with ties, the first row of its run
        *reinterpret_cast<u16*>(o + BLK_N) = rk[0];
        if(rounds) {
            rounds[blk] = nrounds | (distinct < BLK_N ? 0x80000000u : 0u) | (distinct == 1u ? 0x40000000u : 0u);
        }
        if(ties && distinct < BLK_N && distinct != 1u) {
            ties[1u + atomicAdd(ties, 1u)] = tie_base + (u32)blk;
        }
    }
}

// ------------------------------------------------------------------- B2 inverse --
// src: nblocks x 32 770 bytes (2-byte aligned); dst: nblocks x 32 768 bytes (4-byte aligned).
__global__ void __launch_bounds__(BLK_THREADS, 1) k_blk_inv(const u8* __restrict__ src, u8* __restrict__ dst, int* err)
{
    extern __shared__ __align__(16) u8 blk_sm[];
    u16* ja = reinterpret_cast<u16*>(blk_sm + BI_JA);
    u16* jb = reinterpret_cast<u16*>(blk_sm + BI_JB);
    u16* walk = reinterpret_cast<u16*>(blk_sm + BI_P);
    u8* col = blk_sm + BI_L;
    u32* misc = reinterpret_cast<u32*>(blk_sm + BI_MISC);
    const u32 t = threadIdx.x;
    const u64 blk = blockIdx.x;
    const u8* c = src + blk * BLK_CODED;
    {
        const u16* g = reinterpret_cast<const u16*>(c);
        u16* d2 = reinterpret_cast<u16*>(col);
#pragma unroll 4
        for(u32 k = 0; k < 16u; ++k) {
            d2[k * BLK_THREADS + t] = __ldg(g + k * BLK_THREADS + t);
        }
    }
    u32 top = __ldg(reinterpret_cast<const u16*>(c + BLK_N));
    if(top >= BLK_N) {  // the reference reads out of bounds here (blksort.h:652-654)
        if(t == 0u) {
            atomicOr(err, ERR_CORRUPT);
        }
        top = 0;
    }
    __syncthreads();
    // counting_sort (blksort.h:379-402): next[k] = place in the column of the k-th smallest byte, stable
    blk_pass<8>([&](u32 i) { return i; }, [&](u32 e) { return (u32)col[e]; }, ja, jb, misc);
    u8* o = dst + blk * BLK_N;
    // ---- the walk, cut at stations.  Every 8th row and the starting row p0 are stations; from each, a
    // thread follows `next` to the following station (8 steps on average, four stations per thread in
    // flight), which leaves a list of <= 4097 legs with their lengths.  Ranking that short list (13 rounds
    // over 4097 entries instead of 14 over 32 768) tells every leg where in the output it starts, and the
    // same threads walk their legs once more, this time writing bytes.  When the cycle through p0 is not
    // the whole permutation (a block with a period) or a leg gets long (a `next` that dodges the
    // stations), the doubling below does the job instead.
    {
        constexpr u32 NODES = 4608u, NIL = 0xFFFFu, LEG_MAX = 1024u;
        u16* nx = jb;
        u16* rr = jb + NODES;
        u16* nx2 = jb + 2u * NODES;
        u16* rr2 = jb + 3u * NODES;
        u8* outb = reinterpret_cast<u8*>(walk);
        const u32 p0 = ja[top];
        const u32 start = (p0 & 7u) ? 4096u : (p0 >> 3);
        const u32 nnodes = (p0 & 7u) ? 4097u : 4096u;
        if(t == 0u) {
            misc[0] = 0u;
        }
        __syncthreads();
        auto station = [&](u32 q) { return (q & 7u) == 0u || q == p0; };
        auto leg_end = [&](u32 q) { return q == p0 ? NIL : (q >> 3); };  // the leg that reaches p0 ends the list
        u32 q[4], len[4];
        bool act[4];
#pragma unroll
        for(u32 k = 0; k < 4u; ++k) {
            q[k] = ja[8u * (t + BLK_THREADS * k)];
            len[k] = 1u;
            act[k] = !station(q[k]);
        }
        while(act[0] || act[1] || act[2] || act[3]) {
#pragma unroll
            for(u32 k = 0; k < 4u; ++k) {
                if(act[k]) {
                    q[k] = ja[q[k]];
                    ++len[k];
                    act[k] = !station(q[k]) && len[k] < LEG_MAX;
                }
            }
        }
        bool lost = false;
#pragma unroll
        for(u32 k = 0; k < 4u; ++k) {
            lost = lost || !station(q[k]);
            nx[t + BLK_THREADS * k] = (u16)leg_end(q[k]);
            rr[t + BLK_THREADS * k] = (u16)len[k];
        }
        u32 xlen = 1u;
        if(t == BLK_THREADS - 1u && nnodes == 4097u) {  // the leg that starts at p0 when p0 is no 8th row
            u32 y = ja[p0];
            while(!station(y) && xlen < LEG_MAX) {
                y = ja[y];
                ++xlen;
            }
            lost = lost || !station(y);
            nx[4096] = (u16)leg_end(y);
            rr[4096] = (u16)xlen;
        }
        if(lost) {
            atomicOr(&misc[0], 1u);
        }
        __syncthreads();
#pragma unroll 1
        for(u32 round = 0; round < 13u; ++round) {
#pragma unroll
            for(u32 k = 0; k < 5u; ++k) {
                const u32 v = t + BLK_THREADS * k;
                if(v < nnodes) {
                    const u32 n = nx[v];
                    const u32 r = rr[v];
                    if(n != NIL) {
                        rr2[v] = (u16)(r + rr[n]);
                        nx2[v] = nx[n];
                    } else {
                        rr2[v] = (u16)r;
                        nx2[v] = (u16)NIL;
                    }
                }
            }
            __syncthreads();
            u16* sw = nx;
            nx = nx2;
            nx2 = sw;
            sw = rr;
            rr = rr2;
            rr2 = sw;
        }
        // rr[v]: bytes from the start of leg v to the end of the walk; the leg of p0 sees all of it
        const bool whole = misc[0] == 0u && (u32)rr[start] == (BLK_N & 0xFFFFu) && nx[start] == NIL;
        if(whole) {
#pragma unroll
            for(u32 k = 0; k < 4u; ++k) {
                const u32 v = t + BLK_THREADS * k;
                u32 at = (BLK_N - (u32)rr[v]) & BLK_M;
                u32 y = 8u * v;
                for(u32 j = 0; j < len[k]; ++j) {
                    outb[at + j] = col[y];
                    y = ja[y];
                }
            }
            if(t == BLK_THREADS - 1u && nnodes == 4097u) {
                u32 y = p0;
                for(u32 j = 0; j < xlen; ++j) {  // this leg opens the walk
                    outb[j] = col[y];
                    y = ja[y];
                }
            }
            __syncthreads();
            const uint4* o4 = reinterpret_cast<const uint4*>(outb);
            reinterpret_cast<uint4*>(o)[t] = o4[t];
            reinterpret_cast<uint4*>(o)[t + BLK_THREADS] = o4[t + BLK_THREADS];
            return;
        }
        __syncthreads();
    }
    // ---- pointer doubling: jump tables J <- J o J, the known part of the walk doubles every round
    if(t == 0u) {
        walk[0] = ja[top];
    }
    __syncthreads();
    u16* cur = ja;
    u16* nxt = jb;
#pragma unroll 1
    for(u32 k = 0; k < 15u; ++k) {
        const u32 len = 1u << k;
        for(u32 i = t; i < len; i += BLK_THREADS) {
            walk[len + i] = cur[walk[i]];
        }
        if(k < 14u) {
#pragma unroll 8
            for(u32 j = 0; j < 32u; ++j) {
                const u32 p = j * BLK_THREADS + t;
                nxt[p] = cur[cur[p]];
            }
        }
        __syncthreads();
        u16* sw = cur;
        cur = nxt;
        nxt = sw;
    }
    const uint2* w4 = reinterpret_cast<const uint2*>(walk);
#pragma unroll 4
    for(u32 k = 0; k < 8u; ++k) {
        const uint2 q = w4[k * BLK_THREADS + t];
        const u32 v = (u32)col[q.x & 0xFFFFu] | ((u32)col[q.x >> 16] << 8) | ((u32)col[q.y & 0xFFFFu] << 16) |
                      ((u32)col[q.y >> 16] << 24);
        *reinterpret_cast<u32*>(o + 4u * (k * BLK_THREADS + t)) = v;
    }
}

// ---------------------------------------------------------------- B3 tie replay --
// A block with a period has runs of EQUAL rotations.  Their last bytes are equal, so the column is settled,
// but the row number the reference reports for rotation 0 depends on where its unstable multikey quicksort
// (mqsort, blksort.h:281-362, level 11, median of three on byte 0, insertion sort below 37 rows, heapsort
// when the levels run out) happens to leave it inside its run.  This kernel replays that sort for such
// blocks, swap for swap, with two shortcuts that cannot change a swap:
//   * comparisons over the whole rotation (insertion sort, heapsort) are rank comparisons -- B1 has
//     already ranked every rotation (rk_in);
//   * a pass in which every row shows the same byte moves nothing, so a range goes straight to the first
//     depth at which its smallest and largest rotation differ (none: all rows equal, nothing ever moves).
// The "<", "=", ">" parts of a partition are disjoint ranges, so the order in which they are finished
// does not matter: warps take ranges from a queue, wave after wave; lane 0 does the serial partition, the
// warp does the range scans.
struct TieRange {
    u16 lo, size_m1, depth, level;
};
constexpr u32 BT_S = 0u, BT_V = 32768u, BT_RK = 98304u, BT_MISC = 163840u;
constexpr u32 BLK_TIES_SMEM = BT_MISC + 64u;
constexpr u32 BLK_TIES_QUEUE = 16384u;  // ranges per wave, at most (a range holds two rows or more)

__device__ __forceinline__ void tie_sift(u16* h1, const u16* rk, int root, int last, u32 x)
{
    int i = root, j;
    while((j = i << 1) <= last) {
        if(j < last && rk[h1[j]] < rk[h1[j + 1]]) {
            ++j;
        }
        if(!(rk[x] < rk[h1[j]])) {
            break;
        }
        h1[i] = h1[j];
        i = j;
    }
    h1[i] = (u16)x;
}

__global__ void __launch_bounds__(BLK_THREADS, 1) k_blk_ties(const u8* __restrict__ src, u8* __restrict__ dst, const u32* __restrict__ list,
                                                              const u16* __restrict__ rk_in, TieRange* __restrict__ queues)
{
    extern __shared__ __align__(16) u8 blk_sm[];
    u8* s8 = blk_sm + BT_S;
    u16* v = reinterpret_cast<u16*>(blk_sm + BT_V);
    u16* rk = reinterpret_cast<u16*>(blk_sm + BT_RK);
    u32* misc = reinterpret_cast<u32*>(blk_sm + BT_MISC);  // [0] next range to take, [1] ranges queued for the next wave
    const u32 t = threadIdx.x, lane = t & 31u;
    const u64 blk = list[blockIdx.x];
    {
        const uint4* g = reinterpret_cast<const uint4*>(src + blk * BLK_N);
        uint4* d4 = reinterpret_cast<uint4*>(s8);
        d4[t] = __ldg(g + t);
        d4[t + BLK_THREADS] = __ldg(g + t + BLK_THREADS);
        const uint4* r4 = reinterpret_cast<const uint4*>(rk_in + (u64)blockIdx.x * BLK_N);
        uint4* k4 = reinterpret_cast<uint4*>(rk);
#pragma unroll
        for(u32 k = 0; k < 4u; ++k) {
            k4[k * BLK_THREADS + t] = r4[k * BLK_THREADS + t];
        }
        for(u32 k = 0; k < 32u; ++k) {
            v[k * BLK_THREADS + t] = (u16)(k * BLK_THREADS + t);
        }
    }
    TieRange* q_cur = queues + (u64)blockIdx.x * 2u * BLK_TIES_QUEUE;
    TieRange* q_nxt = q_cur + BLK_TIES_QUEUE;
    if(t == 0u) {
        q_cur[0] = TieRange{0, (u16)(BLK_N - 1u), 0, 11};
        misc[0] = 0;
        misc[1] = 0;
    }
    u32 ncur = 1;
    __syncthreads();
    while(ncur) {
        for(;;) {
            u32 take = 0;
            if(lane == 0u) {
                take = atomicAdd(&misc[0], 1u);
            }
            take = __shfl_sync(FULL, take, 0);
            if(take >= ncur) {
                break;
            }
            const TieRange r = q_cur[take];
            u16* w = v + r.lo;
            const u32 size = (u32)r.size_m1 + 1u;
            u32 d = r.depth;
            if(r.level == 0u) {  // heapsort, blksort.h:235-279
                if(lane == 0u) {
                    u16* h1 = w - 1;
                    int last = (int)size;
                    for(int k = last >> 1; k >= 1; --k) {
                        tie_sift(h1, rk, k, last, h1[k]);
                    }
                    while(last > 1) {
                        const u32 x = h1[last];
                        h1[last] = h1[1];
                        --last;
                        tie_sift(h1, rk, 1, last, x);
                    }
                }
                __syncwarp();
                continue;
            }
            if(size < 37u) {  // insertionsort, blksort.h:223-233
                if(lane == 0u) {
                    for(u32 i = 1; i < size; ++i) {
                        const u32 x = w[i];
                        int j = (int)i - 1;
                        while(j >= 0 && rk[x] < rk[w[j]]) {
                            w[j + 1] = w[j];
                            --j;
                        }
                        w[j + 1] = (u16)x;
                    }
                }
                __syncwarp();
                continue;
            }
            // smallest and largest rotation of the range (rank in the high half, the row in the low half)
            u32 mn = 0xFFFFFFFFu, mx = 0u;
            for(u32 i = lane; i < size; i += 32u) {
                const u32 row = w[i];
                const u32 key = ((u32)rk[row] << 16) | row;
                mn = key < mn ? key : mn;
                mx = key > mx ? key : mx;
            }
            mn = __reduce_min_sync(FULL, mn);
            mx = __reduce_max_sync(FULL, mx);
            if((mn >> 16) == (mx >> 16)) {
                continue;  // all rows equal: no pass, no insertion sort ever moves one of them
            }
            {
                const u32 a = mn & 0xFFFFu, b = mx & 0xFFFFu;
                u32 off = d;
                for(;;) {
                    const u32 o = off + lane;
                    const bool ne = o < BLK_N && s8[(a + o) & BLK_M] != s8[(b + o) & BLK_M];
                    const u32 bits = __ballot_sync(FULL, ne);
                    if(bits) {
                        d = off + (u32)__ffs((int)bits) - 1u;
                        break;
                    }
                    off += 32u;
                    if(off >= BLK_N) {  // cannot happen for ranks B1 produced (they differ, so do the rotations)
                        d = BLK_N;
                        break;
                    }
                }
            }
            if(d >= BLK_N) {
                continue;
            }
            if(lane == 0u) {
                // one pass of mqsort at depth d, blksort.h:293-358
                const u32 q1 = size >> 2, q2 = q1 + q1, q3 = q1 + q2;
                const u32 ba = s8[w[q1]], bb = s8[w[q2]], bc = s8[w[q3]];
                const u32 pick = ba < bb ? (bb < bc ? w[q2] : (ba < bc ? w[q3] : w[q1])) : (ba < bc ? w[q1] : (bb < bc ? w[q3] : w[q2]));
                const u32 p = s8[(pick + d) & BLK_M];
                const int hi = (int)size - 1;
                int lo = 0, up = hi, eql = 0, eqr = hi;
                for(;;) {
                    while(lo <= up) {
                        const u32 row = w[lo];
                        const u32 c = s8[(row + d) & BLK_M];
                        if(p < c) {
                            break;
                        }
                        if(p == c) {
                            w[lo] = w[eql];
                            w[eql] = (u16)row;
                            ++eql;
                        }
                        ++lo;
                    }
                    while(lo <= up) {
                        const u32 row = w[up];
                        const u32 c = s8[(row + d) & BLK_M];
                        if(c < p) {
                            break;
                        }
                        if(p == c) {
                            w[up] = w[eqr];
                            w[eqr] = (u16)row;
                            --eqr;
                        }
                        --up;
                    }
                    if(up < lo) {
                        break;
                    }
                    const u16 x = w[lo];
                    w[lo] = w[up];
                    w[up] = x;
                    ++lo;
                    --up;
                }
                const int nl = eql < lo - eql ? eql : lo - eql;
                for(int i = 0; i < nl; ++i) {
                    const u16 x = w[i];
                    w[i] = w[up - i];
                    w[up - i] = x;
                }
                const int less_n = lo - eql;
                const int ra = hi - eqr, rb = eqr - up;
                const int nr = ra < rb ? ra : rb;
                for(int i = 0; i < nr; ++i) {
                    const u16 x = w[lo + i];
                    w[lo + i] = w[hi - i];
                    w[hi - i] = x;
                }
                const int gt_at = hi - (eqr - up) + 1;
                if(less_n >= 2) {
                    q_nxt[atomicAdd(&misc[1], 1u)] = TieRange{r.lo, (u16)(less_n - 1), (u16)d, (u16)(r.level - 1u)};
                }
                if((int)size - gt_at >= 2) {
                    q_nxt[atomicAdd(&misc[1], 1u)] = TieRange{(u16)(r.lo + gt_at), (u16)((int)size - gt_at - 1), (u16)d, (u16)(r.level - 1u)};
                }
                if(gt_at - less_n >= 2 && d + 1u < BLK_N) {
                    q_nxt[atomicAdd(&misc[1], 1u)] = TieRange{(u16)(r.lo + less_n), (u16)(gt_at - less_n - 1), (u16)(d + 1u), r.level};
                }
            }
            __syncwarp();
        }
        __syncthreads();
        ncur = misc[1];
        __syncthreads();
        if(t == 0u) {
            misc[0] = 0;
            misc[1] = 0;
        }
        TieRange* sw = q_cur;
        q_cur = q_nxt;
        q_nxt = sw;
        __syncthreads();
    }
    for(u32 k = 0; k < 32u; ++k) {
        const u32 i = k * BLK_THREADS + t;
        if(v[i] == 0u) {
            *reinterpret_cast<u16*>(dst + blk * BLK_CODED + BLK_N) = (u16)i;
        }
    }
}

}  // namespace b2rc
