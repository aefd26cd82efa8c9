// b2rc_adaptseg.cuh -- restart points for the ADAPTIVE range coder: several chains per block for its
// decoder (AdaptiveRangeDecoder::decode, cpprcoder.h:872-924), the slowest kernel of the path.
//
// The adaptive model is part of the coder's state, so a restart point carries it: besides {bytes
// shifted out so far, the encoder's low, range} the 256 symbol counts the
// model holds there (u16: blocks of at most 65536 bytes) -- 524 bytes per point, every 16384 symbols
// by default.  The payloads stay the reference's bytes; the points live behind them in the container.
//
//   k_enc_adaptive (b2rc_kernels.cuh) records the points while it codes (adaptive_mark below);
//   k_dec_adaptive_seg decodes every segment as a chain of its own: one warp = 32 blocks x one segment.
//
// The decoder's model here is the count tree WITHOUT its leaves: a node holds how many symbols so far
// went into its left subtree; the number that went into the subtree the walk is in comes down with
// the walk (total so far at the root, then v or sub - v at every level), so the symbol's own count --
// its frequency -- is what is left at the bottom.  255 nodes instead of 511 entries: 16 KiB of shared
// memory per warp instead of 32, eleven warps per SM instead of six; for a kernel that is bound by
// how many chains an SM holds, that is most of the gain.  A segment builds its tree from the point's
// counts; every segment has to END where the next point stands (records are as untrusted as the
// rest of the container).
#pragma once
#include "b2rc_kernels.cuh"

namespace b2rc
{
// =================================================== K2a, two warps per 32 blocks ==
// k_enc_adaptive runs one warp per scheduler and is bound by the length of what that warp does per
// symbol: 115 instructions at 2.2 cycles each, of which 65 are the model (nine table reads, nine
// writes, the bookkeeping of eight tree levels) and 50 the coder (the exact division by 256 + i, the
// step, the commit).  The model does not depend on the coder at all, so the two halves go to two
// warps of one CTA: warp 0 walks the tree for tile k + 1 and leaves (cum - symbol, freq - 1), 16 bits
// each, in shared memory while warp 1 codes tile k from what warp 0 left a barrier ago.  Same tables,
// same arithmetic, same bytes; the time per symbol is the longer half instead of the sum.
// Blocks of at most 65536 bytes (the pairs are 16 + 16 bits); longer blocks keep k_enc_adaptive<u32>.
// MEASURED (1 GiB mixed stream): 9.2 ms against 7.9 ms for k_enc_adaptive -- SLOWER.  The model half is
// not 65 instructions' worth of time but a read-modify-write chain through shared memory (every symbol
// reads the nodes the previous symbol wrote, the root always): about 250 cycles per symbol on its own,
// as long as the whole one-warp loop, plus a barrier per tile.  Kept behind B2RC_ADAPTIVE_TWO_WARPS=1
// (same bytes, tests/test_gpu_adaptseg.py::test_two_warp_encoder_writes_the_same_container).
constexpr u32 ENC_AD2_TAB = 512u * 32u * 2u;
constexpr u32 ENC_AD2_SRC = 3u * TILE_BYTES;        // three input tiles: being staged / modelled / coded
constexpr u32 ENC_AD2_PAIRS = 2u * TILE * 128u;     // two tiles of [symbol][lane] u32
constexpr u32 ENC_AD2_SMEM = ENC_AD2_TAB + ENC_AD2_SRC + ENC_AD2_PAIRS;

__global__ void __launch_bounds__(64) k_enc_adaptive2(EncArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    const u32 sbase = smem_addr(smem);
    const u32 src_a = sbase + ENC_AD2_TAB;
    const u32 pairs_a = src_a + ENC_AD2_SRC;
    const u32 warp = threadIdx.x >> 5, lane = lane_id();
    const u64 b0 = (u64)blockIdx.x * 32u;
    const u64 b = b0 + lane;
    const bool has = b < a.nblocks;
    u32 n_b = 0;
    if(has) {
        const u64 lo = b * (u64)a.block;
        n_b = (u32)((a.n - lo < a.block) ? (a.n - lo) : a.block);
    }
    u8* slot = a.slots + b * a.slot_stride;
    const u32 n_max = __reduce_max_sync(FULL, n_b);
    const u32 ntiles = (n_max + TILE - 1) / TILE;
    const u32 nrec = a.restart ? (a.block + a.seg_syms - 1u) / a.seg_syms - 1u : 0u;
    const u32 seg_tiles = a.restart ? a.seg_syms / TILE : 0xFFFFFFFFu;

    {  // initialize(): all counts zero (the ones are implicit), cpprcoder.h:1094-1132
        uint4* z = reinterpret_cast<uint4*>(smem);
        for(u32 i = threadIdx.x; i < ENC_AD2_TAB / 16u; i += 64u) {
            z[i] = make_uint4(0, 0, 0, 0);
        }
    }
    LaneTab<u16> tab{sbase + lane * 2u};
    if(warp == 0) {
        stage_tile(src_a, a.src, a.n, b0, a.block, 0, lane);
        cp_async_commit();
        if(ntiles > 1) {
            stage_tile(src_a + TILE_BYTES, a.src, a.n, b0, a.block, TILE, lane);
        }
        cp_async_commit();
    } else if(has) {
        *reinterpret_cast<u32*>(slot) = n_b;  // cpprcoder.h:689-694
    }
    __syncthreads();

    // ---- warp 0: the model of one tile -> pairs
    auto model_tile = [&](u32 k) {
        if(nrec && k != 0u && k % seg_tiles == 0u && k * TILE < n_b) {
            // a segment starts with this tile: the model as it stands, for the restart point (the coder adds its three words)
            u32* rec = a.restart + ((b0 + lane) * nrec + k / seg_tiles - 1u) * (u64)ADAPT_REC_WORDS;
#pragma unroll 4
            for(u32 s = 0; s < 256u; s += 2u) {
                rec[3u + s / 2u] = tab.ld(256u + s) | (tab.ld(257u + s) << 16);
            }
        }
        const u32 row = src_a + (k % 3u) * TILE_BYTES + lane * ROW;
        const u32 out = pairs_a + (k & 1u) * (TILE * 128u) + lane * 4u;
#pragma unroll 1
        for(int wi = 0; wi < TILE / 4; ++wi) {
            const u32 word = lds32(row + 4u * (u32)wi);
#pragma unroll
            for(int q = 0; q < 4; ++q) {
                const u32 j = (u32)(wi * 4 + q);
                if(k * TILE + j < n_b) {
                    const u32 sym = (word >> (8 * q)) & 0xFFu;
                    u32 cum, freq;
                    TreeOps<u16>::encode(tab.base, sym, cum, freq);
                    sts32v(out + j * 128u, (cum - sym) | ((freq - 1u) << 16));  // both fit 16 bits: at most 65535 symbols so far
                }
            }
        }
    };

    RcEnc st;
    rc_enc_init(st, RC_ADAPT_RANGE0);
    SlotSink sink;
    sink.out = reinterpret_cast<u32*>(slot + RC_ADAPT_HDR);
    sink.wcount = -1;
    sink.cap_words = has ? (u32)((a.slot_stride - RC_ADAPT_HDR) / 4u) : 0u;
    sink.err = a.err;

    // ---- warp 1: the coder of one tile <- pairs
    auto code_tile = [&](u32 k) {
        if(nrec && k != 0u && k % seg_tiles == 0u && k * TILE < n_b) {
            u32* rec = a.restart + ((b0 + lane) * nrec + k / seg_tiles - 1u) * (u64)ADAPT_REC_WORDS;
            const u32 words = (u32)(sink.wcount + 1) + st.nff;  // words cut off the shift register so far
            rec[0] = 4u * words + (u32)st.ocnt / 8u - 1u;       // bytes shifted out of low, the dummy byte aside
            rec[1] = st.low;
            rec[2] = st.range;
        }
        const u32 d0 = 256u + k * TILE;
        const u32 mg0 = rc_magic(d0 + lane);
        const u32 mg1 = rc_magic(d0 + 32u + lane);
        const u32 row = src_a + (k % 3u) * TILE_BYTES + lane * ROW;
        const u32 in = pairs_a + (k & 1u) * (TILE * 128u) + lane * 4u;
#pragma unroll 1
        for(int wi = 0; wi < TILE / 4; ++wi) {
            const u32 word = lds32(row + 4u * (u32)wi);
            const u32 mg = wi < 8 ? mg0 : mg1;
            RcCut cuts[4];
#pragma unroll
            for(int q = 0; q < 4; ++q) {
                const int j = wi * 4 + q;
                const u32 magic = __shfl_sync(FULL, mg, j & 31);
                const bool active = k * TILE + (u32)j < n_b;
                u32 cum = 0, freq = 1;
                if(active) {
                    const u32 p = lds32v(in + (u32)j * 128u);
                    cum = (p & 0xFFFFu) + ((word >> (8 * q)) & 0xFFu);
                    freq = (p >> 16) + 1u;
                }
                const u32 t = rc_div(st.range, d0 + (u32)j, magic);
                rc_enc_step<3>(st, cum, freq, t, cuts[q], active);
            }
            rc_enc_commit(st, cuts, sink);
        }
    };

    if(warp == 0) {
        cp_async_wait<1>();  // tile 0 has arrived
        __syncwarp();
        model_tile(0);
    }
    __syncthreads();
#pragma unroll 1
    for(u32 k = 0; k < ntiles; ++k) {
        if(warp == 0) {
            if(k + 2 < ntiles) {
                stage_tile(src_a + ((k + 2) % 3u) * TILE_BYTES, a.src, a.n, b0, a.block, (k + 2) * TILE, lane);
            }
            cp_async_commit();
            if(k + 1 < ntiles) {
                cp_async_wait<1>();  // everything but the newest group: tile k + 1 is there
                __syncwarp();
                model_tile(k + 1);
            }
        } else {
            code_tile(k);
        }
        __syncthreads();
    }
    if(warp == 1) {
        finish_block(st, sink, has, RC_ADAPT_HDR, n_b, slot, a.sizes + b);
    }
}

// ------------------------------------------------------------------ leafless tree --
// Layout: the two CHILDREN of node j share one 32-bit word, [j][lane] (bank == lane: no conflicts, the
// u16 [node][lane] layout of the full tree has two lanes per bank word); node 2j in the low half, 2j+1
// in the high half, so the root (node 1) is the high half of word 0.  Arriving at a node the walk
// holds its count and the word with both children; one level is: multiply, compare, the bookkeeping
// of the side taken, one 16-bit store (the node), one 32-bit load (the children of the child).
struct Leafless {
    static constexpr u32 S = 128u;  // bytes between consecutive words of one lane (u32 [128][32])
    static constexpr u32 LANE = 4u;
    static constexpr u32 BYTES = 256u * 32u * 2u;
    static constexpr u32 REC = ADAPT_REC_WORDS;
    static __device__ __forceinline__ u32 total(const u32* leaves)
    {
        u32 sum = 0;
        for(u32 k = 0; k < 128u; ++k) {
            const u32 pair = __ldg(leaves + k);
            sum += (pair & 0xFFFFu) + (pair >> 16);
        }
        return sum;
    }
    static __device__ __forceinline__ u32 node_addr(u32 base, u32 id) { return base + (id >> 1) * S + (id & 1u) * 2u; }

    // `rem` = low minus everything known to lie below the symbol (times t); `v` the node's count, `kids`
    // the word with the counts of its two children, `ka` that word's address, `na` the node's own;
    // `sub` how many symbols so far lie in this node's subtree.  The words with the GRANDchildren are
    // requested here, two levels ahead of their use: no shared-memory latency is left on the chain of
    // the walk, which is per level multiply-add, compare, select.
    template <int L>
    static __device__ __forceinline__ void level(u32 base, u32 t, u32& rem, u32& v, u32& kids, u32& na, u32& ka, u32& sub)
    {
        u32 ga, gl = 0, gr = 0;  // ga = base + 128 * (2 id): the children of the left child; the right child's follow
        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(ga) : "r"(ka), "r"(0u - base));
        if(L >= 2) {
            gl = lds32v(ga);
            gr = lds32v(ga + S);
        }
        const u32 prod = v * t + (t << L);  // left subtree: counts + the implicit one per symbol, times t
        u32 vnext, knext, nan, kan;
        // outputs that are written before the last input is read are early-clobber: they must not share a register
        asm("{ .reg .pred p;\n\t.reg .u32 hi, lo;\n\t"
            "setp.le.u32 p, %7, %0;\n\t"
            "@p sub.u32 %0, %0, %7;\n\t@p sub.u32 %2, %2, %1;\n\t@!p mov.u32 %2, %1;\n\t@!p add.u32 %1, %1, 1;\n\t"
            "shr.u32 hi, %8, 16;\n\tand.b32 lo, %8, 0xFFFF;\n\tselp.u32 %3, hi, lo, p;\n\t"
            "selp.u32 %4, %12, %11, p;\n\t"
            "selp.u32 %5, 2, 0, p;\n\tadd.u32 %5, %5, %9;\n\t"
            "selp.u32 %6, 128, 0, p;\n\tadd.u32 %6, %6, %10; }"
            : "+r"(rem), "+r"(v), "+r"(sub), "=&r"(vnext), "=&r"(knext), "=&r"(nan), "=&r"(kan)
            : "r"(prod), "r"(kids), "r"(ka), "r"(ga), "r"(gl), "r"(gr));
        sts16v(na, v);  // incremented when the symbol went left, unchanged otherwise
        na = nan;
        ka = kan;
        v = vnext;
        kids = knext;
    }

    // AdaptiveFrequencyTable::find (cpprcoder.h:1221-1241) in the product domain + update (:1134-1177,
    // for a model that never halves).  `seen` = symbols decoded so far in the block.  Returns the symbol;
    // `rem` comes back as low - cum*t, freq as the symbol's frequency.
    static __device__ __forceinline__ u32 decode(u32 base, u32 t, u32 seen, u32& rem, u32& freq)
    {
        u32 na = base + 2u, ka = base + S;  // node 1, the root, and the word of its children (nodes 2, 3)
        u32 v = lds16v(na), kids = lds32v(ka), sub = seen;
        level<7>(base, t, rem, v, kids, na, ka, sub);
        level<6>(base, t, rem, v, kids, na, ka, sub);
        level<5>(base, t, rem, v, kids, na, ka, sub);
        level<4>(base, t, rem, v, kids, na, ka, sub);
        level<3>(base, t, rem, v, kids, na, ka, sub);
        level<2>(base, t, rem, v, kids, na, ka, sub);
        level<1>(base, t, rem, v, kids, na, ka, sub);
        // the last level: the node's two children are symbols; ka = base + 128 * (its number)
        const u32 prod = (v + 1u) * t;
        u32 bit;
        asm("{ .reg .pred p;\n\tsetp.le.u32 p, %4, %0;\n\t@p sub.u32 %0, %0, %4;\n\t@p sub.u32 %2, %2, %1;\n\t"
            "@!p mov.u32 %2, %1;\n\t@!p add.u32 %1, %1, 1;\n\tselp.u32 %3, 1, 0, p; }"
            : "+r"(rem), "+r"(v), "+r"(sub), "=r"(bit)
            : "r"(prod));
        sts16v(na, v);
        freq = sub + 1u;
        return (((ka - base) >> 6) | bit) & 255u;
    }

    // The lane's tree from 256 symbol counts (two per word at `leaves`; null: an empty model).
    static __device__ __forceinline__ void build(u32 base, const u32* leaves)
    {
        if(!leaves) {
            for(u32 j = 0; j < 128u; ++j) {
                sts32v(base + j * S, 0u);
            }
            return;
        }
        // totals bottom up, in place ...
        for(u32 id = 255u; id >= 128u; --id) {
            const u32 pair = __ldg(leaves + (id - 128u));
            sts16v(node_addr(base, id), (pair & 0xFFFFu) + (pair >> 16));
        }
        for(u32 id = 127u; id >= 1u; --id) {
            const u32 kids = lds32v(base + id * S);  // nodes 2 id and 2 id + 1
            sts16v(node_addr(base, id), (kids & 0xFFFFu) + (kids >> 16));
        }
        // ... then every node takes its LEFT child's total, parents before their children
        for(u32 id = 1u; id < 128u; ++id) {
            sts16v(node_addr(base, id), lds32v(base + id * S) & 0xFFFFu);
        }
        for(u32 id = 128u; id < 256u; ++id) {
            sts16v(node_addr(base, id), __ldg(leaves + (id - 128u)) & 0xFFFFu);
        }
        sts16v(base, 0u);
    }
};

// The same tree with 32-bit counts, for blocks above 65536 bytes: the two children of node j are the two words
// of an 8-byte pair, [j][lane] (one 64-bit load; the two half-warps take a wavefront each, no conflicts).
// 32 KiB per warp, six warps per SM.  Written plainly: the narrow tree is the hot one.
struct LeaflessW {
    static constexpr u32 S = 256u;  // bytes between consecutive pairs of one lane (u32 [128][32][2])
    static constexpr u32 LANE = 8u;
    static constexpr u32 BYTES = 128u * 32u * 8u;
    static constexpr u32 REC = ADAPT_REC_WORDS_WIDE;
    static __device__ __forceinline__ u32 node_addr(u32 base, u32 id) { return base + (id >> 1) * S + (id & 1u) * 4u; }
    static __device__ __forceinline__ void ld2(u32 a, u32& lo, u32& hi)
    {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a));
    }

    template <int L>
    static __device__ __forceinline__ void level(u32 base, u32 t, u32& rem, u32& v, u32& k_lo, u32& k_hi, u32& na,
                                                 u32& ka, u32& sub)
    {
        const u32 ga = 2u * ka - base;  // pair 2 id: the children of the left child; the right child's follow
        u32 gl_lo = 0, gl_hi = 0, gr_lo = 0, gr_hi = 0;
        if(L >= 2) {
            ld2(ga, gl_lo, gl_hi);
            ld2(ga + S, gr_lo, gr_hi);
        }
        const u32 prod = v * t + (t << L);  // left subtree: counts + the implicit one per symbol, times t
        const bool right = prod <= rem;
        rem -= right ? prod : 0u;
        sub = right ? sub - v : v;
        v += right ? 0u : 1u;
        sts32v(na, v);  // incremented when the symbol went left, unchanged otherwise
        v = right ? k_hi : k_lo;
        na = ka + (right ? 4u : 0u);
        ka = ga + (right ? S : 0u);
        k_lo = right ? gr_lo : gl_lo;
        k_hi = right ? gr_hi : gl_hi;
    }

    static __device__ __forceinline__ u32 decode(u32 base, u32 t, u32 seen, u32& rem, u32& freq)
    {
        u32 na = base + 4u, ka = base + S;  // node 1, the root, and the pair of its children (nodes 2, 3)
        u32 v = lds32v(na), k_lo, k_hi, sub = seen;
        ld2(ka, k_lo, k_hi);
        level<7>(base, t, rem, v, k_lo, k_hi, na, ka, sub);
        level<6>(base, t, rem, v, k_lo, k_hi, na, ka, sub);
        level<5>(base, t, rem, v, k_lo, k_hi, na, ka, sub);
        level<4>(base, t, rem, v, k_lo, k_hi, na, ka, sub);
        level<3>(base, t, rem, v, k_lo, k_hi, na, ka, sub);
        level<2>(base, t, rem, v, k_lo, k_hi, na, ka, sub);
        level<1>(base, t, rem, v, k_lo, k_hi, na, ka, sub);
        // the last level: the node's two children are symbols; ka = base + 256 * (its number)
        const u32 prod = (v + 1u) * t;
        const bool right = prod <= rem;
        rem -= right ? prod : 0u;
        sub = right ? sub - v : v;
        v += right ? 0u : 1u;
        sts32v(na, v);
        freq = sub + 1u;
        return (((ka - base) >> 7) | (right ? 1u : 0u)) & 255u;
    }

    // The lane's tree from 256 symbol counts (one per word at `leaves`; null: an empty model).
    static __device__ __forceinline__ void build(u32 base, const u32* leaves)
    {
        if(!leaves) {
            for(u32 j = 0; j < 256u; ++j) {
                sts32v(base + (j >> 1) * S + (j & 1u) * 4u, 0u);
            }
            return;
        }
        // totals bottom up, in place ...
        for(u32 id = 255u; id >= 128u; --id) {
            sts32v(node_addr(base, id), __ldg(leaves + 2u * (id - 128u)) + __ldg(leaves + 2u * (id - 128u) + 1u));
        }
        for(u32 id = 127u; id >= 1u; --id) {
            sts32v(node_addr(base, id), lds32v(node_addr(base, 2u * id)) + lds32v(node_addr(base, 2u * id + 1u)));
        }
        // ... then every node takes its LEFT child's total, parents before their children
        for(u32 id = 1u; id < 128u; ++id) {
            sts32v(node_addr(base, id), lds32v(node_addr(base, 2u * id)));
        }
        for(u32 id = 128u; id < 256u; ++id) {
            sts32v(node_addr(base, id), __ldg(leaves + 2u * (id - 128u)));
        }
        sts32v(base, 0u);
    }
    // sum of the counts of a point (must equal its position in the block)
    static __device__ __forceinline__ u32 total(const u32* leaves)
    {
        u32 sum = 0;
        for(u32 k = 0; k < 256u; ++k) {
            sum += __ldg(leaves + k);
        }
        return sum;
    }
};

constexpr u32 DEC_ADAPT_SEG_SMEM = 256u * 32u * 2u + TILE_BYTES + INQ_BYTES;
constexpr u32 DEC_ADAPT_SEG_SMEM_WIDE = LeaflessW::BYTES + TILE_BYTES + INQ_BYTES;

template <class Tree, class Src>
__device__ __forceinline__ void dec_adaptive_seg_tile(u32 tbase, RcDec& d, Src& src, u32 otile_a, u32 tile_off, u32 n_b,
                                                      u32 lane)
{
    const u32 d0 = 256u + tile_off;
    const u32 mg0 = rc_magic(d0 + lane);
    const u32 mg1 = rc_magic(d0 + 32u + lane);
#pragma unroll 1
    for(int wi = 0; wi < TILE / 4; ++wi) {
        const u32 mg = wi < 8 ? mg0 : mg1;
        u32 word = 0;
#pragma unroll
        for(int k = 0; k < 4; ++k) {
            const int j = wi * 4 + k;
            const u32 magic = __shfl_sync(FULL, mg, j & 31);
            if(tile_off + j < n_b) {
                const u32 t = rc_div(d.range, d0 + j, magic);
                u32 freq;
                const u32 sym = Tree::decode(tbase, t, tile_off + j, d.low, freq);  // d.low -= cum * t on the way
                rc_dec_advance(d, 0u, freq, t, src);
                word |= sym << (8 * k);
            }
        }
        sts32v(otile_a + lane * ROW + wi * 4, word);
    }
}

// grid = (ceil(nblocks / 32), segments per block); one warp per CTA.  Tree = Leafless for blocks of at most
// 65536 bytes (16-bit counts), LeaflessW above.
template <class Tree>
__global__ void __launch_bounds__(32) k_dec_adaptive_seg(DecArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    constexpr u32 TAB_BYTES = Tree::BYTES;
    const u32 sbase = smem_addr(smem);
    u8* otile = smem + TAB_BYTES;
    const u32 otile_a = sbase + TAB_BYTES;
    const u32 queue_a = otile_a + TILE_BYTES;

    const u32 lane = lane_id();
    const u64 b0 = (u64)blockIdx.x * 32u;
    const u64 b = b0 + lane;
    const bool has = b < a.nblocks;
    u32 n_b = 0;
    if(has) {
        const u64 lo = b * (u64)a.block;
        n_b = (u32)((a.n - lo < a.block) ? (a.n - lo) : a.block);
    }
    const u32 nseg = (a.block + a.seg_syms - 1u) / a.seg_syms;
    const u32 nrec = nseg - 1u;
    const u32 seg = blockIdx.y;

    const u8* pay = a.payload;
    u64 len = 0;
    if(has) {
        const u64 o0 = a.offsets[b], o1 = a.offsets[b + 1];
        if(o0 <= o1 && o1 <= a.payload_len) {
            pay = a.payload + o0;
            len = o1 - o0;
        }
    }
    bool ok = has && len >= (u64)RC_ADAPT_HDR + 5u;
    if(ok) {
        const u32 want = (u32)pay[0] | ((u32)pay[1] << 8) | ((u32)pay[2] << 16) | ((u32)pay[3] << 24);
        ok = want == n_b;  // the container, not the payload, says how long block b is
    }
    if(has && !ok && seg == 0u) {
        atomicOr(a.err, ERR_CORRUPT);
    }
    const u32 seg_lo = seg * a.seg_syms;
    u32 seg_hi = seg_lo + a.seg_syms;
    seg_hi = seg_hi < n_b ? seg_hi : n_b;  // my symbols: [seg_lo, seg_hi)
    bool mine_ok = ok && seg_lo < n_b;
    const u32 skip0 = (u32)((uintptr_t)(pay + RC_ADAPT_HDR) & 3u);
    u32 skip = skip0, word0 = 0, range0 = RC_ADAPT_RANGE0, enc_low = 0;
    const u32* leaves = nullptr;
    if(seg != 0u && mine_ok) {
        const u32* rec = a.restart + (b * nrec + seg - 1u) * (u64)Tree::REC;
        const u32 m = rec[0];
        enc_low = rec[1];
        range0 = rec[2];
        leaves = rec + 3;
        if(m == 0xFFFFFFFFu || (u64)m + RC_ADAPT_HDR + 5u > len) {
            mine_ok = false;
            atomicOr(a.err, ERR_CORRUPT);
        } else {
            // the byte at offset m of the coded stream plays the part of the dummy first byte
            word0 = (skip + m) >> 2;
            skip = (skip + m) & 3u;
        }
    }
    const u32 tbase = sbase + lane * Tree::LANE;
    Tree::build(tbase, mine_ok ? leaves : nullptr);
    if(mine_ok && leaves) {
        // the counts of a point must add up to the symbols in front of it, or the walk's bookkeeping
        // (and with it every frequency) would be off: the root's total is checked against the position
        if(Tree::total(leaves) != seg_lo) {
            mine_ok = false;
            atomicOr(a.err, ERR_CORRUPT);
        }
    }
    WordSrc src;
    {
        const u8* coded = pay + RC_ADAPT_HDR;
        const u8* wbase = (const u8*)((uintptr_t)coded & ~(uintptr_t)3);
        src.base = reinterpret_cast<const u32*>(wbase);
        const u64 room = (u64)((a.payload + a.payload_len) - wbase);
        src.lim = mine_ok ? (u32)(room < 0xFFFFFFF0ull ? room : 0xFFFFFFF0ull) : 0u;
        src.q = queue_a + lane * 4u;
        src.prime(mine_ok ? word0 : 0u);
    }
    __syncwarp();
    const u32 n_eff = mine_ok ? seg_hi : 0u;
    RcDec d;
    rc_dec_init(d, range0, skip, src);
    d.low -= enc_low;

    const u32 n_max = __reduce_max_sync(FULL, n_eff);
    const u32 tix0 = seg_lo / TILE;
    const u32 ntiles = (n_max + TILE - 1) / TILE;
    const u32 tix1 = ntiles > tix0 ? ntiles : tix0;
#pragma unroll 1
    for(u32 tix = tix0; tix < tix1; ++tix) {
        if(__all_sync(FULL, src.tile_is_inside())) {
            WordSrcInside in{src};
            dec_adaptive_seg_tile<Tree>(tbase, d, in, otile_a, tix * TILE, n_eff, lane);
        } else {
            dec_adaptive_seg_tile<Tree>(tbase, d, src, otile_a, tix * TILE, n_eff, lane);
        }
        __syncwarp();
        // only this segment's columns of the tile are this warp's to store
        store_tile(otile, a.dst, a.n, b0, a.block, tix * TILE, lane);
        __syncwarp();
    }

    // a segment has to end where the next point stands; the last one with the coded bytes used up
    if(mine_ok) {
        const u32 used = 4u * src.rd - skip0 - (u32)d.wbits / 8u;
        bool good;
        if(seg_hi < n_b) {
            const u32* rec = a.restart + (b * nrec + seg) * (u64)Tree::REC;
            const u32 m = rec[0];
            good = m != 0xFFFFFFFFu && (u64)m + RC_ADAPT_HDR + 5u <= len && used == m + 5u;
            if(good) {
                const u8* p = pay + RC_ADAPT_HDR + 1u + m;
                const u32 be = ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3];
                good = d.low == be - rec[1] && d.range == rec[2];
            }
        } else {
            good = (u64)used + RC_ADAPT_HDR == len;
        }
        if(!good) {
            atomicOr(a.err, ERR_CORRUPT);
        }
    }
}

}  // namespace b2rc
