// b2rc_kernels.cuh -- the sm_100a kernels of the block range coder.
//
//   K1  k_hist            per-block byte histogram + the reference's count scaling
//   K2  k_enc_static      static encode      (RangeEncoder::encode,          cpprcoder.h:375-458)
//       k_enc_adaptive    adaptive encode    (AdaptiveRangeEncoder,          cpprcoder.h:678-802)
//   K3  k_dec_static      static decode      (RangeEncoder::decode,          cpprcoder.h:460-535)
//       k_dec_adaptive    adaptive decode    (AdaptiveRangeDecoder,          cpprcoder.h:859-940)
//   K4  k_scan, k_compact exclusive scan of payload sizes + compaction into one stream
//
// Mapping: one block per LANE, one warp per CTA, 32
// consecutive blocks per warp.  The coder's per-symbol chain is serial per block,
// so throughput comes from the number of chains in flight and from how few
// instructions one step of a (lone) warp needs -- not from splitting a block over a
// warp.  Per-warp shared memory holds
//   * the model table, interleaved by lane ([entry][lane]: bank == lane),
//   * input tiles staged by cp.async (encode) / an output tile (decode), moved by the
//     whole warp in 16-byte vectors, 8 rows of 64 bytes per instruction.
// The variable-rate side of each coder (coded words out of the encoder, into the decoder)
// is per lane: 4-byte accesses that L2 / L1 merge into sectors (see SlotSink, WordSrc).
// Hot-loop shared-memory accesses use explicit ld/st.shared on 32-bit shared
// addresses: generic pointers made the compiler rebuild the shared window base
// (S2UR SR_CgaCtaId) on every step and turned ring stores into generic ST.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "rc_lane.cuh"

namespace b2rc
{
constexpr u32 FULL = 0xFFFFFFFFu;
constexpr int TILE = 64;        // symbols per lane per staged tile
constexpr int ROW = 80;         // bytes per lane row of a tile (64 + 16 pad; rows stay 16-byte aligned)
constexpr int TILE_BYTES = 32 * ROW;

enum : int {
    ERR_SLOT_OVERFLOW = 1,  // a payload outgrew its slot
    ERR_CORRUPT = 2,        // decoder met an impossible header / state
    ERR_DST_SMALL = 4,      // compaction ran out of destination
};

// --------------------------------------------------------------- small helpers --
__device__ __forceinline__ u32 smem_addr(const void* p)
{
    return (u32)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(u32 dst, const void* src, u32 bytes)
{
    // 16-byte async copy global -> shared, zero-filling beyond `bytes` (LDGSTS)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ u32 lane_id()
{
    return threadIdx.x & 31u;
}
// read-only data (static tables, staged input): plain asm, free to be hoisted
__device__ __forceinline__ u32 lds32(u32 a)
{
    u32 v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ u32 lds16(u32 a)
{
    u32 v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));  // a wider destination is zero extended: no cvt
    return v;
}
__device__ __forceinline__ u32 lds8(u32 a)
{
    u32 v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// read-write data (adaptive model, rings): volatile keeps program order among them
__device__ __forceinline__ u32 lds32v(u32 a)
{
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ u32 lds16v(u32 a)
{
    u32 v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32v(u32 a, u32 v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v));
}
__device__ __forceinline__ void sts16v(u32 a, u32 v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v));  // a wider source is truncated: no cvt
}

// Stage tile `tile_off .. tile_off+TILE` of 32 consecutive blocks into shared memory.
// Row r of the tile is block b0+r; 4 lanes cover one 64-byte row, 8 rows per instruction.
// Bytes past the end of the block / of the input are zero-filled by cp.async's src-size.
__device__ __forceinline__ void stage_tile(u32 tile, const u8* src, u64 n, u64 b0, u32 block, u32 tile_off, u32 lane)
{
    constexpr int CH = TILE / 16;
    constexpr int RPI = 32 / CH;
#pragma unroll
    for(int it = 0; it < 32 / RPI; ++it) {
        const u32 row = it * RPI + lane / CH;
        const u32 ch = lane % CH;
        const u64 blk_lo = (b0 + row) * (u64)block;
        u64 blk_hi = blk_lo + block;
        if(blk_hi > n) {
            blk_hi = n;
        }
        const u64 at = blk_lo + tile_off + ch * 16u;
        u32 bytes = 0;
        if(at < blk_hi) {
            const u64 rem = blk_hi - at;
            bytes = rem >= 16 ? 16u : (u32)rem;
        }
        cp_async16(tile + row * ROW + ch * 16u, bytes ? (const void*)(src + at) : (const void*)src, bytes);
    }
}

// ------------------------------------------------------------ encoder output --
// A committed word is final (carries only ever reach the deferred word, rc_lane.cuh), so
// each lane stores its words straight into its own staging slot: 4-byte stores to 32
// different lines per instruction, which the L2 merges into full sectors long before they
// are evicted.  An earlier version staged words in a shared-memory ring and flushed all
// 32 rings cooperatively; the 32 serial shuffle/load/store rounds of that flush cost a
// third of the kernel (profiles/r1_ncu_notes.md).
// `wcount` starts at -1: the encoder's first push is its placeholder word (Sink contract)
// and lands on the 4 bytes in front of the coded stream, which finish_block() restores.
struct CheckedSlotSink;
struct SlotSink {
    typedef CheckedSlotSink Checked;
    u32* out;  // coded stream start of this lane's block (slot + header), 4-byte aligned
    s32 wcount;
    u32 cap_words;
    int* err;

    __device__ __forceinline__ bool tight(int n) const { return (u32)(wcount + n) > cap_words; }
    __device__ __forceinline__ void push(u32 w)  // hot path: rc_enc_commit checked tight() first
    {
        out[wcount] = rc_bswap(w);
        ++wcount;
    }
};

struct CheckedSlotSink {
    SlotSink s;
    __device__ explicit CheckedSlotSink(const SlotSink& r) : s(r) {}
    __device__ void push(u32 w)
    {
        if(s.wcount < 0 || (u32)s.wcount < s.cap_words) {
            s.out[s.wcount] = rc_bswap(w);
        } else {
            atomicOr(s.err, ERR_SLOT_OVERFLOW);
        }
        ++s.wcount;
    }
    __device__ void settle(SlotSink& r) { r.wcount = s.wcount; }
};

struct EncArgs {
    const u8* src;
    u64 n;
    u32 block;
    u64 nblocks;
    const u16* freq16;
    u8* slots;
    u64 slot_stride;
    u32* sizes;
    int* err;
    // Restart points of the static coder: before symbol j * seg_syms of block
    // b, j >= 1, the encoder records {bytes shifted out so far, low, range} at
    // restart[(b * nrec + j - 1) * 3], nrec = ceil(block / seg_syms) - 1.  Null: none.
    u32* restart;
    u32 seg_syms;
};

// Tail of a block: the 4..7 bytes that do not fill a word, then the size.  `front` is what
// the 4 bytes in front of the coded stream must hold (the placeholder word overwrote them).
__device__ __forceinline__ void finish_block(RcEnc& st, SlotSink& sink, bool has, u32 hdr, u32 front, u8* slot,
                                             u32* size_out)
{
    if(has) {
        u8 tail[8];
        CheckedSlotSink cs(sink);
        const u32 ntail = rc_enc_finish(st, cs, tail);
        sink.wcount = cs.s.wcount;
        sink.out[-1] = front;
        const u32 at = hdr + 4u * (u32)sink.wcount;
        if((u32)sink.wcount <= sink.cap_words && (u64)at + ntail <= (u64)hdr + 4ull * sink.cap_words) {
            for(u32 k = 0; k < ntail; ++k) {
                slot[at + k] = tail[k];
            }
        } else {
            atomicOr(sink.err, ERR_SLOT_OVERFLOW);
        }
        *size_out = at + ntail;
    }
}

// ======================================================================= K2s ==
// Static encode.  WIDE = false: blocks <= 65536 bytes; cum and freq of a coded symbol
// both fit 16 bits and live in two u16 arrays built from K1's
// frequencies.  WIDE = true: larger blocks; the lane counts its own block with the
// reference's order-dependent rule (cpprcoder.h:543-571) and keeps a 257-entry u32
// cum table (freq = difference of neighbours).
constexpr u32 ENC_STATIC_TAB_NARROW = 2u * 256u * 32u * 2u;  // cum16[256][32] then freq16[256][32]
constexpr u32 ENC_STATIC_TAB_WIDE = 257u * 32u * 4u;

template <bool WIDE>
struct StaticTab {
    u32 base;  // shared address, already offset by the lane
    __device__ __forceinline__ void get(u32 sym, u32& cum, u32& freq) const
    {
        if(WIDE) {
            const u32 a = base + sym * 128u;
            cum = lds32(a);
            freq = lds32(a + 128u) - cum;
        } else {
            const u32 a = base + sym * 64u;
            cum = lds16(a);
            freq = lds16(a + 256u * 64u);
        }
    }
};

template <bool WIDE, bool POW2, bool RAGGED>
__device__ __forceinline__ void enc_static_tiles(const EncArgs& a, u32 tiles, const StaticTab<WIDE>& tab, RcEnc& st,
                                                 SlotSink& sink, u64 b0, u32 n_b, u32 n_max, u32 total, u32 magic,
                                                 u32 shift, u32 lane)
{
    const u32 ntiles = (n_max + TILE - 1) / TILE;
    u32 tcur = POW2 ? (st.range >> shift) : 0u;  // the power-of-two chain carries t, not range
    const u32 seg_tiles = a.restart ? a.seg_syms / TILE : 0u;
    u32 next_mark = a.restart ? seg_tiles : 0xFFFFFFFFu;  // tile index of the next restart point
    // tile 0 was staged (and committed) by the caller
#pragma unroll 1
    for(u32 tix = 0; tix < ntiles; ++tix) {
        if(tix + 1 < ntiles) {
            stage_tile(tiles + ((tix + 1) & 1u) * TILE_BYTES, a.src, a.n, b0, a.block, (tix + 1) * TILE, lane);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        if(tix == next_mark) {
            // a segment starts here: what a decoder needs to start here too
            next_mark += seg_tiles;
            if(tix * TILE < n_b) {
                const u32 nrec = (a.block + a.seg_syms - 1u) / a.seg_syms - 1u;
                u32* rec = a.restart + ((b0 + lane) * nrec + tix / seg_tiles - 1u) * 3u;
                const u32 words = (u32)(sink.wcount + 1) + st.nff;  // words cut off the shift register so far
                rec[0] = 4u * words + (u32)st.ocnt / 8u - 1u;       // bytes shifted out of low, the dummy byte aside
                rec[1] = st.low;
                rec[2] = POW2 ? (tcur << shift) : ((st.range >> shift) << shift);  // one form whatever the warp's path (b2rc_encseg.cuh)
            }
        }
        const u32 row = tiles + (tix & 1u) * TILE_BYTES + lane * ROW;
        // Four symbols per trip.  The trip stays small on purpose: with one warp per SM
        // nothing hides an instruction fetch.  Table entries of the NEXT four symbols are
        // requested before the current four are coded, so their latency is off the chain.
        u32 word = lds32(row);
        u32 cum[4], freq[4];
#pragma unroll
        for(int k = 0; k < 4; ++k) {
            tab.get((word >> (8 * k)) & 0xFFu, cum[k], freq[k]);
        }
#pragma unroll 1
        for(int wi = 0; wi < TILE / 4; ++wi) {
            const u32 wnext = lds32(row + 4u * (u32)((wi + 1) & (TILE / 4 - 1)));
            u32 ncum[4], nfreq[4];
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                tab.get((wnext >> (8 * k)) & 0xFFu, ncum[k], nfreq[k]);
            }
            RcCut cuts[4];
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                const bool active = !RAGGED || tix * TILE + wi * 4 + k < n_b;
                if(POW2) {
                    rc_enc_step_pow2<WIDE ? 3 : 2>(st, tcur, shift, cum[k], freq[k], cuts[k], active);
                } else {
                    const u32 t = rc_div(st.range, total, magic);
                    rc_enc_step<WIDE ? 3 : 2>(st, cum[k], freq[k], t, cuts[k], active);
                }
            }
            rc_enc_commit(st, cuts, sink);
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                cum[k] = ncum[k];
                freq[k] = nfreq[k];
            }
        }
        __syncwarp();
    }
}

template <bool WIDE>
__global__ void __launch_bounds__(32) k_enc_static(EncArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    constexpr u32 TAB_BYTES = WIDE ? ENC_STATIC_TAB_WIDE : ENC_STATIC_TAB_NARROW;
    const u32 sbase = smem_addr(smem);
    const u32 tiles = sbase + TAB_BYTES;

    const u32 lane = lane_id();
    const u64 b0 = (u64)blockIdx.x * 32u;
    const u64 b = b0 + lane;
    const bool has = b < a.nblocks;
    u32 n_b = 0;
    if(has) {
        const u64 lo = b * (u64)a.block;
        n_b = (u32)((a.n - lo < a.block) ? (a.n - lo) : a.block);
    }
    u8* slot = a.slots + b * a.slot_stride;
    u32 total = 0;
    u32 front = 0;  // header bytes 512..515 (frequencies of symbols 254, 255)

    if(!WIDE) {
        // first input tile in flight while the tables are built
        stage_tile(tiles, a.src, a.n, b0, a.block, 0, lane);
        cp_async_commit();
        u16* cum16 = reinterpret_cast<u16*>(smem);
        u16* frq16 = cum16 + 256 * 32;
#pragma unroll 1
        for(u32 r = 0; r < 32; ++r) {
            if(b0 + r >= a.nblocks) {
                break;
            }
            // lane j holds the frequencies of symbols 8j .. 8j+7 of block b0+r
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.freq16 + (b0 + r) * 256u) + lane);
            const u32 f[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16,
                              v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
            u32 sum = 0;
#pragma unroll
            for(int k = 0; k < 8; ++k) {
                sum += f[k];
            }
            u32 incl = sum;
#pragma unroll
            for(int d = 1; d < 32; d <<= 1) {
                const u32 up = __shfl_up_sync(FULL, incl, d);
                if(lane >= (u32)d) {
                    incl += up;
                }
            }
            u32 run = incl - sum;  // calcCumulatives (cpprcoder.h:573-583)
#pragma unroll
            for(int k = 0; k < 8; ++k) {
                // a symbol that occurs has cum <= total - freq <= 65535; for one that does not,
                // the truncated value is never read
                cum16[(8u * lane + k) * 32u + r] = (u16)run;
                frq16[(8u * lane + k) * 32u + r] = (u16)f[k];
                run += f[k];
            }
            const u32 tot = __shfl_sync(FULL, incl, 31);
            const u32 last = __shfl_sync(FULL, v.w, 31);
            if(lane == r) {
                total = tot;
                front = last;
            }
            // payload header: u32 LE size, then write16 (cpprcoder.h:386-395, :604-619)
            u8* slot_r = a.slots + (b0 + r) * a.slot_stride;
            const u32 n_r = __shfl_sync(FULL, n_b, (int)r);
            u32* hw = reinterpret_cast<u32*>(slot_r + 4u + 16u * lane);
            hw[0] = v.x;
            hw[1] = v.y;
            hw[2] = v.z;
            hw[3] = v.w;
            if(lane == 0) {
                *reinterpret_cast<u32*>(slot_r) = n_r;
            }
        }
        __syncwarp();
    } else if(a.freq16) {
        // counts from k_hist_wide (order-dependent halving applied there); the cumulative table
        // needs 32 bits here, so each lane sums its own block's 256 counts
        u32* mine = reinterpret_cast<u32*>(smem) + lane;
        if(has) {
            *reinterpret_cast<u32*>(slot) = n_b;
            const u16* fq = a.freq16 + b * 256u;
            u32 run = 0;
#pragma unroll 1
            for(u32 s = 0; s < 256; s += 2) {
                const u32 pair = __ldg(reinterpret_cast<const u32*>(fq + s));
                *reinterpret_cast<u32*>(slot + 4u + 2u * s) = pair;
                mine[s * 32u] = run;
                run += pair & 0xFFFFu;
                mine[(s + 1u) * 32u] = run;
                run += pair >> 16;
            }
            mine[256u * 32u] = run;
            total = run;
            front = *reinterpret_cast<const u32*>(slot + RC_STATIC_HDR - 4u);
        } else {
            for(u32 s = 0; s < 257; ++s) {
                mine[s * 32u] = 0;
            }
        }
        __syncwarp();
        stage_tile(tiles, a.src, a.n, b0, a.block, 0, lane);
        cp_async_commit();
    } else {
        // no table given: count() here, each lane walking its own block once (the halving is
        // order dependent)
        u32* mine = reinterpret_cast<u32*>(smem) + lane;
        for(u32 s = 0; s < 257; ++s) {
            mine[s * 32u] = 0;
        }
        const u32 n_max0 = __reduce_max_sync(FULL, n_b);
        const u32 ntiles = (n_max0 + TILE - 1) / TILE;
        stage_tile(tiles, a.src, a.n, b0, a.block, 0, lane);
        cp_async_commit();
#pragma unroll 1
        for(u32 tix = 0; tix < ntiles; ++tix) {
            if(tix + 1 < ntiles) {
                stage_tile(tiles + ((tix + 1) & 1u) * TILE_BYTES, a.src, a.n, b0, a.block, (tix + 1) * TILE, lane);
            }
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            const u8* row = smem + TAB_BYTES + (tix & 1u) * TILE_BYTES + lane * ROW;
#pragma unroll 1
            for(u32 k = 0; k < (u32)TILE; ++k) {
                if(tix * TILE + k < n_b) {
                    const u32 c = row[k];
                    u32 f = mine[c * 32u];
                    if(f >= 0xFFFFu) {
                        for(u32 s = 0; s < 256; ++s) {
                            const u32 x = mine[s * 32u];
                            if(x) {
                                mine[s * 32u] = (x >> 1) | 1u;
                            }
                        }
                        f = mine[c * 32u];
                    }
                    mine[c * 32u] = f + 1u;
                }
            }
            __syncwarp();
        }
        if(has) {
            *reinterpret_cast<u32*>(slot) = n_b;
            u32 run = 0;
            for(u32 s = 0; s < 256; ++s) {
                const u32 f = mine[s * 32u];
                *reinterpret_cast<u16*>(slot + 4u + 2u * s) = (u16)f;
                mine[s * 32u] = run;
                run += f;
            }
            mine[256u * 32u] = run;
            total = run;
            front = *reinterpret_cast<const u32*>(slot + RC_STATIC_HDR - 4u);
        }
        __syncwarp();
        stage_tile(tiles, a.src, a.n, b0, a.block, 0, lane);
        cp_async_commit();
    }

    RcEnc st;
    rc_enc_init(st, RC_STATIC_RANGE0);
    SlotSink sink;
    sink.out = reinterpret_cast<u32*>(slot + RC_STATIC_HDR);
    sink.wcount = -1;
    sink.cap_words = has ? (u32)((a.slot_stride - RC_STATIC_HDR) / 4u) : 0u;
    sink.err = a.err;

    const StaticTab<WIDE> tab{sbase + lane * (WIDE ? 4u : 2u)};
    const u32 magic = rc_magic(total);
    const bool is_pow2 = total != 0 && (total & (total - 1u)) == 0;
    const u32 shift = is_pow2 ? 31u - rc_clz(total) : 0u;
    const u32 n_max = __reduce_max_sync(FULL, n_b);
    const bool all_pow2 = __all_sync(FULL, is_pow2 || !has);
    const bool ragged = __any_sync(FULL, n_b != n_max);

    if(all_pow2 && !ragged) {
        enc_static_tiles<WIDE, true, false>(a, tiles, tab, st, sink, b0, n_b, n_max, total, magic, shift, lane);
    } else if(!ragged) {
        enc_static_tiles<WIDE, false, false>(a, tiles, tab, st, sink, b0, n_b, n_max, total, magic, shift, lane);
    } else {
        enc_static_tiles<WIDE, false, true>(a, tiles, tab, st, sink, b0, n_b, n_max, total, magic, shift, lane);
    }

    // cpprcoder.h:439-451: when the block ends on low_ == 0xFFFFFFFF the reference bumps
    // the held byte but still writes low_ as FF FF FF FF -- not the big-endian sum.
    // Probability 2^-32 per block; redo such a block with the reference-shaped encoder.
    bool exact = false;
    if(has && st.low == 0xFFFFFFFFu) {
        exact = true;
        sink.out[-1] = front;
        u32 at = RC_STATIC_HDR;
        const u8* blk = a.src + b * (u64)a.block;
        const u32 cap = (u32)a.slot_stride;
        bool over = false;
        rc_static_encode_exact(
            n_b, total,
            [&](u32 c) -> u32 {
                if(c >= 256u) {
                    return total;
                }
                // walk the frequencies: the 16-bit cum array is only exact for symbols that occur
                u32 cum, freq;
                if(WIDE) {
                    tab.get(c, cum, freq);
                    return cum;
                }
                u32 run = 0;
                for(u32 s = 0; s < c; ++s) {
                    run += lds16(tab.base + s * 64u + 256u * 64u);
                }
                return run;
            },
            [&](u32 i) -> u32 { return blk[i]; },
            [&](u8 byte) {
                if(at < cap) {
                    slot[at] = byte;
                } else {
                    over = true;
                }
                ++at;
            },
            [&](u32 i, u32 shifted, u32 low, u32 range) {
                if(a.restart && i != 0u && i % a.seg_syms == 0u) {
                    const u32 nrec = (a.block + a.seg_syms - 1u) / a.seg_syms - 1u;
                    u32* rec = a.restart + (b * nrec + i / a.seg_syms - 1u) * 3u;
                    rec[0] = shifted;
                    rec[1] = low;
                    rec[2] = (range >> shift) << shift;
                }
            });
        if(over) {
            atomicOr(a.err, ERR_SLOT_OVERFLOW);
        }
        a.sizes[b] = at;
    }
    finish_block(st, sink, has && !exact, RC_STATIC_HDR, front, slot, a.sizes + b);
}

// ======================================================================= K2a ==
// Adaptive encode.  The model is the count tree of rc_model_encode, one per lane
// ([node][lane], u16 for blocks <= 65536 bytes, u32 above).  total_i = 256 + i is the
// same for every lane, so the 64 magics of a tile are computed once per warp (two per
// lane) and handed round by shuffle.
template <class W>
struct LaneTab {
    u32 base;  // shared address, already offset by the lane
    __device__ __forceinline__ u32 ld(u32 i) const
    {
        return sizeof(W) == 2 ? lds16v(base + i * 64u) : lds32v(base + i * 128u);
    }
    __device__ __forceinline__ void st(u32 i, u32 v) const
    {
        if(sizeof(W) == 2) {
            sts16v(base + i * 64u, v);
        } else {
            sts32v(base + i * 128u, v);
        }
    }
};

// Hand-scheduled forms of rc_model_encode / rc_model_decode (rc_lane.cuh holds the portable
// statements of the same walks, which tests/sim checks against the oracle).  The compiler
// turned the per-level "bit set? add : increment" into 7 mask-arithmetic instructions and
// rebuilt every node address from shifts and masks (163 instructions per symbol in the
// encoder, profiles/r1_ncu_notes.md); written out with predicates a level is: address (2),
// load, test+two predicated adds (3), store.
template <class W>
struct TreeOps {
    static constexpr u32 S = 32u * sizeof(W);  // bytes between consecutive nodes of one lane
    static __device__ __forceinline__ u32 ld(u32 a) { return sizeof(W) == 2 ? lds16v(a) : lds32v(a); }
    static __device__ __forceinline__ void st(u32 a, u32 v)
    {
        if(sizeof(W) == 2) {
            sts16v(a, v);
        } else {
            sts32v(a, v);
        }
    }
    template <int L>
    static __device__ __forceinline__ u32 node_addr(u32 base, u32 leaf)  // node (leaf >> (L+1)) of the lane at `base`
    {
        u32 idx, a;
        asm("shr.u32 %0, %1, %2;" : "=r"(idx) : "r"(leaf), "n"(L + 1));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(idx), "n"(S), "r"(base));
        return a;
    }
    template <int L>
    static __device__ __forceinline__ void enc_level(u32 b, u32& below, u32& x)
    {
        asm("{ .reg .pred p; .reg .u32 m;\n\tand.b32 m, %2, %3;\n\tsetp.ne.u32 p, m, 0;\n\t@p add.u32 %0, %0, %1;\n\t@!p add.u32 %1, %1, 1; }"
            : "+r"(below), "+r"(x)
            : "r"(b), "n"(1u << L));
    }
    // cum / freq of b under the current counts, then count b (AdaptiveFrequencyTable::cumulative
    // + update, cpprcoder.h:1134-1187, for a model that never halves)
    static __device__ __forceinline__ void encode(u32 base, u32 b, u32& cum, u32& freq)
    {
        const u32 leaf = 256u | b;
        u32 a[8], v[8];
        a[7] = node_addr<7>(base, leaf);
        a[6] = node_addr<6>(base, leaf);
        a[5] = node_addr<5>(base, leaf);
        a[4] = node_addr<4>(base, leaf);
        a[3] = node_addr<3>(base, leaf);
        a[2] = node_addr<2>(base, leaf);
        a[1] = node_addr<1>(base, leaf);
        a[0] = node_addr<0>(base, leaf);
        u32 al;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(al) : "r"(leaf), "n"(S), "r"(base));
#pragma unroll
        for(int l = 7; l >= 0; --l) {
            v[l] = ld(a[l]);
        }
        const u32 f = ld(al);
        u32 below = b;  // the implicit one per symbol below b
        enc_level<7>(b, below, v[7]);
        enc_level<6>(b, below, v[6]);
        enc_level<5>(b, below, v[5]);
        enc_level<4>(b, below, v[4]);
        enc_level<3>(b, below, v[3]);
        enc_level<2>(b, below, v[2]);
        enc_level<1>(b, below, v[1]);
        enc_level<0>(b, below, v[0]);
#pragma unroll
        for(int l = 7; l >= 0; --l) {
            st(a[l], v[l]);
        }
        st(al, f + 1u);
        cum = below;
        freq = f + 1u;
    }

    // One level of the decoder's walk.  `rem` = low minus everything already known to lie
    // below the symbol (times t); `aid` = address of the current node, `v` its count, `cl`/`cr`
    // the counts of its two children.  The four GRANDCHILDREN are read here, two levels ahead
    // of their use, so that no table latency is left on the compare chain of the walk.
    template <int L>
    static __device__ __forceinline__ void dec_level(u32 base, u32 t, u32& rem, u32& aid, u32& v, u32& cl, u32& cr)
    {
        u32 g0 = 0, g1 = 0, g2 = 0, g3 = 0;
        if(L >= 1) {
            u32 ga;  // base + 4*id*S = 4*aid - 3*base
            asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(ga) : "r"(aid), "r"(0u - 3u * base));
            g0 = ld(ga);
            g1 = ld(ga + S);
            g2 = ld(ga + 2u * S);
            g3 = ld(ga + 3u * S);
        }
        u32 ca;  // address of the left child: base + 2*id*S = 2*aid - base
        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(ca) : "r"(aid), "r"(0u - base));
        const u32 prod = (v + (1u << L)) * t;  // left subtree: counts + the implicit one per symbol
        u32 anext, vnext, nl, nr;
        asm("{ .reg .pred p;\n\tsetp.le.u32 p, %6, %0;\n\t@p sub.u32 %0, %0, %6;\n\t@!p add.u32 %1, %1, 1;\n\t"
            "selp.u32 %2, %7, %8, p;\n\tselp.u32 %3, %9, %10, p;\n\t"
            "selp.u32 %4, %11, %12, p;\n\tselp.u32 %5, %13, %14, p; }"
            : "+r"(rem), "+r"(v), "=r"(anext), "=r"(vnext), "=r"(nl), "=r"(nr)
            : "r"(prod), "r"(ca + S), "r"(ca), "r"(cr), "r"(cl), "r"(g2), "r"(g0), "r"(g3), "r"(g1));
        st(aid, v);  // incremented when the symbol went left, unchanged otherwise
        aid = anext;
        v = vnext;
        cl = nl;
        cr = nr;
    }
    // AdaptiveFrequencyTable::find (cpprcoder.h:1221-1241) in the product domain + update.
    // Returns the symbol; `rem` comes back as low - cum*t, freq as the symbol's frequency.
    static __device__ __forceinline__ u32 decode(u32 base, u32 t, u32& rem, u32& freq)
    {
        u32 aid = base + S;  // node 1, the root
        u32 v = ld(aid), cl = ld(base + 2u * S), cr = ld(base + 3u * S);
        dec_level<7>(base, t, rem, aid, v, cl, cr);
        dec_level<6>(base, t, rem, aid, v, cl, cr);
        dec_level<5>(base, t, rem, aid, v, cl, cr);
        dec_level<4>(base, t, rem, aid, v, cl, cr);
        dec_level<3>(base, t, rem, aid, v, cl, cr);
        dec_level<2>(base, t, rem, aid, v, cl, cr);
        dec_level<1>(base, t, rem, aid, v, cl, cr);
        dec_level<0>(base, t, rem, aid, v, cl, cr);
        st(aid, v + 1u);  // aid is now the leaf 256 | symbol, v its count
        freq = v + 1u;
        return ((aid - base) / S) & 255u;
    }
};

constexpr u32 ADAPT_REC_WORDS = 3u + 128u;  // B2RC_ADAPTIVE_RESTART_WORDS: a restart point of the adaptive coder
constexpr u32 ADAPT_REC_WORDS_WIDE = 3u + 256u;  // ... for blocks above 65536 bytes: the counts take 32 bits
template <class W>
constexpr u32 adapt_rec_words() { return sizeof(W) == 2 ? ADAPT_REC_WORDS : ADAPT_REC_WORDS_WIDE; }

// Called by k_enc_adaptive at the tile that starts segment j >= 1 of the lane's block: the point's three
// words, then the model's 256 leaf counts (entries 256..511 of the lane's tree), two per word (u16) or one.
template <class W>
__device__ __forceinline__ void adaptive_mark(u32* rec, const RcEnc& st, s32 wcount, const LaneTab<W>& tab)
{
    const u32 words = (u32)(wcount + 1) + st.nff;       // words cut off the shift register so far
    rec[0] = 4u * words + (u32)st.ocnt / 8u - 1u;       // bytes shifted out of low, the dummy byte aside
    rec[1] = st.low;
    rec[2] = st.range;
    if(sizeof(W) == 2) {
#pragma unroll 4
        for(u32 s = 0; s < 256u; s += 2u) {
            rec[3u + s / 2u] = tab.ld(256u + s) | (tab.ld(257u + s) << 16);
        }
    } else {
#pragma unroll 4
        for(u32 s = 0; s < 256u; ++s) {
            rec[3u + s] = tab.ld(256u + s);
        }
    }
}

template <class W, bool RAGGED>
__device__ __forceinline__ void enc_adaptive_tiles(const EncArgs& a, u32 tiles, LaneTab<W>& tab, RcEnc& st,
                                                   SlotSink& sink, u64 b0, u32 n_b, u32 n_max, u32 lane)
{
    const u32 ntiles = (n_max + TILE - 1) / TILE;
#pragma unroll 1
    for(u32 tix = 0; tix < ntiles; ++tix) {
        if(tix + 1 < ntiles) {
            stage_tile(tiles + ((tix + 1) & 1u) * TILE_BYTES, a.src, a.n, b0, a.block, (tix + 1) * TILE, lane);
        }
        cp_async_commit();
        const u32 d0 = 256u + tix * TILE;
        const u32 mg0 = rc_magic(d0 + lane);
        const u32 mg1 = rc_magic(d0 + 32u + lane);
        cp_async_wait<1>();
        __syncwarp();
        if(a.restart && tix != 0u && (tix * TILE) % a.seg_syms == 0u && tix * TILE < n_b) {
            // a segment starts here: what a decoder needs to start here too, the model included (b2rc_adaptseg.cuh)
            const u32 nrec = (a.block + a.seg_syms - 1u) / a.seg_syms - 1u;
            adaptive_mark(a.restart + ((b0 + lane) * nrec + tix * TILE / a.seg_syms - 1u) * (u64)adapt_rec_words<W>(), st,
                          sink.wcount, tab);
        }
        const u32 row = tiles + (tix & 1u) * TILE_BYTES + lane * ROW;
#pragma unroll 1
        for(int wi = 0; wi < TILE / 4; ++wi) {
            const u32 word = lds32(row + 4u * (u32)wi);
            const u32 mg = wi < 8 ? mg0 : mg1;
            RcCut cuts[4];
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                const int j = wi * 4 + k;
                const u32 sym = (word >> (8 * k)) & 0xFFu;
                const u32 magic = __shfl_sync(FULL, mg, j & 31);
                const bool active = !RAGGED || tix * TILE + j < n_b;
                u32 cum = 0, freq = 1;
                if(active) {
                    TreeOps<W>::encode(tab.base, sym, cum, freq);
                }
                const u32 t = rc_div(st.range, d0 + j, magic);
                // the funnel-shift form of the step: this kernel runs one warp per scheduler and is bound by
                // the latency of its chain; the multiplier form (fewer instructions, longer latencies) was
                // measured here and lost 7 % (profiles/r2_ncu_notes.md)
                rc_enc_step<3>(st, cum, freq, t, cuts[k], active);
            }
            rc_enc_commit(st, cuts, sink);
        }
        __syncwarp();
    }
}

template <class W>
__global__ void __launch_bounds__(32) k_enc_adaptive(EncArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    constexpr u32 TAB_BYTES = 512u * 32u * sizeof(W);
    const u32 sbase = smem_addr(smem);
    const u32 tiles = sbase + TAB_BYTES;

    const u32 lane = lane_id();
    const u64 b0 = (u64)blockIdx.x * 32u;
    const u64 b = b0 + lane;
    const bool has = b < a.nblocks;
    u32 n_b = 0;
    if(has) {
        const u64 lo = b * (u64)a.block;
        n_b = (u32)((a.n - lo < a.block) ? (a.n - lo) : a.block);
    }
    u8* slot = a.slots + b * a.slot_stride;

    stage_tile(tiles, a.src, a.n, b0, a.block, 0, lane);
    cp_async_commit();
    {  // initialize(): all counts zero (the ones are implicit), cpprcoder.h:1094-1132
        uint4* z = reinterpret_cast<uint4*>(smem);
        for(u32 i = lane; i < TAB_BYTES / 16u; i += 32u) {
            z[i] = make_uint4(0, 0, 0, 0);
        }
    }
    if(has) {
        *reinterpret_cast<u32*>(slot) = n_b;  // cpprcoder.h:689-694
    }
    __syncwarp();

    RcEnc st;
    rc_enc_init(st, RC_ADAPT_RANGE0);
    SlotSink sink;
    sink.out = reinterpret_cast<u32*>(slot + RC_ADAPT_HDR);
    sink.wcount = -1;
    sink.cap_words = has ? (u32)((a.slot_stride - RC_ADAPT_HDR) / 4u) : 0u;
    sink.err = a.err;
    LaneTab<W> tab{sbase + lane * (u32)sizeof(W)};

    const u32 n_max = __reduce_max_sync(FULL, n_b);
    if(__any_sync(FULL, n_b != n_max)) {
        enc_adaptive_tiles<W, true>(a, tiles, tab, st, sink, b0, n_b, n_max, lane);
    } else {
        enc_adaptive_tiles<W, false>(a, tiles, tab, st, sink, b0, n_b, n_max, lane);
    }
    finish_block(st, sink, has, RC_ADAPT_HDR, n_b, slot, a.sizes + b);
}

// ============================================================= decoder input ==
// Each lane reads its own payload at its own pace (WordSrc below); it pops one aligned
// word whenever its bit window runs low.
struct DecArgs {
    const u8* payload;    // payload base
    u64 payload_len;
    const u64* offsets;   // nblocks + 1, relative to payload
    u64 nblocks;
    u32 block;
    u8* dst;
    u64 n;
    int* err;
    // Phased decode (host pipeline): symbols [sym0, sym0 + nsym) of every block in this launch,
    // coder state carried between launches in `state` (8 words per block).  nsym == 0: the
    // whole block in one launch, `state` unused.
    u32 sym0, nsym;
    u32* state;
    u8* model;  // adaptive coder only: the count tables of 32 blocks (one CTA) per slot, parked alongside
    // Restart points (k_dec_static_seg): the records the encoder left, see EncArgs
    const u32* restart;
    u32 seg_syms;
};

// Each lane reads its own payload at its own pace, one aligned word at a time.  Words
// travel global -> shared by 4-byte cp.async into an 8-deep queue per lane ([slot][lane],
// bank == lane) and shared -> register one word ahead of use.  No register ever waits on
// a global load: in lock step the scoreboard is per register, not per lane, so a plain
// prefetching LDG made every step wait for the previous step's load (profiles/).
constexpr int INQ = 8;
constexpr int INQ_BYTES = INQ * 128;
struct WordSrc {
    const u32* base;  // aligned word holding coded byte 0 of this lane's payload
    u32 lim;          // readable bytes from `base` to the end of the stream buffer
    u32 q;            // shared address of this lane's queue column
    u32 rd;           // index of the word held in `ahead`
    u32 ahead;        // that word (raw little endian), already in a register

    // word i -> queue slot i % INQ, when `on`; bytes past the buffer arrive as zeros
    __device__ __forceinline__ void request(u32 i, bool on) const
    {
        const u32 last = lim ? (lim - 1u) >> 2 : 0u;
        const u32 at = i < last ? i : last;                                 // keep the address inside the buffer
        const u32 left = (i << 2) < lim ? lim - (i << 2) : 0u;
        const u32 bytes = left < 4u ? left : 4u;                            // src-size: the rest is zero filled
        const u32 slot = q + (i & (INQ - 1)) * 128u;
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 4, %2; }" ::"r"(slot),
                     "l"(base + at), "r"(bytes), "r"((u32)on)
                     : "memory");
        cp_async_commit();
    }
    // the same without bounds: only for words known to lie wholly inside the buffer
    __device__ __forceinline__ void request_inside(u32 i, bool on) const
    {
        const u32 slot = q + (i & (INQ - 1)) * 128u;
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 4; }" ::"r"(slot),
                     "l"(base + i), "r"((u32)on)
                     : "memory");
        cp_async_commit();
    }
    // true when every word a tile of TILE symbols can request (3 bytes per symbol at most, plus
    // the queue's look-ahead) lies inside the buffer
    __device__ __forceinline__ bool tile_is_inside() const { return (u64)(rd + TILE + 2u * INQ) * 4u <= (u64)lim; }
    __device__ __forceinline__ void prime(u32 start = 0)
    {
#pragma unroll
        for(u32 i = 0; i < (u32)INQ; ++i) {
            request(start + i, true);
        }
        cp_async_wait<0>();
        rd = start;
        ahead = lds32v(q + (rd & (INQ - 1)) * 128u);
    }
    // Next stream word, big endian, when `need`; otherwise nothing moves.  Every call commits
    // one (possibly empty) copy group, so "all but the newest INQ-1 groups" always covers the
    // word that is read ahead here: it was requested INQ-1 or more calls ago.
    template <bool INSIDE>
    __device__ __forceinline__ u32 take_(bool need)
    {
        const u32 w = rc_bswap(ahead);
        if(INSIDE) {
            request_inside(rd + INQ, need);  // reuses the slot of the word just handed out
        } else {
            request(rd + INQ, need);
        }
        rd += need ? 1u : 0u;
        cp_async_wait<INQ - 1>();
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0;\n\t@p ld.shared.u32 %0, [%1]; }"
                     : "+r"(ahead)
                     : "r"(q + (rd & (INQ - 1)) * 128u), "r"((u32)need));
        return w;
    }
    __device__ __forceinline__ u32 take(bool need) { return take_<false>(need); }
    __device__ __forceinline__ u32 operator()() { return take(true); }
};
// View of a WordSrc for a stretch of symbols whose words are all inside the buffer.
struct WordSrcInside {
    WordSrc& s;
    __device__ __forceinline__ u32 take(bool need) { return s.take_<true>(need); }
};

// Output tile -> global, 8 rows of 64 bytes per instruction.
__device__ __forceinline__ void store_tile(const u8* tile, u8* dst, u64 n, u64 b0, u32 block, u32 tile_off, u32 lane)
{
    constexpr int CH = TILE / 16;
    constexpr int RPI = 32 / CH;
#pragma unroll
    for(int it = 0; it < 32 / RPI; ++it) {
        const u32 row = it * RPI + lane / CH;
        const u32 ch = lane % CH;
        const u64 blk_lo = (b0 + row) * (u64)block;
        u64 blk_hi = blk_lo + block;
        if(blk_hi > n) {
            blk_hi = n;
        }
        const u64 at = blk_lo + tile_off + ch * 16u;
        if(at >= blk_hi) {
            continue;
        }
        const uint4 v = *reinterpret_cast<const uint4*>(tile + row * ROW + ch * 16u);
        if(at + 16 <= blk_hi) {
            *reinterpret_cast<uint4*>(dst + at) = v;
        } else {
            const u32 w4[4] = {v.x, v.y, v.z, v.w};
            for(u32 k = 0; at + k < blk_hi; ++k) {
                dst[at + k] = (u8)(w4[k >> 2] >> (8 * (k & 3)));
            }
        }
    }
}

__device__ __forceinline__ void dec_setup(const DecArgs& a, u32 hdr, u64 b, bool has, u32 n_b, u32 queue, WordSrc& src,
                                          const u8*& pay, bool& ok)
{
    pay = a.payload;
    u64 len = 0;
    if(has) {
        const u64 o0 = a.offsets[b], o1 = a.offsets[b + 1];
        if(o0 <= o1 && o1 <= a.payload_len) {
            pay = a.payload + o0;
            len = o1 - o0;
        }
    }
    ok = has && len >= (u64)hdr + 5u;
    if(ok) {
        const u32 want = (u32)pay[0] | ((u32)pay[1] << 8) | ((u32)pay[2] << 16) | ((u32)pay[3] << 24);
        ok = want == n_b;  // the container, not the payload, says how long block b is
    }
    if(has && !ok) {
        atomicOr(a.err, ERR_CORRUPT);
    }
    const u8* coded = pay + hdr;
    const u8* wbase = (const u8*)((uintptr_t)coded & ~(uintptr_t)3);
    src.base = reinterpret_cast<const u32*>(wbase);
    // reading past this block's own payload is harmless (those bits are never decisive for a
    // valid stream); reading past the buffer is not: `lim` bounds every copy
    const u64 room = (u64)((a.payload + a.payload_len) - wbase);
    src.lim = ok ? (u32)(room < 0xFFFFFFF0ull ? room : 0xFFFFFFF0ull) : 0u;  // !ok: everything reads as zero
    src.q = queue;
    src.prime();
}

// ======================================================================= K3s ==
// Static decode.  Symbol search is two 16-way steps in the product domain
// (cum * t <= low, no second divide): level 1 compares the 15
// chunk boundaries cum[16j] held in registers, level 2 the 15 boundaries inside the
// chunk from the lane's table.  Equivalent to RangeEncoder::find (cpprcoder.h:521-535).
// The table is the plain 257-entry u32 cum array: trailing symbols that never occur
// have cum == total, which can be 65536 and does not fit a packed 16-bit field.
constexpr u32 DEC_STATIC_TAB = 257u * 32u * 4u;  // 32896, a multiple of 16

struct CumTab {
    enum : u32 { UNIT = 128 };  // position = symbol * 128 = byte offset inside the lane's column
    u32 base;                   // shared address of cum[0] for this lane
    __device__ __forceinline__ u32 at(u32 pos) const { return lds32(base + pos); }
};

// The same table in 16 bits (k_dec_static_seg: half the shared memory, seven CTAs per SM instead of
// four).  Only one value does not fit: cum == total == 65536, which the entries BEHIND the last
// symbol that occurs hold; they are stored as 65535.  The search is exact whenever it lands in
// front of that symbol; where it lands on or behind it, the symbol IS the last one that occurs
// (symbols behind it have no interval), so its cum and freq come from registers (Last below).
struct CumTab16 {
    enum : u32 { UNIT = 64 };
    u32 base;
    __device__ __forceinline__ u32 at(u32 pos) const { return lds16(base + pos); }
};
struct Last {
    u32 sym, cum, freq;  // the last symbol with a non-zero count, its cum and freq
};

// One tile (TILE symbols per lane).  MODE: 0 general divide, 2 / 3 power-of-two total with at
// most 2 / 3 renormalisation rounds per symbol.
template <int MODE, bool RAGGED, bool PAIR = false, class Tab = CumTab, class Src>
__device__ __forceinline__ void dec_static_tile(const Tab& tab, const u32 (&k1)[8], RcDec& d, u32& t, Src& src,
                                                u32 otile_a, u32 tile_off, u32 n_b, u32 total, u32 magic, u32 shift,
                                                u32 lane, const Last* last = nullptr)
{
#pragma unroll 1
    for(int wi = 0; wi < TILE / 4; ++wi) {
        u32 word = 0;
#pragma unroll
        for(int k = 0; k < 4; ++k) {
            if(!RAGGED || tile_off + wi * 4 + k < n_b) {
                if(MODE == 0) {
                    t = rc_div(d.range, total, magic);
                }
                u32 sym, cum, freq;
                if(PAIR) {  // the segmented kernel: fewer instructions beat a shorter chain there
                    const u32 k0[4] = {k1[0], k1[2], k1[4], k1[6]};
                    rc_static_find4(tab, k0, t, d.low, sym, cum, freq);
                    if(Tab::UNIT == CumTab16::UNIT) {  // see CumTab16
                        const bool at_last = sym >= last->sym;
                        sym = at_last ? last->sym : sym;
                        cum = at_last ? last->cum : cum;
                        freq = at_last ? last->freq : freq;
                    }
                } else {
                    rc_static_find(tab, k1, t, d.low, sym, cum, freq);
                }
                if(MODE == 0) {
                    rc_dec_advance(d, cum, freq, t, src);
                } else if(MODE == 2 && PAIR) {
                    rc_dec_advance_pow2_pair(d, t, shift, cum, freq, src, (k & 1) != 0);  // the window lasts for two
                } else {
                    rc_dec_advance_pow2<MODE>(d, t, shift, cum, freq, src);
                }
                word |= sym << (8 * k);
            }
        }
        sts32v(otile_a + lane * ROW + wi * 4, word);
    }
}

template <int MODE, bool RAGGED, bool PAIR = false, class Tab = CumTab>
__device__ __forceinline__ void dec_static_tiles(const DecArgs& a, const Tab& tab, const u32 (&k1)[8], RcDec& d,
                                                 WordSrc& src, u8* otile, u32 otile_a, u64 b0, u32 n_b, u32 tix0,
                                                 u32 tix1, bool resume, u32 total, u32 magic, u32 shift, u32 lane,
                                                 const Last* last = nullptr)
{
    // a resumed power-of-two chain finds t where the previous launch left it (in d.range).
    // In the one-launch kernel tix0, tix1 and resume are compile-time constants / ntiles.
    u32 t = MODE ? (resume ? d.range : (d.range >> shift)) : 0u;
#pragma unroll 1
    for(u32 tix = tix0; tix < tix1; ++tix) {
        // all but the last tile or two of the last block of the stream read words that lie wholly
        // inside the buffer: those tiles skip the bounds arithmetic of the copy requests
        if(__all_sync(FULL, src.tile_is_inside())) {
            WordSrcInside in{src};
            dec_static_tile<MODE, RAGGED, PAIR>(tab, k1, d, t, in, otile_a, tix * TILE, n_b, total, magic, shift, lane, last);
        } else {
            dec_static_tile<MODE, RAGGED, PAIR>(tab, k1, d, t, src, otile_a, tix * TILE, n_b, total, magic, shift, lane, last);
        }
        __syncwarp();
        store_tile(otile, a.dst, a.n, b0, a.block, tix * TILE, lane);
        __syncwarp();
    }
    if(MODE) {
        d.range = t;  // only its being non-zero is looked at afterwards
    }
}

// PHASED = false: the whole block in one launch (the device API, what bench.py times).
// PHASED = true: symbols [a.sym0, a.sym0 + a.nsym) with the coder state parked in a.state --
// its own instantiation, so that the one-launch kernel's hot loop is compiled exactly as it
// was before phases existed (it is sensitive to the slightest change, profiles/r1_ncu_notes.md).
template <bool PHASED>
__global__ void __launch_bounds__(32) k_dec_static(DecArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    const u32 sbase = smem_addr(smem);
    u32* table = reinterpret_cast<u32*>(smem);
    u8* otile = smem + DEC_STATIC_TAB;
    const u32 otile_a = sbase + DEC_STATIC_TAB;
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
    WordSrc src;
    const u8* pay;
    bool ok;
    dec_setup(a, RC_STATIC_HDR, b, has, n_b, queue_a + lane * 4u, src, pay, ok);

    // read16 + calcCumulatives (cpprcoder.h:585-602, :573-583); payloads are unaligned
    u32* mine = table + lane;
    u32 total = 0;
    u32 k1[8];
    {
        u32 run = 0;
#pragma unroll 1
        for(u32 s = 0; s < 256; ++s) {
            u32 f = 0;
            if(ok) {
                f = (u32)pay[4u + 2u * s] | ((u32)pay[5u + 2u * s] << 8);
            }
            mine[s * 32u] = run;
            run += f;
        }
        mine[256u * 32u] = run;
        total = run;
        if(ok && total == 0) {
            ok = false;
            atomicOr(a.err, ERR_CORRUPT);
        }
#pragma unroll
        for(int j = 0; j < 8; ++j) {
            k1[j] = mine[(32u * j) * 32u];
        }
    }
    if(!ok) {
        n_b = 0;
        total = 1;
    }
    __syncwarp();
    const CumTab tab{sbase + lane * 4u};
    const u32 magic = rc_magic(total);
    const bool is_pow2 = (total & (total - 1u)) == 0;
    const u32 shift = is_pow2 ? 31u - rc_clz(total) : 0u;
    RcDec d;
    const bool resume = PHASED && a.sym0 != 0u;
    u32* saved = PHASED ? a.state + (has ? b : b0) * 8u : nullptr;
    if(!resume) {
        rc_dec_init(d, RC_STATIC_RANGE0, (u32)((uintptr_t)(pay + RC_STATIC_HDR) & 3u), src);
    } else {
        d.low = saved[0];
        d.range = saved[1];
        d.w_hi = saved[2];
        d.w_lo = saved[3];
        d.wbits = (s32)saved[4];
        src.prime(ok ? saved[5] : 0u);
    }

    const u32 n_max = __reduce_max_sync(FULL, n_b);
    const bool all_pow2 = __all_sync(FULL, is_pow2);
    const bool ragged = __any_sync(FULL, n_b != n_max);
    const u32 ntiles = (n_max + TILE - 1) / TILE;
    const u32 tix0 = PHASED ? a.sym0 / TILE : 0u;
    u32 tix1 = PHASED ? (a.sym0 + a.nsym) / TILE : ntiles;
    tix1 = tix1 < ntiles ? tix1 : ntiles;
    if(all_pow2 && !ragged && a.block <= 65536u) {
        dec_static_tiles<2, false>(a, tab, k1, d, src, otile, otile_a, b0, n_b, tix0, tix1, resume, total, magic, shift,
                                   lane);
    } else if(all_pow2 && !ragged) {
        dec_static_tiles<3, false>(a, tab, k1, d, src, otile, otile_a, b0, n_b, tix0, tix1, resume, total, magic, shift,
                                   lane);
    } else if(!ragged) {
        dec_static_tiles<0, false>(a, tab, k1, d, src, otile, otile_a, b0, n_b, tix0, tix1, resume, total, magic, shift,
                                   lane);
    } else {
        dec_static_tiles<0, true>(a, tab, k1, d, src, otile, otile_a, b0, n_b, tix0, tix1, resume, total, magic, shift,
                                  lane);
    }
    if(PHASED && has) {  // park the chain for the next launch (also when this block is already done)
        saved[0] = d.low;
        saved[1] = d.range;
        saved[2] = d.w_hi;
        saved[3] = d.w_lo;
        saved[4] = (u32)d.wbits;
        saved[5] = src.rd;
    }
    if(tix1 >= ntiles && ok && d.range == 0) {
        atomicOr(a.err, ERR_CORRUPT);
    }
}

// ---------------------------------------------------------------- K3s, segmented --
// Static decode from restart points: the container carries, for every block and every
// seg_syms-th symbol, what the decoder's state is there -- {bytes shifted so far, the encoder's
// low, range}; the decoder's own low is the stream's next four bytes minus the encoder's low
// (DESIGN.md section 10).  A block is then seg-many independent chains instead of one.
// Mapping: a CTA of SEG_WARPS warps takes 32 blocks; lane = block as everywhere, WARP = segment,
// so the cumulative tables of the 32 blocks are built once and shared by the CTA with the usual
// bank == lane layout.  blockIdx.y walks groups of SEG_WARPS segments for long blocks.
constexpr u32 SEG_WARPS = 4;       // the default CTA; the launch may take any count up to SEG_WARPS_MAX (blockDim.x / 32)
constexpr u32 SEG_WARPS_MAX = 12;
// NARROW: blocks <= 65536 bytes, u16 table (CumTab16); else the u32 table of k_dec_static.
constexpr u32 dec_seg_tab(bool narrow)
{
    return 257u * 32u * (narrow ? 2u : 4u);
}
constexpr u32 dec_seg_smem(bool narrow, u32 warps = SEG_WARPS)
{
    return dec_seg_tab(narrow) + 3u * 128u + warps * (TILE_BYTES + INQ_BYTES);
}

template <bool NARROW>
__global__ void __launch_bounds__(32 * SEG_WARPS_MAX, 3) k_dec_static_seg(DecArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    constexpr u32 DEC_SEG_TAB = dec_seg_tab(NARROW);
    typedef typename std::conditional<NARROW, u16, u32>::type Entry;
    typedef typename std::conditional<NARROW, CumTab16, CumTab>::type Tab;
    constexpr u32 CLAMP = NARROW ? 65535u : 0xFFFFFFFFu;
    const u32 sbase = smem_addr(smem);
    Entry* table = reinterpret_cast<Entry*>(smem);
    u32* tots = reinterpret_cast<u32*>(smem + DEC_SEG_TAB);  // per block: total, last symbol that occurs, its cum
    const u32 warp = threadIdx.x >> 5, lane = lane_id();
    u8* otile = smem + DEC_SEG_TAB + 384u + warp * (TILE_BYTES + INQ_BYTES);
    const u32 otile_a = sbase + DEC_SEG_TAB + 384u + warp * (TILE_BYTES + INQ_BYTES);
    const u32 queue_a = otile_a + TILE_BYTES;

    const u64 b0 = (u64)blockIdx.x * 32u;
    const u64 b = b0 + lane;
    const bool has = b < a.nblocks;
    u32 n_b = 0;
    if(has) {
        const u64 lo = b * (u64)a.block;
        n_b = (u32)((a.n - lo < a.block) ? (a.n - lo) : a.block);
    }
    const u32 nseg = (a.block + a.seg_syms - 1u) / a.seg_syms;
    const u32 seg = blockIdx.y * (blockDim.x >> 5) + warp;

    // where this lane's payload is; every warp needs that, warp 0 also reports a bad one
    const u8* pay = a.payload;
    u64 len = 0;
    if(has) {
        const u64 o0 = a.offsets[b], o1 = a.offsets[b + 1];
        if(o0 <= o1 && o1 <= a.payload_len) {
            pay = a.payload + o0;
            len = o1 - o0;
        }
    }
    bool ok = has && len >= (u64)RC_STATIC_HDR + 5u;
    if(ok) {
        const u32 want = (u32)pay[0] | ((u32)pay[1] << 8) | ((u32)pay[2] << 16) | ((u32)pay[3] << 24);
        ok = want == n_b;
    }
    Entry* mine = table + lane;
    if(warp == 0) {
        // read16 + calcCumulatives (cpprcoder.h:585-602, :573-583), once for the CTA
        u32 run = 0, lsym = 0, lcum = 0;
#pragma unroll 1
        for(u32 s = 0; s < 256; ++s) {
            u32 f = 0;
            if(ok) {
                f = (u32)pay[4u + 2u * s] | ((u32)pay[5u + 2u * s] << 8);
            }
            mine[s * 32u] = (Entry)(run < CLAMP ? run : CLAMP);
            if(f) {
                lsym = s;
                lcum = run;
            }
            run += f;
        }
        mine[256u * 32u] = (Entry)(run < CLAMP ? run : CLAMP);
        tots[lane] = run;
        tots[32u + lane] = lsym;
        tots[64u + lane] = lcum;
        // a table that sums to more than 2^16 is not one an encoder of <= 64 KiB blocks writes, and
        // the 16-bit entries could not hold it
        if(has && (!ok || run == 0 || (NARROW && run > 65536u)) && blockIdx.y == 0) {
            atomicOr(a.err, ERR_CORRUPT);
        }
    }
    __syncthreads();
    if(seg >= nseg) {
        return;
    }
    u32 total = tots[lane];
    if(total == 0 || (NARROW && total > 65536u)) {
        ok = false;
        total = 1;
    }
    const Last last{tots[32u + lane], tots[64u + lane], total - tots[64u + lane]};
    const Tab tab{sbase + lane * (u32)sizeof(Entry)};
    u32 k1[8];
#pragma unroll
    for(int j = 0; j < 8; ++j) {
        k1[j] = tab.at((32u * j) * Tab::UNIT);
        k1[j] = (32u * (u32)j > last.sym) ? total : k1[j];  // exact in registers: 65536 fits here
    }
    const u32 seg_lo = seg * a.seg_syms;
    u32 seg_hi = seg_lo + a.seg_syms;
    seg_hi = seg_hi < n_b ? seg_hi : n_b;  // my symbols: [seg_lo, seg_hi)
    bool mine_ok = ok && seg_lo < n_b;
    const u32 skip0 = (u32)((uintptr_t)(pay + RC_STATIC_HDR) & 3u);
    u32 skip = skip0, word0 = 0, range0 = RC_STATIC_RANGE0, enc_low = 0;
    const u32 nrec = nseg - 1u;
    if(seg != 0u && mine_ok) {
        const u32* rec = a.restart + (b * nrec + seg - 1u) * 3u;
        const u32 m = rec[0];
        enc_low = rec[1];
        range0 = rec[2];
        if(m == 0xFFFFFFFFu || (u64)m + RC_STATIC_HDR + 5u > len) {
            mine_ok = false;
            atomicOr(a.err, ERR_CORRUPT);
        } else {
            // the byte at offset m of the coded stream plays the part of the dummy first byte
            word0 = (skip + m) >> 2;
            skip = (skip + m) & 3u;
        }
    }
    WordSrc src;
    {
        const u8* coded = pay + RC_STATIC_HDR;
        const u8* wbase = (const u8*)((uintptr_t)coded & ~(uintptr_t)3);
        src.base = reinterpret_cast<const u32*>(wbase);
        const u64 room = (u64)((a.payload + a.payload_len) - wbase);
        src.lim = mine_ok ? (u32)(room < 0xFFFFFFF0ull ? room : 0xFFFFFFF0ull) : 0u;
        src.q = queue_a + lane * 4u;
        src.prime(mine_ok ? word0 : 0u);
    }
    const u32 n_eff = mine_ok ? seg_hi : 0u;  // symbols at or beyond this are not mine
    const u32 magic = rc_magic(total);
    const bool is_pow2 = (total & (total - 1u)) == 0;
    const u32 shift = is_pow2 ? 31u - rc_clz(total) : 0u;
    RcDec d;
    rc_dec_init(d, range0, skip, src);
    d.low -= enc_low;

    const u32 n_max = __reduce_max_sync(FULL, n_eff);
    const bool all_pow2 = __all_sync(FULL, is_pow2);
    const bool ragged = __any_sync(FULL, n_eff != n_max);
    const u32 tix0 = seg_lo / TILE;
    const u32 ntiles = (n_max + TILE - 1) / TILE;
    const u32 tix1 = ntiles > tix0 ? ntiles : tix0;
    const bool pow2_path = all_pow2 && !ragged;
    if(pow2_path && NARROW) {
        dec_static_tiles<2, false, true>(a, tab, k1, d, src, otile, otile_a, b0, n_eff, tix0, tix1, false, total, magic, shift,
                                         lane, &last);
    } else if(pow2_path) {
        dec_static_tiles<3, false, true>(a, tab, k1, d, src, otile, otile_a, b0, n_eff, tix0, tix1, false, total, magic, shift,
                                         lane, &last);
    } else if(!ragged) {
        dec_static_tiles<0, false, true>(a, tab, k1, d, src, otile, otile_a, b0, n_eff, tix0, tix1, false, total, magic, shift,
                                         lane, &last);
    } else {
        dec_static_tiles<0, true, true>(a, tab, k1, d, src, otile, otile_a, b0, n_eff, tix0, tix1, false, total, magic, shift,
                                        lane, &last);
    }

    // ---- the records are part of the container and as untrusted as the rest of it: a chain that was
    //      started from a damaged record must not pass for a decode.  A segment has to END where the
    //      next record says the coder stands -- the decoder's low is the stream's next four bytes minus
    //      the encoder's low there, and range / total is the same on both sides; the last segment of a
    //      block has to end with exactly the block's coded bytes used up (5 + shifts, cpprcoder.h:494-517).
    if(mine_ok) {
        const u32 used = 4u * src.rd - skip0 - (u32)d.wbits / 8u;  // bytes of the coded stream taken into low so far
        bool good;
        if(seg_hi < n_b) {
            const u32* rec = a.restart + (b * nrec + seg) * 3u;
            const u32 m = rec[0];
            good = m != 0xFFFFFFFFu && (u64)m + RC_STATIC_HDR + 5u <= len && used == m + 5u;
            if(good) {
                const u8* p = pay + RC_STATIC_HDR + 1u + m;
                const u32 be = ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3];
                const u32 t_mine = pow2_path ? d.range : rc_div(d.range, total, magic);
                const u32 t_rec = is_pow2 ? (rec[2] >> shift) : rc_div(rec[2], total, magic);
                good = d.low == be - rec[1] && t_mine == t_rec;
            }
        } else {
            good = (u64)used + RC_STATIC_HDR == len;
        }
        if(!good) {
            atomicOr(a.err, ERR_CORRUPT);
        }
    }
}

// ======================================================================= K3a ==
template <class W, class Src>
__device__ __forceinline__ void dec_adaptive_tile(LaneTab<W>& tab, RcDec& d, Src& src, u32 otile_a, u32 tile_off,
                                                  u32 n_b, u32 lane)
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
                const u32 sym = TreeOps<W>::decode(tab.base, t, d.low, freq);  // d.low -= cum * t on the way
                rc_dec_advance(d, 0u, freq, t, src);
                word |= sym << (8 * k);
            }
        }
        sts32v(otile_a + lane * ROW + wi * 4, word);
    }
}

// PHASED: as for k_dec_static -- symbols [a.sym0, a.sym0 + a.nsym) per launch, the coder state in
// a.state and the warp's count tables (the model after sym0 symbols) in a.model between launches.
template <class W, bool PHASED>
__global__ void __launch_bounds__(32) k_dec_adaptive(DecArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    constexpr u32 TAB_BYTES = 512u * 32u * sizeof(W);
    const u32 sbase = smem_addr(smem);
    u8* otile = smem + TAB_BYTES;
    const u32 otile_a = sbase + TAB_BYTES;

    const u32 lane = lane_id();
    const u64 b0 = (u64)blockIdx.x * 32u;
    const u64 b = b0 + lane;
    const bool has = b < a.nblocks;
    u32 n_b = 0;
    if(has) {
        const u64 lo = b * (u64)a.block;
        n_b = (u32)((a.n - lo < a.block) ? (a.n - lo) : a.block);
    }
    WordSrc src;
    const u8* pay;
    bool ok;
    dec_setup(a, RC_ADAPT_HDR, b, has, n_b, otile_a + TILE_BYTES + lane * 4u, src, pay, ok);
    const bool resume = PHASED && a.sym0 != 0u;
    uint4* parked = PHASED ? reinterpret_cast<uint4*>(a.model + (u64)blockIdx.x * TAB_BYTES) : nullptr;
    {
        uint4* z = reinterpret_cast<uint4*>(smem);
        for(u32 i = lane; i < TAB_BYTES / 16u; i += 32u) {
            z[i] = resume ? parked[i] : make_uint4(0, 0, 0, 0);
        }
    }
    if(!ok) {
        n_b = 0;
    }
    __syncwarp();
    LaneTab<W> tab{sbase + lane * (u32)sizeof(W)};
    RcDec d;
    u32* saved = PHASED ? a.state + (has ? b : b0) * 8u : nullptr;
    if(!resume) {
        rc_dec_init(d, RC_ADAPT_RANGE0, (u32)((uintptr_t)(pay + RC_ADAPT_HDR) & 3u), src);
    } else {
        d.low = saved[0];
        d.range = saved[1];
        d.w_hi = saved[2];
        d.w_lo = saved[3];
        d.wbits = (s32)saved[4];
        src.prime(ok ? saved[5] : 0u);
    }

    const u32 n_max = __reduce_max_sync(FULL, n_b);
    const u32 ntiles = (n_max + TILE - 1) / TILE;
    const u32 tix0 = PHASED ? a.sym0 / TILE : 0u;
    u32 tix1 = PHASED ? (a.sym0 + a.nsym) / TILE : ntiles;
    tix1 = tix1 < ntiles ? tix1 : ntiles;
#pragma unroll 1
    for(u32 tix = tix0; tix < tix1; ++tix) {
        if(__all_sync(FULL, src.tile_is_inside())) {
            WordSrcInside in{src};
            dec_adaptive_tile<W>(tab, d, in, otile_a, tix * TILE, n_b, lane);
        } else {
            dec_adaptive_tile<W>(tab, d, src, otile_a, tix * TILE, n_b, lane);
        }
        __syncwarp();
        store_tile(otile, a.dst, a.n, b0, a.block, tix * TILE, lane);
        __syncwarp();
    }
    if(PHASED) {
        const uint4* z = reinterpret_cast<const uint4*>(smem);
        for(u32 i = lane; i < TAB_BYTES / 16u; i += 32u) {
            parked[i] = z[i];
        }
        if(has) {
            saved[0] = d.low;
            saved[1] = d.range;
            saved[2] = d.w_hi;
            saved[3] = d.w_lo;
            saved[4] = (u32)d.wbits;
            saved[5] = src.rd;
        }
    }
    if(tix1 >= ntiles && ok && d.range == 0) {
        atomicOr(a.err, ERR_CORRUPT);
    }
}

// ======================================================================== K1 ==
// Per-block histogram for blocks <= 65536 bytes: one warp per block, 16-byte loads,
// warp-private shared-memory bins.  For such blocks RangeEncoder::count never halves
// except when all 65536 bytes are equal, where it ends at 0x8000.
constexpr int HIST_WARPS = 8;

// Exact byte counts of one block into 256 warp-private shared-memory bins (one warp).
__device__ __forceinline__ void hist_block(u32* h, const u8* p, u32 len, u32 lane)
{
#pragma unroll
    for(int k = 0; k < 8; ++k) {
        h[lane + 32 * k] = 0;
    }
    __syncwarp();
    const u32 vec = len & ~15u;
    // four 16-byte loads in flight per lane before the first bin is touched: the kernel was
    // stalled on load latency (long scoreboard), not on the shared-memory atomics
    u32 off = lane * 16u;
    for(; off + 3u * 512u < vec; off += 4u * 512u) {
        uint4 v[4];
#pragma unroll
        for(int q = 0; q < 4; ++q) {
            v[q] = __ldg(reinterpret_cast<const uint4*>(p + off + q * 512u));
        }
#pragma unroll
        for(int q = 0; q < 4; ++q) {
            const u32 w4[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
            for(int k = 0; k < 16; ++k) {
                atomicAdd(&h[(w4[k >> 2] >> (8 * (k & 3))) & 0xFFu], 1u);
            }
        }
    }
    for(; off < vec; off += 512u) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + off));
        const u32 w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for(int k = 0; k < 16; ++k) {
            atomicAdd(&h[(w4[k >> 2] >> (8 * (k & 3))) & 0xFFu], 1u);
        }
    }
    for(u32 o2 = vec + lane; o2 < len; o2 += 32u) {
        atomicAdd(&h[p[o2]], 1u);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(HIST_WARPS * 32) k_hist(const u8* src, u64 n, u32 block, u64 nblocks, u16* freq16)
{
    __shared__ u32 bins[HIST_WARPS][256];
    const u32 lane = lane_id();
    const u32 warp = threadIdx.x >> 5;
    u32* h = bins[warp];
    for(u64 b = (u64)blockIdx.x * HIST_WARPS + warp; b < nblocks; b += (u64)gridDim.x * HIST_WARPS) {
        const u64 lo = b * (u64)block;
        const u32 len = (u32)((n - lo < block) ? (n - lo) : block);
        hist_block(h, src + lo, len, lane);
        u32 f[8];
#pragma unroll
        for(int k = 0; k < 8; ++k) {
            f[k] = h[8u * lane + k];
            if(f[k] >= 0x10000u) {
                f[k] = 0x8000u;  // 65535 hits, then halve-all-and-increment: (0xFFFF >> 1 | 1) + 1
            }
        }
        uint4 o;
        o.x = f[0] | (f[1] << 16);
        o.y = f[2] | (f[3] << 16);
        o.z = f[4] | (f[5] << 16);
        o.w = f[6] | (f[7] << 16);
        reinterpret_cast<uint4*>(freq16 + b * 256u)[lane] = o;
        __syncwarp();
    }
}

// (Tried and taken out: LANE-private bins, u32 [bin][lane], 32 KiB per warp, seven warps per SM, the 32 partial
// counts summed once per block -- no two lanes ever in one bank.  0.595 ms against 0.338 ms for 1 GiB.  The
// shared-memory atomic here is ATOMS.POPC.INC: lanes with the SAME address are merged by the hardware and cost
// nothing; what costs is the number of DISTINCT addresses per instruction, and private bins make that 32
// every time.  With shared bins a zipf stream has ~20 distinct bytes per instruction, which fall into the 32
// banks like balls into bins: 2.45 wavefronts per instruction, the shared-memory pipe 96 % busy -- that, not
// HBM, is this kernel's roofline: 0.49 of the measured copy bandwidth.)

// K1 for blocks above 65536 bytes, where RangeEncoder::count is order dependent: whenever the
// symbol about to be counted already stands at 0xFFFF, EVERY non-zero count becomes
// (x >> 1) | 1 first (cpprcoder.h:549-555).  One warp per block walks it in segments: the
// segment's histogram is taken in parallel, and if no count can reach the limit inside the
// segment (count + hits <= 0xFFFF for every symbol) it is simply added -- exact, because
// then no halving happens in it.  Otherwise one lane replays the segment byte by byte with
// the reference's rule.  A halving leaves the largest count at 0x8000, so at most one segment
// in eight takes the slow road even on a block of equal bytes.  Counts end <= 0xFFFF.
constexpr u32 HIST_SEG = 4096u;
__global__ void __launch_bounds__(HIST_WARPS * 32) k_hist_wide(const u8* src, u64 n, u32 block, u64 nblocks, u16* freq16)
{
    __shared__ u32 bins[HIST_WARPS][256];
    __shared__ u32 tots[HIST_WARPS][256];
    const u32 lane = lane_id();
    const u32 warp = threadIdx.x >> 5;
    u32* h = bins[warp];
    u32* tot = tots[warp];
    for(u64 b = (u64)blockIdx.x * HIST_WARPS + warp; b < nblocks; b += (u64)gridDim.x * HIST_WARPS) {
        const u64 lo = b * (u64)block;
        const u32 len = (u32)((n - lo < block) ? (n - lo) : block);
        u32 f[8];
#pragma unroll
        for(int k = 0; k < 8; ++k) {
            f[k] = 0;
        }
        for(u32 at = 0; at < len; at += HIST_SEG) {
            const u32 seg = len - at < HIST_SEG ? len - at : HIST_SEG;
            hist_block(h, src + lo + at, seg, lane);
            bool fits = true;
            u32 add[8];
#pragma unroll
            for(int k = 0; k < 8; ++k) {
                add[k] = h[8u * lane + k];
                fits = fits && (f[k] + add[k] <= 0xFFFFu);
            }
            if(__all_sync(FULL, fits)) {
#pragma unroll
                for(int k = 0; k < 8; ++k) {
                    f[k] += add[k];
                }
            } else {
#pragma unroll
                for(int k = 0; k < 8; ++k) {
                    tot[8u * lane + k] = f[k];
                }
                __syncwarp();
                if(lane == 0) {
                    const u8* p = src + lo + at;
                    for(u32 i = 0; i < seg; ++i) {
                        const u32 c = p[i];
                        u32 x = tot[c];
                        if(x >= 0xFFFFu) {
                            for(u32 s2 = 0; s2 < 256u; ++s2) {
                                const u32 y = tot[s2];
                                if(y) {
                                    tot[s2] = (y >> 1) | 1u;
                                }
                            }
                            x = tot[c];
                        }
                        tot[c] = x + 1u;
                    }
                }
                __syncwarp();
#pragma unroll
                for(int k = 0; k < 8; ++k) {
                    f[k] = tot[8u * lane + k];
                }
            }
            __syncwarp();
        }
        uint4 o;
        o.x = f[0] | (f[1] << 16);
        o.y = f[2] | (f[3] << 16);
        o.z = f[4] | (f[5] << 16);
        o.w = f[6] | (f[7] << 16);
        reinterpret_cast<uint4*>(freq16 + b * 256u)[lane] = o;
    }
}

// ======================================================================== K4 ==
// Exclusive scan of the payload sizes (one CTA; the index is tiny next to the data)
// and, fused into it, the 32-byte container header when `header` is not null.
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) k_scan(const u32* sizes, u64 nblocks, u64* offsets, u64* total_out,
                                                       u8* header, u32 mode, u32 block, u64 n, const u64* base_in,
                                                       u32 flags)
{
    // base_in: where this range of blocks starts in the payload area (chunked host pipeline:
    // chunk c continues where chunk c-1 ended; total_out then receives the new end)
    const u64 base = base_in ? *base_in : 0ull;
    __shared__ u64 part[SCAN_THREADS];
    const u32 t = threadIdx.x;
    const u64 per = (nblocks + SCAN_THREADS - 1) / SCAN_THREADS;
    u64 lo = per * t, hi = lo + per;
    if(lo > nblocks) {
        lo = nblocks;
    }
    if(hi > nblocks) {
        hi = nblocks;
    }
    u64 sum = 0;
    for(u64 i = lo; i < hi; ++i) {
        sum += sizes[i];
    }
    part[t] = sum;
    __syncthreads();
    for(u32 d = 1; d < SCAN_THREADS; d <<= 1) {
        u64 add = 0;
        if(t >= d) {
            add = part[t - d];
        }
        __syncthreads();
        part[t] += add;
        __syncthreads();
    }
    u64 run = base + part[t] - sum;
    for(u64 i = lo; i < hi; ++i) {
        offsets[i] = run;
        run += sizes[i];
    }
    if(t == SCAN_THREADS - 1) {
        offsets[nblocks] = base + part[t];
        if(total_out) {
            *total_out = base + part[t];
        }
    }
    if(header && t == 0) {
        u32* h = reinterpret_cast<u32*>(header);
        h[0] = 0x43523242u;  // 'B','2','R','C'
        h[1] = 1u | (mode << 16);
        h[2] = block;
        h[3] = flags;
        h[4] = (u32)n;
        h[5] = (u32)(n >> 32);
        h[6] = (u32)nblocks;
        h[7] = (u32)(nblocks >> 32);
    }
}

// The restart table goes behind the payloads, at the next 4-byte boundary; where that is, is only
// known on the device when this runs (*total).
__global__ void __launch_bounds__(256) k_put_table(const u32* table, u64 words, u8* payload, const u64* total, u64 cap,
                                                   int* err)
{
    const u64 at = (*total + 3ull) & ~3ull;
    if(at + 4ull * words > cap) {
        if(blockIdx.x == 0 && threadIdx.x == 0) {
            atomicOr(err, ERR_DST_SMALL);
        }
        return;
    }
    if(blockIdx.x == 0 && threadIdx.x < (u32)(at - *total)) {
        payload[*total + threadIdx.x] = 0;  // the padding in front of the table
    }
    u32* d = reinterpret_cast<u32*>(payload + at);
    for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (u64)gridDim.x * blockDim.x) {
        d[i] = table[i];
    }
}

// Compaction: payload b (16-byte aligned slot) -> payload + offsets[b] (any alignment).
// One CTA per block at a time; destination-aligned 4-byte words assembled from two
// neighbouring source words with a funnel shift.
constexpr int COMPACT_THREADS = 256;
__global__ void __launch_bounds__(COMPACT_THREADS) k_compact(const u8* slots, u64 slot_stride, const u32* sizes,
                                                             const u64* offsets, u64 nblocks, u8* payload,
                                                             u64 payload_cap, int* err)
{
    for(u64 b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const u8* s = slots + b * slot_stride;
        const u32 len = sizes[b];
        const u64 off = offsets[b];
        if(off + len > payload_cap) {
            if(threadIdx.x == 0) {
                atomicOr(err, ERR_DST_SMALL);
            }
            continue;
        }
        u8* d = payload + off;
        u32 head = (u32)((4u - ((uintptr_t)d & 3u)) & 3u);
        if(head > len) {
            head = len;
        }
        if(threadIdx.x < head) {
            d[threadIdx.x] = s[threadIdx.x];
        }
        const u32 nwords = (len - head) >> 2;
        u32* dw = reinterpret_cast<u32*>(d + head);
        const u32* sw = reinterpret_cast<const u32*>(s);
        const u32 sh = head * 8u;  // source byte offset of dst word j is head + 4j
        u32 j = threadIdx.x;
        for(; j + 3u * COMPACT_THREADS < nwords; j += 4u * COMPACT_THREADS) {  // four loads in flight per thread
            u32 a0[4], a1[4];
#pragma unroll
            for(int q = 0; q < 4; ++q) {
                a0[q] = __ldg(sw + j + q * COMPACT_THREADS);
                a1[q] = sh ? __ldg(sw + j + q * COMPACT_THREADS + 1) : 0u;
            }
#pragma unroll
            for(int q = 0; q < 4; ++q) {
                dw[j + q * COMPACT_THREADS] = __funnelshift_r(a0[q], a1[q], sh);
            }
        }
        for(; j < nwords; j += COMPACT_THREADS) {
            const u32 a0 = __ldg(sw + j);
            const u32 a1 = sh ? __ldg(sw + j + 1) : 0u;
            dw[j] = __funnelshift_r(a0, a1, sh);
        }
        const u32 done = head + 4u * nwords;
        if(done + threadIdx.x < len) {
            d[done + threadIdx.x] = s[done + threadIdx.x];
        }
    }
}

}  // namespace b2rc
