// b2rc_encseg.cuh -- static encode as many chains per block, written straight into the container.
//
//   K2r k_enc_ranges   the range-only chain of every block (cpprcoder.h:401-404 without `low`):
//                      before every P-th symbol, the bytes shifted out so far and the range; the
//                      exact payload size of every block BEFORE a byte is coded
//   K2s k_enc_seg      every segment of P symbols coded from low = 0 into its own bytes of the final
//                      payload (RangeEncoder::encode, cpprcoder.h:400-436); also the payload header
//   K2m k_enc_seams    the four bytes of low each segment is left with, added where the next segment
//                      begins, carries and all; the restart points; the low_ == 0xFFFFFFFF flush quirk
//
// Why: k_enc_static runs ONE chain of 65 536 links per 64 KiB block, so its time is links x cycles
// per link whatever the number of blocks, a lone warp per scheduler issues every second cycle, and
// a GPU with an eighth of the blocks (8-GPU run of a 1 GiB stream) takes as long as one with all of
// them.  `range` never depends on `low`, so the only truly serial part is the range chain -- a
// multiply, two compares and a shift per symbol (K2r).  With every segment's starting range and
// byte position known, K2s has blocks x segments independent chains (131 072 lanes x 4 per 64 KiB
// block at P = 2048): enough warps per scheduler to be bound by issue slots, at any block size and
// any number of GPUs.  Sizes are exact up front, so the payloads are written once, in place: no
// staging slots, no compaction pass.  rc_lane.cuh ("segmented static encode") holds the arithmetic
// and tests/sim runs exactly that code on the CPU against the oracle.
#pragma once
#include "b2rc_kernels.cuh"

namespace b2rc
{
constexpr int ERR_INTERNAL = 8;  // a segment did not end where the range pass said it would

struct SegArgs {
    const u8* src;
    u64 n;
    u32 block;
    u64 nblocks;
    const u16* freq16;   // K1's counts, [nblocks][256]
    u32 P;               // symbols per segment, a multiple of TILE
    u32 nseg;            // ceil(block / P)
    u32* recs;           // [nblocks][nseg + 1][2]: bytes shifted out before symbol j*P, range there
                         // (entry j = number of segments of the block: the totals)
    u32* lows;           // [nblocks][nseg]: the low every segment ended with
    u32* sizes;          // payload bytes per block (K2r writes, the scan reads)
    const u64* offsets;  // where payload b starts, relative to `payload`
    u8* payload;
    u64 payload_cap;
    u32* restart;        // restart table (EncArgs::restart layout) or null
    u32 seg_syms;
    int* err;
    u32 force_exact;     // tests: every block takes the reference-shaped path of the flush quirk
};

__device__ __forceinline__ u32 seg_block_len(const SegArgs& a, u64 b)
{
    const u64 lo = b * (u64)a.block;
    return (u32)((a.n - lo < a.block) ? (a.n - lo) : a.block);
}

// ======================================================================= K2r ==
// One block per lane (as everywhere).  Shared memory: the 32 blocks' frequencies, u16 [256][32]
// (a count never exceeds 0xFFFF, cpprcoder.h:549-555), and two input tiles.
constexpr u32 ENC_RANGES_SMEM = 256u * 32u * 2u + 2u * TILE_BYTES;

template <int MAXSH, bool POW2, bool RAGGED>
__device__ __forceinline__ void enc_range_tiles(const SegArgs& a, u32 tiles, u32 ftab, u64 b0, u32 n_b, u32 n_max,
                                                u32 total, u32 magic, u32 shift, u32 lane, u32* rec)
{
    u32 range = RC_STATIC_RANGE0;
    u32 t = POW2 ? (range >> shift) : 0u;
    u32 bits = 0;
    u32 ntotal = 0u - total;
    asm volatile("" : "+r"(ntotal));  // opaque: see rc_range_step_div
    const u32 ntiles = (n_max + TILE - 1) / TILE;
    const u32 seg_tiles = a.P / TILE;
    u32 next_mark = 0;
    // (a deeper ring of input tiles was tried for 1 MiB blocks, whose rows lie 1 MiB apart: no change -- the
    // chain, not the staging, is what the lone warp waits for)
#pragma unroll 1
    for(u32 tix = 0; tix < ntiles; ++tix) {
        if(tix + 1 < ntiles) {
            stage_tile(tiles + ((tix + 1) & 1u) * TILE_BYTES, a.src, a.n, b0, a.block, (tix + 1) * TILE, lane);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        if(tix == next_mark) {  // a segment starts here
            next_mark += seg_tiles;
            if(tix * TILE < n_b) {
                const u32 j = tix / seg_tiles;
                // the general chain carries the range before its renormalisation (rc_range_step_div)
                const u32 shn = POW2 ? 0u : rc_norm_shift_flo(range);
                rec[2u * j] = (bits + shn) >> 3;
                // any range with the same range / total serves a decoder; ONE form is written whatever path the
                // warp took (the path depends on which blocks share a warp, which a multi-device split changes):
                // power-of-two total (shift != 0, or total 1): the quotient shifted back; else the range itself
                rec[2u * j + 1u] = POW2 ? (t << shift) : (((range << shn) >> shift) << shift);
            }
        }
        const u32 row = tiles + (tix & 1u) * TILE_BYTES + lane * ROW;
        // frequencies of the NEXT four symbols are requested before the current four are chained
        u32 word = lds32(row);
        u32 f[4];
#pragma unroll
        for(int k = 0; k < 4; ++k) {
            f[k] = lds16(ftab + ((word >> (8 * k)) & 0xFFu) * 64u);
        }
#pragma unroll 1
        for(int wi = 0; wi < TILE / 4; ++wi) {
            const u32 wnext = lds32(row + 4u * (u32)((wi + 1) & (TILE / 4 - 1)));
            u32 nf[4];
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                nf[k] = lds16(ftab + ((wnext >> (8 * k)) & 0xFFu) * 64u);
            }
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                if(!RAGGED || tix * TILE + wi * 4 + k < n_b) {
                    if(POW2) {
                        bits += rc_range_step_pow2<MAXSH>(t, shift, f[k]);
                    } else {
                        rc_range_step_div(range, bits, f[k], total, ntotal, magic);
                    }
                }
            }
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                f[k] = nf[k];
            }
        }
        __syncwarp();
    }
    if(n_b) {
        const u32 nseg_b = (n_b + a.P - 1u) / a.P;
        rec[2u * nseg_b] = (bits + (POW2 ? 0u : rc_norm_shift_flo(range))) >> 3;
        rec[2u * nseg_b + 1u] = 0u;
    }
}

// The chain for total == 65536 (every full 64 KiB block that is not one repeated byte).
// t = range >> 16 lies in [2^8, 2^16).  With r = freq * t the next t is r >> 16, r >> 8 or r, whichever
// lies in [2^8, 2^16) (cpprcoder.h:418: shift left by 8 while range < 2^24).  Subtracting the lower
// end of each candidate's home range BEFORE shifting
//     A = (r - 2^24) >> 16     B = (r - 2^16) >> 8     C = r - 2^8
// makes the candidate that applies come out as t' - 256 (below 65280) and the others wrap around to
// 65280 or more, so t' - 256 is the minimum of the three (VIMNMX3) -- no compares, no selects.  The
// three r - K are three multiply-adds with an immediate addend, side by side.  A lone warp per
// scheduler pays for every instruction it issues (about 2.3 cycles each, profiles/r2_ncu_notes.md),
// so the count matters more than the depth: 4 multiply-adds, 2 shifts, the minimum, one add, and the
// shift count from the leading zero bytes of r (conversion pipe) -- 14 instructions per symbol with
// the two loads and their address.
__device__ __forceinline__ void range_step16(u32& t, u32& bits, u32 f)
{
    u32 r0, r1, r2, r3, top;
    asm("mad.lo.u32 %0, %1, %2, 0xFF000000;" : "=r"(r1) : "r"(f), "r"(t));
    asm("mad.lo.u32 %0, %1, %2, 0xFFFF0000;" : "=r"(r2) : "r"(f), "r"(t));
    asm("mad.lo.u32 %0, %1, %2, 0xFFFFFF00;" : "=r"(r3) : "r"(f), "r"(t));
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(r0) : "r"(f), "r"(t));
    t = min(min(r1 >> 16, r2 >> 8), r3) + 256u;
    // shift = 8 * leading zero bytes of r (2^8 <= r): 24 & (31 - top), top the highest set bit
    asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(r0));
    bits += ~top & 24u;
}

template <bool RAGGED>
__device__ __forceinline__ void enc_range_tiles16(const SegArgs& a, u32 tiles, u32 ftab, u64 b0, u32 n_b, u32 n_max,
                                                  u32 lane, u32* rec)
{
    u32 u = RC_STATIC_RANGE0 >> 16;  // t
    u32 bits = 0;
    const u32 ntiles = (n_max + TILE - 1) / TILE;
    const u32 seg_tiles = a.P / TILE;
    u32 next_mark = 0;
#pragma unroll 1
    for(u32 tix = 0; tix < ntiles; ++tix) {
        if(tix + 1 < ntiles) {
            stage_tile(tiles + ((tix + 1) & 1u) * TILE_BYTES, a.src, a.n, b0, a.block, (tix + 1) * TILE, lane);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        if(tix == next_mark) {  // a segment starts here
            next_mark += seg_tiles;
            if(tix * TILE < n_b) {
                const u32 j = tix / seg_tiles;
                rec[2u * j] = bits >> 3;
                rec[2u * j + 1u] = u << 16;  // any range with the same range >> 16 serves
            }
        }
        const u32 row = tiles + (tix & 1u) * TILE_BYTES + lane * ROW;
        // symbols by byte loads (the load pipe is idle, the integer pipe is not); frequencies of the
        // NEXT four symbols are requested before the current four are chained
        u32 f[4];
#pragma unroll
        for(int k = 0; k < 4; ++k) {
            f[k] = lds16(ftab + lds8(row + k) * 64u);
        }
#pragma unroll 2
        for(u32 at = 0; at < (u32)TILE; at += 4u) {
            const u32 nx = row + ((at + 4u) & (TILE - 1u));
            u32 nf[4];
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                nf[k] = lds16(ftab + lds8(nx + k) * 64u);
            }
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                if(!RAGGED || tix * TILE + at + k < n_b) {
                    range_step16(u, bits, f[k]);
                }
            }
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                f[k] = nf[k];
            }
        }
        __syncwarp();
    }
    if(n_b) {
        const u32 nseg_b = (n_b + a.P - 1u) / a.P;
        rec[2u * nseg_b] = bits >> 3;
        rec[2u * nseg_b + 1u] = 0u;
    }
}

// What a warp of the range pass knows about its 32 blocks once the frequencies are in shared memory.
struct RangesWarp {
    u32 lane, n_b, n_max, total, ftab;
    u64 b0, b;
    bool has, ragged;
    u32* rec;
};

// Stages input tile 0 (committed, not waited for) and builds the 32 blocks' frequency table, u16 [256][32].
__device__ __forceinline__ void ranges_prepare(const SegArgs& a, u8* smem, u32 sbase, u32 tiles, RangesWarp& w)
{
    w.lane = lane_id();
    w.b0 = (u64)blockIdx.x * 32u;
    w.b = w.b0 + w.lane;
    w.has = w.b < a.nblocks;
    w.n_b = w.has ? seg_block_len(a, w.b) : 0u;
    stage_tile(tiles, a.src, a.n, w.b0, a.block, 0, w.lane);
    cp_async_commit();
    u16* frq16 = reinterpret_cast<u16*>(smem);
    u32 total = 0;
#pragma unroll 1
    for(u32 r = 0; r < 32; ++r) {
        if(w.b0 + r >= a.nblocks) {
            break;
        }
        // lane j holds the frequencies of symbols 8j .. 8j+7 of block b0+r
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.freq16 + (w.b0 + r) * 256u) + w.lane);
        const u32 f[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16,
                          v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
        u32 sum = 0;
#pragma unroll
        for(int k = 0; k < 8; ++k) {
            sum += f[k];
            frq16[(8u * w.lane + k) * 32u + r] = (u16)f[k];
        }
        sum = __reduce_add_sync(FULL, sum);
        if(w.lane == r) {
            total = sum;
        }
    }
    __syncwarp();
    w.total = total;
    w.n_max = __reduce_max_sync(FULL, w.n_b);
    w.ragged = __any_sync(FULL, w.n_b != w.n_max);
    w.rec = a.recs + (w.has ? w.b : w.b0) * (u64)(a.nseg + 1u) * 2u;
    w.ftab = sbase + w.lane * 2u;
}

// The pass with ONE warp per 32 blocks (every path), sizes included.
template <bool WIDE>
__device__ __forceinline__ void ranges_one_warp(const SegArgs& a, u32 tiles, const RangesWarp& w)
{
    const u32 total = w.total, lane = w.lane, n_b = w.n_b, n_max = w.n_max, ftab = w.ftab;
    const u64 b0 = w.b0;
    u32* rec = w.rec;
    const u32 magic = rc_magic(total);
    const bool is_pow2 = total != 0 && (total & (total - 1u)) == 0;
    const u32 shift = is_pow2 ? 31u - rc_clz(total) : 0u;
    const bool all_pow2 = __all_sync(FULL, is_pow2 || !w.has);
    const bool ragged = w.ragged;
    constexpr int MAXSH = WIDE ? 3 : 2;
    const bool all16 = !WIDE && __all_sync(FULL, total == 65536u || !w.has);
    if(all16 && !ragged) {
        enc_range_tiles16<false>(a, tiles, ftab, b0, n_b, n_max, lane, rec);
    } else if(all16) {
        enc_range_tiles16<true>(a, tiles, ftab, b0, n_b, n_max, lane, rec);
    } else if(all_pow2 && !ragged) {
        enc_range_tiles<MAXSH, true, false>(a, tiles, ftab, b0, n_b, n_max, total, magic, shift, lane, rec);
    } else if(all_pow2) {
        enc_range_tiles<MAXSH, true, true>(a, tiles, ftab, b0, n_b, n_max, total, magic, shift, lane, rec);
    } else if(!ragged) {
        enc_range_tiles<3, false, false>(a, tiles, ftab, b0, n_b, n_max, total, magic, shift, lane, rec);
    } else {
        enc_range_tiles<3, false, true>(a, tiles, ftab, b0, n_b, n_max, total, magic, shift, lane, rec);
    }
    if(w.has) {
        const u32 nseg_b = (n_b + a.P - 1u) / a.P;
        a.sizes[w.b] = RC_STATIC_HDR + 5u + rec[2u * nseg_b];  // cpprcoder.h:386-395 header, buffer_, shifted bytes, low
    }
}

template <bool WIDE>
__global__ void __launch_bounds__(32) k_enc_ranges(SegArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    const u32 sbase = smem_addr(smem);
    const u32 tiles = sbase + 256u * 32u * 2u;
    RangesWarp w;
    ranges_prepare(a, smem, sbase, tiles, w);
    ranges_one_warp<WIDE>(a, tiles, w);
}

// --------------------------------------- K2r for 64 KiB blocks, two or three warps per 32 blocks --
// A lone warp issues an instruction every other cycle at best (2.0 - 2.3 cycles each in every capture of this
// round), so k_enc_ranges's 16 instructions per symbol cost 32 cycles although the chain for total == 65536 is
// only multiply-add, minimum, add.  Most of the 16 do not depend on the chain: the symbol's byte, its table
// address, its frequency, and the count of the bytes shifted out (multiply, highest set bit, mask, add).  Here
// other warps of the CTA do that work:
//   chain warp    per symbol: load {freq, 256 freq - 256}, two adds, three multiply-adds, two shifts, minimum,
//                 store r - 256 -- ten instructions, immediate addresses, a chain three deep (ranges2_chain)
//   freqs warp    turns input tile k + 1 into the tile of {freq, addend} pairs the chain warp reads next
//   shifts warp   counts what tile k - 1 shifted out from the products the chain warp left, writes the byte
//                 counts and the sizes, fetches the input tiles (cp.async) three tiles ahead
// (k_enc_ranges2<2>: one helper warp does both helper jobs).  One CTA barrier per tile of 64 symbols.  Same
// arithmetic, same records (tests/test_gpu_encseg.py runs every case with 1, 2 and 3 warps).
// MEASURED, 1 GiB Zipf / 128 MiB: one warp 1.08 / 1.06 ms; two warps 2.9 / 1.37 (the helper is the longer half);
// three warps 2.1 / 0.75 ms (22 cycles per link).  With more CTAs than SMs the extra warps share schedulers with
// other chains and everything slows down; with at most one CTA per SM (4736 blocks, 296 MiB) three warps win,
// and that is when they run (static_ranges_launch; env B2RC_RANGES_WARPS = 1, 2, 3 forces one form).
// A warp whose blocks do not all have total 65536 (a block of one repeated byte: 0x8000) falls back to the
// one-warp paths; the other warps leave.
constexpr u32 ENC_RANGES2_F = TILE * 256u;  // one tile of {freq, 256 * freq - 256}, 2 x u32 [symbol][lane]
constexpr u32 ENC_RANGES2_T = TILE * 128u;  // one tile of products, u32 [symbol][lane]
constexpr u32 ENC_RANGES2_SRC = 4u;         // input tiles in the second warp's ring: three tiles of lead on DRAM
constexpr u32 ENC_RANGES2_SMEM =
    256u * 32u * 2u + ENC_RANGES2_SRC * TILE_BYTES + 2u * ENC_RANGES2_F + 2u * ENC_RANGES2_T + 16u;
constexpr u32 RANGES2_IDLE = 0x7FFFFFFFu;   // "no symbol here": + 256 has its top bit set, shift 0

// The chain carries m = t - 256.  With r = freq * t = freq * m + 256 * freq the three candidates of range_step16
// are multiply-adds whose addends (256 * freq - K) do not depend on the chain, and their minimum IS the next m:
// the link is multiply-add, shift, minimum -- three deep, no add.
template <bool RAGGED>
__device__ __forceinline__ void ranges2_chain(const SegArgs& a, u32 fbuf, u32 tbuf, const RangesWarp& w)
{
    u32 m = (RC_STATIC_RANGE0 >> 16) - 256u;
    const u32 ntiles = (w.n_max + TILE - 1) / TILE;
    const u32 seg_tiles = a.P / TILE;
    u32 next_mark = 0;
    __syncthreads();  // the first tile of frequencies is there
#pragma unroll 1
    for(u32 tix = 0; tix < ntiles; ++tix) {
        if(tix == next_mark) {  // a segment starts here; another warp writes the byte count beside it
            next_mark += seg_tiles;
            if(tix * TILE < w.n_b) {
                w.rec[2u * (tix / seg_tiles) + 1u] = (m + 256u) << 16;  // any range with the same range >> 16 serves
            }
        }
        const u32 F = fbuf + (tix & 1u) * ENC_RANGES2_F + w.lane * 8u;
        const u32 T = tbuf + (tix & 1u) * ENC_RANGES2_T + w.lane * 4u;
        u32 f[8], c[8];
#pragma unroll
        for(int k = 0; k < 8; ++k) {
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(f[k]), "=r"(c[k]) : "r"(F + (u32)k * 256u));
        }
#pragma unroll
        for(u32 at = 0; at < (u32)TILE; at += 8u) {
            u32 nf[8], nc[8];
#pragma unroll
            for(int k = 0; k < 8; ++k) {
                nf[k] = nc[k] = 0u;
                if(at + 8u < (u32)TILE) {
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                                 : "=r"(nf[k]), "=r"(nc[k])
                                 : "r"(F + (at + 8u + (u32)k) * 256u));
                }
            }
#pragma unroll
            for(int k = 0; k < 8; ++k) {
                u32 r1, r2, r3 = RANGES2_IDLE;
                if(!RAGGED || tix * TILE + at + (u32)k < w.n_b) {
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r1) : "r"(f[k]), "r"(m), "r"(c[k] - 0x00FFFF00u));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r2) : "r"(f[k]), "r"(m), "r"(c[k] - 0x0000FF00u));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r3) : "r"(f[k]), "r"(m), "r"(c[k]));
                    m = min(min(r1 >> 16, r2 >> 8), r3);
                }
                sts32v(T + (at + (u32)k) * 128u, r3);
            }
#pragma unroll
            for(int k = 0; k < 8; ++k) {
                f[k] = nf[k];
                c[k] = nc[k];
            }
        }
        __syncthreads();
    }
}

// bytes * 8 the coder shifts out for one tile of products
__device__ __forceinline__ u32 ranges2_shifts(u32 T)
{
#if defined(B2RC_RANGES2_NO_SHIFTS)  // timing experiment only: wrong sizes
    return lds32v(T) & 8u;
#else
    u32 bits = 0;
#pragma unroll 1
    for(u32 j0 = 0; j0 < (u32)TILE; j0 += 16u) {
        u32 r[16];
#pragma unroll
        for(int k = 0; k < 16; ++k) {
            r[k] = lds32v(T + (j0 + (u32)k) * 128u);
        }
#pragma unroll
        for(int k = 0; k < 16; ++k) {
            u32 top;
            asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(r[k] + 256u));
            bits += ~top & 24u;
        }
    }
    return bits;
#endif
}

// one input tile -> one tile of frequencies.  Thirty-two at a time: the loads of a batch go out together (volatile
// accesses keep their order, and a store behind every load would wait for two shared-memory round trips)
__device__ __forceinline__ void ranges2_freqs(u32 tile_row, u32 ftab, u32 F)
{
#pragma unroll 1
    for(u32 j0 = 0; j0 < (u32)TILE; j0 += 32u) {
        u32 sym[32], f[32];
#pragma unroll
        for(int k = 0; k < 32; ++k) {
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(sym[k]) : "r"(tile_row + j0 + (u32)k));
        }
#pragma unroll
        for(int k = 0; k < 32; ++k) {
            f[k] = lds16(ftab + sym[k] * 64u);
        }
#pragma unroll
        for(int k = 0; k < 32; ++k) {
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(F + (j0 + (u32)k) * 256u), "r"(f[k]), "r"(f[k] * 256u - 256u));
        }
    }
}

// FREQS: this warp turns input tiles into frequency tiles; SHIFTS: it counts what the chain warp's tiles shift out
// and writes the byte counts and the sizes.  One helper warp does both (k_enc_ranges2<2>), or one each (<3>).
template <bool FREQS, bool SHIFTS>
__device__ __forceinline__ void ranges2_helper(const SegArgs& a, u32 tiles, u32 fbuf, u32 tbuf, const RangesWarp& w)
{
    // With two helper warps the one that counts the shifts has time to spare: it also fetches the input tiles
    // (cp.async) and waits for them, so that the warp that converts them never waits for DRAM: a tile is
    // confirmed one barrier before it is converted.
    constexpr bool STAGES = SHIFTS;
    constexpr bool LEAD = FREQS && SHIFTS ? 0 : 1;  // barriers between "tile has arrived" and its conversion
    const u32 lane = w.lane;
    const u32 ntiles = (w.n_max + TILE - 1) / TILE;
    const u32 seg_tiles = a.P / TILE;
    u32 next_mark = 0, bits = 0;
    constexpr u32 R = ENC_RANGES2_SRC;
    if(STAGES) {
        // tile 0 was staged by the chain warp and is complete (the barrier in the kernel); tiles 1 .. R - 1 go out now
        for(u32 k = 1; k < R; ++k) {
            if(k < ntiles) {
                stage_tile(tiles + k * TILE_BYTES, a.src, a.n, w.b0, a.block, k * TILE, lane);
            }
            cp_async_commit();
        }
        if(LEAD) {
            cp_async_wait<R - 2>();  // tile 1, for the other helper's first trip
            __syncwarp();
        }
    }
    if(FREQS) {
        ranges2_freqs(tiles + lane * ROW, w.ftab, fbuf + lane * 8u);
    }
    __syncthreads();
#pragma unroll 1
    for(u32 tix = 0; tix < ntiles; ++tix) {
        if(STAGES && !LEAD && tix + 1 < ntiles) {
            cp_async_wait<R - 2>();  // all but the newest R - 2 groups: tile tix + 1 has arrived
            __syncwarp();
        }
        if(FREQS && tix + 1 < ntiles) {  // frequencies of the tile the chain warp walks next
            ranges2_freqs(tiles + ((tix + 1) % R) * TILE_BYTES + lane * ROW, w.ftab,
                          fbuf + ((tix + 1) & 1u) * ENC_RANGES2_F + lane * 8u);
            __syncwarp();
        }
        if(STAGES) {
            // tile tix + R - LEAD goes into the buffer of tile tix - LEAD, which was converted an iteration ago
            if(tix + 1u > (u32)LEAD && tix + R - LEAD < ntiles) {
                stage_tile(tiles + ((tix - LEAD) % R) * TILE_BYTES, a.src, a.n, w.b0, a.block, (tix + R - LEAD) * TILE, lane);
            }
            cp_async_commit();
            if(LEAD && tix + 2 < ntiles) {
                cp_async_wait<R - 3>();  // tile tix + 2 has arrived: the other helper converts it after the barrier
                __syncwarp();
            }
        }
        if(SHIFTS) {
            if(tix >= 1) {  // what the tile the chain warp just left shifted out
                bits += ranges2_shifts(tbuf + ((tix - 1) & 1u) * ENC_RANGES2_T + lane * 4u);
            }
            if(tix == next_mark) {
                next_mark += seg_tiles;
                if(tix * TILE < w.n_b) {
                    w.rec[2u * (tix / seg_tiles)] = bits >> 3;
                }
            }
        }
        __syncthreads();
    }
    if(SHIFTS) {
        if(ntiles) {
            bits += ranges2_shifts(tbuf + ((ntiles - 1) & 1u) * ENC_RANGES2_T + lane * 4u);
        }
        if(w.n_b) {
            const u32 nseg_b = (w.n_b + a.P - 1u) / a.P;
            w.rec[2u * nseg_b] = bits >> 3;
            w.rec[2u * nseg_b + 1u] = 0u;
            a.sizes[w.b] = RC_STATIC_HDR + 5u + (bits >> 3);  // cpprcoder.h:386-395 header, buffer_, shifted bytes, low
        }
    }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_enc_ranges2(SegArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    const u32 sbase = smem_addr(smem);
    const u32 tiles = sbase + 256u * 32u * 2u;
    const u32 fbuf = tiles + ENC_RANGES2_SRC * TILE_BYTES;
    const u32 tbuf = fbuf + 2u * ENC_RANGES2_F;
    volatile u32* flag = reinterpret_cast<volatile u32*>(smem + 256u * 32u * 2u + ENC_RANGES2_SRC * TILE_BYTES +
                                                         2u * ENC_RANGES2_F + 2u * ENC_RANGES2_T);
    const u32 warp = threadIdx.x >> 5;
    RangesWarp w;
    if(warp == 0) {
        ranges_prepare(a, smem, sbase, tiles, w);
        const bool all16 = __all_sync(FULL, w.total == 65536u || !w.has);
        cp_async_wait<0>();  // tile 0: the other warp reads it first
        if(w.lane == 0) {
            flag[0] = all16 ? 1u : 0u;
            flag[1] = w.n_max;
            flag[2] = w.ragged ? 1u : 0u;
        }
    }
    __syncthreads();
    const bool two = flag[0] != 0u;
    if(!two) {
        if(warp == 0) {
            ranges_one_warp<false>(a, tiles, w);
        }
        return;
    }
    if(warp == 0) {
        if(w.ragged) {
            ranges2_chain<true>(a, fbuf, tbuf, w);
        } else {
            ranges2_chain<false>(a, fbuf, tbuf, w);
        }
    } else {
        w.lane = lane_id();
        w.b0 = (u64)blockIdx.x * 32u;
        w.b = w.b0 + w.lane;
        w.has = w.b < a.nblocks;
        w.n_b = w.has ? seg_block_len(a, w.b) : 0u;
        w.n_max = flag[1];
        w.ragged = flag[2] != 0u;
        w.total = 0;
        w.rec = a.recs + (w.has ? w.b : w.b0) * (u64)(a.nseg + 1u) * 2u;
        w.ftab = sbase + w.lane * 2u;
        if(WARPS == 2) {
            ranges2_helper<true, true>(a, tiles, fbuf, tbuf, w);
        } else if(warp == 1) {
            ranges2_helper<true, false>(a, tiles, fbuf, tbuf, w);
        } else {
            ranges2_helper<false, true>(a, tiles, fbuf, tbuf, w);
        }
    }
}

// ======================================================================= K2s ==
// A CTA takes 32 blocks (lane = block) and up to ENC_SEG_WARPS consecutive segments of them
// (warp = segment); blockIdx.y walks further groups of segments.  The 32 tables are built once
// per CTA with the bank == lane layout of k_enc_static.
constexpr u32 ENC_SEG_WARPS = 8;
constexpr u32 enc_seg_smem(bool wide, u32 warps)
{
    return (wide ? ENC_STATIC_TAB_WIDE : ENC_STATIC_TAB_NARROW) + 128u + warps * 2u * TILE_BYTES;  // tables, totals, tiles
}

// Which form of the encoder step the segment kernel runs (rc_lane.cuh).  Measured on 1 GiB zipf, P = 2048:
// funnel-shift form as k_enc_static has it 2.07 ms (integer pipe 84 % busy); multiplier form (RcEnc2) 2.03 ms
// (integer pipe 45 %, but IMAD.WIDE is dear); funnel-shift form with the byte loads, the one-compare test for
// all-ones words and the edge path that the multiplier version brought 1.88 ms: the default.
#if defined(B2RC_SEG_MULTIPLIER)
typedef RcEnc2 SegEnc;
#else
typedef RcEnc SegEnc;
#endif
#if defined(B2RC_SEG_NO_FLO)
constexpr bool SEG_FLO = false;
#else
constexpr bool SEG_FLO = true;  // the renormalisation shift from the highest set bit (rc_norm_shift_flo)
#endif
__device__ __forceinline__ u32 seg_low(const RcEnc& e) { return e.low; }
__device__ __forceinline__ u32 seg_low(const RcEnc2& e) { return (u32)e.x; }

template <bool WIDE, bool POW2, bool RAGGED>
__device__ __forceinline__ void enc_seg_tiles(const SegArgs& a, u32 tiles, const StaticTab<WIDE>& tab, SegEnc& st,
                                              RcSegSink& sink, u64 b0, u32 n_eff, u32 tix0, u32 tix1, u32 total,
                                              u32 magic, u32 shift, u32 lane)
{
    u32 tcur = POW2 ? (st.range >> shift) : 0u;  // the power-of-two chain carries t, not range
    // tile tix0 was staged (and committed) by the caller
#pragma unroll 1
    for(u32 tix = tix0; tix < tix1; ++tix) {
        if(tix + 1 < tix1) {
            stage_tile(tiles + ((tix + 1) & 1u) * TILE_BYTES, a.src, a.n, b0, a.block, (tix + 1) * TILE, lane);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const u32 row = tiles + (tix & 1u) * TILE_BYTES + lane * ROW;
        // symbols by byte loads: the load pipe has room, the integer pipe (which a word load's shifts and
        // masks would use) is what bounds this kernel.  Table entries of the NEXT four symbols are
        // requested before the current four are coded.
        u32 cum[4], freq[4];
#pragma unroll
        for(int k = 0; k < 4; ++k) {
            tab.get(lds8(row + k), cum[k], freq[k]);
        }
#pragma unroll 1
        for(u32 at = 0; at < (u32)TILE; at += 4u) {
            const u32 nx = row + ((at + 4u) & (TILE - 1u));
            u32 ncum[4], nfreq[4];
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                tab.get(lds8(nx + k), ncum[k], nfreq[k]);
            }
            RcCut cuts[4];
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                const bool active = !RAGGED || tix * TILE + at + k < n_eff;
#if !defined(B2RC_SEG_MULTIPLIER)
                if(POW2) {
                    rc_enc_step_pow2<WIDE ? 3 : 2, SEG_FLO>(st, tcur, shift, cum[k], freq[k], cuts[k], active);
                } else {
                    const u32 t = rc_div(st.range, total, magic);
                    rc_enc_step<WIDE ? 3 : 2, SEG_FLO>(st, cum[k], freq[k], t, cuts[k], active);
                }
            }
            rc_enc_commit_edge(st, cuts, sink);
#else
                if(POW2) {
                    rc_enc2_step_pow2<WIDE ? 3 : 2>(st, tcur, shift, cum[k], freq[k], cuts[k], active);
                } else {
                    const u32 t = rc_div(st.range, total, magic);
                    rc_enc2_step<WIDE ? 3 : 2>(st, cum[k], freq[k], t, cuts[k], active);
                }
            }
            rc_enc2_commit(st, cuts, sink);
#endif
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
__global__ void __launch_bounds__(32 * ENC_SEG_WARPS) k_enc_seg(SegArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    constexpr u32 TAB_BYTES = WIDE ? ENC_STATIC_TAB_WIDE : ENC_STATIC_TAB_NARROW;
    const u32 sbase = smem_addr(smem);
    const u32 warp = threadIdx.x >> 5, lane = lane_id(), nwarps = blockDim.x >> 5;
    u32* tots = reinterpret_cast<u32*>(smem + TAB_BYTES);  // total of each of the 32 blocks
    const u32 tiles = sbase + TAB_BYTES + 128u + warp * 2u * TILE_BYTES;
    const u64 b0 = (u64)blockIdx.x * 32u;
    const u64 b = b0 + lane;
    const bool has = b < a.nblocks;
    const u32 n_b = has ? seg_block_len(a, b) : 0u;
    const u32 seg = blockIdx.y * nwarps + warp;
    const u32 tix0 = seg * (a.P / TILE);

    // this warp's first input tile in flight while the tables are built
    if(seg < a.nseg) {
        stage_tile(tiles + (tix0 & 1u) * TILE_BYTES, a.src, a.n, b0, a.block, tix0 * TILE, lane);  // buffer = tile index & 1
    }
    cp_async_commit();

    // ---- tables of the 32 blocks, rows dealt round the warps; the first group of segments also
    //      writes the payload header: u32 LE size, write16 (cpprcoder.h:386-395, :604-619), and the
    //      coder's first byte, the initial buffer_ = 0 (cpprcoder.h:385), which stays 0
#pragma unroll 1
    for(u32 r = warp; r < 32; r += nwarps) {
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
        if(WIDE) {
            u32* cum32 = reinterpret_cast<u32*>(smem);
#pragma unroll
            for(int k = 0; k < 8; ++k) {
                cum32[(8u * lane + k) * 32u + r] = run;
                run += f[k];
            }
            if(lane == 31) {
                cum32[256u * 32u + r] = run;
                tots[r] = run;
            }
        } else {
            u16* cum16 = reinterpret_cast<u16*>(smem);
            u16* frq16 = cum16 + 256 * 32;
#pragma unroll
            for(int k = 0; k < 8; ++k) {
                // a symbol that occurs has cum <= total - freq <= 65535; for one that does not,
                // the truncated value is never read
                cum16[(8u * lane + k) * 32u + r] = (u16)run;
                frq16[(8u * lane + k) * 32u + r] = (u16)f[k];
                run += f[k];
            }
            if(lane == 31) {
                tots[r] = run;
            }
        }
        if(blockIdx.y == 0) {
            const u64 off = a.offsets[b0 + r];
            const u32 size_r = a.sizes[b0 + r];
            if(off + size_r <= a.payload_cap) {
                u8* pay = a.payload + off;
                const u32 n_r = seg_block_len(a, b0 + r);
                if(((uintptr_t)pay & 3u) == 0) {
                    u32* hw = reinterpret_cast<u32*>(pay + 4u + 16u * lane);
                    hw[0] = v.x;
                    hw[1] = v.y;
                    hw[2] = v.z;
                    hw[3] = v.w;
                    if(lane == 0) {
                        *reinterpret_cast<u32*>(pay) = n_r;
                    }
                } else {
                    const u32 w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for(int k = 0; k < 16; ++k) {
                        pay[4u + 16u * lane + k] = (u8)(w4[k >> 2] >> (8 * (k & 3)));
                    }
                    if(lane < 4) {
                        pay[lane] = (u8)(n_r >> (8u * lane));
                    }
                }
                if(lane == 0) {
                    pay[RC_STATIC_HDR] = 0;
                }
            } else if(lane == 0) {
                atomicOr(a.err, ERR_DST_SMALL);
            }
        }
    }
    __syncthreads();
    if(seg >= a.nseg) {
        return;
    }

    // ---- this lane's segment: where it starts, what it owns
    const u32 nseg_b = (n_b + a.P - 1u) / a.P;
    bool mine = has && seg < nseg_b;
    u64 off = 0;
    if(mine) {
        off = a.offsets[b];
        mine = off + a.sizes[b] <= a.payload_cap;  // the header writer reports it
    }
    u32 S0 = 0, S1 = 0, range0 = RC_STATIC_RANGE0;
    if(mine) {
        const u32* rec = a.recs + (b * (u64)(a.nseg + 1u) + seg) * 2u;
        S0 = rec[0];
        range0 = rec[1];
        S1 = rec[2];
    }
    const bool last = seg + 1u == nseg_b;
    const u32 own = (last ? S1 + 5u : S1 + 1u) - (S0 + 1u);
    SegEnc st;
    RcSegSink sink;
    rc_seg_begin(st, sink, a.payload + off + RC_STATIC_HDR + S0 + 1u, mine ? own : 0u, range0);
    const u32 total = has ? tots[lane] : 0u;
    const StaticTab<WIDE> tab{sbase + lane * (WIDE ? 4u : 2u)};
    const u32 magic = rc_magic(total);
    const bool is_pow2 = total != 0 && (total & (total - 1u)) == 0;
    const u32 shift = is_pow2 ? 31u - rc_clz(total) : 0u;
    u32 seg_hi = (seg + 1u) * a.P;
    seg_hi = seg_hi < n_b ? seg_hi : n_b;
    const u32 n_eff = mine ? seg_hi : 0u;  // symbols at or beyond this are not this lane's
    const u32 n_max = __reduce_max_sync(FULL, n_eff);
    const bool all_pow2 = __all_sync(FULL, is_pow2 || !mine);
    const bool ragged = __any_sync(FULL, n_eff != n_max);
    const u32 ntiles = (n_max + TILE - 1) / TILE;
    const u32 tix1 = ntiles > tix0 ? ntiles : tix0;

    if(all_pow2 && !ragged) {
        enc_seg_tiles<WIDE, true, false>(a, tiles, tab, st, sink, b0, n_eff, tix0, tix1, total, magic, shift, lane);
    } else if(all_pow2) {
        enc_seg_tiles<WIDE, true, true>(a, tiles, tab, st, sink, b0, n_eff, tix0, tix1, total, magic, shift, lane);
    } else if(!ragged) {
        enc_seg_tiles<WIDE, false, false>(a, tiles, tab, st, sink, b0, n_eff, tix0, tix1, total, magic, shift, lane);
    } else {
        enc_seg_tiles<WIDE, false, true>(a, tiles, tab, st, sink, b0, n_eff, tix0, tix1, total, magic, shift, lane);
    }
    if(mine) {
        if(!rc_seg_end(st, sink, last)) {
            atomicOr(a.err, ERR_INTERNAL);
        }
        a.lows[b * (u64)a.nseg + seg] = seg_low(st);
    }
}

// ======================================================================= K2m ==
// One thread per block, its seams in order (a carry may run back across earlier seams, and two
// seams less than four bytes apart overlap: one thread, one order).
constexpr int SEAM_THREADS = 64;

__global__ void __launch_bounds__(SEAM_THREADS) k_enc_seams(SegArgs a)
{
    const u64 b = (u64)blockIdx.x * SEAM_THREADS + threadIdx.x;
    if(b >= a.nblocks) {
        return;
    }
    const u32 n_b = seg_block_len(a, b);
    const u32 nseg_b = (n_b + a.P - 1u) / a.P;
    const u64 off = a.offsets[b];
    if(off + a.sizes[b] > a.payload_cap) {
        return;  // reported by k_enc_seg
    }
    u8* coded = a.payload + off + RC_STATIC_HDR;
    const u32* rec = a.recs + b * (u64)(a.nseg + 1u) * 2u;
    const u32* lows = a.lows + b * (u64)a.nseg;
    const u32 nrec = a.restart ? (a.block + a.seg_syms - 1u) / a.seg_syms - 1u : 0u;
    u32* rrow = a.restart ? a.restart + b * (u64)nrec * 3u : nullptr;
    u32 low = 0;
#pragma unroll 1
    for(u32 j = 0; j < nseg_b; ++j) {
        const u32 S0 = rec[2u * j], S1 = rec[2u * j + 2u];
        const u32 own = lows[j];
        low = rc_seam_low(low, own, S1 - S0);
        if(j + 1u < nseg_b) {
            rc_seam_add(coded + S1 + 1u, own);
            const u32 at = (j + 1u) * a.P;
            if(rrow && at % a.seg_syms == 0u) {
                u32* r = rrow + (at / a.seg_syms - 1u) * 3u;
                r[0] = S1;
                r[1] = low;
                r[2] = rec[2u * j + 3u];
            }
        }
    }
    for(u32 k = 0; k < nrec; ++k) {  // points the block ends before
        if((u64)(k + 1u) * a.seg_syms >= n_b) {
            rrow[3u * k] = 0xFFFFFFFFu;
            rrow[3u * k + 1u] = 0xFFFFFFFFu;
            rrow[3u * k + 2u] = 0xFFFFFFFFu;
        }
    }
    // cpprcoder.h:439-451: when the block ends on low_ == 0xFFFFFFFF the reference bumps the held
    // byte but still writes low_ as FF FF FF FF -- not the big-endian sum.  Probability 2^-32 per
    // block; such a block is coded again by the reference-shaped encoder (same size).
    if(low == 0xFFFFFFFFu || a.force_exact) {
        u32 cum[257];
        const u16* fq = a.freq16 + b * 256u;
        u32 run = 0;
        for(u32 s = 0; s < 256u; ++s) {
            cum[s] = run;
            run += fq[s];
        }
        cum[256] = run;
        const u8* blk = a.src + b * (u64)a.block;
        u32 at = 0;
        const u32 cap = a.sizes[b] - RC_STATIC_HDR;
        bool over = false;
        rc_static_encode_exact(
            n_b, run, [&](u32 c) -> u32 { return cum[c]; }, [&](u32 i) -> u32 { return blk[i]; },
            [&](u8 byte) {
                if(at < cap) {
                    coded[at] = byte;
                } else {
                    over = true;
                }
                ++at;
            },
            [&](u32 i, u32 shifted, u32 lo2, u32 range) {
                if(rrow && i != 0u && i % a.seg_syms == 0u) {
                    u32* r = rrow + (i / a.seg_syms - 1u) * 3u;
                    const u32 sh2 = (run & (run - 1u)) == 0u ? 31u - rc_clz(run) : 0u;  // as k_enc_ranges writes it
                    r[0] = shifted;
                    r[1] = lo2;
                    r[2] = (range >> sh2) << sh2;
                }
            });
        if(over || at != cap) {
            atomicOr(a.err, ERR_INTERNAL);
        }
    }
}

}  // namespace b2rc
