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

template <bool WIDE>
__global__ void __launch_bounds__(32) k_enc_ranges(SegArgs a)
{
    extern __shared__ __align__(16) u8 smem[];
    const u32 sbase = smem_addr(smem);
    const u32 tiles = sbase + 256u * 32u * 2u;
    const u32 lane = lane_id();
    const u64 b0 = (u64)blockIdx.x * 32u;
    const u64 b = b0 + lane;
    const bool has = b < a.nblocks;
    const u32 n_b = has ? seg_block_len(a, b) : 0u;

    stage_tile(tiles, a.src, a.n, b0, a.block, 0, lane);
    cp_async_commit();
    u16* frq16 = reinterpret_cast<u16*>(smem);
    u32 total = 0;
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
            frq16[(8u * lane + k) * 32u + r] = (u16)f[k];
        }
        sum = __reduce_add_sync(FULL, sum);
        if(lane == r) {
            total = sum;
        }
    }
    __syncwarp();

    const u32 magic = rc_magic(total);
    const bool is_pow2 = total != 0 && (total & (total - 1u)) == 0;
    const u32 shift = is_pow2 ? 31u - rc_clz(total) : 0u;
    const u32 n_max = __reduce_max_sync(FULL, n_b);
    const bool all_pow2 = __all_sync(FULL, is_pow2 || !has);
    const bool ragged = __any_sync(FULL, n_b != n_max);
    u32* rec = a.recs + (has ? b : b0) * (u64)(a.nseg + 1u) * 2u;
    const u32 ftab = sbase + lane * 2u;
    constexpr int MAXSH = WIDE ? 3 : 2;
    const bool all16 = !WIDE && __all_sync(FULL, total == 65536u || !has);
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
    if(has) {
        const u32 nseg_b = (n_b + a.P - 1u) / a.P;
        a.sizes[b] = RC_STATIC_HDR + 5u + rec[2u * nseg_b];  // cpprcoder.h:386-395 header, buffer_, shifted bytes, low
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
