// ans_lane.cuh -- per-block arithmetic of the byte-wise rANS coder (cppans::rANS::encode /
// ::decode, cppans.h:497-564), one block per GPU lane like the range coder in rc_lane.cuh.
//
// The byte variant carries ONE 32-bit state per stream, so a block is a single serial
// chain and gets a lane; the word variant's eight interleaved states get eight lanes and
// live in b2rc_ans.cuh.  Written __host__ __device__ so tests/sim drives the same code on
// the CPU against the oracle.  Restates behaviour of the reference; no reference code is
// reused: put() here is branch free (two compares pick the 0, 1 or 2 bytes to emit), the
// division is rc_div's multiply + fix-up instead of the reference's per-symbol
// (rcp_freq, rcp_shift, bias) triple, and the decoder finds its symbol with a three-level
// 8 x 8 x 4 search over the cumulative table instead of a 16 KiB slot-to-symbol array
// (32 blocks per warp cannot afford one each).
#pragma once
#include "rc_lane.cuh"

#define ANS_HDR_BYTES 1032u        // u32 size + 257 x u32 cumulative counts (cppans.h:521-527)
#define ANS_BYTE_SCALE_BITS 14u    // rANS::ProbBits (cppans.h:27)
#define ANS_BYTE_LOW (1u << 23)    // rANS::rANSByteLowBounds (cppans.h:29)

// -------------------------------------------------------------------- encoder --
// What one put() sends to the stream: k bytes (0..2), the first one emitted -- which lands
// on the HIGHER address, the coder writes backwards -- in the higher byte of `bytes`.
struct AnsPut {
    u32 bytes, k;
};

// put (cppans.h:265-287): renormalise, then x = C(s, x) = (x / f) << 14 | x % f, + start.
// x_max = ((2^23 >> 14) << 8) * f = f << 17 (cppans.h:203); x < 2^31 always, so at most two
// bytes leave.  (x / f << 14) + x % f + start == x + start + (x / f) * (2^14 - f).
RC_HD void ans_byte_put(u32& x, u32 start, u32 f, u32 magic, AnsPut& out, bool active)
{
    const u32 x_max = f << 17;
    const bool p1 = active && x >= x_max;
    const bool p2 = active && (x >> 8) >= x_max;  // implies p1
    out.k = (p1 ? 1u : 0u) + (p2 ? 1u : 0u);
    out.bytes = p2 ? (((x & 0xFFu) << 8) | ((x >> 8) & 0xFFu)) : (p1 ? (x & 0xFFu) : 0u);
    const u32 xs = p2 ? (x >> 16) : (p1 ? (x >> 8) : x);
    u32 quo = rc_umulhi(xs, magic);
    const u32 rem = xs - quo * f;
    quo += rem >= f ? 1u : 0u;
    const u32 xn = xs + start + quo * ((1u << ANS_BYTE_SCALE_BITS) - f);
    x = active ? xn : x;
}

// Bytes on their way out, oldest in the most significant position.  `Out::word(w)` receives
// four of them as one little-endian u32 for the next lower aligned address (the oldest
// byte of the four is the most significant: highest address), `Out::byte(b)` one byte for
// the next lower address.
struct AnsByteAcc {
    u64 acc;
    u32 cnt;
};
RC_HD void ans_acc_init(AnsByteAcc& a)
{
    a.acc = 0;
    a.cnt = 0;
}
RC_HD void ans_acc_push(AnsByteAcc& a, const AnsPut& p)
{
    a.acc = (a.acc << (8u * p.k)) | p.bytes;  // p.bytes holds exactly p.k bytes
    a.cnt += p.k;
}
// at most 3 + 2 * 2 = 7 bytes are pending when this runs after every second put
template <class Out>
RC_HD void ans_acc_commit(AnsByteAcc& a, Out& out)
{
    const bool full = a.cnt >= 4u;
    const u32 w = (u32)(a.acc >> (8u * ((a.cnt - 4u) & 7u)));
    out.word(w, full);
    a.cnt -= full ? 4u : 0u;
}
template <class Out>
RC_HD void ans_acc_finish(AnsByteAcc& a, u32 x, Out& out)
{
    for(u32 i = 0; i < a.cnt; ++i) {
        out.byte((u8)(a.acc >> (8u * (a.cnt - 1u - i))));
    }
    // flush (cppans.h:289-299): the state, little endian, below everything else
    out.byte((u8)(x >> 24));
    out.byte((u8)(x >> 16));
    out.byte((u8)(x >> 8));
    out.byte((u8)x);
}

// -------------------------------------------------------------------- decoder --
// Symbol of a slot: the s with cum[s] <= v < cum[s+1], with its start and width.  Same
// shape as rc_static_find (rc_lane.cuh) minus the products: seven boundaries cum[32 j] in
// registers, seven cum[32 a + 4 j] and then five neighbours from the lane's table; 17
// carry-counted compares, two table round trips.  Tab::at(p) reads cum at POSITION
// p = symbol * Tab::UNIT.
template <class Tab>
RC_HD void ans_find(const Tab& tab, const u32 (&k1)[8], u32 v, u32& sym, u32& cum, u32& freq)
{
    constexpr u32 U = Tab::UNIT;
    const u32 nv = ~v;
    u32 a = 0, b = 0;  // boundaries ABOVE v
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int j = 1; j < 8; j += 2) {
        rc_count_gt(a, k1[j], nv);
        if(j + 1 < 8) {
            rc_count_gt(b, k1[j + 1], nv);
        }
    }
    const u32 p1 = (7u - (a + b)) * (32u * U);
    u32 e2[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int j = 1; j < 8; ++j) {
        e2[j] = tab.at(p1 + 4u * U * j);
    }
    a = 0;
    b = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int j = 1; j < 8; j += 2) {
        rc_count_gt(a, e2[j], nv);
        if(j + 1 < 8) {
            rc_count_gt(b, e2[j + 1], nv);
        }
    }
    const u32 p2 = p1 + (7u - (a + b)) * (4u * U);
    const u32 f0 = tab.at(p2), f1 = tab.at(p2 + U), f2 = tab.at(p2 + 2 * U), f3 = tab.at(p2 + 3 * U),
              f4 = tab.at(p2 + 4 * U);
    const bool q1 = f1 <= v, q2 = f2 <= v, q3 = f3 <= v;  // monotone: q1 >= q2 >= q3
    sym = p2 / U + (q1 ? 1u : 0u) + (q2 ? 1u : 0u) + (q3 ? 1u : 0u);
    cum = q3 ? f3 : (q2 ? f2 : (q1 ? f1 : f0));
    const u32 nxt = q3 ? f4 : (q2 ? f3 : (q1 ? f2 : f1));
    freq = nxt - cum;
}

// The stream window is RcDec's (w_hi, w_lo, wbits kept >= 32, topped up by rc_dec_refill);
// low/range are not used.  `skip` = bytes of the first word that precede the coded bytes.
// Returns the initial state: init_decode reads four bytes little endian (cppans.h:303-310).
template <class Next>
RC_HD u32 ans_byte_dec_init(RcDec& d, u32 skip, Next& next)
{
    d.low = 0;
    d.range = 0;
    d.w_hi = next();
    d.w_lo = next();
    const u32 drop = skip * 8u;
    d.w_hi = rc_funnel_l(d.w_lo, d.w_hi, drop);
    d.w_lo <<= drop;
    const u32 first = d.w_hi;  // coded bytes 0..3, big endian
    d.w_hi = d.w_lo;
    d.w_lo = 0;
    d.wbits = 32 - (s32)drop;
    rc_dec_refill_pair(d, next);
    return rc_bswap(first);
}

// The window alone, for a decoder that starts inside the stream with a state it was given
// (restart points): `skip` bytes of the first word precede the next coded byte.
template <class Next>
RC_HD void ans_byte_win_init(RcDec& d, u32 skip, Next& next)
{
    d.low = 0;
    d.range = 0;
    d.w_hi = next();
    d.w_lo = next();
    const u32 drop = skip * 8u;
    d.w_hi = rc_funnel_l(d.w_lo, d.w_hi, drop);
    d.w_lo <<= drop;
    d.wbits = 64 - (s32)drop;
}

// advance (cppans.h:321-334): x = f * (x >> 14) + slot - start, then 0, 1 or 2 bytes in.
// A valid stream never needs a third (x >= 2^9 after the update); a corrupt one may, and
// then x stays below 2^23 -- the caller checks.  At most 16 bits per symbol: the window is
// topped up after every second symbol (REFILL, rc_dec_refill_pair).
template <bool REFILL = true, class Next>
RC_HD void ans_byte_advance(RcDec& d, u32& x, u32 slot, u32 start, u32 f, Next& next)
{
    const u32 xn = f * (x >> ANS_BYTE_SCALE_BITS) + slot - start;
    const u32 sh = xn < (1u << 15) ? 16u : (xn < ANS_BYTE_LOW ? 8u : 0u);
    x = rc_funnel_l(d.w_hi, xn, sh);
    d.w_hi = rc_funnel_l(d.w_lo, d.w_hi, sh);
    d.w_lo <<= sh;
    d.wbits -= (s32)sh;
    if(REFILL) {
        rc_dec_refill_pair(d, next);
    }
}
