// rc_lane.cuh -- the per-block coder arithmetic, one block per GPU lane.
//
// Every kernel in this library maps ONE independent block to ONE lane: a warp
// advances 32 blocks in lock step, tables live in shared memory interleaved by
// lane ([entry][lane], bank == lane, conflict free) and nothing in the hot loop
// needs a shuffle.  The only serial dependency of an order-0 range coder is
//     range' = norm(freq * (range / total))
// per block; mapping blocks to lanes keeps 32 of those chains in flight per warp
// instead of one.
//
// This header holds the arithmetic of one lane and is written __host__ __device__
// so tests/sim can drive exactly the same code on the CPU against the oracle.
// It restates behaviour of the reference (cpprcoder.h, cited per function); no
// reference code is reused: the encoder here is a carry-save formulation (a wide
// shift register with deferred word emission) instead of the reference's
// byte-at-a-time buffer_/count_ state machine.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RC_HD __host__ __device__ __forceinline__
#define RC_COLD __host__ __device__ __noinline__
#else
#define RC_HD inline
#define RC_COLD inline
#endif

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t s32;
typedef int64_t s64;

#define RC_MIN_RANGE 0x01000000u      // cpprcoder.h:327, :631
#define RC_STATIC_RANGE0 0xFFFFFFFFu  // cpprcoder.h:326
#define RC_ADAPT_RANGE0 0xFFFFFF00u   // cpprcoder.h:630 (the decoder reaches it after its first normalize, :813,:929)
#define RC_STATIC_HDR 516u            // cpprcoder.h:331
#define RC_ADAPT_HDR 4u               // cpprcoder.h:689-694

// ------------------------------------------------------------------ bit helpers --
RC_HD u32 rc_funnel_l(u32 lo, u32 hi, u32 s)  // high word of (hi:lo) << s, s in [0,31]
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s ? ((hi << s) | (lo >> (32u - s))) : hi;
#endif
}
RC_HD u32 rc_funnel_r(u32 lo, u32 hi, u32 s)  // low word of (hi:lo) >> s, s in [0,31]
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? ((lo >> s) | (hi << (32u - s))) : lo;
#endif
}
RC_HD u32 rc_clz(u32 x)
{
#if defined(__CUDA_ARCH__)
    return (u32)__clz((int)x);
#else
    return x ? (u32)__builtin_clz(x) : 32u;
#endif
}
RC_HD u32 rc_umulhi(u32 a, u32 b)
{
#if defined(__CUDA_ARCH__)
    // mul.wide + take the upper register: __umulhi becomes IMAD.HI.U32, which on sm_100a is a
    // slow scoreboarded instruction (about 19 cycles measured, profiles/r1_ncu_notes.md)
    unsigned long long p;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
    return (u32)(p >> 32);
#else
    return (u32)(((u64)a * b) >> 32);
#endif
}
RC_HD u32 rc_bswap(u32 x)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}

// ------------------------------------------------------------------- division --
// t = range / total is the one divide per symbol (cpprcoder.h:401, :501, :703, :904).
// magic = floor(2^32 / d) gives q or q-1 from one mulhi; one compare repairs it.
// d == 1 uses magic 0xFFFFFFFF and is repaired by the same compare.
RC_HD u32 rc_magic(u32 d)
{
    if(d <= 1) {
        return 0xFFFFFFFFu;
    }
    u32 m = 0xFFFFFFFFu / d;
    if((d & (d - 1)) == 0) {  // d divides 2^32
        m += 1;
    }
    return m;
}
RC_HD u32 rc_div(u32 x, u32 d, u32 magic)
{
    u32 q = rc_umulhi(x, magic);
    u32 r = x - q * d;
    return q + (r >= d ? 1u : 0u);
}

// -------------------------------------------------------------------- encoder --
// Carry-save restatement of RangeEncoder::encode (cpprcoder.h:400-457) and
// AdaptiveRangeEncoder::encode/normalize/finish (cpprcoder.h:697-802).
//
// The coded bytes are one big-endian integer: sum over symbols of cum*t placed at
// the byte position the renormalisation shifts define.  The lane keeps the tail of
// that integer in a shift register  o(64) : low(32);  `ocnt` bits of o are valid.
// Whenever 32 valid bits have left `low` they are cut off as one stream word.
// A carry out of `low` ripples into o by plain addition; a carry out of o's valid
// bits is parked just above them and is applied, when the next word is cut, to the
// newest word that is not all ones (`pend`) -- all-ones words after it are only
// counted (`nff`), which is the reference's buffer_/count_ idea at word size.
struct RcEnc {
    u32 low, range;
    u32 o_lo, o_hi;
    s32 ocnt;       // valid bits in o (multiple of 8, < 32 between steps)
    u32 pend, nff;  // deferred word and the number of 0xFFFFFFFF words behind it
};

// Sink contract: push(word) appends one big-endian stream word.  The FIRST word an
// encoder pushes is a placeholder (the initial `pend`) and must be discarded by the
// sink -- that spares the hot path a "has a deferred word yet?" test.  push() on the
// hot path may assume room: rc_enc_commit asks tight(n) first and takes the rare path
// when fewer than n words are left.  Sink::Checked is the variant the rare path uses
// (runs of any length, bounds checked); Checked::settle(sink) writes its bookkeeping back.
RC_HD void rc_enc_init(RcEnc& e, u32 range0)
{
    e.low = 0;
    e.range = range0;
    e.o_lo = 0;
    e.o_hi = 0;
    e.ocnt = 8;  // the reference's initial buffer_ = 0 is the first stream byte (cpprcoder.h:385, :687)
    e.pend = 0;
    e.nff = 0;
}

// (o_hi : o_lo : low) += term
RC_HD void rc_add96(u32& low, u32& o_lo, u32& o_hi, u32 term)
{
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.u32 %2, %2, 0;"
        : "+r"(low), "+r"(o_lo), "+r"(o_hi)
        : "r"(term));
#else
    const u32 nl = low + term;
    const u32 c = (nl < low) ? 1u : 0u;
    low = nl;
    const u32 ol = o_lo + c;
    o_hi += (ol < c) ? 1u : 0u;
    o_lo = ol;
#endif
}

// Rare part of cutting a word: all-ones words in play (cpprcoder.h:405-435 / :767-800 at
// word granularity).  Kept out of line; callers pass copies so that the lane state proper
// never has its address taken and stays in registers.
template <class Sink, class Enc>
RC_COLD void rc_enc_word_slow(Enc& e, u32 word, u32 ovf, Sink& s)
{
    if(ovf) {  // a carry reaches the deferred word: the 0xFF run behind it rolls over to zeros
        e.pend += ovf;
        if(e.nff) {
            s.push(e.pend);
            for(u32 i = 1; i < e.nff; ++i) {
                s.push(0u);
            }
            e.pend = 0;
            e.nff = 0;
        }
    }
    if(word == 0xFFFFFFFFu) {
        e.nff += 1;  // cpprcoder.h:431 / :796
        return;
    }
    s.push(e.pend);  // cpprcoder.h:420-428 / :785-794
    for(u32 i = 0; i < e.nff; ++i) {
        s.push(0xFFFFFFFFu);
    }
    e.pend = word;
    e.nff = 0;
}

#if defined(__CUDA_ARCH__)
#define RC_WARP_ANY(p) __any_sync(0xFFFFFFFFu, (p))
#else
#define RC_WARP_ANY(p) (p)
#endif

// A stream word cut off the shift register by one step, waiting to be committed.
struct RcCut {
    u32 word, ovf;  // ovf: carry parked above the valid bits; it belongs to the word BEFORE this one
    bool on;
};

// Renormalisation shift of a fresh range: 8 bits per round while range < 2^24
// (cpprcoder.h:418, :783).  Compare-and-select instead of count-leading-zeros: the
// selects are shorter than FLO on the serial chain.  MAXSH = 2 when range >= 2^8 is
// guaranteed (total <= 2^16), else 3.
template <int MAXSH>
RC_HD u32 rc_norm_shift(u32 r)
{
    u32 sh = (r < 0x01000000u) ? 8u : 0u;
    sh = (r < 0x00010000u) ? 16u : sh;
    if(MAXSH >= 3) {
        sh = (r < 0x00000100u) ? 24u : sh;
    }
    return sh;
}

// The same shift from the position of the highest set bit: one instruction on the conversion pipe and
// one on the integer pipe instead of four (six) on the integer pipe, at a longer latency -- for kernels
// that run many warps per scheduler and are bound by the integer pipe (k_enc_seg), not for the
// one-warp-per-scheduler kernels, whose chain it would lengthen.  r >= 1.
RC_HD u32 rc_norm_shift_flo(u32 r)
{
#if defined(__CUDA_ARCH__)
    u32 top;
    asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(r));
    return ~top & 24u;
#else
    return rc_clz(r) & 24u;
#endif
}

// And as three compares in parallel and a TREE of selects, two deep (the plain form above compiles to a
// chain of three): the shortest way from r to its shift, for a lone warp whose chain runs through it.
RC_HD u32 rc_norm_shift_tree(u32 r)
{
#if defined(__CUDA_ARCH__)
    u32 sh;
    asm("{ .reg .pred p8, p16, p24;\n\t.reg .u32 a, b;\n\t"
        "setp.lt.u32 p24, %1, 0x1000000;\n\tsetp.lt.u32 p16, %1, 0x10000;\n\tsetp.lt.u32 p8, %1, 0x100;\n\t"
        "selp.u32 a, 8, 0, p24;\n\tselp.u32 b, 24, 16, p8;\n\tselp.u32 %0, b, a, p16; }"
        : "=r"(sh)
        : "r"(r));
    return sh;
#else
    return rc_clz(r) & 24u;
#endif
}

// One symbol, branch free: cum/freq from the model, t = range / total already divided.
// Leaves at most one cut word in `c`; rc_enc_commit() takes care of it later, off the chain.
// FLO: the shift by rc_norm_shift_flo.
template <int MAXSH, bool FLO = false>
RC_HD void rc_enc_step(RcEnc& e, u32 cum, u32 freq, u32 t, RcCut& c, bool active = true)
{
    u32 sh = 0;
    if(active) {
        rc_add96(e.low, e.o_lo, e.o_hi, cum * t);
        e.range = freq * t;
        sh = FLO ? rc_norm_shift_flo(e.range) : rc_norm_shift<MAXSH>(e.range);
        e.range <<= sh;
    }
    e.o_hi = rc_funnel_l(e.o_lo, e.o_hi, sh);
    e.o_lo = rc_funnel_l(e.low, e.o_lo, sh);
    e.low <<= sh;
    e.ocnt += (s32)sh;
    const bool cut = e.ocnt >= 32;
    const u32 k = (u32)(e.ocnt - 32) & 31u;  // 0..23 when cutting
    c.word = rc_funnel_r(e.o_lo, e.o_hi, k);
    c.ovf = e.o_hi >> k;
    c.on = cut;
    e.o_hi = cut ? 0u : e.o_hi;
    e.o_lo = cut ? (e.o_lo & ((1u << k) - 1u)) : e.o_lo;
    e.ocnt = cut ? (s32)k : e.ocnt;
}

// Power-of-two total (every full block of the static coder that never halved): the
// chain is carried by t = range >> shift alone: t' = (freq * t << sh) >> shift.  One link of
// the chain is IMAD -> compare -> select -> shift -> shift; a lone warp issues roughly one
// instruction every two cycles, so the instruction count, not the last few cycles of chain,
// decides (profiles/r1_ncu_notes.md).  `e.range` is not maintained on this path.
template <int MAXSH, bool FLO = false>
RC_HD void rc_enc_step_pow2(RcEnc& e, u32& t, u32 shift, u32 cum, u32 freq, RcCut& c, bool active = true)
{
    u32 sh = 0;
    if(active) {
        rc_add96(e.low, e.o_lo, e.o_hi, cum * t);
        const u32 r = freq * t;
        sh = FLO ? rc_norm_shift_flo(r) : rc_norm_shift<MAXSH>(r);
        t = (r << sh) >> shift;  // two shifts on the chain instead of three candidates and two selects: fewer instructions
    }
    e.o_hi = rc_funnel_l(e.o_lo, e.o_hi, sh);
    e.o_lo = rc_funnel_l(e.low, e.o_lo, sh);
    e.low <<= sh;
    e.ocnt += (s32)sh;
    const bool cut = e.ocnt >= 32;
    const u32 k = (u32)(e.ocnt - 32) & 31u;
    c.word = rc_funnel_r(e.o_lo, e.o_hi, k);
    c.ovf = e.o_hi >> k;
    c.on = cut;
    e.o_hi = cut ? 0u : e.o_hi;
    e.o_lo = cut ? (e.o_lo & ((1u << k) - 1u)) : e.o_lo;
    e.ocnt = cut ? (s32)k : e.ocnt;
}

// ---------------------------------------------------- the encoder step, multiplier form --
// The same step as rc_enc_step_pow2 / rc_enc_step with the shifting done by MULTIPLYING: the integer
// pipe of an SM issues half as many warp instructions per cycle as the scheduler can hand out, and the
// kernels that run many chains per scheduler (k_enc_seg) are bound by exactly that pipe; the
// multiplier pipe idles.  With m = 2^sh (1, 2^8, 2^16 or 2^24):
//     (o : low) += cum * t                    one 32 x 32 + 64 multiply-add, the carry ripples into o
//     low * m                                 low half: the new low; high half: the bytes leaving low
//     o * m + those bytes                     the new o, up to 25 + 24 bits
//     2^ocnt * m                              high half non-zero <=> 32 valid bits have left low: cut a
//                                             word; it IS 2^k, k the bits that stay behind
// so that o, kept below 2^32 between steps, and 2^ocnt replace the three-word shift register, its
// bit count, the variable funnel shifts, the mask and the selects of rc_enc_step.
struct RcEnc2 {
    u64 x;          // o : low
    u32 oc;         // 2^ocnt, ocnt in {0, 8, 16, 24}: valid bits in o between steps
    u32 range;
    u32 pend, nff;  // as RcEnc
};

RC_HD u32 rc_norm_mult(u32 r, int maxsh)  // 2^(renormalisation shift of r), see rc_norm_shift
{
#if defined(__CUDA_ARCH__) && !defined(RC_NORM_BY_COMPARES)
    // 8 * (leading zero bytes of r) from the position of its highest bit (one instruction on the
    // conversion pipe instead of two compares and two selects on the integer pipe), and through asm:
    // a compiler that sees "multiply by a power of two" turns the multiplications of the step back
    // into variable shifts, which is the integer-pipe work this form exists to avoid
    (void)maxsh;  // 2^8 <= r is the caller's promise when maxsh == 2; the same code serves both
    u32 m;
    asm("{ .reg .u32 top, sh;\n\tbfind.u32 top, %1;\n\tlop3.b32 sh, top, 24, 0, 0x0C;\n\tshl.b32 %0, 1, sh; }"
        : "=r"(m)
        : "r"(r));  // lop3 0x0C: ~a & b
    return m;
#elif defined(__CUDA_ARCH__)
    u32 m;  // two compares and two selects, through asm for the same reason
    asm("{ .reg .pred p, q;\n\tsetp.lt.u32 p, %1, 0x01000000;\n\tsetp.lt.u32 q, %1, 0x00010000;\n\t"
        "selp.u32 %0, 0x100, 1, p;\n\tselp.u32 %0, 0x10000, %0, q; }"
        : "=r"(m)
        : "r"(r));
    if(maxsh >= 3) {
        asm("{ .reg .pred p;\n\tsetp.lt.u32 p, %1, 0x00000100;\n\tselp.u32 %0, 0x1000000, %0, p; }" : "+r"(m) : "r"(r));
    }
    return m;
#else
    u32 m = (r < 0x01000000u) ? 0x100u : 1u;
    m = (r < 0x00010000u) ? 0x10000u : m;
    if(maxsh >= 3) {
        m = (r < 0x00000100u) ? 0x1000000u : m;
    }
    return m;
#endif
}

// 32 x 32 -> (lo, hi) and 32 x 32 + 32 -> (lo, hi), the halves as separate 32-bit values so that no
// 64-bit compare or shift is ever made of them
RC_HD void rc_mul_wide(u32 a, u32 b, u32& lo, u32& hi)
{
#if defined(__CUDA_ARCH__)
    asm("{ .reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#else
    const u64 p = (u64)a * b;
    lo = (u32)p;
    hi = (u32)(p >> 32);
#endif
}
RC_HD void rc_mad_wide(u32 a, u32 b, u32 c, u32& lo, u32& hi)
{
#if defined(__CUDA_ARCH__)
    asm("{ .reg .u64 p, q;\n\tcvt.u64.u32 q, %4;\n\tmad.wide.u32 p, %2, %3, q;\n\tmov.b64 {%0, %1}, p; }"
        : "=r"(lo), "=r"(hi)
        : "r"(a), "r"(b), "r"(c));
#else
    const u64 p = (u64)a * b + c;
    lo = (u32)p;
    hi = (u32)(p >> 32);
#endif
}
RC_HD u64 rc_mad_wide64(u32 a, u32 b, u64 c)
{
#if defined(__CUDA_ARCH__)
    u64 p;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(a), "r"(b), "l"(c));
    return p;
#else
    return (u64)a * b + c;
#endif
}

// everything of a step behind "which multiplier": shift, cut.  When nothing is cut, c.word is o
// itself (below 2^25): never all ones, which rc_enc2_commit relies on.
template <int MAXSH>
RC_HD void rc_enc2_shift(RcEnc2& e, u32 m, RcCut& c)
{
    u32 l_lo, l_hi, o_lo, o_hi, c_lo, chi;
    rc_mul_wide((u32)e.x, m, l_lo, l_hi);
    rc_mad_wide((u32)(e.x >> 32), m, l_hi, o_lo, o_hi);
    rc_mul_wide(e.oc, m, c_lo, chi);  // chi: 0, or 2^k with k = ocnt + sh - 32 in {0, 8, 16}
    const u32 k = MAXSH >= 3 ? (((chi >> 5) & 8u) | ((chi >> 12) & 16u)) : (chi >> 5);
    c.word = rc_funnel_r(o_lo, o_hi, k);
    c.ovf = o_hi >> k;
    c.on = chi != 0u;
    e.x = ((u64)(o_lo & (chi - 1u)) << 32) | l_lo;  // chi == 0: nothing is cut, the mask is all ones
    e.oc = chi | c_lo;
}

// power-of-two total: the chain carries t = range >> shift (rc_enc_step_pow2)
template <int MAXSH>
RC_HD void rc_enc2_step_pow2(RcEnc2& e, u32& t, u32 shift, u32 cum, u32 freq, RcCut& c, bool active = true)
{
    u32 m = 1u;
    if(active) {
        e.x = rc_mad_wide64(cum, t, e.x);
        const u32 r = freq * t;
        m = rc_norm_mult(r, MAXSH);
        t = (r * m) >> shift;
    }
    rc_enc2_shift<MAXSH>(e, m, c);
}

template <int MAXSH>
RC_HD void rc_enc2_step(RcEnc2& e, u32 cum, u32 freq, u32 t, RcCut& c, bool active = true)
{
    u32 m = 1u;
    if(active) {
        e.x = rc_mad_wide64(cum, t, e.x);
        const u32 r = freq * t;
        m = rc_norm_mult(r, MAXSH);
        e.range = r * m;
    }
    rc_enc2_shift<MAXSH>(e, m, c);
}

RC_HD RcEnc rc_enc2_view(const RcEnc2& e)  // the same state as the three-word register of RcEnc
{
    RcEnc v;
    v.low = (u32)e.x;
    v.range = e.range;
    v.o_lo = (u32)(e.x >> 32);
    v.o_hi = 0;
    v.ocnt = (s32)(31u - rc_clz(e.oc));
    v.pend = e.pend;
    v.nff = e.nff;
    return v;
}

// Commits the cuts of up to N consecutive steps, in order.  The common case is straight
// line: the parked carry goes into the deferred word, which is pushed, and the new word
// becomes the deferred one.  All-ones words (cpprcoder.h:431, :796) take the rare path,
// entered by the whole warp on one vote.  CONVERGED: every lane must call this together.
template <int N, class Sink, class Enc>
RC_HD void rc_enc_commit(Enc& e, const RcCut (&c)[N], Sink& s)
{
    bool rare = e.nff != 0u || s.tight(N);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int k = 0; k < N; ++k) {
        rare = rare || (c[k].on && c[k].word == 0xFFFFFFFFu);
    }
    if(RC_WARP_ANY(rare)) {
        if(rare) {
            Enc te = e;
            typename Sink::Checked ts(s);  // may push a long run: this one checks for room
            for(int k = 0; k < N; ++k) {
                if(c[k].on) {
                    rc_enc_word_slow(te, c[k].word, c[k].ovf, ts);
                }
            }
            e.pend = te.pend;
            e.nff = te.nff;
            ts.settle(s);
            return;
        }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int k = 0; k < N; ++k) {
        if(c[k].on) {
            s.push(e.pend + c[k].ovf);
            e.pend = c[k].word;
        }
    }
}

// rc_enc_commit for cuts made by rc_enc2_shift: a word that was not cut is never all ones there, so
// "some cut word is all ones" is one maximum over the words and one compare.
template <class Sink>
RC_HD void rc_enc2_commit(RcEnc2& e, const RcCut (&c)[4], Sink& s)
{
    u32 top = c[0].word > c[1].word ? c[0].word : c[1].word;
    top = top > c[2].word ? top : c[2].word;
    top = top > c[3].word ? top : c[3].word;
    const bool rare = e.nff != 0u || top == 0xFFFFFFFFu;  // all-ones words in play (cpprcoder.h:431)
    const bool edge = s.tight(4);                         // first words of a segment / little room left
    if(RC_WARP_ANY(rare || edge)) {
        if(RC_WARP_ANY(rare)) {
            if(rare) {
                RcEnc2 te = e;
                typename Sink::Checked ts(s);  // may push a long run: this one checks for room
                for(int k = 0; k < 4; ++k) {
                    if(c[k].on) {
                        rc_enc_word_slow(te, c[k].word, c[k].ovf, ts);
                    }
                }
                e.pend = te.pend;
                e.nff = te.nff;
                ts.settle(s);
                return;
            }
        }
        // no all-ones word anywhere near: the straight-line commit through the checked sink
        typename Sink::Checked ts(s);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int k = 0; k < 4; ++k) {
            if(c[k].on) {
                ts.push(e.pend + c[k].ovf);
                e.pend = c[k].word;
            }
        }
        ts.settle(s);
        return;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int k = 0; k < 4; ++k) {
        if(c[k].on) {
            s.push(e.pend + c[k].ovf);
            e.pend = c[k].word;
        }
    }
}

// The same commit for cuts made by rc_enc_step / rc_enc_step_pow2 (the funnel-shift form), whose
// c.word is arbitrary when nothing was cut: masked first.
template <class Enc, class Sink>
RC_HD void rc_enc_commit_edge(Enc& e, const RcCut (&c)[4], Sink& s)
{
    const u32 w0 = c[0].on ? c[0].word : 0u, w1 = c[1].on ? c[1].word : 0u, w2 = c[2].on ? c[2].word : 0u,
              w3 = c[3].on ? c[3].word : 0u;
    u32 top = w0 > w1 ? w0 : w1;
    top = top > w2 ? top : w2;
    top = top > w3 ? top : w3;
    const bool rare = e.nff != 0u || top == 0xFFFFFFFFu;
    const bool edge = s.tight(4);
    if(RC_WARP_ANY(rare || edge)) {
        if(RC_WARP_ANY(rare)) {
            if(rare) {
                Enc te = e;
                typename Sink::Checked ts(s);
                for(int k = 0; k < 4; ++k) {
                    if(c[k].on) {
                        rc_enc_word_slow(te, c[k].word, c[k].ovf, ts);
                    }
                }
                e.pend = te.pend;
                e.nff = te.nff;
                ts.settle(s);
                return;
            }
        }
        typename Sink::Checked ts(s);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int k = 0; k < 4; ++k) {
            if(c[k].on) {
                ts.push(e.pend + c[k].ovf);
                e.pend = c[k].word;
            }
        }
        ts.settle(s);
        return;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int k = 0; k < 4; ++k) {
        if(c[k].on) {
            s.push(e.pend + c[k].ovf);
            e.pend = c[k].word;
        }
    }
}

// End of block (cpprcoder.h:439-457 / :744-762).  Pushes the deferred words and
// returns the bytes that do not fill a word: `tail[0..ntail)`, ntail in 4..7.
// The caller must have dealt with the static coder's low_ == 0xFFFFFFFF quirk
// (cpprcoder.h:440-443) BEFORE calling: that case is re-encoded by rc_static_encode_exact.
template <class Sink>
RC_HD u32 rc_enc_finish(RcEnc& e, Sink& s, u8 tail[8])
{
    const u64 o = ((u64)e.o_hi << 32) | e.o_lo;
    const u32 ovf = (u32)(o >> e.ocnt);
    if(ovf) {
        e.pend += ovf;
        if(e.nff) {
            s.push(e.pend);
            for(u32 i = 1; i < e.nff; ++i) {
                s.push(0u);
            }
            e.pend = 0;
            e.nff = 0;
        }
    }
    s.push(e.pend);
    for(u32 i = 0; i < e.nff; ++i) {
        s.push(0xFFFFFFFFu);
    }
    u32 n = 0;
    for(s32 b = e.ocnt - 8; b >= 0; b -= 8) {
        tail[n++] = (u8)(e.o_lo >> b);
    }
    tail[n++] = (u8)(e.low >> 24);
    tail[n++] = (u8)(e.low >> 16);
    tail[n++] = (u8)(e.low >> 8);
    tail[n++] = (u8)(e.low);
    return n;
}

// Reference-shaped byte-at-a-time static encoder for ONE block.  Used only for the
// 2^-32 flush quirk (final low_ == 0xFFFFFFFF, cpprcoder.h:439-451), where the
// reference's output is NOT the big-endian sum, and as a cross-check in tests.
// cum[257] is the exclusive prefix of the block's frequencies.  `put(byte)` appends.
// `mark(i, shifted, low, range)` sees the coder before symbol i: the bytes shifted out of low so
// far, low and range -- what a restart point records.
template <class CumAt, class SymAt, class Put, class Mark>
RC_HD void rc_static_encode_exact(u32 n, u32 total, CumAt cum_at, SymAt sym_at, Put put, Mark mark)
{
    u32 range = RC_STATIC_RANGE0, low = 0, run = 0, held = 0, shifted = 0;
    const u32 magic = rc_magic(total);
    for(u32 i = 0; i < n; ++i) {
        mark(i, shifted, low, range);
        const u32 c = sym_at(i);
        const u32 t = rc_div(range, total, magic);
        const u32 lo_c = cum_at(c);
        const u32 next = low + lo_c * t;
        range = (cum_at(c + 1) - lo_c) * t;
        if(next < low) {
            ++held;
            for(; run != 0; --run) {
                put((u8)held);
                held = 0;
            }
        }
        low = next;
        while(range < RC_MIN_RANGE) {
            if(low < 0xFF000000u) {
                put((u8)held);
                for(; run != 0; --run) {
                    put((u8)0xFF);
                }
                held = low >> 24;
            } else {
                ++run;
            }
            low <<= 8;
            range <<= 8;
            ++shifted;
        }
    }
    u8 fill = 0xFF;
    if(low == 0xFFFFFFFFu) {
        ++held;
        fill = 0;
    }
    put((u8)held);
    for(; run != 0; --run) {
        put(fill);
    }
    put((u8)(low >> 24));
    put((u8)(low >> 16));
    put((u8)(low >> 8));
    put((u8)low);
}

template <class CumAt, class SymAt, class Put>
RC_HD void rc_static_encode_exact(u32 n, u32 total, CumAt cum_at, SymAt sym_at, Put put)
{
    rc_static_encode_exact(n, total, cum_at, sym_at, put, [](u32, u32, u32, u32) {});
}

// cnt += (prod > low), without a predicate: prod + ~low carries out exactly when
// prod >= low + 1; the carry is added in.  `nlow` is ~low.  (add.cc/addc rather than
// sub.cc/addc: mixing a borrow with a carry-in is not something PTX pins down.)
RC_HD void rc_count_gt(u32& cnt, u32 prod, u32 nlow)
{
#if defined(__CUDA_ARCH__)
    asm("{ .reg .u32 d; add.cc.u32 d, %1, %2;\n\taddc.u32 %0, %0, 0; }" : "+r"(cnt) : "r"(prod), "r"(nlow));
#else
    cnt += ((u64)prod + nlow > 0xFFFFFFFFull) ? 1u : 0u;
#endif
}

// Static symbol search: smallest s with cum[s+1] * t > low, i.e. RangeEncoder::find
// (cpprcoder.h:521-535) applied to low / t, done in the product domain so the decoder's
// second divide (cpprcoder.h:502) disappears: cum*t <= low  <=>  cum <= low / t, and
// cum*t never overflows because cum <= total and total * t <= range.
// Three levels, 8 x 8 x 4: seven boundaries cum[32j] held in registers, seven cum[32a+4j]
// and then five neighbours cum[4g .. 4g+4] from the table -- 17 compares instead of the
// reference's 8 dependent steps, two table round trips, and the last level hands back
// cum and freq of the symbol without another lookup.  (IMAD.HI is avoided on purpose:
// it is a slow, scoreboarded instruction on sm_100a.)
// Tab::at(p) reads cum at POSITION p = symbol * Tab::UNIT, so that on the device the
// position is the shared-memory byte offset and no index arithmetic sits on the chain.
template <class Tab>
RC_HD void rc_static_find(const Tab& tab, const u32 (&k1)[8], u32 t, u32 low, u32& sym, u32& cum, u32& freq)
{
    constexpr u32 U = Tab::UNIT;
    const u32 nlow = ~low;
    u32 a = 0, b = 0;  // boundaries ABOVE low
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int j = 1; j < 8; j += 2) {
        rc_count_gt(a, k1[j] * t, nlow);
        if(j + 1 < 8) {
            rc_count_gt(b, k1[j + 1] * t, nlow);
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
        rc_count_gt(a, e2[j] * t, nlow);
        if(j + 1 < 8) {
            rc_count_gt(b, e2[j + 1] * t, nlow);
        }
    }
    const u32 p2 = p1 + (7u - (a + b)) * (4u * U);
    const u32 f0 = tab.at(p2), f1 = tab.at(p2 + U), f2 = tab.at(p2 + 2 * U), f3 = tab.at(p2 + 3 * U),
              f4 = tab.at(p2 + 4 * U);
    const bool q1 = f1 * t <= low, q2 = f2 * t <= low, q3 = f3 * t <= low;  // monotone: q1 >= q2 >= q3
    sym = p2 / U + (q1 ? 1u : 0u) + (q2 ? 1u : 0u) + (q3 ? 1u : 0u);
    cum = q3 ? f3 : (q2 ? f2 : (q1 ? f1 : f0));
    const u32 nxt = q3 ? f4 : (q2 ? f3 : (q1 ? f2 : f1));
    freq = nxt - cum;
}

// The same search in four levels of four (4 x 4 x 4 x 4): 12 compares and three table round
// trips instead of 17 and two.  Fewer instructions, a longer chain: for the segmented decoder,
// whose warps share schedulers and are bound by issue slots, not by the latency of one chain.
// k0[j] = cum[64 j], j = 1..3.
// (Round 2 tried the compares as wide multiply-adds -- the high half of k * t + (2^64 - 1 - low) is
// all ones exactly when k * t <= low: one instruction on the multiplier pipe instead of a multiply
// and two adds on the integer pipe, which ncu shows 68 % busy in k_dec_static_seg.  95 -> 80
// instructions per symbol and 3.97 -> 4.9 ms: IMAD.WIDE with a 64-bit addend is far slower than
// its count suggests.  profiles/r2_ncu_notes.md.)
template <class Tab>
RC_HD void rc_static_find4(const Tab& tab, const u32 (&k0)[4], u32 t, u32 low, u32& sym, u32& cum, u32& freq)
{
    constexpr u32 U = Tab::UNIT;
    const u32 nlow = ~low;
    u32 a = 0;  // boundaries ABOVE low
    rc_count_gt(a, k0[1] * t, nlow);
    rc_count_gt(a, k0[2] * t, nlow);
    rc_count_gt(a, k0[3] * t, nlow);
    const u32 p1 = (3u - a) * (64u * U);
    const u32 e1 = tab.at(p1 + 16u * U), e2 = tab.at(p1 + 32u * U), e3 = tab.at(p1 + 48u * U);
    a = 0;
    rc_count_gt(a, e1 * t, nlow);
    rc_count_gt(a, e2 * t, nlow);
    rc_count_gt(a, e3 * t, nlow);
    const u32 p2 = p1 + (3u - a) * (16u * U);
    const u32 g1 = tab.at(p2 + 4u * U), g2 = tab.at(p2 + 8u * U), g3 = tab.at(p2 + 12u * U);
    a = 0;
    rc_count_gt(a, g1 * t, nlow);
    rc_count_gt(a, g2 * t, nlow);
    rc_count_gt(a, g3 * t, nlow);
    const u32 p3 = p2 + (3u - a) * (4u * U);
    const u32 f0 = tab.at(p3), f1 = tab.at(p3 + U), f2 = tab.at(p3 + 2 * U), f3 = tab.at(p3 + 3 * U),
              f4 = tab.at(p3 + 4 * U);
    const bool q1 = f1 * t <= low, q2 = f2 * t <= low, q3 = f3 * t <= low;  // monotone: q1 >= q2 >= q3
    sym = p3 / U + (q1 ? 1u : 0u) + (q2 ? 1u : 0u) + (q3 ? 1u : 0u);
    cum = q3 ? f3 : (q2 ? f2 : (q1 ? f1 : f0));
    const u32 nxt = q3 ? f4 : (q2 ? f3 : (q1 ? f2 : f1));
    freq = nxt - cum;
}

// -------------------------------------------------------------------- decoder --
// Restates the state handling of RangeEncoder::decode (cpprcoder.h:494-517) and
// AdaptiveRangeDecoder::decode/normalize (cpprcoder.h:889-917, :926-940).
// Input bytes wait in a 64-bit big-endian window (`wbits` valid bits, kept >= 32)
// that is topped up one aligned stream word at a time.
struct RcDec {
    u32 low, range;
    u32 w_hi, w_lo;
    s32 wbits;
};

// `next()` returns the next 4 stream bytes as a big-endian word (zeros past the end).
// `skip` = bytes of the first word that precede the coded stream (0..3).
// Both coders start with low = coded bytes 1..4: the static decoder skips byte 0
// outright (cpprcoder.h:494-498); the adaptive one loads bytes 0..3 and its first
// normalize shifts byte 0 (always 0) out while pulling byte 4 (cpprcoder.h:889-892, :929-938).
template <class Next>
RC_HD void rc_dec_init(RcDec& d, u32 range0, u32 skip, Next& next)
{
    d.w_hi = next();
    d.w_lo = next();
    const u32 drop = (skip + 1u) * 8u;  // alignment bytes + the dummy first byte
    if(drop == 32u) {
        d.w_hi = d.w_lo;
        d.w_lo = next();
    } else {
        d.w_hi = rc_funnel_l(d.w_lo, d.w_hi, drop);
        d.w_lo <<= drop;
        const u32 w = next();  // window holds 64 - drop bits; append 32 more below them
        d.w_lo |= w >> (32u - drop);
        d.low = d.w_hi;
        d.w_hi = d.w_lo;
        d.w_lo = w << drop;
        d.wbits = 64 - (s32)drop;
        d.range = range0;
        return;
    }
    d.low = d.w_hi;
    d.w_hi = d.w_lo;
    d.w_lo = 0;
    d.wbits = 32;
    d.range = range0;
}

// Tops the window up when fewer than 32 bits are left (then 8, 16 or 24 valid bits sit in
// w_hi and w_lo is empty).  Branch free: `next.take(need)` hands out the next stream word
// and moves on only when `need` is set, so the whole warp runs the same few instructions
// whether or not a given lane refills on this symbol.
template <class Next>
RC_HD void rc_dec_refill(RcDec& d, Next& next)
{
    const bool need = d.wbits < 32;
    const u32 w = next.take(need);
    const u32 have = (u32)d.wbits & 31u;
    d.w_hi = need ? (d.w_hi | (w >> have)) : d.w_hi;
    d.w_lo = need ? (w << (32u - have)) : d.w_lo;
    d.wbits += need ? 32 : 0;
}

// The same for a window that is topped up after every SECOND symbol (a symbol takes at most 16
// bits, so >= 32 bits last for two): the pair may drain it exactly, have == 0, and the shift
// above would be by 32.  Kept apart from rc_dec_refill on purpose: k_dec_static's hot loop
// lost 7 % when this form replaced the plain shift there (profiles/r1_ncu_notes.md).
template <class Next>
RC_HD void rc_dec_refill_pair(RcDec& d, Next& next)
{
    const bool need = d.wbits < 32;
    const u32 w = next.take(need);
    const u32 have = (u32)d.wbits & 31u;
    d.w_hi = need ? (d.w_hi | (w >> have)) : d.w_hi;
    d.w_lo = need ? rc_funnel_r(0u, w, have) : d.w_lo;  // w << (32 - have), and 0 for have == 0
    d.wbits += need ? 32 : 0;
}

// After the symbol is known: low -= cum*t, range = freq*t, renormalise, refill.
template <class Next>
RC_HD void rc_dec_advance(RcDec& d, u32 cum, u32 freq, u32 t, Next& next)
{
    d.low -= cum * t;
    d.range = freq * t;
    const u32 sh = rc_norm_shift<3>(d.range);
    d.range <<= sh;
    d.low = rc_funnel_l(d.w_hi, d.low, sh);
    d.w_hi = rc_funnel_l(d.w_lo, d.w_hi, sh);
    d.w_lo <<= sh;
    d.wbits -= (s32)sh;
    rc_dec_refill(d, next);
}

// Same as rc_dec_advance for a power-of-two total: the chain is carried by
// t = range >> shift (see rc_enc_step_pow2); d.range is not maintained.
// MAXSH = 2 when range >= 2^8 is guaranteed (total <= 2^16), else 3.
template <int MAXSH, class Next>
RC_HD void rc_dec_advance_pow2(RcDec& d, u32& t, u32 shift, u32 cum, u32 freq, Next& next)
{
    d.low -= cum * t;
    const u32 r = freq * t;
    const u32 sh = rc_norm_shift<MAXSH>(r);
    t = (r << sh) >> shift;
    d.low = rc_funnel_l(d.w_hi, d.low, sh);
    d.w_hi = rc_funnel_l(d.w_lo, d.w_hi, sh);
    d.w_lo <<= sh;
    d.wbits -= (s32)sh;
    rc_dec_refill(d, next);
}

// The two halves of rc_dec_advance_pow2<2> for a window that is topped up after every second
// symbol (k_dec_static_seg, where issue slots and not latency are the limit).
template <class Next>
RC_HD void rc_dec_advance_pow2_pair(RcDec& d, u32& t, u32 shift, u32 cum, u32 freq, Next& next, bool top_up)
{
    d.low -= cum * t;
    const u32 r = freq * t;
    const u32 sh = rc_norm_shift<2>(r);
    t = (r << sh) >> shift;
    d.low = rc_funnel_l(d.w_hi, d.low, sh);
    d.w_hi = rc_funnel_l(d.w_lo, d.w_hi, sh);
    d.w_lo <<= sh;
    d.wbits -= (s32)sh;
    if(top_up) {
        rc_dec_refill_pair(d, next);
    }
}

// ------------------------------------------------------------- adaptive model --
// AdaptiveFrequencyTable (cpprcoder.h:256-314, :1094-1261) for blocks short enough
// that it never halves (total = 256 + i stays below 2^24, cpprcoder.h:1138): then
//     freq_i(c) = 1 + #{j < i : b_j == c}        cum_i(c) = c + #{j < i : b_j < c}.
// Only the counts are stored; the "+1 per symbol" of initialize() is implicit.
//
// Encoder side: a binary tree of left-subtree counts.  Node (256|b) >> (l+1) holds
// how many symbols so far share b's bits above l and have bit l clear.  One walk
// both answers #{b_j < b} (sum the nodes where b's bit is set) and records b
// (increment the nodes where it is clear): 8 table touches instead of the
// reference's up-to-15 adds plus up-to-16 prefix increments (cpprcoder.h:1156-1159,
// :1179-1187).  Entries 256..511 are the per-symbol counts.
template <class Tab>
RC_HD void rc_model_encode(Tab& tab, u32 b, u32& cum, u32& freq)
{
    // All nine reads first, then the writes: the nine nodes are distinct, so nothing inside one
    // symbol depends on a store, and the table latency is paid once instead of once per level.
    const u32 leaf = 256u | b;
    u32 v[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(s32 l = 7; l >= 0; --l) {
        v[l] = tab.ld(leaf >> (l + 1));
    }
    const u32 f = tab.ld(leaf);
    u32 below = b;  // the implicit one per symbol below b
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(s32 l = 7; l >= 0; --l) {
        u32 x = v[l];
        if(b & (1u << l)) {
            below += x;  // everything in the left subtree is below b
        } else {
            x += 1u;     // b goes into the left subtree
        }
        tab.st(leaf >> (l + 1), x);  // unchanged where the bit is set
    }
    tab.st(leaf, f + 1u);
    cum = below;
    freq = 1u + f;
}

// Decoder side, same tree: walk down from the root comparing in the product
// domain -- smallest symbol whose upper bound exceeds low (the scalar branch of
// AdaptiveFrequencyTable::find, cpprcoder.h:1221-1241, for target < total).
// (base + left) * t <= low  <=>  base + left <= low / t, and never overflows
// because (base + left) <= total and total * t <= range.
template <class Tab>
RC_HD void rc_model_decode(Tab& tab, u32 low, u32 t, u32& sym, u32& cum, u32& freq)
{
    // Both children are read while the compare that chooses between them is still in flight,
    // so one level of the walk costs the compare chain, not a table round trip plus the chain.
    u32 id = 1, base = 0;
    u32 v = tab.ld(1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(s32 l = 7; l >= 0; --l) {
        const u32 vl = tab.ld(2 * id), vr = tab.ld(2 * id + 1);  // level l-1 nodes, or the two leaves
        const u32 left = v + (1u << l);                          // counts + the implicit one per symbol
        const bool right = (base + left) * t <= low;
        tab.st(id, v + (right ? 0u : 1u));
        base += right ? left : 0u;
        id = 2 * id + (right ? 1u : 0u);
        v = right ? vr : vl;
    }
    tab.st(id, v + 1u);  // id is now 256 | symbol, v its count
    sym = id & 255u;
    cum = base;
    freq = 1u + v;
}

// The same walk over a tree WITHOUT its leaves (k_dec_adaptive_seg): how many symbols so far lie in
// the subtree the walk is in comes down with it -- everything seen so far at the root, then the left
// count or the rest -- and what is left at the bottom is the symbol's own count.  Nodes 1..255 only.
template <class Tab>
RC_HD void rc_model_decode_leafless(Tab& tab, u32 low, u32 t, u32 seen, u32& sym, u32& cum, u32& freq)
{
    u32 id = 1, base = 0, sub = seen;
    for(s32 l = 7; l >= 0; --l) {
        const u32 v = tab.ld(id);
        const u32 left = v + (1u << l);
        const bool right = (base + left) * t <= low;
        tab.st(id, v + (right ? 0u : 1u));
        base += right ? left : 0u;
        sub = right ? sub - v : v;
        id = 2 * id + (right ? 1u : 0u);
    }
    sym = id & 255u;
    cum = base;
    freq = 1u + sub;
}

// ================================================== segmented static encode ==
// RangeEncoder::encode (cpprcoder.h:400-457) as MANY chains per block.  `range` never
// depends on `low` (cpprcoder.h:401-404: t = range / total; low += cum * t; range = freq * t),
// so a range-only pass (rc_range_step*) knows, for every P-th symbol, the range there and how
// many bytes the coder has shifted out by then.  A segment then codes its P symbols from low = 0
// straight into the bytes of the payload that are its own; the coded bytes are one big-endian sum
// of terms, so the payload is the sum of the segments' outputs: the four bytes of `low` a segment
// is left with overlap the first four bytes of the next segment's output and are added there
// afterwards (rc_seam_add), carry and all.
//
// Byte positions, in coded-stream coordinates (byte 0 = the reference's initial buffer_ = 0,
// always 0 in the result): a segment that starts after S bytes have been shifted out owns
// [S + 1, S' + 1), S' the same count at its end; the last segment also owns the final four bytes
// of low: [S + 1, S_total + 5).  Payload size = header + 5 + S_total, known before a byte is coded.

// One link of the range-only chain; returns the renormalisation shift in bits.
template <int MAXSH>
RC_HD u32 rc_range_step(u32& range, u32 freq, u32 t)
{
    const u32 r = freq * t;
    const u32 sh = rc_norm_shift<MAXSH>(r);
    range = r << sh;
    return sh;
}
// The same for a power-of-two total: the chain carries t = range >> shift (see rc_enc_step_pow2).
template <int MAXSH>
RC_HD u32 rc_range_step_pow2(u32& t, u32 shift, u32 freq)
{
    const u32 r = freq * t;
    const u32 sh = rc_norm_shift<MAXSH>(r);
    t = (r << sh) >> shift;
    return sh;
}

// The same chain for a total that is not a power of two, with the renormalisation taken OFF the
// dependent chain.  The link is  range' = freq * (range / total)  shifted left by whole bytes until it
// is at least 2^24; done in that order it reads  multiply -> three compares -> three selects -> shift ->
// multiply-high -> multiply -> compare -> select  (65 cycles per link for a lone warp).  Here the chain
// carries x = freq * t BEFORE the shift: (x << sh) * magic = (x * magic) << sh as 64-bit integers
// (x << sh < 2^32), so the estimate of the quotient is the high word of the wide product funnel-shifted
// by sh, and sh (from the highest set bit of x) is worked out while the multiplier runs.  Same
// arithmetic as rc_div, same results.  `bits` counts the shifts of the symbols BEFORE the newest one:
// a reader of the chain's state adds rc_norm_shift_flo(x) and takes x << that as the range.
// `ntotal` = 0 - total, handed in by the caller (as a value the compiler cannot see through, or it
// negates the quotient on the chain instead).
RC_HD void rc_range_step_div(u32& x, u32& bits, u32 freq, u32 total, u32 ntotal, u32 magic)
{
    const u32 sh = rc_norm_shift_tree(x);  // FLO here: 59 cycles per link instead of 65; it is a slow instruction
    u32 p_lo, p_hi;
    rc_mul_wide(x, magic, p_lo, p_hi);
    const u32 q = rc_funnel_l(p_lo, p_hi, sh);  // = umulhi(x << sh, magic), at most one below the quotient
    const u32 rem = (x << sh) + q * ntotal;
    const u32 x0 = q * freq;
    x = rem >= total ? x0 + freq : x0;
    bits += sh;
}

// Sink of one segment: aligned 4-byte words of the destination, which the segment shares with its
// neighbours at both ends.  The shift register starts with as many phantom (zero) bytes as the
// segment's first byte lies behind a word boundary, so every word it cuts is a destination word.
// Words that lie wholly inside the segment's own bytes are stored whole on the hot path; the
// first and last words go through the checked sink, which stores own bytes only.
struct RcSegSinkChecked;
struct RcSegSink {
    typedef RcSegSinkChecked Checked;
    u32* out;     // the aligned word that holds the segment's first own byte
    s32 wcount;   // index of the next word; -1: the encoder's placeholder push is still to come
    u32 lo, hi;   // own bytes, as offsets from `out`: [lo, hi), lo < 4
    // Only the first two pushes need care: the placeholder (index -1) and the word that holds the
    // phantom bytes (index 0).  Every later word a step cuts is a whole own word: the bytes cut
    // off the shift register are exactly the bytes the range pass counted, the deferred word lags
    // them by one, and what does not fill a word stays behind for rc_seg_end.  rc_seg_end checks
    // that the count came out as promised.
    RC_HD bool tight(int) const { return wcount < 1; }
    RC_HD void push(u32 w)
    {
        out[wcount] = rc_bswap(w);
        ++wcount;
    }
};
struct RcSegSinkChecked {
    RcSegSink s;
    RC_HD explicit RcSegSinkChecked(const RcSegSink& r) : s(r) {}
    RC_HD void push(u32 w)
    {
        if(s.wcount >= 0) {
            const u32 at = 4u * (u32)s.wcount;
            if(at >= s.lo && at + 4u <= s.hi) {
                s.out[s.wcount] = rc_bswap(w);
            } else {
                u8* ob = reinterpret_cast<u8*>(s.out);
                for(u32 k = 0; k < 4u; ++k) {
                    if(at + k >= s.lo && at + k < s.hi) {
                        ob[at + k] = (u8)(w >> (24u - 8u * k));
                    }
                }
            }
        }
        ++s.wcount;
    }
    RC_HD void settle(RcSegSink& r) { r.wcount = s.wcount; }
};

// Start of a segment whose first own byte is *first and which owns `nbytes` bytes.
RC_HD void rc_seg_begin(RcEnc& e, RcSegSink& s, u8* first, u32 nbytes, u32 range0)
{
    const u32 ph = (u32)((uintptr_t)first & 3u);
    s.out = reinterpret_cast<u32*>(first - ph);
    s.wcount = -1;
    s.lo = ph;
    s.hi = ph + nbytes;
    e.low = 0;
    e.range = range0;
    e.o_lo = 0;
    e.o_hi = 0;
    e.ocnt = (s32)(8u * ph);  // phantom bytes: they belong to whoever owns the bytes in front
    e.pend = 0;
    e.nff = 0;
}

RC_HD void rc_seg_begin(RcEnc2& e, RcSegSink& s, u8* first, u32 nbytes, u32 range0)
{
    RcEnc v;
    rc_seg_begin(v, s, first, nbytes, range0);
    e.x = 0;
    e.oc = 1u << (u32)v.ocnt;
    e.range = range0;
    e.pend = 0;
    e.nff = 0;
}

// End of a segment: the deferred words, then the bytes that do not fill a word; the last segment
// of a block also writes the four bytes of low (cpprcoder.h:453-456).  Returns false when the
// bytes written do not end where the range-only pass said they would.
RC_HD bool rc_seg_end(RcEnc& e, RcSegSink& s, bool last)
{
    RcSegSinkChecked cs(s);
    u8 tail[8];
    const u32 nt = rc_enc_finish(e, cs, tail);
    const u32 at = 4u * (u32)cs.s.wcount;
    const u32 nw = last ? nt : nt - 4u;
    u8* ob = reinterpret_cast<u8*>(s.out);
    for(u32 k = 0; k < nw; ++k) {
        if(at + k >= s.lo && at + k < s.hi) {
            ob[at + k] = tail[k];
        }
    }
    s.wcount = cs.s.wcount;
    return at + nw == s.hi;
}

RC_HD bool rc_seg_end(RcEnc2& e, RcSegSink& s, bool last)
{
    RcEnc v = rc_enc2_view(e);
    return rc_seg_end(v, s, last);
}

// (big-endian 32 bits at p) += tail; a carry runs towards lower addresses through 0xFF bytes
// (the reference's buffer_ / count_ bookkeeping, cpprcoder.h:405-416, done after the fact).
// It never runs past the first coded byte: every partial sum is below the full one.
RC_HD void rc_seam_add(u8* p, u32 tail)
{
    const u32 have = ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3];
    const u32 sum = have + tail;
    p[0] = (u8)(sum >> 24);
    p[1] = (u8)(sum >> 16);
    p[2] = (u8)(sum >> 8);
    p[3] = (u8)sum;
    if(sum < have) {
        u8* q = p - 1;
        while(*q == 0xFFu) {
            *q = 0;
            --q;
        }
        *q = (u8)(*q + 1u);
    }
}

// The encoder's 32-bit low at the end of a segment that shifted `d` bytes out, from its own
// final low and the low at its start (what a restart point records).
RC_HD u32 rc_seam_low(u32 low_before, u32 own_low, u32 d)
{
    return own_low + (d >= 4u ? 0u : (low_before << (8u * d)));
}
