// sim_lane.cpp -- drives cpprcoder_b200/csrc/rc_lane.cuh on the CPU, one block at
// a time, so the lane arithmetic the kernels run (carry-save encoder, windowed
// decoder, count-tree model, magic division) can be checked against the oracle
// without a GPU.  TEST INFRASTRUCTURE: built and loaded by tests/test_sim_lane.py only.
#include "../../cpprcoder_b200/csrc/rc_lane.cuh"
#include "../../cpprcoder_b200/csrc/ans_lane.cuh"

#include <cstring>
#include <algorithm>
#include <vector>

namespace
{
struct VecSink {
    typedef VecSink Checked;
    std::vector<u8>* out;
    bool first = true;
    void settle(VecSink& s) const { s = *this; }
    bool tight(int) const { return false; }
    void push(u32 w)
    {
        if(first) {  // the encoder's placeholder word (rc_lane.cuh, Sink contract)
            first = false;
            return;
        }
        out->push_back((u8)(w >> 24));
        out->push_back((u8)(w >> 16));
        out->push_back((u8)(w >> 8));
        out->push_back((u8)w);
    }
};

struct ArrTab {
    u32 v[512];
    u32 ld(u32 i) const { return v[i]; }
    void st(u32 i, u32 x) { v[i] = x; }
};

// static count for blocks of at most 65536 bytes (the K1 rule)
void count64k(const u8* src, u32 n, u32 freq[256])
{
    memset(freq, 0, 256 * sizeof(u32));
    for(u32 i = 0; i < n; ++i) {
        freq[src[i]]++;
    }
    for(int s = 0; s < 256; ++s) {
        if(freq[s] >= 0x10000u) {
            freq[s] = 0x8000u;
        }
    }
}

struct WordReader {
    const u8* p;
    size_t n, pos;  // pos = byte index of the next aligned word
    u32 operator()() { return take(true); }
    u32 take(bool need)
    {
        if(!need) {
            return 0xDEADBEEFu;  // must be ignored by the caller
        }
        u32 w = 0;
        for(int k = 0; k < 4; ++k) {
            const size_t at = pos + k;
            w = (w << 8) | (at < n ? p[at] : 0u);
        }
        pos += 4;
        return w;
    }
};
}  // namespace

extern "C" {

// mode 0 static (n <= 65536), 1 adaptive.  force_exact: use the reference-shaped path.
long sim_encode(int mode, const u8* src, u32 n, u8* dst, size_t cap, int force_exact)
{
    std::vector<u8> out;
    out.push_back((u8)n);
    out.push_back((u8)(n >> 8));
    out.push_back((u8)(n >> 16));
    out.push_back((u8)(n >> 24));
    VecSink sink{&out};
    RcEnc e;
    u8 tail[8];
    if(mode == 0) {
        u32 freq[256], cum[257];
        count64k(src, n, freq);
        u32 run = 0;
        for(int s = 0; s < 256; ++s) {
            out.push_back((u8)freq[s]);
            out.push_back((u8)(freq[s] >> 8));
            cum[s] = run;
            run += freq[s];
        }
        cum[256] = run;
        const u32 total = run;
        const u32 magic = rc_magic(total);
        const bool pow2 = total && (total & (total - 1)) == 0;
        const u32 shT = pow2 ? 31u - rc_clz(total) : 0;
        rc_enc_init(e, RC_STATIC_RANGE0);
        if(!force_exact) {
            // groups of four steps, then one commit -- the shape of the kernel's hot loop
            u32 t = pow2 ? (e.range >> shT) : 0;
            for(u32 i = 0; i < n; i += 4) {
                RcCut cuts[4];
                for(u32 k = 0; k < 4; ++k) {
                    const bool active = i + k < n;
                    const u32 c = active ? src[i + k] : 0;
                    if(pow2) {
                        rc_enc_step_pow2<2>(e, t, shT, cum[c], cum[c + 1] - cum[c], cuts[k], active);
                    } else {
                        const u32 tt = active ? rc_div(e.range, total, magic) : 0;
                        rc_enc_step<3>(e, cum[c], cum[c + 1] - cum[c], tt, cuts[k], active);
                    }
                }
                rc_enc_commit(e, cuts, sink);
            }
        }
        if(force_exact || e.low == 0xFFFFFFFFu) {
            out.resize(RC_STATIC_HDR);
            rc_static_encode_exact(
                n, total, [&](u32 c) { return cum[c]; }, [&](u32 i) { return (u32)src[i]; },
                [&](u8 b) { out.push_back(b); });
        } else {
            const u32 nt = rc_enc_finish(e, sink, tail);
            out.insert(out.end(), tail, tail + nt);
        }
    } else {
        ArrTab tab;
        memset(tab.v, 0, sizeof tab.v);
        RcEnc2 e2;  // the multiplier form of the step, as k_enc_adaptive runs it
        e2.x = 0;
        e2.oc = 1u << 8;
        e2.range = RC_ADAPT_RANGE0;
        e2.pend = 0;
        e2.nff = 0;
        for(u32 i = 0; i < n; i += 4) {
            RcCut cuts[4];
            for(u32 k = 0; k < 4; ++k) {
                const bool active = i + k < n;
                u32 cum = 0, freq = 1, t = 0;
                if(active) {
                    rc_model_encode(tab, src[i + k], cum, freq);
                    const u32 d = 256u + i + k;
                    t = rc_div(e2.range, d, rc_magic(d));
                }
                rc_enc2_step<3>(e2, cum, freq, t, cuts[k], active);
            }
            rc_enc2_commit(e2, cuts, sink);
        }
        e = rc_enc2_view(e2);
        const u32 nt = rc_enc_finish(e, sink, tail);
        out.insert(out.end(), tail, tail + nt);
    }
    if(out.size() > cap) {
        return -1;
    }
    memcpy(dst, out.data(), out.size());
    return (long)out.size();
}

// Static encode with restart points: besides the payload, reports the
// encoder's (bytes shifted out so far, low, range) before symbol j * seg for j = 1 .. nseg-1,
// taken exactly where the kernel takes them -- at a commit boundary, from the words pushed so far
// and the bits still in the shift register.  rec[(j-1)*3 + {0,1,2}]; 0xFFFFFFFF when not reached.
long sim_encode_restart(const u8* src, u32 n, u8* dst, size_t cap, u32 seg, u32 nseg, u32* rec)
{
    struct CountSink {
        typedef CountSink Checked;
        std::vector<u8>* out;
        s32 wcount = -1;  // as SlotSink: the first push is the placeholder
        void settle(CountSink& s) const { s = *this; }
        bool tight(int) const { return false; }
        void push(u32 w)
        {
            if(wcount >= 0) {
                out->push_back((u8)(w >> 24));
                out->push_back((u8)(w >> 16));
                out->push_back((u8)(w >> 8));
                out->push_back((u8)w);
            }
            ++wcount;
        }
    };
    for(u32 i = 0; i < 3 * (nseg - 1); ++i) {
        rec[i] = 0xFFFFFFFFu;
    }
    std::vector<u8> out;
    for(int k = 0; k < 4; ++k) {
        out.push_back((u8)(n >> (8 * k)));
    }
    u32 freq[256], cum[257];
    count64k(src, n, freq);
    u32 run = 0;
    for(int s2 = 0; s2 < 256; ++s2) {
        out.push_back((u8)freq[s2]);
        out.push_back((u8)(freq[s2] >> 8));
        cum[s2] = run;
        run += freq[s2];
    }
    cum[256] = run;
    const u32 total = run, magic = rc_magic(total);
    const bool pow2 = total && (total & (total - 1)) == 0;
    const u32 shT = pow2 ? 31u - rc_clz(total) : 0;
    CountSink sink;
    sink.out = &out;
    RcEnc e;
    rc_enc_init(e, RC_STATIC_RANGE0);
    u32 t = pow2 ? (e.range >> shT) : 0;
    for(u32 i = 0; i < n; i += 4) {
        if(i && i % seg == 0 && i / seg < nseg) {
            u32* r = rec + 3 * (i / seg - 1);
            const u32 words = (u32)(sink.wcount + 1) + e.nff;      // words cut off the shift register so far
            r[0] = 4u * words + (u32)e.ocnt / 8u - 1u;              // bytes shifted out of low after the dummy byte
            r[1] = e.low;
            r[2] = pow2 ? (t << shT) : e.range;                     // any range with the same range >> shT serves
        }
        RcCut cuts[4];
        for(u32 k = 0; k < 4; ++k) {
            const bool active = i + k < n;
            const u32 c = active ? src[i + k] : 0;
            if(pow2) {
                rc_enc_step_pow2<2>(e, t, shT, cum[c], cum[c + 1] - cum[c], cuts[k], active);
            } else {
                const u32 tt = active ? rc_div(e.range, total, magic) : 0;
                rc_enc_step<3>(e, cum[c], cum[c + 1] - cum[c], tt, cuts[k], active);
            }
        }
        rc_enc_commit(e, cuts, sink);
    }
    if(e.low == 0xFFFFFFFFu) {
        return -2;  // the flush quirk: the kernel re-encodes such a block and drops its restart points
    }
    u8 tail[8];
    const u32 nt = rc_enc_finish(e, sink, tail);
    out.insert(out.end(), tail, tail + nt);
    if(out.size() > cap) {
        return -1;
    }
    memcpy(dst, out.data(), out.size());
    return (long)out.size();
}

// Static encode as the segmented kernels do it (rc_lane.cuh, "segmented static encode"): a range-only
// pass records, before every P-th symbol, the bytes shifted out so far and the range; every segment
// is then coded from low = 0 straight into its own bytes of the payload at dst + lead (any
// alignment; dst itself 4-byte aligned), LAST SEGMENT FIRST so that a segment that touched a
// neighbour's bytes would be found out; the seams are added afterwards.  rec as in
// sim_encode_restart (restart points every `seg` symbols, seg a multiple of P).  Returns the
// payload size, -1 no room, -2 flush quirk, -3 a segment did not end where the range pass said.
long sim_encode_segmented(const u8* src, u32 n, u8* dst, size_t cap, u32 lead, u32 P, u32 seg, u32 nrec, u32* rec)
{
    for(u32 i = 0; i < 3 * nrec; ++i) {
        rec[i] = 0xFFFFFFFFu;
    }
    u32 freq[256], cum[257];
    count64k(src, n, freq);
    u32 run = 0;
    for(int s2 = 0; s2 < 256; ++s2) {
        cum[s2] = run;
        run += freq[s2];
    }
    cum[256] = run;
    const u32 total = run, magic = rc_magic(total);
    const bool pow2 = total && (total & (total - 1)) == 0;
    const u32 shT = pow2 ? 31u - rc_clz(total) : 0;
    const u32 nseg = n ? (n + P - 1) / P : 1;
    // ---- the range-only pass
    std::vector<u32> S(nseg + 1), R(nseg + 1);
    {
        u32 range = RC_STATIC_RANGE0, t = pow2 ? (range >> shT) : 0, bits = 0;
        for(u32 i = 0; i < n; ++i) {
            // the general chain carries the range before its renormalisation (rc_range_step_div)
            const u32 shn = pow2 ? 0u : rc_norm_shift_flo(range);
            if(i % P == 0) {
                S[i / P] = (bits + shn) / 8;
                R[i / P] = pow2 ? (t << shT) : (range << shn);
            }
            const u32 f = cum[src[i] + 1] - cum[src[i]];
            if(pow2) {
                bits += rc_range_step_pow2<2>(t, shT, f);
            } else {
                // ... and must agree with the plain form of the link at every symbol
                u32 plain = range << shn;
                const u32 plain_sh = rc_range_step<3>(plain, f, rc_div(plain, total, magic));
                rc_range_step_div(range, bits, f, total, 0u - total, magic);
                if(range << rc_norm_shift_flo(range) != plain || rc_norm_shift_flo(range) != plain_sh) {
                    return -2;
                }
            }
        }
        if(n == 0) {
            S[0] = 0;
            R[0] = RC_STATIC_RANGE0;
        }
        S[nseg] = (bits + ((pow2 || n == 0) ? 0u : rc_norm_shift_flo(range))) / 8;
    }
    const size_t size = RC_STATIC_HDR + 5u + S[nseg];
    if(lead + size > cap) {
        return -1;
    }
    u8* pay = dst + lead;
    memset(pay, 0xAA, size);  // every byte must be written by somebody
    for(int k = 0; k < 4; ++k) {
        pay[k] = (u8)(n >> (8 * k));
    }
    for(int s2 = 0; s2 < 256; ++s2) {
        pay[4 + 2 * s2] = (u8)freq[s2];
        pay[5 + 2 * s2] = (u8)(freq[s2] >> 8);
    }
    u8* coded = pay + RC_STATIC_HDR;
    coded[0] = 0;  // the reference's initial buffer_ (cpprcoder.h:385): written with the header
    // ---- the segments, last first
    std::vector<u32> lows(nseg);
    for(u32 j = nseg; j-- > 0;) {
        const bool last = j + 1 == nseg;
        const u32 own = (last ? S[nseg] + 5u : S[j + 1] + 1u) - (S[j] + 1u);
        RcEnc2 e;  // the multiplier form of the step (k_enc_seg with B2RC_SEG_MULTIPLIER; the default funnel form is sim_encode's)
        RcSegSink sink;
        rc_seg_begin(e, sink, coded + S[j] + 1u, own, R[j]);
        u32 t = pow2 ? (e.range >> shT) : 0;
        const u32 hi = (j + 1) * P < n ? (j + 1) * P : n;
        for(u32 i = j * P; i < hi; i += 4) {
            RcCut cuts[4];
            for(u32 k = 0; k < 4; ++k) {
                const bool active = i + k < hi;
                const u32 c = active ? src[i + k] : 0;
                if(pow2) {
                    rc_enc2_step_pow2<2>(e, t, shT, cum[c], cum[c + 1] - cum[c], cuts[k], active);
                } else {
                    const u32 tt = active ? rc_div(e.range, total, magic) : 0;
                    rc_enc2_step<3>(e, cum[c], cum[c + 1] - cum[c], tt, cuts[k], active);
                }
            }
            rc_enc2_commit(e, cuts, sink);
        }
        if(!rc_seg_end(e, sink, last)) {
            return -3;
        }
        lows[j] = (u32)e.x;
    }
    // ---- the seams, in order, and the restart points
    u32 low = 0;
    for(u32 j = 0; j < nseg; ++j) {
        low = rc_seam_low(low, lows[j], S[j + 1] - S[j]);
        if(j + 1 < nseg) {
            rc_seam_add(coded + S[j + 1] + 1u, lows[j]);
            const u32 at = (j + 1) * P;
            if(seg && at % seg == 0 && at / seg - 1 < nrec) {
                u32* r = rec + 3 * (at / seg - 1);
                r[0] = S[j + 1];
                r[1] = low;
                r[2] = R[j + 1];
            }
        }
    }
    if(low == 0xFFFFFFFFu) {
        return -2;
    }
    return (long)size;
}

// Static decode of symbols [k, k + count) from a restart point (m, low, range) of sim_encode_restart.
long sim_decode_from(const u8* stream, size_t stream_len, u32 lead, u32 k, u32 m, u32 enc_low, u32 range, u32 count,
                     u8* dst)
{
    const u8* pay = stream + lead;
    u32 cum[257];
    u32 run = 0;
    for(int s2 = 0; s2 < 256; ++s2) {
        cum[s2] = run;
        run += (u32)pay[4 + 2 * s2] | ((u32)pay[5 + 2 * s2] << 8);
    }
    cum[256] = run;
    const u32 total = run, magic = rc_magic(total);
    struct CumTab {
        enum : u32 { UNIT = 1 };
        const u32* c;
        u32 at(u32 i) const { return c[i]; }
    } ctab{cum};
    u32 k1[8];
    for(int j = 0; j < 8; ++j) {
        k1[j] = cum[32 * j];
    }
    // the byte at offset m of the coded stream plays the part of the dummy first byte
    const size_t at = lead + RC_STATIC_HDR + m;
    WordReader rd{stream, stream_len, at & ~(size_t)3};
    RcDec d;
    rc_dec_init(d, range, (u32)(at & 3), rd);
    d.low -= enc_low;
    (void)k;
    const bool pow2 = (total & (total - 1)) == 0;
    const u32 shT = pow2 ? 31u - rc_clz(total) : 0;
    u32 t = pow2 ? (d.range >> shT) : 0;
    for(u32 i = 0; i < count; ++i) {
        if(!pow2) {
            t = rc_div(d.range, total, magic);
        }
        // the segmented kernel's search (four levels of four) and, for totals up to 2^16, its
        // every-second-symbol top-up
        u32 sym, c0, fr, s2, c2, f2;
        const u32 k0[4] = {k1[0], k1[2], k1[4], k1[6]};
        rc_static_find4(ctab, k0, t, d.low, sym, c0, fr);
        rc_static_find(ctab, k1, t, d.low, s2, c2, f2);
        if(sym != s2 || c0 != c2 || fr != f2) {
            return -2;
        }
        dst[i] = (u8)sym;
        if(!pow2) {
            rc_dec_advance(d, c0, fr, t, rd);
        } else if(total <= 65536u && (count & 1u) == 0) {
            rc_dec_advance_pow2_pair(d, t, shT, c0, fr, rd, (i & 1u) != 0);
        } else {
            rc_dec_advance_pow2<3>(d, t, shT, c0, fr, rd);
        }
    }
    return (long)count;
}

// `lead` = how many bytes precede the payload inside an (aligned) stream, to exercise
// the misaligned-start path of rc_dec_init.  src points at the aligned stream start.
long sim_decode(int mode, const u8* stream, size_t stream_len, u32 lead, u8* dst, size_t cap)
{
    const u8* pay = stream + lead;
    const u32 want = (u32)pay[0] | ((u32)pay[1] << 8) | ((u32)pay[2] << 16) | ((u32)pay[3] << 24);
    if(want > cap) {
        return -1;
    }
    if(want == 0) {
        return 0;
    }
    const u32 hdr = mode == 0 ? RC_STATIC_HDR : RC_ADAPT_HDR;
    const size_t coded = lead + hdr;  // byte index of coded byte 0 within the stream
    WordReader rd{stream, stream_len, coded & ~(size_t)3};
    RcDec d;
    if(mode == 0) {
        u32 cum[257];
        u32 run = 0;
        for(int s = 0; s < 256; ++s) {
            cum[s] = run;
            run += (u32)pay[4 + 2 * s] | ((u32)pay[5 + 2 * s] << 8);
        }
        cum[256] = run;
        const u32 total = run;
        if(total == 0) {
            return -1;
        }
        const u32 magic = rc_magic(total);
        struct CumTab {
            enum : u32 { UNIT = 1 };
            const u32* c;
            u32 at(u32 i) const { return c[i]; }
        } ctab{cum};
        u32 k1[8];
        for(int j = 0; j < 8; ++j) {
            k1[j] = cum[32 * j];
        }
        rc_dec_init(d, RC_STATIC_RANGE0, (u32)(coded & 3), rd);
        // the kernel's three chains: general divide, power of two with 2 or 3 renormalisation rounds
        const bool pow2 = (total & (total - 1)) == 0;
        const u32 shT = pow2 ? 31u - rc_clz(total) : 0;
        u32 t = pow2 ? (d.range >> shT) : 0;
        for(u32 i = 0; i < want; ++i) {
            if(!pow2) {
                t = rc_div(d.range, total, magic);
            }
            u32 sym, c0, fr;
            rc_static_find(ctab, k1, t, d.low, sym, c0, fr);
            if(c0 != cum[sym] || fr != cum[sym + 1] - cum[sym]) {
                return -2;
            }
            dst[i] = (u8)sym;
            if(!pow2) {
                rc_dec_advance(d, c0, fr, t, rd);
            } else if(total <= 65536u && (i & 1)) {
                rc_dec_advance_pow2<2>(d, t, shT, c0, fr, rd);
            } else {
                rc_dec_advance_pow2<3>(d, t, shT, c0, fr, rd);
            }
        }
    } else {
        ArrTab tab;
        memset(tab.v, 0, sizeof tab.v);
        rc_dec_init(d, RC_ADAPT_RANGE0, (u32)(coded & 3), rd);
        ArrTab bare;  // the leafless tree of k_dec_adaptive_seg, walked side by side: same answers
        memset(bare.v, 0, sizeof bare.v);
        for(u32 i = 0; i < want; ++i) {
            const u32 dd = 256u + i;
            const u32 t = rc_div(d.range, dd, rc_magic(dd));
            u32 sym, cum, freq, s2, c2, f2;
            rc_model_decode_leafless(bare, d.low, t, i, s2, c2, f2);
            rc_model_decode(tab, d.low, t, sym, cum, freq);
            if(s2 != sym || c2 != cum || f2 != freq) {
                return -3;
            }
            dst[i] = (u8)sym;
            rc_dec_advance(d, cum, freq, t, rd);
        }
    }
    return (long)want;
}

// ---- byte-wise rANS (ans_lane.cuh).  `cum` = the normalised 257-entry cumulative table the
// model step produced (tests take it from the oracle; the kernels from k_ans_model).
long sim_ans_byte_encode(const u8* src, u32 n, const u32* cum, u8* dst, size_t cap)
{
    struct BackSink {
        std::vector<u8> rev;  // bytes in emission order = descending address
        void word(u32 w, bool on)
        {
            if(on) {
                rev.push_back((u8)(w >> 24));
                rev.push_back((u8)(w >> 16));
                rev.push_back((u8)(w >> 8));
                rev.push_back((u8)w);
            }
        }
        void byte(u8 b) { rev.push_back(b); }
    } sink;
    u32 magic[256];
    for(int s = 0; s < 256; ++s) {
        magic[s] = rc_magic(cum[s + 1] - cum[s]);
    }
    u32 x = ANS_BYTE_LOW;
    AnsByteAcc acc;
    ans_acc_init(acc);
    // last symbol first, two puts per commit -- the shape of the kernel's loop
    const u32 n2 = (n + 1u) & ~1u;
    for(u32 i = n2; i > 0; i -= 2) {
        for(u32 k = 0; k < 2; ++k) {
            const u32 p = i - 1u - k;
            const bool active = p < n;
            const u32 c = active ? src[p] : 0;
            AnsPut put;
            ans_byte_put(x, cum[c], cum[c + 1] - cum[c], magic[c], put, active);
            ans_acc_push(acc, put);
        }
        ans_acc_commit(acc, sink);
    }
    ans_acc_finish(acc, x, sink);
    const size_t total = ANS_HDR_BYTES + sink.rev.size();
    if(total > cap) {
        return -1;
    }
    memcpy(dst, &n, 4);
    memcpy(dst + 4, cum, 257 * 4);
    for(size_t i = 0; i < sink.rev.size(); ++i) {
        dst[ANS_HDR_BYTES + i] = sink.rev[sink.rev.size() - 1 - i];
    }
    return (long)total;
}

long sim_ans_byte_decode(const u8* stream, size_t stream_len, u32 lead, u8* dst, size_t cap)
{
    const u8* pay = stream + lead;
    u32 want;
    memcpy(&want, pay, 4);
    if(want > cap) {
        return -1;
    }
    u32 cum[257];
    memcpy(cum, pay + 4, sizeof cum);
    struct CumTab {
        enum : u32 { UNIT = 1 };
        const u32* c;
        u32 at(u32 i) const { return c[i]; }
    } ctab{cum};
    u32 k1[8];
    for(int j = 0; j < 8; ++j) {
        k1[j] = cum[32 * j];
    }
    const size_t coded = lead + ANS_HDR_BYTES;
    WordReader rd{stream, stream_len, coded & ~(size_t)3};
    RcDec d;
    u32 x = ans_byte_dec_init(d, (u32)(coded & 3), rd);
    for(u32 i = 0; i < want; ++i) {
        const u32 slot = x & ((1u << ANS_BYTE_SCALE_BITS) - 1u);
        u32 sym, start, f;
        ans_find(ctab, k1, slot, sym, start, f);
        if(start != cum[sym] || f != cum[sym + 1] - cum[sym] || slot < start || slot - start >= f) {
            return -2;
        }
        dst[i] = (u8)sym;
        if(i & 1) {
            ans_byte_advance<true>(d, x, slot, start, f, rd);
        } else {
            ans_byte_advance<false>(d, x, slot, start, f, rd);
        }
        if(x < ANS_BYTE_LOW) {
            return -3;
        }
    }
    return (long)want;
}

// The range chain for total == 65536 as the range-pass kernels compute it (b2rc_encseg.cuh), for one frequency and
// EVERY t in [2^8, 2^16), against the plain form of the link (rc_range_step_pow2): the next t as the minimum of three
// wrapped candidates (range_step16), the same with the chain carrying t - 256 and the addends 256 f - K prepared
// beside it (ranges2_chain), and the shift read off r - 256 afterwards (ranges2_shifts).  Returns the mismatches.
u64 sim_check_range16(u32 f)
{
    u64 bad = 0;
    for(u32 t = 256u; t < 65536u; ++t) {
        u32 want = t;
        const u32 want_sh = rc_range_step_pow2<2>(want, 16u, f);
        const u32 r = f * t;
        const u32 a = (r - 0x01000000u) >> 16, b = (r - 0x00010000u) >> 8, c = r - 0x00000100u;
        const u32 t1 = std::min(std::min(a, b), c) + 256u;
        const u32 m = t - 256u, c3 = f * 256u - 256u;
        const u32 r1 = f * m + (c3 - 0x00FFFF00u), r2 = f * m + (c3 - 0x0000FF00u), r3 = f * m + c3;
        const u32 m1 = std::min(std::min(r1 >> 16, r2 >> 8), r3);
        const u32 sh2 = rc_clz(r3 + 256u) & 24u;
        bad += (t1 != want ? 1u : 0u) + (m1 + 256u != want ? 1u : 0u) + (sh2 != want_sh ? 1u : 0u) + (r3 != c ? 1u : 0u);
    }
    return bad;
}

// exhaustive-ish check of the magic division; returns the number of mismatches
u64 sim_check_div(u32 d, u32 x0, u32 step, u32 count)
{
    const u32 m = rc_magic(d);
    u64 bad = 0;
    u32 x = x0;
    for(u32 i = 0; i < count; ++i, x += step) {
        bad += (rc_div(x, d, m) != x / d) ? 1u : 0u;
    }
    return bad;
}
}
