// b2rc_ans.cuh -- the sm_100a kernels of the block rANS coder (cppans::rANS of the
// reference's cppans.h), on the same block framework as the
// range coder: same container, same scan, one payload per block that is byte for byte
// what the reference emits for that block.
//
//   A1  k_ans_model<BITS>   count + cumulative + normalize     (cppans.h:102-177)
//   A2  k_ans_enc_word      rANS::encode_simd                   (cppans.h:567-607)
//   A3  k_ans_dec_word      rANS::decode_simd                   (cppans.h:609-649)
//   A4  k_compact_split     header + tail-aligned coded bytes -> contiguous payload
//   A5  k_ans_enc_byte      rANS::encode                        (cppans.h:497-530)
//   A6  k_ans_dec_byte      rANS::decode                        (cppans.h:532-564)
//
// Mapping of the word variant: the reference interleaves EIGHT rANS states per stream so
// that SSE can step four at a time; here the eight states of a block are eight LANES, a
// warp carries four blocks, and the only communication is the one the format demands --
// who emits / refills in this round, in lane order -- which is a ballot and a popcount.
// That is eight times the parallelism per block the range coder offers, and no carries.
//
// The byte variant has ONE state per stream: a block is one serial chain and gets one lane,
// 32 blocks per warp, tables interleaved by lane -- the mapping of the range coder kernels
// (b2rc_kernels.cuh), whose input staging, stream queue and output tiles it reuses; its
// arithmetic is in ans_lane.cuh and is also run on the CPU by tests/sim.
//
// Slot layout while encoding: the model kernel leaves the 1032-byte header (u32 size,
// u32 cum[257]) at the START of the block's slot; the coder writes its words from the END
// of the slot downwards, as the reference writes from the end of dst (cppans.h:591);
// k_compact_split joins the two pieces.
#pragma once
#include "ans_lane.cuh"
#include "b2rc_kernels.cuh"

namespace b2rc
{
constexpr u32 ANS_HDR = 1032u;          // 258 x u32 (cppans.h:598-604)
constexpr u32 ANS_WORD_BITS = 12u;      // rANS::WordScaleBits (cppans.h:31)
constexpr u32 ANS_BYTE_BITS = 14u;      // rANS::ProbBits      (cppans.h:27)
constexpr u32 ANS_WORD_LOW = 1u << 16;  // rANS::WordLowBounds (cppans.h:30)

__device__ __forceinline__ void lds64(u32 a, u32& lo, u32& hi)
{
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a));
}
// little-endian u32 / u16 at a 2-byte aligned address
__device__ __forceinline__ u32 ld32_a2(const u8* p)
{
    const u16* q = reinterpret_cast<const u16*>(p);
    return (u32)__ldg(q) | ((u32)__ldg(q + 1) << 16);
}

// ======================================================================== A1 ==
// One warp per block: exact byte counts (as K1, without its 16-bit scaling), then the
// reference's normalize -- scale the cumulative table to 2^BITS, and wherever a symbol
// that occurs was squeezed to nothing, take one unit from the narrowest slice wider than
// one (first such slice on ties), sliding the boundaries in between (cppans.h:138-177).
// The repair loop runs in the reference's order; each repair is a warp-wide arg-min.
template <int BITS>
__global__ void __launch_bounds__(HIST_WARPS * 32) k_ans_model(const u8* src, u64 n, u32 block, u64 nblocks, u8* slots,
                                                               u64 slot_stride)
{
    __shared__ u32 bins[HIST_WARPS][256];
    __shared__ u32 cums[HIST_WARPS][264];
    const u32 lane = lane_id();
    const u32 warp = threadIdx.x >> 5;
    u32* h = bins[warp];
    u32* sc = cums[warp];
    for(u64 b = (u64)blockIdx.x * HIST_WARPS + warp; b < nblocks; b += (u64)gridDim.x * HIST_WARPS) {
        const u64 lo = b * (u64)block;
        const u32 len = (u32)((n - lo < block) ? (n - lo) : block);
        hist_block(h, src + lo, len, lane);
        u32 f[8], sum = 0;
#pragma unroll
        for(int k = 0; k < 8; ++k) {
            f[k] = h[8u * lane + k];
            sum += f[k];
        }
        u32 inc = sum;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(FULL, inc, d);
            if(lane >= (u32)d) {
                inc += t;
            }
        }
        u32 run = inc - sum;
        if(lane == 0) {
            sc[0] = 0;
        }
#pragma unroll
        for(int k = 0; k < 8; ++k) {
            run += f[k];
            sc[8u * lane + k + 1u] = (u32)((((u64)run) << BITS) / len);
        }
        __syncwarp();
        for(u32 i = 0; i < 256u; ++i) {
            if(h[i] == 0u || sc[i + 1u] != sc[i]) {  // same words for every lane: uniform
                continue;
            }
            u32 key = FULL;
            u32 prev = sc[8u * lane];
#pragma unroll
            for(int k = 0; k < 8; ++k) {
                const u32 nx = sc[8u * lane + k + 1u];
                const u32 w = nx - prev;
                prev = nx;
                const u32 cand = (w << 8) | (8u * lane + k);
                if(w > 1u && cand < key) {
                    key = cand;
                }
            }
            key = __reduce_min_sync(FULL, key);
            if(key == FULL) {
                continue;
            }
            const u32 donor = key & 255u;
            __syncwarp();
            if(donor < i) {
                for(u32 j = donor + 1u + lane; j <= i; j += 32u) {
                    sc[j] -= 1u;
                }
            } else {
                for(u32 j = i + 1u + lane; j <= donor; j += 32u) {
                    sc[j] += 1u;
                }
            }
            __syncwarp();
        }
        u32* hdr = reinterpret_cast<u32*>(slots + b * slot_stride);
        for(u32 w = lane; w < 258u; w += 32u) {
            hdr[w] = w == 0u ? len : sc[w - 1u];
        }
        __syncwarp();
    }
}

// ======================================================================== A2 ==
// Word-variant encode: lane j of an 8-lane group is state j of its block and takes the
// symbols at positions p with p & 7 == j, last round first (cppans.h:591-594).  A state
// emits its low 16 bits when it reaches freq << 20 -- a u32 product in the reference
// (wordEncPut, cppans.h:357), so a symbol that owns the whole scale emits every time --
// and within a round the states emit in descending lane order at descending addresses.
//
// Everything that does not depend on the state runs ahead of it, one quad (4 rounds, 32
// input bytes per block) per loop trip:
//   A  the lane's four symbols of a quad, byte loads through L1, five quads ahead;
//   B  {start, freq, reciprocal of freq} of each symbol in one 8-byte load from the
//      block's table in shared memory, one quad ahead;
//   D  the state chain itself: compare, ballot, store, divide, update.
// (The reciprocals first lived in one device-wide table indexed by freq and read through
// L1: 13 sectors per request, and the L1 tag stage became the bottleneck.)
__device__ __forceinline__ void st_u16(u8* p, u32 v)
{
    asm volatile("st.global.u16 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool RAGGED>
__device__ __forceinline__ void ans_enc_word_loop(const u8* __restrict__ sp, u32 len, s32 K4, u32 tb, u8* slot, u32& x,
                                                  u32& w, u32 lane)
{
    const u32 j = lane & 7u, gbase = lane & 24u;
    const u32 above = (0xFEu << j) & 0xFFu;          // the states of my block that emit before me
    const u32 gmask = 0xFFu << gbase, amask = above << gbase;
    const u8* __restrict__ pl = sp + j;              // my state's symbols: pl[8 * round]

    // A: the four symbols of quad q for this lane (one byte each; the eight lanes of a group
    // share a sector, the four rounds of a quad share it too)
    auto fetch = [&](s32 q, u32 (&sy)[4]) {
        const u8* __restrict__ pq = pl + 32 * (s64)q;
#pragma unroll
        for(int rr = 0; rr < 4; ++rr) {
            const u32 pos = 32u * (u32)q + 8u * (u32)rr;
            sy[rr] = (q >= 0 && (!RAGGED || pos + j < len)) ? (u32)__ldg(pq + 8 * rr) : 0u;
        }
    };
    // B: table entry of a symbol: {start << 16 | freq, reciprocal of freq}, one 8-byte load
    auto lookup = [&](const u32 (&sy)[4], u32 (&e)[4], u32 (&m)[4]) {
#pragma unroll
        for(int rr = 0; rr < 4; ++rr) {
            lds64(tb + 8u * sy[rr], e[rr], m[rr]);
        }
    };

    // Quad q keeps its symbols in S[q & 3] and its table entries in e/m[q & 1]; the trip
    // count is a multiple of four quads and the body is unrolled four times, so every
    // index below is a compile-time constant and nothing is moved between registers (a
    // register-to-register rotation would wait for the newest load every trip).
    u32 e[2][4], m[2][4], S[4][4];
    fetch(K4 - 1, S[3]);
    fetch(K4 - 2, S[2]);
    fetch(K4 - 3, S[1]);
    fetch(K4 - 4, S[0]);
    lookup(S[3], e[1], m[1]);
    fetch(K4 - 5, S[3]);
    for(s32 k = K4 - 1; k >= 0; k -= 4) {
#pragma unroll
        for(int u = 0; u < 4; ++u) {
            const s32 q = k - u;                      // q & 3 == 3 - u
            const int par = (3 - u) & 1, nxt = (2 - u) & 3;
            lookup(S[nxt], e[par ^ 1], m[par ^ 1]);   // B: quad q-1
            fetch(q - 5, S[nxt]);                     // A: quad q-5
#pragma unroll
            for(int rr = 3; rr >= 0; --rr) {          // D: quad q, last round first
                const u32 f = e[par][rr] & 0xFFFFu, start = e[par][rr] >> 16;
                const bool act = !RAGGED || (32u * (u32)q + 8u * (u32)rr + j) < len;
                const bool emit = act && ((f << 20) <= x);
                const u32 bal = __ballot_sync(FULL, emit);
                const u32 at = w - 2u * __popc(bal & amask);
                const u32 xl = x & 0xFFFFu;           // its own register: the store must not hold up the chain
                if(emit) {
                    st_u16(slot + at - 2, xl);
                }
                x = emit ? (x >> 16) : x;
                w -= 2u * __popc(bal & gmask);
                u32 quo = rc_umulhi(x, m[par][rr]);
                const u32 rem = x - quo * f;
                quo += rem >= f ? 1u : 0u;
                const u32 xn = x + start + quo * ((1u << ANS_WORD_BITS) - f);  // (quo << 12) + x % f + start, cppans.h:363
                x = act ? xn : x;
            }
        }
    }
}

__global__ void __launch_bounds__(32) k_ans_enc_word(EncArgs a)
{
    __shared__ uint2 tab[4][256];
    const u32 lane = lane_id(), grp = lane >> 3, j = lane & 7u;
    const u64 b = (u64)blockIdx.x * 4u + grp;
    const bool live = b < a.nblocks;
    const u64 lo = live ? b * (u64)a.block : 0ull;
    const u32 len = live ? (u32)((a.n - lo < a.block) ? (a.n - lo) : a.block) : 0u;
    u8* slot = a.slots + (live ? b : 0ull) * a.slot_stride;
    if(live) {
        const u32* hdr = reinterpret_cast<const u32*>(slot);
        for(u32 s = j; s < 256u; s += 8u) {
            const u32 c0 = hdr[1u + s], c1 = hdr[2u + s];
            tab[grp][s] = make_uint2((c0 << 16) | (c1 - c0), rc_magic(c1 - c0));
        }
    }
    __syncwarp();
    const u32 tb = smem_addr(tab[grp]);
    const u8* __restrict__ sp = a.src + lo;
    u32 x = ANS_WORD_LOW;        // wordEncInit (cppans.h:336-339)
    u32 w = (u32)a.slot_stride;  // write cursor, bytes from the slot start
    const s32 K4 = (s32)(((__reduce_max_sync(FULL, len) + 127u) >> 7) << 2);  // quads, a multiple of four
    if(__all_sync(FULL, live && len == 32u * (u32)K4)) {
        ans_enc_word_loop<false>(sp, len, K4, tb, slot, x, w, lane);
    } else {
        ans_enc_word_loop<true>(sp, len, K4, tb, slot, x, w, lane);
    }
    if(live) {
        // wordEncFlush, states 7 .. 0 at descending addresses (cppans.h:595-597)
        const u32 at = w - 32u + 4u * j;
        st_u16(slot + at, x);
        st_u16(slot + at + 2u, x >> 16);
        if(j == 0u) {
            a.sizes[b] = ANS_HDR + ((u32)a.slot_stride - (w - 32u));
        }
    }
}

// ======================================================================== A3 ==
// Word-variant decode.  Two tables per block in shared memory: slot -> symbol (4096 x u8,
// initSymbols' slot2symbol_, cppans.h:342-351) and symbol -> (start, freq); freq and bias
// of a slot follow from its symbol, so the reference's 16 KiB WordSlot array is not kept
// and 10 warps fit an SM instead of 3.  After each round of eight symbols the states below
// 2^16 take the next u16s of the stream in lane order (simdDecRenorm, cppans.h:443-488).
//
// The stream reaches the states through a 512-byte ring per block in shared memory, filled
// by cp.async a quarter (128 bytes, 16 per lane) at a time: when the read cursor enters
// quarter q the group requests quarter q+3 and waits for everything older, so q, q+1 and
// q+2 are always complete and the request is ~32 rounds ahead of its use.  No register
// ever waits for global memory inside the loop.
constexpr u32 ANS_RING = 512u, ANS_QUARTER = 128u;
constexpr u32 ANS_DEC_WARPS = 2u;  // per CTA: halves the per-CTA shared-memory reserve per warp
constexpr u32 ANS_DEC_WARP_SMEM = 4u * (4096u + 1024u + ANS_RING);
constexpr u32 ANS_DEC_WORD_SMEM = ANS_DEC_WARPS * ANS_DEC_WARP_SMEM;

struct AnsRing {
    const u8* abase;  // 16-byte aligned address at or below the block's payload
    u32 lim;          // readable bytes from abase to the end of the whole payload area
    u32 ring;         // shared address of the block's ring
    u32 j;

    // quarter `qi` (bytes [128 qi, 128 qi + 128) from abase) -> ring; 16 bytes per lane
    __device__ __forceinline__ void request(u32 qi) const
    {
        const u32 at = qi * ANS_QUARTER + 16u * j;
        const u32 bytes = at < lim ? (lim - at < 16u ? lim - at : 16u) : 0u;
        cp_async16(ring + (at & (ANS_RING - 1u)), abase + (bytes ? at : 0u), bytes);
        cp_async_commit();
    }
};

template <bool RAGGED>
__device__ __forceinline__ void ans_dec_word_loop(const AnsRing& rg, u32 full, u32 rmax4, u32 sb, u32 tb, u8* out, u32& x,
                                                  u32& rp, u32 lane)
{
    const u32 j = lane & 7u, gbase = lane & 24u;
    const u32 below = ((1u << j) - 1u) << gbase, gmask = 0xFFu << gbase;
    // the block's next eight u16s, one per lane, read from the ring as soon as the cursor is
    // known; a refilling lane takes its word from the lane that holds it (8-lane shuffle)
    auto window = [&](u32 pos) -> u32 { return lds16v(rg.ring + ((pos + 2u * j) & (ANS_RING - 1u))); };
    u32 wnext = window(rp);
    for(u32 r = 0; r < rmax4; r += 4u) {
        const u32 rp0 = rp;
#pragma unroll
        for(u32 t = 0; t < 4u; ++t) {
            const bool act = !RAGGED || r + t < full;
            const u32 slt = x & ((1u << ANS_WORD_BITS) - 1u);
            const u32 s = lds8(sb + slt);
            const u32 e = lds32(tb + 4u * s);
            u32 xn = (e >> 16) * (x >> ANS_WORD_BITS) + slt - (e & 0xFFFFu);  // simdDecSym (cppans.h:412-440)
            if(act) {
                out[8u * (r + t)] = (u8)s;
            }
            const bool need = act && xn < ANS_WORD_LOW;
            const u32 bal = __ballot_sync(FULL, need);
            const u32 wv = __shfl_sync(FULL, wnext, __popc(bal & below), 8);
            rp += 2u * __popc(bal & gmask);
            wnext = window(rp);
            xn = need ? ((xn << 16) | wv) : xn;
            x = act ? xn : x;
        }
        // entered a new quarter (at most 64 bytes ago): request the one three ahead and wait
        // for the older ones; the rounds above never read past quarter q+2
        const bool cross = ((rp ^ rp0) & ANS_QUARTER) != 0u;
        if(__any_sync(FULL, cross)) {
            if(cross) {
                rg.request((rp >> 7) + 3u);
            }
            cp_async_wait<1>();
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(32 * ANS_DEC_WARPS) k_ans_dec_word(DecArgs a)
{
    extern __shared__ __align__(16) u8 ans_sm[];
    const u32 lane = lane_id(), grp = lane >> 3, j = lane & 7u, warp = threadIdx.x >> 5;
    const u64 b = ((u64)blockIdx.x * ANS_DEC_WARPS + warp) * 4u + grp;
    const bool has = b < a.nblocks;
    const u64 lo = has ? b * (u64)a.block : 0ull;
    const u32 n_b = has ? (u32)((a.n - lo < a.block) ? (a.n - lo) : a.block) : 0u;
    u8* wsm = ans_sm + warp * ANS_DEC_WARP_SMEM;
    u8* s2s = wsm + grp * 4096u;
    u32* st = reinterpret_cast<u32*>(wsm + 4u * 4096u + grp * 1024u);
    const u8* pay = a.payload;
    u32 paylen = 0;
    bool ok = false;
    if(has) {
        const u64 o0 = a.offsets[b], o1 = a.offsets[b + 1];
        if(o0 <= o1 && o1 <= a.payload_len && ((o0 | o1) & 1u) == 0u && o1 - o0 >= ANS_HDR + 32u &&
           o1 - o0 < 0xFFFFFFF0ull) {
            pay = a.payload + o0;
            paylen = (u32)(o1 - o0);
            ok = ld32_a2(pay) == n_b;  // the container, not the payload, says how long block b is
        }
    }
    if(ok) {
        for(u32 s = j; s < 256u; s += 8u) {
            const u32 c0 = ld32_a2(pay + 4u + 4u * s), c1 = ld32_a2(pay + 8u + 4u * s);
            const bool sane = c0 <= c1 && c1 <= (1u << ANS_WORD_BITS) && (s != 0u || c0 == 0u) &&
                              (s != 255u || c1 == (1u << ANS_WORD_BITS));
            ok = ok && sane;
            st[s] = c0 | ((c1 - c0) << 16);
            if(sane) {
                for(u32 k = c0; k < c1; ++k) {
                    s2s[k] = (u8)s;
                }
            }
        }
    }
    const u32 gshift = grp * 8u;
    ok = ((__ballot_sync(FULL, ok) >> gshift) & 0xFFu) == 0xFFu;
    if(has && !ok && j == 0u) {
        atomicOr(a.err, ERR_CORRUPT);
    }
    const u32 want = ok ? n_b : 0u;
    const u32 full = want >> 3;
    const u32 sb = smem_addr(s2s), tb = smem_addr(st);
    u32 x = ok ? ld32_a2(pay + ANS_HDR + 4u * j) : 0u;  // simdDecInit (cppans.h:405-409)
    // the ring mirrors memory from a 16-byte aligned base; rp counts from that base
    AnsRing rg;
    const u32 mis = (u32)((uintptr_t)pay & 15u);
    rg.abase = pay - mis;
    const u64 room = (u64)((a.payload + a.payload_len) - rg.abase);
    rg.lim = ok ? (u32)(room < 0xFFFFFFF0ull ? room : 0xFFFFFFF0ull) : 0u;
    rg.ring = smem_addr(wsm + 4u * (4096u + 1024u) + grp * ANS_RING);
    rg.j = j;
    u32 rp = mis + ANS_HDR + 32u;
    const u32 end = mis + paylen;
    rg.request((rp >> 7) + 0u);
    rg.request((rp >> 7) + 1u);
    rg.request((rp >> 7) + 2u);
    rg.request((rp >> 7) + 3u);
    cp_async_wait<0>();
    __syncwarp();
    u8* out = a.dst + lo + j;
    const u32 rmax4 = (__reduce_max_sync(FULL, full) + 3u) & ~3u;
    if(__all_sync(FULL, full == rmax4)) {
        ans_dec_word_loop<false>(rg, full, rmax4, sb, tb, out, x, rp, lane);
    } else {
        ans_dec_word_loop<true>(rg, full, rmax4, sb, tb, out, x, rp, lane);
    }
    // the last (size & 7) symbols: one more symbol from states 0.. without a refill (cppans.h:643-647)
    if(8u * full + j < want) {
        out[8u * full] = (u8)lds8(sb + (x & ((1u << ANS_WORD_BITS) - 1u)));
    }
    if(ok && rp > end) {
        atomicOr(a.err, ERR_CORRUPT);  // the coder ran past this block's payload
    }
}

// ======================================================================== A4 ==
// Copies n bytes with a whole CTA; any alignment on either side.  Destination-aligned
// 4-byte words are assembled from the two aligned source words they straddle.
__device__ __forceinline__ void cta_copy(u8* d, const u8* s, u32 n)
{
    u32 head = (u32)((4u - ((uintptr_t)d & 3u)) & 3u);
    if(head > n) {
        head = n;
    }
    if(threadIdx.x < head) {
        d[threadIdx.x] = s[threadIdx.x];
    }
    const u32 nwords = (n - head) >> 2;
    u32* dw = reinterpret_cast<u32*>(d + head);
    const u8* s0 = s + head;
    const u32 mis = (u32)((uintptr_t)s0 & 3u);
    const u32* sw = reinterpret_cast<const u32*>(s0 - mis);
    const u32 sh = mis * 8u;
    for(u32 k = threadIdx.x; k < nwords; k += blockDim.x) {
        const u32 a0 = __ldg(sw + k);
        const u32 a1 = sh ? __ldg(sw + k + 1) : 0u;
        dw[k] = __funnelshift_r(a0, a1, sh);
    }
    const u32 done = head + 4u * nwords;
    if(done + threadIdx.x < n) {
        d[done + threadIdx.x] = s[done + threadIdx.x];
    }
}

// payload b = [head_bytes from the slot start][sizes[b] - head_bytes from the slot end]
__global__ void __launch_bounds__(COMPACT_THREADS) k_compact_split(const u8* slots, u64 slot_stride, const u32* sizes,
                                                                   const u64* offsets, u64 nblocks, u8* payload,
                                                                   u64 payload_cap, int* err, u32 head_bytes)
{
    for(u64 b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const u8* s = slots + b * slot_stride;
        const u32 len = sizes[b];
        const u64 off = offsets[b];
        if(off + len > payload_cap || len < head_bytes) {
            if(threadIdx.x == 0) {
                atomicOr(err, len < head_bytes ? ERR_SLOT_OVERFLOW : ERR_DST_SMALL);
            }
            continue;
        }
        cta_copy(payload + off, s, head_bytes);
        cta_copy(payload + off + head_bytes, s + slot_stride - (len - head_bytes), len - head_bytes);
    }
}

// ======================================================================== A5 ==
// Byte-variant encode, one block per lane.  Per warp in shared memory: the 257 cumulative
// counts as u16 (<= 2^14) and the 256 reciprocals as u32, both [entry][lane], and two input
// tiles staged by cp.async.  The block is walked from its last tile to its first and each
// tile from its last symbol to its first (cppans.h:516-519).  Emitted bytes collect in a
// 64-bit register and leave as aligned 32-bit words, downwards from the end of the slot.
constexpr u32 ANS_ENC_BYTE_CUM = 257u * 64u;                 // u16 x 32 lanes per entry
constexpr u32 ANS_ENC_BYTE_MAGIC = 256u * 128u;
constexpr u32 ANS_ENC_BYTE_SMEM = ANS_ENC_BYTE_CUM + ANS_ENC_BYTE_MAGIC + 2u * TILE_BYTES;

struct AnsSlotOut {
    u8* slot;
    u32 w;  // write cursor: bytes from the slot start; everything at and above it is written
    __device__ __forceinline__ void word(u32 v, bool on)
    {
        w -= on ? 4u : 0u;
        if(on) {
            *reinterpret_cast<u32*>(slot + w) = v;
        }
    }
    __device__ __forceinline__ void byte(u8 v)
    {
        w -= 1u;
        slot[w] = v;
    }
};

template <bool RAGGED>
__device__ __forceinline__ void ans_enc_byte_tiles(const EncArgs& a, u32 cum_a, u32 mag_a, u32 tiles, u64 b0, u32 n_b,
                                                   u32 ntiles, u32& x, AnsByteAcc& acc, AnsSlotOut& out, u32 lane)
{
    auto entry = [&](u32 sym, u32& start, u32& f, u32& mg) {
        start = lds16(cum_a + sym * 64u);
        f = lds16(cum_a + sym * 64u + 64u) - start;
        mg = lds32(mag_a + sym * 128u);
    };
    // the last tile was staged (and committed) by the caller into buffer (ntiles - 1) & 1
#pragma unroll 1
    for(u32 tix = ntiles; tix-- > 0;) {
        if(tix > 0) {
            stage_tile(tiles + ((tix - 1) & 1u) * TILE_BYTES, a.src, a.n, b0, a.block, (tix - 1) * TILE, lane);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const u32 row = tiles + (tix & 1u) * TILE_BYTES + lane * ROW;
        // four symbols per trip, last word of the tile first; the table entries of the next
        // trip's symbols are fetched before this trip's are coded
        u32 word = lds32(row + 4u * (TILE / 4 - 1));
        u32 st[4], fr[4], mg[4];
#pragma unroll
        for(int k = 0; k < 4; ++k) {
            entry((word >> (8 * k)) & 0xFFu, st[k], fr[k], mg[k]);
        }
#pragma unroll 1
        for(int wi = TILE / 4 - 1; wi >= 0; --wi) {
            const u32 wnext = lds32(row + 4u * (u32)((wi - 1) & (TILE / 4 - 1)));
            u32 nst[4], nfr[4], nmg[4];
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                entry((wnext >> (8 * k)) & 0xFFu, nst[k], nfr[k], nmg[k]);
            }
#pragma unroll
            for(int k = 3; k >= 0; --k) {
                const bool active = !RAGGED || tix * TILE + wi * 4 + k < n_b;
                AnsPut put;
                ans_byte_put(x, st[k], fr[k], mg[k], put, active);
                ans_acc_push(acc, put);
                if((k & 1) == 0) {
                    ans_acc_commit(acc, out);
                }
            }
#pragma unroll
            for(int k = 0; k < 4; ++k) {
                st[k] = nst[k];
                fr[k] = nfr[k];
                mg[k] = nmg[k];
            }
        }
        if(a.restart && tix != 0u && (tix * TILE) % a.seg_syms == 0u && tix * TILE < n_b) {
            // everything from symbol tix * TILE on is coded: a decoder that has x and knows how many
            // bytes were emitted so far can start at that symbol (the bytes it has consumed by then
            // are exactly the ones emitted from here on)
            const u32 nrec = (a.block + a.seg_syms - 1u) / a.seg_syms - 1u;
            u32* rec = a.restart + ((b0 + lane) * nrec + (tix * TILE) / a.seg_syms - 1u) * 3u;
            rec[0] = ((u32)a.slot_stride - out.w) + acc.cnt;  // bytes emitted, the ones still in the register included
            rec[1] = x;
            rec[2] = 0;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32) k_ans_enc_byte(EncArgs a)
{
    extern __shared__ __align__(16) u8 ans_sm[];
    const u32 sbase = smem_addr(ans_sm);
    const u32 lane = lane_id();
    const u32 cum_a = sbase + lane * 2u, mag_a = sbase + ANS_ENC_BYTE_CUM + lane * 4u;
    const u32 tiles = sbase + ANS_ENC_BYTE_CUM + ANS_ENC_BYTE_MAGIC;
    const u64 b0 = (u64)blockIdx.x * 32u;
    const u64 b = b0 + lane;
    const bool has = b < a.nblocks;
    u32 n_b = 0;
    if(has) {
        const u64 lo = b * (u64)a.block;
        n_b = (u32)((a.n - lo < a.block) ? (a.n - lo) : a.block);
    }
    u8* slot = a.slots + (has ? b : b0) * a.slot_stride;
    const u32 n_max = __reduce_max_sync(FULL, n_b);
    const u32 ntiles = (n_max + TILE - 1) / TILE;
    stage_tile(tiles + ((ntiles - 1) & 1u) * TILE_BYTES, a.src, a.n, b0, a.block, (ntiles - 1) * TILE, lane);
    cp_async_commit();
    {
        // the normalised model the model kernel left at the head of the slot
        const u32* hdr = reinterpret_cast<const u32*>(slot);
        u32 c0 = 0;
#pragma unroll 1
        for(u32 s = 0; s < 256u; ++s) {
            const u32 c1 = has ? hdr[2u + s] : 0u;
            sts16v(cum_a + s * 64u, c0);
            sts32v(mag_a + s * 128u, rc_magic(c1 - c0));
            c0 = c1;
        }
        sts16v(cum_a + 256u * 64u, c0);
    }
    __syncwarp();
    u32 x = ANS_BYTE_LOW;  // init (cppans.h:260-263)
    AnsByteAcc acc;
    ans_acc_init(acc);
    AnsSlotOut out{slot, (u32)a.slot_stride};
    if(__any_sync(FULL, n_b != ntiles * TILE)) {
        ans_enc_byte_tiles<true>(a, cum_a, mag_a, tiles, b0, n_b, ntiles, x, acc, out, lane);
    } else {
        ans_enc_byte_tiles<false>(a, cum_a, mag_a, tiles, b0, n_b, ntiles, x, acc, out, lane);
    }
    if(has) {
        ans_acc_finish(acc, x, out);
        a.sizes[b] = ANS_HDR + ((u32)a.slot_stride - out.w);
    }
}

// ======================================================================== A6 ==
// Byte-variant decode, one block per lane: cumulative counts as u16 [entry][lane] in shared
// memory, symbol by the 8 x 8 x 4 search of ans_lane.cuh, stream through the per-lane word
// queue of the range decoder (WordSrc), output through its 64-symbol tiles.
constexpr u32 ANS_DEC_BYTE_CUM = 257u * 64u;
constexpr u32 ANS_DEC_BYTE_SMEM = ANS_DEC_BYTE_CUM + TILE_BYTES + INQ_BYTES;

struct AnsCumTab {
    enum : u32 { UNIT = 64 };  // position = symbol * 64 = byte offset inside the lane's column
    u32 base;
    __device__ __forceinline__ u32 at(u32 pos) const { return lds16(base + pos); }
};

template <bool RAGGED, class Src>
__device__ __forceinline__ void ans_dec_byte_tile(const AnsCumTab& tab, const u32 (&k1)[8], RcDec& d, u32& x, Src& src,
                                                  u32 otile_a, u32 tile_off, u32 n_b, u32 lane)
{
#pragma unroll 1
    for(int wi = 0; wi < TILE / 4; ++wi) {
        u32 word = 0;
#pragma unroll
        for(int k = 0; k < 4; ++k) {
            if(!RAGGED || tile_off + wi * 4 + k < n_b) {
                const u32 slot = x & ((1u << ANS_BYTE_SCALE_BITS) - 1u);
                u32 sym, start, f;
                ans_find(tab, k1, slot, sym, start, f);
                if(k & 1) {
                    ans_byte_advance<true>(d, x, slot, start, f, src);
                } else {
                    ans_byte_advance<false>(d, x, slot, start, f, src);
                }
                word |= sym << (8 * k);
            }
        }
        sts32v(otile_a + lane * ROW + wi * 4, word);
    }
}

__global__ void __launch_bounds__(32) k_ans_dec_byte(DecArgs a)
{
    extern __shared__ __align__(16) u8 ans_sm[];
    const u32 sbase = smem_addr(ans_sm);
    u8* otile = ans_sm + ANS_DEC_BYTE_CUM;
    const u32 otile_a = sbase + ANS_DEC_BYTE_CUM;
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
    const u8* pay = a.payload;
    u64 len = 0;
    if(has) {
        const u64 o0 = a.offsets[b], o1 = a.offsets[b + 1];
        if(o0 <= o1 && o1 <= a.payload_len) {
            pay = a.payload + o0;
            len = o1 - o0;
        }
    }
    bool ok = has && len >= (u64)ANS_HDR + 4u;
    auto ld32u = [&](u32 off) -> u32 {
        return (u32)pay[off] | ((u32)pay[off + 1u] << 8) | ((u32)pay[off + 2u] << 16) | ((u32)pay[off + 3u] << 24);
    };
    if(ok) {
        ok = ld32u(0) == n_b;  // the container, not the payload, says how long block b is
    }
    const u32 cum_a = sbase + lane * 2u;
    u32 k1[8];
    {
        u32 prev = 0;
        bool sane = true;
#pragma unroll 1
        for(u32 s = 0; s <= 256u; ++s) {
            const u32 c = ok ? ld32u(4u + 4u * s) : (s == 256u ? (1u << ANS_BYTE_SCALE_BITS) : 0u);
            sane = sane && c >= prev && c <= (1u << ANS_BYTE_SCALE_BITS) && (s != 0u || c == 0u);
            sts16v(cum_a + s * 64u, c);
            prev = c;
        }
        sane = sane && prev == (1u << ANS_BYTE_SCALE_BITS);
        if(ok && !sane) {
            ok = false;
            // keep the search well defined: one symbol owning the whole scale
            for(u32 s = 0; s <= 256u; ++s) {
                sts16v(cum_a + s * 64u, s == 0u ? 0u : (1u << ANS_BYTE_SCALE_BITS));
            }
        }
#pragma unroll
        for(int j = 0; j < 8; ++j) {
            k1[j] = lds16v(cum_a + (32u * j) * 64u);
        }
    }
    if(has && !ok) {
        atomicOr(a.err, ERR_CORRUPT);
    }
    if(!ok) {
        n_b = 0;
    }
    WordSrc src;
    {
        const u8* coded = pay + ANS_HDR;
        const u8* wbase = (const u8*)((uintptr_t)coded & ~(uintptr_t)3);
        src.base = reinterpret_cast<const u32*>(wbase);
        const u64 room = (u64)((a.payload + a.payload_len) - wbase);
        src.lim = ok ? (u32)(room < 0xFFFFFFF0ull ? room : 0xFFFFFFF0ull) : 0u;
        src.q = queue_a + lane * 4u;
        src.prime();
    }
    __syncwarp();
    const AnsCumTab tab{cum_a};
    RcDec d;
    u32 x = ans_byte_dec_init(d, (u32)((uintptr_t)(pay + ANS_HDR) & 3u), src);
    const u32 n_max = __reduce_max_sync(FULL, n_b);
    const bool ragged = __any_sync(FULL, n_b != n_max) || (n_max % TILE) != 0u;
    const u32 ntiles = (n_max + TILE - 1) / TILE;
#pragma unroll 1
    for(u32 tix = 0; tix < ntiles; ++tix) {
        const bool inside = __all_sync(FULL, src.tile_is_inside());
        if(!ragged && inside) {
            WordSrcInside in{src};
            ans_dec_byte_tile<false>(tab, k1, d, x, in, otile_a, tix * TILE, n_b, lane);
        } else {
            ans_dec_byte_tile<true>(tab, k1, d, x, src, otile_a, tix * TILE, n_b, lane);
        }
        __syncwarp();
        store_tile(otile, a.dst, a.n, b0, a.block, tix * TILE, lane);
        __syncwarp();
    }
    // a valid stream hands the state back where the encoder started it (cppans.h:260-263)
    if(ok && x != ANS_BYTE_LOW) {
        atomicOr(a.err, ERR_CORRUPT);
    }
}

// ---------------------------------------------------------------- A6, segmented --
// Byte-variant decode from restart points: the encoder recorded, after
// coding everything from symbol j * seg_syms on, its state x and the number of bytes it had
// emitted; a decoder that reaches that symbol holds the same x and has consumed everything
// emitted AFTER that moment, so it stands E bytes before the end of the payload.  CTA = 4 warps
// over the same 32 blocks, warp = segment, one set of cumulative tables per CTA.
constexpr u32 ANS_SEG_WARPS = 4;
constexpr u32 ANS_DEC_BYTE_SEG_SMEM = ANS_DEC_BYTE_CUM + ANS_SEG_WARPS * (TILE_BYTES + INQ_BYTES);

__global__ void __launch_bounds__(32 * ANS_SEG_WARPS) k_ans_dec_byte_seg(DecArgs a)
{
    extern __shared__ __align__(16) u8 ans_sm[];
    const u32 sbase = smem_addr(ans_sm);
    const u32 warp = threadIdx.x >> 5, lane = lane_id();
    u8* otile = ans_sm + ANS_DEC_BYTE_CUM + warp * (TILE_BYTES + INQ_BYTES);
    const u32 otile_a = sbase + ANS_DEC_BYTE_CUM + warp * (TILE_BYTES + INQ_BYTES);
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
    const u32 seg = blockIdx.y * ANS_SEG_WARPS + warp;
    const u8* pay = a.payload;
    u64 len = 0;
    if(has) {
        const u64 o0 = a.offsets[b], o1 = a.offsets[b + 1];
        if(o0 <= o1 && o1 <= a.payload_len && o1 - o0 < 0xFFFFFFF0ull) {
            pay = a.payload + o0;
            len = o1 - o0;
        }
    }
    bool ok = has && len >= (u64)ANS_HDR + 4u;
    auto ld32u = [&](u32 off) -> u32 {
        return (u32)pay[off] | ((u32)pay[off + 1u] << 8) | ((u32)pay[off + 2u] << 16) | ((u32)pay[off + 3u] << 24);
    };
    if(ok) {
        ok = ld32u(0) == n_b;
    }
    const u32 cum_a = sbase + lane * 2u;
    if(warp == 0) {
        u32 prev = 0;
        bool sane = true;
#pragma unroll 1
        for(u32 s = 0; s <= 256u; ++s) {
            const u32 c = ok ? ld32u(4u + 4u * s) : (s == 256u ? (1u << ANS_BYTE_SCALE_BITS) : 0u);
            sane = sane && c >= prev && c <= (1u << ANS_BYTE_SCALE_BITS) && (s != 0u || c == 0u);
            sts16v(cum_a + s * 64u, c);
            prev = c;
        }
        sane = sane && prev == (1u << ANS_BYTE_SCALE_BITS);
        if(!sane) {
            for(u32 s = 0; s <= 256u; ++s) {
                sts16v(cum_a + s * 64u, s == 0u ? 0u : (1u << ANS_BYTE_SCALE_BITS));
            }
        }
        if(has && (!ok || !sane) && blockIdx.y == 0) {
            atomicOr(a.err, ERR_CORRUPT);
        }
    }
    __syncthreads();
    if(seg >= nseg) {
        return;
    }
    // a table that warp 0 had to replace marks the block as bad for everyone
    ok = ok && !(lds16v(cum_a + 1u * 64u) == (1u << ANS_BYTE_SCALE_BITS) && lds16v(cum_a + 256u * 64u) == (1u << ANS_BYTE_SCALE_BITS) &&
                 ld32u(4u + 4u) != (1u << ANS_BYTE_SCALE_BITS));
    u32 k1[8];
#pragma unroll
    for(int j = 0; j < 8; ++j) {
        k1[j] = lds16v(cum_a + (32u * j) * 64u);
    }
    const u32 seg_lo = seg * a.seg_syms;
    u32 seg_hi = seg_lo + a.seg_syms;
    seg_hi = seg_hi < n_b ? seg_hi : n_b;
    bool mine_ok = ok && seg_lo < n_b;
    const u32 coded_len = ok ? (u32)len - ANS_HDR : 0u;  // the 4-byte state and everything emitted
    u32 x = ANS_BYTE_LOW, off = 0;                       // off: coded bytes in front of my first byte
    u32 x_end = ANS_BYTE_LOW;                            // where a sound stream leaves me
    const u32 nrec = nseg - 1u;
    if(mine_ok && seg != 0u) {
        const u32* rec = a.restart + (b * nrec + seg - 1u) * 3u;
        const u32 emitted = rec[0];
        x = rec[1];
        if(emitted == 0xFFFFFFFFu || emitted > coded_len - 4u) {
            mine_ok = false;
            atomicOr(a.err, ERR_CORRUPT);
        } else {
            off = coded_len - emitted;
        }
    }
    if(mine_ok && seg + 1u < nseg && (seg + 1u) * a.seg_syms < n_b) {
        x_end = a.restart[(b * nrec + seg) * 3u + 1u];
    }
    WordSrc src;
    const u32 skip0 = (u32)((uintptr_t)(pay + ANS_HDR) & 3u);
    {
        const u8* coded = pay + ANS_HDR;
        const u8* wbase = (const u8*)((uintptr_t)coded & ~(uintptr_t)3);
        src.base = reinterpret_cast<const u32*>(wbase);
        const u64 room = (u64)((a.payload + a.payload_len) - wbase);
        src.lim = mine_ok ? (u32)(room < 0xFFFFFFF0ull ? room : 0xFFFFFFF0ull) : 0u;
        src.q = queue_a + lane * 4u;
        src.prime(mine_ok ? (skip0 + off) >> 2 : 0u);
    }
    const AnsCumTab tab{cum_a};
    RcDec d;
    if(seg == 0u) {
        x = ans_byte_dec_init(d, skip0, src);
    } else {
        ans_byte_win_init(d, (skip0 + off) & 3u, src);
    }
    const u32 n_eff = mine_ok ? seg_hi : 0u;
    const u32 n_max = __reduce_max_sync(FULL, n_eff);
    const bool ragged = __any_sync(FULL, n_eff != n_max) || (n_max % TILE) != 0u;
    const u32 tix0 = seg_lo / TILE;
    const u32 ntiles = (n_max + TILE - 1) / TILE;
#pragma unroll 1
    for(u32 tix = tix0; tix < ntiles; ++tix) {
        const bool inside = __all_sync(FULL, src.tile_is_inside());
        if(!ragged && inside) {
            WordSrcInside in{src};
            ans_dec_byte_tile<false>(tab, k1, d, x, in, otile_a, tix * TILE, n_eff, lane);
        } else {
            ans_dec_byte_tile<true>(tab, k1, d, x, src, otile_a, tix * TILE, n_eff, lane);
        }
        __syncwarp();
        store_tile(otile, a.dst, a.n, b0, a.block, tix * TILE, lane);
        __syncwarp();
    }
    // a sound stream hands the state on to the next segment exactly as the encoder recorded it
    if(mine_ok && x != x_end) {
        atomicOr(a.err, ERR_CORRUPT);
    }
}

}  // namespace b2rc
