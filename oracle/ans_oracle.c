/*
 * ans_oracle.c -- see ans_oracle.h.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates cppans.h (cppans::rANS) in plain C.  The arithmetic is the reference's,
 * including the places where its u32 expressions wrap; the code is not.
 */
#include "ans_oracle.h"

#include <stdlib.h>
#include <string.h>

enum {
    BYTE_BITS = 14,          /* rANS::ProbBits          cppans.h:27 */
    BYTE_LOW = 1u << 23,     /* rANS::rANSByteLowBounds cppans.h:29 */
    WORD_BITS = 12,          /* rANS::WordScaleBits     cppans.h:31 */
    WORD_LOW = 1u << 16,     /* rANS::WordLowBounds     cppans.h:30 */
    WORD_STATES = 8          /* cppans.h:586 */
};

size_t rao_slot_bytes(uint32_t n)
{
    size_t s = (size_t)RAO_HEADER + 4u * WORD_STATES + 2u * (size_t)n;
    return (s + 15u) & ~(size_t)15u;
}

/* cppans.h:138-177 */
void rao_normalize(uint32_t freq[256], uint32_t cum[257], uint32_t target)
{
    const uint32_t total = cum[256];
    for(int i = 1; i <= 256; ++i) {
        cum[i] = (uint32_t)(((uint64_t)target * cum[i]) / total);
    }
    for(int i = 0; i < 256; ++i) {
        if(freq[i] == 0 || cum[i + 1] != cum[i]) {
            continue;
        }
        /* symbol i occurs but lost its slice: take one unit from the narrowest slice
         * that can spare it (first one wins ties), sliding the boundaries in between */
        uint32_t narrowest = 0xFFFFFFFFu;
        int donor = -1;
        for(int j = 0; j < 256; ++j) {
            const uint32_t w = cum[j + 1] - cum[j];
            if(w > 1 && w < narrowest) {
                narrowest = w;
                donor = j;
            }
        }
        if(donor < 0) {
            continue; /* cannot happen with 256 symbols and target >= 512 (cppans.h:154) */
        }
        if(donor < i) {
            for(int j = donor + 1; j <= i; ++j) {
                cum[j]--;
            }
        } else {
            for(int j = i + 1; j <= donor; ++j) {
                cum[j]++;
            }
        }
    }
    for(int i = 0; i < 256; ++i) {
        freq[i] = cum[i + 1] - cum[i];
    }
}

/* count + cumulative + normalize (cppans.h:102-136, :504-508) */
void rao_model(const uint8_t* src, uint32_t n, uint32_t target, uint32_t freq[256], uint32_t cum[257])
{
    memset(freq, 0, 256 * sizeof(uint32_t));
    for(uint32_t i = 0; i < n; ++i) {
        freq[src[i]]++;
    }
    cum[0] = 0;
    for(int i = 0; i < 256; ++i) {
        cum[i + 1] = cum[i] + freq[i];
    }
    rao_normalize(freq, cum, target);
}

static void put_u32(uint8_t* p, uint32_t v)
{
    p[0] = (uint8_t)v;
    p[1] = (uint8_t)(v >> 8);
    p[2] = (uint8_t)(v >> 16);
    p[3] = (uint8_t)(v >> 24);
}

static uint32_t get_u32(const uint8_t* p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* Both encoders fill a scratch buffer from its end towards its start, like the
 * reference fills dst (cppans.h:515, :591); `w` is the write cursor. */
static long finish_payload(uint8_t* scratch, size_t scratch_len, size_t w, uint32_t n, const uint32_t cum[257],
                           uint8_t* dst, size_t cap)
{
    if(w < RAO_HEADER) {
        return -1;
    }
    w -= RAO_HEADER; /* cppans.h:521-527: size, then the 257 cumulative counts */
    put_u32(scratch + w, n);
    for(int i = 0; i <= 256; ++i) {
        put_u32(scratch + w + 4 + 4 * (size_t)i, cum[i]);
    }
    const size_t len = scratch_len - w;
    if(len > cap) {
        return -1;
    }
    memcpy(dst, scratch + w, len);
    return (long)len;
}

/* rANS::encode (cppans.h:497-530) with put (:265-287) and flush (:289-299).
 * put() divides by multiplying with a rounded-up reciprocal; for every state the
 * coder can reach (x < 2^31) that is the exact quotient, and its freq == 1 special
 * case (:205-233) lands on the same x*M + start, so plain division restates it. */
static long byte_encode(const uint8_t* src, uint32_t n, uint8_t* dst, size_t cap)
{
    uint32_t freq[256], cum[257];
    rao_model(src, n, 1u << BYTE_BITS, freq, cum);
    const size_t scratch_len = rao_slot_bytes(n);
    uint8_t* scratch = (uint8_t*)malloc(scratch_len);
    if(!scratch) {
        return -1;
    }
    size_t w = scratch_len;
    uint32_t x = BYTE_LOW;
    for(uint32_t i = n; i > 0; --i) {
        const uint32_t s = src[i - 1];
        const uint32_t f = freq[s];
        const uint32_t x_max = ((BYTE_LOW >> BYTE_BITS) << 8) * f; /* :203 */
        while(x >= x_max) {
            scratch[--w] = (uint8_t)x;
            x >>= 8;
        }
        x = ((x / f) << BYTE_BITS) + (x % f) + cum[s];
    }
    w -= 4;
    put_u32(scratch + w, x);
    const long r = finish_payload(scratch, scratch_len, w, n, cum, dst, cap);
    free(scratch);
    return r;
}

/* rANS::decode (cppans.h:532-564) with init_decode (:303-310), get (:313-316),
 * advance (:321-334).  The reference does not bound its reads; the oracle does. */
static long byte_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
    if(n < RAO_HEADER + 4u) {
        return -1;
    }
    const uint32_t want = get_u32(src);
    if(want > cap) {
        return -1;
    }
    uint32_t cum[257];
    for(int i = 0; i <= 256; ++i) {
        cum[i] = get_u32(src + 4 + 4 * (size_t)i);
    }
    if(cum[0] != 0 || cum[256] != (1u << BYTE_BITS)) {
        return -1;
    }
    uint8_t* map = (uint8_t*)malloc(1u << BYTE_BITS);
    if(!map) {
        return -1;
    }
    for(int s = 0; s < 256; ++s) {
        if(cum[s + 1] < cum[s] || cum[s + 1] > (1u << BYTE_BITS)) {
            free(map);
            return -1;
        }
        for(uint32_t k = cum[s]; k < cum[s + 1]; ++k) {
            map[k] = (uint8_t)s;
        }
    }
    size_t r = RAO_HEADER;
    uint32_t x = get_u32(src + r);
    r += 4;
    const uint32_t mask = (1u << BYTE_BITS) - 1u;
    for(uint32_t i = 0; i < want; ++i) {
        const uint32_t slot = x & mask;
        const uint32_t s = map[slot];
        dst[i] = (uint8_t)s;
        x = (cum[s + 1] - cum[s]) * (x >> BYTE_BITS) + slot - cum[s];
        while(x < BYTE_LOW) {
            if(r >= n) {
                free(map);
                return -1;
            }
            x = (x << 8) | src[r++];
        }
    }
    free(map);
    return (long)want;
}

/* rANS::encode_simd (cppans.h:567-607) with wordEncPut (:353-364) and wordEncFlush
 * (:367-373).  All products are u32, so the
 * renormalisation bound wraps to 0 when one symbol holds the whole scale. */
static long word_encode(const uint8_t* src, uint32_t n, uint8_t* dst, size_t cap)
{
    uint32_t freq[256], cum[257];
    rao_model(src, n, 1u << WORD_BITS, freq, cum);
    const size_t scratch_len = rao_slot_bytes(n);
    uint8_t* scratch = (uint8_t*)malloc(scratch_len);
    if(!scratch) {
        return -1;
    }
    size_t w = scratch_len;
    uint32_t st[WORD_STATES];
    for(int k = 0; k < WORD_STATES; ++k) {
        st[k] = WORD_LOW;
    }
    for(uint32_t i = n; i > 0; --i) {
        const uint32_t p = i - 1;
        const uint32_t s = src[p];
        const uint32_t f = freq[s];
        uint32_t x = st[p & (WORD_STATES - 1)];
        const uint32_t bound = (uint32_t)(((WORD_LOW >> WORD_BITS) << 16) * f);
        if(bound <= x) {
            w -= 2;
            scratch[w] = (uint8_t)x;
            scratch[w + 1] = (uint8_t)(x >> 8);
            x >>= 16;
        }
        st[p & (WORD_STATES - 1)] = ((x / f) << WORD_BITS) + (x % f) + cum[s];
    }
    for(int k = WORD_STATES; k > 0; --k) {
        w -= 4;
        put_u32(scratch + w, st[k - 1]);
    }
    const long r = finish_payload(scratch, scratch_len, w, n, cum, dst, cap);
    free(scratch);
    return r;
}

/* rANS::decode_simd (cppans.h:609-649).  The SSE code decodes states 0..3 and 4..7 as
 * two vectors, then refills them in lane order (simdDecRenorm, :443-488: the shuffle
 * table hands the next u16s to the lanes that fell below 2^16, lowest lane first);
 * the tail symbols are decoded without a refill (:643-647). */
static long word_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
    if(n < RAO_HEADER + 4u * WORD_STATES) {
        return -1;
    }
    const uint32_t want = get_u32(src);
    if(want > cap) {
        return -1;
    }
    uint32_t cum[257];
    for(int i = 0; i <= 256; ++i) {
        cum[i] = get_u32(src + 4 + 4 * (size_t)i);
    }
    if(cum[0] != 0 || cum[256] != (1u << WORD_BITS)) {
        return -1;
    }
    uint8_t sym_of[1u << WORD_BITS];
    for(int s = 0; s < 256; ++s) {
        if(cum[s + 1] < cum[s] || cum[s + 1] > (1u << WORD_BITS)) {
            return -1;
        }
        for(uint32_t k = cum[s]; k < cum[s + 1]; ++k) {
            sym_of[k] = (uint8_t)s;
        }
    }
    size_t r = RAO_HEADER;
    uint32_t st[WORD_STATES];
    for(int k = 0; k < WORD_STATES; ++k) {
        st[k] = get_u32(src + r);
        r += 4;
    }
    const uint32_t mask = (1u << WORD_BITS) - 1u;
    const uint32_t full = want & ~(uint32_t)(WORD_STATES - 1);
    for(uint32_t i = 0; i < want; ++i) {
        const uint32_t k = i & (WORD_STATES - 1);
        const uint32_t slot = st[k] & mask;
        const uint32_t s = sym_of[slot];
        dst[i] = (uint8_t)s;
        st[k] = (cum[s + 1] - cum[s]) * (st[k] >> WORD_BITS) + (slot - cum[s]);
        if(i < full && k == WORD_STATES - 1) {
            /* all eight of this round are decoded; now refill in lane order */
            for(int j = 0; j < WORD_STATES; ++j) {
                if(st[j] < WORD_LOW) {
                    if(r + 2 > n) {
                        return -1;
                    }
                    st[j] = (st[j] << 16) | (uint32_t)src[r] | ((uint32_t)src[r + 1] << 8);
                    r += 2;
                }
            }
        }
    }
    return (long)want;
}

long rao_encode(int variant, const uint8_t* src, uint32_t n, uint8_t* dst, size_t cap)
{
    if(n == 0 || !src || !dst) {
        return -1; /* the reference asserts 0 < src_size (cppans.h:502, :571) */
    }
    return variant == RAO_WORD ? word_encode(src, n, dst, cap) : byte_encode(src, n, dst, cap);
}

long rao_decode(int variant, const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
    if(!src || !dst) {
        return -1;
    }
    return variant == RAO_WORD ? word_decode(src, n, dst, cap) : byte_decode(src, n, dst, cap);
}
