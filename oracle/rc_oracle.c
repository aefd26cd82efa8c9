/*
 * rc_oracle.c -- CPU restatement of cpprcoder.h's static and adaptive order-0
 * range coder.  TEST INFRASTRUCTURE ONLY (see rc_oracle.h): the product path
 * never calls this file.  Parity status: PINNED against the unmodified
 * reference (oracle/_ref) and the golden vectors in tests/golden/.
 *
 * Written from the behaviour of /root/reference/cpprcoder.h; each routine names
 * the lines it restates.  All arithmetic is uint32_t, as in the reference.
 */
#include "rc_oracle.h"
#include "ans_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#define MIN_RANGE 0x01000000u /* cpprcoder.h:327, :631 */
#define TOP_BYTE 0xFF000000u  /* (0xFFU << SHIFT), cpprcoder.h:419, :784 */

/* ------------------------------------------------------------------ sink -- */
/* Bounded byte sink.  The reference writes through MemoryStream::writeByte,
 * which fails at capacity (cpprcoder.h:1047-1054); `bad` models that `false`. */
typedef struct {
    uint8_t* p;
    size_t n, cap;
    int bad;
} sink;

static void put(sink* s, uint32_t byte)
{
    if(s->n < s->cap) {
        s->p[s->n++] = (uint8_t)byte;
    } else {
        s->bad = 1;
    }
}

static void put_le32(sink* s, uint32_t v)
{
    put(s, v);
    put(s, v >> 8);
    put(s, v >> 16);
    put(s, v >> 24);
}

static void put_be32(sink* s, uint32_t v)
{
    put(s, v >> 24);
    put(s, v >> 16);
    put(s, v >> 8);
    put(s, v);
}

size_t rco_slot_bytes(uint32_t n)
{
    size_t s = (size_t)n + n / 8 + 1024;
    return (s + 127) & ~(size_t)127;
}

uint64_t rco_fnv1a64(const uint8_t* p, size_t n, uint64_t seed)
{
    uint64_t h = seed ? seed : 1469598103934665603ull;
    for(size_t i = 0; i < n; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

/* ---------------------------------------------------------------- static -- */

/* cpprcoder.h:543-571.  The test `0xFFFF <= freq[c]` is made on the symbol that
 * is about to be counted, before the increment; a hit halves EVERY non-zero
 * count with (x >> 1) | 1.  Blocks above 2^24 bytes get one more scaling pass
 * that starts at symbol 1 (cpprcoder.h:567), leaving symbol 0 untouched. */
void rco_static_count(const uint8_t* src, uint32_t n, uint32_t freq[256], uint32_t* rescales)
{
    uint32_t events = 0;
    memset(freq, 0, 256 * sizeof(uint32_t));
    for(uint32_t i = 0; i < n; ++i) {
        const uint32_t c = src[i];
        if(freq[c] >= 0xFFFFu) {
            for(int s = 0; s < 256; ++s) {
                if(freq[s] != 0) {
                    freq[s] = (freq[s] >> 1) | 1u;
                }
            }
            ++events;
        }
        freq[c] += 1;
    }
    if(n > MIN_RANGE) {
        uint32_t shift = 0, m = n;
        while(m > MIN_RANGE) {
            m >>= 1;
            ++shift;
        }
        for(int s = 1; s < 256; ++s) {
            freq[s] = freq[s] ? ((freq[s] >> shift) | 1u) : 0u;
        }
    }
    if(rescales) {
        *rescales = events;
    }
}

/* cpprcoder.h:573-583: cum[s] = sum of freq below s, cum[256] = total. */
static void exclusive_prefix(const uint32_t freq[256], uint32_t cum[257])
{
    uint32_t run = 0;
    for(int s = 0; s < 256; ++s) {
        cum[s] = run;
        run += freq[s];
    }
    cum[256] = run;
}

/* cpprcoder.h:375-458. */
long rco_static_encode(const uint8_t* src, uint32_t n, uint8_t* dst, size_t cap, rco_stats* st)
{
    uint32_t freq[256], cum[257];
    rco_stats local;
    memset(&local, 0, sizeof local);
    sink out = {dst, 0, cap, 0};

    rco_static_count(src, n, freq, &local.rescales);
    put_le32(&out, n); /* :386-393 */
    for(int s = 0; s < 256; ++s) { /* write16, :604-619: u16 in host order (x86: little endian) */
        put(&out, freq[s] & 0xFFu);
        put(&out, (freq[s] >> 8) & 0xFFu);
    }
    exclusive_prefix(freq, cum);
    const uint32_t total = cum[256];

    uint32_t range = 0xFFFFFFFFu, low = 0, run = 0 /* count_ */, held = 0 /* buffer_, u32 */;
    for(uint32_t i = 0; i < n; ++i) { /* :400-436 */
        const uint32_t c = src[i];
        const uint32_t t = range / total;
        const uint32_t next = low + cum[c] * t;
        range = (cum[c + 1] - cum[c]) * t;
        if(next < low) { /* carry out of the 32-bit window, :405-415 */
            ++held;
            ++local.carries;
            if(run > 0) {
                ++local.carries_with_run;
                for(; run != 0; --run) {
                    put(&out, held);
                    held = 0;
                }
            }
        }
        low = next;
        while(range < MIN_RANGE) { /* :418-435 */
            if(low < TOP_BYTE) {
                put(&out, held);
                for(; run != 0; --run) {
                    put(&out, 0xFFu);
                }
                held = low >> 24;
            } else {
                ++run;
                if(run > local.max_pending_run) {
                    local.max_pending_run = run;
                }
            }
            low <<= 8;
            range <<= 8;
        }
    }
    /* flush, :439-457 -- note the low_ == 0xFFFFFFFF quirk: the held byte is bumped
     * and the pending run becomes zeros, yet low_ itself is still written as FF FF FF FF. */
    local.final_low = low;
    uint32_t fill = 0xFFu;
    if(low >= 0xFFFFFFFFu) {
        ++held;
        fill = 0;
    }
    put(&out, held);
    for(; run != 0; --run) {
        put(&out, fill);
    }
    put_be32(&out, low);
    if(st) {
        *st = local;
    }
    return out.bad ? -1 : (long)out.n;
}

/* cpprcoder.h:521-535: smallest s with cum[s+1] > target, never above 255. */
static uint32_t static_find(const uint32_t cum[257], uint32_t target)
{
    uint32_t lo = 0, hi = 255;
    while(lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if(cum[mid + 1] <= target) {
            lo = mid + 1;
        } else {
            hi = mid;
        }
    }
    return lo;
}

/* cpprcoder.h:460-519 (+ read16 :585-602). */
long rco_static_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
    if(n < RCO_STATIC_HEADER) { /* :468-476 */
        return -1;
    }
    const uint32_t want = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
    if(want == 0) { /* :481-483 */
        return 0;
    }
    uint32_t freq[256], cum[257];
    for(int s = 0; s < 256; ++s) {
        freq[s] = (uint32_t)src[4 + 2 * s] | ((uint32_t)src[5 + 2 * s] << 8);
    }
    const uint8_t* p = src + RCO_STATIC_HEADER;
    const uint8_t* end = src + n;
    if(!(p < end)) { /* read16 returns bytes < end, :601 */
        return -1;
    }
    exclusive_prefix(freq, cum);
    const uint32_t total = cum[256];
    if((size_t)(end - p) < 5 || total == 0) { /* :491-493; total==0 would divide by zero in the reference */
        return -1;
    }
    /* the first coded byte is the encoder's dummy buffer_ and is skipped, :494-498 */
    uint32_t low = ((uint32_t)p[1] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 8) | p[4];
    p += 5;
    uint32_t range = 0xFFFFFFFFu;
    size_t made = 0;
    for(uint32_t i = 0; i < want; ++i) { /* :500-517 */
        const uint32_t t = range / total;
        if(t == 0) {
            return -1; /* unreachable on valid streams (range >= 2^24 > total) */
        }
        const uint32_t c = static_find(cum, low / t);
        low -= cum[c] * t;
        range = (cum[c + 1] - cum[c]) * t;
        while(range < MIN_RANGE) {
            if(p >= end) {
                return -1;
            }
            range <<= 8;
            low = (low << 8) | *p++;
        }
        if(made >= cap) {
            return -1;
        }
        dst[made++] = (uint8_t)c;
    }
    return (long)made;
}

/* -------------------------------------------------------------- adaptive -- */

/* AdaptiveFrequencyTable, cpprcoder.h:256-314 and the scalar branches of
 * :1094-1261 (the SSE2 branches compute the same values). */
typedef struct {
    uint32_t total;
    uint32_t freq[256];
    uint32_t upto[16]; /* prefix_: inclusive sums of the 16-symbol chunks */
} model;

static void model_chunks(model* m) /* countChunks, :1245-1261 */
{
    uint32_t run = 0;
    for(int k = 0; k < 16; ++k) {
        for(int j = 0; j < 16; ++j) {
            run += m->freq[16 * k + j];
        }
        m->upto[k] = run;
    }
}

static void model_init(model* m) /* :1094-1132 */
{
    m->total = 256;
    for(int s = 0; s < 256; ++s) {
        m->freq[s] = 1;
    }
    model_chunks(m);
}

static int model_update(model* m, uint32_t c) /* :1134-1177; returns 1 when it halved */
{
    m->freq[c] += 1;
    m->total += 1;
    if(m->total >= MIN_RANGE) {
        uint32_t sum = 0;
        for(int s = 0; s < 256; ++s) {
            m->freq[s] = (m->freq[s] >> 1) | 1u;
            sum += m->freq[s];
        }
        m->total = sum;
        model_chunks(m);
        return 1;
    }
    for(uint32_t k = c >> 4; k < 16; ++k) {
        m->upto[k] += 1;
    }
    return 0;
}

static uint32_t model_below(const model* m, uint32_t c) /* cumulative, :1179-1187 */
{
    const uint32_t k = c >> 4;
    uint32_t acc = k ? m->upto[k - 1] : 0;
    for(uint32_t s = k << 4; s < c; ++s) {
        acc += m->freq[s];
    }
    return acc;
}

/* find, scalar branch :1221-1241.  When target >= total no chunk matches, the
 * scan restarts at symbol 0 and falls off the end with code 0 / below = total. */
static void model_find(const model* m, uint32_t target, uint32_t* below, uint32_t* code)
{
    uint32_t acc = 0, k = 0;
    for(uint32_t i = 0; i < 16; ++i) {
        if(target < m->upto[i]) {
            k = i;
            acc = k ? m->upto[k - 1] : 0;
            break;
        }
    }
    uint32_t sym = k << 4;
    for(uint32_t s = sym; s < 256; ++s) {
        const uint32_t next = acc + m->freq[s];
        if(target < next) {
            *below = acc;
            *code = s;
            return;
        }
        acc = next;
    }
    *below = acc;
    *code = sym;
}

/* cpprcoder.h:678-802. */
long rco_adaptive_encode(const uint8_t* src, uint32_t n, uint8_t* dst, size_t cap, rco_stats* st)
{
    model m;
    rco_stats local;
    memset(&local, 0, sizeof local);
    sink out = {dst, 0, cap, 0};
    model_init(&m);
    put_le32(&out, n); /* initialize, :689-694 */

    uint32_t range = 0xFFFFFF00u, low = 0, run = 0 /* carry_ */;
    uint8_t held = 0; /* buffer_ is u8 here, :657 */
    for(uint32_t i = 0; i < n; ++i) { /* :702-713 */
        const uint32_t c = src[i];
        const uint32_t t = range / m.total;
        const uint32_t before = low;
        low += model_below(&m, c) * t;
        range = m.freq[c] * t;
        /* normalize(prevLow), :764-802 */
        if(low < before) {
            held = (uint8_t)(held + 1);
            ++local.carries;
            if(run > 0) {
                ++local.carries_with_run;
                put(&out, held);
                for(uint32_t k = 1; k < run; ++k) {
                    put(&out, 0);
                }
                held = 0;
                run = 0;
            }
        }
        while(range < MIN_RANGE) {
            if(low < TOP_BYTE) {
                put(&out, held);
                for(uint32_t k = 0; k < run; ++k) {
                    put(&out, 0xFFu);
                }
                held = (uint8_t)(low >> 24);
                run = 0;
            } else {
                ++run;
                if(run > local.max_pending_run) {
                    local.max_pending_run = run;
                }
            }
            low <<= 8;
            range <<= 8;
        }
        local.rescales += (uint32_t)model_update(&m, c);
    }
    /* finish, :744-762 */
    local.final_low = low;
    put(&out, held);
    for(uint32_t k = 0; k < run; ++k) {
        put(&out, 0xFFu);
    }
    put_be32(&out, low);
    if(st) {
        *st = local;
    }
    return out.bad ? -1 : (long)out.n;
}

/* cpprcoder.h:859-940. */
long rco_adaptive_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
    if(n < 8) { /* State_Init asks for 8 bytes, :878-880 */
        return -1;
    }
    model m;
    model_init(&m);
    const uint32_t want = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
    uint32_t low = ((uint32_t)src[4] << 24) | ((uint32_t)src[5] << 16) | ((uint32_t)src[6] << 8) | src[7];
    const uint8_t* p = src + 8;
    const uint8_t* end = src + n;
    uint32_t range = 0x00FFFFFFu; /* :813, :867: the first normalize pulls the 5th byte */
    size_t made = 0;
    for(;;) { /* :900-917 */
        while(range < MIN_RANGE) { /* normalize, :926-940 */
            if(p >= end) {
                return -1;
            }
            range <<= 8;
            low = (low << 8) + *p++;
        }
        const uint32_t t = range / m.total;
        if(t == 0) {
            return -1;
        }
        uint32_t below, code;
        model_find(&m, low / t, &below, &code);
        low -= t * below;
        range = t * m.freq[code];
        if(made >= cap) {
            return -1;
        }
        dst[made++] = (uint8_t)code;
        if(want <= made) { /* tested AFTER the write: want == 0 still emits one byte, :909-914 */
            return (long)made;
        }
        model_update(&m, code);
    }
}

/* ---------------------------------------------------------- block drivers -- */

typedef struct {
    int mode, decode, failed;
    const uint8_t* src;
    uint64_t n;
    uint32_t block;
    uint8_t* slots;
    uint64_t slot_stride;
    uint32_t* sizes;
    const uint64_t* offsets;
    uint8_t* dst;
    uint64_t nblocks;
    atomic_ullong next;
} job;

/* modes 0/1 are the range coders of this file, 2/3 the rANS variants of ans_oracle.c */
static long one_encode(int mode, const uint8_t* src, uint32_t n, uint8_t* dst, size_t cap)
{
    switch(mode) {
    case RCO_STATIC: return rco_static_encode(src, n, dst, cap, NULL);
    case RCO_ADAPTIVE: return rco_adaptive_encode(src, n, dst, cap, NULL);
    case RAO_BYTE:
    case RAO_WORD: return rao_encode(mode, src, n, dst, cap);
    default: return -1;
    }
}

static long one_decode(int mode, const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
    switch(mode) {
    case RCO_STATIC: return rco_static_decode(src, n, dst, cap);
    case RCO_ADAPTIVE: return rco_adaptive_decode(src, n, dst, cap);
    case RAO_BYTE:
    case RAO_WORD: return rao_decode(mode, src, n, dst, cap);
    default: return -1;
    }
}

static void* worker(void* arg)
{
    job* j = (job*)arg;
    for(;;) {
        const uint64_t b = atomic_fetch_add(&j->next, 1);
        if(b >= j->nblocks) {
            break;
        }
        const uint64_t at = b * (uint64_t)j->block;
        const uint32_t len = (uint32_t)((j->n - at < j->block) ? (j->n - at) : j->block);
        long r;
        if(!j->decode) {
            uint8_t* slot = j->slots + b * j->slot_stride;
            r = one_encode(j->mode, j->src + at, len, slot, j->slot_stride);
            if(r < 0) {
                j->failed = 1;
            } else {
                j->sizes[b] = (uint32_t)r;
            }
        } else {
            const uint8_t* pay = j->src + j->offsets[b];
            const size_t paylen = (size_t)(j->offsets[b + 1] - j->offsets[b]);
            r = one_decode(j->mode, pay, paylen, j->dst + at, len);
            if(r != (long)len) {
                j->failed = 1;
            }
        }
    }
    return NULL;
}

static int run_job(job* j, int threads)
{
    atomic_init(&j->next, 0);
    j->failed = 0;
    if(threads <= 1) {
        worker(j);
        return j->failed ? -1 : 0;
    }
    if(threads > 256) {
        threads = 256;
    }
    pthread_t tid[256];
    int started = 0;
    for(int i = 0; i < threads; ++i) {
        if(pthread_create(&tid[i], NULL, worker, j) != 0) {
            break;
        }
        ++started;
    }
    if(started == 0) {
        worker(j);
    }
    for(int i = 0; i < started; ++i) {
        pthread_join(tid[i], NULL);
    }
    return j->failed ? -1 : 0;
}

int rco_encode_blocks(int mode, const uint8_t* src, uint64_t n, uint32_t block, uint8_t* slots, uint64_t slot_stride,
                      uint32_t* sizes, int threads)
{
    if(block == 0) {
        return -1;
    }
    job j;
    memset(&j, 0, sizeof j);
    j.mode = mode;
    j.src = src;
    j.n = n;
    j.block = block;
    j.slots = slots;
    j.slot_stride = slot_stride;
    j.sizes = sizes;
    j.nblocks = (n + block - 1) / block;
    return run_job(&j, threads);
}

int rco_decode_blocks(int mode, const uint8_t* stream, const uint64_t* offsets, uint64_t nblocks, uint32_t block,
                      uint8_t* dst, uint64_t n, int threads)
{
    if(block == 0 || nblocks != (n + block - 1) / block) {
        return -1;
    }
    job j;
    memset(&j, 0, sizeof j);
    j.mode = mode;
    j.decode = 1;
    j.src = stream;
    j.offsets = offsets;
    j.n = n;
    j.block = block;
    j.dst = dst;
    j.nblocks = nblocks;
    return run_job(&j, threads);
}
