/*
 * blk_oracle.c -- plain-C restatement of blksort::BlkSort (see blk_oracle.h).
 * TEST INFRASTRUCTURE ONLY.  Every routine names the blksort.h lines it follows.
 *
 * The reference sorts `Item {u8* str_; u16 id_}` records whose str_ points into a
 * doubled copy of the block (blksort.h:451-487); str_ is always buffer + id_, so a
 * row is carried here as its u16 id alone and "str_[d]" is twin[id + d].
 */
#include "blk_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define N BSO_BLOCK

typedef struct {
    const uint8_t* twin; /* the block twice in a row: rotation r is twin + r */
} Rows;

static inline uint8_t at(const Rows* R, uint16_t id, uint32_t d)
{
    return R->twin[(uint32_t)id + d];
}

/* less, blksort.h:183-208 (the scalar branch; the SSE branch is compiled out) */
static int row_less(const Rows* R, uint16_t a, uint16_t b, uint32_t depth)
{
    const uint8_t* x = R->twin + a;
    const uint8_t* y = R->twin + b;
    for(uint32_t d = 0; d < depth; ++d) {
        if(x[d] != y[d]) {
            return x[d] < y[d];
        }
    }
    return 0;
}

/* median, blksort.h:166-181.  Looks at byte 0 of the rows at the three quartile
 * positions whatever the depth being partitioned -- kept, it decides the swaps. */
static uint16_t pick_pivot(const Rows* R, uint32_t size, const uint16_t* v)
{
    const uint32_t q1 = size >> 2, q2 = q1 + q1, q3 = q1 + q2;
    const uint8_t a = at(R, v[q1], 0), b = at(R, v[q2], 0), c = at(R, v[q3], 0);
    if(a < b) {
        return b < c ? v[q2] : (a < c ? v[q3] : v[q1]);
    }
    return a < c ? v[q1] : (b < c ? v[q3] : v[q2]);
}

/* insertionsort, blksort.h:223-233 */
static void insertion(const Rows* R, uint32_t size, uint16_t* v, uint32_t depth)
{
    for(uint32_t i = 1; i < size; ++i) {
        const uint16_t x = v[i];
        int64_t j = (int64_t)i - 1;
        while(j >= 0 && row_less(R, x, v[j], depth)) {
            v[j + 1] = v[j];
            --j;
        }
        v[j + 1] = x;
    }
}

/* the sift-down both phases of heapsort share, blksort.h:244-256 and :264-276 (1-based) */
static void sift(const Rows* R, uint16_t* h1, int32_t root, int32_t last, uint16_t x, uint32_t depth)
{
    int32_t i = root, j;
    while((j = i << 1) <= last) {
        if(j < last && row_less(R, h1[j], h1[j + 1], depth)) {
            ++j;
        }
        if(!row_less(R, x, h1[j], depth)) {
            break;
        }
        h1[i] = h1[j];
        i = j;
    }
    h1[i] = x;
}

/* heapsort, blksort.h:235-279 */
static void heap(const Rows* R, uint32_t n, uint16_t* v, uint32_t depth)
{
    uint16_t* h1 = v - 1;
    int32_t last = (int32_t)n;
    for(int32_t k = last >> 1; k >= 1; --k) {
        sift(R, h1, k, last, h1[k], depth);
    }
    while(last > 1) {
        const uint16_t x = h1[last];
        h1[last] = h1[1];
        --last;
        sift(R, h1, 1, last, x, depth);
    }
}

static inline void swap16(uint16_t* a, uint16_t* b)
{
    const uint16_t t = *a;
    *a = *b;
    *b = t;
}

/* mqsort, blksort.h:281-362: three-way partition on byte d with the equal keys parked at
 * both ends, then swung to the middle; "<" and ">" parts recurse AT THE SAME d with one
 * level less, the "=" part goes on at d + 1 in the loop. */
static void mkq(const Rows* R, uint32_t size, uint16_t* v, uint32_t d, uint32_t depth, int32_t level)
{
    if(level <= 0) {
        heap(R, size, v, depth);
        return;
    }
    while(d < depth) {
        if(size < 37u) {
            insertion(R, size, v, depth);
            return;
        }
        const uint8_t p = at(R, pick_pivot(R, size, v), d);
        const int32_t hi = (int32_t)size - 1;
        int32_t lo = 0, up = hi;   /* scanning cursors (i0, i1) */
        int32_t eql = 0, eqr = hi; /* next free place of the parked equal keys (m0, m1) */
        for(;;) {
            while(lo <= up) {
                const uint8_t c = at(R, v[lo], d);
                if(p < c) {
                    break;
                }
                if(p == c) {
                    swap16(&v[lo], &v[eql]);
                    ++eql;
                }
                ++lo;
            }
            while(lo <= up) {
                const uint8_t c = at(R, v[up], d);
                if(c < p) {
                    break;
                }
                if(p == c) {
                    swap16(&v[up], &v[eqr]);
                    --eqr;
                }
                --up;
            }
            if(up < lo) {
                break;
            }
            swap16(&v[lo], &v[up]);
            ++lo;
            --up;
        }
        const int32_t nl = eql < lo - eql ? eql : lo - eql;
        for(int32_t i = 0; i < nl; ++i) {
            swap16(&v[i], &v[up - i]);
        }
        const int32_t less_n = lo - eql; /* rows with a smaller byte (new m0) */
        const int32_t a = hi - eqr, b = eqr - up;
        const int32_t nr = a < b ? a : b;
        for(int32_t i = 0; i < nr; ++i) {
            swap16(&v[lo + i], &v[hi - i]);
        }
        const int32_t gt_at = hi - (eqr - up) + 1; /* first row with a larger byte (new m1) */
        if(0 < less_n - 1) {
            mkq(R, (uint32_t)less_n, v, d, depth, level - 1);
        }
        if(gt_at < hi) {
            mkq(R, (uint32_t)((int32_t)size - gt_at), v + gt_at, d, depth, level - 1);
        }
        if(gt_at <= less_n) {
            break;
        }
        v += less_n;
        size = (uint32_t)(gt_at - less_n);
        ++d;
    }
}

uint32_t bso_encode_bound(uint32_t size)
{
    const uint32_t blocks = size >> 15;
    return blocks * BSO_CODED + (size - (blocks << 15));
}

uint32_t bso_decode_bound(uint32_t size)
{
    /* as written in the reference: blocks = size >> 15, the coded size is NOT divided by
     * EncodedSize here (blksort.h:411-416); it is an upper bound, decode() divides properly */
    const uint32_t blocks = size >> 15;
    return blocks * N + (size - (blocks << 15));
}

/* encode_internal, blksort.h:444-543: sort(size_, strings, size_) with level 11 (:364-377) */
void bso_encode_block(uint8_t* dst, const uint8_t* src)
{
    uint8_t* twin = (uint8_t*)malloc(2u * N);
    uint16_t* rows = (uint16_t*)malloc(N * sizeof(uint16_t));
    memcpy(twin, src, N);
    memcpy(twin + N, src, N);
    for(uint32_t i = 0; i < N; ++i) {
        rows[i] = (uint16_t)i;
    }
    Rows R = {twin};
    mkq(&R, N, rows, 0, N, 11);
    uint16_t pos = 0;
    for(uint32_t i = 0; i < N; ++i) {
        dst[i] = twin[(uint32_t)rows[i] + N - 1];
        if(rows[i] == 0) {
            pos = (uint16_t)i;
        }
    }
    memcpy(dst + N, &pos, sizeof pos);
    free(rows);
    free(twin);
}

/* decode_internal, blksort.h:545-672: identity ids, counting_sort keyed by the column
 * (:379-402, stable, filled from the back), then the walk from out_id[top]. */
int bso_decode_block(uint8_t* dst, const uint8_t* src)
{
    uint16_t top;
    memcpy(&top, src + N, sizeof top);
    if(top >= N) {
        return -1;
    }
    uint32_t start[256] = {0};
    for(uint32_t i = 0; i < N; ++i) {
        ++start[src[i]];
    }
    uint32_t run = 0;
    for(uint32_t s = 0; s < 256; ++s) { /* the reference keeps inclusive ends and pre-decrements */
        run += start[s];
        start[s] = run;
    }
    uint16_t* next = (uint16_t*)malloc(N * sizeof(uint16_t));
    for(int32_t i = (int32_t)N - 1; i >= 0; --i) {
        next[--start[src[i]]] = (uint16_t)i;
    }
    uint16_t p = next[top];
    for(uint32_t i = 0; i < N; ++i) {
        dst[i] = src[p];
        p = next[p];
    }
    free(next);
    return 0;
}

typedef struct {
    const uint8_t* src;
    uint8_t* dst;
    uint32_t blocks;
    uint32_t* cursor;
    pthread_mutex_t* mu;
    int encode;
    int failed;
} Job;

static void* worker(void* arg)
{
    Job* j = (Job*)arg;
    for(;;) {
        pthread_mutex_lock(j->mu);
        const uint32_t b = (*j->cursor)++;
        pthread_mutex_unlock(j->mu);
        if(b >= j->blocks) {
            break;
        }
        if(j->encode) {
            bso_encode_block(j->dst + (size_t)b * BSO_CODED, j->src + (size_t)b * N);
        } else if(bso_decode_block(j->dst + (size_t)b * N, j->src + (size_t)b * BSO_CODED) != 0) {
            j->failed = 1;
        }
    }
    return NULL;
}

static int run_blocks(int encode, uint32_t blocks, uint8_t* dst, const uint8_t* src, int threads)
{
    uint32_t cursor = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    if(threads < 1) {
        threads = 1;
    }
    if(threads > 256) {
        threads = 256;
    }
    Job jobs[256];
    pthread_t tid[256];
    for(int t = 0; t < threads; ++t) {
        jobs[t] = (Job){src, dst, blocks, &cursor, &mu, encode, 0};
    }
    if(threads == 1) {
        worker(&jobs[0]);
    } else {
        for(int t = 0; t < threads; ++t) {
            pthread_create(&tid[t], NULL, worker, &jobs[t]);
        }
        for(int t = 0; t < threads; ++t) {
            pthread_join(tid[t], NULL);
        }
    }
    int failed = 0;
    for(int t = 0; t < threads; ++t) {
        failed |= jobs[t].failed;
    }
    return failed ? -1 : 0;
}

/* encode, blksort.h:418-428 */
void bso_encode(uint32_t size, uint8_t* dst, const uint8_t* src, int threads)
{
    const uint32_t blocks = size >> 15;
    run_blocks(1, blocks, dst, src, threads);
    memcpy(dst + (size_t)blocks * BSO_CODED, src + (size_t)blocks * N, size - blocks * N);
}

/* decode, blksort.h:430-442 */
int bso_decode(uint32_t size, uint8_t* dst, const uint8_t* src, int threads)
{
    const uint32_t blocks = size / BSO_CODED;
    const int rc = run_blocks(0, blocks, dst, src, threads);
    memcpy(dst + (size_t)blocks * N, src + (size_t)blocks * BSO_CODED, size - blocks * BSO_CODED);
    return rc;
}

int bso_block_is_periodic(const uint8_t* src)
{
    for(uint32_t p = 1; p < N; p <<= 1) { /* a period of a 2^15 block divides it */
        if(memcmp(src, src + p, N - p) == 0) {
            return 1;
        }
    }
    return 0;
}
