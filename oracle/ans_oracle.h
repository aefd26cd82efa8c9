/*
 * ans_oracle.h -- CPU restatement (plain C) of the static rANS coder in
 * taqu/cpprcoder's cppans.h (cppans::rANS), one independent block at a time.
 * This is the "next" row N3 of SURVEY.md section 8f: the sibling coder of the range
 * coder, carried by the same block framework.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as rc_oracle.h): only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.
 *
 * Parity status: PINNED.  tests/test_ans_oracle.py checks this restatement byte for
 * byte against the unmodified cppans.h compiled into oracle/_ref/libcpprcoder_ref.so
 * (ref_shim.cpp, modes 2 and 3) and against the golden vectors that library generated
 * (tests/golden/golden.json, keys "rans_byte" / "rans_word").  The reference ships no
 * known-answer vectors of its own for this coder.
 *
 * Two variants, as in the reference:
 *   RAO_BYTE  rANS::encode / ::decode            (cppans.h:497-564)  one 32-bit state,
 *             14-bit probabilities, byte renormalisation;
 *   RAO_WORD  rANS::encode_simd / ::decode_simd  (cppans.h:567-649)  eight interleaved
 *             states, 12-bit probabilities, 16-bit renormalisation (the SSE4 variant).
 * Both write  u32 size | u32 cum[257] | coded  (cppans.h:521-527, :598-604).  The
 * reference leaves the payload at the END of the caller's buffer; these functions
 * return it at the START of dst, nothing else differs.
 */
#ifndef ANS_ORACLE_H_
#define ANS_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { RAO_BYTE = 2, RAO_WORD = 3 }; /* container mode numbers (include/b2rc.h) */

#define RAO_HEADER 1032u /* 258 x u32, cppans.h:521 */

/* normalize (cppans.h:138-177): scales the cumulative table to `target` and repairs
 * symbols whose slice vanished by stealing from the narrowest slice wider than 1. */
void rao_normalize(uint32_t freq[256], uint32_t cum[257], uint32_t target);

/* Builds the normalised model of a block the way both encoders do
 * (count + cumulative + normalize, cppans.h:504-508 / :574-578). */
void rao_model(const uint8_t* src, uint32_t n, uint32_t target, uint32_t freq[256], uint32_t cum[257]);

/* Returns payload bytes (written at dst[0..)), or -1 when cap is too small. n > 0. */
long rao_encode(int variant, const uint8_t* src, uint32_t n, uint8_t* dst, size_t cap);
/* Returns the number of symbols written (the header's size field), or -1 on a short or
 * inconsistent payload. */
long rao_decode(int variant, const uint8_t* src, size_t n, uint8_t* dst, size_t cap);

/* Payload bound for an n-byte block.  The word variant emits 16 bits per symbol when
 * one symbol owns the whole 12-bit scale (the bound at cppans.h:357 wraps to zero), so 2n it is. */
size_t rao_slot_bytes(uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
