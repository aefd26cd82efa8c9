/*
 * blk_oracle.h -- CPU restatement (plain C) of the block-sort transform in
 * taqu/cpprcoder's blksort.h (blksort::BlkSort), SURVEY.md section 8f row N4: the
 * pre-transform the reference's own pipelines put in front of a coder
 * (run_zlib_blk / run_zstd_blk, test/main.cpp:944-1002, :1057-1110).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as rc_oracle.h): only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.
 *
 * Parity status: PINNED.  tests/test_blk_oracle.py checks this restatement byte for
 * byte against the unmodified blksort.h compiled into oracle/_ref/libcpprcoder_ref.so
 * (ref_shim.cpp, ref_blk_*) and against golden vectors that library generated
 * (tests/golden/golden_blk.json).  The reference ships no known answers for it.
 *
 * The transform (blksort.h:418-455): the input is cut into 32 KiB blocks; each FULL
 * block becomes 32 770 bytes -- the last column of its sorted cyclic rotations
 * followed by the u16 (host order, little endian here) row number of the unrotated
 * block; whatever is left behind the last full block is copied as it is.
 * BLOCKSORT_MTF is 0 in the reference (blksort.h:55): no move-to-front stage.
 *
 * Rotations that are EQUAL (a block with a period) have equal last bytes, so the
 * column never depends on how the sort breaks ties -- but the row number does.  The
 * reference's sort is an unstable multikey quicksort; bso_encode_block replays it
 * step for step (same pivots, same swaps, same heap / insertion fallbacks), which is
 * the only way to name the row it reports for such a block.
 */
#ifndef BLK_ORACLE_H_
#define BLK_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSO_BLOCK 32768u   /* BlkSort::BlockSize, blksort.h:80 */
#define BSO_CODED 32770u   /* BlkSort::EncodedSize, blksort.h:83 */

uint32_t bso_encode_bound(uint32_t size); /* blksort.h:404-409 */
uint32_t bso_decode_bound(uint32_t size); /* blksort.h:411-416 */

/* One full block: dst gets BSO_CODED bytes. blksort.h:444-543. */
void bso_encode_block(uint8_t* dst, const uint8_t* src);
/* One coded block back: dst gets BSO_BLOCK bytes; -1 when the row number is out of range
 * (the reference reads out of bounds there). blksort.h:545-672. */
int bso_decode_block(uint8_t* dst, const uint8_t* src);

/* Whole buffers, blksort.h:418-442.  `threads` > 1 spreads the blocks over that many threads. */
void bso_encode(uint32_t size, uint8_t* dst, const uint8_t* src, int threads);
int bso_decode(uint32_t size, uint8_t* dst, const uint8_t* src, int threads);

/* 1 when the block has a period shorter than itself (some rotations are equal). */
int bso_block_is_periodic(const uint8_t* src);

#ifdef __cplusplus
}
#endif
#endif
