/*
 * b2rc.h -- C ABI of the B200-native block range coder (libb2rc.so).
 *
 * This is the drop-in boundary for the range-coder hot path of taqu/cpprcoder.
 * The reference has no FFI: its boundary is the C++ class API of cpprcoder.h
 *     RangeEncoder<T>::encode / ::decode                 (cpprcoder.h:336-337)
 *     AdaptiveRangeEncoder<T>::initialize / ::encode     (cpprcoder.h:636-638)
 *     AdaptiveRangeDecoder<T>::initialize / ::decode     (cpprcoder.h:819-820)
 * cpprcoder_b200/include/cpprcoder_b200.h re-creates those classes on top of the
 * functions below (INTEGRATION.md shows the binding).  Plain pointers and sizes
 * only; no CUDA or torch types appear in the signatures (a CUDA stream is passed
 * as void*).  There is NO CPU fallback: every entry point that does work needs a
 * CUDA device and returns B2RC_E_CUDA without one.
 *
 * Framing ("B2RC" container, little endian; ours -- the reference codes one stream):
 *     0   u32  magic 'B','2','R','C'
 *     4   u16  version (1)          6  u16 mode (0 static, 1 adaptive)
 *     8   u32  block_size          12  u32 flags (0, or 1 | (seg_syms / 64) << 8: restart table)
 *     16  u64  total_uncompressed  24  u64 nblocks
 *     32  u64  offsets[nblocks+1]  relative to the payload base, offsets[0] = 0
 *     32 + 8*(nblocks+1)           payloads, back to back
 *     [at the next 4-byte boundary, when flags say so: the restart table (static range coder and
 *      byte rANS), 3 x u32 per point, see b2rc_k_encode_blocks_r]
 * Payload b is byte-for-byte what the reference encoder emits for block b alone:
 * static  = u32 LE size, 256 x u16 LE frequencies, coded bytes  (cpprcoder.h:386-457)
 * adaptive = u32 LE size, coded bytes                            (cpprcoder.h:689-762)
 * The last block may be short; an empty input has zero blocks.
 */
#ifndef B2RC_H_
#define B2RC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2RC_MODE_STATIC 0   /* RangeEncoder<T>          cpprcoder.h:321-619 */
#define B2RC_MODE_ADAPTIVE 1 /* AdaptiveRangeEncoder/Decoder<T>  cpprcoder.h:626-940 */
/* The sibling static rANS coder of the reference (cppans.h), on the same container:
 * payload = u32 LE size, 257 x u32 LE normalised cumulative counts, coded bytes -- what
 * cppans::rANS leaves at the END of its dst buffer (cppans.h:521-529, :598-605). */
#define B2RC_MODE_RANS_BYTE 2 /* rANS::encode / ::decode            cppans.h:497-564 */
#define B2RC_MODE_RANS_WORD 3 /* rANS::encode_simd / ::decode_simd  cppans.h:567-649 */

#define B2RC_DEFAULT_BLOCK 65536u
#define B2RC_MIN_BLOCK 64u          /* block_size must be a multiple of 64 ... */
#define B2RC_MAX_BLOCK (1u << 23)   /* ... and small enough that neither model rescales by size
                                       (cpprcoder.h:561, :1138) */
#define B2RC_HEADER_BYTES 32u
#define B2RC_DEFAULT_RESTART_SYMS 8192u /* restart points of the byte rANS coder, see b2rc_k_encode_blocks_r */
/* ... of the static range coder: sixteen chains per 64 KiB block, decoded by CTAs of eight warps: 1 GiB in 3.80 ms
 * against 3.97 ms at 8192 (180 instead of 84 bytes per block: +0.15 % size, counted in every ratio) */
#define B2RC_DEFAULT_STATIC_RESTART_SYMS 4096u
#define B2RC_MIN_RESTART_SYMS 1024u     /* shortest segment a context accepts (env B2RC_RESTART_SYMS; 0 = none) */
/* The adaptive coder's restart points carry the model as well (256 u16 symbol counts): 524 bytes each.
 * Default every 21888 symbols: a 64 KiB block is three chains for the decoder and carries two points,
 * 1048 B (+1.6 % of the input size, counted in every ratio the bench prints).  Three, because the
 * decoder keeps 11 warps of 32 blocks per SM (16 KiB of model each): 1628 on a B200, and a 1 GiB stream
 * is 512 warps per segment -- four segments would need a second, mostly empty, wave (measured: 10.4 ms
 * against 6.9).  env B2RC_ADAPTIVE_RESTART_SYMS, 0 = none (the reference's stream size). */
#define B2RC_DEFAULT_ADAPTIVE_RESTART_SYMS 21888u
#define B2RC_MIN_ADAPTIVE_RESTART_SYMS 4096u
#define B2RC_ADAPTIVE_RESTART_WORDS 131u /* u32 per point: bytes shifted, low, range, then 256 x u16 counts */
/* Blocks above 65536 bytes: the counts take 32 bits (259 words = 1036 bytes a point) and the decoder's tree 32 KiB
 * per warp, six warps per SM = 888 on a B200.  Default every 43712 symbols: 1 GiB in blocks of 128 KiB .. 1 MiB is
 * then 768 warps, one wave (+2.4 % size).  env B2RC_ADAPTIVE_RESTART_SYMS_WIDE, 0 = none. */
#define B2RC_DEFAULT_ADAPTIVE_RESTART_SYMS_WIDE 43712u
#define B2RC_ADAPTIVE_RESTART_WORDS_WIDE 259u

/* Status codes.  0 = the reference's `true` / Status_Success (cpprcoder.h:112-117);
 * negatives map to `false` / Status_Error in the C++ header. */
#define B2RC_OK 0
#define B2RC_E_ARG (-1)      /* null pointer, bad mode, bad block size, misaligned device pointer */
#define B2RC_E_DST_SMALL (-2) /* dst capacity too small; *out_n holds the size needed (when known) */
#define B2RC_E_CORRUPT (-3)  /* container or payload fails validation */
#define B2RC_E_CUDA (-4)     /* CUDA runtime error, or no device */
#define B2RC_E_EXPAND (-5)   /* a block's payload outgrew its staging slot (pathological input);
                                the reference fails the same way when its MemoryStream is full
                                (cpprcoder.h:409-411, :1047-1051) */
#define B2RC_E_NOMEM (-6)
#define B2RC_E_INTERNAL (-7) /* the library contradicted itself (a bug); nothing usable was written */

typedef struct b2rc_ctx b2rc_ctx;

/* One context per host thread and device.  Owns the device scratch (staging slots,
 * per-block sizes, frequency tables) and a stream of its own for the host-pointer calls. */
int b2rc_ctx_create(int device, b2rc_ctx** out);
void b2rc_ctx_destroy(b2rc_ctx* ctx);
/* One context over several devices of ONE process ("multi-GPU = single process, one
 * stream per device"): the host-pointer calls -- b2rc_encode, b2rc_decode and the *_staged pair, i.e.
 * everything the C++ drop-in classes use -- shard the blocks by contiguous range over the devices
 * (SURVEY.md 8e), one host thread per device.  The payload sizes meet on the host, so this path
 * needs no collective at all; the container is byte for byte the one a single device writes.
 * The device-pointer calls and the kernel doors of such a context run on devices[0].
 * A device may be listed more than once (two streams of work on one GPU; also how the path is
 * tested on a one-GPU box).  1 <= ndev <= 16. */
int b2rc_ctx_create_multi(const int* devices, int ndev, b2rc_ctx** out);
int b2rc_ctx_devices(const b2rc_ctx* ctx);
/* Restart points are written every B2RC_DEFAULT_RESTART_SYMS symbols for a stream that
 * fills the GPU; a stream of fewer blocks gets them closer together (down to B2RC_MIN_RESTART_SYMS), so that its
 * decoder still has a warp per scheduler slot.  b2rc_restart_for: the spacing this context would write a stream of
 * `nblocks` blocks with (0: no restart table).  b2rc_ctx_force_restart: one spacing for every stream from now on
 * (a multiple of 64, B2RC_MIN_RESTART_SYMS ..; 0 = automatic again) -- what a caller that stitches containers of
 * several contexts sets (cpprcoder_b200/dist.py); env B2RC_RESTART_SYMS does the same at context creation. */
uint32_t b2rc_restart_for(const b2rc_ctx* ctx, int mode, uint32_t block_size, uint64_t nblocks);
int b2rc_ctx_force_restart(b2rc_ctx* ctx, uint32_t seg_syms);
const char* b2rc_strerror(int code);
/* Last CUDA error text seen by this context ("" if none). */
const char* b2rc_last_cuda_error(const b2rc_ctx* ctx);

/* Sizes.  b2rc_bound: capacity that always suffices for b2rc_encode*(n bytes).
 * b2rc_slot_bytes: staging slot (and per-block payload bound) for a block of n bytes. */
uint64_t b2rc_bound(int mode, uint64_t n, uint32_t block_size);
uint64_t b2rc_slot_bytes(uint32_t n);
/* The same for any mode (the rANS payload bound is 2n + 1064: cppans.h:492-495 and the
 * word coder's wrap at :357). */
uint64_t b2rc_slot_bytes_for(int mode, uint32_t n);
uint64_t b2rc_nblocks(uint64_t n, uint32_t block_size);

/* ---- whole-container calls, HOST pointers (what the C++ drop-in classes call) ----
 * replaces RangeEncoder<T>::encode (cpprcoder.h:375) and
 * AdaptiveRangeEncoder<T>::initialize+encode (cpprcoder.h:678, :697). */
int b2rc_encode(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* src, uint64_t n, uint8_t* dst,
                uint64_t dst_cap, uint64_t* out_n);
/* replaces RangeEncoder<T>::decode (cpprcoder.h:460) and AdaptiveRangeDecoder<T>::decode
 * (cpprcoder.h:872).  mode and block size come from the container header. */
int b2rc_decode(b2rc_ctx* ctx, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* out_n);
/* Parses a container header held in host memory. */
int b2rc_peek(const uint8_t* src, uint64_t n, int* mode, uint32_t* block_size, uint64_t* total, uint64_t* nblocks);
/* How long is the container that starts at `prefix`?  *need = its total length once `have` bytes of it
 * tell (header, index, restart table and all); otherwise the shortest prefix that tells more.
 * B2RC_E_CORRUPT for something that is not a container.  (Streaming callers: AdaptiveRangeDecoder.) */
int b2rc_container_bytes(const uint8_t* prefix, uint64_t have, uint64_t* need);
/* b2rc_peek + the whole index, on the host, before anything is allocated for the output: offsets
 * monotone and inside the container, every payload at least as long as its mode's header, the
 * restart table (if any) inside the container.  *total = bytes b2rc_decode will write.  A container
 * of a kilobyte cannot make a caller allocate gigabytes: total <= nblocks * block_size, and
 * nblocks is bounded by the index that is really there. */
int b2rc_check(const uint8_t* src, uint64_t n, uint64_t* total);
/* The same two calls with the RESULT left in host memory that the context owns (pinned, never
 * initialised, reused by the next staged call on this context): what the C++ drop-in classes use
 * when they cannot code straight into the caller's stream.  *out stays valid until the next
 * b2rc_*_staged call or b2rc_ctx_destroy. */
int b2rc_encode_staged(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* src, uint64_t n,
                       const uint8_t** out, uint64_t* out_n);
int b2rc_decode_staged(b2rc_ctx* ctx, const uint8_t* src, uint64_t n, const uint8_t** out, uint64_t* out_n);
/* Page-locked host memory for callers that want the copies of the host-pointer calls at full PCIe
 * speed (with pageable memory the driver stages every copy through its own buffers).
 * cpprcoder::PinnedStream in the C++ header sits on these.  NULL / B2RC_E_NOMEM without a device. */
int b2rc_host_alloc(uint64_t bytes, void** out);
void b2rc_host_free(void* p);
/* memcpy over a few threads for large sizes (the staged results above are copied out with it). */
void b2rc_host_copy(void* dst, const void* src, uint64_t n);

/* ---- whole-container calls, DEVICE pointers (what bench.py times as `value`) ----
 * d_src / d_dst must be 16-byte aligned.  `cuda_stream` is a cudaStream_t (NULL = the
 * legacy default stream, as everywhere in CUDA).  The calls return after the stream has drained. */
int b2rc_encode_device(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_src, uint64_t n,
                       uint8_t* d_dst, uint64_t dst_cap, uint64_t* out_n, void* cuda_stream);
int b2rc_decode_device(b2rc_ctx* ctx, const uint8_t* d_src, uint64_t n, uint8_t* d_dst, uint64_t dst_cap,
                       uint64_t* out_n, void* cuda_stream);

/* ---- per-kernel entry points (parity tests, ncu, multi-GPU sharding) ----
 * All pointers are device pointers; launches are asynchronous on `cuda_stream`
 * unless stated.  `d_err` is one int the kernels OR error bits into (0 = clean).
 *
 * K1  b2rc_k_histogram: per-block symbol counts with the reference's scaling rule,
 *     RangeEncoder::count (cpprcoder.h:543-571), including its order-dependent halving for
 *     blocks above 65536 bytes.  d_freq16[b*256 + s] = u16 frequency of symbol s in block b. */
int b2rc_k_histogram(b2rc_ctx* ctx, const uint8_t* d_src, uint64_t n, uint32_t block_size, uint16_t* d_freq16,
                     void* cuda_stream);
/* K2  b2rc_k_encode_blocks: block b's payload -> d_slots + b*slot_stride, its length ->
 *     d_sizes[b].  Static mode with block_size <= 65536 needs d_freq16 from K1; larger
 *     static blocks take it when given and otherwise count inside the kernel (NULL, slower);
 *     the rANS modes ignore it (their
 *     model kernel, cppans.h:504-508, runs as part of this call).
 *     slot_stride >= b2rc_slot_bytes_for(mode, block_size), multiple of 16. */
int b2rc_k_encode_blocks(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_src, uint64_t n,
                         const uint16_t* d_freq16, uint8_t* d_slots, uint64_t slot_stride, uint32_t* d_sizes,
                         int* d_err, void* cuda_stream);
/* Restart points of the static coder: while encoding, K2 can record, before every seg_syms-th
 * symbol of every block, what a decoder needs to start there -- {bytes shifted out so far, the
 * encoder's low, range}, 3 x u32 per point, b2rc_restart_records(block, seg) points per block
 * (0xFFFFFFFF in the first word: the block ends before that point).  The payloads do not change;
 * a block becomes several independent chains for K3.  seg_syms: a multiple of 64 below block_size.
 * B2RC_MODE_RANS_BYTE records the same way (its encoder walks the block backwards, so a point is
 * taken AFTER everything from that symbol on is coded): {coded bytes emitted so far = bytes still
 * ahead of a decoder that stands at that symbol, the 32-bit state x, 0}.  Other modes: B2RC_E_ARG.
 * d_restart: nblocks * records * 3 u32, 4-byte aligned; NULL = the calls above. */
uint32_t b2rc_restart_records(uint32_t block_size, uint32_t seg_syms);
int b2rc_k_encode_blocks_r(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_src, uint64_t n,
                           const uint16_t* d_freq16, uint8_t* d_slots, uint64_t slot_stride, uint32_t* d_sizes,
                           uint32_t* d_restart, uint32_t seg_syms, int* d_err, void* cuda_stream);
int b2rc_k_decode_blocks_r(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_payload,
                           uint64_t payload_len, const uint64_t* d_offsets, uint64_t nblocks, uint8_t* d_dst, uint64_t n,
                           const uint32_t* d_restart, uint32_t seg_syms, int* d_err, void* cuda_stream);
/* K4  b2rc_k_scan: d_offsets[0..nblocks] = exclusive prefix of d_sizes (u64).
 *     b2rc_k_compact: payload b -> d_payload + d_offsets[b] (payload_cap bytes available). */
int b2rc_k_scan(b2rc_ctx* ctx, const uint32_t* d_sizes, uint64_t nblocks, uint64_t* d_offsets, void* cuda_stream);
int b2rc_k_compact(b2rc_ctx* ctx, const uint8_t* d_slots, uint64_t slot_stride, const uint32_t* d_sizes,
                   const uint64_t* d_offsets, uint64_t nblocks, uint8_t* d_payload, uint64_t payload_cap, int* d_err,
                   void* cuda_stream);
/* The same for any mode: rANS slots hold the header at their start and the coded bytes at
 * their end (the reference codes backwards from the end of dst, cppans.h:515, :591). */
int b2rc_k_compact_for(b2rc_ctx* ctx, int mode, const uint8_t* d_slots, uint64_t slot_stride, const uint32_t* d_sizes,
                       const uint64_t* d_offsets, uint64_t nblocks, uint8_t* d_payload, uint64_t payload_cap,
                       int* d_err, void* cuda_stream);
/* K3  b2rc_k_decode_blocks: payload b = d_payload[d_offsets[b] .. d_offsets[b+1]) -> block b of d_dst
 *     (n bytes in all, 16-byte aligned).  Offsets beyond payload_len mark the block corrupt. */
int b2rc_k_decode_blocks(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_payload,
                         uint64_t payload_len, const uint64_t* d_offsets, uint64_t nblocks, uint8_t* d_dst, uint64_t n,
                         int* d_err, void* cuda_stream);

/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
uint64_t b2rc_launch_count(const b2rc_ctx* ctx);
/* Per-kernel device timing.  b2rc_profile(ctx, 1) makes every launch record CUDA events on
 * its stream; b2rc_kernel_ms returns the duration of the LAST launch of kernel `which`
 * (-1 if none since profiling was switched on).  Measurement only; off by default. */
#define B2RC_K_HISTOGRAM 0
#define B2RC_K_ENCODE 1
#define B2RC_K_SCAN 2
#define B2RC_K_COMPACT 3
#define B2RC_K_DECODE 4
#define B2RC_K_BLK_FORWARD 5
#define B2RC_K_BLK_INVERSE 6
#define B2RC_K_RANGES 7 /* static encode: the range-only pass (sizes up front) */
#define B2RC_K_SEAMS 8  /* static encode: the seams between segments */
#define B2RC_K_COUNT 9
int b2rc_profile(b2rc_ctx* ctx, int enable);
int b2rc_kernel_ms(b2rc_ctx* ctx, int which, float* ms);
/* ------------------------------------------------------------------ block sort --
 * The reference's block-sort transform (blksort::BlkSort, blksort.h; SURVEY.md section 8f row
 * N4), the pre-transform its own pipelines run in front of a coder (test/main.cpp:944-1002).
 * Output format = the reference's, byte for byte: every FULL 32 KiB block of the input becomes
 * 32 770 bytes -- the last column of its sorted cyclic rotations, then the u16 (little endian)
 * row of the unrotated block -- and the bytes behind the last full block are copied as they
 * are (blksort.h:418-442).  No container, no index: sizes are a function of n alone.
 *   b2rc_blk_encode_bound   replaces BlkSort::encodeBound (blksort.h:404-409), 64-bit;
 *   b2rc_blk_decoded_size   the exact size decode() writes for n coded bytes (the reference's
 *                           decodeBound, blksort.h:411-416, returns n itself: an upper bound);
 *   b2rc_blk_encode[_device] replaces BlkSort::encode (blksort.h:418-428);
 *   b2rc_blk_decode[_device] replaces BlkSort::decode (blksort.h:430-442); B2RC_E_CORRUPT when a
 *                           block's row number is >= 32 768 (the reference reads out of bounds).
 * Device variants: d_src / d_dst 16-byte aligned, work queued on `cuda_stream`, the call returns
 * after the stream has drained (it reads the error word).  Host variants pipeline chunks of
 * blocks over copy and kernel streams.  *out_n = bytes written (or needed, on DST_SMALL). */
#define B2RC_BLK_BLOCK 32768u
#define B2RC_BLK_CODED 32770u
uint64_t b2rc_blk_encode_bound(uint64_t n);
uint64_t b2rc_blk_decoded_size(uint64_t n);
int b2rc_blk_encode_device(b2rc_ctx* ctx, const uint8_t* d_src, uint64_t n, uint8_t* d_dst, uint64_t dst_cap,
                           uint64_t* out_n, void* cuda_stream);
int b2rc_blk_decode_device(b2rc_ctx* ctx, const uint8_t* d_src, uint64_t n, uint8_t* d_dst, uint64_t dst_cap,
                           uint64_t* out_n, void* cuda_stream);
int b2rc_blk_encode(b2rc_ctx* ctx, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* out_n);
int b2rc_blk_decode(b2rc_ctx* ctx, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* out_n);
/* Block sort in front of a coder in ONE call, everything on the device -- the shape of the reference's
 * run_zlib_blk / run_zstd_blk pipelines (test/main.cpp:944-1110: BlkSort::encode, then a compressor over its
 * output) with this library's coders where zlib / zstd stand.  Output = 16 bytes {'B','2','B','S', u32 0,
 * u64 original size} and behind them the B2RC container of BlkSort::encode's output (b2rc_encode_device, `mode`
 * and `block_size` as there); b2rc_blkrc_decode_device undoes both.  d_src / d_dst 16-byte aligned; the calls
 * return after the stream has drained. */
uint64_t b2rc_blkrc_bound(int mode, uint64_t n, uint32_t block_size);
int b2rc_blkrc_encode_device(b2rc_ctx* ctx, int mode, uint32_t block_size, const uint8_t* d_src, uint64_t n,
                             uint8_t* d_dst, uint64_t dst_cap, uint64_t* out_n, void* cuda_stream);
int b2rc_blkrc_decode_device(b2rc_ctx* ctx, const uint8_t* d_src, uint64_t n, uint8_t* d_dst, uint64_t dst_cap,
                             uint64_t* out_n, void* cuda_stream);
/* Doubling rounds the forward kernel took per block in the last b2rc_blk_encode_device call of
 * this context (bit 31: the block has a period; bit 30: one repeated byte; bits 8..15: how many of the rounds were short ones); copies min(cap, blocks) words to host memory.
 * Measurement / tests only. */
int b2rc_blk_rounds(b2rc_ctx* ctx, uint32_t* rounds, uint64_t cap, uint64_t* nblocks);

/* "sm_100a" etc.: the architecture the kernels were compiled for. */
const char* b2rc_build_arch(void);

#ifdef __cplusplus
}
#endif
#endif
