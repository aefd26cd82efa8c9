// blksort_b200.h -- drop-in C++ host header for the block-sort transform of taqu/cpprcoder (blksort.h).
//
// Same namespace, class, entry points, argument meaning and output format as blksort::BlkSort in the
// reference, so that its callers -- run_blksort, run_zlib_blk, run_zstd_blk (test/main.cpp:791-816,
// :944-1002, :1057-1110) -- compile unchanged against this header:
//
//   blksort::BlkSort::encodeBound / ::decodeBound        blksort.h:86-87, :404-416
//   blksort::BlkSort::encode(size, dst, data)             blksort.h:89, :418-428
//   blksort::BlkSort::decode(size, dst, data)             blksort.h:90, :430-442
//
// The bytes written are the reference's, bit for bit (every full 32 KiB block -> last column + u16 row
// number, the tail copied), including the row number of blocks whose rotations tie.  The work is done by
// CUDA kernels through libb2rc.so (b2rc_blk_encode / b2rc_blk_decode, include/b2rc.h).  There is no CPU
// path: the reference's functions return void, so a failure (no CUDA device, a row number out of range)
// is reported through status() / ok() and leaves dst untouched or partly written -- callers that never
// look behave as they would with a reference that had crashed.  New code written against the
// reference's interface; no reference source is reused.
#ifndef INC_BLKSORT_B200_H_
#define INC_BLKSORT_B200_H_

#include "cpprcoder_b200.h"

namespace blksort
{
class BlkSort
{
public:
    inline static constexpr uint32_t Align = 16;                  // blksort.h:79
    inline static constexpr uint32_t BlockSize = B2RC_BLK_BLOCK;  // blksort.h:80
    inline static constexpr uint32_t BlockShift = 15;             // blksort.h:81
    inline static constexpr uint32_t BlockMask = BlockSize - 1;   // blksort.h:82
    inline static constexpr uint32_t EncodedSize = B2RC_BLK_CODED;  // blksort.h:83

    BlkSort() : status_(B2RC_OK) {}
    ~BlkSort() {}

    static uint32_t encodeBound(uint32_t size) { return static_cast<uint32_t>(b2rc_blk_encode_bound(size)); }
    // as in the reference: the argument shifted, not divided by EncodedSize -- an upper bound (blksort.h:411-416)
    static uint32_t decodeBound(uint32_t size)
    {
        const uint32_t blocks = size >> BlockShift;
        return blocks * BlockSize + (size - (blocks << BlockShift));
    }

    void encode(uint32_t size, uint8_t* dst, const uint8_t* data)
    {
        b2rc_ctx* ctx = cpprcoder::detail::context();
        uint64_t made = 0;
        status_ = ctx ? b2rc_blk_encode(ctx, data, size, dst, b2rc_blk_encode_bound(size), &made) : B2RC_E_CUDA;
    }
    // `data` is not const in the reference (its optional move-to-front stage, compiled out, works in place)
    void decode(uint32_t size, uint8_t* dst, uint8_t* data)
    {
        b2rc_ctx* ctx = cpprcoder::detail::context();
        uint64_t made = 0;
        status_ = ctx ? b2rc_blk_decode(ctx, data, size, dst, b2rc_blk_decoded_size(size), &made) : B2RC_E_CUDA;
    }

    int status() const { return status_; }  // B2RC_OK or the error of the last call (include/b2rc.h)
    bool ok() const { return B2RC_OK == status_; }

private:
    BlkSort(const BlkSort&) = delete;
    BlkSort& operator=(const BlkSort&) = delete;
    int status_;
};
}  // namespace blksort
#endif
