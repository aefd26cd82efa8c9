// cppans_b200.h -- drop-in C++ host header for the rANS coder of taqu/cpprcoder (cppans.h).
//
// Same namespace, class, static entry points, argument meaning, buffer convention and
// error behaviour as cppans::rANS in the reference, so that its callers -- run_ans and
// run_ans_simd (test/main.cpp:367-545) -- compile unchanged against this header:
//
//   cppans::rANS::calc_encoded_size(u32)                         cppans.h:71, :492-495
//   cppans::rANS::encode / ::decode            (byte variant)    cppans.h:72-73, :497-564
//   cppans::rANS::encode_simd / ::decode_simd  (8 interleaved)   cppans.h:75-76, :567-649
//
// Kept from the reference: the encoders fill dst from its END -- the coded bytes are
// dst + dst_size - <return value> .. dst + dst_size (cppans.h:515, :529) -- and every
// function returns 0 on failure (destination too small, cppans.h:523-525, :541-543).
// What differs, by design: the coded bytes are a B2RC container (include/b2rc.h) of
// independent 64 KiB blocks, each block's payload being exactly what the reference writes
// for that block, and the work is done by CUDA kernels through libb2rc.so.  A container
// needs a little more room than one stream, so calc_encoded_size returns more than the
// reference's 2 * size + 1032.  There is no CPU coding path: without a CUDA device every
// call returns 0.  New code written against the reference's interface; no reference source
// is reused.
#ifndef INC_CPPANS_B200_H_
#define INC_CPPANS_B200_H_

#include "cpprcoder_b200.h"

namespace cppans
{
using s8 = int8_t;
using s16 = int16_t;
using s32 = int32_t;
using s64 = int64_t;
using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;

class rANS
{
public:
    inline static constexpr u32 MaxSize = 0x7FFFFFFFUL;  // cppans.h:26
    inline static constexpr u32 ProbBits = 14;           // cppans.h:27
    inline static constexpr u32 WordScaleBits = 12;      // cppans.h:31
    inline static constexpr u32 BlockSize = B2RC_DEFAULT_BLOCK;

    static u64 calc_encoded_size(u32 size) { return b2rc_bound(B2RC_MODE_RANS_WORD, size, BlockSize); }

    static u32 encode(u32 dst_size, u8* dst, u32 src_size, const u8* src)
    {
        return encode_as(B2RC_MODE_RANS_BYTE, dst_size, dst, src_size, src);
    }
    static u32 encode_simd(u32 dst_size, u8* dst, u32 src_size, const u8* src)
    {
        return encode_as(B2RC_MODE_RANS_WORD, dst_size, dst, src_size, src);
    }
    // the reference returns the number of coded bytes it consumed (cppans.h:562-563) ...
    static u32 decode(u32 dst_size, u8* dst, u32 src_size, const u8* src)
    {
        u64 nblocks = 0;
        const u64 made = decode_as(B2RC_MODE_RANS_BYTE, dst_size, dst, src_size, src, nblocks);
        if(0 == nblocks) {
            return 0;
        }
        (void)made;
        const u64 coded = static_cast<u64>(src_size) - (B2RC_HEADER_BYTES + 8 * (nblocks + 1)) - 1032 * nblocks;
        return static_cast<u32>(coded);
    }
    // ... and here the number of symbols it produced (cppans.h:648)
    static u32 decode_simd(u32 dst_size, u8* dst, u32 src_size, const u8* src)
    {
        u64 nblocks = 0;
        return static_cast<u32>(decode_as(B2RC_MODE_RANS_WORD, dst_size, dst, src_size, src, nblocks));
    }

private:
    rANS(const rANS&) = delete;
    rANS& operator=(const rANS&) = delete;

    static u32 encode_as(int mode, u32 dst_size, u8* dst, u32 src_size, const u8* src)
    {
        b2rc_ctx* ctx = cpprcoder::detail::context();
        if(nullptr == ctx || nullptr == dst || nullptr == src || 0 == src_size) {
            return 0;
        }
        // code into the front of a scratch buffer, then place the container at the end of dst
        std::vector<u8> out(static_cast<size_t>(b2rc_bound(mode, src_size, BlockSize)));
        u64 made = 0;
        if(B2RC_OK != b2rc_encode(ctx, mode, BlockSize, src, src_size, out.data(), out.size(), &made)) {
            return 0;
        }
        if(made > dst_size || made > MaxSize) {
            return 0;  // cppans.h:523-525
        }
        ::memcpy(dst + dst_size - made, out.data(), static_cast<size_t>(made));
        return static_cast<u32>(made);
    }

    static u64 decode_as(int mode, u32 dst_size, u8* dst, u32 src_size, const u8* src, u64& nblocks)
    {
        nblocks = 0;
        b2rc_ctx* ctx = cpprcoder::detail::context();
        if(nullptr == ctx || nullptr == dst || nullptr == src) {
            return 0;
        }
        int got = -1;
        u64 total = 0, nb = 0;
        if(B2RC_OK != b2rc_peek(src, src_size, &got, nullptr, &total, &nb) || got != mode) {
            return 0;
        }
        if(dst_size < total) {
            return 0;  // cppans.h:541-543
        }
        u64 made = 0;
        if(B2RC_OK != b2rc_decode(ctx, src, src_size, dst, dst_size, &made)) {
            return 0;
        }
        nblocks = nb;
        return made;
    }
};
}  // namespace cppans
#endif
