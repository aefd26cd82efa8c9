// ref_shim.cpp -- thin extern "C" door onto the UNMODIFIED reference header.
//
// TEST / BASELINE INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile from the
// reference sources where they lie (CPPRCODER_H, default /root/reference/cpprcoder.h;
// nothing is copied into this repository) into oracle/_ref/libcpprcoder_ref.so,
// which is git-ignored and travels to the GPU box as a built artefact.  It is
// used to (a) validate oracle/rc_oracle.c, (b) generate tests/golden/, and
// (c) time the reference's own CPU coder for bench.py's cpu_baseline and
// `--impl reference` legs.  The product library never links it.
//
// Per-block calls follow how the reference harness drives the coder:
// run_rangecoder (test/main.cpp:254-301) and run_adaptive (test/main.cpp:305-363).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define CPPRCODER_IMPLEMENTATION
#include CPPRCODER_H

// The sibling rANS coder (cppans.h), modes 2 and 3 below.
// The header does not compile with g++ as written: its restrict macro is defined for
// MSVC and clang only (it tests `__gnuc__`, cppans.h:92-99), and `static alignas(16)
// const` (cppans.h:445) is ill-formed for GCC.  Two macros around the include repair
// both; the header is used as it lies.
#ifdef CPPANS_H
#define CPPANS_RESTRICT __restrict
#define alignas(x) __attribute__((aligned(x)))
#define CPPANS_IMPLEMENTATION
#include CPPANS_H
#undef alignas
#endif

// The block-sort transform (blksort.h), ref_blk_* below.
#ifdef BLKSORT_H
#define BLKSORT_IMPLEMENTATION
#include BLKSORT_H
#endif

namespace
{
using namespace cpprcoder;

// MemoryStream::writeByte never grows (cpprcoder.h:1047-1054), so the stream is
// pre-sized to the caller's slot, as the harness pre-sizes it to the input size.
#ifdef CPPANS_H
// rANS::encode / encode_simd write the payload at the END of the buffer they are given
// (cppans.h:515-529, :591-605); it is moved to the front of dst here.
long ans_encode_one(int mode, const u8* src, u32 n, u8* dst, size_t cap, std::vector<u8>& buf)
{
    if(n == 0) {
        return -1;  // asserted against by the reference (cppans.h:502, :572)
    }
    buf.resize(static_cast<size_t>(n) * 2 + 1032 + 64);
    const u32 got = mode == 3 ? cppans::rANS::encode_simd(static_cast<u32>(buf.size()), buf.data(), n, src)
                              : cppans::rANS::encode(static_cast<u32>(buf.size()), buf.data(), n, src);
    if(got == 0 || got > cap) {
        return -1;
    }
    memcpy(dst, buf.data() + buf.size() - got, got);
    return static_cast<long>(got);
}

// decode_simd reads up to 6 bytes past the coded words (cppans.h:476-478), so the payload
// is copied into a padded buffer first.  Returns the symbol count of the header.
long ans_decode_one(int mode, const u8* src, size_t n, u8* dst, size_t cap, std::vector<u8>& buf)
{
    if(n < 1032 + 4) {
        return -1;
    }
    u32 want;
    memcpy(&want, src, 4);
    if(want > cap) {
        return -1;
    }
    buf.assign(src, src + n);
    buf.resize(n + 16, 0);
    const u32 got = mode == 3 ? cppans::rANS::decode_simd(static_cast<u32>(cap), dst, static_cast<u32>(n), buf.data())
                              : cppans::rANS::decode(static_cast<u32>(cap), dst, static_cast<u32>(n), buf.data());
    return got == 0 ? -1 : static_cast<long>(want);
}
#endif

long encode_one(int mode, const u8* src, u32 n, u8* dst, size_t cap, MemoryStream& scratch)
{
#ifdef CPPANS_H
    if(mode == 2 || mode == 3) {
        thread_local std::vector<u8> buf;
        return ans_encode_one(mode, src, n, dst, cap, buf);
    }
#endif
    scratch.resize(0);
    if(mode == 0) {
        RangeEncoder<> coder;
        if(!coder.encode(scratch, n, src)) {
            return -1;
        }
    } else {
        AdaptiveRangeEncoder<> coder;
        if(!coder.initialize(scratch, n)) {
            return -1;
        }
        if(n == 0) {
            // encode(0, ...) reaches finish() immediately (cpprcoder.h:714-717)
        }
        Result r = coder.encode(static_cast<s32>(n), src);
        if(r.status_ != Status_Success) {
            return -1;
        }
    }
    const size_t got = static_cast<size_t>(scratch.size());
    if(got > cap) {
        return -1;
    }
    memcpy(dst, scratch.get(), got);
    return static_cast<long>(got);
}

long decode_one(int mode, const u8* src, size_t n, u8* dst, size_t cap, MemoryStream& scratch)
{
#ifdef CPPANS_H
    if(mode == 2 || mode == 3) {
        thread_local std::vector<u8> buf;
        return ans_decode_one(mode, src, n, dst, cap, buf);
    }
#endif
    scratch.resize(0);
    if(mode == 0) {
        RangeEncoder<> coder;
        if(!coder.decode(scratch, static_cast<u32>(n), src)) {
            return -1;
        }
    } else {
        AdaptiveRangeDecoder<> coder;
        if(!coder.initialize(scratch)) {
            return -1;
        }
        Result r = coder.decode(static_cast<s32>(n), src);
        if(r.status_ != Status_Success) {
            return -1;
        }
    }
    const size_t got = static_cast<size_t>(scratch.size());
    if(got > cap) {
        return -1;
    }
    memcpy(dst, scratch.get(), got);
    return static_cast<long>(got);
}

size_t slot_for(u32 n)
{
    size_t s = static_cast<size_t>(n) + n / 8 + 1024;
    return (s + 127) & ~static_cast<size_t>(127);
}
} // namespace

extern "C" {

long ref_encode(int mode, const uint8_t* src, uint32_t n, uint8_t* dst, size_t cap)
{
    cpprcoder::MemoryStream scratch(static_cast<cpprcoder::s32>(cap));
    static const uint8_t nothing = 0;
    return encode_one(mode, src ? src : &nothing, n, dst, cap, scratch);
}

long ref_decode(int mode, const uint8_t* src, size_t n, uint8_t* dst, size_t cap)
{
    cpprcoder::MemoryStream scratch(static_cast<cpprcoder::s32>(cap ? cap : 16));
    return decode_one(mode, src, n, dst, cap, scratch);
}

// Block-parallel std::thread drivers: workers pull block
// indices from an atomic counter; each owns one pre-allocated MemoryStream so no
// allocation happens inside the timed loop.
int ref_encode_blocks(int mode, const uint8_t* src, uint64_t n, uint32_t block, uint8_t* slots, uint64_t slot_stride,
                      uint32_t* sizes, int threads)
{
    if(block == 0) {
        return -1;
    }
    const uint64_t nblocks = (n + block - 1) / block;
    std::atomic<uint64_t> next(0);
    std::atomic<int> failed(0);
    auto body = [&]() {
        cpprcoder::MemoryStream scratch(static_cast<cpprcoder::s32>(slot_for(block)));
        for(;;) {
            const uint64_t b = next.fetch_add(1);
            if(b >= nblocks) {
                break;
            }
            const uint64_t at = b * static_cast<uint64_t>(block);
            const uint32_t len = static_cast<uint32_t>((n - at < block) ? (n - at) : block);
            const long r = encode_one(mode, src + at, len, slots + b * slot_stride, slot_stride, scratch);
            if(r < 0) {
                failed = 1;
            } else {
                sizes[b] = static_cast<uint32_t>(r);
            }
        }
    };
    if(threads <= 1) {
        body();
    } else {
        std::vector<std::thread> pool;
        for(int i = 0; i < threads; ++i) {
            pool.emplace_back(body);
        }
        for(auto& t : pool) {
            t.join();
        }
    }
    return failed ? -1 : 0;
}

int ref_decode_blocks(int mode, const uint8_t* stream, const uint64_t* offsets, uint64_t nblocks, uint32_t block,
                      uint8_t* dst, uint64_t n, int threads)
{
    if(block == 0 || nblocks != (n + block - 1) / block) {
        return -1;
    }
    std::atomic<uint64_t> next(0);
    std::atomic<int> failed(0);
    auto body = [&]() {
        cpprcoder::MemoryStream scratch(static_cast<cpprcoder::s32>(block + 16));
        for(;;) {
            const uint64_t b = next.fetch_add(1);
            if(b >= nblocks) {
                break;
            }
            const uint64_t at = b * static_cast<uint64_t>(block);
            const uint32_t len = static_cast<uint32_t>((n - at < block) ? (n - at) : block);
            const long r = decode_one(mode, stream + offsets[b], static_cast<size_t>(offsets[b + 1] - offsets[b]),
                                      dst + at, len, scratch);
            if(r != static_cast<long>(len)) {
                failed = 1;
            }
        }
    };
    if(threads <= 1) {
        body();
    } else {
        std::vector<std::thread> pool;
        for(int i = 0; i < threads; ++i) {
            pool.emplace_back(body);
        }
        for(auto& t : pool) {
            t.join();
        }
    }
    return failed ? -1 : 0;
}

int ref_hardware_threads(void)
{
    const unsigned h = std::thread::hardware_concurrency();
    return h ? static_cast<int>(h) : 1;
}

#ifdef BLKSORT_H
// BlkSort::encode / ::decode over whole buffers (blksort.h:418-442).  With threads > 1 the full
// blocks are spread over that many BlkSort objects -- the block-parallel CPU baseline.
uint32_t ref_blk_encode_bound(uint32_t size)
{
    return blksort::BlkSort::encodeBound(size);
}
uint32_t ref_blk_decode_bound(uint32_t size)
{
    return blksort::BlkSort::decodeBound(size);
}
static void blk_run(bool enc, uint32_t blocks, uint8_t* dst, const uint8_t* src, int threads)
{
    std::atomic<uint32_t> next(0);
    auto body = [&]() {
        blksort::BlkSort bs;
        for(;;) {
            const uint32_t b = next.fetch_add(1);
            if(b >= blocks) {
                break;
            }
            if(enc) {
                bs.encode(blksort::BlkSort::BlockSize, dst + static_cast<size_t>(b) * blksort::BlkSort::EncodedSize,
                          src + static_cast<size_t>(b) * blksort::BlkSort::BlockSize);
            } else {
                bs.decode(blksort::BlkSort::EncodedSize, dst + static_cast<size_t>(b) * blksort::BlkSort::BlockSize,
                          const_cast<uint8_t*>(src) + static_cast<size_t>(b) * blksort::BlkSort::EncodedSize);
            }
        }
    };
    if(threads <= 1) {
        body();
        return;
    }
    std::vector<std::thread> pool;
    for(int i = 0; i < threads; ++i) {
        pool.emplace_back(body);
    }
    for(auto& t : pool) {
        t.join();
    }
}
void ref_blk_encode(uint32_t size, uint8_t* dst, const uint8_t* src, int threads)
{
    if(threads <= 1) {
        blksort::BlkSort bs;
        bs.encode(size, dst, src);
        return;
    }
    const uint32_t blocks = size >> blksort::BlkSort::BlockShift;
    blk_run(true, blocks, dst, src, threads);
    memcpy(dst + static_cast<size_t>(blocks) * blksort::BlkSort::EncodedSize,
           src + static_cast<size_t>(blocks) * blksort::BlkSort::BlockSize, size - blocks * blksort::BlkSort::BlockSize);
}
void ref_blk_decode(uint32_t size, uint8_t* dst, const uint8_t* src, int threads)
{
    if(threads <= 1) {
        blksort::BlkSort bs;
        bs.decode(size, dst, const_cast<uint8_t*>(src));
        return;
    }
    const uint32_t blocks = size / blksort::BlkSort::EncodedSize;
    blk_run(false, blocks, dst, src, threads);
    memcpy(dst + static_cast<size_t>(blocks) * blksort::BlkSort::BlockSize,
           src + static_cast<size_t>(blocks) * blksort::BlkSort::EncodedSize, size - blocks * blksort::BlkSort::EncodedSize);
}
#endif
}
