// harness.cpp -- drives the drop-in header the way the reference's own harness drives
// cpprcoder.h: per file, encode -> decode -> byte-compare, one markdown row
// |file|ratio|enc MiB/s|dec MiB/s| (the format of test/main.cpp:104-107).  The call
// sequences follow run_rangecoder / run_adaptive / test_rangecoder (test/main.cpp:254-363,
// :1170-1197); unlike the reference's harness, a mismatch or a failed call is an error exit.
//
//   harness <file>...            static + adaptive round trip of every file
//   harness --ans <file>...      rANS byte + word round trip (run_ans / run_ans_simd, test/main.cpp:367-545)
//   harness --blk <file>...      block sort alone, and block sort in front of the static coder (run_blksort,
//                                run_zlib_blk with this repository's coder where zlib stood, test/main.cpp:791-1002)
//   harness --selftest           the 127-nibble static test and the chunked adaptive API
#include "../../cpprcoder_b200/include/cpprcoder_b200.h"
#include "../../cpprcoder_b200/include/cppans_b200.h"
#include "../../cpprcoder_b200/include/blksort_b200.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <string>
#include <vector>

namespace
{
double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void print(const char* filepath, double ratio, double deflateSpeed, double inflateSpeed)
{
    printf("|%s|%f|%f|%f|\n", filepath, ratio, deflateSpeed, inflateSpeed);
}

bool same(const cpprcoder::MemoryStream& got, const std::vector<cpprcoder::u8>& want)
{
    if(got.size() != static_cast<cpprcoder::s32>(want.size())) {
        printf("size %d != %zu\n", got.size(), want.size());
        return false;
    }
    for(size_t i = 0; i < want.size(); ++i) {
        if(got[static_cast<cpprcoder::s32>(i)] != want[i]) {
            printf("[%zu] %d != %d\n", i, got[static_cast<cpprcoder::s32>(i)], want[i]);
            return false;
        }
    }
    return true;
}

bool run_rangecoder(const char* filepath, const std::vector<cpprcoder::u8>& src)
{
    const cpprcoder::u32 size = static_cast<cpprcoder::u32>(src.size());
    cpprcoder::MemoryStream encstream(size);
    cpprcoder::MemoryStream decstream(size);
    cpprcoder::RangeEncoder<> encoder;
    double t = now();
    if(!encoder.encode(encstream, size, src.data())) {
        printf("%s: static encode failed\n", filepath);
        return false;
    }
    const double deflateTime = now() - t;
    t = now();
    if(!encoder.decode(decstream, static_cast<cpprcoder::u32>(encstream.size()), encstream.get())) {
        printf("%s: static decode failed\n", filepath);
        return false;
    }
    const double inflateTime = now() - t;
    print(filepath, (double)size / encstream.size(), size / deflateTime / (1024.0 * 1024.0),
          size / inflateTime / (1024.0 * 1024.0));
    return same(decstream, src);
}

bool run_adaptive(const char* filepath, const std::vector<cpprcoder::u8>& src)
{
    const cpprcoder::u32 size = static_cast<cpprcoder::u32>(src.size());
    cpprcoder::MemoryStream encstream(size);
    cpprcoder::MemoryStream decstream(size);
    double t = now();
    cpprcoder::AdaptiveRangeEncoder<> encoder;
    if(!encoder.initialize(encstream, size)) {
        return false;
    }
    cpprcoder::Result result0 = encoder.encode(static_cast<cpprcoder::s32>(size), src.data());
    if(cpprcoder::Status_Success != result0.status_) {
        printf("%s: adaptive encode status %d\n", filepath, result0.status_);
        return false;
    }
    const double deflateTime = now() - t;
    t = now();
    cpprcoder::AdaptiveRangeDecoder<> decoder;
    if(!decoder.initialize(decstream)) {
        return false;
    }
    cpprcoder::Result result1 = decoder.decode(encstream.size(), &encstream[0]);
    if(cpprcoder::Status_Success != result1.status_) {
        printf("%s: adaptive decode status %d\n", filepath, result1.status_);
        return false;
    }
    const double inflateTime = now() - t;
    print(filepath, (double)size / encstream.size(), size / deflateTime / (1024.0 * 1024.0),
          size / inflateTime / (1024.0 * 1024.0));
    return same(decstream, src);
}

bool test_rangecoder()
{
    std::mt19937 mt(12345);
    static const int Size = 127;
    std::vector<cpprcoder::u8> src(Size);
    for(int i = 0; i < Size; ++i) {
        src[i] = static_cast<cpprcoder::u8>(mt() & 0x0FU);
    }
    cpprcoder::RangeEncoder<> encoder;
    cpprcoder::MemoryStream encstream(Size);
    if(!encoder.encode(encstream, Size, src.data())) {
        return false;
    }
    cpprcoder::MemoryStream decstream(Size);
    if(!encoder.decode(decstream, static_cast<cpprcoder::u32>(encstream.size()), encstream.get())) {
        return false;
    }
    return same(decstream, src);
}

// the streaming contract: pieces in, Pending with the remaining count, Success on the last
bool test_adaptive_chunked()
{
    std::mt19937 mt(777);
    static const int Size = 300000;
    std::vector<cpprcoder::u8> src(Size);
    for(int i = 0; i < Size; ++i) {
        src[i] = static_cast<cpprcoder::u8>((mt() & 0xFFU) < 200 ? 'a' + (mt() % 7) : mt() & 0xFFU);
    }
    cpprcoder::AdaptiveRangeEncoder<> encoder;
    cpprcoder::MemoryStream encstream(16);
    if(!encoder.initialize(encstream, Size)) {
        return false;
    }
    int fed = 0;
    while(fed < Size) {
        const int piece = (Size - fed) < 70001 ? (Size - fed) : 70001;
        cpprcoder::Result r = encoder.encode(piece, src.data() + fed);
        fed += piece;
        if(fed < Size) {
            if(r.status_ != cpprcoder::Status_Pending || r.requestSize_ != static_cast<cpprcoder::u32>(Size - fed)) {
                printf("chunked encode: expected Pending/%d, got %d/%u\n", Size - fed, r.status_, r.requestSize_);
                return false;
            }
        } else if(r.status_ != cpprcoder::Status_Success) {
            return false;
        }
    }
    // one whole-buffer encode must give the same bytes
    cpprcoder::AdaptiveRangeEncoder<> once;
    cpprcoder::MemoryStream encstream2(16);
    once.initialize(encstream2, Size);
    if(once.encode(Size, src.data()).status_ != cpprcoder::Status_Success || encstream2.size() != encstream.size() ||
       0 != memcmp(encstream2.get(), encstream.get(), static_cast<size_t>(encstream.size()))) {
        printf("chunked and whole-buffer encodes differ\n");
        return false;
    }
    cpprcoder::AdaptiveRangeDecoder<> decoder;
    cpprcoder::MemoryStream decstream(16);
    decoder.initialize(decstream);
    int given = 0;
    const int total = encstream.size();
    cpprcoder::Result r = {cpprcoder::Status_Pending, 0};
    while(given < total) {
        const int piece = (total - given) < 50000 ? (total - given) : 50000;
        r = decoder.decode(piece, encstream.get() + given);
        given += piece;
        if(given < total && r.status_ != cpprcoder::Status_Pending) {
            printf("chunked decode: expected Pending, got %d\n", r.status_);
            return false;
        }
    }
    return r.status_ == cpprcoder::Status_Success && same(decstream, src);
}
// run_ans / run_ans_simd (test/main.cpp:397-455, :487-545): the encoder fills `encoded` from its
// end, the decoder is handed encoded + dst_size - result0.
bool run_ans(const char* filepath, const std::vector<cpprcoder::u8>& src, bool simd)
{
    using namespace cppans;
    const u32 size = static_cast<u32>(src.size());
    if(0 == size) {
        return true;  // the reference asserts 0 < src_size
    }
    const u64 dst_size = rANS::calc_encoded_size(size);
    std::vector<u8> encoded(dst_size), decoded(size);
    double t = now();
    const u32 result0 = simd ? rANS::encode_simd(static_cast<u32>(dst_size), encoded.data(), size, src.data())
                             : rANS::encode(static_cast<u32>(dst_size), encoded.data(), size, src.data());
    if(0 == result0) {
        printf("%s: rANS encode failed\n", filepath);
        return false;
    }
    const double deflateTime = now() - t;
    t = now();
    const u8* encoded_start = encoded.data() + dst_size - result0;
    const u32 result1 = simd ? rANS::decode_simd(size, decoded.data(), result0, encoded_start)
                             : rANS::decode(size, decoded.data(), result0, encoded_start);
    if(0 == result1) {
        printf("%s: rANS decode failed\n", filepath);
        return false;
    }
    const double inflateTime = now() - t;
    print(filepath, (double)size / result0, size / deflateTime / (1024.0 * 1024.0), size / inflateTime / (1024.0 * 1024.0));
    // too small a destination is refused with 0, as in the reference (cppans.h:523-525, :541-543)
    if(0 != (simd ? rANS::encode_simd(result0 - 1, encoded.data(), size, src.data())
                  : rANS::encode(result0 - 1, encoded.data(), size, src.data()))) {
        printf("%s: rANS encode into too small a buffer did not fail\n", filepath);
        return false;
    }
    if(size > 1 && 0 != (simd ? rANS::decode_simd(size - 1, decoded.data(), result0, encoded_start)
                              : rANS::decode(size - 1, decoded.data(), result0, encoded_start))) {
        printf("%s: rANS decode into too small a buffer did not fail\n", filepath);
        return false;
    }
    for(u32 i = 0; i < size; ++i) {
        if(decoded[i] != src[i]) {
            printf("[%u] %d != %d\n", i, decoded[i], src[i]);
            return false;
        }
    }
    return true;
}
}  // namespace

// run_blksort (test/main.cpp:791-833): the transform and its inverse, ratio = size / encodeBound
bool run_blksort(const char* filepath, const std::vector<cpprcoder::u8>& src)
{
    const uint32_t size = static_cast<uint32_t>(src.size());
    blksort::BlkSort blk;
    const uint32_t encoded_size = blksort::BlkSort::encodeBound(size);
    std::vector<uint8_t> encoded(encoded_size + 1), decoded(size + 1);
    double t = now();
    blk.encode(size, encoded.data(), src.data());
    const double deflateTime = now() - t;
    if(!blk.ok()) {
        printf("%s: block sort failed (%s)\n", filepath, b2rc_strerror(blk.status()));
        return false;
    }
    t = now();
    blk.decode(encoded_size, decoded.data(), encoded.data());
    const double inflateTime = now() - t;
    if(!blk.ok() || 0 != memcmp(decoded.data(), src.data(), size)) {
        printf("%s: block sort round trip failed\n", filepath);
        return false;
    }
    print(filepath, encoded_size ? static_cast<double>(size) / encoded_size : 1.0, size / deflateTime / (1024.0 * 1024.0),
          size / inflateTime / (1024.0 * 1024.0));
    return true;
}

// run_zlib_blk (test/main.cpp:944-1002) with the static range coder in the place of zlib:
// block sort -> encode, decode -> inverse block sort.
bool run_rc_blk(const char* filepath, const std::vector<cpprcoder::u8>& src)
{
    const uint32_t size = static_cast<uint32_t>(src.size());
    blksort::BlkSort blk;
    const uint32_t encoded_size = blksort::BlkSort::encodeBound(size);
    std::vector<uint8_t> encoded(encoded_size + 1), decoded(size + 1);
    cpprcoder::MemoryStream encstream(size);
    cpprcoder::MemoryStream decstream(size);
    cpprcoder::RangeEncoder<> coder;
    double t = now();
    blk.encode(size, encoded.data(), src.data());
    if(!blk.ok() || !coder.encode(encstream, encoded_size, encoded.data())) {
        printf("%s: block sort + encode failed\n", filepath);
        return false;
    }
    const double deflateTime = now() - t;
    t = now();
    if(!coder.decode(decstream, static_cast<cpprcoder::u32>(encstream.size()), encstream.get()) ||
       decstream.size() != static_cast<cpprcoder::s32>(encoded_size)) {
        printf("%s: decode behind the block sort failed\n", filepath);
        return false;
    }
    blk.decode(encoded_size, decoded.data(), &decstream[0]);
    const double inflateTime = now() - t;
    if(!blk.ok() || 0 != memcmp(decoded.data(), src.data(), size)) {
        printf("%s: block sort + coder round trip failed\n", filepath);
        return false;
    }
    print(filepath, static_cast<double>(size) / encstream.size(), size / deflateTime / (1024.0 * 1024.0),
          size / inflateTime / (1024.0 * 1024.0));
    return true;
}

int main(int argc, char** argv)
{
    if(argc >= 2 && std::string(argv[1]) == "--selftest") {
        const bool a = test_rangecoder();
        const bool b = test_adaptive_chunked();
        printf("test_rangecoder %s\ntest_adaptive_chunked %s\n", a ? "ok" : "FAILED", b ? "ok" : "FAILED");
        return (a && b) ? 0 : 1;
    }
    int bad = 0;
    const bool ans = argc >= 2 && std::string(argv[1]) == "--ans";
    const bool blk = argc >= 2 && std::string(argv[1]) == "--blk";
    for(int i = (ans || blk) ? 2 : 1; i < argc; ++i) {
        std::ifstream file(argv[i], std::ios::binary);
        if(!file.is_open()) {
            printf("cannot open %s\n", argv[i]);
            ++bad;
            continue;
        }
        std::vector<cpprcoder::u8> src((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
        if(blk) {
            bad += run_blksort(argv[i], src) ? 0 : 1;
            bad += run_rc_blk(argv[i], src) ? 0 : 1;
            continue;
        }
        if(ans) {
            bad += run_ans(argv[i], src, false) ? 0 : 1;
            bad += run_ans(argv[i], src, true) ? 0 : 1;
            continue;
        }
        bad += run_rangecoder(argv[i], src) ? 0 : 1;
        bad += run_adaptive(argv[i], src) ? 0 : 1;
    }
    return bad ? 1 : 0;
}
