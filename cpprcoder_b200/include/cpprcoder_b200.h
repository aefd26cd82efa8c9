// cpprcoder_b200.h -- drop-in C++ host header for the range-coder path of taqu/cpprcoder.
//
// Same namespace, class names, entry points, argument meaning and error behaviour as
// the reference's cpprcoder.h for this path, so that its callers -- run_rangecoder /
// run_adaptive (test/main.cpp:254-363), test_rangecoder / test_adaptive (:1170-1237) --
// compile unchanged against this header:
//
//   cpprcoder::RangeEncoder<T>::encode(T&, u32, const u8*) / ::decode     cpprcoder.h:336-337
//   cpprcoder::AdaptiveRangeEncoder<T>::initialize / ::encode(s32,..) / ::encode(u8)   :636-638
//   cpprcoder::AdaptiveRangeDecoder<T>::initialize / ::decode             :819-820
//   cpprcoder::MemoryStream, IStream<T>, Status, Result, u8..u64          :83-247
//
// What differs, by design: the bytes written to the stream are a B2RC container
// (include/b2rc.h) of independent 64 KiB blocks, each block's payload being exactly
// the reference's output for that block; the work is done by CUDA kernels through the
// C ABI of libb2rc.so.  There is no CPU coding path: without a CUDA device encode /
// decode return false / Status_Error.
//
// The three coder classes are new code written against the reference's interface.  The
// small host-side types below them in the file -- the typedef block, Status / Result,
// IStream and MemoryStream -- FOLLOW the reference's bodies (cpprcoder.h:83-247, :964-1077)
// statement for statement: MemoryStream's growth policy is observable through capacity(),
// so reserve / resize / writeByte / expand have to behave identically.  Stream concept kept:
// encoders need `s32 write(s32, const u8*)`, decoders are given the whole output through
// write() as well (a superset of the reference, whose writeByte() fails where capacity <
// output, cpprcoder.h:1047-1054).
#ifndef INC_CPPRCODER_B200_H_
#define INC_CPPRCODER_B200_H_

#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#if defined(__linux__)
#include <sys/mman.h>
#endif

#include "../../include/b2rc.h"

#ifndef CPPRCODER_ASSERT
#    define CPPRCODER_ASSERT(exp) assert(exp)
#endif
#ifndef CPPRCODER_NULL
#    define CPPRCODER_NULL nullptr
#endif

namespace cpprcoder
{
#ifndef CPPRCODER_TYPES
#    define CPPRCODER_TYPES
typedef int8_t s8;
typedef int16_t s16;
typedef int32_t s32;
typedef int64_t s64;
typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef float f32;
typedef double f64;
typedef char Char;
using ::size_t;
using ::uintptr_t;
#endif

enum Status  // cpprcoder.h:112-117
{
    Status_Success = 0,
    Status_Pending = 1,
    Status_Error = -1,
};

struct Result  // cpprcoder.h:119-123
{
    Status status_;
    u32 requestSize_;
};

// CRTP byte stream, cpprcoder.h:130-166
template<class T>
class IStream
{
public:
    s32 read(s32 size, u8* bytes) { return static_cast<T*>(this)->read(size, bytes); }
    bool readByte(u8& byte) { return static_cast<T*>(this)->readByte(byte); }
    s32 write(s32 size, const u8* bytes) { return static_cast<T*>(this)->write(size, bytes); }
    bool writeByte(u8 byte) { return static_cast<T*>(this)->writeByte(byte); }

protected:
    IStream() {}
    ~IStream() {}

private:
    IStream(const IStream&) = delete;
    IStream& operator=(const IStream&) = delete;
};

// Growable memory buffer with the reference's conventions (cpprcoder.h:185-247, :964-1077):
// capacity rounded up to 16, write() grows (x2 below 16 KiB, +16 KiB above), writeByte()
// never grows, reserve() discards, size() is both cursors.
namespace detail
{
// malloc for a stream's buffer.  A buffer of many megabytes is fresh memory from the kernel, and the first write to
// every 4 KiB page of it is a page fault (a quarter of a million of them for a 1 GiB stream): ask for huge pages where
// the system hands them out on request (transparent huge pages in "madvise" mode; nothing happens elsewhere).
inline u8* big_malloc(size_t bytes)
{
    u8* p = static_cast<u8*>(::malloc(bytes));
#if defined(__linux__) && defined(MADV_HUGEPAGE)
    const size_t huge = size_t(2) << 20;
    if(p && bytes >= 4 * huge) {
        const uintptr_t lo = (reinterpret_cast<uintptr_t>(p) + huge - 1) & ~(uintptr_t)(huge - 1);
        const uintptr_t hi = (reinterpret_cast<uintptr_t>(p) + bytes) & ~(uintptr_t)(huge - 1);
        if(hi > lo) {
            ::madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
        }
    }
#endif
    return p;
}
}  // namespace detail

class MemoryStream : public IStream<MemoryStream>
{
public:
    MemoryStream() : capacity_(0), size_(0), buffer_(CPPRCODER_NULL) {}
    explicit MemoryStream(s32 capacity) : capacity_(capacity), size_(0)
    {
        capacity_ = (capacity_ <= 0) ? 16 : static_cast<s32>((static_cast<u32>(capacity_) + 15U) & ~15U);
        buffer_ = detail::big_malloc(static_cast<size_t>(capacity_));
    }
    ~MemoryStream() { ::free(buffer_); }

    s32 capacity() const { return capacity_; }
    s32 size() const { return size_; }
    const u8* get() const { return buffer_; }
    const u8& operator[](s32 index) const
    {
        CPPRCODER_ASSERT(0 <= index && index < size_);
        return buffer_[index];
    }
    u8& operator[](s32 index)
    {
        CPPRCODER_ASSERT(0 <= index && index < size_);
        return buffer_[index];
    }

    void reserve(s32 capacity)
    {
        capacity = static_cast<s32>((static_cast<u32>(capacity) + 15U) & ~15U);
        if(capacity < capacity_) {
            return;
        }
        ::free(buffer_);
        capacity_ = capacity;
        buffer_ = detail::big_malloc(static_cast<size_t>(capacity_));
    }
    void resize(s32 size)
    {
        CPPRCODER_ASSERT(0 <= size);
        if(capacity_ < size) {
            reserve(size);
        }
        size_ = size;
    }
    s32 read(s32 size, u8* bytes)
    {
        const s32 end = size_ + size;
        if(capacity_ < end) {
            return -1;
        }
        ::memcpy(bytes, buffer_ + size_, static_cast<size_t>(size));
        size_ = end;
        return size;
    }
    bool readByte(u8& byte)
    {
        if(capacity_ < size_ + 1) {
            return false;
        }
        byte = buffer_[size_++];
        return true;
    }
    s32 write(s32 size, const u8* bytes)
    {
        CPPRCODER_ASSERT(0 <= size);
        const s32 end = size_ + size;
        if(capacity_ < end && !expand(end)) {
            return -1;
        }
        ::memcpy(buffer_ + size_, bytes, static_cast<size_t>(size));
        size_ = end;
        return size;
    }
    bool writeByte(u8 byte)
    {
        if(capacity_ <= size_) {
            return false;
        }
        buffer_[size_++] = byte;
        return true;
    }

private:
    MemoryStream(const MemoryStream&) = delete;
    MemoryStream& operator=(const MemoryStream&) = delete;
    static const s32 EXPAND_LIMIT_SIZE = 4096 * 4;
    friend struct detail_access;  // the coder classes below write their output straight into buffer_

    bool expand(s32 size)
    {
        s32 prev = capacity_, capacity = 0;
        do {
            if(prev <= 0) {
                prev = capacity = 1024;
            } else if(prev < EXPAND_LIMIT_SIZE) {
                prev = capacity = prev << 1;
            } else {
                prev = capacity = prev + EXPAND_LIMIT_SIZE;
            }
        } while(capacity < size);
        capacity = static_cast<s32>((static_cast<u32>(capacity) + 15U) & ~15U);
        u8* grown = detail::big_malloc(static_cast<size_t>(capacity));
        if(CPPRCODER_NULL == grown) {
            return false;
        }
        if(buffer_) {
            ::memcpy(grown, buffer_, static_cast<size_t>(size_));
        }
        ::free(buffer_);
        capacity_ = capacity;
        buffer_ = grown;
        return true;
    }

    s32 capacity_;
    s32 size_;
    u8* buffer_;
};

// A stream over page-locked host memory (not in the reference): same concept as MemoryStream, fixed
// capacity, for callers that want the host <-> device copies at full PCIe speed.  The coders write
// straight into it, as they do into an empty MemoryStream.
class PinnedStream : public IStream<PinnedStream>
{
public:
    explicit PinnedStream(u64 capacity) : capacity_(0), size_(0), buffer_(CPPRCODER_NULL)
    {
        void* p = CPPRCODER_NULL;
        if(B2RC_OK == b2rc_host_alloc(capacity, &p)) {
            buffer_ = static_cast<u8*>(p);
            capacity_ = capacity;
        }
    }
    ~PinnedStream() { b2rc_host_free(buffer_); }
    u64 capacity() const { return capacity_; }
    u64 size() const { return size_; }
    const u8* get() const { return buffer_; }
    u8* data() { return buffer_; }
    void resize(u64 size)
    {
        CPPRCODER_ASSERT(size <= capacity_);
        size_ = size;
    }
    const u8& operator[](u64 index) const { return buffer_[index]; }
    s32 write(s32 size, const u8* bytes)
    {
        if(size < 0 || capacity_ < size_ + static_cast<u64>(size)) {
            return -1;
        }
        ::memcpy(buffer_ + size_, bytes, static_cast<size_t>(size));
        size_ += static_cast<u64>(size);
        return size;
    }
    bool writeByte(u8 byte)
    {
        if(capacity_ <= size_) {
            return false;
        }
        buffer_[size_++] = byte;
        return true;
    }

private:
    PinnedStream(const PinnedStream&) = delete;
    PinnedStream& operator=(const PinnedStream&) = delete;
    u64 capacity_;
    u64 size_;
    u8* buffer_;
};

// Lets the coders use an EMPTY MemoryStream's buffer as the destination of the C ABI call itself:
// reserve() (which may discard, cpprcoder.h:985-994 -- nothing is lost in an empty stream), code, resize().
struct detail_access {
    static u8* writable(MemoryStream& s, u64 need)
    {
        if(0 != s.size_ || 0x7FFFFFF0ULL < need) {
            return CPPRCODER_NULL;
        }
        s.reserve(static_cast<s32>(need));
        return (CPPRCODER_NULL != s.buffer_ && static_cast<u64>(s.capacity_) >= need) ? s.buffer_ : CPPRCODER_NULL;
    }
    static void written(MemoryStream& s, u64 made) { s.size_ = static_cast<s32>(made); }
};

namespace detail
{
// One b2rc context per host thread, created on first use (device 0 unless
// CPPRCODER_B200_DEVICE or CPPRCODER_B200_DEVICES is set).  Null when there is no CUDA device.
inline b2rc_ctx* context()
{
    struct Holder {
        b2rc_ctx* ctx;
        Holder() : ctx(CPPRCODER_NULL)
        {
            // CPPRCODER_B200_DEVICES=0,1,2,3: one context over several devices (b2rc_ctx_create_multi);
            // the host-pointer calls these classes make then shard the blocks over them
            if(const char* list = ::getenv("CPPRCODER_B200_DEVICES")) {
                int devs[16];
                int n = 0;
                for(const char* p = list; *p && n < 16;) {
                    devs[n++] = ::atoi(p);
                    while(*p && *p != ',') {
                        ++p;
                    }
                    if(*p == ',') {
                        ++p;
                    }
                }
                if(0 < n) {
                    b2rc_ctx_create_multi(devs, n, &ctx);
                    return;
                }
            }
            const char* dev = ::getenv("CPPRCODER_B200_DEVICE");
            b2rc_ctx_create(dev ? ::atoi(dev) : 0, &ctx);
        }
        ~Holder() { b2rc_ctx_destroy(ctx); }
    };
    static thread_local Holder holder;
    return holder.ctx;
}

// the stream API is s32-sized (cpprcoder.h:143); hand large outputs over in slices
template<class T>
bool write_all(T& stream, const u8* bytes, u64 size)
{
    while(0 < size) {
        const s32 piece = static_cast<s32>(size < 0x40000000ULL ? size : 0x40000000ULL);
        if(stream.write(piece, bytes) <= 0) {
            return false;
        }
        bytes += piece;
        size -= static_cast<u64>(piece);
    }
    return true;
}

// Any stream: the container is made in pinned host memory that the context owns (never initialised,
// reused from call to call) and handed to write() from there.
template<class T>
bool encode_to(T& stream, int mode, u32 blockSize, u64 size, const u8* bytes)
{
    b2rc_ctx* ctx = context();
    if(CPPRCODER_NULL == ctx) {
        return false;
    }
    static const u8 nothing = 0;
    const u8* out = CPPRCODER_NULL;
    u64 made = 0;
    if(B2RC_OK != b2rc_encode_staged(ctx, mode, blockSize, bytes ? bytes : &nothing, size, &out, &made)) {
        return false;
    }
    return write_all(stream, out, made);
}

// MemoryStream, empty (what run_rangecoder / run_adaptive pass, test/main.cpp:270-271): the container
// is made in the context's pinned staging (a device-to-host copy into fresh malloc memory runs at a
// fraction of the PCIe rate) and moved into the stream's own buffer, reserved at its exact size, by
// a few threads -- no intermediate vector, no zero fill, no second pass through write().
inline bool encode_to(MemoryStream& stream, int mode, u32 blockSize, u64 size, const u8* bytes)
{
    b2rc_ctx* ctx = context();
    if(CPPRCODER_NULL == ctx) {
        return false;
    }
    static const u8 nothing = 0;
    const u8* out = CPPRCODER_NULL;
    u64 made = 0;
    if(B2RC_OK != b2rc_encode_staged(ctx, mode, blockSize, bytes ? bytes : &nothing, size, &out, &made)) {
        return false;
    }
    u8* dst = detail_access::writable(stream, made);
    if(CPPRCODER_NULL == dst) {
        return write_all(stream, out, made);
    }
    b2rc_host_copy(dst, out, made);
    detail_access::written(stream, made);
    return true;
}

inline bool encode_to(PinnedStream& stream, int mode, u32 blockSize, u64 size, const u8* bytes)
{
    b2rc_ctx* ctx = context();
    if(CPPRCODER_NULL == ctx || CPPRCODER_NULL == stream.data()) {
        return false;
    }
    static const u8 nothing = 0;
    u64 made = 0;
    if(B2RC_OK != b2rc_encode(ctx, mode, blockSize, bytes ? bytes : &nothing, size, stream.data() + stream.size(),
                              stream.capacity() - stream.size(), &made)) {
        return false;
    }
    stream.resize(stream.size() + made);
    return true;
}

template<class T>
int decode_to(T& stream, u64 size, const u8* bytes)
{
    b2rc_ctx* ctx = context();
    if(CPPRCODER_NULL == ctx) {
        return B2RC_E_CUDA;
    }
    const u8* out = CPPRCODER_NULL;
    u64 made = 0;
    const int rc = b2rc_decode_staged(ctx, bytes, size, &out, &made);  // validates the index before it allocates
    if(B2RC_OK != rc) {
        return rc;
    }
    return write_all(stream, out, made) ? B2RC_OK : B2RC_E_DST_SMALL;
}

inline int decode_to(PinnedStream& stream, u64 size, const u8* bytes)
{
    b2rc_ctx* ctx = context();
    if(CPPRCODER_NULL == ctx || CPPRCODER_NULL == stream.data()) {
        return B2RC_E_CUDA;
    }
    u64 made = 0;
    const int rc = b2rc_decode(ctx, bytes, size, stream.data() + stream.size(), stream.capacity() - stream.size(), &made);
    if(B2RC_OK == rc) {
        stream.resize(stream.size() + made);
    }
    return rc;
}

inline int decode_to(MemoryStream& stream, u64 size, const u8* bytes)
{
    b2rc_ctx* ctx = context();
    if(CPPRCODER_NULL == ctx) {
        return B2RC_E_CUDA;
    }
    const u8* out = CPPRCODER_NULL;
    u64 made = 0;
    const int rc = b2rc_decode_staged(ctx, bytes, size, &out, &made);  // validates the index before it allocates
    if(B2RC_OK != rc) {
        return rc;
    }
    u8* dst = detail_access::writable(stream, made ? made : 1);
    if(CPPRCODER_NULL == dst) {
        return write_all(stream, out, made) ? B2RC_OK : B2RC_E_DST_SMALL;
    }
    b2rc_host_copy(dst, out, made);
    detail_access::written(stream, made);
    return B2RC_OK;
}
} // namespace detail

//----------------------------------------------
//--- RangeEncoder            (cpprcoder.h:321-619)
//----------------------------------------------
template<class T = MemoryStream>
class RangeEncoder
{
public:
    static const u32 MAX_SIZE = 0x7FFFFFFFU;  // cpprcoder.h:329
    static const u32 FREQUENCY_SIZE = 256;

    RangeEncoder() : blockSize_(B2RC_DEFAULT_BLOCK) {}
    ~RangeEncoder() {}

    /// Block size of the container written by encode (multiple of 64, 64 .. 2^23).
    void setBlockSize(u32 blockSize) { blockSize_ = blockSize; }
    u32 blockSize() const { return blockSize_; }

    /// cpprcoder.h:375: false when the stream cannot take the output (or no CUDA device).
    bool encode(T& stream, u32 size, const u8* bytes)
    {
        CPPRCODER_ASSERT(size <= MAX_SIZE);
        CPPRCODER_ASSERT(CPPRCODER_NULL != bytes || 0 == size);
        return detail::encode_to(stream, B2RC_MODE_STATIC, blockSize_, size, bytes);
    }
    /// 64-bit entry for streams above 2 GiB (the C ABI is 64-bit throughout).  A separate name keeps
    /// `encode(stream, int, ptr)` calls of existing code unambiguous.
    bool encode64(T& stream, u64 size, const u8* bytes)
    {
        return detail::encode_to(stream, B2RC_MODE_STATIC, blockSize_, size, bytes);
    }
    /// cpprcoder.h:460: false on short or corrupt input.
    bool decode(T& stream, u32 size, const u8* bytes)
    {
        CPPRCODER_ASSERT(CPPRCODER_NULL != bytes);
        if(size < 1) {
            return false;  // cpprcoder.h:468-470
        }
        return B2RC_OK == detail::decode_to(stream, size, bytes);
    }
    bool decode64(T& stream, u64 size, const u8* bytes) { return B2RC_OK == detail::decode_to(stream, size, bytes); }

private:
    RangeEncoder(const RangeEncoder&) = delete;
    RangeEncoder& operator=(const RangeEncoder&) = delete;
    u32 blockSize_;
};

//----------------------------------------------
//--- AdaptiveRangeEncoder    (cpprcoder.h:626-802)
//----------------------------------------------
// Streaming contract kept: initialize(stream, total) then encode(size, bytes) any number
// of times; each call returns {Status_Pending, bytes still expected} until the last byte
// arrives, which returns {Status_Success, 0} (cpprcoder.h:714-719).  The pieces are
// gathered on the host and coded on the GPU when the stream is complete.
template<class T = MemoryStream>
class AdaptiveRangeEncoder
{
public:
    AdaptiveRangeEncoder() : stream_(CPPRCODER_NULL), umcompressedSize_(0), inSize_(0), blockSize_(B2RC_DEFAULT_BLOCK) {}
    ~AdaptiveRangeEncoder() {}

    void setBlockSize(u32 blockSize) { blockSize_ = blockSize; }

    bool initialize(T& stream, u32 umcompressedSize)  // cpprcoder.h:678
    {
        stream_ = &stream;
        umcompressedSize_ = umcompressedSize;
        inSize_ = 0;
        pending_.clear();
        if(0 == umcompressedSize_) {  // the reference writes its 4-byte size here; we write the empty container at once
            return detail::encode_to(*stream_, B2RC_MODE_ADAPTIVE, blockSize_, 0, CPPRCODER_NULL);
        }
        return true;
    }

    Result encode(s32 size, const u8* bytes)  // cpprcoder.h:697
    {
        CPPRCODER_ASSERT(CPPRCODER_NULL != stream_);
        CPPRCODER_ASSERT((static_cast<u64>(inSize_) + static_cast<u64>(size)) <= umcompressedSize_);
        if(CPPRCODER_NULL == stream_ || size < 0) {
            return {Status_Error, 0};
        }
        if(0 == inSize_ && static_cast<u32>(size) == umcompressedSize_) {  // the common whole-buffer call: no copy
            inSize_ = umcompressedSize_;
            return finish(bytes);
        }
        pending_.insert(pending_.end(), bytes, bytes + size);
        inSize_ += static_cast<u32>(size);
        if(umcompressedSize_ <= inSize_) {
            return finish(pending_.data());
        }
        return {Status_Pending, umcompressedSize_ - inSize_};
    }

    Result encode(u8 byte) { return encode(1, &byte); }  // cpprcoder.h:722

private:
    AdaptiveRangeEncoder(const AdaptiveRangeEncoder&) = delete;
    AdaptiveRangeEncoder& operator=(const AdaptiveRangeEncoder&) = delete;

    Result finish(const u8* bytes)
    {
        const bool ok = detail::encode_to(*stream_, B2RC_MODE_ADAPTIVE, blockSize_, umcompressedSize_, bytes);
        pending_.clear();
        if(ok) {
            return {Status_Success, 0};
        }
        return {Status_Pending, 0};  // the reference reports a full stream as Pending (cpprcoder.h:708-711)
    }

    T* stream_;
    u32 umcompressedSize_;
    u32 inSize_;
    u32 blockSize_;
    std::vector<u8> pending_;
};

//----------------------------------------------
//--- AdaptiveRangeDecoder    (cpprcoder.h:809-940)
//----------------------------------------------
// decode(size, bytes) may be fed the container in pieces: it answers
// {Status_Pending, bytes still missing} until the whole container has arrived
// (the reference resumes the same way when its input starves, cpprcoder.h:901-903).
template<class T = MemoryStream>
class AdaptiveRangeDecoder
{
public:
    AdaptiveRangeDecoder() : stream_(CPPRCODER_NULL) {}
    ~AdaptiveRangeDecoder() {}

    bool initialize(T& stream)  // cpprcoder.h:859
    {
        stream_ = &stream;
        pending_.clear();
        return true;
    }

    Result decode(s32 size, const u8* bytes)  // cpprcoder.h:872
    {
        if(CPPRCODER_NULL == stream_ || size < 0) {
            return {Status_Error, 0};
        }
        const u8* all = bytes;
        u64 have = static_cast<u64>(size);
        if(!pending_.empty()) {
            pending_.insert(pending_.end(), bytes, bytes + size);
            all = pending_.data();
            have = pending_.size();
        }
        const u64 need = containerBytes(all, have);
        if(0 == need) {
            return {Status_Error, 0};
        }
        if(have < need) {
            if(pending_.empty()) {
                pending_.assign(bytes, bytes + size);
            }
            const u64 missing = need - have;
            return {Status_Pending, static_cast<u32>(missing < 0xFFFFFFFFULL ? missing : 0xFFFFFFFFULL)};
        }
        const int rc = detail::decode_to(*stream_, need, all);
        pending_.clear();
        if(B2RC_OK == rc) {
            return {Status_Success, 0};
        }
        return {B2RC_E_DST_SMALL == rc ? Status_Pending : Status_Error, 0};
    }

private:
    AdaptiveRangeDecoder(const AdaptiveRangeDecoder&) = delete;
    AdaptiveRangeDecoder& operator=(const AdaptiveRangeDecoder&) = delete;

    // Total container length once enough of it is known (restart table included); the smallest prefix
    // that tells more when it is not; 0 for garbage.
    static u64 containerBytes(const u8* bytes, u64 have)
    {
        u64 need = 0;
        return B2RC_OK == b2rc_container_bytes(bytes, have, &need) ? need : 0;
    }

    T* stream_;
    std::vector<u8> pending_;
};

} // namespace cpprcoder
#endif // INC_CPPRCODER_B200_H_
